"""Synthetic tracks in the reference's JSON schema (track_editor.py:50-56) for exercising segment counts
the two shipped tracks do not have (odd counts -> generic kernel loop, even -> 2-segment unroll)."""
import json
import math


def ring_track(path, n_outer, n_inner, n_gates=12, wobble=0.08, seed_phase=0.3):
    cx, cy = 0.5, 0.5

    def poly(n, rx, ry, phase):
        pts = []
        for i in range(n):
            a = 2 * math.pi * i / n + phase
            r = 1.0 + wobble * math.sin(3 * a + phase)
            pts.append([cx + rx * r * math.cos(a), cy + ry * r * math.sin(a)])
        return pts + [pts[0]]                                  # closed polyline, first == last

    outer = poly(n_outer, 0.42, 0.40, seed_phase)
    inner = poly(n_inner, 0.22, 0.18, seed_phase * 0.5)
    gates = []
    for g in range(n_gates):
        a = 2 * math.pi * (g + 0.5) / n_gates
        gates.append([cx + 0.15 * math.cos(a), cy + 0.12 * math.sin(a)])
        gates.append([cx + 0.48 * math.cos(a), cy + 0.46 * math.sin(a)])
    data = {"outer_track_points": outer, "inner_track_points": inner, "reward_gates": gates,
            "initial_position": [cx + 0.32, cy], "initial_angle": 93.7}
    with open(path, "w") as fh:
        json.dump(data, fh)
    return path


def near_wall_track(path, gap_px=6.0, n_outer=10, n_inner=6):
    """A ring track whose start pose is `gap_px` (< 10) pixels from an outer wall along ray 0, so that the very
    first Car.update inside CarEnv.reset already sets `destroyed` (lib/car_env.py:682-686) and EVERY step
    terminates with the -3 penalty (lib/car_env.py:745-748; `destroyed` is only cleared by reset)."""
    ring_track(path, n_outer, n_inner)
    with open(path) as fh:
        data = json.load(fh)
    W, H = 1280.0, 720.0
    a, b = data["outer_track_points"][2], data["outer_track_points"][3]
    ax, ay, bx, by = a[0] * W, a[1] * H, b[0] * W, b[1] * H
    mx, my = 0.5 * (ax + bx), 0.5 * (ay + by)
    nx, ny = -(by - ay), bx - ax                               # a normal of the wall ...
    if nx * (0.5 * W - mx) + ny * (0.5 * H - my) < 0:          # ... turned towards the centre of the ring
        nx, ny = -nx, -ny
    ln = math.hypot(nx, ny)
    nx, ny = nx / ln, ny / ln
    sx, sy = mx + gap_px * nx, my + gap_px * ny
    data["initial_position"] = [sx / W, sy / H]
    data["initial_angle"] = math.degrees(math.atan2(-ny, -nx))  # ray 0 points straight at the wall
    with open(path, "w") as fh:
        json.dump(data, fh)
    return path
