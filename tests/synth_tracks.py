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
