"""Track files: the reference's JSON schema, loaded unchanged.

Mirrors CarEnv.load_track and the geometry construction in CarEnv.reset
(/root/reference/lib/car_env.py:535-567, 651-676): normalised coordinates are scaled by
1280 x 720, wall segments are consecutive point pairs of the outer polyline followed by
those of the inner polyline, gates are consecutive PAIRS of gate points, the start angle
stays in degrees.  The schema itself is defined by the reference's track editor
(track_editor.py:50-56).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass

import numpy as np

WIDTH, HEIGHT = 1280, 720          # lib/car_env.py:488-489
TRACK_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tracks")


@dataclass
class Track:
    walls: np.ndarray      # [S, 4] float64  x1 y1 x2 y2 (pixels), outer first then inner
    gates: np.ndarray      # [G, 4] float64
    start: tuple           # (x, y) pixels
    angle: float           # degrees
    path: str
    n_outer: int = 0       # how many of the wall segments belong to the outer polyline (rendering)


def builtin_track(name: str) -> str:
    """Path of a track shipped with the package ('track' or 'big_track')."""
    return os.path.join(TRACK_DIR, name if name.endswith(".json") else name + ".json")


def load_track(path: str) -> Track:
    with open(path, "r") as fh:            # FileNotFoundError propagates (the reference prints and returns None)
        raw = json.load(fh)
    for key in ("outer_track_points", "inner_track_points", "reward_gates", "initial_position", "initial_angle"):
        if key not in raw:
            raise ValueError(f"{path}: missing key {key!r}")
    scale = np.array([WIDTH, HEIGHT], np.float64)
    outer = np.asarray(raw["outer_track_points"], np.float64).reshape(-1, 2) * scale
    inner = np.asarray(raw["inner_track_points"], np.float64).reshape(-1, 2) * scale
    gpts = np.asarray(raw["reward_gates"], np.float64).reshape(-1, 2) * scale
    if len(outer) < 2 or len(inner) < 2:
        raise ValueError(f"{path}: a border polyline needs at least two points")
    n_g = len(gpts) // 2                   # zip(points[::2], points[1::2]) drops an unpaired last point
    if n_g < 1:
        raise ValueError(f"{path}: no reward gates")
    walls = np.concatenate([np.hstack([outer[:-1], outer[1:]]), np.hstack([inner[:-1], inner[1:]])])
    gates = np.hstack([gpts[0:2 * n_g:2], gpts[1:2 * n_g:2]])
    start = (raw["initial_position"][0] * WIDTH, raw["initial_position"][1] * HEIGHT)
    return Track(np.ascontiguousarray(walls), np.ascontiguousarray(gates), start, float(raw["initial_angle"]), path,
                 len(outer) - 1)


def _point_in_polygon(pt, poly) -> bool:
    """Even-odd rule; poly is a closed [n, 2] polyline."""
    x, y = pt
    inside = False
    for (x1, y1), (x2, y2) in zip(poly[:-1], poly[1:]):
        if (y1 > y) != (y2 > y) and x < x1 + (y - y1) * (x2 - x1) / (y2 - y1):
            inside = not inside
    return inside


def validate_track(path: str) -> list[str]:
    """Checks of a track file that the reference performs nowhere (its editor can write all of these,
    track_editor.py:211-255) but that decide whether training on it can work.  Returns a list of
    human-readable problems; an empty list means the track is fine.  SURVEY §8 f-4."""
    problems = []
    with open(path, "r") as fh:
        raw = json.load(fh)
    for key in ("outer_track_points", "inner_track_points", "reward_gates", "initial_position", "initial_angle"):
        if key not in raw:
            return [f"missing key {key!r}"]
    scale = np.array([WIDTH, HEIGHT], np.float64)
    outer = np.asarray(raw["outer_track_points"], np.float64).reshape(-1, 2) * scale
    inner = np.asarray(raw["inner_track_points"], np.float64).reshape(-1, 2) * scale
    gpts = np.asarray(raw["reward_gates"], np.float64).reshape(-1, 2) * scale
    for name, poly in (("outer", outer), ("inner", inner)):
        if len(poly) < 4:
            problems.append(f"{name} border has fewer than 3 segments")
        elif not np.array_equal(poly[0], poly[-1]):
            problems.append(f"{name} border is not closed (first point != last point): cars can leave through the gap")
        if len(poly) > 1 and (np.diff(poly, axis=0) == 0).all(axis=1).any():
            problems.append(f"{name} border has a zero-length segment")
    if len(gpts) % 2:
        problems.append("odd number of gate points: the last one is ignored (gates are consecutive pairs)")
    if len(gpts) < 2:
        problems.append("no reward gates")
    n_seg = max(len(outer) - 1, 0) + max(len(inner) - 1, 0)
    if n_seg > 2048:
        problems.append(f"{n_seg} wall segments: more than the 2,048 the kernels support")
    start = np.array([raw["initial_position"][0] * WIDTH, raw["initial_position"][1] * HEIGHT])
    if len(outer) >= 4 and not _point_in_polygon(start, outer):
        problems.append("start position is outside the outer border")
    if len(inner) >= 4 and _point_in_polygon(start, inner):
        problems.append("start position is inside the inner border")
    if not problems:                                         # clearance of the four cardinal rays at the start pose
        walls = np.concatenate([np.hstack([outer[:-1], outer[1:]]), np.hstack([inner[:-1], inner[1:]])])
        for k in range(4):
            a = np.radians(float(raw["initial_angle"]) + 90.0 * k)
            d = np.array([np.cos(a), np.sin(a)])
            e = walls[:, 2:] - walls[:, :2]
            den = e[:, 0] * d[1] - e[:, 1] * d[0]
            rel = walls[:, :2] - start
            with np.errstate(divide="ignore", invalid="ignore"):
                t = (rel[:, 1] * d[0] - rel[:, 0] * d[1]) / den
                u = (rel[:, 1] * e[:, 0] - rel[:, 0] * e[:, 1]) / den
            hit = (den != 0) & (t > 0) & (t < 1) & (u > 0)
            if hit.any() and u[hit].min() < 10.0:
                problems.append("a wall is closer than 10 px to the start pose: every episode ends at once")
                break
    return problems
