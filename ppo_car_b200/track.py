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
    return Track(np.ascontiguousarray(walls), np.ascontiguousarray(gates), start, float(raw["initial_angle"]), path)
