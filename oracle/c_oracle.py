"""TEST INFRASTRUCTURE — ctypes wrapper around oracle/libcarenv_oracle.so.

Checker and CPU baseline only (see the header of carenv_oracle.c).  Builds the
shared object with oracle/Makefile on first use.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .carenv_port import load_track

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    so = os.path.join(_HERE, "libcarenv_oracle.so")
    src = os.path.join(_HERE, "carenv_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_env_bytes.restype = C.c_size_t
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class COracleVecEnv:
    """N float64 environments stepped by the plain-C oracle, with same-step autoreset."""

    def __init__(self, n_envs: int, track_path: str, threads: int | None = None, scan_all_gates: bool = True):
        tr = load_track(track_path)
        self.n = n_envs
        self.walls = np.ascontiguousarray(tr["walls"], np.float64)
        self.gates = np.ascontiguousarray(tr["gates"], np.float64)
        self.sx, self.sy, self.angle = float(tr["start"][0]), float(tr["start"][1]), float(tr["angle"])
        self.scan = int(scan_all_gates)
        self.lib = _lib()
        self.state = np.zeros(n_envs * self.lib.oracle_env_bytes(), np.uint8)
        self.threads = max(1, min(threads or os.cpu_count() or 1, n_envs))

    def _shards(self):
        edges = np.linspace(0, self.n, self.threads + 1).astype(int)
        return [(int(a), int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a]

    def _run(self, fn):
        sh = self._shards()
        if len(sh) == 1:
            rcs = [fn(*sh[0])]
        else:
            with ThreadPoolExecutor(len(sh)) as ex:
                rcs = list(ex.map(lambda ab: fn(*ab), sh))
        if any(rcs):
            raise RuntimeError(f"oracle call failed: {rcs}")

    def reset(self, return_dist: bool = False):
        obs = np.zeros((self.n, 18), np.float32)
        dist = np.zeros((self.n, 12), np.float64) if return_dist else None
        self._run(lambda lo, hi: self.lib.oracle_reset(
            _p(self.state), lo, hi, _p(self.walls), len(self.walls), _p(self.gates), len(self.gates),
            C.c_double(self.sx), C.c_double(self.sy), C.c_double(self.angle), _p(obs), _p(dist)))
        return (obs, dist) if return_dist else obs

    def rollout(self, actions: np.ndarray, want=("obs", "fobs", "rew", "term", "trunc", "gates_passed",
                                                  "time_passed", "next_gate_index", "pose")) -> dict:
        """actions [T,N] (any int dtype) -> dict of [T,N,...] arrays."""
        a = np.ascontiguousarray(actions, np.uint8)
        T, n = a.shape
        assert n == self.n
        mk = {
            "obs": lambda: np.zeros((T, n, 18), np.float32), "fobs": lambda: np.zeros((T, n, 18), np.float32),
            "rew": lambda: np.zeros((T, n), np.float64), "term": lambda: np.zeros((T, n), np.uint8),
            "trunc": lambda: np.zeros((T, n), np.uint8), "gates_passed": lambda: np.zeros((T, n), np.int32),
            "time_passed": lambda: np.zeros((T, n), np.int32), "next_gate_index": lambda: np.zeros((T, n), np.int32),
            "pose": lambda: np.zeros((T, n, 5), np.float64),
        }
        out = {k: (mk[k]() if k in want else None) for k in mk}
        self._run(lambda lo, hi: self.lib.oracle_rollout(
            _p(self.state), n, lo, hi, T, _p(a), _p(self.walls), len(self.walls), _p(self.gates), len(self.gates),
            C.c_double(self.sx), C.c_double(self.sy), C.c_double(self.angle), self.scan,
            _p(out["obs"]), _p(out["fobs"]), _p(out["rew"]), _p(out["term"]), _p(out["trunc"]),
            _p(out["gates_passed"]), _p(out["time_passed"]), _p(out["next_gate_index"]), _p(out["pose"])))
        return {k: v for k, v in out.items() if v is not None}

    def step(self, actions):
        r = self.rollout(np.asarray(actions).reshape(1, -1))
        return (r["obs"][0], r["rew"][0], r["term"][0].astype(bool), r["trunc"][0].astype(bool),
                dict(gates_passed=r["gates_passed"][0], time_passed=r["time_passed"][0],
                     next_gate_index=r["next_gate_index"][0], final_obs=r["fobs"][0]))
