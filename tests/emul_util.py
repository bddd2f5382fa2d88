"""ctypes wrapper of tests/host_emul/libcarenv_emul.so: the kernel's per-environment arithmetic
(ppo_car_b200/csrc/carenv_core.cuh) compiled for the host.  TEST HARNESS only."""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from ppo_car_b200.track import load_track

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _DIR, "-s"])
        _lib = C.CDLL(os.path.join(_DIR, "libcarenv_emul.so"))
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def emul_rollout(track_path, actions, reward_scale=1.0, threads=None, unrolled=True, state_pv=None, state_i=None):
    """Replay [T,N] actions through the host build of the kernel arithmetic, from reset or (state_pv [N,4]
    float64 px py vx vy, state_i [N,4] int32 heading index, time step, next gate, gates passed) from a given state."""
    L = lib()
    tr = load_track(track_path)
    walls, gates = np.ascontiguousarray(tr.walls), np.ascontiguousarray(tr.gates)
    a = np.ascontiguousarray(actions, np.uint8)
    T, N = a.shape
    threads = max(1, min(threads or os.cpu_count() or 1, N))
    from_state = state_pv is not None
    pv = np.ascontiguousarray(state_pv, np.float64).copy() if from_state else np.zeros((N, 4))
    si = np.ascontiguousarray(state_i, np.int32).copy() if from_state else np.zeros((N, 4), np.int32)
    out = dict(obs=np.zeros((T, N, 18), np.float32), rew=np.zeros((T, N), np.float32),
               term=np.zeros((T, N), np.uint8), trunc=np.zeros((T, N), np.uint8),
               info=np.zeros((T, N, 4), np.int32), stats=np.zeros(4, np.uint64), reset_obs=np.zeros(18, np.float32))
    edges = np.linspace(0, N, threads + 1).astype(int)

    def run(i):
        return L.emul_rollout(_p(walls), len(walls), _p(gates), len(gates), C.c_double(tr.start[0]),
                              C.c_double(tr.start[1]), C.c_double(tr.angle), N, int(edges[i]), int(edges[i + 1]), T,
                              _p(a), C.c_double(reward_scale), _p(pv), _p(si), 0 if from_state else 1, int(unrolled), _p(out["reset_obs"]), _p(out["obs"]),
                              _p(out["rew"]), _p(out["term"]), _p(out["trunc"]), _p(out["info"]), _p(out["stats"]))

    with ThreadPoolExecutor(threads) as ex:
        rcs = list(ex.map(run, range(threads)))
    assert not any(rcs), rcs
    out["state_pv"], out["state_i"] = pv, si
    return out
