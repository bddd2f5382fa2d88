"""Where the time of the host-buffer step goes: sub-range count, float64 reward widening, record vs array layout.
    python benchmarks/e2e_breakdown.py [--envs 1048576] [--out profiles/x.jsonl]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200  # noqa: E402
from ppo_car_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1_048_576)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    n = args.envs
    L = _lib.lib()
    track = ppo_car_b200.builtin_track("big_track")
    rng = np.random.default_rng(0)
    acts = [rng.integers(0, 9, size=n).astype(np.int64) for _ in range(4)]
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
    obs, rec, rew64 = pin((n, 18), torch.float32), pin((n, 16), torch.uint8), pin((n,), torch.float64)
    rew, te, tr = pin((n,), torch.float32), pin((n,), torch.uint8), pin((n,), torch.uint8)
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    rows = []

    def run(name, ranges, fn):
        env = ppo_car_b200.VecCarEnv(n, track)
        env.set_option("host_ranges", ranges)
        env.reset()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for i in range(3):
            fn(env, acts[i % 4], st)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(args.steps):
            fn(env, acts[i % 4], st)
        dt = (time.perf_counter() - t0) / args.steps
        row = {"variant": name, "ranges": ranges, "envs": n, "ms_per_step": dt * 1e3, "env_steps_per_s": n / dt}
        print(json.dumps(row), flush=True)
        rows.append(row)

    def records(widen):
        def f(env, a, st):
            rc = L.carenv_step_host_records(env._handle, n, p(env.pos), p(env.vel), p(env.ints), C.c_void_p(a.ctypes.data),
                                            _lib.ACT_I64, 1.0, p(obs), p(rec), None, st)
            assert rc == 0
        return f

    def arrays(env, a, st):
        rc = L.carenv_step_host(env._handle, n, p(env.pos), p(env.vel), p(env.ints), C.c_void_p(a.ctypes.data),
                                _lib.ACT_I64, 1.0, p(obs), p(rew), p(te), p(tr), _lib.FLAG_U8, None, st)
        assert rc == 0

    def api(env, a, st):
        env.step(a)

    for ranges in (2, 4, 8, 16):
        run("16-byte records with float64 rewards (C ABI)", ranges, records(False))
    for ranges in (8, 16):
        run("separate arrays, no info (C ABI)", ranges, arrays)
    run("VecCarEnv.step(numpy)", 0, api)
    run("VecCarEnv.step(numpy)", 8, api)
    if args.out:
        with open(args.out, "w") as fh:
            for r in rows:
                fh.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
