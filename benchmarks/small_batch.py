#!/usr/bin/env python
"""Small batches: the warp-per-environment kernel (k_rollout_warp) versus the thread-per-environment kernel,
1,024-step rollouts on big_track (BASELINE config 2 is the 24-env row).  One JSON line per size."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200  # noqa: E402

track = ppo_car_b200.builtin_track(sys.argv[1] if len(sys.argv) > 1 else "big_track")
T = 1024
for n in (24, 64, 256, 592, 1024, 2048, 4096, 8192, 16384):
    res = {"n_envs": n, "steps": T}
    a = torch.randint(0, 9, (T, n), device="cuda", dtype=torch.uint8)
    for tag, mode in (("warp_per_env", 1), ("thread_per_env", -1)):
        env = ppo_car_b200.VecCarEnv(n, track)
        env.set_option("warp_per_env", mode)
        env.reset()
        o = env.rollout(a)
        kw = dict(obs_out=o["obs"], reward_out=o["reward"], term_out=o["terminated"], trunc_out=o["truncated"])
        for _ in range(2):
            env.rollout(a, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            env.rollout(a, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res[tag + "_us_per_step"] = round(ms / T * 1e3, 3)
        res[tag + "_env_steps_per_s"] = n * T / ms * 1e3
    print(json.dumps(res), flush=True)
