"""Rollout-kernel time per launch versus block size, resident blocks per SM and steps per launch
(where the per-launch overhead at small shards comes from).
Usage: python benchmarks/sweep_launch.py [n_envs ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200

track = ppo_car_b200.builtin_track("big_track")
sizes = [int(a) for a in sys.argv[1:]] or [131072, 1048576]


def time_rollout(env, a, o, reps):
    kw = dict(obs_out=o["obs"], reward_out=o["reward"], term_out=o["terminated"], trunc_out=o["truncated"])
    for _ in range(3):
        env.rollout(a, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        env.rollout(a, **kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n in sizes:
    for chunk in (32, 128) if n > 262144 else (32, 128, 256):
        env = ppo_car_b200.VecCarEnv(n, track)
        env.reset()
        a = torch.randint(0, 9, (chunk, n), device="cuda", dtype=torch.uint8)
        o = env.rollout(a)
        for block, pad in ((128, 0), (96, 0), (64, 0), (32, 0), (64, 24 * 1024), (64, 28 * 1024), (32, 10 * 1024),
                           (32, 12 * 1024)):
            env.set_option("block", block)
            env.set_option("smem_pad", pad)
            ms = time_rollout(env, a, o, max(5, 2 * 1048576 * 32 // (n * chunk)))
            print(json.dumps({"n_envs": n, "steps_per_launch": chunk, "block": block, "smem_pad": pad,
                              "ms_per_launch": round(ms, 4), "env_steps_per_s": n * chunk / ms * 1e3,
                              "us_per_32_steps": round(ms / chunk * 32 * 1e3, 1)}), flush=True)
        del env, a, o
        torch.cuda.empty_cache()
