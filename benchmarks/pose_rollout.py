#!/usr/bin/env python
"""Rollout storage format: 72-byte observations versus 32-byte pose records per env-step, and the
observation recompute (all records / a 512-sample minibatch gather).  One JSON line per size."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200  # noqa: E402

track = ppo_car_b200.builtin_track("big_track")


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n, K in ((65536, 1024), (1048576, 128)):
    env = ppo_car_b200.VecCarEnv(n, track)
    env.reset()
    a = torch.randint(0, 9, (K, n), device="cuda", dtype=torch.uint8)
    o = env.rollout(a)
    kw = dict(reward_out=o["reward"], term_out=o["terminated"], trunc_out=o["truncated"])
    ms_obs = timed(lambda: env.rollout(a, obs_out=o["obs"], **kw), 3)
    obs_bytes = o["obs"].numel() * 4
    del o["obs"]
    p = env.rollout(a, store_poses=True, **kw)["poses"]
    ms_pose = timed(lambda: env.rollout(a, pose_out=p, **kw), 3)
    ms_none = timed(lambda: env.rollout(a, store_obs=False, **kw), 3)
    m = min(K * n, 32 * 1048576)
    out = torch.empty((m, 18), device="cuda")
    ms_all = timed(lambda: env.observe(p.view(-1, 4)[:m], out=out), 3)
    idx = torch.randint(0, K * n, (512,), device="cuda")
    mb = torch.empty((512, 18), device="cuda")
    ms_mb = timed(lambda: env.observe(p, idx, out=mb), 50)
    print(json.dumps({"n_envs": n, "steps": K, "rollout_obs_ms": ms_obs, "rollout_poses_ms": ms_pose,
                      "rollout_no_obs_ms": ms_none, "obs_buffer_GB": obs_bytes / 1e9, "pose_buffer_GB": p.numel() * 8 / 1e9,
                      "env_steps_per_s_obs": n * K / ms_obs * 1e3, "env_steps_per_s_poses": n * K / ms_pose * 1e3,
                      "observe_records": m, "observe_ms": ms_all, "observe_records_per_s": m / ms_all * 1e3,
                      "observe_minibatch512_us": ms_mb * 1e3}), flush=True)
    del env, a, o, p, out
    torch.cuda.empty_cache()
