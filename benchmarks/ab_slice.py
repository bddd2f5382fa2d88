#!/usr/bin/env python
"""k_rollout_tab: block rounds (option tab_slice = -1) against the time-sliced distribution (k_rollout_tab_sliced,
default) at the shard sizes of the 1 M-env job on 1 / 2 / 4 / 8 GPUs; checks bit-identity, then times 256-step launches.
    python benchmarks/ab_slice.py [n_envs ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200  # noqa: E402

steps = 256
for n in [int(a) for a in sys.argv[1:]] or [131072, 262144, 524288, 1048576]:
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = torch.randint(0, 9, (steps, n), generator=g, device="cuda", dtype=torch.uint8)
    obs = torch.empty((steps, n, 18), device="cuda")
    rew = torch.empty((steps, n), device="cuda")
    te = torch.empty((steps, n), dtype=torch.uint8, device="cuda")
    tr = torch.empty((steps, n), dtype=torch.uint8, device="cuda")
    res, ref = {"n_envs": n, "steps_per_launch": steps}, None
    for tag, opt in (("block_rounds", -1), ("time_sliced", 0)):
        env = ppo_car_b200.VecCarEnv(n, ppo_car_b200.builtin_track("big_track"))
        env.set_option("tab_slice", opt)
        env.reset()
        env.rollout(acts, obs_out=obs, reward_out=rew, term_out=te, trunc_out=tr)
        torch.cuda.synchronize()
        sig = (obs.view(torch.int32).sum(dtype=torch.int64).item(), rew.view(torch.int32).sum(dtype=torch.int64).item(),
               int(te.sum()), int(tr.sum()), env.pos.clone(), env.ints.clone())
        if ref is None:
            ref = sig
        else:
            res["bit_identical"] = bool(sig[:4] == ref[:4] and torch.equal(sig[4], ref[4]) and torch.equal(sig[5], ref[5]))
        for _ in range(2):
            env.rollout(acts, obs_out=obs, reward_out=rew, term_out=te, trunc_out=tr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            env.rollout(acts, obs_out=obs, reward_out=rew, term_out=te, trunc_out=tr)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res[tag + "_ms"] = round(ms, 4)
        res[tag + "_env_steps_per_s"] = n * steps / ms * 1e3
        del env
    print(json.dumps(res), flush=True)
