#!/usr/bin/env python
"""Fused rollout at small batch sizes: warp-per-environment kernel (k_policy_rollout_warp) against the tensor-core
kernel (k_policy_rollout_tc3), us per step and env-steps/s.  One JSON line per size."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200  # noqa: E402
from ppo_car_b200.train_ppo import ActorCritic  # noqa: E402

dev = torch.device("cuda")
track = ppo_car_b200.builtin_track("big_track")
torch.manual_seed(0)
net = ActorCritic(18, 9).to(dev)
packed = ppo_car_b200.pack_policy_weights_tc(net.actor, net.critic)
for n in (24, 128, 592, 1024, 2048, 4096, 8192):
    T = 1024 if n <= 1024 else 256
    env = ppo_car_b200.VecCarEnv(n, track, reward_scaling=0.1, float_flags=True, with_info=False)
    buf = ppo_car_b200.Buffer((18,), T, n, dev)
    obs = env.reset()[0].clone()
    term, trunc, lv = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.empty(n, device=dev)
    res = {"n_envs": n, "steps_per_launch": T}
    for tag in ("warp_per_env", "tensor_core_tc3"):
        def run(i):
            if tag == "warp_per_env":
                ppo_car_b200.fused_rollout_warp(env, net.actor, net.critic, buf, obs, term, trunc, seed=1, step0=i * T, last_val=lv)
            else:
                ppo_car_b200.fused_rollout(env, packed, buf, obs, term, trunc, seed=1, step0=i * T, last_val=lv)
        for i in range(2):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(4):
            run(i + 2)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 4
        res[tag + "_us_per_step"] = round(ms / T * 1e3, 3)
        res[tag + "_env_steps_per_s"] = n * T / ms * 1e3
    print(json.dumps(res), flush=True)
