#!/usr/bin/env python
"""k_policy_rollout_tc2: start delay of the second 128-env group (option tc_stagger, cycles) versus step time.
    python benchmarks/tc2_stagger.py [n_envs ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200  # noqa: E402
from ppo_car_b200.train_ppo import ActorCritic  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
net = ActorCritic(18, 9).to(dev)
packed = ppo_car_b200.pack_policy_weights_tc(net.actor, net.critic)
sizes = [int(a) for a in sys.argv[1:]] or [32768]
for n in sizes:
    T = 256 if n <= 65536 else 32
    env = ppo_car_b200.VecCarEnv(n, ppo_car_b200.builtin_track("big_track"), reward_scaling=0.1, float_flags=True, with_info=False)
    env.set_option("tc_tiles", 3)
    buf = ppo_car_b200.Buffer((18,), T, n, dev)
    obs = env.reset()[0].clone()
    term, trunc, lv = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.empty(n, device=dev)
    for stagger in (0, 3000, 6000, 9000, 12000, 15000):
        env.set_option("tc_stagger", stagger)
        for i in range(2):
            ppo_car_b200.fused_rollout(env, packed, buf, obs, term, trunc, seed=1, step0=i * T, last_val=lv)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(4):
            ppo_car_b200.fused_rollout(env, packed, buf, obs, term, trunc, seed=1, step0=(i + 2) * T, last_val=lv)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 4
        print(json.dumps({"n_envs": n, "steps_per_launch": T, "stagger_cycles": stagger, "us_per_step": round(ms / T * 1e3, 2),
                          "env_steps_per_s": n * T / ms * 1e3}), flush=True)
