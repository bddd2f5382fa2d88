#!/usr/bin/env python
"""Throughput of the fused rollout kernel (policy forward + sampling + env step + buffer rows) versus the
PyTorch-eager rollout loop of train_ppo.py at several env counts.  One JSON line per size."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200  # noqa: E402
from ppo_car_b200.train_ppo import ActorCritic  # noqa: E402

dev = torch.device("cuda")
track = ppo_car_b200.builtin_track("big_track")
torch.manual_seed(0)
net = ActorCritic(18, 9).to(dev)
packed_cc = ppo_car_b200.pack_policy_weights(net.actor, net.critic)
packed_tc = ppo_car_b200.pack_policy_weights_tc(net.actor, net.critic)
for n, T in ((24, 1024), (32768, 64), (262144, 32), (1048576, 16)):
    env = ppo_car_b200.VecCarEnv(n, track, reward_scaling=0.1, float_flags=True, with_info=False)
    buf = ppo_car_b200.Buffer((18,), T, n, dev)
    obs = env.reset()[0].clone()
    term, trunc, lv = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.empty(n, device=dev)
    res = {}
    for tag, packed in (("cuda_core", packed_cc), ("tensor_core", packed_tc)):
        for i in range(2):
            ppo_car_b200.fused_rollout(env, packed, buf, obs, term, trunc, seed=1, step0=i * T, last_val=lv)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for i in range(reps):
            ppo_car_b200.fused_rollout(env, packed, buf, obs, term, trunc, seed=1, step0=(i + 2) * T, last_val=lv)
        e1.record()
        torch.cuda.synchronize()
        res[tag] = e0.elapsed_time(e1) / reps
    ms_fused = res["tensor_core"]
    # eager loop (what train_ppo.py does without --fused-rollout), a few steps only
    steps = min(T, 32)
    with torch.no_grad():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            act, logp, _, val = net.act(obs)
            buf.ptr = 0
            buf.store(obs, act, 0.0, val.view(-1), term, trunc, logp)
            o, rew, te, tr, _ = env.step(act)
            obs.copy_(o); term.copy_(te); trunc.copy_(tr)
        torch.cuda.synchronize()
        ms_eager = (time.perf_counter() - t0) * 1e3 / steps
    print(json.dumps({"n_envs": n, "steps_per_launch": T, "fused_ms_per_step": ms_fused / T,
                      "fused_env_steps_per_s": n * T / ms_fused * 1e3,
                      "fused_cuda_core_ms_per_step": res["cuda_core"] / T,
                      "fused_cuda_core_env_steps_per_s": n * T / res["cuda_core"] * 1e3, "eager_ms_per_step": ms_eager,
                      "eager_env_steps_per_s": n / ms_eager * 1e3}), flush=True)
    del buf
