#!/usr/bin/env python
"""Time the BASELINE.json configs that bench.py does not headline (one JSON object per config).

  config 2   tracks/big_track.json, 24 envs x 1024 steps, random actions, one rollout launch
  config 3   big_track.json, 65,536 envs x 1024 steps rollout (one 1,024-step launch, every step stored in
             the [1024, 65536, 18] buffer) + GAE over [1024, 65536]
  step API   VecCarEnv.step with device actions at 24 / 65,536 / 1,048,576 envs (one launch per step)

Device timed with CUDA events after warm-up.  Usage: python benchmarks/configs.py > gpurun_out/configs.jsonl
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200  # noqa: E402

dev = torch.device("cuda", 0)
track = ppo_car_b200.builtin_track("big_track")
g = torch.Generator(device=dev).manual_seed(0)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ---- config 2
n, T = 24, 1024
env = ppo_car_b200.VecCarEnv(n, track)
env.reset()
acts = torch.randint(0, 9, (T, n), generator=g, device=dev, dtype=torch.uint8)
ms = timed(lambda: env.rollout(acts), 5)
print(json.dumps({"config": "2: big_track, 24 envs x 1024 steps, one launch", "ms": ms, "env_steps_per_s": n * T / ms * 1e3}))

# ---- config 3
n, T, chunk = 65536, 1024, 1024
env = ppo_car_b200.VecCarEnv(n, track, float_flags=True)
env.reset()
buf = ppo_car_b200.Buffer((18,), T, n, dev)
acts = torch.randint(0, 9, (T, n), generator=g, device=dev, dtype=torch.uint8)
buf.val_buf.normal_(generator=g)


def rollout_into_buffer():
    for t0 in range(0, T, chunk):
        env.rollout(acts[t0:t0 + chunk], obs_out=buf.obs_buf[t0:t0 + chunk], reward_out=buf.rew_buf[t0:t0 + chunk],
                    term_out=buf.term_buf[t0:t0 + chunk], trunc_out=buf.trunc_buf[t0:t0 + chunk])


ms_roll = timed(rollout_into_buffer, 3)
buf.ptr = T
z = torch.zeros(1, n, device=dev)
ms_gae = timed(lambda: buf.calculate_advantages(z, z, z), 10)
print(json.dumps({"config": "3: big_track, 65536 envs x 1024 steps rollout + GAE", "rollout_ms": ms_roll,
                  "gae_ms": ms_gae, "env_steps_per_s_rollout": n * T / ms_roll * 1e3,
                  "env_steps_per_s_with_gae": n * T / (ms_roll + ms_gae) * 1e3,
                  "gae_gbs": T * n * 24 / ms_gae / 1e6}))
del buf, acts

# ---- single-step API
for n in (24, 65536, 1048576):
    env = ppo_car_b200.VecCarEnv(n, track)
    env.reset()
    a = torch.randint(0, 9, (n,), generator=g, device=dev)          # int64 like Categorical.sample()
    ms = timed(lambda: env.step(a), 200)
    print(json.dumps({"config": f"step API, {n} envs, int64 device actions", "ms_per_step": ms,
                      "env_steps_per_s": n / ms * 1e3}))
