"""Rollout-kernel throughput versus environments per GPU.  Usage: python benchmarks/sweep_envs.py"""
import torch, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200
track = ppo_car_b200.builtin_track("big_track")
for n in (32768, 65536, 131072, 262144, 524288, 1048576):
    env = ppo_car_b200.VecCarEnv(n, track); env.reset()
    a = torch.randint(0, 9, (32, n), device="cuda", dtype=torch.uint8)
    o = env.rollout(a); 
    for _ in range(3): env.rollout(a, obs_out=o["obs"], reward_out=o["reward"], term_out=o["terminated"], trunc_out=o["truncated"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(10, 4*1048576//n)
    e0.record()
    for _ in range(reps): env.rollout(a, obs_out=o["obs"], reward_out=o["reward"], term_out=o["terminated"], trunc_out=o["truncated"])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/reps
    print(n, f"{ms:.3f} ms/launch  {32*n/ms*1e3:.4g} env-steps/s", flush=True)
