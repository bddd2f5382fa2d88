"""Target for `ncu --set full -k regex:k_rollout_warp`: 1,024-step rollouts of 592 envs (one warp per scheduler)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200
n, T = 592, 1024
env = ppo_car_b200.VecCarEnv(n, ppo_car_b200.builtin_track("big_track")); env.reset()
a = torch.randint(0, 9, (T, n), device="cuda", dtype=torch.uint8)
for _ in range(3): env.rollout(a)
torch.cuda.synchronize()
print("ok")
