"""ncu target: a few launches of the step kernel at the bench workload (1,048,576 envs x 256 steps, big_track).
    python benchmarks/step_profile_target.py [tab]   # tab: 1 = k_rollout_tab (default), -1 = k_rollout"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200  # noqa: E402

tab = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_048_576
steps = 256
env = ppo_car_b200.VecCarEnv(n, ppo_car_b200.builtin_track("big_track"))
env.set_option("tab", tab)
env.reset()
g = torch.Generator(device="cuda").manual_seed(1)
acts = torch.randint(0, 9, (steps, n), generator=g, device="cuda", dtype=torch.uint8)
obs = torch.empty((steps, n, 18), device="cuda")
rew = torch.empty((steps, n), device="cuda")
te = torch.empty((steps, n), dtype=torch.uint8, device="cuda")
tr = torch.empty((steps, n), dtype=torch.uint8, device="cuda")
for _ in range(3):
    env.rollout(acts, obs_out=obs, reward_out=rew, term_out=te, trunc_out=tr)
torch.cuda.synchronize()
print("ok")
