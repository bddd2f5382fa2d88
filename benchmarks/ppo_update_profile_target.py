"""Target for `ncu -k regex:k_ppo`: two short epochs of train_ppo with the fused rollout and the fused update."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200, ppo_car_b200.train_ppo as tp
args = tp.parse_args(["--track", "big_track", "--n-envs", "4096", "--n-epochs", "2", "--n-steps", "64", "--fused-rollout", "--fused-update"])
tp.train(args)
