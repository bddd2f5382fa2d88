"""Target for `ncu --set full -k regex:k_ppo_epoch`: three epoch updates (80 minibatches of 512 samples each) on one GPU."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppo_car_b200.ppo_update import FusedPPOUpdate
from ppo_car_b200.train_ppo import ActorCritic
dev = torch.device("cuda"); torch.manual_seed(0)
net = ActorCritic(18, 9).to(dev)
M, B, U = 1 << 22, 512, 80
obs = torch.rand((M, 18), device=dev); act = torch.randint(0, 9, (M,), device=dev).float()
old_logp = torch.full((M,), -2.2, device=dev); adv, ret = torch.randn(M, device=dev), torch.randn(M, device=dev)
idx = torch.randint(0, M, (U, B), device=dev)
upd = FusedPPOUpdate(net.actor, net.critic, B, lr=3e-4)
for _ in range(3):
    upd.run_epoch(obs, idx, act, old_logp, adv, ret)
upd.check_epoch()
print("ok")
