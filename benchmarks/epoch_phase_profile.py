"""Finer phase stamps of carenv_ppo_epoch (variant build with -DCARENV_EPOCH_PROF2): inside phase C, when the clip
coefficient is known and when the Adam loop is done.  CARENV_LIB=build/variants/libcarenv_epochprof.so python benchmarks/epoch_phase_profile.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppo_car_b200.ppo_update import FusedPPOUpdate
from ppo_car_b200.train_ppo import ActorCritic
dev = torch.device("cuda"); torch.manual_seed(0)
net = ActorCritic(18, 9).to(dev)
M, B, U = 1 << 20, 512, 40
obs = torch.rand((M, 18), device=dev); act = torch.randint(0, 9, (M,), device=dev).float()
old_logp = torch.full((M,), -2.2, device=dev); adv, ret = torch.randn(M, device=dev), torch.randn(M, device=dev)
idx = torch.randint(0, M, (U, B), device=dev)
upd = FusedPPOUpdate(net.actor, net.critic, B, lr=3e-4)
prof = torch.zeros((2 * U, 4), dtype=torch.int64, device=dev)
for _ in range(2):
    upd.run_epoch(obs, idx, act, old_logp, adv, ret, prof=prof)
torch.cuda.synchronize()
p = prof.cpu().double()
a, b = p[:U], p[U:]
sl = slice(5, U)
print("us: A+bar1 %.2f  B+bar2 %.2f  C total %.2f | C: loads+fold %.2f  adam %.2f  tail %.2f" % (
    (a[sl, 1] - a[sl, 0]).mean() / 1e3, (a[sl, 2] - a[sl, 1]).mean() / 1e3, (a[sl, 3] - a[sl, 2]).mean() / 1e3,
    (b[sl, 0] - a[sl, 2]).mean() / 1e3, (b[sl, 1] - b[sl, 0]).mean() / 1e3, (a[sl, 3] - b[sl, 1]).mean() / 1e3))
