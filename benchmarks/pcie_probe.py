"""Raw pinned D2H bandwidth, host cast cost and step() latencies behind the e2e number.  Usage: python benchmarks/pcie_probe.py"""
import torch, time, numpy as np, sys
sys.path.insert(0,'/root/repo')
import ppo_car_b200
n=1048576
d = torch.empty(n*94//4, dtype=torch.float32, device='cuda')
h = torch.empty(n*94//4, dtype=torch.float32, pin_memory=True)
for _ in range(3): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(10): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
dt=(time.perf_counter()-t0)/10
print(f"D2H {d.numel()*4/1e6:.1f} MB pinned: {dt*1e3:.2f} ms -> {d.numel()*4/dt/1e9:.1f} GB/s")
a64 = np.random.randint(0,9,size=n).astype(np.int64)
pin = torch.empty(n, dtype=torch.uint8, pin_memory=True)
t0=time.perf_counter()
for _ in range(10): np.copyto(pin.numpy(), a64, casting='unsafe')
print(f"host cast int64->u8 pinned: {(time.perf_counter()-t0)/10*1e3:.2f} ms")
env = ppo_car_b200.VecCarEnv(n, ppo_car_b200.builtin_track('big_track'))
env.reset()
for _ in range(3): env.step(a64)
t0=time.perf_counter()
for _ in range(10): env.step(a64)
print(f"step(numpy) total: {(time.perf_counter()-t0)/10*1e3:.2f} ms")
ad = torch.from_numpy(a64).cuda()
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(10): env.step(ad)
torch.cuda.synchronize()
print(f"step(device int64): {(time.perf_counter()-t0)/10*1e3:.3f} ms")
