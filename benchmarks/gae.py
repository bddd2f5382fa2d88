"""GAE kernel bandwidth at several [T, N] shapes.  Usage: python benchmarks/gae.py"""
import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200
dev="cuda"
for T,N in ((1024,65536),(1024,24),(1024,1048576//8),(128,1048576)):
    g = torch.Generator(device=dev).manual_seed(7)
    rew = torch.rand((T,N),generator=g,device=dev); val=torch.randn((T,N),generator=g,device=dev)
    term=(torch.rand((T,N),generator=g,device=dev)<0.004).float(); trunc=(torch.rand((T,N),generator=g,device=dev)<0.001).float()
    lv=torch.randn(N,generator=g,device=dev); z=torch.zeros(N,device=dev)
    adv=torch.empty_like(rew); ret=torch.empty_like(rew)
    for _ in range(3): ppo_car_b200.gae_reverse_scan(rew,val,term,trunc,lv,z,z,adv_out=adv,ret_out=ret)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ppo_car_b200.gae_reverse_scan(rew,val,term,trunc,lv,z,z,adv_out=adv,ret_out=ret)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print(T,N,f"{ms:.4f} ms  {T*N*24/ms/1e6:.0f} GB/s", flush=True)
