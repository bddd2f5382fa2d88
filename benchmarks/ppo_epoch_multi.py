#!/usr/bin/env python
"""The 80 minibatch updates of a PPO epoch on N GPUs (torchrun): three launches + one NCCL all-reduce per minibatch
(FusedPPOUpdate.grad / apply) against ONE persistent launch whose gradient all-reduce runs over NVLink peer memory
inside the kernel (FusedPPOUpdate.run_epoch, csrc/ppo_epoch.cuh).  Checks that both give the same parameters and that
every rank holds the same bits, then times both with CUDA events (max over ranks).
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/ppo_epoch_multi.py [--ctas 64]"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppo_car_b200.ppo_update import FusedPPOUpdate  # noqa: E402
from ppo_car_b200.train_ppo import ActorCritic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ctas", type=int, nargs="*", default=[0])
    ap.add_argument("--updates", type=int, default=80)
    ap.add_argument("--rows", type=int, default=1 << 22)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B, U, M = 512, args.updates, args.rows
    torch.manual_seed(0)
    net0 = ActorCritic(18, 9).to(dev)
    torch.manual_seed(100 + rank)
    obs = torch.rand((M, 18), device=dev)
    act = torch.randint(0, 9, (M,), device=dev).float()
    old_logp = torch.log_softmax(net0.actor(obs), -1).gather(-1, act.long().unsqueeze(-1)).squeeze(-1).detach()
    old_logp = old_logp + torch.randn(M, device=dev) * 0.1
    adv, ret = torch.randn(M, device=dev), torch.randn(M, device=dev)
    idx = torch.randint(0, M, (U, B), device=dev)

    def note(msg):
        print(f"[rank {rank}] {msg}", file=sys.stderr, flush=True)

    def fresh():
        net = ActorCritic(18, 9).to(dev)
        net.load_state_dict(net0.state_dict())
        return net, FusedPPOUpdate(net.actor, net.critic, B, lr=3e-4)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    net_a, ua = fresh()

    def per_minibatch():
        for u in range(U):
            ua.grad(obs, idx[u], act, old_logp, adv, ret)
            ua.apply(world)

    # correctness first: one epoch each from the same start
    per_minibatch()
    torch.cuda.synchronize()
    note("per-minibatch reference epoch done")
    rows = []
    for ctas in args.ctas:
        net_b, ub = fresh()
        if world > 1:
            ub.connect()
        ub.run_epoch(obs, idx, act, old_logp, adv, ret, world=world, n_ctas=ctas)
        ub.check_epoch()
        note(f"epoch kernel done (ctas {ctas})")
        err = max(float((pa.detach() - pb.detach()).abs().max()) for pa, pb in zip(ua.params, ub.params))
        flat = torch.cat([p.detach().reshape(-1) for p in ub.params])
        same = True
        if world > 1:
            gathered = [torch.empty_like(flat) for _ in range(world)]
            dist.all_gather(gathered, flat)
            same = all(torch.equal(gathered[0], g) for g in gathered)
        ms_epoch = timed(lambda: ub.run_epoch(obs, idx, act, old_logp, adv, ret, world=world, n_ctas=ctas), args.reps)
        ub.check_epoch()
        note(f"epoch kernel timed: {ms_epoch:.3f} ms")
        prof = torch.zeros((U, 4), dtype=torch.int64, device=dev)
        ub.run_epoch(obs, idx, act, old_logp, adv, ret, world=world, n_ctas=ctas, prof=prof)
        torch.cuda.synchronize()
        pr = prof.cpu().double()
        phase_us = [float((pr[5:, 1] - pr[5:, 0]).mean()) / 1e3, float((pr[5:, 2] - pr[5:, 1]).mean()) / 1e3,
                    float((pr[5:, 3] - pr[5:, 2]).mean()) / 1e3]
        rows.append({"n_ctas": ctas or 64, "phase_us_forward_backward|reduce_exchange|adam": phase_us, "epoch_kernel_ms": ms_epoch, "epoch_kernel_us_per_update": 1e3 * ms_epoch / U,
                     "max_abs_param_diff_vs_per_minibatch": err, "ranks_bit_identical": same})
        ub.close()
    ms_mb = timed(per_minibatch, max(1, args.reps // 2))
    note(f"per-minibatch timed: {ms_mb:.3f} ms")
    g = torch.cuda.CUDAGraph()                             # the per-minibatch path replayed from a CUDA graph (train_ppo --graph-update)
    idx1 = idx[0].clone()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        ua.grad(obs, idx1, act, old_logp, adv, ret); ua.apply(world)
    torch.cuda.current_stream(dev).wait_stream(side)
    with torch.cuda.graph(g):
        ua.grad(obs, idx1, act, old_logp, adv, ret); ua.apply(world)

    def graphed():
        for _ in range(U):
            g.replay()

    note("graph captured")
    ms_graph = timed(graphed, max(1, args.reps // 2))
    if rank == 0:
        for r in rows:
            print(json.dumps(dict(r, n_gpus=world, updates=U, batch=B, per_minibatch_ms=ms_mb,
                                  per_minibatch_us_per_update=1e3 * ms_mb / U, per_minibatch_graph_ms=ms_graph,
                                  per_minibatch_graph_us_per_update=1e3 * ms_graph / U)), flush=True)
    del g                                                   # a live graph with captured NCCL work blocks the teardown
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
