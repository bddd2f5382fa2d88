"""Concurrent pinned device->host bandwidth of ONE box, all ranks copying at the same time: the ceiling of the
end-to-end (host-buffer) step, which ships 88 B per environment to the host every step.

    python benchmarks/d2h_probe.py                                                    # 1 rank
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        benchmarks/d2h_probe.py --out profiles/r2_d2h_probe_8gpu.jsonl

Per chunk size (1.4 MB = one 16,384-env sub-range, 12 MB, 92 MB = one 1,048,576-env step) every rank enqueues `reps`
back-to-back copies on one stream between two barriers; reported: each rank's own rate, the aggregate (all bytes /
slowest rank's time) and, per box, what that allows in env-steps/s at 88 B per env-step.  `--streams 2` splits the
copies over two streams per rank (do the copy engines of one GPU overlap?).
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--reps", type=int, default=40)
    ap.add_argument("--streams", type=int, default=1)
    args = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rows = []
    streams = [torch.cuda.Stream(dev) for _ in range(args.streams)]
    for direction in ("d2h", "h2d"):
        for mb in (1.44, 12.0, 92.3):
            nbytes = int(mb * 1e6) // 16 * 16
            d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            hs = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
            reps = max(4, min(args.reps, int(2e9 / nbytes)))
            for _ in range(3):
                (hs[0].copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(hs[0], non_blocking=True))
            barrier()
            t0 = time.perf_counter()
            for i in range(reps):
                with torch.cuda.stream(streams[i % len(streams)]):
                    if direction == "d2h":
                        hs[i % 2].copy_(d, non_blocking=True)
                    else:
                        d.copy_(hs[i % 2], non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            rates = torch.zeros(world, dtype=torch.float64, device=dev)
            rates[rank] = reps * nbytes / dt / 1e9
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dist.all_reduce(rates)
            agg = world * reps * nbytes / float(t.item()) / 1e9
            if rank == 0:
                row = {"direction": direction, "ranks": world, "chunk_mb": nbytes / 1e6, "reps": reps,
                       "streams_per_rank": args.streams, "per_rank_gbs": [round(float(x), 2) for x in rates.tolist()],
                       "aggregate_gbs": round(agg, 2),
                       "env_steps_per_s_at_88B": round(agg * 1e9 / 88.0) if direction == "d2h" else None}
                print(json.dumps(row), flush=True)
                rows.append(row)
            del d, hs
    if rank == 0 and args.out:
        with open(args.out, "w") as fh:
            for r in rows:
                fh.write(json.dumps(r) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
