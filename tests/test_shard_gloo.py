"""CPU, world_size 2 over gloo: the env-range partition and the rollout-statistics all-reduce that
the N > 1 bench / training path uses.  The kernels themselves need no collective; their
shard invariance on the device is covered by test_full_size_properties_65536_envs."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.c_oracle import COracleVecEnv
from ppo_car_b200.shard import allreduce_rollout_stats, shard_range


def test_shard_range_is_a_partition():
    for n in (1, 7, 24, 65536, 1_048_576, 1_000_003):
        for w in (1, 2, 3, 4, 8):
            edges = [shard_range(n, w, r) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, track, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_total, T = 48, 200
    acts = np.random.default_rng(0).integers(0, 9, size=(T, n_total)).astype(np.uint8)   # same stream on every rank
    lo, hi = shard_range(n_total, world, rank)
    env = COracleVecEnv(hi - lo, track, threads=1, scan_all_gates=False)                  # checker stands in for the GPU shard
    env.reset()
    r = env.rollout(acts[:, lo:hi], want=("rew", "term", "trunc"))
    stats = allreduce_rollout_stats(torch.tensor(r["rew"].sum()), torch.tensor(float(r["rew"].size)),
                                    torch.tensor(float(r["term"].sum() + r["trunc"].sum())))
    q.put((rank, lo, hi, r["rew"], stats))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_rollout_equals_single_rank(tracks_dir):
    track = os.path.join(tracks_dir, "big_track.json")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, track, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_total, T = 48, 200
    acts = np.random.default_rng(0).integers(0, 9, size=(T, n_total)).astype(np.uint8)
    env = COracleVecEnv(n_total, track, threads=1, scan_all_gates=False)
    env.reset()
    full = env.rollout(acts, want=("rew", "term", "trunc"))
    stitched = np.concatenate([r[3] for r in res], axis=1)
    assert np.array_equal(stitched, full["rew"])                       # shard invariance
    for r in res:                                                       # both ranks hold the global statistics
        assert np.isclose(r[4][0], full["rew"].sum()) and r[4][1] == full["rew"].size
        assert r[4][2] == full["term"].sum() + full["trunc"].sum()


def _handle_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ppo_car_b200.ppo_update import exchange_ipc_handles

    mine = bytes([rank + 1]) * 64                          # stands in for this rank's cudaIpcMemHandle_t
    everyone = exchange_ipc_handles(mine)
    bad = None
    try:
        exchange_ipc_handles(b"short")
    except ValueError as e:
        bad = str(e)
    q.put((rank, everyone, bad))
    dist.barrier()
    dist.destroy_process_group()


def test_ipc_handle_exchange_is_in_rank_order():
    """Host side of the in-kernel gradient all-reduce (FusedPPOUpdate.connect): every rank ends up with all the
    64-byte handles concatenated in rank order — what carenv_ppo_comm_connect indexes by rank."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_handle_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = bytes([1]) * 64 + bytes([2]) * 64
    for rank, everyone, bad in res:
        assert everyone == want and bad is not None
