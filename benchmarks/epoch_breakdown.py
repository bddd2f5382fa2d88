"""Where one PPO epoch of BASELINE config 5 goes on ONE GPU's shard (32,768 envs x 1024 steps, fused rollout +
fused update, no NCCL): CUDA-event times of pack, rollout, GAE, statistics and the 80 minibatch updates.
    python benchmarks/epoch_breakdown.py [--envs 32768] [--out profiles/x.jsonl]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200  # noqa: E402
from ppo_car_b200 import Buffer, VecCarEnv, builtin_track  # noqa: E402
from ppo_car_b200.policy import fused_rollout, pack_policy_weights_tc  # noqa: E402
from ppo_car_b200.ppo_update import FusedPPOUpdate  # noqa: E402
from ppo_car_b200.train_ppo import ActorCritic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=32768)
    ap.add_argument("--epochs", type=int, default=6)
    ap.add_argument("--out", default=None)
    ap.add_argument("--per-minibatch", action="store_true", help="three launches per minibatch instead of carenv_ppo_epoch")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    n, T, B = args.envs, 1024, 512
    torch.manual_seed(0)
    envs = VecCarEnv(n, builtin_track("big_track"), device=dev, reward_scaling=0.1, float_flags=True, with_info=False)
    agent = ActorCritic(18, 9).to(dev)
    upd = FusedPPOUpdate(agent.actor, agent.critic, B, 3e-4)
    buf = Buffer((18,), T, n, dev, 0.99, 0.95)
    next_obs = envs.reset()[0].clone()
    next_term, next_trunc = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    packed = pack_policy_weights_tc(agent.actor, agent.critic)
    last_val = torch.empty(n, device=dev)
    idx = torch.zeros(B, dtype=torch.int64, device=dev)
    rows = []
    for epoch in range(args.epochs):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        with torch.no_grad():
            ev[0].record()
            pack_policy_weights_tc(agent.actor, agent.critic, out=packed)
            ev[1].record()
            fused_rollout(envs, packed, buf, next_obs, next_term, next_trunc, seed=0, step0=epoch * T, env_offset=0,
                          last_val=last_val)
            ev[2].record()
            adv, ret = buf.calculate_advantages(last_val.reshape(1, -1), next_term.reshape(1, -1), next_trunc.reshape(1, -1))
            ev[3].record()
            stats = (buf.rew_buf.sum(), buf.term_buf.sum() + buf.trunc_buf.sum())
            ev[4].record()
            obs_b, act_b, val_b, logp_b = buf.get()
            obs_f, act_f, logp_f, adv_f, ret_f = obs_b.view(-1, 18), act_b.view(-1), logp_b.view(-1), adv.view(-1), ret.view(-1)
            if args.per_minibatch:
                for _ in range(80):
                    torch.randint(0, T * n, (B,), device=dev, out=idx)
                    upd.grad(obs_f, idx, act_f, logp_f, adv_f, ret_f)
                    upd.apply(1)
            else:
                upd.run_epoch(obs_f, torch.randint(0, T * n, (80, B), device=dev), act_f, logp_f, adv_f, ret_f)
            ev[5].record()
        torch.cuda.synchronize()
        t = [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
        row = {"epoch": epoch, "envs": n, "pack_ms": t[0], "rollout_ms": t[1], "gae_ms": t[2], "stats_ms": t[3],
               "updates_ms": t[4], "update_path": "per-minibatch" if args.per_minibatch else "epoch kernel", "total_ms": sum(t), "rollout_env_steps_per_s": n * T / (t[1] * 1e-3)}
        print(json.dumps(row), flush=True)
        rows.append(row)
    if args.out:
        with open(args.out, "w") as fh:
            for r in rows:
                fh.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
