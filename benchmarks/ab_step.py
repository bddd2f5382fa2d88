"""A/B of the step kernels: k_rollout (denominators recomputed) vs k_rollout_tab (shared-memory table).

    python benchmarks/ab_step.py [--envs 1048576,131072] [--steps 256] [--reps 5] [--out profiles/x.jsonl]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200  # noqa: E402


def measure(n, steps, reps, track, options):
    env = ppo_car_b200.VecCarEnv(n, ppo_car_b200.builtin_track(track))
    for k, v in options.items():
        env.set_option(k, v)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = torch.randint(0, 9, (steps, n), generator=g, device="cuda", dtype=torch.uint8)
    obs = torch.empty((steps, n, 18), device="cuda")
    rew = torch.empty((steps, n), device="cuda")
    te = torch.empty((steps, n), dtype=torch.uint8, device="cuda")
    tr = torch.empty((steps, n), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        env.rollout(acts, obs_out=obs, reward_out=rew, term_out=te, trunc_out=tr)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        env.rollout(acts, obs_out=obs, reward_out=rew, term_out=te, trunc_out=tr)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    med = ms[len(ms) // 2]
    return {"envs": n, "steps": steps, "track": track, "options": options, "ms_median": med, "ms_min": ms[0],
            "env_steps_per_s": n * steps / (med * 1e-3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", default="1048576,131072")
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--track", default="big_track")
    ap.add_argument("--variants", default="tab=-1;tab=1")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rows = []
    for n in (int(x) for x in args.envs.split(",")):
        for var in args.variants.split(";"):
            opts = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in var.split(",") if kv}
            r = measure(n, args.steps, args.reps, args.track, opts)
            print(json.dumps(r), flush=True)
            rows.append(r)
    if args.out:
        with open(args.out, "w") as fh:
            for r in rows:
                fh.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
