"""Parity soak at the bench workload: 1,048,576 envs x 1024 steps on big_track with the bench's kind of action
stream, EVERY integer output (terminated, truncated, gates_passed, time_passed, next_gate_index) and every reward of
every env-step compared with the float64 C oracle, chunk by chunk; ray-distance errors measured on a sub-slice of
every chunk.  Reports counts, not just pass / fail.  TEST INFRASTRUCTURE (uses oracle/).

    python benchmarks/soak_parity.py [--envs 1048576] [--steps 1024] [--chunk 65536] [--out profiles/r2_soak_parity.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200  # noqa: E402
from oracle.c_oracle import COracleVecEnv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1_048_576)
    ap.add_argument("--steps", type=int, default=1024)
    ap.add_argument("--chunk", type=int, default=65_536)
    ap.add_argument("--obs-slice", type=int, default=1024)
    ap.add_argument("--track", default="big_track")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    path = ppo_car_b200.builtin_track(args.track)
    T, total = args.steps, args.envs
    res = {"track": args.track, "envs": total, "steps": T, "env_steps": 0, "episodes": 0, "gate_events": 0,
           "mismatch": {k: 0 for k in ("terminated", "truncated", "gates_passed", "time_passed", "next_gate_index", "reward")},
           "envs_with_any_mismatch": 0, "obs_elements": 0, "obs_worst_rel_ray": 0.0, "obs_over_1e-5_ray": 0,
           "obs_worst_abs_other": 0.0, "slow_path": {"line": 0, "band": 0, "gate": 0, "tiny": 0}}
    t0 = time.time()
    for lo in range(0, total, args.chunk):
        n = min(args.chunk, total - lo)
        g = torch.Generator(device="cuda").manual_seed(1234 + lo)
        acts = torch.randint(0, 9, (T, n), generator=g, device="cuda", dtype=torch.uint8)
        env = ppo_car_b200.VecCarEnv(n, path)
        env.set_option("tab", 1)                                  # the benchmarked kernel
        env.reset()
        so = args.obs_slice
        obs = torch.empty((T, so, 18), device="cuda")
        out = env.rollout(acts, store_obs=False, store_info=True)
        env2 = ppo_car_b200.VecCarEnv(so, path)                   # observations of the first `so` envs of the chunk
        env2.reset()
        o2 = env2.rollout(acts[:, :so].contiguous())
        a_host = acts.cpu().numpy()
        ora = COracleVecEnv(n, path, scan_all_gates=False)
        ora.reset()
        ref = ora.rollout(a_host, want=("rew", "term", "trunc", "gates_passed", "time_passed", "next_gate_index"))
        bad_env = np.zeros(n, bool)
        for key, got in (("terminated", out["terminated"]), ("truncated", out["truncated"]),
                         ("gates_passed", out["info"]["gates_passed"]), ("time_passed", out["info"]["time_passed"]),
                         ("next_gate_index", out["info"]["next_gate_index"])):
            rk = {"terminated": "term", "truncated": "trunc"}.get(key, key)
            diff = got.cpu().numpy().astype(np.int64) != ref[rk].astype(np.int64)
            res["mismatch"][key] += int(diff.sum())
            bad_env |= diff.any(0)
        diff = out["reward"].cpu().numpy() != ref["rew"].astype(np.float32)
        res["mismatch"]["reward"] += int(diff.sum())
        bad_env |= diff.any(0)
        res["envs_with_any_mismatch"] += int(bad_env.sum())
        res["env_steps"] += n * T
        res["episodes"] += int(ref["term"].sum() + ref["trunc"].sum())
        res["gate_events"] += int((out["info"]["events"] & 1).sum())
        ora2 = COracleVecEnv(so, path, scan_all_gates=False)
        ora2.reset()
        ro = ora2.rollout(a_host[:, :so].copy(), want=("obs",))["obs"].astype(np.float64)
        go = o2["obs"].cpu().numpy().astype(np.float64)
        rel = np.abs(go[..., 6:] - ro[..., 6:]) / np.maximum(np.abs(ro[..., 6:]), 1e-300)
        res["obs_worst_rel_ray"] = max(res["obs_worst_rel_ray"], float(rel.max()))
        res["obs_over_1e-5_ray"] += int((rel > 1e-5).sum())
        res["obs_worst_abs_other"] = max(res["obs_worst_abs_other"], float(np.abs(go[..., :6] - ro[..., :6]).max()))
        res["obs_elements"] += int(go.size)
        for k, v in env.slow_path_counts().items():
            res["slow_path"][k] += int(v)
        print(f"chunk {lo // args.chunk}: {res['env_steps']:,} env-steps, mismatches {sum(res['mismatch'].values())}, "
              f"worst ray rel {res['obs_worst_rel_ray']:.2e}, {time.time() - t0:.0f} s", flush=True)
        del env, env2, out, o2, acts
    res["seconds"] = time.time() - t0
    print(json.dumps(res))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
