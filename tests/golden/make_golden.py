"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Everything written here comes out of the reference's own code
(/root/reference/lib/car_env.py and lib/buffer.py imported through
oracle/ref_import.py); nothing is computed by this repository's oracle or
kernels.  The fixtures pin the oracle (tests/test_oracle_golden.py) and the CUDA
path (tests/test_gpu_parity.py) on the GPU box, where /root/reference does not
exist.

Files written:
  carenv_<track>.npz   reset observation + four trajectory groups, each with the
                       action tensor that was played and every per-step output
  gae.npz              Buffer.calculate_advantages on seeded inputs
  ../../ppo_car_b200/tracks/<track>.json   the two track files (input data in the
                       reference's schema; BASELINE.json names them as workloads)
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_import import RefVecEnv, import_reference, track_path  # noqa: E402

T = 1024


def follower_action(obs: np.ndarray) -> int:
    """Closed-loop ray follower: steers towards the side with more room."""
    d = obs[6:18]
    speed = 10.0 * float(np.hypot(obs[2], obs[3]))
    left = d[11] + 0.5 * d[10]
    right = d[1] + 0.5 * d[2]
    turn = 0 if abs(left - right) < 0.005 else (1 if right > left else -1)
    if speed < 3:
        return {0: 0, 1: 5, -1: 4}[turn]
    return {0: 8, 1: 3, -1: 2}[turn]


def rollout(track: str, n_envs: int, policy) -> dict:
    env = RefVecEnv(n_envs, track)
    obs = env.reset()
    rec = {k: [] for k in ("actions", "final_obs", "rew", "term", "trunc", "gates_passed", "time_passed", "next_gate_index")}
    for t in range(T):
        a = np.asarray(policy(t, obs), dtype=np.uint8)
        obs, rew, term, trunc, info = env.step(a)
        rec["actions"].append(a)
        rec["final_obs"].append(info["final_obs"])
        rec["rew"].append(rew)
        rec["term"].append(term)
        rec["trunc"].append(trunc)
        for k in ("gates_passed", "time_passed", "next_gate_index"):
            rec[k].append(info[k])
    return {k: np.stack(v) for k, v in rec.items()}


def make_track(name: str) -> None:
    path = track_path(name + ".json")
    CarEnv, _ = import_reference()
    env = CarEnv(track_path=path)
    reset_obs, _ = env.reset(options={"track_path": path})
    car = env._CarEnv__car
    reset_dist = np.array(car.get_distances(env._CarEnv__boundaries), np.float64)
    out = dict(reset_obs=reset_obs, reset_dist=reset_dist)

    rng = np.random.default_rng(20240)
    groups = {
        # env i plays constant action i: collisions (0,1,4-7) and the 1000-step truncation (2,3,8)
        "const": (9, lambda t, obs: np.arange(9)),
        # closed-loop controller: completes laps (+10 branch) and reaches truncation
        "lap": (1, lambda t, obs: [follower_action(obs[0])]),
        # i.i.d. uniform actions (the benchmark's action distribution)
        "random": (8, lambda t, obs: rng.integers(0, 9, size=8)),
        # forward-biased random actions: many more gate hits per episode
        "fwd": (8, lambda t, obs: rng.choice(9, size=8, p=[.3, .02, .1, .1, .2, .2, .02, .02, .04])),
    }
    for gname, (n, pol) in groups.items():
        rec = rollout(path, n, pol)
        for k, v in rec.items():
            out[f"{gname}_{k}"] = v
        print(name, gname, "episodes ended:", int(rec["term"].sum()), "term,", int(rec["trunc"].sum()),
              "trunc; gate hits:", int((np.diff(rec["gates_passed"], axis=0) > 0).sum()),
              "lap rewards:", int((rec["rew"] > 10).sum()), flush=True)
    np.savez_compressed(os.path.join(HERE, f"carenv_{name}.npz"), **out)

    # re-serialise the track (input data, reference schema) next to the package
    with open(path) as fh:
        data = json.load(fh)
    dst = os.path.join(ROOT, "ppo_car_b200", "tracks", name + ".json")
    with open(dst, "w") as fh:
        json.dump(data, fh)


def make_gae() -> None:
    import torch

    _, Buffer = import_reference()
    out = {}
    for tag, (T_, N_, seed) in {"small": (37, 5, 1), "train": (1024, 24, 2), "wide": (64, 1000, 3)}.items():
        g = torch.Generator().manual_seed(seed)
        buf = Buffer((18,), T_, N_, "cpu", gamma=0.99, gae_lambda=0.95)
        buf.rew_buf = torch.rand((T_, N_), generator=g) * 1.4 - 0.3
        buf.val_buf = torch.randn((T_, N_), generator=g)
        buf.term_buf = (torch.rand((T_, N_), generator=g) < 0.05).float()
        buf.trunc_buf = (torch.rand((T_, N_), generator=g) < 0.02).float()
        buf.ptr = T_
        lv = torch.randn((1, N_), generator=g)
        lt = (torch.rand((1, N_), generator=g) < 0.1).float()
        lu = (torch.rand((1, N_), generator=g) < 0.1).float()
        adv, ret = buf.calculate_advantages(lv, lt, lu)
        for k, v in dict(rew=buf.rew_buf, val=buf.val_buf, term=buf.term_buf, trunc=buf.trunc_buf,
                         last_val=lv, last_term=lt, last_trunc=lu, adv=adv, ret=ret).items():
            out[f"{tag}_{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "gae.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["track", "big_track", "gae"]
    for w in which:
        if w == "gae":
            make_gae()
        else:
            make_track(w)
    print("done")
