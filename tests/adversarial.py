"""Adversarial poses for the guard bands: rays passing within 1e-8..1e-4 px of polyline vertices and cardinal
rays hitting a wall at 10 px +- 1e-9..1e-4 (see tests/test_host_replay.py)."""
import numpy as np

from oracle.c_oracle import COracleVecEnv, _lib as oracle_lib
from ppo_car_b200.track import load_track


def make_poses(path, seed=5, per_item=40):
    tr = load_track(path)
    ang = np.radians(tr.angle + 5.0 * np.arange(72))
    dirs = np.stack([np.cos(ang), np.sin(ang)], 1)
    rng = np.random.default_rng(seed)
    poses = []                                                # (px, py, heading index)
    for v in np.unique(tr.walls[:, :2], axis=0):
        for _ in range(per_item):
            k, i, u = rng.integers(0, 72), rng.integers(0, 12), rng.uniform(15, 400)
            d = dirs[(k + 6 * i) % 72]
            off = rng.choice([1e-8, -1e-8, 1e-6, -1e-6, 1e-4, -1e-4])
            poses.append((v[0] - u * d[0] - off * d[1], v[1] - u * d[1] + off * d[0], k))
    for w in tr.walls:
        a, b = w[:2], w[2:]
        for _ in range(per_item):
            k, i = rng.integers(0, 72), rng.choice([0, 3, 6, 9])
            d = dirs[(k + 6 * i) % 72]
            hit = a + rng.uniform(0.1, 0.9) * (b - a)
            dist = 10.0 + rng.choice([1e-9, -1e-9, 1e-6, -1e-6, 1e-4, -1e-4])
            poses.append((hit[0] - dist * d[0], hit[1] - dist * d[1], k))
    return np.array(poses), tr


def oracle_at_poses(path, poses, tr):
    """One no-op step of the float64 oracle from each pose: (terminated [n], pre-reset observation [n,18])."""
    n = len(poses)
    ora = COracleVecEnv(n, path, threads=1, scan_all_gates=False)
    ora.reset()
    stride = oracle_lib().oracle_env_bytes()
    blob = ora.state.view(np.uint8).reshape(n, stride)
    f64 = blob[:, :56].copy().view(np.float64)                # px py vx vy ax ay rot
    f64[:, 0], f64[:, 1] = poses[:, 0], poses[:, 1]
    f64[:, 6] = tr.angle + 5.0 * poses[:, 2]
    blob[:, :56] = f64.view(np.uint8)
    ref = ora.rollout(np.full((1, n), 8, np.uint8), want=("fobs", "term"))
    return ref["term"][0], ref["fobs"][0]
