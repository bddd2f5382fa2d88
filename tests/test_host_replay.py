"""CPU-only: replay the kernel's arithmetic (float32 fast path + guard bands + float64 fallbacks,
compiled for the host from the very same header the GPU kernel uses) against the golden
trajectories of the reference and against the float64 C oracle.  This is what lets the integer
traces be checked at scale without a GPU; the -m gpu tests repeat it on the device."""
import os

import numpy as np
import pytest

from oracle.c_oracle import COracleVecEnv
from tests.emul_util import emul_rollout
from tests.parity import RTOL_SYNTHETIC, assert_floats_close, assert_trajectory_matches

GROUPS = ["const", "lap", "random", "fwd"]


@pytest.mark.parametrize("name", ["track", "big_track"])
def test_kernel_arithmetic_replays_reference_golden(golden_dir, tracks_dir, name):
    g = np.load(os.path.join(golden_dir, f"carenv_{name}.npz"))
    for group in GROUPS:
        e = emul_rollout(os.path.join(tracks_dir, name + ".json"), g[f"{group}_actions"])
        assert_floats_close(e["reset_obs"], g["reset_obs"], "reset obs")
        done = (g[f"{group}_term"] | g[f"{group}_trunc"]).astype(bool)
        ref = dict(obs=np.where(done[..., None], g["reset_obs"], g[f"{group}_final_obs"]), rew=g[f"{group}_rew"],
                   term=g[f"{group}_term"], trunc=g[f"{group}_trunc"], gates_passed=g[f"{group}_gates_passed"],
                   time_passed=g[f"{group}_time_passed"], next_gate_index=g[f"{group}_next_gate_index"])
        got = dict(obs=e["obs"], rew=e["rew"], term=e["term"], trunc=e["trunc"], gates_passed=e["info"][..., 0],
                   time_passed=e["info"][..., 1], next_gate_index=e["info"][..., 2])
        assert_trajectory_matches(got, ref, what=f"{name}/{group}")
        if group == "lap":
            assert (e["info"][..., 3] >> 1).sum() == 1          # exactly one lap (+10) event


@pytest.mark.parametrize("name,n_envs,biased", [("big_track", 2048, False), ("track", 2048, True)])
def test_kernel_arithmetic_matches_oracle_at_scale(tracks_dir, name, n_envs, biased):
    T = 512
    rng = np.random.default_rng(7)
    p = [.3, .02, .1, .1, .2, .2, .02, .02, .04] if biased else None
    acts = rng.choice(9, size=(T, n_envs), p=p).astype(np.uint8)
    path = os.path.join(tracks_dir, name + ".json")
    ora = COracleVecEnv(n_envs, path, scan_all_gates=False)
    ora.reset()
    ref = ora.rollout(acts, want=("obs", "rew", "term", "trunc", "gates_passed", "time_passed", "next_gate_index"))
    e = emul_rollout(path, acts)
    got = dict(obs=e["obs"], rew=e["rew"], term=e["term"], trunc=e["trunc"], gates_passed=e["info"][..., 0],
               time_passed=e["info"][..., 1], next_gate_index=e["info"][..., 2])
    assert_trajectory_matches(got, ref, what=name)
    assert ref["term"].sum() > 1000
    # the float64 fallback must stay rare (it is a correctness device, not the main path)
    assert e["stats"].sum() < 2e-3 * acts.size


def test_reward_scaling_is_float32_of_float64_product(tracks_dir):
    rng = np.random.default_rng(3)
    acts = rng.choice(9, size=(300, 64), p=[.3, .02, .1, .1, .2, .2, .02, .02, .04]).astype(np.uint8)
    path = os.path.join(tracks_dir, "big_track.json")
    a, b = emul_rollout(path, acts, reward_scale=1.0), emul_rollout(path, acts, reward_scale=0.1)
    # TransformReward multiplies the float64 reward (train.py:65); Buffer stores float32 (lib/buffer.py:14)
    r64 = np.select([np.isclose(a["rew"], v) for v in (0.01, 1.0, 1.01, -3.0, -2.99, -2.0, -1.99)],
                    [0.01, 1.0, 0.01 + 1.0, -3.0, 0.01 - 3.0, 1.0 - 3.0, (0.01 + 1.0) - 3.0], 0.0)
    assert np.array_equal(b["rew"], (r64 * 0.1).astype(np.float32))


def test_packed_and_scalar_wall_paths_are_bit_identical(tracks_dir):
    """The FFMA2-packed pair path (U = 2, 4, 6), its table variant (denominators read from the per-track table
    instead of recomputed) and the scalar segment loop (U = 1) perform the same IEEE operations per component:
    observations, rewards, flags and final state must be identical bits."""
    rng = np.random.default_rng(17)
    acts = rng.choice(9, size=(400, 256), p=[.3, .02, .1, .1, .2, .2, .02, .02, .04]).astype(np.uint8)
    for name in ("track", "big_track"):
        path = os.path.join(tracks_dir, name + ".json")
        a, b = emul_rollout(path, acts, unrolled=True), emul_rollout(path, acts, unrolled=False)
        c = emul_rollout(path, acts, unrolled=2)              # denominators from the per-track table (k_rollout_tab)
        for k in ("obs", "rew", "term", "trunc", "info", "state_pv", "state_i"):
            assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k]), (name, k)
        assert np.array_equal(a["stats"], b["stats"]) and np.array_equal(a["stats"], c["stats"])


@pytest.mark.parametrize("n_outer,n_inner", [(7, 5), (10, 6), (16, 12), (31, 29), (90, 80)])
def test_other_segment_counts_match_oracle(tmp_path, n_outer, n_inner):
    """Tracks with other polyline sizes pick other loop unrollings (1, 2, 4): all must follow the oracle."""
    from tests.synth_tracks import ring_track

    path = ring_track(str(tmp_path / "ring.json"), n_outer, n_inner)
    rng = np.random.default_rng(n_outer)
    acts = rng.choice(9, size=(400, 512), p=[.3, .02, .1, .1, .2, .2, .02, .02, .04]).astype(np.uint8)
    ora = COracleVecEnv(512, path, scan_all_gates=True)
    ora.reset()
    ref = ora.rollout(acts, want=("obs", "rew", "term", "trunc", "gates_passed", "time_passed", "next_gate_index"))
    e = emul_rollout(path, acts)
    got = dict(obs=e["obs"], rew=e["rew"], term=e["term"], trunc=e["trunc"], gates_passed=e["info"][..., 0],
               time_passed=e["info"][..., 1], next_gate_index=e["info"][..., 2])
    assert_trajectory_matches(got, ref, what=f"ring {n_outer}+{n_inner}", rtol=RTOL_SYNTHETIC)
    assert ref["term"].sum() > 50 and ref["gates_passed"].max() > 0


def test_adversarial_poses_rays_through_vertices_and_at_the_collision_distance(tracks_dir):
    """Poses constructed so that a ray passes within 1e-8 .. 1e-4 px of a polyline vertex, or a cardinal ray hits
    a wall at 10 px +- 1e-9 .. 1e-4: far inside the float32 error (the guard bands are 1.5e-3 px and 2e-3
    relative) but far outside float64 noise, so the oracle's answer is well defined and the float32 answer is
    not.  The guard bands must send these to the float64 evaluation and every flag / distance must follow the
    oracle.  (Offsets of exactly 0 are not used: there the float64 answer itself depends on the last bit of
    cos/sin, which is the one thing oracle and kernel do not share.)"""
    from tests.adversarial import make_poses, oracle_at_poses

    path = os.path.join(tracks_dir, "big_track.json")
    poses, tr = make_poses(path)
    n = len(poses)
    term_ref, fobs_ref = oracle_at_poses(path, poses, tr)
    ref = {"term": [term_ref], "fobs": [fobs_ref]}
    acts = np.full((1, n), 8, np.uint8)
    pv = np.zeros((n, 4)); pv[:, 0], pv[:, 1] = poses[:, 0], poses[:, 1]
    si = np.zeros((n, 4), np.int32); si[:, 0] = poses[:, 2].astype(np.int32)
    e = emul_rollout(path, acts, state_pv=pv, state_i=si)
    # the emulation's first step already evaluates the poked pose; compare with the oracle's first step
    assert np.array_equal(e["term"][0], ref["term"][0])
    alive = ref["term"][0] == 0
    reset_obs = e["reset_obs"]
    got = np.where(alive[:, None], e["obs"][0], 0.0)
    want = np.where(alive[:, None], ref["fobs"][0], 0.0)
    assert_floats_close(got[:, 6:], want[:, 6:], "ray distances at adversarial poses")
    assert np.array_equal(e["obs"][0][~alive], np.broadcast_to(reset_obs, (int((~alive).sum()), 18)))
    assert e["stats"][0] + e["stats"][1] > 200                # the float64 path really was exercised
    assert 0 < alive.sum() < n


@pytest.mark.parametrize("seed", range(6))
def test_random_ring_tracks_match_oracle(tmp_path, seed):
    """Randomly shaped tracks (segment counts 5..40 per border, wobble, phase): whatever unroll factor and guard
    bands the host picks for them, the kernel arithmetic follows the float64 oracle."""
    from tests.synth_tracks import ring_track

    rng = np.random.default_rng(1000 + seed)
    n_outer, n_inner = int(rng.integers(5, 41)), int(rng.integers(5, 41))
    path = ring_track(str(tmp_path / "ring.json"), n_outer, n_inner, n_gates=int(rng.integers(4, 30)),
                      wobble=float(rng.uniform(0.0, 0.12)), seed_phase=float(rng.uniform(0.0, 1.5)))
    p = rng.dirichlet(np.ones(9) * 2.0)
    acts = rng.choice(9, size=(500, 384), p=p).astype(np.uint8)
    ora = COracleVecEnv(384, path, scan_all_gates=True)
    ora.reset()
    ref = ora.rollout(acts, want=("obs", "rew", "term", "trunc", "gates_passed", "time_passed", "next_gate_index"))
    e = emul_rollout(path, acts)
    got = dict(obs=e["obs"], rew=e["rew"], term=e["term"], trunc=e["trunc"], gates_passed=e["info"][..., 0],
               time_passed=e["info"][..., 1], next_gate_index=e["info"][..., 2])
    assert_trajectory_matches(got, ref, what=f"ring {n_outer}+{n_inner} (seed {seed})", rtol=RTOL_SYNTHETIC)


def test_start_pose_inside_the_collision_distance(tmp_path):
    """`start_destroyed` (lib/car_env.py:682-686, 745-748): the start pose is 6 px from a wall, so Car.update inside
    reset already sets `destroyed`; every step then terminates with -3 (+0.01 when thrusting forward) and the
    autoreset puts the car back on the same spot.  Kernel arithmetic vs the port (= the reference's behaviour,
    see test_port_against_live_reference_when_available) and vs the C oracle."""
    from oracle.carenv_port import PortVecEnv
    from tests.synth_tracks import near_wall_track

    path = near_wall_track(str(tmp_path / "near_wall.json"))
    rng = np.random.default_rng(12)
    T, N = 30, 48
    acts = rng.integers(0, 9, size=(T, N)).astype(np.uint8)
    e = emul_rollout(path, acts)
    assert e["term"].all() and not e["trunc"].any() and (e["info"][..., 1] == 1).all()
    fwd = np.isin(acts, (0, 4, 5))
    assert np.array_equal(e["rew"], np.where(fwd, np.float32(0.01 - 3.0), np.float32(-3.0)))
    assert np.array_equal(e["obs"], np.broadcast_to(e["reset_obs"], e["obs"].shape))
    ora = COracleVecEnv(N, path, scan_all_gates=True)
    obs0 = ora.reset()
    ref = ora.rollout(acts, want=("obs", "rew", "term", "trunc", "gates_passed", "time_passed", "next_gate_index"))
    assert_floats_close(np.broadcast_to(e["reset_obs"], obs0.shape), obs0, "reset obs")
    got = dict(obs=e["obs"], rew=e["rew"], term=e["term"], trunc=e["trunc"], gates_passed=e["info"][..., 0],
               time_passed=e["info"][..., 1], next_gate_index=e["info"][..., 2])
    assert_trajectory_matches(got, ref, what="near-wall start")
    port = PortVecEnv(4, path)
    assert_floats_close(np.broadcast_to(e["reset_obs"], (4, 18)), port.reset(), "reset obs vs port")
    for t in range(8):
        o, r, te, tr, info = port.step(acts[t, :4])
        assert te.all() and np.array_equal(r.astype(np.float32), e["rew"][t, :4])
        assert np.array_equal(info["time_passed"], e["info"][t, :4, 1])


def test_config1_shape_track_json_24_envs_1024_uniform_steps(tracks_dir):
    """BASELINE configs[0] exactly: tracks/track.json, 24 envs x 1024 steps, i.i.d. uniform actions."""
    path = os.path.join(tracks_dir, "track.json")
    acts = np.random.default_rng(101).integers(0, 9, size=(1024, 24)).astype(np.uint8)
    ora = COracleVecEnv(24, path, scan_all_gates=True)
    ora.reset()
    ref = ora.rollout(acts, want=("obs", "rew", "term", "trunc", "gates_passed", "time_passed", "next_gate_index"))
    e = emul_rollout(path, acts)
    got = dict(obs=e["obs"], rew=e["rew"], term=e["term"], trunc=e["trunc"], gates_passed=e["info"][..., 0],
               time_passed=e["info"][..., 1], next_gate_index=e["info"][..., 2])
    assert_trajectory_matches(got, ref, what="config 1")
    assert ref["term"].sum() > 50


@pytest.mark.parametrize("name", ["track", "big_track"])
def test_next_gate_only_test_equals_the_full_ordered_gate_scan(tracks_dir, name):
    """The kernel tests gate `next_gate_index` only; the reference scans every active gate in index order
    (lib/car_env.py:394-408).  Oracle in its literal full-scan mode, forward-biased actions (many gate events)."""
    path = os.path.join(tracks_dir, name + ".json")
    n, T = 1024, 512
    acts = np.random.default_rng(77).choice(9, size=(T, n), p=[.3, .02, .1, .1, .2, .2, .02, .02, .04]).astype(np.uint8)
    ora = COracleVecEnv(n, path, scan_all_gates=True)
    ora.reset()
    ref = ora.rollout(acts, want=("obs", "rew", "term", "trunc", "gates_passed", "time_passed", "next_gate_index"))
    e = emul_rollout(path, acts)
    got = dict(obs=e["obs"], rew=e["rew"], term=e["term"], trunc=e["trunc"], gates_passed=e["info"][..., 0],
               time_passed=e["info"][..., 1], next_gate_index=e["info"][..., 2])
    assert_trajectory_matches(got, ref, what=f"{name}, full gate scan")
    assert (e["info"][..., 3] & 1).sum() > 2000                # gate events


def test_heading_table_equals_numpy_trigonometry(tracks_dir):
    """The kernels take cos / sin of the heading from a 72-entry table built with libm; the reference evaluates
    np.cos(np.radians(rotation)) (lib/car_env.py:430, 155-160) and the C oracle shares libm with the table — so this is
    the one place where kernel and oracle could jointly differ from the reference.  Checked directly: for the initial
    angles of the shipped tracks (and a few odd ones) the table is bit-identical to numpy on the wrapped angle, and
    within 4e-15 of numpy on the UNWRAPPED angles the reference reaches after turning (rotation is never wrapped
    there: the documented deviation)."""
    import ctypes as C
    import json

    from tests.emul_util import lib

    L = lib()
    angles = {0.0, 90.0, 180.0, 270.0, 33.0, 12.5}
    for name in ("big_track.json", "track.json"):
        with open(os.path.join(tracks_dir, name)) as fh:
            angles.add(float(json.load(fh)["initial_angle"]))
    for ang in sorted(angles):
        tab = np.zeros((72, 2))
        assert L.emul_heading_table(C.c_double(ang), tab.ctypes.data_as(C.c_void_p)) == 0
        k = np.arange(72)
        rad = np.radians(ang + 5.0 * k)
        assert np.array_equal(tab[:, 0], np.cos(rad)) and np.array_equal(tab[:, 1], np.sin(rad)), ang
        # unwrapped: two full turns either way, accumulated by +-5.0 in float64 exactly as Car.move_car does
        for turns in (-2, -1, 1, 2):
            rot = ang + 5.0 * (k + 72 * turns)
            for col, fn in ((0, np.cos), (1, np.sin)):
                ref = fn(np.radians(rot))
                err = np.abs(tab[:, col] - ref)
                # the argument itself is only known to half a unit in the last place of ~13 rad: a few 1e-15
                assert err.max() <= 4.0e-15, (ang, turns, err.max())


def test_time_slice_work_lists_cover_every_step_once_and_in_order():
    """slice_items (carenv_core.cuh, what every warp of k_rollout_tab_sliced computes for itself): over all warps the
    items cover every (job, step) exactly once; a job cut between two warps is split into first steps / last steps
    with the first part scheduled before the second starts (the property the in-kernel flag only double-checks)."""
    import ctypes as C

    from tests.emul_util import lib

    L = lib()
    rng = np.random.default_rng(5)
    cases = [(28, 256, 16), (27, 256, 16), (16, 4, 16), (222, 256, 16), (221, 1024, 16), (17, 5, 16), (16, 7, 16), (40, 1, 16)]
    cases += [(int(rng.integers(16, 400)), int(rng.integers(1, 600)), 16) for _ in range(60)]
    out = (C.c_int * 4)()
    for n_jobs, n_steps, n_slots in cases:
        total = n_jobs * n_steps
        quota = -(-total // n_slots)
        seen = np.zeros((n_jobs, n_steps), np.int32)
        head_end, tail_start = {}, {}
        for slot in range(n_slots):
            L.emul_slice_items(n_jobs, n_steps, slot, n_slots, out)
            first_job, head_len, last_full, tail_len = list(out)
            clock = 0
            job = first_job
            if head_len > 0:
                seen[job, :head_len] += 1
                clock += head_len
                head_end[job] = clock
                job += 1
            while job <= last_full:
                seen[job] += 1
                clock += n_steps
                job += 1
            if tail_len > 0:
                tail_start[job] = clock
                seen[job, n_steps - tail_len:] += 1
                clock += tail_len
            assert clock <= quota
        assert np.all(seen == 1), (n_jobs, n_steps)
        assert set(head_end) == set(tail_start)
        for job in head_end:
            assert head_end[job] <= tail_start[job], (n_jobs, n_steps, job)
