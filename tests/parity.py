"""Shared parity assertions: integer outputs bit-exact, float outputs within the tolerance
BASELINE.json's north_star states (1e-5 relative), with a 1e-6 absolute floor on the
normalised observation because velocities and sin/cos pass through zero."""
import numpy as np

RTOL, ATOL = 1e-5, 1e-6


def assert_floats_close(got, want, what=""):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    err = np.abs(got - want)
    tol = RTOL * np.maximum(np.abs(got), np.abs(want)) + ATOL
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.size} elements outside 1e-5 rel + 1e-6 abs; max err {err.max():.3e}"


def assert_trajectory_matches(got, ref, reset_obs=None, what=""):
    """got / ref: dicts with obs [T,N,18], rew, term, trunc, gates_passed, time_passed, next_gate_index."""
    for k in ("term", "trunc", "gates_passed", "time_passed", "next_gate_index"):
        a, b = np.asarray(got[k]).astype(np.int64), np.asarray(ref[k]).astype(np.int64)
        bad = a != b
        assert not bad.any(), f"{what}: integer output {k} differs in {int(bad.sum())} elements ({int(bad.any(0).sum())} envs)"
    # float32(reward_f64 * scale) is an exactly specified value: bit-exact
    assert np.array_equal(np.asarray(got["rew"], np.float32), np.asarray(ref["rew"]).astype(np.float32)), f"{what}: rewards"
    assert_floats_close(got["obs"], ref["obs"], what + " obs")
