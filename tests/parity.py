"""Shared parity assertions: integer outputs bit-exact, float outputs within the tolerance
BASELINE.json's north_star states — 1e-5 RELATIVE, with no absolute floor — on every observation
column except the two velocity columns (obs[2], obs[3] = v/10), which pass through exactly 0 and
therefore keep a 1e-6 absolute floor.  Position (obs[0:2]), heading cos/sin
(obs[4:6]) and the twelve ray distances (obs[6:18], the columns the north star names) are pure
rtol: a regression of the ray casting near walls (small d/1000) cannot hide behind a floor."""
import numpy as np

RTOL = 1e-5
# Synthetic ring tracks (tests/synth_tracks.py): the same 1e-5 as the reference's tracks.  (Round 2 first needed 2e-5
# here: with float32 denominators a ray grazing a wall at < 0.4 degrees lost 6e-8 / sin(incidence); the denominators
# are now rounded once from float64 in every kernel — see seg_den in csrc/carenv_core.cuh.)
RTOL_SYNTHETIC = RTOL
ATOL_VELOCITY = 1e-6
VELOCITY_COLUMNS = (2, 3)


def _column_atol(shape):
    """Per-element absolute floor for an array whose last axis is the 18 observation columns."""
    atol = np.zeros(shape[-1], np.float64)
    if shape[-1] == 18:
        atol[list(VELOCITY_COLUMNS)] = ATOL_VELOCITY
    return atol


def assert_floats_close(got, want, what="", atol=None, rtol=RTOL):
    """|got - want| <= 1e-5 * max(|got|, |want|) (+ `atol`, default: 1e-6 on obs[2:4] only when the last
    axis has 18 columns, else 0)."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, f"{what}: shapes {got.shape} vs {want.shape}"
    if atol is None:
        atol = _column_atol(got.shape) if got.ndim else 0.0
    err = np.abs(got - want)
    tol = rtol * np.maximum(np.abs(got), np.abs(want)) + atol
    bad = err > tol
    if bad.any():
        rel = err / np.maximum(np.maximum(np.abs(got), np.abs(want)), 1e-300)
        raise AssertionError(f"{what}: {int(bad.sum())} of {bad.size} elements outside {rtol:g} relative "
                             f"(velocity columns: + 1e-6 absolute); worst relative error {rel[bad].max():.3e}, "
                             f"worst absolute error {err[bad].max():.3e}")


def assert_ray_columns_close(got, want, what=""):
    """Ray-distance columns only (obs[..., 6:18] or a [..., 12] array): pure 1e-5 relative."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    if got.shape[-1] == 18:
        got, want = got[..., 6:], want[..., 6:]
    assert_floats_close(got, want, what, atol=0.0)


def assert_trajectory_matches(got, ref, reset_obs=None, what="", rtol=RTOL):
    """got / ref: dicts with obs [T,N,18], rew, term, trunc, gates_passed, time_passed, next_gate_index."""
    for k in ("term", "trunc", "gates_passed", "time_passed", "next_gate_index"):
        a, b = np.asarray(got[k]).astype(np.int64), np.asarray(ref[k]).astype(np.int64)
        bad = a != b
        assert not bad.any(), f"{what}: integer output {k} differs in {int(bad.sum())} elements ({int(bad.any(0).sum())} envs)"
    # float32(reward_f64 * scale) is an exactly specified value: bit-exact
    assert np.array_equal(np.asarray(got["rew"], np.float32), np.asarray(ref["rew"]).astype(np.float32)), f"{what}: rewards"
    assert_floats_close(got["obs"], ref["obs"], what + " obs", rtol=rtol)
