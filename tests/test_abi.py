"""CPU-only: the C-ABI library builds for sm_100a, loads, exports every symbol the header declares,
and fails loudly (no CPU fallback) when there is no GPU.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import ppo_car_b200
from ppo_car_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "carenv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|const char \*)\s*\*?\s*(carenv_\w+|gae_\w+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = ppo_car_b200.build()
    assert os.path.exists(path)
    L = C.CDLL(path)
    names = declared_functions()
    assert {"carenv_create", "carenv_destroy", "carenv_reset", "carenv_step", "carenv_rollout", "carenv_reset_obs",
            "carenv_stats", "gae_reverse_scan", "carenv_last_error", "carenv_abi_version"} <= set(names)
    for n in names:
        assert hasattr(L, n), n
    assert _lib.lib().carenv_abi_version() == 1


def test_ctypes_bindings_have_the_declared_argument_counts():
    """Every prototype of include/carenv_b200.h that ppo_car_b200/_lib.py binds with argtypes has as many
    parameters in the header as in the binding (guards against the two drifting apart)."""
    text = open(os.path.join(ROOT, "include", "carenv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = dict(re.findall(r"\b(?:int|const char \*)\s*\*?\s*(carenv_\w+|gae_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S))
    L = _lib.lib()
    checked = 0
    for name, params in protos.items():
        fn = getattr(L, name)
        if fn.argtypes is None:
            continue
        params = params.strip()
        n_decl = 0 if params in ("", "void") else params.count(",") + 1
        assert len(fn.argtypes) == n_decl, (name, len(fn.argtypes), n_decl)
        checked += 1
    assert checked >= 15


def test_library_contains_sm100a_code():
    out = os.popen(f"cuobjdump -lelf {ppo_car_b200.build()} 2>/dev/null").read()
    assert "sm_100a" in out


def test_bad_arguments_return_error_codes():
    L = _lib.lib()
    h = C.c_void_p()
    assert L.carenv_create(None, 4, None, 1, 0.0, 0.0, 0.0, 0, C.byref(h)) == -1          # CARENV_E_INVAL
    walls = np.zeros((3000, 4))
    gates = np.zeros((1, 4))
    rc = L.carenv_create(walls.ctypes.data_as(C.c_void_p), 3000, gates.ctypes.data_as(C.c_void_p), 1, 0.0, 0.0, 0.0,
                         0, C.byref(h))
    assert rc == -2 and b"segments" in L.carenv_last_error()                               # CARENV_E_TRACK
    assert L.carenv_step(None, 1, None, None, None, None, 0, 1.0, None, None, None, None, 0, None, None) < 0
    assert L.gae_reverse_scan(*([None] * 9), 4, 4, 0.99, 0.95, None) == -1
    assert L.carenv_rollout_poses(None, 1, 1, None, None, None, None, 0, 1.0, None, None, None, None, 0, None, None) < 0
    assert L.carenv_observe(None, 1, None, None, None, None) == -1
    assert L.carenv_step_host(None, 1, None, None, None, None, 0, 1.0, None, None, None, None, 0, None, None) == -1
    assert L.carenv_step_host_records(None, 1, None, None, None, None, 0, 1.0, None, None, None, None) == -1
    assert L.carenv_step_records(None, 1, None, None, None, None, 0, 1.0, None, None, None, None) == -1
    assert L.carenv_host_alloc(16, None) == -1 and L.carenv_host_free(None) == 0
    # the policy / PPO entry points
    comm, hbuf = C.c_void_p(), C.create_string_buffer(64)
    assert L.carenv_ppo_comm_create(0, 0, C.byref(comm), hbuf) == -1 and not comm.value      # world out of range
    assert L.carenv_ppo_comm_create(9, 0, C.byref(comm), hbuf) == -1
    assert L.carenv_ppo_comm_create(2, 2, C.byref(comm), hbuf) == -1                         # rank out of range
    assert L.carenv_ppo_comm_create(2, 0, None, hbuf) == -1
    assert L.carenv_ppo_comm_connect(None, hbuf) == -1 and L.carenv_ppo_comm_destroy(None) == 0
    assert L.carenv_ppo_epoch_workspace_floats() > 160 * 12298
    epoch_args = [None] * 14 + [512, 80, 0.2, 0.5, 0.001, None, None, None, None, 0.9, 0.999, 1e-5, 1.0, None, None, None,
                                None, 0, None, None]
    assert L.carenv_ppo_epoch(*epoch_args) == -1 and b"null" in L.carenv_last_error()
    epoch_args[14] = 5000                                                                    # batch out of range
    assert L.carenv_ppo_epoch(*epoch_args) == -1 and b"batch" in L.carenv_last_error()
    assert L.carenv_policy_rollout_warp(*([None] * 9), 4, 4, 0, 0, 0, *([None] * 6), 1.0, *([None] * 10)) == -1
    assert L.carenv_policy_rollout_tc(None, None, 4, 4, 0, 0, 0, *([None] * 6), 1.0, *([None] * 10)) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(ppo_car_b200.CarEnvError):
        ppo_car_b200.VecCarEnv(4, ppo_car_b200.builtin_track("track"))
    with pytest.raises(ppo_car_b200.CarEnvError):
        ppo_car_b200.Buffer((18,), 8, 4, "cpu")
    with pytest.raises(ppo_car_b200.CarEnvError):
        z = torch.zeros(4, 4)
        ppo_car_b200.gae_reverse_scan(z, z, z, z, z[0], z[0], z[0])
    L = _lib.lib()
    tr = ppo_car_b200.load_track(ppo_car_b200.builtin_track("track"))
    h = C.c_void_p()
    rc = L.carenv_create(tr.walls.ctypes.data_as(C.c_void_p), len(tr.walls), tr.gates.ctypes.data_as(C.c_void_p),
                         len(tr.gates), tr.start[0], tr.start[1], tr.angle, 0, C.byref(h))
    assert rc == -3 and not h.value                                                        # CARENV_E_NOGPU


def test_track_loader_matches_reference_construction(tracks_dir):
    from oracle.carenv_port import load_track as port_load

    for name in ("track", "big_track"):
        p = os.path.join(tracks_dir, name + ".json")
        a, b = ppo_car_b200.load_track(p), port_load(p)
        assert np.array_equal(a.walls, np.asarray(b["walls"])) and np.array_equal(a.gates, np.asarray(b["gates"]))
        assert a.start == tuple(b["start"]) and a.angle == b["angle"]
    with pytest.raises(FileNotFoundError):
        ppo_car_b200.load_track(os.path.join(tracks_dir, "nope.json"))


def test_track_validation(tmp_path, tracks_dir):
    import json

    from ppo_car_b200.track import validate_track

    for name in ("track", "big_track"):
        assert validate_track(os.path.join(tracks_dir, name + ".json")) == []
    raw = json.load(open(os.path.join(tracks_dir, "track.json")))

    def variant(**changes):
        d = dict(raw)
        d.update(changes)
        p = tmp_path / "v.json"
        p.write_text(json.dumps(d))
        return validate_track(str(p))

    assert any("not closed" in m for m in variant(outer_track_points=raw["outer_track_points"][:-1]))
    assert any("odd number" in m for m in variant(reward_gates=raw["reward_gates"][:-1]))
    assert any("outside the outer" in m for m in variant(initial_position=[0.001, 0.001]))
    assert any("inside the inner" in m for m in variant(initial_position=[0.5, 0.5]))
    wall_y = raw["outer_track_points"][0][1]
    assert any("closer than 10 px" in m or "outside" in m
               for m in variant(initial_position=[raw["initial_position"][0], wall_y - 0.005]))
