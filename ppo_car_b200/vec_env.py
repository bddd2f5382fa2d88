"""VecCarEnv — the batched, auto-resetting CarEnv on one B200.

Drop-in for what train.py does with ``gym.vector.AsyncVectorEnv([make_env]*n_envs)``
(/root/reference/train.py:138-142, 159, 185, 296): ``single_observation_space``,
``single_action_space``, ``reset(options={"track_path": ...})``, ``step(actions)`` returning
the 5-tuple with same-step autoreset, ``close()``.  ``reward_scaling`` stands in for the
``TransformReward`` wrapper (train.py:65, 68).

All state and outputs are PyTorch CUDA tensors; the arithmetic runs in the hand-written
sm_100a kernels of libcarenv_b200.so.  There is no CPU path: constructing a VecCarEnv
without the built library or without a CUDA device raises.

Two calling conventions for ``step``:
  * CUDA tensor in  -> CUDA tensors out (no host round trip; int64/int32/uint8 actions; float32 rewards).
  * numpy array in  -> numpy arrays out (the reference's convention, train.py:185) with the reference's dtypes:
    float32 observations, FLOAT64 rewards, bool flags, int info.  Actions are staged through pinned host memory,
    results come back as an observation array plus one 16-byte record per environment (carenv_step_host_records);
    flags and info are strided numpy views of the record array.

BUFFER LIFETIME (differs from the reference, which returns fresh arrays): everything ``reset`` / ``step`` return
is a view of buffers owned by this object and is OVERWRITTEN BY THE NEXT ``step`` / ``reset``.  Code that keeps a
result across a step (``train.py:176-195`` stores the previous ``next_obs`` / ``next_terminateds`` after calling
``envs.step``) must ``.clone()`` / ``.copy()`` it first, or construct the env with ``copy_outputs=True``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .track import Track, builtin_track, load_track

OBS_DIM = 18
N_ACTIONS = 9


class Box:
    """Minimal stand-in for gymnasium.spaces.Box (lib/car_env.py:513-524)."""

    def __init__(self, low, high, dtype=np.float32):
        self.low, self.high, self.dtype = np.asarray(low, dtype), np.asarray(high, dtype), dtype
        self.shape = self.low.shape


class Discrete:
    """Minimal stand-in for gymnasium.spaces.Discrete (lib/car_env.py:525)."""

    def __init__(self, n: int):
        self.n = n
        self.shape = ()


_ACT_CODES = {torch.uint8: _lib.ACT_U8, torch.int32: _lib.ACT_I32, torch.int64: _lib.ACT_I64}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class VecCarEnv:
    def __init__(self, n_envs: int, track_path: str | None = None, device="cuda", reward_scaling: float = 1.0,
                 float_flags: bool = False, with_info: bool = True, _out: dict | None = None,
                 copy_outputs: bool = False, debug_info: bool = False, final_observation: bool = False):
        """``copy_outputs=True``: step / reset return fresh copies (the reference's semantics) instead of views of the
        internal buffers.  ``debug_info=True``: the numpy step also returns ``next_gate_index`` and ``events`` (16 more
        bytes per environment over PCIe; the parity tests use them).  ``final_observation=True``: step() with CUDA actions
        also returns ``info["final_observation"]`` [N, 18] — the observation each step ended in BEFORE the autoreset — and
        the mask ``info["_final_observation"]`` (gymnasium's vector-env convention)."""
        if n_envs < 1:
            raise ValueError("n_envs must be >= 1")
        self._L = _lib.lib()                       # raises if the CUDA extension is missing
        if not torch.cuda.is_available():
            raise _lib.CarEnvError("VecCarEnv needs a CUDA device (there is no CPU implementation)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CarEnvError("VecCarEnv only runs on CUDA devices")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs = int(n_envs)
        self.reward_scaling = float(reward_scaling)
        self.float_flags = bool(float_flags)
        self.with_info = bool(with_info)
        self.copy_outputs = bool(copy_outputs)
        self.debug_info = bool(debug_info)
        self.final_observation = bool(final_observation)
        self._fobs = None
        low = np.array([0, 0, -1, -1, -1, -1] + [0] * 12, np.float32)
        high = np.ones(OBS_DIM, np.float32)
        self.single_observation_space = Box(low, high)
        self.single_action_space = Discrete(N_ACTIONS)
        self.observation_space, self.action_space = self.single_observation_space, self.single_action_space
        self._handle = None
        self.track: Track | None = None
        n, dev = self.num_envs, self.device
        # structure-of-arrays state, owned by PyTorch (include/carenv_b200.h)
        self.pos = torch.zeros((n, 2), dtype=torch.float64, device=dev)
        self.vel = torch.zeros((n, 2), dtype=torch.float64, device=dev)
        self.ints = torch.zeros((n, 4), dtype=torch.int32, device=dev)
        fdt = torch.float32 if self.float_flags else torch.uint8
        if _out is not None:                      # output rows owned by a MultiTrackVecEnv (contiguous slices)
            self._obs, self._rew, self._term, self._trunc = _out["obs"], _out["rew"], _out["term"], _out["trunc"]
            self._info = _out.get("info") if self.with_info else None
        else:
            self._obs = torch.empty((n, OBS_DIM), dtype=torch.float32, device=dev)
            self._rew = torch.empty((n,), dtype=torch.float32, device=dev)
            self._term = torch.empty((n,), dtype=fdt, device=dev)
            self._trunc = torch.empty((n,), dtype=fdt, device=dev)
            self._info = torch.empty((n, 4), dtype=torch.int32, device=dev) if self.with_info else None
        self._host = None                          # pinned staging buffers, created on first numpy call
        self._needs_reset = True
        self._set_track(track_path or builtin_track("track"))   # reference default: tracks/track.json

    # ------------------------------------------------------------------ track / handle
    def _set_track(self, path: str) -> None:
        track = load_track(path)
        handle = C.c_void_p()
        walls = np.ascontiguousarray(track.walls, np.float64)
        gates = np.ascontiguousarray(track.gates, np.float64)
        rc = self._L.carenv_create(walls.ctypes.data_as(C.c_void_p), len(walls), gates.ctypes.data_as(C.c_void_p),
                                   len(gates), track.start[0], track.start[1], track.angle, self.device.index,
                                   C.byref(handle))
        _lib.check(rc, "carenv_create")
        if self._handle is not None:
            self._L.carenv_destroy(self._handle)
        self._handle, self.track = handle, track
        self._needs_reset = True

    @property
    def reset_observation(self) -> np.ndarray:
        out = np.zeros(OBS_DIM, np.float32)
        _lib.check(self._L.carenv_reset_obs(self._handle, out.ctypes.data_as(C.c_void_p)), "carenv_reset_obs")
        return out

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ gym-style API
    def reset(self, seed=None, options=None):
        """CarEnv.reset for every environment (lib/car_env.py:605-691).  ``seed`` is accepted and
        ignored exactly like in the reference: the environment has no randomness."""
        if options and "track_path" in options:
            self._set_track(options["track_path"])
        with torch.cuda.device(self.device):
            rc = self._L.carenv_reset(self._handle, self.num_envs, _ptr(self.pos), _ptr(self.vel), _ptr(self.ints),
                                      _ptr(self._obs), self._stream())
        _lib.check(rc, "carenv_reset")
        self._needs_reset = False
        zeros = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        obs = self._obs.clone() if self.copy_outputs else self._obs
        return obs, {"gates_passed": zeros, "time_passed": zeros.clone()}

    def _info_dict(self, info):
        if info is None:
            return {}
        # views only (no per-step arithmetic on the host): "events" has bit 0 = gate hit, bit 1 = lap (+10)
        return {"gates_passed": info[..., 0], "time_passed": info[..., 1], "next_gate_index": info[..., 2],
                "events": info[..., 3]}

    def step(self, actions):
        """One step of every environment with same-step autoreset (lib/car_env.py:693-760 +
        gymnasium AsyncVectorEnv).  Returns (obs, rewards, terminateds, truncateds, info).

        The returned tensors / arrays are views of internal buffers that the NEXT step() overwrites (see the
        module docstring); clone / copy what must survive a step, or use ``copy_outputs=True``."""
        if self._needs_reset:
            raise _lib.CarEnvError("call reset() before step()")
        if isinstance(actions, torch.Tensor) and actions.is_cuda:
            return self._step_device(actions)
        return self._step_host(np.asarray(actions))

    def _step_device(self, actions: torch.Tensor):
        if actions.dtype not in _ACT_CODES:
            actions = actions.to(torch.int64)
        actions = actions.reshape(-1)
        if actions.numel() != self.num_envs:
            raise ValueError(f"expected {self.num_envs} actions, got {actions.numel()}")
        if not actions.is_contiguous():
            actions = actions.contiguous()
        flag_code = _lib.FLAG_F32 if self.float_flags else _lib.FLAG_U8
        with torch.cuda.device(self.device):
            if self.final_observation:
                if self._fobs is None:
                    self._fobs = torch.empty((self.num_envs, OBS_DIM), dtype=torch.float32, device=self.device)
                rc = self._L.carenv_step_final(self._handle, self.num_envs, _ptr(self.pos), _ptr(self.vel),
                                               _ptr(self.ints), _ptr(actions), _ACT_CODES[actions.dtype],
                                               self.reward_scaling, _ptr(self._obs), _ptr(self._fobs), _ptr(self._rew),
                                               _ptr(self._term), _ptr(self._trunc), flag_code, _ptr(self._info),
                                               self._stream())
            else:
                rc = self._L.carenv_step(self._handle, self.num_envs, _ptr(self.pos), _ptr(self.vel), _ptr(self.ints),
                                         _ptr(actions), _ACT_CODES[actions.dtype], self.reward_scaling, _ptr(self._obs),
                                         _ptr(self._rew), _ptr(self._term), _ptr(self._trunc), flag_code,
                                         _ptr(self._info), self._stream())
        _lib.check(rc, "carenv_step")
        term = self._term if self.float_flags else self._term.view(torch.bool)
        trunc = self._trunc if self.float_flags else self._trunc.view(torch.bool)
        info = self._info_dict(self._info.clone() if (self.copy_outputs and self._info is not None) else self._info)
        if self.final_observation:
            info["final_observation"] = self._fobs.clone() if self.copy_outputs else self._fobs
            info["_final_observation"] = (self._term != 0) | (self._trunc != 0)
        if self.copy_outputs:
            return self._obs.clone(), self._rew.clone(), term.clone(), trunc.clone(), info
        return self._obs, self._rew, term, trunc, info

    _REC_DTYPE = np.dtype([("reward", "<f8"), ("gates_passed", "<i4"), ("time_passed", "<u2"),
                           ("terminated", "?"), ("truncated", "?")])              # carenv_step_record, 16 bytes

    def _step_host(self, actions: np.ndarray):
        """numpy in -> numpy out through carenv_step_host_records: the library narrows the actions into a pinned
        buffer and pipelines H2D copy, kernel and D2H copies over sub-ranges of the batch (the D2H copy is the
        bottleneck: 88 B per env over PCIe — the observation and one 16-byte record that already holds the float64
        reward, the bool flags and the info counters).  Rewards, flags and info are strided views of the pinned
        record array (float_flags=True returns float32 copies of the flags instead)."""
        n = self.num_envs
        if actions.size != n:
            raise ValueError(f"expected {n} actions, got {actions.size}")
        if self._host is None:
            pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
            self._host = dict(obs=pin((n, OBS_DIM), torch.float32), rec=pin((n, 16), torch.uint8),
                              dbg=pin((n, 4), torch.int32) if self.debug_info else None)
        h = self._host
        flat = np.ascontiguousarray(actions.reshape(-1))
        if flat.dtype == np.uint8:
            code = _lib.ACT_U8
        elif flat.dtype == np.int32:
            code = _lib.ACT_I32
        else:
            flat = flat.astype(np.int64, copy=False)
            code = _lib.ACT_I64
        with torch.cuda.device(self.device):
            rc = self._L.carenv_step_host_records(self._handle, n, _ptr(self.pos), _ptr(self.vel), _ptr(self.ints),
                                                  C.c_void_p(flat.ctypes.data), code, self.reward_scaling,
                                                  _ptr(h["obs"]), _ptr(h["rec"]), _ptr(h["dbg"]), self._stream())
        _lib.check(rc, "carenv_step_host_records")
        rec = h["rec"].numpy().view(self._REC_DTYPE).reshape(n)
        obs, rew, term, trunc = h["obs"].numpy(), rec["reward"], rec["terminated"], rec["truncated"]
        info = {}
        if self.with_info:
            info = {"gates_passed": rec["gates_passed"], "time_passed": rec["time_passed"]}
            if self.debug_info:
                dbg = h["dbg"].numpy()
                info["next_gate_index"], info["events"] = dbg[:, 2], dbg[:, 3]
        if self.float_flags:
            term, trunc = term.astype(np.float32), trunc.astype(np.float32)
        if self.copy_outputs:
            obs, rew, term, trunc = obs.copy(), rew.copy(), term.copy(), trunc.copy()
            info = {k: v.copy() for k, v in info.items()}
        return obs, rew, term, trunc, info

    # ------------------------------------------------------------------ multi-step launch
    def rollout(self, actions: torch.Tensor, obs_out=None, reward_out=None, term_out=None, trunc_out=None,
                info_out=None, store_obs: bool = True, store_info: bool = False, store_poses: bool = False,
                pose_out=None):
        """n_steps consecutive steps in one kernel launch (the rollout loop of train.py:173-195 with
        the actions given up front).  ``actions`` is a CUDA tensor [n_steps, n_envs]; outputs are
        [n_steps, n_envs(, 18)] CUDA tensors (allocated here unless passed in).

        ``store_poses=True`` (or ``pose_out``) stores a 32-byte pose record per env-step instead of the
        72-byte observation (out["poses"], an opaque float64 tensor [n_steps, n_envs, 4]); ``observe`` turns
        records back into observations, bit-identical to the ones this call would have stored."""
        if store_poses or pose_out is not None:
            store_obs, store_poses = False, True
            if obs_out is not None:
                raise ValueError("pass either obs_out or pose_out / store_poses")
        if self._needs_reset:
            raise _lib.CarEnvError("call reset() before rollout()")
        if not (isinstance(actions, torch.Tensor) and actions.is_cuda and actions.dim() == 2):
            raise ValueError("rollout expects a CUDA tensor of shape [n_steps, n_envs]")
        if actions.dtype not in _ACT_CODES:
            actions = actions.to(torch.int64)
        actions = actions.contiguous()
        K, n = actions.shape
        if n != self.num_envs:
            raise ValueError(f"expected {self.num_envs} environments, got {n}")
        dev = self.device
        fdt = torch.float32 if self.float_flags else torch.uint8
        if obs_out is None and store_obs:
            obs_out = torch.empty((K, n, OBS_DIM), dtype=torch.float32, device=dev)
        if reward_out is None:
            reward_out = torch.empty((K, n), dtype=torch.float32, device=dev)
        if term_out is None:
            term_out = torch.empty((K, n), dtype=fdt, device=dev)
        if trunc_out is None:
            trunc_out = torch.empty((K, n), dtype=fdt, device=dev)
        if info_out is None and store_info:
            info_out = torch.empty((K, n, 4), dtype=torch.int32, device=dev)
        if pose_out is None and store_poses:
            pose_out = torch.empty((K, n, 4), dtype=torch.float64, device=dev)
        for name, t, shape, dt in (("obs_out", obs_out, (K, n, OBS_DIM), torch.float32),
                                   ("pose_out", pose_out, (K, n, 4), torch.float64),
                                   ("reward_out", reward_out, (K, n), torch.float32),
                                   ("term_out", term_out, (K, n), fdt), ("trunc_out", trunc_out, (K, n), fdt),
                                   ("info_out", info_out, (K, n, 4), torch.int32)):
            if t is not None and (tuple(t.shape) != shape or t.dtype != dt or not t.is_contiguous() or not t.is_cuda):
                raise ValueError(f"{name} must be a contiguous CUDA tensor of shape {shape} and dtype {dt}")
        with torch.cuda.device(dev):
            fn = self._L.carenv_rollout_poses if store_poses else self._L.carenv_rollout
            rc = fn(self._handle, n, K, _ptr(self.pos), _ptr(self.vel), _ptr(self.ints), _ptr(actions),
                    _ACT_CODES[actions.dtype], self.reward_scaling, _ptr(pose_out if store_poses else obs_out),
                    _ptr(reward_out), _ptr(term_out), _ptr(trunc_out),
                    _lib.FLAG_F32 if self.float_flags else _lib.FLAG_U8, _ptr(info_out), self._stream())
        _lib.check(rc, "carenv_rollout_poses" if store_poses else "carenv_rollout")
        out = {"obs": obs_out, "reward": reward_out, "terminated": term_out, "truncated": trunc_out}
        if store_poses:
            out["poses"] = pose_out
        if info_out is not None:
            out["info"] = self._info_dict(info_out)
        return out

    def observe(self, poses: torch.Tensor, index: torch.Tensor | None = None, out: torch.Tensor | None = None):
        """Observations [M, 18] float32 from pose records (``rollout(..., store_poses=True)["poses"]``, any leading
        shape): all records in order, or the records ``index`` (int64, into the flattened record array) names —
        the minibatch gather of train.py:233 without ever materialising the [T, N, 18] observation buffer."""
        if not (poses.is_cuda and poses.dtype == torch.float64 and poses.shape[-1] == 4 and poses.is_contiguous()):
            raise ValueError("poses must be a contiguous CUDA float64 tensor [..., 4] from rollout(store_poses=True)")
        if index is not None:
            if not (index.is_cuda and index.dtype == torch.int64):
                raise ValueError("index must be a CUDA int64 tensor")
            index = index.reshape(-1).contiguous()
        m = index.numel() if index is not None else poses.numel() // 4
        if out is None:
            out = torch.empty((m, OBS_DIM), dtype=torch.float32, device=self.device)
        elif tuple(out.shape) != (m, OBS_DIM) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous float32 tensor of shape {(m, OBS_DIM)}")
        with torch.cuda.device(self.device):
            rc = self._L.carenv_observe(self._handle, m, _ptr(poses), _ptr(index), _ptr(out), self._stream())
        _lib.check(rc, "carenv_observe")
        return out

    def render(self, env_indices=(0,), width: int = 1280, height: int = 720) -> torch.Tensor:
        """Headless ``render_mode="rgb_array"`` frames (lib/car_env.py:762-812) of a few environments in their CURRENT
        state: uint8 CUDA tensor [k, height, width, 3].  ``.cpu().numpy()`` is what train.py:23-50 (log_video) feeds
        to cv2.  The picture shows background, corridor, walls, active gates (next one yellow), rays and the car as a
        box; it is a diagnostic rendering, not pygame's rasteriser."""
        if self._needs_reset:
            raise _lib.CarEnvError("call reset() before render()")
        idx = torch.as_tensor(list(env_indices), dtype=torch.int32, device=self.device).reshape(-1)
        if idx.numel() == 0 or int(idx.min()) < 0 or int(idx.max()) >= self.num_envs:
            raise ValueError("env_indices must name existing environments")
        if (width, height) != (1280, 720):
            raise ValueError("the canvas is the reference's 1280 x 720 (track coordinates are pixels of that canvas)")
        out = torch.empty((idx.numel(), height, width, 3), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            rc = self._L.carenv_render(self._handle, self.track.n_outer, idx.numel(), _ptr(idx), _ptr(self.pos),
                                       _ptr(self.ints), _ptr(self._obs), width, height, _ptr(out), self._stream())
        _lib.check(rc, "carenv_render")
        return out

    def set_option(self, name: str, value: int) -> None:
        _lib.check(self._L.carenv_set_option(self._handle, name.encode(), int(value)), "carenv_set_option")

    def slow_path_counts(self, reset: bool = True) -> dict:
        """How often the float64 re-evaluation ran since the last call (synchronises)."""
        out = (C.c_ulonglong * 4)()
        _lib.check(self._L.carenv_stats(self._handle, out, int(reset)), "carenv_stats")
        return {"line": out[0], "band": out[1], "gate": out[2], "tiny": out[3]}

    def close(self):
        if self._handle is not None:
            self._L.carenv_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiTrackVecEnv:
    """Several tracks in one vector env (the reference selects a track per env with
    ``reset(options={"track_path": ...})``, lib/car_env.py:621-628).  Every environment carries a track id and ALL of
    them are stepped by ONE kernel launch (carenv_multi_rollout); environments of different tracks may be mixed in
    any order.  Rows are bit-identical to a single-track VecCarEnv stepped with the same actions.

        MultiTrackVecEnv([(path_a, 300), (path_b, 500)])                 # contiguous groups
        MultiTrackVecEnv(track_paths=[path_a, path_b], track_ids=ids)     # ids: [N] ints, any assignment
    """

    def __init__(self, groups=None, device="cuda", reward_scaling: float = 1.0, float_flags: bool = False,
                 track_paths=None, track_ids=None):
        self._L = _lib.lib()
        if not torch.cuda.is_available():
            raise _lib.CarEnvError("MultiTrackVecEnv needs a CUDA device (there is no CPU implementation)")
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if groups is not None:
            if not groups:
                raise ValueError("need at least one (track_path, n_envs) group")
            track_paths = [p for p, _ in groups]
            track_ids = np.concatenate([np.full(int(c), i, np.int32) for i, (_, c) in enumerate(groups)])
        if track_paths is None or track_ids is None:
            raise ValueError("pass groups=[(track_path, n_envs), ...] or track_paths=[...] and track_ids=[...]")
        self.reward_scaling, self.float_flags = float(reward_scaling), bool(float_flags)
        self._multi = None
        self._tracks: list[VecCarEnv] = []
        self._set_tracks(list(track_paths))
        ids = torch.as_tensor(np.asarray(track_ids), dtype=torch.int32).reshape(-1)
        self.num_envs = int(ids.numel())
        self._set_ids(ids)
        n, dev = self.num_envs, self.device
        fdt = torch.float32 if float_flags else torch.uint8
        self.pos = torch.zeros((n, 2), dtype=torch.float64, device=dev)
        self.vel = torch.zeros((n, 2), dtype=torch.float64, device=dev)
        self.ints = torch.zeros((n, 4), dtype=torch.int32, device=dev)
        self._obs = torch.empty((n, OBS_DIM), dtype=torch.float32, device=dev)
        self._rew = torch.empty((n,), dtype=torch.float32, device=dev)
        self._term = torch.empty((n,), dtype=fdt, device=dev)
        self._trunc = torch.empty((n,), dtype=fdt, device=dev)
        self._info = torch.empty((n, 4), dtype=torch.int32, device=dev)
        self.single_observation_space = self._tracks[0].single_observation_space
        self.single_action_space = self._tracks[0].single_action_space
        self._needs_reset = True

    def _set_tracks(self, paths):
        """One single-track handle per path (they own the tables), then the multi-track handle over them."""
        tracks = [VecCarEnv(1, p, device=self.device) for p in paths]
        arr = (C.c_void_p * len(tracks))(*[t._handle for t in tracks])
        multi = C.c_void_p()
        _lib.check(self._L.carenv_multi_create(arr, len(tracks), C.byref(multi)), "carenv_multi_create")
        if self._multi is not None:
            self._L.carenv_multi_destroy(self._multi)
            for t in self._tracks:
                t.close()
        self._multi, self._tracks, self.track_paths = multi, tracks, list(paths)
        self._needs_reset = True

    def _set_ids(self, ids):
        ids = torch.as_tensor(np.asarray(ids.cpu() if isinstance(ids, torch.Tensor) else ids), dtype=torch.int32).reshape(-1)
        if ids.numel() != self.num_envs:
            raise ValueError(f"expected {self.num_envs} track ids, got {ids.numel()}")
        if int(ids.min()) < 0 or int(ids.max()) >= len(self._tracks):
            raise ValueError(f"track ids must be in 0..{len(self._tracks) - 1}")
        self.track_ids = ids.to(self.device)
        self._needs_reset = True

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self, seed=None, options=None):
        """``options={"track_paths": [...]}`` replaces the track list (lib/car_env.py:621-628 for every env),
        ``options={"track_ids": [N ints]}`` re-assigns environments to tracks; both may be given.  Any other option
        is refused rather than silently ignored.  ``seed`` is accepted and ignored like in the reference."""
        if options:
            unknown = set(options) - {"track_paths", "track_ids"}
            if unknown:
                raise ValueError(f"MultiTrackVecEnv.reset does not understand options {sorted(unknown)}; "
                                 "use 'track_paths' (one path per track) and / or 'track_ids' (one id per env)")
            if "track_paths" in options:
                paths = list(options["track_paths"])
                if "track_ids" not in options and len(paths) != len(self._tracks):
                    raise ValueError(f"expected {len(self._tracks)} track paths (or pass new track_ids too), got {len(paths)}")
                self._set_tracks(paths)
            if "track_ids" in options:
                self._set_ids(options["track_ids"])
        with torch.cuda.device(self.device):
            rc = self._L.carenv_multi_reset(self._multi, self.num_envs, _ptr(self.track_ids), _ptr(self.pos),
                                            _ptr(self.vel), _ptr(self.ints), _ptr(self._obs), self._stream())
        _lib.check(rc, "carenv_multi_reset")
        self._needs_reset = False
        zeros = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        return self._obs, {"gates_passed": zeros, "time_passed": zeros.clone()}

    def step(self, actions):
        """One step of every environment, all tracks in one launch.  Returns views of internal buffers that the next
        step overwrites (see VecCarEnv)."""
        if self._needs_reset:
            raise _lib.CarEnvError("call reset() before step()")
        if not (isinstance(actions, torch.Tensor) and actions.is_cuda):
            actions = torch.as_tensor(np.asarray(actions), device=self.device)
        if actions.dtype not in _ACT_CODES:
            actions = actions.to(torch.int64)
        actions = actions.reshape(-1).contiguous()
        if actions.numel() != self.num_envs:
            raise ValueError(f"expected {self.num_envs} actions, got {actions.numel()}")
        with torch.cuda.device(self.device):
            rc = self._L.carenv_multi_rollout(self._multi, self.num_envs, 1, _ptr(self.track_ids), _ptr(self.pos),
                                              _ptr(self.vel), _ptr(self.ints), _ptr(actions), _ACT_CODES[actions.dtype],
                                              self.reward_scaling, _ptr(self._obs), _ptr(self._rew), _ptr(self._term),
                                              _ptr(self._trunc), _lib.FLAG_F32 if self.float_flags else _lib.FLAG_U8,
                                              _ptr(self._info), self._stream())
        _lib.check(rc, "carenv_multi_rollout")
        term = self._term if self.float_flags else self._term.view(torch.bool)
        trunc = self._trunc if self.float_flags else self._trunc.view(torch.bool)
        info = {"gates_passed": self._info[:, 0], "time_passed": self._info[:, 1], "next_gate_index": self._info[:, 2],
                "events": self._info[:, 3]}
        return self._obs, self._rew, term, trunc, info

    def close(self):
        if self._multi is not None:
            self._L.carenv_multi_destroy(self._multi)
            self._multi = None
        for t in self._tracks:
            t.close()
        self._tracks = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
