"""TEST INFRASTRUCTURE: import the UNMODIFIED reference CarEnv / Buffer.

Two places can hold the reference files:

  * ``/root/reference`` (or ``$PPO_CAR_REFERENCE``) — the read-only mount of the build container;
  * ``oracle/_ref/`` — a byte-for-byte copy of the four hot-path files made by ``oracle/make_ref.py``
    (git-ignored, travels to the GPU box with the snapshot), verified against its MANIFEST.json.

Used by ``tests/`` (golden-vector generation, live cross-checks of the port AND of the CUDA path),
and by ``bench.py``'s CPU legs (``cpu_baseline``, ``--impl reference``: the reference's own classes
timed on the host cores).  It is never imported by the product package ``ppo_car_b200``.

The reference's ``lib/car_env.py`` imports ``gymnasium`` and ``pygame``
(lib/car_env.py:4-7), neither of which is installed.  Only a base class, two
space constructors, ``register`` and three pygame names are touched on the
headless path (lib/car_env.py:250-255, 472, 485, 522-525, 617, 815-816), so two
inert ``types.ModuleType`` stubs are injected into ``sys.modules``.  No arithmetic
lives in the stubs: every number the harness produces comes from the reference's
own code running on numpy.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_MOUNT = os.environ.get("PPO_CAR_REFERENCE", "/root/reference")
_COPY = os.path.join(_HERE, "_ref")


def _copy_is_intact() -> bool:
    try:
        man = json.load(open(os.path.join(_COPY, "MANIFEST.json")))
        for rel, digest in man["files"].items():
            with open(os.path.join(_COPY, rel), "rb") as fh:
                if hashlib.sha256(fh.read()).hexdigest() != digest:
                    return False
        return "lib/car_env.py" in man["files"]
    except Exception:
        return False


def reference_root(prefer_copy: bool = False) -> str | None:
    """Directory holding lib/car_env.py: the mount when present, else the verified oracle/_ref copy."""
    mount_ok = os.path.isfile(os.path.join(_MOUNT, "lib", "car_env.py"))
    if mount_ok and not prefer_copy:
        return _MOUNT
    if _copy_is_intact():
        return _COPY
    return _MOUNT if mount_ok else None


REFERENCE_ROOT = reference_root() or _MOUNT


def reference_available() -> bool:
    return reference_root() is not None


def reference_kind() -> str:
    """"mount" (/root/reference), "copy" (oracle/_ref) or "absent"."""
    root = reference_root()
    return "absent" if root is None else ("mount" if root == _MOUNT else "copy")


def _install_stubs() -> None:
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")

        class Env:  # gymnasium.Env: reset(seed=, options=) is a no-op on this path
            metadata: dict = {}

            def reset(self, seed=None, options=None):
                return None

            def close(self):
                return None

        class _Space:
            def __init__(self, **kw):
                self.__dict__.update(kw)

        spaces = types.ModuleType("gymnasium.spaces")

        def Box(low, high, dtype=None, shape=None):
            return _Space(low=low, high=high, dtype=dtype, shape=getattr(low, "shape", shape))

        def Discrete(n):
            return _Space(n=n)

        spaces.Box, spaces.Discrete = Box, Discrete
        gym.Env, gym.spaces = Env, spaces
        gym.register = lambda *a, **k: None
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = spaces
    if "pygame" not in sys.modules:
        pg = types.ModuleType("pygame")

        class Surface:  # only evaluated inside annotations
            pass

        class _Img:
            def get_rect(self):
                return None

        pg.Surface = Surface
        pg.image = types.SimpleNamespace(load=lambda path: _Img())
        pg.transform = types.SimpleNamespace(scale=lambda img, size: img)
        pg.draw = types.SimpleNamespace()
        sys.modules["pygame"] = pg


def import_reference():
    """Return (CarEnv class, Buffer class) from the unmodified reference tree."""
    root = reference_root()
    if root is None:
        raise RuntimeError(f"reference not found: neither {_MOUNT} nor an intact {_COPY} (python oracle/make_ref.py)")
    _install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    from lib.car_env import CarEnv  # type: ignore
    from lib.buffer import Buffer  # type: ignore

    return CarEnv, Buffer


def track_path(name: str) -> str:
    return os.path.join(reference_root() or _MOUNT, "tracks", name)


class RefVecEnv:
    """N reference CarEnv objects + the same-step autoreset contract of
    gymnasium 0.29.1's AsyncVectorEnv worker (un-vendored): when a
    step ends the episode the env is reset at once and the RESET observation is
    returned together with the finished step's reward and flags."""

    def __init__(self, n_envs: int, track: str):
        CarEnv, _ = import_reference()
        self.envs = [CarEnv(track_path=track) for _ in range(n_envs)]
        self.track = track

    def reset(self):
        import numpy as np

        out = [e.reset(options={"track_path": self.track}) for e in self.envs]
        return np.stack([o for o, _ in out])

    def step(self, actions):
        import numpy as np

        n = len(self.envs)
        obs = np.zeros((n, 18), np.float32)
        rew = np.zeros(n, np.float64)
        term = np.zeros(n, np.bool_)
        trunc = np.zeros(n, np.bool_)
        gates = np.zeros(n, np.int32)   # info["gates_passed"] of the finished step
        tpass = np.zeros(n, np.int32)   # info["time_passed"] of the finished step
        nxt = np.zeros(n, np.int32)     # private next_gate_index after the step (pre-reset)
        fobs = np.zeros((n, 18), np.float32)  # pre-reset ("final") observation
        for i, (e, a) in enumerate(zip(self.envs, actions)):
            o, r, te, tr, info = e.step(int(a))
            fobs[i] = o
            rew[i], term[i], trunc[i] = r, te, tr
            gates[i], tpass[i] = info["gates_passed"], info["time_passed"]
            nxt[i] = e._CarEnv__next_gate_index
            if te or tr:
                o, _ = e.reset()
            obs[i] = o
        return obs, rew, term, trunc, dict(gates_passed=gates, time_passed=tpass,
                                           next_gate_index=nxt, final_obs=fobs)
