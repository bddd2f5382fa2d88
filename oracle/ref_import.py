"""TEST INFRASTRUCTURE (container-only): import the UNMODIFIED reference CarEnv.

This module is used only by ``tests/golden/make_golden.py`` and by the optional
``-m "not gpu"`` cross-checks that run when ``/root/reference`` is mounted.  It is
never imported by the product package and never runs on the GPU box (the reference
tree does not exist there).

The reference's ``lib/car_env.py`` imports ``gymnasium`` and ``pygame``
(lib/car_env.py:4-7), neither of which is installed.  Only a base class, two
space constructors, ``register`` and three pygame names are touched on the
headless path (lib/car_env.py:250-255, 472, 485, 522-525, 617, 815-816), so two
inert ``types.ModuleType`` stubs are injected into ``sys.modules``.  No arithmetic
lives in the stubs: every number the harness produces comes from the reference's
own code running on numpy.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PPO_CAR_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "lib", "car_env.py"))


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
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from lib.car_env import CarEnv  # type: ignore
    from lib.buffer import Buffer  # type: ignore

    return CarEnv, Buffer


def track_path(name: str) -> str:
    return os.path.join(REFERENCE_ROOT, "tracks", name)


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
