"""TEST INFRASTRUCTURE — CPU oracle ("port") for the CarEnv step/reset path.

A from-scratch Python restatement of the algorithm in the reference's
``lib/car_env.py``.  It exists to CHECK the CUDA path and to serve as the timed
CPU baseline (``bench.py``'s ``cpu_baseline`` leg / ``--impl reference``).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` may import it; the
product package ``ppo_car_b200`` never does and has no CPU fallback.

Parity status: PINNED.  ``tests/golden/*.npz`` were produced by the unmodified
reference (``tests/golden/make_golden.py``, imported through
``oracle/ref_import.py``) and ``tests/test_oracle_golden.py`` requires this port
to reproduce them BIT-EXACTLY (float64 rewards, float32 observations, all flags
and counters).  The reference itself ships no tests or golden vectors
(SURVEY §4); the same-step autoreset contract comes from gymnasium 0.29.1's
AsyncVectorEnv (un-vendored dependency) and is restated in
``PortVecEnv.step``.

To stay bit-identical with the reference every floating-point operation is kept
in the reference's order and the same numpy entry points are used for the
operations whose rounding is implementation-defined (``np.cos``, ``np.sin``,
``np.radians``, ``np.linalg.norm``); everything else is IEEE double arithmetic
on Python floats.  The redundant work of the reference (full gate scan,
separate collision casts) is kept on purpose so the port is also a fair cost
model of the reference's CPU path.

Reference map (file:line in /root/reference):
  ray_hit            lib/car_env.py:155-184   Ray.cast
  ray_distance       lib/car_env.py:186-213   Ray.get_distance
  refresh_rays       lib/car_env.py:151-153, 463-466
  touches            lib/car_env.py:376-392   Car.check_collision
  first_touched_gate lib/car_env.py:394-408   Car.get_passed_gate
  thrust / turn      lib/car_env.py:416-442   Car.move_car
  integrate          lib/car_env.py:444-469   Car.update
  load_track         lib/car_env.py:535-567
  observe            lib/car_env.py:569-597
  PortCarEnv.reset   lib/car_env.py:605-691
  PortCarEnv.step    lib/car_env.py:693-760
"""
from __future__ import annotations

import json

import numpy as np

WIDTH, HEIGHT = 1280, 720          # lib/car_env.py:488-489
TIME_LIMIT = 1000                  # lib/car_env.py:491
TURN_DEG, V_MAX, ACCEL, FRICTION = 5.0, 10.0, 0.8, 0.2   # lib/car_env.py:223-226
N_RAYS = 12                        # lib/car_env.py:227
NO_HIT = 1000.0                    # lib/car_env.py:198
TOUCH = 10.0                       # lib/car_env.py:387
CARDINAL = (0, 3, 6, 9)            # lib/car_env.py:389


def load_track(path: str) -> dict:
    """JSON -> pixel coordinates (x*1280, y*720); the angle is left in degrees."""
    with open(path, "r") as fh:
        raw = json.load(fh)
    px = lambda pts: [[x * WIDTH, y * HEIGHT] for x, y in pts]
    outer, inner, gpts = px(raw["outer_track_points"]), px(raw["inner_track_points"]), px(raw["reward_gates"])
    walls = [(float(a[0]), float(a[1]), float(b[0]), float(b[1])) for a, b in zip(outer[:-1], outer[1:])]
    walls += [(float(a[0]), float(a[1]), float(b[0]), float(b[1])) for a, b in zip(inner[:-1], inner[1:])]
    gates = [(float(a[0]), float(a[1]), float(b[0]), float(b[1])) for a, b in zip(gpts[::2], gpts[1::2])]
    return dict(walls=walls, gates=gates,
                start=(raw["initial_position"][0] * WIDTH, raw["initial_position"][1] * HEIGHT),
                angle=raw["initial_angle"])


def ray_hit(ox, oy, dx, dy, seg):
    """Line/line intersection; returns the hit point or None."""
    x1, y1, x2, y2 = seg
    x3, y3 = ox, oy
    x4, y4 = ox + dx, oy + dy
    den = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4)
    if den == 0:
        return None
    t = ((x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4)) / den
    u = -((x1 - x2) * (y1 - y3) - (y1 - y2) * (x1 - x3)) / den
    if 0 < t < 1 and u > 0:
        return x1 + t * (x2 - x1), y1 + t * (y2 - y1)
    return None


def ray_distance(ox, oy, dx, dy, segs):
    best = NO_HIT
    for seg in segs:
        hit = ray_hit(ox, oy, dx, dy, seg)
        if hit is not None:
            d = float(np.linalg.norm(np.array([ox - hit[0], oy - hit[1]])))
            if d < best:
                best = d
    return best


class PortCarEnv:
    """Single environment, same call surface as the reference CarEnv (reset/step)."""

    def __init__(self, track_path: str):
        self.track = load_track(track_path)
        self.walls, self.gates = self.track["walls"], self.track["gates"]
        self.pos = [0.0, 0.0]
        self.vel = [0.0, 0.0]
        self.acc = [0.0, 0.0]
        self.rot = 0.0
        self.rays = [(0.0, 0.0, 1.0, 0.0)] * N_RAYS     # (ox, oy, dx, dy)
        self.destroyed = False
        self.t = 0
        self.next_gate = 0
        self.passed = 0
        self.remaining = len(self.gates)
        self.active = [True] * len(self.gates)

    # -- car ---------------------------------------------------------------
    def refresh_rays(self):
        ox, oy = self.pos
        self.rays = []
        for k in range(N_RAYS):
            ang = self.rot + k * (360 // N_RAYS)
            self.rays.append((ox, oy, float(np.cos(np.radians(ang))), float(np.sin(np.radians(ang)))))

    def touches(self, segs) -> bool:
        for k in CARDINAL:
            if ray_distance(*self.rays[k], segs) < TOUCH:
                return True
        return False

    def first_touched_gate(self):
        for g, seg in enumerate(self.gates):
            if self.active[g] and self.touches([seg]):
                return g
        return None

    def thrust(self, sign: float):
        c, s = float(np.cos(np.radians(self.rot))), float(np.sin(np.radians(self.rot)))
        if sign > 0:
            self.acc = [c * ACCEL, s * ACCEL]
        else:
            self.acc = [-c * ACCEL, -s * ACCEL]

    def integrate(self):
        self.vel[0] += self.acc[0]
        self.vel[1] += self.acc[1]
        if float(np.linalg.norm(np.array(self.acc))) == 0:
            self.vel[0] *= 1 - FRICTION
            self.vel[1] *= 1 - FRICTION
        self.vel = [min(max(v, -V_MAX), V_MAX) for v in self.vel]
        self.pos[0] += self.vel[0]
        self.pos[1] += self.vel[1]
        self.acc = [0.0, 0.0]
        self.refresh_rays()
        if self.touches(self.walls):
            self.destroyed = True

    # -- env ---------------------------------------------------------------
    def observe(self) -> np.ndarray:
        o = [self.pos[0] / WIDTH, self.pos[1] / HEIGHT, self.vel[0] / V_MAX, self.vel[1] / V_MAX,
             float(np.cos(np.radians(self.rot))), float(np.sin(np.radians(self.rot)))]
        for r in self.rays:
            o.append(ray_distance(*r, self.walls) / 1000.0)
        return np.array(o, dtype=np.float32)

    def info(self) -> dict:
        return {"gates_passed": self.passed, "time_passed": self.t}

    def reset(self):
        self.t = 0
        self.pos = [self.track["start"][0], self.track["start"][1]]
        self.rot = self.track["angle"]
        self.vel, self.acc = [0.0, 0.0], [0.0, 0.0]
        self.passed, self.next_gate, self.remaining = 0, 0, len(self.gates)
        self.destroyed = False
        self.active = [True] * len(self.gates)
        self.integrate()
        return self.observe(), self.info()

    def step(self, action: int):
        reward = 0.0
        fwd, bwd = action in (0, 4, 5), action in (1, 6, 7)
        left, right = action in (2, 4, 6), action in (3, 5, 7)
        if fwd:
            self.thrust(+1.0)
            reward += 0.01
        elif bwd:
            self.thrust(-1.0)
        if left:
            self.rot -= TURN_DEG
        elif right:
            self.rot += TURN_DEG

        g = self.first_touched_gate()
        if g is not None and g == self.next_gate:
            reward += 1.0
            self.remaining -= 1
            self.passed += 1
            if self.remaining == 0:
                reward += 10.0
                self.active = [True] * len(self.gates)
                self.remaining = len(self.gates)
                self.next_gate = 0
            else:
                self.active[g] = False
                self.next_gate += 1

        self.integrate()
        self.t += 1
        terminated = truncated = False
        if self.destroyed:
            terminated = True
            reward -= 3.0
        elif self.t >= TIME_LIMIT:
            truncated = True
        return self.observe(), reward, terminated, truncated, self.info()


class PortVecEnv:
    """N independent PortCarEnv objects with same-step autoreset."""

    def __init__(self, n_envs: int, track_path: str):
        self.envs = [PortCarEnv(track_path) for _ in range(n_envs)]

    def reset(self) -> np.ndarray:
        return np.stack([e.reset()[0] for e in self.envs])

    def step(self, actions):
        n = len(self.envs)
        obs = np.zeros((n, 6 + N_RAYS), np.float32)
        fobs = np.zeros((n, 6 + N_RAYS), np.float32)
        rew = np.zeros(n, np.float64)
        term = np.zeros(n, np.bool_)
        trunc = np.zeros(n, np.bool_)
        gates = np.zeros(n, np.int32)
        tpass = np.zeros(n, np.int32)
        nxt = np.zeros(n, np.int32)
        for i, (e, a) in enumerate(zip(self.envs, actions)):
            o, r, te, tr, info = e.step(int(a))
            fobs[i] = o
            rew[i], term[i], trunc[i] = r, te, tr
            gates[i], tpass[i], nxt[i] = info["gates_passed"], info["time_passed"], e.next_gate
            if te or tr:
                o, _ = e.reset()
            obs[i] = o
        return obs, rew, term, trunc, dict(gates_passed=gates, time_passed=tpass,
                                           next_gate_index=nxt, final_obs=fobs)


def gae_port(rew, val, term, trunc, last_val, last_term, last_trunc, gamma=0.99, gae_lambda=0.95):
    """float32 GAE in the reference's own operation order (lib/buffer.py:36-64).

    numpy float32 restatement of the torch loop: python-float constants are
    rounded to float32 when they meet a float32 tensor (torch scalar semantics),
    ``gamma * gae_lambda`` is first evaluated in double (lib/buffer.py:61) and
    every array operation is individually rounded to float32.
    """
    f = np.float32
    rew, val, term, trunc = (np.asarray(a, f) for a in (rew, val, term, trunc))
    T = rew.shape[0]
    g, gl, one = f(gamma), f(gamma * gae_lambda), f(1.0)
    adv = np.zeros_like(rew)
    last = np.zeros(rew.shape[1:], f)
    for t in range(T - 1, -1, -1):
        nv = np.asarray(last_val, f).reshape(-1) if t == T - 1 else val[t + 1]
        tm = one - (np.asarray(last_term, f).reshape(-1) if t == T - 1 else term[t + 1])
        um = one - (np.asarray(last_trunc, f).reshape(-1) if t == T - 1 else trunc[t + 1])
        delta = (rew[t] + (g * nv) * tm) - val[t]
        last = delta + ((gl * tm) * um) * last
        adv[t] = last
    return adv, adv + val
