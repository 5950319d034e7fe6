"""CPU restatement of the classic-control envs behind jorldy/core/env/gym_env.py.

JORLDY wrapper semantics (pinned by reading the reference):
  * Cartpole.step (gym_env.py:70-83): action.item(); continuous variant thresholds `action < 0`
    (:74-75); score += gym reward; reward := -1 if done else 0.1 (:78); outputs expand to (1, ?).
  * _Gym.reset (gym_env.py:32-36): score = 0, state expanded to (1, state_size).
Physics = gym==0.23.0 classic_control (third-party, not vendored: PARITY UNPINNED, see package
docstring), restated from its published equations with python-float (f64) state and an f32 copy
returned as the observation, wrapped in TimeLimit (CartPole-v1: 500, Pendulum-v1 / MountainCar-v0:
200) whose truncation also reports done=True.

Batched over N envs with numpy; reset draws come from oracle.philox so they match the CUDA env.
"""
import math

import numpy as np

from . import philox


def _reset_uniforms(seed, stream_base, env_ids, episode, n_draws):
    """Same counter scheme as csrc/env_classic.cu: ctr = 2*episode (+1), 2 doubles per Philox call."""
    out = []
    for k in range((n_draws + 1) // 2):
        a, b, c, d = philox.philox4x32(seed, stream_base + env_ids.astype(np.uint64), 2 * episode.astype(np.uint64) + np.uint64(k))
        out.append(philox.u01_double(a, b))
        out.append(philox.u01_double(c, d))
    return out[:n_draws]


class CartPoleBatch:
    gravity = 9.8
    masscart = 1.0
    masspole = 0.1
    total_mass = masspole + masscart
    length = 0.5
    polemass_length = masspole * length
    force_mag = 10.0
    tau = 0.02
    theta_threshold_radians = 12 * 2 * math.pi / 360
    x_threshold = 2.4
    max_steps = 500

    def __init__(self, n, seed=0, stream_base=0, auto_reset=True):
        self.n, self.seed, self.stream_base, self.auto_reset = n, seed, stream_base, auto_reset
        self.ids = np.arange(n, dtype=np.int64)
        self.phys = np.zeros((n, 4), dtype=np.float64)
        self.elapsed = np.zeros(n, dtype=np.int32)
        self.episode = np.zeros(n, dtype=np.int64)
        self.score = np.zeros(n, dtype=np.float32)

    def _draw(self, mask):
        u = _reset_uniforms(self.seed, np.uint64(self.stream_base), self.ids[mask], self.episode[mask], 4)
        s = np.stack([-0.05 + 0.1 * x for x in u], axis=1)
        self.phys[mask] = s
        self.episode[mask] += 1
        self.elapsed[mask] = 0
        self.score[mask] = 0

    def reset(self):
        self._draw(np.ones(self.n, dtype=bool))
        return self.phys.astype(np.float32)

    def step(self, action):
        a = np.asarray(action).reshape(self.n)
        if a.dtype.kind == "f":
            a = np.where(a < 0, 0, 1)          # gym_env.py:74-75
        x, x_dot, theta, theta_dot = (self.phys[:, i].copy() for i in range(4))
        force = np.where(a == 1, self.force_mag, -self.force_mag)
        costheta, sintheta = np.cos(theta), np.sin(theta)
        temp = (force + self.polemass_length * (theta_dot * theta_dot) * sintheta) / self.total_mass
        thetaacc = (self.gravity * sintheta - costheta * temp) / (
            self.length * (4.0 / 3.0 - self.masspole * (costheta * costheta) / self.total_mass))
        xacc = temp - self.polemass_length * thetaacc * costheta / self.total_mass
        x = x + self.tau * x_dot
        x_dot = x_dot + self.tau * xacc
        theta = theta + self.tau * theta_dot
        theta_dot = theta_dot + self.tau * thetaacc
        term = (x < -self.x_threshold) | (x > self.x_threshold) | (theta < -self.theta_threshold_radians) | (
            theta > self.theta_threshold_radians)
        self.elapsed += 1
        done = term | (self.elapsed >= self.max_steps)
        self.score += 1.0
        self.phys = np.stack([x, x_dot, theta, theta_dot], axis=1)
        next_obs = self.phys.astype(np.float32)
        reward = np.where(done, -1.0, 0.1).astype(np.float32)     # gym_env.py:78
        if self.auto_reset and done.any():
            self._draw(done)
        return next_obs, reward, done

    @property
    def obs(self):
        return self.phys.astype(np.float32)


class PendulumBatch:
    """gym 0.23.0 Pendulum-v1 (max_speed 8, max_torque 2, dt 0.05, g 10, m = l = 1; 200-step TimeLimit) behind
    _Gym.step's action rescale ((a+1)/2*(high-low)+low, gym_env.py:41-45).  PARITY UNPINNED (third-party)."""
    max_speed, max_torque, dt, g, m, l, max_steps = 8.0, 2.0, 0.05, 10.0, 1.0, 1.0, 200

    def __init__(self, n, seed=0, stream_base=0, auto_reset=True):
        self.n, self.seed, self.stream_base, self.auto_reset = n, seed, stream_base, auto_reset
        self.ids = np.arange(n, dtype=np.int64)
        self.phys = np.zeros((n, 2), dtype=np.float64)
        self.elapsed = np.zeros(n, dtype=np.int32)
        self.episode = np.zeros(n, dtype=np.int64)

    def _draw(self, mask):
        u = _reset_uniforms(self.seed, np.uint64(self.stream_base), self.ids[mask], self.episode[mask], 2)
        self.phys[mask, 0] = -math.pi + (2 * math.pi) * u[0]
        self.phys[mask, 1] = -1.0 + 2.0 * u[1]
        self.episode[mask] += 1
        self.elapsed[mask] = 0

    def _obs(self):
        th, thdot = self.phys[:, 0], self.phys[:, 1]
        return np.stack([np.cos(th), np.sin(th), thdot], axis=1).astype(np.float32)

    def reset(self):
        self._draw(np.ones(self.n, dtype=bool))
        return self._obs()

    def step(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(self.n)
        scaled = ((a + np.float32(1.0)) / np.float32(2.0)) * np.float32(4.0) + np.float32(-2.0)
        u = np.clip(scaled.astype(np.float64), -self.max_torque, self.max_torque)
        th, thdot = self.phys[:, 0].copy(), self.phys[:, 1].copy()
        an = np.mod(th + math.pi, 2 * math.pi) - math.pi
        costs = an * an + 0.1 * (thdot * thdot) + 0.001 * (u * u)
        newthdot = thdot + (3 * self.g / (2 * self.l) * np.sin(th) + 3.0 / (self.m * (self.l * self.l)) * u) * self.dt
        newthdot = np.clip(newthdot, -self.max_speed, self.max_speed)
        newth = th + newthdot * self.dt
        self.phys = np.stack([newth, newthdot], axis=1)
        self.elapsed += 1
        done = self.elapsed >= self.max_steps
        next_obs = self._obs()
        reward = (-costs).astype(np.float32)
        if self.auto_reset and done.any():
            self._draw(done)
        return next_obs, reward, done


class MountainCarBatch:
    """gym 0.23.0 MountainCar-v0 (200-step TimeLimit).  PARITY UNPINNED (third-party)."""
    min_position, max_position, max_speed, goal_position, goal_velocity = -1.2, 0.6, 0.07, 0.5, 0.0
    force, gravity, max_steps = 0.001, 0.0025, 200

    def __init__(self, n, seed=0, stream_base=0, auto_reset=True):
        self.n, self.seed, self.stream_base, self.auto_reset = n, seed, stream_base, auto_reset
        self.ids = np.arange(n, dtype=np.int64)
        self.phys = np.zeros((n, 2), dtype=np.float64)
        self.elapsed = np.zeros(n, dtype=np.int32)
        self.episode = np.zeros(n, dtype=np.int64)

    def _draw(self, mask):
        u = _reset_uniforms(self.seed, np.uint64(self.stream_base), self.ids[mask], self.episode[mask], 2)
        self.phys[mask, 0] = -0.6 + 0.2 * u[0]
        self.phys[mask, 1] = 0.0
        self.episode[mask] += 1
        self.elapsed[mask] = 0

    def reset(self):
        self._draw(np.ones(self.n, dtype=bool))
        return self.phys.astype(np.float32)

    def step(self, action):
        a = np.asarray(action).reshape(self.n).astype(np.float64)
        pos, vel = self.phys[:, 0].copy(), self.phys[:, 1].copy()
        vel = vel + ((a - 1) * self.force + np.cos(3 * pos) * (-self.gravity))
        vel = np.clip(vel, -self.max_speed, self.max_speed)
        pos = pos + vel
        pos = np.clip(pos, self.min_position, self.max_position)
        vel = np.where((pos == self.min_position) & (vel < 0), 0.0, vel)
        term = (pos >= self.goal_position) & (vel >= self.goal_velocity)
        self.phys = np.stack([pos, vel], axis=1)
        self.elapsed += 1
        done = term | (self.elapsed >= self.max_steps)
        next_obs = self.phys.astype(np.float32)
        reward = np.full(self.n, -1.0, dtype=np.float32)
        if self.auto_reset and done.any():
            self._draw(done)
        return next_obs, reward, done


class SyntheticControlBatch:
    """numpy restatement of csrc/env_synth.cu (synthetic generator with Hopper-v3 DIMENSIONS standing in for
    jorldy/core/env/mujoco.py: our own definition, not a restatement of MuJoCo — "parity" here means the CUDA kernel
    computes what its header says).  float32 arithmetic in the kernel's operation order; Philox counters:
    16 t + k for the 4 normals of dims 4k..4k+3, 16 t + 15 for the done draw, 2^40 + 16 episode + k for reset."""

    def __init__(self, n, D=11, A=3, seed=0, stream_base=0, auto_reset=True, p_done=1e-3, max_steps=1000, Ws=None, Wa=None):
        self.n, self.D, self.A, self.seed, self.stream_base, self.auto_reset = n, D, A, seed, stream_base, auto_reset
        self.p_done, self.max_steps = np.float32(p_done), max_steps
        self.Ws, self.Wa = np.asarray(Ws, np.float32), np.asarray(Wa, np.float32)
        self.ids = np.arange(n, dtype=np.uint64)
        self.obs = np.zeros((n, D), np.float32)
        self.elapsed = np.zeros(n, np.int32)
        self.episode = np.zeros(n, np.int64)
        self.tcount = np.zeros(n, np.int64)

    def _reset_draw(self, mask):
        ids, ep = self.ids[mask], self.episode[mask].astype(np.uint64)
        cols = []
        for k in range((self.D + 3) // 4):
            w = philox.philox4x32(self.seed, np.uint64(self.stream_base) + ids, (np.uint64(1) << np.uint64(40)) + np.uint64(16) * ep + np.uint64(k))
            cols += [np.float32(-0.05) + np.float32(0.1) * philox.u01_float(x) for x in w]
        self.obs[mask] = np.stack(cols[:self.D], axis=1).astype(np.float32)
        self.episode[mask] += 1
        self.elapsed[mask] = 0

    def reset(self):
        self._reset_draw(np.ones(self.n, bool))
        return self.obs.copy()

    def step(self, action):
        a = np.asarray(action, np.float32).reshape(self.n, self.A)
        s, t = self.obs, self.tcount.astype(np.uint64)
        stream = np.uint64(self.stream_base) + self.ids
        sn = np.zeros_like(s)
        sq = np.zeros(self.n, np.float32)
        for k0 in range(0, self.D, 4):
            w = philox.philox4x32(self.seed, stream, np.uint64(16) * t + np.uint64(k0 >> 2))
            nrm = []
            for h in range(2):
                u1 = ((w[2 * h] >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)
                u2 = philox.u01_float(w[2 * h + 1])
                rad = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
                ang = (np.float32(2.0) * u2).astype(np.float64) * np.pi
                nrm += [(rad * np.cos(ang).astype(np.float32)).astype(np.float32), (rad * np.sin(ang).astype(np.float32)).astype(np.float32)]
            for q in range(4):
                k = k0 + q
                if k >= self.D:
                    break
                z = np.zeros(self.n, np.float32)
                for j in range(self.D):
                    z = (z + self.Ws[k, j] * s[:, j]).astype(np.float32)
                for j in range(self.A):
                    z = (z + self.Wa[k, j] * a[:, j]).astype(np.float32)
                v = (np.tanh(z).astype(np.float32) + np.float32(0.01) * nrm[q]).astype(np.float32)
                sn[:, k] = v
                sq = (sq + v * v).astype(np.float32)
        reward = (-sq / np.float32(self.D)).astype(np.float32)
        wd = philox.philox4x32(self.seed, stream, np.uint64(16) * t + np.uint64(15))
        self.tcount += 1
        self.elapsed += 1
        done = (philox.u01_float(wd[0]) < self.p_done) | (self.elapsed >= self.max_steps)
        next_obs = sn.copy()
        self.obs = sn
        if self.auto_reset and done.any():
            self._reset_draw(done)
        return next_obs, reward, done
