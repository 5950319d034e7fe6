"""Batched classic-control environments on the GPU (csrc/env_classic.cu).

Mirrors jorldy/core/env/gym_env.py: `Cartpole` (:61-83, reward -1 on done else 0.1),
`Pendulum` (:86-89), `MountainCar` (:92-95).  `num_envs` instances are stepped by ONE kernel
launch; the numpy-facing reset()/step() keep the reference's shapes with the leading dimension
= num_envs (1 by default, i.e. exactly the reference's (1, D) / (1, 1) arrays), while
reset_device()/step_device() hand out device tensors for the resident collect pipeline.
"""
import numpy as np
import torch

from ..dev import C, ptr, require_cuda, stream_ptr
from .base import BaseEnv

_KIND = {"cartpole": 0, "pendulum": 1, "mountain_car": 2}
_PHYS = {0: 4, 1: 2, 2: 2}
_OBS = {0: 4, 1: 3, 2: 2}
_LIMIT = {0: 500, 1: 200, 2: 200}        # gym TimeLimit of CartPole-v1 / Pendulum-v1 / MountainCar-v0


class _Classic(BaseEnv):
    kind_name = None

    def __init__(self, num_envs=1, seed=0, id=0, device=None, auto_reset=None, render=False, train_mode=True,
                 **kwargs):
        self.device = require_cuda(device)
        self.kind = _KIND[self.kind_name]
        self.num_envs = int(num_envs)
        self.seed = int(seed)
        self.id = int(id) if id is not None else 0
        self.stream_base = self.id << 32
        self.auto_reset = (self.num_envs > 1) if auto_reset is None else bool(auto_reset)
        self.state_size = _OBS[self.kind]
        self.max_steps = _LIMIT[self.kind]
        n, dev = self.num_envs, self.device
        self.phys = torch.zeros(n, _PHYS[self.kind], dtype=torch.float64, device=dev)
        self.obs = torch.zeros(n, self.state_size, dtype=torch.float32, device=dev)
        self.elapsed = torch.zeros(n, dtype=torch.int32, device=dev)
        self.episode = torch.zeros(n, dtype=torch.int64, device=dev)
        self._score = torch.zeros(n, dtype=torch.float32, device=dev)
        self.next_obs = torch.zeros(n, self.state_size, dtype=torch.float32, device=dev)
        self.reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self.done = torch.zeros(n, dtype=torch.float32, device=dev)
        self.stats = torch.zeros(2, dtype=torch.float32, device=dev)   # episodes finished, sum of scores
        self.render = render

    # ---- device API (resident pipeline) ---------------------------------------------------------
    def reset_device(self, mask=None):
        C.jb_env_classic_reset(self.kind, ptr(self.phys), ptr(self.obs), ptr(self.elapsed), ptr(self.episode),
                               ptr(self._score), ptr(mask), self.seed, self.stream_base, self.num_envs, stream_ptr())
        return self.obs

    def _action_kind(self, action):
        if action.dtype == torch.int64:
            return 0
        if action.dtype == torch.int32:
            return 1
        if action.dtype == torch.float32:
            return 2
        raise TypeError(f"unsupported action dtype {action.dtype}")

    def step_device(self, action):
        """action: device tensor [N] / [N,1].  Returns (next_obs, reward, done) device tensors that are
        overwritten by the next call; `self.obs` then holds the observation to act on next
        (post auto-reset where done)."""
        C.jb_env_classic_step(self.kind, ptr(self.phys), ptr(self.obs), ptr(self.elapsed), ptr(self.episode),
                              ptr(self._score), ptr(action), self._action_kind(action), ptr(self.next_obs),
                              ptr(self.reward), ptr(self.done), ptr(self.stats), int(self.auto_reset),
                              self.max_steps, self.seed, self.stream_base, self.num_envs, stream_ptr())
        return self.next_obs, self.reward, self.done

    # ---- reference-shaped numpy API ---------------------------------------------------------------
    @property
    def score(self):
        s = self._score.cpu().numpy()
        return float(s[0]) if self.num_envs == 1 else s

    @score.setter
    def score(self, v):
        self._score.fill_(float(v))

    def reset(self):
        return self.reset_device().cpu().numpy()

    def _to_device_action(self, action):
        a = np.asarray(action)
        if self.action_type == "continuous":
            return torch.as_tensor(a.reshape(self.num_envs, -1)[:, 0].astype(np.float32), device=self.device)
        return torch.as_tensor(a.reshape(self.num_envs).astype(np.int64), device=self.device)

    def step(self, action):
        next_obs, reward, done = self.step_device(self._to_device_action(action))
        n = self.num_envs
        packed = torch.cat([next_obs.reshape(n, -1), reward.view(n, 1), done.view(n, 1)], dim=1).cpu().numpy()
        d = self.state_size
        return (packed[:, :d].copy(), packed[:, d:d + 1].astype(np.float64), packed[:, d + 1:d + 2] > 0.5)

    def close(self):
        pass


class Cartpole(_Classic):
    kind_name = "cartpole"

    def __init__(self, action_type="discrete", **kwargs):
        self.action_type = action_type
        assert action_type in ("discrete", "continuous")
        self.action_size = 1 if action_type == "continuous" else 2
        super().__init__(**kwargs)


class Pendulum(_Classic):
    kind_name = "pendulum"

    def __init__(self, **kwargs):
        self.action_type = "continuous"
        self.action_size = 1
        super().__init__(**kwargs)


class MountainCar(_Classic):
    kind_name = "mountain_car"

    def __init__(self, **kwargs):
        self.action_type = "discrete"
        self.action_size = 3
        super().__init__(**kwargs)
