"""Synthetic continuous-control env with the dimensions of a MuJoCo task (csrc/env_synth.cu), registered under the
reference's MuJoCo env names (jorldy/core/env/mujoco.py:61-105: `hopper`, `half_cheetah`, ...) so
`--config config.ppo.mujoco --env.name hopper` resolves.  MuJoCo itself is a third-party simulator that cannot be
installed here; per the north star the physics is replaced by a synthetic generator with the SAME observation /
action dtype, layout and dimensions (Hopper-v3: obs 11, act 3, 1000-step TimeLimit), see the kernel's header."""
import numpy as np
import torch

from ..dev import C, ptr, require_cuda, stream_ptr
from .base import BaseEnv

_DIMS = {"hopper": (11, 3), "half_cheetah": (17, 6), "walker": (17, 6), "inverted_double_pendulum": (11, 1),
         "reacher": (11, 2), "synthetic_control": (11, 3)}


def synth_weights(D, A, seed=0):
    """The fixed dynamics matrices (shared with oracle/classic_control.py::SyntheticControlBatch)."""
    rs = np.random.RandomState(1000 + seed)
    Ws = (rs.standard_normal((D, D)) * (0.9 / np.sqrt(D))).astype(np.float32)
    Wa = (rs.standard_normal((D, A)) * 0.5).astype(np.float32)
    return Ws, Wa


class SyntheticControl(BaseEnv):
    action_type = "continuous"

    def __init__(self, name="hopper", num_envs=1, seed=0, id=0, device=None, auto_reset=None, p_done=1e-3,
                 max_steps=1000, render=False, train_mode=True, **kwargs):
        self.device = require_cuda(device)
        self.name = name
        self.state_size, self.action_size = _DIMS[name]
        self.num_envs, self.seed = int(num_envs), int(seed)
        self.id = int(id) if id is not None else 0
        self.stream_base = self.id << 32
        self.auto_reset = (self.num_envs > 1) if auto_reset is None else bool(auto_reset)
        self.p_done, self.max_steps = float(p_done), int(max_steps)
        n, dev, D = self.num_envs, self.device, self.state_size
        Ws, Wa = synth_weights(D, self.action_size, 0)
        self.Ws, self.Wa = torch.from_numpy(Ws).to(dev), torch.from_numpy(Wa).to(dev)
        self.obs = torch.zeros(n, D, dtype=torch.float32, device=dev)
        self.next_obs = torch.zeros(n, D, dtype=torch.float32, device=dev)
        self.elapsed = torch.zeros(n, dtype=torch.int32, device=dev)
        self.episode = torch.zeros(n, dtype=torch.int64, device=dev)
        self.tcount = torch.zeros(n, dtype=torch.int64, device=dev)
        self._score = torch.zeros(n, dtype=torch.float32, device=dev)
        self.reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self.done = torch.zeros(n, dtype=torch.float32, device=dev)
        self.stats = torch.zeros(2, dtype=torch.float32, device=dev)

    def reset_device(self, mask=None):
        C.jb_env_synth_reset(ptr(self.obs), ptr(self.elapsed), ptr(self.episode), ptr(self._score), self.seed,
                             self.stream_base, self.num_envs, self.state_size, stream_ptr())
        return self.obs

    def step_device(self, action):
        """action f32 [N, A] in (-1, 1) (Hopper's action bounds are [-1, 1]: the reference's rescale is the identity)."""
        assert action.dtype == torch.float32 and action.is_contiguous()
        C.jb_env_synth_step(ptr(self.obs), ptr(self.elapsed), ptr(self.episode), ptr(self.tcount), ptr(self._score),
                            ptr(action), ptr(self.Ws), ptr(self.Wa), ptr(self.next_obs), ptr(self.reward), ptr(self.done),
                            ptr(self.stats), int(self.auto_reset), self.max_steps, self.p_done, self.seed, self.stream_base,
                            self.num_envs, self.state_size, self.action_size, stream_ptr())
        return self.next_obs, self.reward, self.done

    @property
    def score(self):
        s = self._score.cpu().numpy()
        return float(s[0]) if self.num_envs == 1 else s

    def reset(self):
        return self.reset_device().cpu().numpy()

    def step(self, action):
        a = torch.as_tensor(np.asarray(action, dtype=np.float32).reshape(self.num_envs, self.action_size), device=self.device)
        next_obs, reward, done = self.step_device(a.contiguous())
        n = self.num_envs
        return (next_obs.cpu().numpy(), reward.view(n, 1).cpu().numpy().astype(np.float64),
                done.view(n, 1).cpu().numpy() > 0.5)

    def close(self):
        pass
