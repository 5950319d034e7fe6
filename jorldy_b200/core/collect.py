"""Resident rollout collection: act -> env.step -> write transition, for every env at once, with
no host hop.  Replaces the per-actor loop of jorldy/manager/distributed_manager.py:76-92
(`Actor.run`) and its gather in run_mode.py:180-187: the N ray actors become N rows of one batched
launch sequence, and the whole T-step rollout is captured in ONE CUDA graph (per-row Philox
counters and the env's own episode counters advance on the device, so every replay draws fresh
randomness).
"""
import torch

from .buffer import DeviceRollout


class RolloutCollector:
    def __init__(self, env, agent, n_step=None, use_cuda_graph=True):
        self.env, self.agent = env, agent
        self.T = n_step or agent.n_step
        self.rollout = DeviceRollout(env.num_envs, self.T, env.state_size, env.action_size, env.action_type,
                                     device=agent.device)
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        # kernels of OUR library per env step: mlp_in_fwd + gemm + heads_fwd + act + env_step (the
        # rollout-row copies are torch plumbing and not counted)
        self.launches_per_collect = 5 * self.T
        env.reset_device()

    def _collect_eager(self):
        env, agent, ro = self.env, self.agent, self.rollout
        ro.clear()
        for t in range(self.T):
            ro.state[:, t].copy_(env.obs)                      # state acted on (pre-step observation)
            action = agent.act_device(env.obs, training=True)
            next_obs, reward, done = env.step_device(action)   # env.obs <- post-reset observation
            ro.t = t
            ro.write_after_step(action, reward, done, next_obs)

    def collect(self):
        """Fills self.rollout with T steps of all envs; returns it."""
        if not self.use_cuda_graph:
            self._collect_eager()
            return self.rollout
        if self._graph is None:
            # warm-up (allocates workspaces) on a side stream, then capture the T-step sequence
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._collect_eager()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._collect_eager()
            self._graph = g
        self._graph.replay()
        self.rollout.t = self.T
        return self.rollout
