"""Prioritised replay with the sum-tree resident in HBM (jorldy/core/buffer/per_buffer.py:7-105).

Same attributes as the reference (`tree_size`, `first_leaf_index`, `tree_index`, `sum_tree`,
`max_priority`, `uniform_sample_prob`) and the same sequential update semantics, bit for bit
(see csrc/per_tree.cu), but a batch of B priority writes / B samples is one launch pair instead of
B python loops, and the B `.item()` device->host syncs of rainbow.py:230-231 / ape_x.py:111-112
disappear: `update_priorities` takes device tensors.
"""
import numpy as np
import torch

from ..dev import C, ptr, stream_ptr
from .replay_buffer import ReplayBuffer

_MAX_UPDATE = 4096    # per-launch batch bound (csrc/per_tree.cu keeps the batch in shared memory)


class PERBuffer(ReplayBuffer):
    def __init__(self, buffer_size, uniform_sample_prob=1e-3, device=None, seed=0):
        super().__init__(buffer_size, device=device)
        self.tree_size = (self.buffer_size * 2) - 1
        self.first_leaf_index = self.buffer_size - 1
        self._tree = torch.zeros(self.tree_size, dtype=torch.float64, device=self.device)
        self.tree_index = self.first_leaf_index
        self._max_priority = torch.ones(1, dtype=torch.float64, device=self.device)
        self.uniform_sample_prob = uniform_sample_prob
        self.seed = int(seed)
        self._sample_ctr = 0

    # ---- reference-visible attributes -----------------------------------------------------------
    @property
    def sum_tree(self):
        return self._tree.cpu().numpy()

    @property
    def max_priority(self):
        return float(self._max_priority.item())

    # ---- tree writes ----------------------------------------------------------------------------
    def _update(self, tree_idx, new_p, first_idx, n):
        """Sequential-semantics batch write, split into launches of <= _MAX_UPDATE entries."""
        s = stream_ptr()
        for off in range(0, n, _MAX_UPDATE):
            b = min(_MAX_UPDATE, n - off)
            ti = ptr(tree_idx[off:off + b]) if tree_idx is not None else 0
            npp = ptr(new_p[off:off + b]) if new_p is not None else 0
            fi = 0
            if tree_idx is None:
                fi = self.first_leaf_index + (first_idx - self.first_leaf_index + off) % self.buffer_size
            C.jb_per_update(ptr(self._tree), self.buffer_size, ti, fi, npp, ptr(self._max_priority),
                            ptr(self._max_priority), b, s)

    def store(self, transitions):
        """per_buffer.py:19-40: ring write + leaf := transition["priority"] if present else max_priority."""
        if self.first_store:
            self.check_dim({k: v for k, v in transitions[0].items()})
        has_p = "priority" in transitions[0]
        prio = None
        if has_p:
            vals = [t["priority"] for t in transitions]
            if torch.is_tensor(vals[0]):
                prio = torch.cat([v.reshape(-1).to(self.device, torch.float64) for v in vals])
            else:
                prio = torch.as_tensor(np.concatenate([np.asarray(v, dtype=np.float64).reshape(-1) for v in vals]),
                                       device=self.device)
            transitions = [{k: v for k, v in t.items() if k != "priority"} for t in transitions]
        n = self._write(transitions)
        self._update(None, prio, self.tree_index, n)
        self.tree_index = self.first_leaf_index + (self.tree_index - self.first_leaf_index + n) % self.buffer_size
        self.buffer_counter = min(self.buffer_counter + n, self.buffer_size)
        self.buffer_index = (self.buffer_index + n) % self.buffer_size

    def update_priority(self, new_priority, index):
        """Scalar API of the reference (per_buffer.py:42-48)."""
        idx = torch.tensor([int(index)], dtype=torch.int64, device=self.device)
        p = torch.tensor([float(np.asarray(new_priority).reshape(-1)[0])], dtype=torch.float64, device=self.device)
        self._update(idx, p, 0, 1)

    def update_priorities(self, indices, priorities):
        """Batched device API: indices int64 [B] (tree coordinates), priorities f64 [B]; applied in
        order, duplicates included (last write wins), like the python loop of the reference agents."""
        self._update(indices, priorities.to(torch.float64), 0, indices.shape[0])

    # ---- sampling -------------------------------------------------------------------------------
    def sample_device(self, beta, batch_size, u_a=None, u_b=None):
        """Returns (transitions dict of device tensors, weights f64 [B], tree indices int64 [B],
        stats f64 [4] = {sampled_p, mean_p, max raw weight, #uniform}).  u_a/u_b: injected uniforms.

        Sharded replay (Ape-X over G GPUs): when `self.shard_world > 1` this tree is one shard of
        a G-way replay; each rank draws its B/G share locally.  Item i of shard r is therefore drawn with
        probability (1/G) [(1-usp) p_i / total_r + usp / count_r] and the importance weight is
        ((1/N_global) / that)^beta, normalised by the GLOBAL max weight (per_buffer.py:88-94 over the union of
        the shards): two tiny all-reduces (2 x f64 SUM for N_global and the logged mean priority, 1 x f64 MAX)."""
        B = batch_size
        idx = torch.empty(B, dtype=torch.int64, device=self.device)
        w = torch.empty(B, dtype=torch.float64, device=self.device)
        p = torch.empty(B, dtype=torch.float64, device=self.device)
        stats = torch.empty(4, dtype=torch.float64, device=self.device)
        self._sample_ctr += 1
        sharded = getattr(self, "shard_world", 1) > 1
        g_total = g_count = shard_p = None
        if sharded:
            import torch.distributed as dist
            tot = torch.stack([self._tree[0], torch.tensor(float(self.buffer_counter), dtype=torch.float64, device=self.device)])
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            g_total = tot[0:1].contiguous()
            g_count = tot[1:2].to(torch.int64).contiguous()
            shard_p = torch.full((1,), 1.0 / self.shard_world, dtype=torch.float64, device=self.device)
        C.jb_per_sample(ptr(self._tree), self.buffer_size, self.buffer_counter, B, float(beta),
                        float(self.uniform_sample_prob), ptr(u_a), ptr(u_b), self.seed, self._sample_ctr, ptr(shard_p),
                        ptr(g_count), ptr(idx), ptr(w), ptr(p), ptr(stats), 0 if sharded else 1, stream_ptr())
        if sharded:
            import torch.distributed as dist
            wmax = stats[2:3].clone()
            dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
            C.jb_per_scale_weights(ptr(w), ptr(wmax), B, stream_ptr())
            stats[1] = g_total[0] / g_count[0].to(torch.float64)
        transitions = self.gather_device(idx - self.first_leaf_index)
        return transitions, w, idx, stats

    def sample(self, beta, batch_size):
        """Reference-shaped host API (per_buffer.py:70-101)."""
        assert float(self._tree[0].item()) > 0.0
        transitions, w, idx, stats = self.sample_device(beta, batch_size)
        st = stats.cpu().numpy()
        return (self._to_numpy(transitions), w.cpu().numpy(), idx.cpu().numpy(), float(st[0]), float(st[1]))
