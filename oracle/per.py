"""CPU oracle of the prioritised-replay sum-tree (jorldy/core/buffer/per_buffer.py), numpy f64.

update()   per_buffer.py:42-54  (sequential, incremental delta up the array heap)
search()   per_buffer.py:56-68  (`num <= left` goes left, else subtract and go right)
sample()   per_buffer.py:70-101 with the random draws injected: u_a decides uniform-vs-prioritised
           per slot (count only — uniform slots come first in the output, :84), u_b[s] is the ring
           draw (floor(u*counter), i.e. np.random.randint) for uniform slots and the target
           fraction (u * tree[0]) for prioritised slots.
"""
import numpy as np


class SumTree:
    def __init__(self, capacity, uniform_sample_prob=1e-3):
        self.capacity = capacity
        self.tree_size = 2 * capacity - 1
        self.first_leaf = capacity - 1
        self.tree = np.zeros(self.tree_size)
        self.tree_index = self.first_leaf
        self.max_priority = 1.0
        self.counter = 0
        self.usp = uniform_sample_prob

    def update(self, new_priority, index):
        delta = new_priority - self.tree[index]
        self.tree[index] = new_priority
        while index > 0:
            index = (index - 1) // 2
            self.tree[index] += delta
        self.max_priority = max(self.max_priority, new_priority)

    def store(self, n, priorities=None):
        for i in range(n):
            p = self.max_priority if priorities is None else float(priorities[i])
            self.update(p, self.tree_index)
            self.tree_index += 1
            if self.tree_index == self.tree_size:
                self.tree_index = self.first_leaf
            self.counter = min(self.counter + 1, self.capacity)

    def search(self, num):
        index = 0
        while index < self.first_leaf:
            left = 2 * index + 1
            if num <= self.tree[left]:
                index = left
            else:
                num -= self.tree[left]
                index = left + 1
        return index

    def sample(self, beta, u_a, u_b):
        B = len(u_a)
        K = int(np.sum(np.asarray(u_a) < self.usp))
        idx = []
        for s in range(B):
            if s < K:
                r = min(int(u_b[s] * self.counter), self.counter - 1)
                idx.append(r + self.first_leaf)
            else:
                idx.append(self.search(u_b[s] * self.tree[0]))
        idx = np.asarray(idx)
        pri = self.tree[idx]
        uniform_prob = 1.0 / self.counter
        prio_prob = pri / self.tree[0]
        sample_prob = (1.0 - self.usp) * prio_prob + self.usp * uniform_prob
        w = (uniform_prob / sample_prob) ** beta
        w = w / np.max(w)
        return idx, w, float(np.mean(pri)), float(self.tree[0] / self.counter)
