"""CPU restatement of the synthetic Atari-shaped generator (jorldy_b200/csrc/env_frames.cu), numpy +
oracle.philox — bit-exact with the CUDA env.  It exists because the real ALE path of
jorldy/core/env/atari.py is deliberately replaced (north star); what is pinned against the reference is
the CONTRACT: state (N,4,84,84) uint8 newest-last (atari.py:56-61,157-160), first state tiled x4
(:112), reward in {-1,0,1} (:151-152)."""
import numpy as np

from . import philox

FRAME = 84 * 84


def gen_frame(seed, stream, fidx):
    q = np.arange(FRAME // 16, dtype=np.uint64)
    a, b, c, d = philox.philox4x32(seed, np.uint64(stream), np.uint64(fidx) * np.uint64(512) + q)
    words = np.stack([a, b, c, d], axis=1).astype("<u4")          # uint4 {x,y,z,w} little-endian in memory
    return words.reshape(-1).view(np.uint8).reshape(84, 84)


def events(seed, stream, fidx):
    a, b, _, _ = philox.philox4x32(seed, np.uint64(stream), np.uint64(fidx) * np.uint64(512) + np.uint64(511))
    u0, u1 = float(philox.u01_float(a)), float(philox.u01_float(b))
    reward = 1.0 if u0 < np.float32(0.05) else (-1.0 if u0 < np.float32(0.10) else 0.0)
    return reward, u1 < np.float32(0.001)


class FramesBatch:
    def __init__(self, n, seed=0, stream_base=0, auto_reset=True):
        self.n, self.seed, self.stream_base, self.auto_reset = n, seed, stream_base, auto_reset
        self.obs = np.zeros((n, 4, 84, 84), dtype=np.uint8)
        self.fcount = np.zeros(n, dtype=np.int64)

    def reset(self):
        for e in range(self.n):
            self.obs[e, :] = gen_frame(self.seed, self.stream_base + e, self.fcount[e])[None]
            self.fcount[e] += 1
        return self.obs.copy()

    def step(self):
        reward = np.zeros(self.n, dtype=np.float32)
        done = np.zeros(self.n, dtype=bool)
        next_obs = np.zeros_like(self.obs)
        for e in range(self.n):
            f = self.fcount[e]
            self.obs[e, :3] = self.obs[e, 1:].copy()
            self.obs[e, 3] = gen_frame(self.seed, self.stream_base + e, f)
            reward[e], done[e] = events(self.seed, self.stream_base + e, f)
            next_obs[e] = self.obs[e]
            self.fcount[e] = f + 1
            if done[e] and self.auto_reset:
                self.obs[e, :] = gen_frame(self.seed, self.stream_base + e, f + 1)[None]
                self.fcount[e] = f + 2
        return next_obs, reward, done
