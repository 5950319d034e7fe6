"""Additional golden fixtures minted from the unmodified reference (called by make_golden.py)."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


class _InjectedRandom:
    """Replaces np.random.uniform / randint inside PERBuffer.sample with pre-drawn numbers so the
    reference consumes exactly the randomness the oracle / CUDA path are given."""

    def __init__(self, u_a, u_b, counter):
        self.u_a, self.u_b, self.counter = u_a, u_b, counter
        self.calls = 0
        self.K = int(np.sum(u_a < 1e-3))

    def uniform(self, size=None, **kw):
        self.calls += 1
        if self.calls == 1:
            return self.u_a.copy()
        return self.u_b[self.K:].copy()

    def randint(self, high, size=None):
        return np.minimum((self.u_b[:self.K] * high).astype(np.int64), high - 1)


def per_scenario(seed, capacity, usp, with_prio=None):
    """Deterministic op sequence shared with the tests: list of ("store", n, prios|None) /
    ("update", idx[], p[]) / ("sample", beta, u_a, u_b)."""
    rs = np.random.RandomState(seed)
    ops = []
    total = 0
    if with_prio is None:
        with_prio = (seed % 2 == 0)     # a buffer holds either all-explicit (Ape-X) or no priorities
    for round_ in range(6):
        n = int(rs.randint(1, capacity))
        if with_prio:
            ops.append(("store", n, rs.uniform(0.01, 3.0, size=n)))
        else:
            ops.append(("store", n, None))
        total = min(total + n, capacity)
        B = int(rs.choice([8, 32, 64]))
        # duplicates on purpose (last write wins, each intermediate delta applied)
        leaf = rs.randint(0, total, size=B) + capacity - 1
        leaf[B // 2] = leaf[0]
        ops.append(("update", leaf.astype(np.int64), rs.uniform(0.0, 2.5, size=B) ** 0.6))
        u_a = rs.uniform(size=B)
        u_a[rs.randint(0, B, size=2)] = 1e-4          # force a couple of uniform slots (usp = 1e-3)
        ops.append(("sample", 0.4 + 0.1 * round_, u_a, rs.uniform(size=B)))
    return ops


PER_CASES = {"per_n10": (1, 10, 1e-3), "per_n1000": (2, 1000, 1e-3), "per_n4096": (3, 4096, 1e-3)}


def gen_per(buffer_mod, name, seed, capacity, usp):
    buf = buffer_mod.PERBuffer(capacity, usp)
    buf.first_store = False
    out = {}
    si = 0
    tr = {"state": np.zeros((1, 2), dtype=np.float32)}
    for op in per_scenario(seed, capacity, usp):
        if op[0] == "store":
            _, n, pr = op
            batch = []
            for i in range(n):
                t = dict(tr)
                if pr is not None:
                    t["priority"] = np.array([[pr[i]]])   # (1,1) like ape_x.py:194-196
                batch.append(t)
            buf.store(batch)
        elif op[0] == "update":
            for i, p in zip(op[1], op[2]):
                buf.update_priority(float(p), int(i))
        else:
            _, beta, u_a, u_b = op
            inj = _InjectedRandom(u_a, u_b, buf.buffer_counter)
            real_u, real_r = np.random.uniform, np.random.randint
            np.random.uniform, np.random.randint = inj.uniform, inj.randint
            try:
                _, w, idx, sp, mp = buf.sample(beta, len(u_a))
            finally:
                np.random.uniform, np.random.randint = real_u, real_r
            out[f"s{si}.idx"] = idx.astype(np.int64)
            out[f"s{si}.w"] = w.astype(np.float64)
            out[f"s{si}.stats"] = np.array([sp, mp])
            out[f"s{si}.root"] = np.float64(buf.sum_tree[0])
            si += 1
    out["final.tree"] = buf.sum_tree.copy() if capacity <= 1000 else buf.sum_tree[::7].copy()
    out["final.max_priority"] = np.float64(buf.max_priority)
    out["final.tree_index"] = np.int64(buf.tree_index)
    out["final.counter"] = np.int64(buf.buffer_counter)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "root", float(buf.sum_tree[0]), "max_p", buf.max_priority)


def main(agent_mod, buffer_mod, network_mod):
    for name, (seed, cap, usp) in PER_CASES.items():
        gen_per(buffer_mod, name, seed, cap, usp)
    try:
        import make_golden_dqn
        make_golden_dqn.main(agent_mod, buffer_mod, network_mod)
    except ImportError:
        pass
