"""oracle/per.py pinned to the UNMODIFIED reference PERBuffer (tests/golden/per_*.npz): tree
contents bit-exact, sampled indices exact, IS weights exact (same numpy pow).  CPU only."""
import numpy as np
import pytest

import make_golden_more as MG
from helpers import load_golden
from oracle.per import SumTree


def replay_oracle(seed, capacity, usp):
    t = SumTree(capacity, usp)
    samples = []
    for op in MG.per_scenario(seed, capacity, usp):
        if op[0] == "store":
            t.store(op[1], op[2])
        elif op[0] == "update":
            for i, p in zip(op[1], op[2]):
                t.update(float(p), int(i))
        else:
            samples.append(t.sample(op[1], op[2], op[3]) + (t.tree[0],))
    return t, samples


@pytest.mark.parametrize("name", list(MG.PER_CASES.keys()))
def test_per_oracle_matches_reference(name):
    seed, cap, usp = MG.PER_CASES[name]
    gold = load_golden(name)
    t, samples = replay_oracle(seed, cap, usp)
    for si, (idx, w, sp, mp, root) in enumerate(samples):
        assert np.array_equal(idx, gold[f"s{si}.idx"])
        assert np.array_equal(w, gold[f"s{si}.w"])
        assert np.array_equal(np.array([sp, mp]), gold[f"s{si}.stats"])
        assert root == float(gold[f"s{si}.root"])
    tree = t.tree if cap <= 1000 else t.tree[::7]
    assert np.array_equal(tree, gold["final.tree"])
    assert t.max_priority == float(gold["final.max_priority"])
    assert t.tree_index == int(gold["final.tree_index"])
    assert t.counter == int(gold["final.counter"])
