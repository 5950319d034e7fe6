"""CUDA sum-tree vs oracle/per.py + the reference golden files.

BIT-EXACT: every tree node (f64), sampled tree indices, max_priority, ring bookkeeping.
IS weights: within 4 ulp (f64) — CUDA pow() vs numpy pow()."""
import numpy as np
import pytest
import torch

import make_golden_more as MG
from helpers import load_golden
from oracle.per import SumTree

pytestmark = pytest.mark.gpu


def _mk(capacity, usp):
    from jorldy_b200.core.buffer import PERBuffer
    buf = PERBuffer(capacity, usp, device="cuda")
    buf.first_store = False
    return buf


@pytest.mark.parametrize("name", list(MG.PER_CASES.keys()))
def test_per_cuda_matches_reference_golden(name):
    seed, cap, usp = MG.PER_CASES[name]
    gold = load_golden(name)
    buf = _mk(cap, usp)
    ora = SumTree(cap, usp)
    si = 0
    for op in MG.per_scenario(seed, cap, usp):
        if op[0] == "store":
            _, n, pr = op
            tr = {"state": np.zeros((n, 2), dtype=np.float32), "reward": np.zeros((n, 1))}
            if pr is not None:
                tr["priority"] = pr.reshape(n, 1)
            buf.store([tr])
            ora.store(n, pr)
        elif op[0] == "update":
            buf.update_priorities(torch.as_tensor(op[1], device="cuda"), torch.as_tensor(op[2], device="cuda"))
            for i, p in zip(op[1], op[2]):
                ora.update(float(p), int(i))
        else:
            _, beta, u_a, u_b = op
            _, w, idx, stats = buf.sample_device(beta, len(u_a), torch.as_tensor(u_a, device="cuda"),
                                                 torch.as_tensor(u_b, device="cuda"))
            assert np.array_equal(idx.cpu().numpy(), gold[f"s{si}.idx"])
            np.testing.assert_allclose(w.cpu().numpy(), gold[f"s{si}.w"], rtol=1e-15, atol=0)
            np.testing.assert_allclose(stats.cpu().numpy()[:2], gold[f"s{si}.stats"], rtol=1e-15)
            assert float(buf.sum_tree[0]) == float(gold[f"s{si}.root"])
            si += 1
        assert np.array_equal(buf.sum_tree, ora.tree), op[0]
    tree = buf.sum_tree if cap <= 1000 else buf.sum_tree[::7]
    assert np.array_equal(tree, gold["final.tree"])
    assert buf.max_priority == float(gold["final.max_priority"])
    assert buf.tree_index == int(gold["final.tree_index"])
    assert buf.buffer_counter == int(gold["final.counter"])


def test_per_reference_contract():
    """jorldy/test/core/buffer/test_per_buffer.py:6-56 re-run against the CUDA buffer."""
    from jorldy_b200.core.buffer import PERBuffer
    mock_transition = [{
        "state": np.random.random((1, 4)), "action": np.random.random((1, 3)), "reward": np.random.random((1, 1)),
        "next_state": np.random.random((1, 4)), "done": np.random.random((1, 1)) < 0.5,
        "multi_modal": [np.random.random((1, 3, 8, 8)), np.random.random((1, 4))], "seq": np.random.random((1, 3, 4)),
    }]
    buffer_size = 10
    memory = PERBuffer(buffer_size=buffer_size, uniform_sample_prob=1e-3)
    assert memory.buffer_size == buffer_size and memory.tree_size == (buffer_size * 2) - 1
    assert memory.buffer_index == 0 and memory.tree_index == buffer_size - 1 and memory.size == 0
    for _ in range(15):
        memory.store(mock_transition)
    assert memory.buffer_index == 15 % buffer_size
    assert memory.tree_index == buffer_size - 1 + (15 % buffer_size)
    assert memory.size == min(buffer_size, 15)
    tr, w, indices, sampled_p, mean_p = memory.sample(beta=0.4, batch_size=8)
    assert isinstance(tr, dict) and isinstance(w, np.ndarray) and (w <= 1.0).all()
    assert isinstance(indices, np.ndarray) and (indices >= buffer_size - 1).all()
    assert isinstance(sampled_p, float) and isinstance(mean_p, float)
    for key, val in tr.items():
        if isinstance(val, list):
            for i, v in enumerate(val):
                assert v.shape == (8, *mock_transition[0][key][i].shape[1:])
        else:
            assert val.shape == (8, *mock_transition[0][key].shape[1:])
    memory.update_priority(2.0, buffer_size - 1 + (buffer_size // 2))
    assert memory.max_priority == 2.0
    assert memory.sum_tree[buffer_size - 1 + (buffer_size // 2)] == 2.0


def test_per_large_tree_descent_exact():
    """1 M-slot tree (BASELINE config #3 size): indices from 4096 prioritised draws equal the oracle's."""
    cap = 1_000_000
    buf = _mk(cap, 1e-3)
    ora = SumTree(cap, 1e-3)
    rs = np.random.RandomState(9)
    n = 20000
    pr = rs.uniform(0.01, 2.0, size=n)
    buf.store([{"state": np.zeros((n, 1), dtype=np.float32), "reward": np.zeros((n, 1)), "priority": pr.reshape(n, 1)}])
    ora.store(n, pr)
    assert np.array_equal(buf.sum_tree, ora.tree)
    u_a = rs.uniform(size=4096); u_b = rs.uniform(size=4096)
    _, w, idx, stats = buf.sample_device(0.5, 4096, torch.as_tensor(u_a, device="cuda"), torch.as_tensor(u_b, device="cuda"))
    oidx, ow, sp, mp = ora.sample(0.5, u_a, u_b)
    assert np.array_equal(idx.cpu().numpy(), oidx)
    np.testing.assert_allclose(w.cpu().numpy(), ow, rtol=1e-15)
