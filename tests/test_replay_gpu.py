"""Replay ring rows (csrc/replay.cu) vs torch indexing, and the ReplayBuffer bookkeeping the reference's tests assert
(jorldy/test/core/buffer/test_replay_buffer.py:11-13,21-22): bit-exact for every dtype / row size the agents store."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,dtype", [((4, 84, 84), torch.uint8), ((4,), torch.float32), ((1,), torch.int64), ((3, 1), torch.float32),
                                         ((11,), torch.float32), ((), torch.float32), ((5,), torch.uint8)])
def test_store_gather_rows_match_torch_indexing(shape, dtype):
    from jorldy_b200.core.dev import C, ptr, stream_ptr
    cap, n = 257, 96
    g = torch.Generator().manual_seed(3)
    ring = torch.zeros((cap,) + shape, dtype=dtype, device="cuda")
    ref = ring.clone()
    mk = lambda k: (torch.randint(0, 255, (k,) + shape, generator=g).to(dtype) if dtype != torch.float32
                    else torch.randn((k,) + shape, generator=g)).cuda()
    batch = mk(n)
    pos = torch.randperm(cap, generator=g)[:n].cuda()
    row_bytes = int(np.prod(shape, dtype=np.int64)) * ring.element_size()
    C.jb_replay_store(ptr(ring), ptr(batch), ptr(pos), n, row_bytes, stream_ptr())
    ref.index_copy_(0, pos, batch)
    assert torch.equal(ring, ref)
    idx = torch.randint(0, cap, (64,), generator=g).cuda()           # with repeats
    out = torch.empty((64,) + shape, dtype=dtype, device="cuda")
    C.jb_replay_gather(ptr(ring), ptr(idx), 64, row_bytes, ptr(out), stream_ptr())
    assert torch.equal(out, ref.index_select(0, idx))


def test_replay_buffer_wraps_like_the_reference():
    from jorldy_b200.core.buffer import ReplayBuffer
    buf = ReplayBuffer(10, device="cuda")
    mk = lambda i, n: {"state": np.full((n, 4), i, np.float32), "action": np.full((n, 1), i, np.int64),
                       "reward": np.full((n, 1), float(i)), "done": np.zeros((n, 1), bool), "next_state": np.full((n, 4), i + 0.5, np.float32)}
    buf.store([mk(1, 4)]); buf.store([mk(2, 4)])
    assert buf.buffer_index == 8 and buf.buffer_counter == 8 and buf.size == 8
    buf.store([mk(3, 4)])                                             # wraps: rows 8, 9, 0, 1
    assert buf.buffer_index == 2 and buf.buffer_counter == 10
    st = buf.fields["state"].cpu().numpy()[:, 0]
    np.testing.assert_array_equal(st, [3, 3, 1, 1, 2, 2, 2, 2, 3, 3])
    s = buf.sample(5)
    assert s["state"].shape == (5, 4) and s["reward"].dtype == np.float64 and s["done"].dtype == np.bool_
    assert np.all(s["next_state"][:, 0] == s["state"][:, 0] + 0.5)
