"""Synthetic Atari-shaped env: CUDA generator bit-exact vs oracle/frames.py, and the observation
contract of jorldy/core/env/atari.py (shape/dtype/layout, first-frame tiling, reward set)."""
import numpy as np
import pytest
import torch

from oracle.frames import FramesBatch

pytestmark = pytest.mark.gpu


def test_frames_bit_exact_vs_oracle():
    from jorldy_b200.core import Env
    n = 5
    env = Env("breakout", num_envs=n, seed=11, id=2)
    ref = FramesBatch(n, seed=11, stream_base=2 << 32)
    obs = env.reset()
    assert obs.shape == (n, 4, 84, 84) and obs.dtype == np.uint8
    robs = ref.reset()
    assert np.array_equal(obs, robs)
    assert all(np.array_equal(obs[:, 0], obs[:, k]) for k in range(1, 4))        # atari.py:112 tile
    for t in range(12):
        ns, r, d = env.step(np.zeros((n, 1), dtype=np.int64))
        rns, rr, rd = ref.step()
        assert np.array_equal(ns, rns) and np.array_equal(r.reshape(-1).astype(np.float32), rr)
        assert np.array_equal(d.reshape(-1), rd)
        assert np.array_equal(env.obs.cpu().numpy(), ref.obs)
        assert set(np.unique(r)).issubset({-1.0, 0.0, 1.0})
        assert np.array_equal(ns[:, :3], obs[:, 1:]) or t > 0
        obs = ns


def test_frames_reset_on_done_tiles_first_frame():
    from jorldy_b200.core import Env
    env = Env("synthetic_atari", num_envs=64, seed=3)
    env.reset()
    seen = 0
    for t in range(400):
        ns, r, d = env.step_device(None)
        dd = d.cpu().numpy() > 0.5
        if dd.any():
            o = env.obs.cpu().numpy()[dd]
            assert all(np.array_equal(o[:, 0], o[:, k]) for k in range(1, 4))
            seen += int(dd.sum())
    assert seen > 0 and env.stats[0].item() == seen


def test_single_env_reference_shapes():
    """jorldy/test/core/env/utils.py contract at N=1."""
    from jorldy_b200.core import Env
    env = Env("pong")
    s = env.reset()
    assert s.shape == (1, 4, 84, 84) and env.state_size == [4, 84, 84] and env.action_size == 6
    ns, r, d = env.step(np.array([[1]]))
    assert ns.shape == (1, 4, 84, 84) and r.shape == (1, 1) and d.shape == (1, 1)
