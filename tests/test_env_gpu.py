"""Batched CartPole kernel vs the CPU restatement (oracle/classic_control.py) on identical seeds.

Bit-exact pieces: Philox reset draws, elapsed/episode bookkeeping, reward/done flags.
Tolerance pieces: f64 physics state within 1e-12 abs per step (CUDA sin/cos are <= 2 ulp from
glibc's), f32 observations within 1 ulp."""
import numpy as np
import pytest
import torch

from oracle.classic_control import CartPoleBatch

pytestmark = pytest.mark.gpu


def test_cartpole_reset_bit_exact():
    from jorldy_b200.core import Env
    env = Env("cartpole", num_envs=257, seed=123, id=3)
    obs = env.reset()
    ref = CartPoleBatch(257, seed=123, stream_base=3 << 32)
    ref_obs = ref.reset()
    assert np.array_equal(env.phys.cpu().numpy(), ref.phys)
    assert np.array_equal(obs, ref_obs)


def test_cartpole_rollout_matches_oracle():
    from jorldy_b200.core import Env
    n = 512
    env = Env("cartpole", num_envs=n, seed=7)
    ref = CartPoleBatch(n, seed=7)
    env.reset(); ref.reset()
    rs = np.random.RandomState(0)
    for t in range(300):
        a = rs.randint(0, 2, size=(n, 1))
        # re-synchronise the f64 state each step so the comparison is one-step (no chaotic drift)
        ref.phys = env.phys.cpu().numpy().copy()
        ref.elapsed = env.elapsed.cpu().numpy().copy()
        ref.episode = env.episode.cpu().numpy().copy()
        ns, r, d = env.step(a)
        rns, rr, rd = ref.step(a)
        np.testing.assert_allclose(ns, rns, rtol=0, atol=1e-6)
        assert np.array_equal(d.reshape(-1), rd), t
        np.testing.assert_array_equal(r.reshape(-1).astype(np.float32), rr)
        np.testing.assert_allclose(env.phys.cpu().numpy(), ref.phys, rtol=0, atol=1e-12)
        assert np.array_equal(env.elapsed.cpu().numpy(), ref.elapsed)
        assert np.array_equal(env.episode.cpu().numpy(), ref.episode)


def test_cartpole_reference_shapes_single_env():
    """jorldy/test/core/env/utils.py:7-16 contract at num_envs=1: (1,4) state, (1,1) reward/done."""
    from jorldy_b200.core import Env
    env = Env("cartpole")
    state = env.reset()
    assert state.shape == (1, 4)
    for _ in range(10):
        ns, r, d = env.step(np.random.randint(0, 2, size=(1, 1)))
        assert ns.shape == (1, 4) and r.shape == (1, 1) and d.shape == (1, 1)
        assert d.dtype == np.bool_
        if d:
            assert r[0, 0] == -1
            env.reset()
        else:
            assert abs(r[0, 0] - 0.1) < 1e-7
    env.close()


def test_cartpole_time_limit_and_score():
    from jorldy_b200.core import Env
    env = Env("cartpole", num_envs=4, seed=1)
    env.reset()
    env.max_steps = 5
    a = torch.zeros(4, dtype=torch.int64, device="cuda")
    for t in range(5):
        a = 1 - a
        ns, r, d = env.step_device(a)
    assert torch.all(d == 1.0) and torch.all(r == -1.0)
    assert torch.all(env.elapsed == 0)
    assert env.stats[0].item() == 4 and env.stats[1].item() == 20.0


def _sync(ref, env):
    ref.phys = env.phys.cpu().numpy().copy()
    ref.elapsed = env.elapsed.cpu().numpy().copy()
    ref.episode = env.episode.cpu().numpy().copy()


def test_pendulum_matches_oracle():
    from jorldy_b200.core import Env
    from oracle.classic_control import PendulumBatch
    n = 256
    env = Env("pendulum", num_envs=n, seed=3)
    ref = PendulumBatch(n, seed=3)
    obs = env.reset(); robs = ref.reset()
    assert np.array_equal(env.phys.cpu().numpy(), ref.phys)
    np.testing.assert_allclose(obs, robs, rtol=0, atol=1e-6)
    assert env.state_size == 3 and env.action_size == 1 and env.action_type == "continuous"
    rs = np.random.RandomState(1)
    for t in range(230):
        a = rs.uniform(-1, 1, size=(n, 1)).astype(np.float32)
        _sync(ref, env)
        ns, r, d = env.step(a)
        rns, rr, rd = ref.step(a)
        np.testing.assert_allclose(ns, rns, rtol=0, atol=2e-6)
        np.testing.assert_allclose(r.reshape(-1), rr, rtol=1e-6, atol=1e-6)
        assert np.array_equal(d.reshape(-1), rd)
        np.testing.assert_allclose(env.phys.cpu().numpy(), ref.phys, rtol=0, atol=1e-11)


def test_mountain_car_matches_oracle():
    from jorldy_b200.core import Env
    from oracle.classic_control import MountainCarBatch
    n = 256
    env = Env("mountain_car", num_envs=n, seed=4)
    ref = MountainCarBatch(n, seed=4)
    obs = env.reset(); robs = ref.reset()
    assert np.array_equal(obs, robs)
    assert env.state_size == 2 and env.action_size == 3
    rs = np.random.RandomState(2)
    for t in range(230):
        a = rs.randint(0, 3, size=(n, 1))
        _sync(ref, env)
        ns, r, d = env.step(a)
        rns, rr, rd = ref.step(a)
        np.testing.assert_allclose(ns, rns, rtol=0, atol=1e-7)
        assert np.array_equal(d.reshape(-1), rd) and np.all(r == -1.0)
        np.testing.assert_allclose(env.phys.cpu().numpy(), ref.phys, rtol=0, atol=1e-14)


def test_synthetic_control_matches_oracle():
    """Hopper-dimension synthetic generator (csrc/env_synth.cu) vs its numpy restatement: reset draws and done flags
    bit-exact (Philox), observations / rewards within 2e-6 (tanhf, Box-Muller logf / cospif differ by an ulp)."""
    from jorldy_b200.core import Env
    from jorldy_b200.core.env.synth import synth_weights
    from oracle.classic_control import SyntheticControlBatch
    n, D, A = 512, 11, 3
    env = Env("hopper", num_envs=n, seed=7, id=2, device="cuda", p_done=0.05, max_steps=20)
    assert env.state_size == 11 and env.action_size == 3 and env.action_type == "continuous"
    Ws, Wa = synth_weights(D, A, 0)
    ref = SyntheticControlBatch(n, D, A, seed=7, stream_base=2 << 32, p_done=0.05, max_steps=20, Ws=Ws, Wa=Wa)
    o = env.reset_device().cpu().numpy()
    np.testing.assert_array_equal(o, ref.reset())
    rs = np.random.RandomState(0)
    n_done = 0
    for t in range(40):
        a = np.tanh(rs.standard_normal((n, A))).astype(np.float32)
        nobs, r, d = env.step_device(torch.from_numpy(a).cuda())
        rn, rr, rd = ref.step(a)
        np.testing.assert_array_equal(d.cpu().numpy() > 0.5, rd, err_msg=f"done step {t}")
        np.testing.assert_allclose(nobs.cpu().numpy(), rn, rtol=0, atol=2e-6, err_msg=f"next_obs step {t}")
        np.testing.assert_allclose(r.cpu().numpy(), rr, rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(env.obs.cpu().numpy(), ref.obs, rtol=0, atol=2e-6)     # post auto-reset observation
        ref.obs = env.obs.cpu().numpy().copy()          # re-synchronise so ulp differences do not accumulate
        n_done += int(rd.sum())
    assert n_done > 100                                  # Bernoulli and TimeLimit terminations both exercised
    assert int(env.elapsed.max().item()) < 20
