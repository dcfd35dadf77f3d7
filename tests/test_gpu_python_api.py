"""GPU: the reference-facing Python API (envpool_b200.make / reset / step / send / recv,
gymnasium- and dm-style adapters) on top of the pybind modules, checked against the oracle.
Mirrors what the reference's own Python tests exercise (classic_control_test.py:34-57
determinism; dummy_py_envpool_test.py:59-132 key lists and _send/_recv plumbing)."""
import numpy as np
import pytest

from helpers import random_actions

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ep(engine_built):
    import torch

    assert torch.cuda.is_available()
    import envpool_b200

    return envpool_b200


def test_gymnasium_api_matches_oracle(ep):
    from oracle.oracle_lib import OraclePool

    N = 64
    env = ep.make("CartPole-v1", env_type="gymnasium", num_envs=N, seed=3)
    orc = OraclePool("CartPole", N, seed=3, max_episode_steps=500)
    obs, info = env.reset()
    w = orc.reset()
    np.testing.assert_allclose(obs, w["obs"], rtol=0, atol=1e-6)
    assert set(info) >= {"env_id", "players", "elapsed_step"}
    np.testing.assert_array_equal(info["env_id"], np.arange(N))
    np.testing.assert_array_equal(info["players"]["env_id"], np.arange(N))
    rng = np.random.default_rng(0)
    for _ in range(100):
        a = random_actions("CartPole", rng, (N,))
        obs, rew, term, trunc, info = env.step(a)
        w = orc.step(a)
        np.testing.assert_allclose(obs, w["obs"], rtol=0, atol=1e-6)
        np.testing.assert_array_equal(rew, w["reward"])
        np.testing.assert_array_equal(trunc, w["trunc"])
        np.testing.assert_array_equal(term, w["done"] & ~w["trunc"])
        np.testing.assert_array_equal(info["elapsed_step"], w["elapsed_step"])
    assert len(env) == N and not env.is_async
    assert env.action_space.n == 2 and env.observation_space.shape == (4,)


def test_dm_api_and_send_recv(ep):
    from oracle.oracle_lib import OraclePool

    N = 32
    env = ep.make_dm("FrozenLake-v1", num_envs=N, seed=5)
    orc = OraclePool("FrozenLake", N, seed=5, max_episode_steps=100, iopt=4)
    env.async_reset()
    ts = env.recv()
    w = orc.reset()
    assert (ts.step_type == 0).all() and ts.first().all()
    np.testing.assert_array_equal(ts.observation.obs, w["obs"])
    np.testing.assert_array_equal(ts.observation.env_id, np.arange(N))
    rng = np.random.default_rng(1)
    for _ in range(60):
        a = random_actions("FrozenLake", rng, (N,))
        env.send(a)
        ts = env.recv()
        w = orc.step(a)
        np.testing.assert_array_equal(ts.observation.obs, w["obs"])
        np.testing.assert_array_equal(ts.reward, w["reward"])
        np.testing.assert_array_equal(ts.discount, w["discount"])
        np.testing.assert_array_equal(ts.step_type, w["step_type"])


def test_partial_env_id_step_and_returned_arrays_stay_valid(ep):
    N = 128
    env = ep.make_gym("Pendulum-v1", num_envs=N, seed=0)
    obs0, _ = env.reset()
    keep = obs0.copy()
    ids = np.array([5, 17, 3, 90], dtype=np.int32)
    a = np.zeros((4, 1), dtype=np.float32)
    obs, rew, term, trunc, info = env.step(a, ids)
    assert obs.shape == (4, 3)
    np.testing.assert_array_equal(info["env_id"], ids)
    rng = np.random.default_rng(2)
    for _ in range(20):   # more batches than pinned slabs would hold if they were reused
        env.step(random_actions("Pendulum", rng, (N,)))
    np.testing.assert_array_equal(obs0, keep)   # zero-copy batch still intact


def test_determinism_and_seed_sensitivity(ep):
    """classic_control_test.py:34-57: same seed => identical, different seed => different."""
    N, T = 16, 200
    rng = np.random.default_rng(3)
    acts = random_actions("Acrobot", rng, (T, N))
    runs = []
    for seed in (0, 0, 1):
        env = ep.make_gym("Acrobot-v1", num_envs=N, seed=seed)
        obs, _ = env.reset()
        traj = [obs]
        for t in range(T):
            traj.append(env.step(acts[t])[0])
        runs.append(np.stack(traj))
    np.testing.assert_array_equal(runs[0], runs[1])
    assert np.abs(runs[0] - runs[2]).sum() > 0
    space = ep.make_spec("Acrobot-v1").observation_space
    assert np.all(runs[0] >= space.low - 1e-6) and np.all(runs[0] <= space.high + 1e-6)


def test_step_device_zero_copy(ep):
    import torch

    N = 4096
    env = ep.make_gym("CartPole-v1", num_envs=N, seed=9)
    ref = ep.make_gym("CartPole-v1", num_envs=N, seed=9)
    out = env.reset_device()
    obs_ref, _ = ref.reset()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out["obs"].cpu().numpy(), obs_ref)
    rng = np.random.default_rng(4)
    for _ in range(30):
        a = random_actions("CartPole", rng, (N,))
        out = env.step_device(torch.from_numpy(a).cuda())
        o, r, term, trunc, info = ref.step(a)
        torch.cuda.synchronize()
        assert out["obs"].is_cuda
        np.testing.assert_array_equal(out["obs"].cpu().numpy(), o)
        np.testing.assert_array_equal(out["done"].cpu().numpy(), term | trunc)


def test_async_mode_send_recv(ep):
    """batch_size < num_envs (async_envpool.h:93-97): the loop of benchmark/test_envpool.py
    (async_reset; recv -> send(action, env_id)).  Every recv returns exactly batch_size
    rows; each row is checked, by env id, against what the oracle says that env's next
    output must be."""
    from oracle.oracle_lib import OraclePool

    N, B = 24, 8
    env = ep.make_gym("CartPole-v1", num_envs=N, batch_size=B, seed=4)
    assert env.is_async
    orc = OraclePool("CartPole", N, seed=4, max_episode_steps=500)
    w = orc.reset()
    exp_obs, exp_step = w["obs"].copy(), w["elapsed_step"].copy()
    env.async_reset()
    rng = np.random.default_rng(6)
    for it in range(80):
        obs, rew, term, trunc, info = env.recv()
        ids = info["env_id"]
        assert obs.shape == (B, 4) and len(set(ids.tolist())) == B
        np.testing.assert_allclose(obs, exp_obs[ids], rtol=0, atol=1e-6)
        np.testing.assert_array_equal(info["elapsed_step"], exp_step[ids])
        if it % 7 == 3:
            # two sends of B/2 rows: the next recv of these envs straddles two submissions
            for part in (ids[: B // 2], ids[B // 2:]):
                a = rng.integers(0, 2, size=len(part)).astype(np.int32)
                w = orc.step(a, part)
                exp_obs[part], exp_step[part] = w["obs"], w["elapsed_step"]
                env.send(a, part)
        else:
            a = rng.integers(0, 2, size=B).astype(np.int32)
            w = orc.step(a, ids)
            exp_obs[ids], exp_step[ids] = w["obs"], w["elapsed_step"]
            env.send(a, ids)


def test_errors_and_engine_kwargs(ep):
    with pytest.raises(ValueError):
        ep.make_gym("CartPole-v1", num_envs=4, batch_size=2, gym_reset_return_info=False)
    with pytest.raises(AssertionError):
        ep.make_gym("NoSuchEnv-v0", num_envs=1)
    env = ep.make_gym("CartPole-v1", num_envs=8, seed=1, precision="f32", env_id_offset=100)
    _, info = env.reset()
    np.testing.assert_array_equal(info["env_id"], np.arange(100, 108))
    with pytest.raises(RuntimeError):
        env.step(np.zeros(8, dtype=np.float64).reshape(8, 1))   # wrong action shape
    with pytest.raises(RuntimeError):
        env.render()
