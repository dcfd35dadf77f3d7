"""GPU: the reference's OWN Python test cases for this path, re-stated against
`envpool_b200.make` -- the ones that do not need gymnasium as a comparator (it is not
installed here; where the reference compares with gymnasium's env the oracle, pinned to the
reference's C++, takes that role in tests/test_gpu_parity.py).

  classic_control_test.py:34-57   run_deterministic_check (same seed -> same obs, other seed ->
                                  different obs, obs inside the observation space)
  classic_control_test.py:25-32   run_space_check (gym and dm specs agree on the bounds)
  toy_text_test.py:30-92          test_catch (a winning and a losing trajectory, gym + dm)
  toy_text_test.py:224-243        test_nchain (return statistics of a random agent)
  toy_text_test.py:258-284        test_cliffwalking (scripted paths along and into the cliff)
  toy_text_test.py:313-333        test_blackjack (return statistics of a random agent)
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CLASSIC = ["CartPole-v0", "CartPole-v1", "Pendulum-v0", "Pendulum-v1", "MountainCar-v0",
           "MountainCarContinuous-v0", "Acrobot-v1"]


@pytest.fixture(scope="module")
def ep(engine_built):
    import torch

    assert torch.cuda.is_available()
    import envpool_b200

    return envpool_b200


def _sample(space, rng, n):
    if hasattr(space, "n"):
        return rng.integers(0, space.n, size=n).astype(np.int32)
    return rng.uniform(space.low, space.high, size=(n,) + tuple(space.shape)).astype(np.float32)


@pytest.mark.parametrize("task", CLASSIC)
def test_classic_control_deterministic_and_bounded(ep, task):
    n = 4
    env0 = ep.make_gym(task, num_envs=n, seed=0)
    env1 = ep.make_gym(task, num_envs=n, seed=0)
    env2 = ep.make_gym(task, num_envs=n, seed=1)
    for e in (env0, env1, env2):
        e.reset()
    eps = np.finfo(np.float32).eps
    space = env0.observation_space
    lo, hi = space.low - eps, space.high + eps
    rng = np.random.default_rng(0)
    for _ in range(400):
        a = _sample(env0.action_space, rng, n)
        obs0, obs1, obs2 = env0.step(a)[0], env1.step(a)[0], env2.step(a)[0]
        np.testing.assert_allclose(obs0, obs1)
        assert not np.allclose(obs0, obs2)
        for o in (obs0, obs2):
            assert np.all(lo <= o) and np.all(o <= hi), o


@pytest.mark.parametrize("task", CLASSIC)
def test_classic_control_spaces_agree(ep, task):
    spec = ep.make_spec(task)
    gym_space = spec.observation_space
    dm_spec = spec.observation_spec()
    dm_obs = dm_spec.obs if hasattr(dm_spec, "obs") else dm_spec
    np.testing.assert_allclose(gym_space.low, np.broadcast_to(dm_obs.minimum, gym_space.shape))
    np.testing.assert_allclose(gym_space.high, np.broadcast_to(dm_obs.maximum, gym_space.shape))


@pytest.mark.parametrize("env_type", ["dm", "gymnasium"])
def test_catch_win_and_lose(ep, env_type):
    num_envs, row, col = 3, 10, 5
    e = ep.make("Catch-v0", env_type=env_type, num_envs=num_envs)

    def reset():
        return e.reset().observation.obs if env_type == "dm" else e.reset()[0]

    def step(action):
        if env_type == "dm":
            ts = e.step(action, np.arange(num_envs))
            return ts.observation.obs, ts.reward, ts.last()
        obs, rew, term, trunc, _ = e.step(action, np.arange(num_envs))
        return obs, rew, np.logical_or(term, trunc)

    for chase, final_reward in ((True, 1), (False, -1)):
        obs = reset()
        assert obs.shape == (num_envs, row, col)
        ball = np.where(obs[:, 0] == 1)[1]
        paddle = np.where(obs[:, -1] == 1)[1]
        for t in range(row - 1):
            if chase:
                action = np.sign(ball - paddle) + 1
            else:
                action = np.sign(paddle - ball) + 1
                action[action == 1] = 0
            obs, rew, done = step(action.astype(np.int32))
            assert obs.shape == (num_envs, row, col)
            paddle = np.where(obs[:, -1] == 1)[1]
            if t != row - 2:
                assert np.all(rew == 0) and np.all(~done)
            else:
                assert np.all(rew == final_reward) and np.all(done)


def test_nchain_random_agent_statistics(ep):
    num_envs = 100
    env = ep.make_gymnasium("NChain-v0", num_envs=num_envs)
    assert env.observation_space.n == 5 and env.action_space.n == 2
    env.reset()
    rng = np.random.default_rng(0)
    reward, done = 0, [False]
    while not done[0]:
        obs, rew, term, trunc, _ = env.step(rng.integers(0, 2, size=num_envs).astype(np.int32))
        done = np.logical_or(term, trunc)
        reward = reward + rew
    assert abs(np.mean(reward) - 1310) < 30 and abs(np.std(reward) - 78) < 15


def test_cliffwalking_scripted_paths(ep):
    """Up 4 (3 effective: the grid has 4 rows), right i, down 4: lands on the cliff for
    1 <= i <= 10 (reward -100, back to the start, no termination), on the start row for i = 0
    and on the goal for i = 11 (terminated)."""
    env = ep.make_gymnasium("CliffWalking-v1")
    assert env.observation_space.n == 48 and env.action_space.n == 4
    for i in range(12):
        obs, info = env.reset()
        assert obs[0] == 36
        np.testing.assert_allclose(info["prob"], 1.0)
        x, y = 3, 0
        for a in [0] * 4 + [1] * i + [2] * 4:
            obs, rew, term, trunc, info = env.step(np.array([a], np.int32))
            if a == 0:
                x = max(x - 1, 0)
            elif a == 1:
                y = min(y + 1, 11)
            else:
                x = min(x + 1, 3)
            want_rew, want_term = -1.0, False
            if x == 3 and 1 <= y <= 10:
                x, y, want_rew = 3, 0, -100.0
            elif x == 3 and y == 11:
                want_term = True
            assert obs[0] == x * 12 + y and rew[0] == want_rew
            assert bool(term[0]) == want_term and not trunc[0]
            np.testing.assert_allclose(info["prob"], 1.0)
            if want_term:
                break


def test_blackjack_random_agent_statistics(ep):
    np.random.seed(0)
    num_envs = 100
    env = ep.make_gymnasium("Blackjack-v1", num_envs=num_envs)
    assert env.observation_space.shape == (3,) and env.action_space.n == 2
    reward, rewards = np.zeros(num_envs), []
    for _ in range(1000):
        obs, rew, term, trunc, _ = env.step(np.random.randint(2, size=(num_envs,)))
        done = np.logical_or(term, trunc)
        reward += rew
        if np.any(done):
            rewards += reward[done].tolist()
            reward[done] = 0
    assert abs(np.mean(rewards) + 0.395) < 0.05
    assert abs(np.std(rewards) - 0.89) < 0.05
