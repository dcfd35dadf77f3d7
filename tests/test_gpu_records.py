"""GPU: the reset-ahead record rings of classic_control (csrc/common.cuh StateView::rec).

An env's Reset() (cartpole.h:82-90, pendulum.h:77-84, acrobot.h:94-103,
mountain_car.h:76-82) is a pure function of its own mt19937 stream, so the engine draws each
env's NEXT initial states ahead of time into a ring (refill_kernel) and the step kernel's
auto-reset takes the next record with a load.  Records are produced and consumed in order,
hence every trajectory must stay the one the reference produces.  These tests stress what the
ring adds: the tightest legal reset spacing against small rings and long refill periods
(the bound rec_q >= refill_every + 2 of capi.cu run_chain), the CUDA-graph chain in which a
refill runs beside the following steps, repeated forced resets, partial-id steps, and the
fused rollout consuming and redrawing records."""
import numpy as np
import pytest

from helpers import REGISTERED, assert_batch_equal, random_actions

pytestmark = pytest.mark.gpu
CLASSIC = ["CartPole", "Pendulum", "Acrobot", "MountainCar", "MountainCarContinuous"]


def _outputs(pool):
    return {k: v.cpu().numpy() for k, v in pool.outputs_torch().items()}


@pytest.mark.parametrize("ring", [None, (4, 2), (8, 6), (16, 14)])
def test_ring_bound_under_the_tightest_reset_spacing(capi, monkeypatch, ring):
    """max_episode_steps = 1: every env resets every second step, the fastest a free-running
    env can consume records.  Default ring (16 records, refill every 8 steps), the smallest
    (4 / 2), 8 / 6 and the longest period a 16-ring allows (14): captured chains against the oracle."""
    import torch
    from oracle.oracle_lib import OraclePool

    if ring:
        monkeypatch.setenv("ENVPOOL_B200_REC_Q", str(ring[0]))
        monkeypatch.setenv("ENVPOOL_B200_REFILL_EVERY", str(ring[1]))
    N, T = 5000, 64
    rng = np.random.default_rng(12)
    pool = capi.CPool("CartPole", N, seed=2, max_episode_steps=1)
    orc = OraclePool("CartPole", N, seed=2, max_episode_steps=1)
    if ring:
        assert pool.state_layout()["rec_q"] == ring[0]
    pool.reset_device()
    orc.reset()
    acts = random_actions("CartPole", rng, (T, N))
    d_acts = torch.from_numpy(acts).cuda()
    t = 0
    for K in (64, 37, 64, 5, 64):
        pool.step_many_device(d_acts, t % T, K, use_graph=True)
        for k in range(K):
            want = orc.step(acts[(t + k) % T])
        t += K
        pool.sync()
        assert_batch_equal(_outputs(pool), want, "CartPole", 1e-6, f"ring={ring} after {t}")
    for k in range(9):            # direct launches: refill every refill_every-th one
        pool.step_device(d_acts[k])
        want = orc.step(acts[k])
    pool.sync()
    assert_batch_equal(_outputs(pool), want, "CartPole", 1e-6, f"ring={ring} direct")


@pytest.mark.parametrize("ms", [1, 2, 3, 17])
@pytest.mark.parametrize("task", CLASSIC)
def test_graph_chain_matches_oracle_with_dense_resets(capi, task, ms):
    """max_episode_steps = 1 / 2 / 3: every env resets every 2nd / 3rd / 4th step, the
    tightest legal spacing (a step that resets cannot be `done`).  The K-step chains are
    replayed from CUDA graphs (refill on a parallel branch) and compared with the oracle
    after each chain; K = 1, 2, 3, odd and even lengths, replayed twice."""
    import torch
    from oracle.oracle_lib import OraclePool

    _, iopt = REGISTERED[task]
    N = 2500
    rng = np.random.default_rng(3)
    pool = capi.CPool(task, N, seed=11, max_episode_steps=ms, iopt=iopt)
    orc = OraclePool(task, N, seed=11, max_episode_steps=ms, iopt=iopt)
    pool.reset_device()
    want = orc.reset()
    pool.sync()
    assert_batch_equal(_outputs(pool), want, task, 1e-6, "reset")
    T = 64
    acts = random_actions(task, rng, (T, N))
    d_acts = torch.from_numpy(acts).cuda()
    t = 0
    for K in (1, 2, 3, 7, 64, 7, 2, 64, 1):
        pool.step_many_device(d_acts, t % T, K, use_graph=True)
        for k in range(K):
            want = orc.step(acts[(t + k) % T])
        t += K
        pool.sync()
        assert_batch_equal(_outputs(pool), want, task, 1e-6, f"{task} ms={ms} after {t} steps")
    # the same steps as direct launches (refill on the step's own stream): bit-identical
    other = capi.CPool(task, N, seed=11, max_episode_steps=ms, iopt=iopt)
    other.reset_device()
    done_steps = 0
    while done_steps < t:
        k = min(T - done_steps % T, t - done_steps)
        other.step_many_device(d_acts, done_steps % T, k, use_graph=False)
        done_steps += k
    other.sync()
    assert_batch_equal(_outputs(other), _outputs(pool), task, 0.0, "graph vs direct launches")


@pytest.mark.parametrize("task", ["CartPole", "Pendulum"])
def test_repeated_forced_resets_and_partial_ids(capi, task):
    """Forced resets consume one record per call, whatever the env's state; partial-id steps
    and resets only touch their own rings."""
    from oracle.oracle_lib import OraclePool

    ms, iopt = REGISTERED[task]
    N = 700
    rng = np.random.default_rng(5)
    pool = capi.CPool(task, N, seed=3, max_episode_steps=4, iopt=iopt)
    orc = OraclePool(task, N, seed=3, max_episode_steps=4, iopt=iopt)
    for r in range(21):   # more forced resets in a row than a ring holds (16): refills in between
        assert_batch_equal(pool.reset(), orc.reset(), task, 1e-6, f"reset #{r}")
    for t in range(12):
        sub = np.sort(rng.choice(N, size=300, replace=False)).astype(np.int32)
        if t % 4 == 3:
            assert_batch_equal(pool.reset(sub), orc.reset(sub), task, 1e-6, f"partial reset {t}")
        a = random_actions(task, rng, (300,))
        assert_batch_equal(pool.step(a, sub), orc.step(a, sub), task, 1e-6, f"partial step {t}")
    a = random_actions(task, rng, (N,))
    assert_batch_equal(pool.step(a), orc.step(a), task, 1e-6, "full step")


@pytest.mark.parametrize("task", ["CartPole", "MountainCar"])
def test_rollout_then_steps_share_the_record_stream(capi, task):
    """A fused rollout takes each env's record at its first reset, draws later resets in
    place and redraws the record before it ends; single steps afterwards continue the same
    per-env draw sequence."""
    import torch
    from oracle.oracle_lib import OraclePool

    _, iopt = REGISTERED[task]
    N, T, ms = 1500, 24, 5
    rng = np.random.default_rng(8)
    pool = capi.CPool(task, N, seed=9, max_episode_steps=ms, iopt=iopt)
    orc = OraclePool(task, N, seed=9, max_episode_steps=ms, iopt=iopt)
    assert_batch_equal(pool.reset(), orc.reset(), task, 1e-6, "reset")
    acts = random_actions(task, rng, (T, N))
    tdt = {np.dtype(np.int32): torch.int32, np.dtype(np.float32): torch.float32,
           np.dtype(np.float64): torch.float64, np.dtype(np.bool_): torch.bool}
    cols = [torch.empty((T, N) + k.shape, dtype=tdt[k.dtype], device="cuda") for k in pool.keys]
    for rep in range(2):
        pool.rollout_device(torch.from_numpy(acts).cuda(), T, cols)
        pool.sync()
        for t in range(T):
            got = {k.name: c[t].cpu().numpy() for k, c in zip(pool.keys, cols)}
            assert_batch_equal(got, orc.step(acts[t]), task, 1e-6, f"rollout {rep} t={t}")
        for t in range(7):
            a = random_actions(task, rng, (N,))
            assert_batch_equal(pool.step(a), orc.step(a), task, 1e-6, f"step after rollout {t}")


def test_large_batch_prefix(capi):
    """A 300000-env pool (state + rings beyond L2 residency): a prefix against the oracle."""
    import torch
    from oracle.oracle_lib import OraclePool

    N, P = 300000, 4096
    pool = capi.CPool("CartPole", N, seed=1, max_episode_steps=6)
    orc = OraclePool("CartPole", P, seed=1, max_episode_steps=6)
    rng = np.random.default_rng(2)
    pool.reset_device()
    want = orc.reset()
    for t in range(20):
        a = rng.integers(0, 2, size=N).astype(np.int32)
        pool.step_device(torch.from_numpy(a).cuda())
        want = orc.step(a[:P])
    pool.sync()
    got = {k: v[:P] for k, v in _outputs(pool).items()}
    assert_batch_equal(got, want, "CartPole", 1e-6, "prefix of a 300000-env pool")
