"""CPU: a timing model of the captured exchange chain (csrc/capi.cu run_chain + the device-side
waits of csrc/exchange.cuh), in the spirit of tests/test_mt_chunked_model.py: the dependency
rules of the engine restated in plain Python and played with random kernel durations and link
latencies, for the world sizes the hardware tests of a round cannot always reach (4 and 8 GPUs).

Per rank and exchanged step k (ring depth D, P push branches):
    step(k)   after step(k-1), wait(k-D+1) (run-ahead bound) and push(k-D) (slot k % D sent)
    push(k)   after step(k) and push(k-P) (same branch); on the device it first needs the CREDIT
              ack[g] >= k-D+1 from every rank g (published by g's wait(k-D+1) at its START), then
              stores into slot k % D of every peer and, at its end, raises flag[k % D] = k+1 there
    wait(k)   after push(k) and wait(k-1); publishes ack = k at its start, then needs
              flag[k % D][g] >= k+1 from every peer g, then re-expands (consumes) slot k % D
What must hold whatever the timings:
  * progress: every kernel of every rank gets a start time (no deadlock);
  * no overwrite before consumption: the stores of step k+D arrive on a receiver only after its
    wait(k) has finished with slot k % D;
  * no overwrite before sending: step(k+D) starts on a rank only after its push(k) has read slot
    k % D;
  * data before flag: a receiver starts consuming step k only after every peer's stores of step k
    have arrived (by construction of the flag time; asserted anyway)."""
import itertools

import numpy as np
import pytest


def simulate(W, D, P, K, rng):
    lat = rng.uniform(0.5, 3.0, size=(W, W))          # one-way latency g -> r
    np.fill_diagonal(lat, 0.0)
    dur = {"step": (1.0, 4.0), "push": (2.0, 12.0), "wait": (1.0, 6.0)}
    d = {(op, r, k): rng.uniform(*dur[op]) for op in dur for r in range(W) for k in range(K)}
    start, end, copy_start = {}, {}, {}
    pending = set(d)

    def known(keys):
        return all(key in end for key in keys)

    progress = True
    while pending and progress:
        progress = False
        for key in sorted(pending, key=lambda x: (x[2], x[1], x[0])):
            op, r, k = key
            if op == "step":
                deps = [("step", r, k - 1)] if k >= 1 else []
                if k >= D - 1:
                    deps.append(("wait", r, k - D + 1))
                if k >= D:
                    deps.append(("push", r, k - D))
                if not known(deps):
                    continue
                start[key] = max([end[x] for x in deps], default=0.0)
                end[key] = start[key] + d[key]
            elif op == "push":
                deps = [("step", r, k)] + ([("push", r, k - P)] if k >= P else [])
                if not known(deps):
                    continue
                t = max(end[x] for x in deps)
                if k >= D:   # credit: every rank (own included) has started wait(k-D+1)
                    need = [("wait", g, k - D + 1) for g in range(W)]
                    if not all(x in start for x in need):
                        continue
                    t = max(t, max(start[("wait", g, k - D + 1)] + lat[g, r] for g in range(W)))
                start[key] = max(end[x] for x in deps)
                copy_start[key] = t
                end[key] = t + d[key]
            else:
                deps = [("push", r, k)] + ([("wait", r, k - 1)] if k >= 1 else [])
                if not known(deps):
                    continue
                s = max(end[x] for x in deps)
                peers = [("push", g, k) for g in range(W) if g != r]
                if not known(peers):
                    # the kernel has started (its ack is out) even though it still spins
                    if key not in start:
                        start[key] = s
                        progress = True
                    continue
                start[key] = s
                flag = max(end[("push", g, k)] + lat[g, r] for g in range(W) if g != r)
                end[key] = max(s, flag) + d[key]
            pending.discard(key)
            progress = True
    return start, end, copy_start, lat, pending


@pytest.mark.parametrize("W,D,P", [(2, 4, 3), (4, 4, 3), (8, 4, 3), (8, 2, 3), (8, 3, 1),
                                   (8, 8, 3), (3, 4, 1), (16, 4, 3)])
def test_exchange_chain_protocol_model(W, D, P):
    rng = np.random.default_rng(100 * W + 10 * D + P)
    for trial in range(6):
        K = int(rng.integers(3 * D, 6 * D + 5))
        start, end, copy_start, lat, pending = simulate(W, D, P, K, rng)
        assert not pending, (W, D, P, K, sorted(pending)[:4])          # no deadlock
        for r, g, k in itertools.product(range(W), range(W), range(K)):
            if g == r:
                continue
            # stores of step k from g reach r between g's copy start and g's flag arrival
            first_arrival = copy_start[("push", g, k)] + lat[g, r]
            flag_arrival = end[("push", g, k)] + lat[g, r]
            consume_start = end[("wait", r, k)] - 0.0
            assert flag_arrival <= end[("wait", r, k)] + 1e-12
            if k + D < K:   # the next user of the slot must not arrive before r is done with it
                nxt = copy_start[("push", g, k + D)] + lat[g, r]
                assert nxt >= end[("wait", r, k)] - 1e-12, (W, D, P, r, g, k, nxt,
                                                             end[("wait", r, k)])
            assert first_arrival <= consume_start
        for r, k in itertools.product(range(W), range(K - D)):
            assert start[("step", r, k + D)] >= end[("push", r, k)] - 1e-12
            assert start[("step", r, k + D)] >= end[("wait", r, k + 1)] - 1e-12


def test_run_ahead_is_bounded_by_the_ring():
    """With instant pushes and waits but one slow rank, the fast ranks' step chains get at most
    D - 1 steps ahead of the slow rank's waits (the credit and the run-ahead edge together)."""
    W, D, P, K = 4, 4, 3, 40
    rng = np.random.default_rng(5)
    start, end, copy_start, lat, pending = simulate(W, D, P, K, rng)
    assert not pending
    for r in range(W):
        for k in range(D - 1, K):
            assert start[("step", r, k)] >= end[("wait", r, k - D + 1)] - 1e-12
