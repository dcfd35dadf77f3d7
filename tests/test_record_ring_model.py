"""CPU: a model of the reset-ahead record rings (csrc/common.cuh env_step / refill_kernel, the
refill placement of csrc/capi.cu launch_batch and run_chain), in the spirit of
tests/test_mt_chunked_model.py: the rules restated in plain Python and played against adversarial
and random episode patterns and refill timings.

Per env: a ring of Q slots; rcons / rprod count consumed / produced records MOD 256 (uint8 on
the device); a reset takes record number rcons from slot rcons % Q; a refill reads rcons (any
time while it runs: a step kernel may be consuming beside it), produces Q - (rprod - rcons)
records in order into slots rprod % Q ... and bumps rprod.
Placement of the refills:
  * direct launches: on the same stream behind every R-th step launch (an env may consume one
    record per launch there: forced resets);
  * engine-captured chains: refill j is launched behind step k_j = (j+1) R - 1 on a parallel
    branch, and the first step after refill j+1 is launched waits for refill j -- so refill j
    executes anywhere between the end of step k_j and the start of step k_{j+1} + 1, beside up
    to R steps; in a chain a reset step is never `done`, so an env consumes at most one record
    every two steps.
What must hold for every (Q, R) the engine accepts (R <= Q - 2): a reset always finds a valid
record (the ring never runs dry), records are consumed in production order (so every env follows
its own std::mt19937 stream), a refill never overwrites a record that has not been consumed,
and the uint8 counters survive wrap-around."""
import numpy as np
import pytest

ENGINE_PAIRS = [(16, 8), (16, 14), (16, 1), (8, 4), (8, 6), (4, 2), (4, 1)]   # R <= Q - 2


class Ring:
    def __init__(self, q):
        self.q = q
        self.slots = [None] * q
        self.rcons = 0          # uint8 on the device
        self.rprod = 0
        self.next_id = 0        # ids of produced records: what the mt19937 stream would give
        self.want = 0           # id the next reset must see
        self.refill(self.rcons)  # epb_create / state import leave the rings full

    def valid(self):
        return (self.rprod - self.rcons) & 255

    def consume(self):
        assert self.valid() >= 1, "ring ran dry"
        rec = self.slots[self.rcons % self.q]
        assert rec == self.want, (rec, self.want)      # production order
        self.want += 1
        self.rcons = (self.rcons + 1) & 255

    def refill(self, rcons_seen):
        """refill_kernel with the rcons value it happened to read."""
        need = self.q - ((self.rprod - rcons_seen) & 255)
        for _ in range(max(need, 0)):
            slot = self.rprod % self.q
            # the slot must not hold a record that is still to be consumed
            assert self.slots[slot] is None or self.slots[slot] < self.want, "overwrote a record"
            self.slots[slot] = self.next_id
            self.next_id += 1
            self.rprod = (self.rprod + 1) & 255
        assert self.valid() <= self.q


@pytest.mark.parametrize("q,r", ENGINE_PAIRS)
def test_direct_launches_with_forced_resets(q, r):
    """One record per launch at worst (every launch a forced reset), refill inline behind every
    R-th launch: never dry, in order, through several wrap-arounds of the uint8 counters."""
    rng = np.random.default_rng(q * 100 + r)
    for pattern in ("always", "random"):
        ring, since = Ring(q), 0
        for launch in range(1500):
            if pattern == "always" or rng.random() < 0.6:
                ring.consume()
            since += 1
            if since >= r:
                ring.refill(ring.rcons)
                since = 0
        assert ring.want > 256 or pattern == "random"


@pytest.mark.parametrize("q,r", ENGINE_PAIRS)
def test_captured_chain_with_refills_on_a_parallel_branch(q, r):
    """Chains: refill j executes at a random point of its window and reads rcons at a random
    earlier point of its own execution; worst-case episodes (done on every non-reset step) and
    random ones."""
    rng = np.random.default_rng(q * 1000 + r)
    for pattern in ("worst", "random", "worst"):
        ring = Ring(q)
        done = True                       # all envs start done (cartpole.h:67)
        K = 600 + int(rng.integers(0, 50))
        launch_after = [k for k in range(K) if k % r == r - 1 or k == K - 1]
        # refill j may execute between "after step launch_after[j]" and "before step
        # launch_after[j+1] + 1"; pick, per refill, the step it completes before and the
        # step after which it reads rcons
        pending = []                      # (read_after_step, complete_before_step)
        for j, k in enumerate(launch_after):
            last = launch_after[j + 1] + 1 if j + 1 < len(launch_after) else K
            complete_before = int(rng.integers(k + 1, last + 1))
            read_after = int(rng.integers(k, complete_before))
            pending.append((read_after, complete_before))
        reads = {}
        for k in range(K):
            for j, (read_after, complete_before) in enumerate(pending):
                if complete_before == k:              # refill j finishes before step k starts
                    ring.refill(reads.pop(j))
            # step k
            if done:
                ring.consume()
                done = False                          # a step that resets is never done
            else:
                done = True if pattern == "worst" else bool(rng.random() < 0.3)
            for j, (read_after, complete_before) in enumerate(pending):
                if read_after == k:                   # refill j samples rcons after step k
                    reads[j] = ring.rcons
        for j, (read_after, complete_before) in enumerate(pending):
            if complete_before == K:                  # joined at the end of the chain
                ring.refill(reads.pop(j))
        assert not reads
        assert ring.valid() == q                      # chains end with full rings
        assert ring.want >= K // 2 - 1 or pattern == "random"


def test_ring_too_small_for_its_refill_period_is_caught():
    """Negative control: R = Q (outside what the engine accepts) runs dry in the model."""
    ring, since = Ring(4), 0
    with pytest.raises(AssertionError):
        for launch in range(100):
            ring.consume()
            since += 1
            if since >= 5:
                ring.refill(ring.rcons)
                since = 0
