"""Model check (CPU, pure Python) of the peer-memory statistics exchange protocol of
csrc/peer_exchange.cu: every fp64 value travels as 8-byte words {data | sequence number}; buffers
have two parity halves that consecutive exchanges alternate; there is no barrier, no flag, no
fence.  The simulation interleaves the ranks' word stores and polling loads in random orders
(stores of one rank may also land out of program order, as NVLink does not order them) and checks
that every rank always collects exactly the vectors of the current exchange -- in particular that
reusing a parity half two exchanges later can never expose stale or too-new words."""
import random

import pytest


class Rank:
    """One rank executing a sequence of exchanges as a generator of atomic word operations."""

    def __init__(self, rank, world, bufs, vectors):
        self.rank, self.world, self.bufs, self.vectors = rank, world, bufs, vectors
        self.results = []
        self.pending = []          # stores issued but not yet delivered (arrive in any order)
        self.gen = self._run()

    def _run(self):
        seq = 0
        for vec in self.vectors:                       # vec = this rank's values for exchange `seq + 1`
            seq += 1
            par = seq & 1
            for p in range(self.world):                # 1. push: stores are only QUEUED here
                for i, v in enumerate(vec):
                    self.pending.append((p, par, i, (v, seq)))
            yield "pushed"
            acc = [0.0] * len(vec)                     # 2. collect in rank order, polling each word
            for i in range(len(vec)):
                for q in range(self.world):
                    while True:
                        word = self.bufs[self.rank][par][q].get(i)
                        if word is not None and word[1] == seq:
                            acc[i] += word[0]
                            break
                        yield "spin"
            self.results.append(acc)
            yield "done"

    def deliver_one(self, rng):
        """Deliver one queued store, in arbitrary order (the fabric does not order them)."""
        if not self.pending:
            return False
        p, par, i, word = self.pending.pop(rng.randrange(len(self.pending)))
        self.bufs[p][par][self.rank][i] = word
        return True


def simulate(world, n_exchanges, seed):
    rng = random.Random(seed)
    bufs = [[[dict() for _ in range(world)] for _ in range(2)] for _ in range(world)]   # [owner][par][src]{i: word}
    lens = [rng.choice([1, 2, 3]) for _ in range(n_exchanges)]
    vecs = [[[float(rng.randrange(1, 50) * (r + 1)) for _ in range(lens[s])] for s in range(n_exchanges)]
            for r in range(world)]
    ranks = [Rank(r, world, bufs, vecs[r]) for r in range(world)]
    alive = set(range(world))
    steps = 0
    while alive or any(rk.pending for rk in ranks):
        steps += 1
        assert steps < 200000, "protocol deadlocked in the model"
        r = rng.randrange(world)
        rk = ranks[r]
        # a queued store must be delivered before its rank can finish collecting, but it may be
        # delayed arbitrarily long relative to OTHER ranks' progress
        if rk.pending and (r not in alive or rng.random() < 0.5):
            rk.deliver_one(rng)
            continue
        if r in alive:
            try:
                next(rk.gen)
            except StopIteration:
                alive.discard(r)
    for s in range(n_exchanges):
        want = [sum(vecs[r][s][i] for r in range(world)) for i in range(lens[s])]
        for rk in ranks:
            assert rk.results[s] == want, (seed, s, rk.rank, rk.results[s], want)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_exchange_protocol_never_reads_stale_or_future_words(world):
    for seed in range(60 if world <= 4 else 15):
        simulate(world, n_exchanges=7, seed=1000 * world + seed)


def test_single_parity_half_would_be_unsafe():
    """Sanity check of the model itself: with ONE buffer half the same schedules do expose a
    too-new word (a fast rank's next exchange overwrites a word a slow rank has not read yet,
    and the slow rank then waits for a sequence number that is gone) -- which is why the kernel
    alternates two halves."""
    class OneHalf(Rank):
        def _run(self):
            seq = 0
            for vec in self.vectors:
                seq += 1
                for p in range(self.world):
                    for i, v in enumerate(vec):
                        self.pending.append((p, 0, i, (v, seq)))
                yield "pushed"
                acc = [0.0] * len(vec)
                for i in range(len(vec)):
                    for q in range(self.world):
                        for _ in range(2000):
                            word = self.bufs[self.rank][0][q].get(i)
                            if word is not None and word[1] == seq:
                                acc[i] += word[0]
                                break
                            yield "spin"
                        else:
                            raise RuntimeError("lost word")
                self.results.append(acc)
                yield "done"

    lost = 0
    for seed in range(40):
        rng = random.Random(seed)
        world = 3
        bufs = [[[dict() for _ in range(world)] for _ in range(2)] for _ in range(world)]
        vecs = [[[1.0] for _ in range(6)] for _ in range(world)]
        ranks = [OneHalf(r, world, bufs, vecs[r]) for r in range(world)]
        alive = set(range(world))
        try:
            for _ in range(100000):
                if not alive:
                    break
                r = rng.randrange(world)
                rk = ranks[r]
                if rk.pending and rng.random() < 0.5:
                    rk.deliver_one(rng)
                    continue
                if r in alive:
                    try:
                        next(rk.gen)
                    except StopIteration:
                        alive.discard(r)
        except RuntimeError:
            lost += 1
    assert lost > 0
