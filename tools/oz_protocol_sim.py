"""Discrete-event model of the mbarrier / tcgen05.commit protocols of the INT8 substitution kernels (csrc/ozaki.cuh:
update_kernel, update128_kernel, and update_stack_kernel<S, 2> -- clusters of two CTAs that share every V digit stage
through TMA multicast).  It replays the producer, MMA-issuer and drain roles with the kernels' own loop structure,
slot / phase arithmetic and barrier counts under random latencies and checks, for every schedule, that

  * nothing deadlocks (every role finishes),
  * every MMA reads a stage that holds the k-step it expects, in every CTA whose shared memory it reads,
  * no stage is overwritten by a bulk copy while MMAs that have to read its previous content are still outstanding,
  * accumulators are never written while a drain of the previous round is reading them, and never drained before the
    round's last MMA has completed.

mbarrier semantics modelled: `count` pending arrivals per phase, phase parity; try_wait.parity(P) succeeds when the phase of
parity P has completed; a TMA stage is one arrival (arrive.expect_tx) plus its bytes (second pending unit); tcgen05.commit
arrives when every MMA issued before it (in order) has completed; the multicast form arrives in both CTAs.

The hardware is the authority (tools/oz_test: all three kernels are bit-identical to the host digit emulation on a B200);
this model exists to catch slot / phase / count mistakes before GPU time is spent on them, and it documents WHY the
cluster protocol is safe: a ring slot is refilled -- by either CTA's multicast copy -- only after BOTH CTAs' MMAs have
read it (`empty` counts two commits, each multicast to both CTAs).

    python tools/oz_protocol_sim.py            # a few hundred random schedules of each kernel
"""
import heapq
import random


class Barrier:
    def __init__(self, count, name):
        self.count, self.pending, self.phase, self.name = count, count, 0, name

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "too many arrivals on " + self.name
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def done(self, parity):
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, rng):
        self.t, self.q, self.n, self.rng = 0.0, [], 0, rng
        self.live = 0

    def at(self, dt, fn):
        self.n += 1
        heapq.heappush(self.q, (self.t + dt, self.n, fn))

    def spawn(self, gen, name):
        self.live += 1

        def step():
            try:
                req = next(gen)
            except StopIteration:
                self.live -= 1
                return
            kind = req[0]
            if kind == "delay":
                self.at(req[1], step)
            else:  # ("wait", barrier, parity): poll with the try_wait latency
                _, bar, par = req

                def poll():
                    if bar.done(par):
                        self.at(self.rng.uniform(0.05, 0.2), step)
                    else:
                        self.at(self.rng.uniform(0.2, 1.0), poll)
                poll()
        self.at(0.0, step)

    def run(self, limit=150_000):
        steps = 0
        while self.q and steps < limit:
            self.t, _, fn = heapq.heappop(self.q)
            fn()
            steps += 1
        return self.live == 0 and not self.q


class TensorPipe:
    """In-order asynchronous queue of one issuing thread: MMAs (with a check callback run at execution) and commits."""

    def __init__(self, sim):
        self.sim, self.queue, self.busy = sim, [], False

    def push(self, dur, fn):
        self.queue.append((dur, fn))
        if not self.busy:
            self._next()

    def _next(self):
        if not self.queue:
            self.busy = False
            return
        self.busy = True
        dur, fn = self.queue.pop(0)

        def fin():
            fn()
            self._next()
        self.sim.at(dur, fin)

    def outstanding(self):
        return len(self.queue) + (1 if self.busy else 0)


def simulate(kind, KT, KT_CHUNK, stages0, stages1, seed):
    """kind: 'narrow' (one pass per chunk: update_kernel, update_stack_kernel<S, 1>), 'wide' (two passes: update128_kernel).
    ('pair' models a cta_group::2 design with a relay warp that was drafted in round 1 and never shipped.)"""
    rng = random.Random(seed)
    sim = Sim(rng)
    ncta = 2 if kind == "pair" else 1
    passes = 1 if kind == "narrow" else 2
    nchunks = (KT + KT_CHUNK - 1) // KT_CHUNK
    rounds = passes * nchunks
    nst = {0: stages0, 1: stages1} if passes == 2 else {0: stages0}
    B = lambda c, n: Barrier(c, n)
    full = [[B(2, "full%d.%d" % (c, i)) for i in range(16)] for c in range(ncta)]      # arrive.expect_tx + bytes
    empty = [[B(1, "empty%d.%d" % (c, i)) for i in range(16)] for c in range(ncta)]
    peer_ready = [B(1, "peer_ready%d" % i) for i in range(16)]                           # leader only
    tmem_full = [B(1, "tmem_full%d" % c) for c in range(ncta)]
    tmem_empty = B(4 * ncta, "tmem_empty")                                              # leader only
    pass_done = [B(1, "pass_done%d" % c) for c in range(ncta)]
    stage_content = [dict() for _ in range(ncta)]     # (pass, slot) -> (round, kt) landed
    stage_readers = [dict() for _ in range(ncta)]     # (pass, slot) -> outstanding MMA reads of the current content
    ring_owner = [None] * ncta                        # which pass's partition the ring currently holds
    acc_state = {"writing_round": -1, "draining": 0, "complete_round": -1}
    pipe = TensorPipe(sim)
    errors = []

    def round_info(r):
        p = r % passes if passes == 2 else 0
        c = r // passes
        kt0 = c * KT_CHUNK
        return p, kt0, min(KT, kt0 + KT_CHUNK)

    def producer(c):
        cnt = {0: 0, 1: 0}
        for r in range(rounds):
            p, kt0, kt1 = round_info(r)
            if passes == 2 and r > 0:
                yield ("wait", pass_done[c], (r - 1) & 1)
            for kt in range(kt0, kt1):
                n = cnt[p]
                cnt[p] += 1
                s = n % nst[p]
                need_wait = (kt - kt0 >= nst[p]) if passes == 2 else (n >= nst[p])
                if need_wait:
                    yield ("wait", empty[c][p * 8 + s], ((n // nst[p]) - 1) & 1)
                yield ("delay", rng.uniform(0.05, 0.3))
                key = (p, s)
                full[c][p * 8 + s].arrive()                      # arrive.expect_tx

                def land(c=c, key=key, r=r, kt=kt, p=p, s=s):
                    if stage_readers[c].get(key, 0) != 0:
                        errors.append("CTA %d: stage %s overwritten with %d reads outstanding" % (c, key, stage_readers[c][key]))
                    if passes == 2 and ring_owner[c] is not None and ring_owner[c] != p:
                        # re-partitioning: no reads of the other pass's stages may be outstanding
                        for k2, v in stage_readers[c].items():
                            if k2[0] != p and v != 0:
                                errors.append("CTA %d: ring re-partitioned with reads of pass %d outstanding" % (c, k2[0]))
                    ring_owner[c] = p
                    stage_content[c][key] = (r, kt)
                    full[c][p * 8 + s].arrive()                  # complete_tx
                sim.at(rng.uniform(1.0, 6.0), land)

    def relay():  # peer CTA 1 -> leader
        cnt = {0: 0, 1: 0}
        for r in range(rounds):
            p, kt0, kt1 = round_info(r)
            for kt in range(kt0, kt1):
                n = cnt[p]
                cnt[p] += 1
                s = n % nst[p]
                yield ("wait", full[1][p * 8 + s], (n // nst[p]) & 1)
                yield ("delay", rng.uniform(0.2, 0.6))           # remote arrive latency
                peer_ready[p * 8 + s].arrive()

    def mma():
        cnt = {0: 0, 1: 0}
        for r in range(rounds):
            p, kt0, kt1 = round_info(r)
            if r > 0:
                yield ("wait", tmem_empty, (r - 1) & 1)
            for kt in range(kt0, kt1):
                n = cnt[p]
                cnt[p] += 1
                s = n % nst[p]
                yield ("wait", full[0][p * 8 + s], (n // nst[p]) & 1)
                if ncta == 2:
                    yield ("wait", peer_ready[p * 8 + s], (n // nst[p]) & 1)
                key = (p, s)
                for c in range(ncta):
                    stage_readers[c][key] = stage_readers[c].get(key, 0) + 1

                def execute(key=key, r=r, kt=kt):
                    if acc_state["draining"]:
                        errors.append("MMA of round %d wrote the accumulators during a drain" % r)
                    acc_state["writing_round"] = r
                    for c in range(ncta):
                        if stage_content[c].get(key) != (r, kt):
                            errors.append("CTA %d: MMA (round %d, kt %d) read stage %s holding %s" %
                                          (c, r, kt, key, stage_content[c].get(key)))
                        stage_readers[c][key] -= 1
                pipe.push(rng.uniform(0.3, 1.5), execute)

                def commit_empty(p=p, s=s):
                    for c in range(ncta):
                        empty[c][p * 8 + s].arrive()
                pipe.push(0.01, commit_empty)
                yield ("delay", rng.uniform(0.02, 0.3))

            def commit_round(r=r):
                acc_state["complete_round"] = r
                for c in range(ncta):
                    tmem_full[c].arrive()
                    pass_done[c].arrive()
            pipe.push(0.01, commit_round)

    def drain(c, w):
        for r in range(rounds):
            yield ("wait", tmem_full[c], r & 1)
            if acc_state["complete_round"] < r:
                errors.append("drain of round %d started before its MMAs completed" % r)
            acc_state["draining"] += 1
            yield ("delay", rng.uniform(0.5, 3.0))
            acc_state["draining"] -= 1
            if c == 0:
                tmem_empty.arrive()
            else:
                yield ("delay", rng.uniform(0.2, 0.6))
                tmem_empty.arrive()

    for c in range(ncta):
        sim.spawn(producer(c), "producer%d" % c)
        for w in range(4):
            sim.spawn(drain(c, w), "drain%d.%d" % (c, w))
    if ncta == 2:
        sim.spawn(relay(), "relay")
    sim.spawn(mma(), "mma")
    finished = sim.run()
    if not finished:
        errors.append("deadlock: %d roles still waiting at t = %.1f" % (sim.live, sim.t))
    return errors


def simulate_stackpair(KT, KT_CHUNK, stages, seed, empty_count=2):
    """update_stack_kernel<S, 2>: two CTAs on the same 64-point V tile, block rows 2j and 2j + 1.
    Per CTA: producer (own L digits + HALF of the V stage, multicast into both CTAs' shared memory at the same offset,
    completing on both CTAs' full[s]), MMA issuer (own tensor pipe, cta_group::1; tcgen05.commit multicast onto both CTAs'
    empty[s]), four drain warps (own tensor memory).  full[s] = 1 arrival (arrive.expect_tx) + three byte deliveries
    (own L, V half of CTA 0, V half of CTA 1); empty[s] counts `empty_count` commits (2 in the kernel)."""
    rng = random.Random(seed)
    sim = Sim(rng)
    nchunks = (KT + KT_CHUNK - 1) // KT_CHUNK
    full = [[Barrier(4, "full%d.%d" % (c, i)) for i in range(stages)] for c in range(2)]
    empty = [[Barrier(empty_count, "empty%d.%d" % (c, i)) for i in range(stages)] for c in range(2)]
    tmem_full = [Barrier(1, "tmem_full%d" % c) for c in range(2)]
    tmem_empty = [Barrier(4, "tmem_empty%d" % c) for c in range(2)]
    content = [[dict(A=None, h0=None, h1=None) for _ in range(stages)] for _ in range(2)]
    readers = [[0] * stages for _ in range(2)]
    acc = [dict(draining=0, complete=-1) for _ in range(2)]
    pipes = [TensorPipe(sim), TensorPipe(sim)]
    errors = []

    def land(c, s, part, kt):
        def fn():
            if readers[c][s] != 0:
                errors.append("CTA %d: stage %d part %s overwritten with %d reads outstanding" % (c, s, part, readers[c][s]))
            content[c][s][part] = kt
            try:
                full[c][s].arrive()
            except AssertionError as e:
                errors.append(str(e))
        return fn

    def producer(c):
        for kt in range(KT):
            s = kt % stages
            if kt >= stages:
                yield ("wait", empty[c][s], ((kt // stages) - 1) & 1)
            yield ("delay", rng.uniform(0.05, 0.3))
            full[c][s].arrive()                                            # arrive.expect_tx(whole stage)
            sim.at(rng.uniform(1.0, 6.0), land(c, s, "A", kt))               # own L digits
            for dst in range(2):                                           # this CTA's half of V, multicast to both
                sim.at(rng.uniform(1.0, 6.0), land(dst, s, "h%d" % c, kt))

    def mma(c):
        kt = 0
        for ch in range(nchunks):
            if ch > 0:
                yield ("wait", tmem_empty[c], (ch - 1) & 1)
            kt_end = min(KT, (ch + 1) * KT_CHUNK)
            while kt < kt_end:
                s = kt % stages
                yield ("wait", full[c][s], (kt // stages) & 1)
                readers[c][s] += 1

                def execute(c=c, s=s, kt=kt, ch=ch):
                    if acc[c]["draining"]:
                        errors.append("CTA %d: MMA of chunk %d wrote the accumulators during a drain" % (c, ch))
                    got = content[c][s]
                    if not (got["A"] == kt and got["h0"] == kt and got["h1"] == kt):
                        errors.append("CTA %d: MMA kt %d read stage %d holding %s" % (c, kt, s, got))
                    readers[c][s] -= 1
                pipes[c].push(rng.uniform(0.3, 1.5), execute)

                def commit_empty(s=s):                                     # tcgen05.commit ... multicast::cluster, mask 0b11
                    for dst in range(2):
                        try:
                            empty[dst][s].arrive()
                        except AssertionError as e:
                            errors.append(str(e))
                pipes[c].push(0.01, commit_empty)
                yield ("delay", rng.uniform(0.02, 0.3))
                kt += 1

            def commit_chunk(c=c, ch=ch):
                acc[c]["complete"] = ch
                tmem_full[c].arrive()
            pipes[c].push(0.01, commit_chunk)

    def drain(c, w):
        for ch in range(nchunks):
            yield ("wait", tmem_full[c], ch & 1)
            if acc[c]["complete"] < ch:
                errors.append("CTA %d: drain of chunk %d started before its MMAs completed" % (c, ch))
            acc[c]["draining"] += 1
            yield ("delay", rng.uniform(0.5, 3.0))
            acc[c]["draining"] -= 1
            tmem_empty[c].arrive()

    for c in range(2):
        sim.spawn(producer(c), "producer%d" % c)
        sim.spawn(mma(c), "mma%d" % c)
        for w in range(4):
            sim.spawn(drain(c, w), "drain%d.%d" % (c, w))
    if not sim.run():
        errors.append("deadlock: %d roles still waiting at t = %.1f" % (sim.live, sim.t))
    return errors


EMPTY_COUNT = 2   # commits an `empty` barrier of update_stack_kernel<S, 2> waits for (tests mutate this to 1)


def campaign(n_seeds=40):
    """Random schedules over the shapes the kernels meet: K below / at / above the ring depth, 1 to 3 drain intervals."""
    failures = []
    shapes = [(4, 512), (5, 512), (12, 512), (40, 16), (37, 16), (96, 32), (64, 64)]   # (k-steps, k-steps per chunk)
    for kind, st0, st1 in (("narrow", 5, 5), ("wide", 7, 3)):
        for KT, chunk in shapes:
            for seed in range(n_seeds):
                errs = simulate(kind, KT, chunk, st0, st1, seed * 7919 + KT)
                if errs:
                    failures.append((kind, KT, chunk, seed, errs[:3]))
    for KT, chunk in shapes:
        for seed in range(n_seeds):
            errs = simulate_stackpair(KT, chunk, 5, seed * 7919 + KT, EMPTY_COUNT)
            if errs:
                failures.append(("stackpair", KT, chunk, seed, errs[:3]))
    return failures


if __name__ == "__main__":
    bad = campaign()
    for f in bad[:10]:
        print("FAIL", f)
    print("%d failing schedules" % len(bad))
