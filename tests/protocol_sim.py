"""Discrete-event model of the mbarrier protocol of the tcgen05 attention kernels (csrc/attention_tc.cuh).

Test infrastructure: it checks the *design* of the producer / MMA-issuer / softmax-warp hand-shake - every
parity wait must only pass once the phase the waiter MEANS has completed, no schedule may deadlock, and the data
hazards (S / P / O / stage reuse) must be ordered - under randomised thread timings. It models

  * mbarriers exactly as the hardware defines them: `try_wait.parity P` passes iff the barrier's current
    (incomplete) phase has the other parity, so a waiter that is two phases behind is released too early and
    one that is two phases ahead blocks forever;
  * the tensor pipe as an in-order queue: `tcgen05.commit` arrives on its mbarrier when every MMA issued before
    it by the same thread has completed.

It does not execute CUDA; it is how the single "O valid" barrier race of the first kernel version was
characterised (tests/test_attention_protocol_cpu.py keeps that version as a negative control).
"""
from __future__ import annotations

import random
from collections import deque


class Barrier:
    def __init__(self, name, count):
        self.name, self.count = name, count
        self.phase = 0            # number of completed phases == index of the current incomplete phase
        self.pending = count

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: more arrivals than the barrier expects"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def passes(self, parity):
        return (self.phase & 1) != parity


class Violation(Exception):
    pass


class Sim:
    """Threads are generators yielding ('wait', barrier, parity, intended_phase) | ('arrive', barrier) |
    ('delay', steps) | ('mma', name, duration) | ('commit', barrier) | ('check', bool, msg)."""

    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.threads = []          # [name, generator, state, payload]
        self.pipe = deque()        # in-order tensor pipe: ('mma', name, remaining) | ('commit', barrier)
        self.done_mma = set()
        self.violations = []

    def add(self, name, gen):
        self.threads.append([name, gen, "run", None])

    def _advance(self, th):
        name, gen, _, _ = th
        try:
            op = next(gen)
        except StopIteration:
            th[2] = "done"
            return
        kind = op[0]
        if kind == "wait":
            th[2], th[3] = "wait", op[1:]
            self._try_wait(th)
        elif kind == "arrive":
            op[1].arrive()
        elif kind == "delay":
            th[2], th[3] = "sleep", op[1]
        elif kind == "mma":
            self.pipe.append(["mma", op[1], op[2]])
        elif kind == "commit":
            self.pipe.append(["commit", op[1], 0])
        elif kind == "check":
            if not op[1]:
                self.violations.append(f"{name}: {op[2]}")
        else:
            raise ValueError(kind)

    def _try_wait(self, th):
        bar, parity, intended = th[3]
        if bar.passes(parity):
            if bar.phase <= intended:
                self.violations.append(
                    f"{th[0]}: wait on {bar.name} for phase {intended} released while only {bar.phase} phases completed")
            th[2], th[3] = "run", None

    def _pipe_tick(self):
        if not self.pipe:
            return
        op = self.pipe[0]
        if op[0] == "commit":
            op[1].arrive()
            self.pipe.popleft()
        else:
            op[2] -= 1
            if op[2] <= 0:
                self.done_mma.add(op[1])
                self.pipe.popleft()

    def run(self, max_ticks=2_000_000):
        for _ in range(max_ticks):
            if all(t[2] == "done" for t in self.threads) and not self.pipe:
                return
            self._pipe_tick()
            progressed = bool(self.pipe)
            for th in self.rng.sample(self.threads, len(self.threads)):
                if th[2] == "sleep":
                    th[3] -= 1
                    progressed = True
                    if th[3] <= 0:
                        th[2] = "run"
                elif th[2] == "wait":
                    self._try_wait(th)
                    progressed = progressed or th[2] == "run"
                elif th[2] == "run":
                    if self.rng.random() < 0.7:
                        self._advance(th)
                    progressed = True
            if not progressed:
                blocked = [(t[0], t[3][0].name, t[3][2]) for t in self.threads if t[2] == "wait"]
                raise Violation(f"deadlock: {blocked}")
        raise Violation("no termination")


def attention_cta(seed, nkt, n_items=1, single_odone=False, stages=3, nsw=8, slow_warp=None):
    """One CTA of attn_tc_kernel (n_items == 1) or attn_tcp_kernel (n_items > 1, running counters across items).
    `single_odone` reproduces the first kernel version (one O-valid barrier advancing once per key tile)."""
    sim = Sim(seed)
    rng = sim.rng
    B = lambda name, count: Barrier(name, count)
    sfull = [B("sfull0", 1), B("sfull1", 1)]
    pready = [B("pready0", nsw), B("pready1", nsw)]
    kvfull = [B(f"kvfull{i}", 1) for i in range(stages)]
    kvfree = [B(f"kvfree{i}", 1) for i in range(stages)]
    odone = [B("odone0", 1), B("odone1", 1)]
    ofree = B("ofree", nsw // 2 if nsw == 8 else 4)
    qfull = [B("qfull0", 1), B("qfull1", 1)]
    qfree = [B("qfree0", nsw + 1), B("qfree1", nsw + 1)]
    ttot = nkt * n_items
    n_half0 = ofree.count

    def od(c):          # barrier and phase index that mean "P.V of tile c retired"
        return (odone[0], c) if single_odone else (odone[c & 1], c >> 1)

    def stager():
        for i in range(n_items):
            qb, u = i & 1, i >> 1
            if u > 0:
                yield ("wait", qfree[qb], (u & 1) ^ 1, u - 1)
            yield ("delay", rng.randint(1, 40))
            yield ("arrive", qfull[qb])

    def producer():
        for c in range(ttot):
            st, use = c % stages, c // stages
            if use > 0:
                yield ("wait", kvfree[st], (use & 1) ^ 1, use - 1)
            yield ("delay", rng.randint(1, 30))          # TMA latency
            yield ("arrive", kvfull[st])

    def mma():
        qk = [0]

        def issue_qk():
            c = qk[0]
            i, g = divmod(c, nkt)
            if g == 0:
                yield ("wait", qfull[i & 1], (i >> 1) & 1, i >> 1)
            yield ("wait", kvfull[c % stages], (c // stages) & 1, c // stages)
            # hazards of overwriting S[c & 1]: P.V(c-2) read P from it (in order before us), every softmax warp
            # finished tile c-2 (it arrived on pready before P.V(c-2) was issued)
            yield ("mma", f"qk{c}", rng.randint(3, 12))
            yield ("commit", sfull[c & 1])
            if g == nkt - 1:
                yield ("commit", qfree[i & 1])
            qk[0] += 1
        yield from issue_qk()
        if ttot > 1:
            yield from issue_qk()
        for c in range(ttot):
            i, g = divmod(c, nkt)
            yield ("wait", pready[c & 1], (c >> 1) & 1, c >> 1)
            if g == 0 and i > 0:
                yield ("wait", ofree, (i - 1) & 1, i - 1)
            yield ("mma", f"pv{c}", rng.randint(5, 40))
            yield ("commit", kvfree[c % stages])
            bar, _ = od(c)
            yield ("commit", bar)
            if c + 2 < ttot:
                yield from issue_qk()

    def softmax(w):
        half0 = w < n_half0
        c = 0
        for i in range(n_items):
            yield ("wait", qfull[i & 1], (i >> 1) & 1, i >> 1)
            yield ("arrive", qfree[i & 1])
            for g in range(nkt):
                yield ("wait", sfull[c & 1], (c >> 1) & 1, c >> 1)
                yield ("check", f"qk{c}" in sim.done_mma, f"reads S of tile {c} before QK^T completed")
                lo, hi = (30, 90) if w == slow_warp else (5, 25)
                yield ("delay", rng.randint(lo, hi))
                if half0 and g > 0 and rng.random() < 0.15:      # rare in-TMEM rescale of O
                    bar, ph = od(c - 1)
                    yield ("wait", bar, ph & 1, ph)
                    yield ("check", all(f"pv{k}" in sim.done_mma for k in range(i * nkt, c)),
                           f"rescales O at tile {c} before the earlier P.V retired")
                yield ("arrive", pready[c & 1])
                c += 1
            if half0:                                            # per-item epilogue
                bar, ph = od(c - 1)
                yield ("wait", bar, ph & 1, ph)
                yield ("check", all(f"pv{k}" in sim.done_mma for k in range(i * nkt, c)),
                       f"reads O of item {i} before all of its P.V retired")
                yield ("check", f"pv{c}" not in sim.done_mma and not any(op[1] == f"pv{c}" for op in sim.pipe if op[0] == "mma"),
                       f"item {i}: next item's first P.V already issued while O is still being read")
                yield ("arrive", ofree)
                yield ("delay", rng.randint(5, 30))

    sim.add("stager", stager())
    sim.add("producer", producer())
    sim.add("mma", mma())
    for w in range(nsw):
        sim.add(f"softmax{w}", softmax(w))
    sim.run()
    return sim.violations
