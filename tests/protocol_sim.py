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


def attention8_cta(seed, nkt, n_items=2, nq=2, stages=8, single_pvdone=False, slow_warp=None, rescale_p=0.2):
    """One persistent CTA of attn8_kernel<NQ> (csrc/attention_v8.cuh): a producer, one MMA-issuing thread per query tile,
    four softmax warps per query tile; S double buffered per tile (QK^T(c+2) is issued right after P.V(c)), K/V ring shared by
    the NQ issuers (kv_free counts NQ arrivals), Q double buffer per item. `single_pvdone` reproduces the version with ONE
    "P.V done" barrier per query tile (advancing once per key tile), which let the epilogue read O one tile early."""
    sim = Sim(seed)
    rng = sim.rng
    B = Barrier
    sfull = [[B(f"sfull{t}{b}", 1) for b in range(2)] for t in range(nq)]
    pready = [[B(f"pready{t}{b}", 4) for b in range(2)] for t in range(nq)]
    pvdone = [[B(f"pvdone{t}{b}", 1) for b in range(2)] for t in range(nq)]
    ofree = [B(f"ofree{t}", 4) for t in range(nq)]
    kvfull = [B(f"kvfull{i}", 1) for i in range(stages)]
    kvfree = [B(f"kvfree{i}", nq) for i in range(stages)]
    qfull = [B("qfull0", 1), B("qfull1", 1)]
    qfree = [B("qfree0", nq), B("qfree1", nq)]
    ttot = nkt * n_items

    def pvd(t, c):      # barrier, phase index meaning "P.V of tile c of query tile t retired"
        return (pvdone[t][0], c) if single_pvdone else (pvdone[t][c & 1], c >> 1)

    def producer():
        c = 0
        for i in range(n_items):
            qb, u = i & 1, i >> 1
            if u > 0:
                yield ("wait", qfree[qb], (u & 1) ^ 1, u - 1)
                yield ("check", all(f"qk{t}_{k}" in sim.done_mma for t in range(nq) for k in range((i - 2) * nkt, (i - 1) * nkt)),
                       f"Q buffer {qb} overwritten for item {i} before every QK^T of item {i - 2} retired")
            yield ("delay", rng.randint(1, 40))
            yield ("arrive", qfull[qb])
            for g in range(nkt):
                st, use = c % stages, c // stages
                if use > 0:
                    yield ("wait", kvfree[st], (use & 1) ^ 1, use - 1)
                    yield ("check", all(f"pv{t}_{c - stages}" in sim.done_mma for t in range(nq)),
                           f"K/V stage {st} overwritten for tile {c} before P.V of tile {c - stages} retired in every query tile")
                yield ("delay", rng.randint(1, 30))
                yield ("arrive", kvfull[st])
                c += 1

    def issuer(t):
        cur = [0]

        def issue_qk():
            c = cur[0]
            i, g = divmod(c, nkt)
            if g == 0:
                yield ("wait", qfull[i & 1], (i >> 1) & 1, i >> 1)
            yield ("wait", kvfull[c % stages], (c // stages) & 1, c // stages)
            yield ("mma", f"qk{t}_{c}", rng.randint(3, 12))
            yield ("commit", sfull[t][c & 1])
            if g == nkt - 1:
                yield ("commit", qfree[i & 1])
            cur[0] += 1
        if ttot > 0:
            yield from issue_qk()
        if ttot > 1:
            yield from issue_qk()
        for c in range(ttot):
            i, g = divmod(c, nkt)
            yield ("wait", pready[t][c & 1], (c >> 1) & 1, c >> 1)
            if g == 0 and i > 0:
                yield ("wait", ofree[t], (i - 1) & 1, i - 1)
            yield ("mma", f"pv{t}_{c}", rng.randint(5, 40))
            bar, _ = pvd(t, c)
            yield ("commit", bar)
            yield ("commit", kvfree[c % stages])
            if c + 2 < ttot:
                yield from issue_qk()

    def softmax(t, qq):
        c = 0
        slow = slow_warp == (t, qq)
        for i in range(n_items):
            for g in range(nkt):
                yield ("wait", sfull[t][c & 1], (c >> 1) & 1, c >> 1)
                yield ("check", f"qk{t}_{c}" in sim.done_mma, f"reads S of tile {c} before QK^T completed")
                yield ("delay", rng.randint(30, 90) if slow else rng.randint(5, 25))
                if g > 0 and rng.random() < rescale_p:           # the reference of some row moved: rescale O in TMEM
                    bar, ph = pvd(t, c - 1)
                    yield ("wait", bar, ph & 1, ph)
                    yield ("check", all(f"pv{t}_{k}" in sim.done_mma for k in range(i * nkt, c)),
                           f"rescales O at tile {c} before the earlier P.V retired")
                yield ("arrive", pready[t][c & 1])
                c += 1
            bar, ph = pvd(t, c - 1)
            yield ("wait", bar, ph & 1, ph)
            yield ("check", all(f"pv{t}_{k}" in sim.done_mma for k in range(i * nkt, c)),
                   f"reads O of item {i} before all of its P.V retired")
            yield ("check", f"pv{t}_{c}" not in sim.done_mma and not any(op[1] == f"pv{t}_{c}" for op in sim.pipe if op[0] == "mma"),
                   f"item {i}: next item's first P.V already issued while O is still being read")
            yield ("arrive", ofree[t])
            yield ("delay", rng.randint(5, 30))

    sim.add("producer", producer())
    for t in range(nq):
        sim.add(f"issuer{t}", issuer(t))
        for qq in range(4):
            sim.add(f"softmax{t}.{qq}", softmax(t, qq))
    sim.run()
    return sim.violations


def gemm_cta(seed, tiles, kblocks, stages=4, n_epi=12, resident_b=False, early_release=True, tempty_count=None):
    """One persistent CTA of gemm_tc_kernel / gemm_tc_ws_kernel (csrc/gemm_tc.cuh): TMA producer, one MMA-issuing thread,
    `n_epi` epilogue warps; operand ring full / empty barriers, TMEM accumulator double buffer tfull / tempty (tempty counts
    the epilogue warps; with `early_release` a warp hands the buffer back as soon as its last tcgen05.ld has returned, before
    the math and the store). `resident_b`: the weight block is loaded once behind its own barrier (weight-stationary form)."""
    sim = Sim(seed)
    rng = sim.rng
    B = Barrier
    full = [B(f"full{i}", 1) for i in range(stages)]
    empty = [B(f"empty{i}", 1) for i in range(stages)]
    tfull = [B("tfull0", 1), B("tfull1", 1)]
    tempty = [B("tempty0", tempty_count or n_epi), B("tempty1", tempty_count or n_epi)]
    bfull = B("bfull", 1)
    read_done = set()              # (warp, tile): accumulator of `tile` is in this warp's registers

    def producer():
        if resident_b:
            yield ("delay", rng.randint(5, 60))
            yield ("arrive", bfull)
        i = 0
        for tile in range(tiles):
            for kb in range(kblocks):
                st, use = i % stages, i // stages
                yield ("wait", empty[st], (use & 1) ^ 1, use - 1)
                if use > 0:
                    prev = i - stages
                    yield ("check", f"mma{prev}" in sim.done_mma, f"ring slot {st} refilled for block {i} before the MMAs of block {prev} retired")
                yield ("delay", rng.randint(1, 40))
                yield ("arrive", full[st])
                i += 1

    def mma():
        if resident_b:
            yield ("wait", bfull, 0, 0)
        i = 0
        for tile in range(tiles):
            buf, use = tile & 1, tile >> 1
            yield ("wait", tempty[buf], (use & 1) ^ 1, use - 1)
            if use > 0:
                yield ("check", all((w, tile - 2) in read_done for w in range(n_epi)),
                       f"accumulator {buf} overwritten for tile {tile} before every epilogue warp read tile {tile - 2}")
            for kb in range(kblocks):
                st = i % stages
                yield ("wait", full[st], (i // stages) & 1, i // stages)
                yield ("mma", f"mma{i}", rng.randint(2, 10))
                yield ("commit", empty[st])
                i += 1
            yield ("commit", tfull[buf])

    def epilogue(w):
        for tile in range(tiles):
            buf, use = tile & 1, tile >> 1
            yield ("wait", tfull[buf], use & 1, use)
            yield ("check", all(f"mma{k}" in sim.done_mma for k in range(tile * kblocks, (tile + 1) * kblocks)),
                   f"epilogue reads tile {tile} before its MMAs retired")
            yield ("delay", rng.randint(2, 12))                   # tcgen05.ld of both chunks
            read_done.add((w, tile))
            if early_release:
                yield ("arrive", tempty[buf])
            yield ("delay", rng.randint(5, 60))                   # bias / GELU / pack / staging / bulk store
            if not early_release:
                yield ("arrive", tempty[buf])

    sim.add("producer", producer())
    sim.add("mma", mma())
    for w in range(n_epi):
        sim.add(f"epi{w}", epilogue(w))
    sim.run()
    return sim.violations
