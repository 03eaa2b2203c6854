#!/usr/bin/env python
"""Discrete-event model of the synchronisation protocol of tc_gemm_kernel (csrc/gemm_tc.cu): one TMA producer, one MMA
issuer, two splitter groups of four warps, the 3-stage shared-memory ring, two TMEM A slots, and every mbarrier of the
kernel with the hardware's semantics (pending-arrival count + transaction bytes per phase, `try_wait.parity` can only
tell the current phase from the one before it).  Agents run as coroutines under randomly drawn delays, including rare
very long stalls (a warp held up behind the memory system), and the model checks on every step that

  * nobody reads a stage / TMEM slot before the right chunk has completely landed in it,
  * nobody overwrites a stage region / TMEM slot that still has a reader,
  * no mbarrier receives more arrivals than its phase expects,
  * the kernel terminates (no agent is left waiting once the event queue is empty).

It exists because the kernel cannot be run in the build container (no GPU) and its hazards are timing-dependent: with
the round-2 "refill duty" (the splitter group that just read a stage's X region re-loads it with cp.async) a group
spends memory-latency-bound time between two of its barrier waits, and the ORIGINAL wait rule -- every splitter warp
observes every chunk's full barrier, also the other group's -- can then be lapped: the barrier completes phase k+1
(whose arrivals do not depend on the late group) before the late group has asked for phase k, its parity wait turns
into a wait for phase k+2, which needs the late group's own work: deadlock.  `--rule observe_all` reproduces that;
`--rule own_only` is the rule the kernel uses now (wait for the stage's EMPTY barrier of chunk c-3, which pins the full
barrier's phase, then for the full barrier of chunk c; own chunks only), which no schedule can lap.  `--rule own_naive` (own chunks only, WITHOUT the empty
barrier) is round 1's first version, whose parity aliasing the model also finds: with an odd stage count a group meets
a stage every other phase and takes "phase k-1 done" for "phase k+1 done".

    python tools/tc_protocol_sim.py --rule own_only --trials 2000
"""
import argparse
import heapq
import random

STAGES, ASLOTS = 3, 2


class ProtocolError(AssertionError):
    pass


class MBar:
    def __init__(self, name, count):
        self.name, self.count = name, count
        self.pending, self.tx, self.phase = count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self, n=1):
        if self.pending < n:
            raise ProtocolError("%s: arrival beyond the phase's count (phase %d)" % (self.name, self.phase))
        self.pending -= n
        self._check()

    def expect_tx_arrive(self, nbytes):
        self.tx += nbytes
        self.arrive()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._check()

    def passed(self, parity):               # mbarrier.try_wait.parity: true iff the phase of that parity is the previous one
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, nchunks, gathered, rule, rng, stall_p=0.02, stall=(4000, 60000), stages=STAGES):
        """gathered(c) -> True when chunk c's X tile is loaded by the refill duty (cp.async), else by the producer's TMA"""
        self.n, self.gathered, self.rule, self.rng = nchunks, gathered, rule, rng
        self.S = stages
        self.stall_p, self.stall = stall_p, stall
        self.gathers = any(gathered(c) for c in range(nchunks))
        S = self.S
        self.full = [MBar("full[%d]" % s, 1 + 4 if self.gathers else 1) for s in range(S)]   # 4 = the 4 warps x 32 threads
        self.empty = [MBar("empty[%d]" % s, 1) for s in range(S)]
        self.yready = [MBar("yready[%d]" % s, 4) for s in range(S)]
        self.accum = MBar("accum", 1)
        self.now, self.events, self.seq = 0, [], 0
        # data model
        self.X = [{"chunk": None, "parts": 0, "readers": 0, "read_done": 4} for _ in range(S)]
        self.W = [{"chunk": None, "parts": 0, "in_use": False, "consumed": True} for _ in range(S)]
        self.A = [{"chunk": None, "parts": 0, "consumed": True} for _ in range(ASLOTS)]
        self.retired = set()
        self.last_retire = 0
        self.bar_sync = [[], []]
        self.blocked = {}
        self.done = set()
        self.agents = {}

    # ---- scheduling
    def delay(self, lo, hi):
        d = self.rng.randint(lo, hi)
        if self.rng.random() < self.stall_p:
            d += self.rng.randint(*self.stall)
        return d

    def at(self, t, fn):
        self.seq += 1
        heapq.heappush(self.events, (t, self.seq, fn))

    def spawn(self, name, gen):
        self.agents[name] = gen
        self.at(0, lambda: self.step(name))

    def step(self, name):
        gen = self.agents[name]
        try:
            cmd = next(gen)
        except StopIteration:
            self.done.add(name)
            return
        self.handle(name, cmd)

    def handle(self, name, cmd):
        kind = cmd[0]
        if kind == "delay":
            self.at(self.now + cmd[1], lambda: self.step(name))
        elif kind == "wait":
            bar, parity = cmd[1], cmd[2]
            if bar.passed(parity):
                self.at(self.now + self.delay(1, 4), lambda: self.step(name))      # a warp may stall anywhere
            else:
                self.blocked[name] = (bar, parity)
        elif kind == "barsync":
            g = cmd[1]
            self.bar_sync[g].append(name)
            if len(self.bar_sync[g]) == 4:
                names, self.bar_sync[g] = self.bar_sync[g], []
                for nm in names:
                    self.at(self.now + self.delay(1, 4), lambda nm=nm: self.step(nm))
        else:
            raise ValueError(kind)

    def wake(self):
        for name, (bar, parity) in list(self.blocked.items()):
            if bar.passed(parity):
                del self.blocked[name]
                self.at(self.now + self.delay(1, 4), lambda name=name: self.step(name))

    def run(self):
        while self.events:
            t, _, fn = heapq.heappop(self.events)
            self.now = t
            fn()
            self.wake()
        if len(self.done) != len(self.agents):
            stuck = {n: "%s parity %d (barrier in phase %d)" % (b.name, p, b.phase) for n, (b, p) in self.blocked.items()}
            raise ProtocolError("deadlock at t=%d: %s; waiting at bar.sync: %s" % (self.now, stuck, self.bar_sync))

    # ---- agents
    def producer(self):
        S = self.S
        for c in range(self.n):
            s, it = c % S, c // S
            yield ("wait", self.empty[s], (it & 1) ^ 1)
            yield ("delay", self.delay(20, 120))
            w = self.W[s]
            if w["in_use"] or not w["consumed"]:
                raise ProtocolError("producer overwrites W of stage %d (chunk %s) still needed" % (s, w["chunk"]))
            self.W[s] = {"chunk": c, "parts": 0, "in_use": False, "consumed": False}
            tma_x = not self.gathered(c)
            units = 3 if tma_x else 2
            self.full[s].expect_tx_arrive(units)
            if tma_x:
                self.begin_x_write(s, c, "TMA")
                self.at(self.now + self.delay(700, 3500), lambda s=s, c=c: (self.land_x(s, c, 4), self.full[s].complete_tx(1)))
            for _ in range(2):
                self.at(self.now + self.delay(700, 3500), lambda s=s, c=c: (self.land_w(s, c), self.full[s].complete_tx(1)))

    def begin_x_write(self, s, c, who):
        x = self.X[s]
        if x["readers"] or x["read_done"] != 4:
            raise ProtocolError("%s overwrites X of stage %d (chunk %s) while it is being read / before it was read "
                                "(readers %d, done %d)" % (who, s, x["chunk"], x["readers"], x["read_done"]))
        self.X[s] = {"chunk": c, "parts": 0, "readers": 0, "read_done": 0}

    def land_x(self, s, c, parts):
        x = self.X[s]
        if x["chunk"] != c:
            raise ProtocolError("X data of chunk %d lands in stage %d that holds chunk %s" % (c, s, x["chunk"]))
        x["parts"] += parts

    def land_w(self, s, c):
        w = self.W[s]
        if w["chunk"] != c:
            raise ProtocolError("W data of chunk %d lands in stage %d that holds %s" % (c, s, w["chunk"]))
        w["parts"] += 1

    def mma(self):
        S = self.S
        for c in range(self.n):
            s, it, a = c % S, c // S, c % ASLOTS
            yield ("wait", self.full[s], it & 1)
            yield ("wait", self.yready[s], it & 1)
            w, slot = self.W[s], self.A[a]
            if w["chunk"] != c or w["parts"] != 2:
                raise ProtocolError("MMA of chunk %d reads W of stage %d = chunk %s, %d/2 landed" % (c, s, w["chunk"], w["parts"]))
            if slot["chunk"] != c or slot["parts"] != 4:
                raise ProtocolError("MMA of chunk %d reads A slot %d = chunk %s, %d/4 written" % (c, a, slot["chunk"], slot["parts"]))
            w["in_use"] = True
            yield ("delay", self.delay(30, 80))
            retire = max(self.last_retire, self.now) + self.delay(700, 900)
            self.last_retire = retire

            def commit(c=c, s=s, a=a):
                self.retired.add(c)
                self.W[s]["in_use"] = False
                self.W[s]["consumed"] = True
                self.A[a]["consumed"] = True
                self.empty[s].arrive()
                if c == self.n - 1:
                    self.accum.arrive()
            self.at(retire, commit)

    def refill(self, c, warp_first):
        """one warp's share of the refill duty for chunk c (generator)"""
        s = c % self.S
        if not self.gathered(c):
            self.full[s].arrive()
            return
        yield ("delay", self.delay(100, 1500))        # ids (TN: global loads) + 8 LDGSTS per thread: memory-latency bound
        if warp_first[0]:
            self.begin_x_write(s, c, "refill")
            warp_first[0] = False
        elif self.X[s]["chunk"] != c:
            raise ProtocolError("refill of chunk %d: stage %d holds %s" % (c, s, self.X[s]["chunk"]))
        self.at(self.now + self.delay(600, 3000), lambda: (self.land_x(s, c, 1), self.full[s].arrive()))

    def splitter(self, g, q, first_flags):
        S = self.S
        if self.gathers:
            for c in range(min(S, self.n)):
                if (c + S) % ASLOTS == g:
                    yield from self.refill(c, first_flags.setdefault(c, [True]))
        for c in range(self.n):
            s, it = c % S, c // S
            if self.rule == "observe_all":
                yield ("wait", self.full[s], it & 1)
                if c % ASLOTS != g:
                    continue
            else:
                if c % ASLOTS != g:
                    continue
                if c >= S and self.rule == "own_only":
                    yield ("wait", self.empty[s], (it - 1) & 1)
                yield ("wait", self.full[s], it & 1)
            x = self.X[s]
            if x["chunk"] != c or x["parts"] != 4:
                raise ProtocolError("splitter g%d w%d reads X of stage %d for chunk %d: holds chunk %s, %d/4 landed"
                                    % (g, q, s, c, x["chunk"], x["parts"]))
            x["readers"] += 1
            yield ("delay", self.delay(80, 300))
            x["readers"] -= 1
            x["read_done"] += 1
            if c >= ASLOTS:
                cp = c - ASLOTS
                yield ("wait", self.empty[cp % S], (cp // S) & 1)
                if cp not in self.retired:
                    raise ProtocolError("splitter writes A slot %d for chunk %d before the MMAs of chunk %d retired" % (g, c, cp))
            slot = self.A[g]
            if slot["chunk"] != c:
                if not slot["consumed"]:
                    raise ProtocolError("A slot %d (chunk %s) overwritten before it was consumed" % (g, slot["chunk"]))
                self.A[g] = slot = {"chunk": c, "parts": 0, "consumed": False}
            yield ("delay", self.delay(40, 150))
            slot["parts"] += 1
            self.yready[s].arrive()
            if self.gathers:
                yield ("barsync", g)
                if c + S < self.n:
                    yield from self.refill(c + S, first_flags.setdefault(c + S, [True]))
        if self.n > 0:
            yield ("wait", self.accum, 0)

    def start(self):
        self.spawn("producer", self.producer())
        self.spawn("mma", self.mma())
        flags = {}
        for g in range(2):
            for q in range(4):
                self.spawn("split g%d w%d" % (g, q), self.splitter(g, q, flags))


def trial(seed, rule, mode, nchunks=None, stall_p=0.02, stages=STAGES):
    rng = random.Random(seed)
    n = nchunks if nchunks is not None else rng.choice([0, 1, 2, 3, 4, 5, 7, 12, 19, 27, 32, 38])
    if mode == "none":
        gathered = lambda c: False
    elif mode == "all":                     # weight gradient, column tile of the self half: every chunk is gathered
        gathered = lambda c: True
    else:                                   # forward: the first chunks (self half) are gathered, the rest comes by TMA
        k = rng.randint(1, max(1, n - 1)) if n > 1 else 1
        gathered = lambda c, k=k: c < k
    sim = Sim(n, gathered, rule, rng, stall_p=stall_p, stages=stages)
    sim.start()
    sim.run()
    return sim.now


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rule", choices=["observe_all", "own_only", "own_naive"], default="own_only")
    ap.add_argument("--trials", type=int, default=2000)
    ap.add_argument("--stall-p", type=float, default=0.02)
    args = ap.parse_args()
    for mode in ("none", "all", "head"):
        bad, first = 0, None
        for seed in range(args.trials):
            try:
                trial(seed, args.rule, mode, stall_p=args.stall_p)
            except ProtocolError as e:
                bad += 1
                first = first or "seed %d: %s" % (seed, e)
        print("rule %-11s  X tiles gathered: %-4s  %d / %d schedules failed%s"
              % (args.rule, mode, bad, args.trials, "" if not bad else "   e.g. " + first))


if __name__ == "__main__":
    main()
