#!/usr/bin/env python
"""Discrete-event model of allreduce_sgd_kernel's cross-GPU protocol (csrc/peer.cu): `world` ranks x `G` CTAs, one
launch per epoch and rank (launches of a rank are stream-ordered: launch e+1 starts when every CTA of launch e has
finished), staging double-buffered by epoch parity, per-(rank, CTA) flags that only ever grow, remote stores that
become visible after random delays (data before the releasing flag store, otherwise unordered), CTAs that start and
run under random delays with rare very long stalls (a rank whose host is late with the next graph launch).

Checked on every access: a reducer reads from every rank's `in` buffer exactly the piece that rank published in THIS
epoch (not yet overwritten by epoch + 2, not a leftover of epoch - 2); the update reads from `out` exactly the sums of
this epoch; every run terminates.  This is DESIGN.md s5's argument, executed.  `--mutate no_flag_b` (update without
waiting for the peers' second flag) and `--mutate early_flag_a` (first flag raised before the pieces are published) show
that the model finds what the flags are there for.  (The model also passes with ONE staging buffer: with the two-phase
exchange the per-CTA flags and the launch boundary already order every reuse; the parity double-buffering is kept as
slack, it costs 1.5 MB.)

    python tools/peer_protocol_sim.py --trials 500
"""
import argparse
import heapq
import random


class ProtocolError(AssertionError):
    pass


class PeerSim:
    def __init__(self, world, G, epochs, rng, buffers=2, stall_p=0.03, mutate=None):
        self.W, self.G, self.E, self.rng, self.nbuf, self.stall_p = world, G, epochs, rng, buffers, stall_p
        self.mutate = mutate
        self.now, self.events, self.seq = 0, [], 0
        W = world
        # staging of rank r: inn[r][buf][(slice, cta)] = (epoch, writer rank) ; out[r][buf][(slice, cta)] = (epoch, owner)
        self.inn = [[{} for _ in range(buffers)] for _ in range(W)]
        self.out = [[{} for _ in range(buffers)] for _ in range(W)]
        self.flag_a = [[[0] * G for _ in range(W)] for _ in range(W)]      # flag_a[dst][src][cta]
        self.flag_b = [[[0] * G for _ in range(W)] for _ in range(W)]
        self.state = [0] * W                                                # last completed epoch per rank
        self.tickets = [0] * W
        self.blocked = []            # (predicate, resume)
        self.finished = 0

    def delay(self, lo=1, hi=60):
        d = self.rng.randint(lo, hi)
        if self.rng.random() < self.stall_p:
            d += self.rng.randint(500, 20000)
        return d

    def at(self, t, fn):
        self.seq += 1
        heapq.heappush(self.events, (t, self.seq, fn))

    def run_gen(self, gen):
        try:
            cmd = next(gen)
        except StopIteration:
            return
        if cmd[0] == "delay":
            self.at(self.now + cmd[1], lambda: self.run_gen(gen))
        else:                                       # ("wait", predicate)
            if cmd[1]():
                self.at(self.now + self.delay(1, 5), lambda: self.run_gen(gen))
            else:
                self.blocked.append((cmd[1], gen))

    def wake(self):
        still = []
        for pred, gen in self.blocked:
            if pred():
                self.at(self.now + self.delay(1, 5), lambda gen=gen: self.run_gen(gen))
            else:
                still.append((pred, gen))
        self.blocked = still

    # remote store: visible after a delay; returns the delivery time
    def remote(self, fn, not_before=0):
        t = max(self.now + self.delay(5, 400), not_before)
        self.at(t, fn)
        return t

    def cta(self, r, b):
        W, G = self.W, self.G
        epoch = self.state[r] + 1
        buf = epoch % self.nbuf
        yield ("delay", self.delay())
        # 1. publish my pieces of every slice into MY in[buf] (local stores, visible to peers before the flag)
        if self.mutate == "early_flag_a":
            for q in range(W):
                self.remote(lambda q=q: self.flag_a[q][r].__setitem__(b, max(self.flag_a[q][r][b], epoch)))
            yield ("delay", self.delay(50, 500))
        for s in range(W):
            self.inn[r][buf][(s, b)] = (epoch, r)
        yield ("delay", self.delay())
        if self.mutate != "early_flag_a":
            for q in range(W):
                self.remote(lambda q=q: self.flag_a[q][r].__setitem__(b, max(self.flag_a[q][r][b], epoch)))
        yield ("wait", lambda: all(self.flag_a[r][q][b] >= epoch for q in range(W)))
        # 2. reduce piece (r, b) over all ranks, broadcast the sum into every rank's out[buf]
        for q in range(W):
            yield ("delay", self.delay(1, 30))
            got = self.inn[q][buf].get((r, b))
            if got != (epoch, q):
                raise ProtocolError("rank %d cta %d epoch %d reads in[%d] of rank %d: holds %s" % (r, b, epoch, buf, q, got))
        last = 0
        for q in range(W):
            last = max(last, self.remote(lambda q=q: self.out[q][buf].__setitem__((r, b), (epoch, r))))
        yield ("delay", self.delay())
        for q in range(W):        # release: the flag becomes visible after the data stores it covers
            self.remote(lambda q=q: self.flag_b[q][r].__setitem__(b, max(self.flag_b[q][r][b], epoch)), not_before=last + 1)
        if self.mutate != "no_flag_b":
            yield ("wait", lambda: all(self.flag_b[r][q][b] >= epoch for q in range(W)))
        # 3. update my pieces of every slice from the broadcast sums
        for s in range(W):
            yield ("delay", self.delay(1, 30))
            got = self.out[r][buf].get((s, b))
            if got != (epoch, s):
                raise ProtocolError("rank %d cta %d epoch %d reads out[%d] slice %d: holds %s" % (r, b, epoch, buf, s, got))
        # 4. the last CTA closes the epoch; the rank's next launch follows in stream order
        self.tickets[r] += 1
        if self.tickets[r] == G:
            self.tickets[r] = 0
            self.state[r] = epoch
            if epoch < self.E:
                self.at(self.now + self.delay(1, 200), lambda: self.launch(r))
            else:
                self.finished += 1

    def launch(self, r):
        for b in range(self.G):               # CTAs become resident at different times
            self.at(self.now + self.delay(1, 100), lambda b=b: self.run_gen(self.cta(r, b)))

    def run(self):
        for r in range(self.W):
            self.at(self.delay(1, 300), lambda r=r: self.launch(r))
        while self.events:
            t, _, fn = heapq.heappop(self.events)
            self.now = t
            fn()
            self.wake()
        if self.finished != self.W:
            raise ProtocolError("deadlock: ranks at epochs %s, %d waiters" % (self.state, len(self.blocked)))


def trial(seed, buffers=2, mutate=None):
    rng = random.Random(seed)
    PeerSim(rng.choice([1, 2, 3, 4, 8]), rng.choice([1, 2, 5]), rng.choice([1, 2, 5, 9]), rng, buffers=buffers, mutate=mutate).run()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=500)
    ap.add_argument("--single-buffer", action="store_true")
    ap.add_argument("--mutate", choices=["no_flag_b", "early_flag_a"], default=None)
    args = ap.parse_args()
    bad, first = 0, None
    for seed in range(args.trials):
        try:
            trial(seed, 1 if args.single_buffer else 2, args.mutate)
        except ProtocolError as e:
            bad += 1
            first = first or "seed %d: %s" % (seed, e)
    print("%s staging%s: %d / %d schedules failed%s" % ("single" if args.single_buffer else "double-buffered",
                                                       ", mutation %s" % args.mutate if args.mutate else "", bad, args.trials,
                                                     "" if not bad else "   e.g. " + first))


if __name__ == "__main__":
    main()
