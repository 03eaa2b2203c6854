#!/usr/bin/env python
"""EXHAUSTIVE companion of tc_protocol_sim.py: the same model of tc_gemm_kernel's synchronisation (producer, MMA
issuer, two splitter groups, 3-stage ring, 2 TMEM A slots, mbarriers with arrival counts + transaction units and
parity waits), but instead of sampling schedules it explores EVERY interleaving of the agents' steps and of the
asynchronous completions (TMA landings, cp.async landings, MMA retirements in issue order) by breadth-first search over
the reachable states -- no notion of time at all, so any agent may be arbitrarily slow relative to any other.  For a
given number of chunks it either proves the four properties (no read before the right chunk has landed, no overwrite
under a reader or before consumption, no surplus arrival, no deadlock) or prints a shortest counter-example.

The protocol repeats with period lcm(3 stages, 2 groups) = 6 chunks; runs of up to 9-10 chunks cover the start-up, two
periods and the drain.  `--wpg 2` models two warps per splitter group (skew inside a group, named barrier) for smaller
runs.

    python tools/tc_protocol_exhaustive.py --rule own_only --mode all --chunks 9
    python tools/tc_protocol_exhaustive.py --rule observe_all --mode all --chunks 7     # finds the round-2 deadlock
"""
import argparse
from collections import deque

S, ASLOTS = 3, 2


class Violation(Exception):
    pass


def build_programs(n, gathered, rule, wpg):
    gathers = any(gathered(c) for c in range(n))
    prod = []
    for c in range(n):
        prod += [("wait", "E", c % S, ((c // S) & 1) ^ 1), ("issue", c)]
    mma = []
    for c in range(n):
        mma += [("wait", "F", c % S, (c // S) & 1), ("wait", "Y", c % S, (c // S) & 1), ("mma", c)]
    warps = []
    for g in range(2):
        for w in range(wpg):
            ops = []
            if gathers:
                for c in range(min(S, n)):
                    if (c + S) % ASLOTS == g:
                        ops.append(("refill", c, g, w))
            for c in range(n):
                s, it = c % S, c // S
                own = c % ASLOTS == g
                if rule == "observe_all" or (rule == "kernel" and not gathers):
                    ops.append(("wait", "F", s, it & 1))
                    if not own:
                        continue
                else:
                    if not own:
                        continue
                    if c >= S and rule != "own_naive":
                        ops.append(("wait", "E", s, (it - 1) & 1))
                    ops.append(("wait", "F", s, it & 1))
                ops += [("read", c), ("readdone", c)]
                if c >= ASLOTS:
                    cp = c - ASLOTS
                    ops += [("wait", "E", cp % S, (cp // S) & 1), ("check_retired", cp)]
                ops.append(("writeA", c, g))
                if gathers:
                    ops += [("bar_arrive", g), ("bar_wait", g, sum(1 for cc in range(c + 1) if cc % ASLOTS == g))]
                    if c + S < n:
                        ops.append(("refill", c + S, g, w))
            if n > 0:
                ops.append(("wait", "A", 0, 0))
            warps.append(tuple(ops))
    return tuple(prod), tuple(mma), tuple(warps), gathers


def explore(n, gathered, rule, wpg=1, limit=30_000_000):
    prod, mma, warps, gathers = build_programs(n, gathered, rule, wpg)
    progs = (prod, mma) + warps
    na = len(progs)
    full_count = 1 + wpg if gathers else 1
    counts = {"F": full_count, "E": 1, "Y": wpg, "A": 1}

    # state: pcs, bars {(kind, s): (phase, pending, tx)}, X[s] (chunk, parts, readers, read_done), W[s] (chunk, landed, inuse,
    # consumed), A[a] (chunk, parts, consumed), pending async events (sorted tuple), retired count, bar arrivals per group
    def bar_init():
        b = {}
        for s in range(S):
            b[("F", s)] = (0, full_count, 0)
            b[("E", s)] = (0, 1, 0)
            b[("Y", s)] = (0, wpg, 0)
        b[("A", 0)] = (0, 1, 0)
        return tuple(sorted(b.items()))

    init = (tuple([0] * na), bar_init(), tuple([(-1, 0, 0, wpg)] * S), tuple([(-1, 0, 0, 1)] * S),
            tuple([(-1, 0, 1)] * ASLOTS), (), 0, (0, 0))

    def arrive(bars, key, n_arr=1, tx=0):
        phase, pending, t = bars[key]
        if pending < n_arr:
            raise Violation("%s%d: arrival beyond the phase's count" % key)
        pending -= n_arr
        t += tx
        if pending == 0 and t == 0:
            phase, pending = phase + 1, counts[key[0]]
        bars[key] = (phase, pending, t)

    def complete_tx(bars, key, tx):
        phase, pending, t = bars[key]
        t -= tx
        if pending == 0 and t == 0:
            phase, pending = phase + 1, counts[key[0]]
        bars[key] = (phase, pending, t)

    def begin_x_write(X, s, c, who):
        chunk, parts, readers, rd = X[s]
        if readers or rd != wpg:
            raise Violation("%s overwrites X of stage %d (chunk %d): %d readers, %d/%d warps have read it" % (who, s, chunk, readers, rd, wpg))
        X[s] = (c, 0, 0, 0)

    def successors(st):
        pcs, bars_t, X_t, W_t, A_t, pend, retired, barc = st
        out = []
        # agent steps
        for a in range(na):
            pc = pcs[a]
            if pc >= len(progs[a]):
                continue
            op = progs[a][pc]
            bars, X, W, A = dict(bars_t), list(X_t), list(W_t), list(A_t)
            pend2, retired2, barc2 = pend, retired, barc
            k = op[0]
            if k == "wait":
                phase = bars[(op[1], op[2])][0]
                if (phase & 1) == op[3]:
                    continue                      # blocked
            elif k == "issue":
                c = op[1]; s = c % S
                wc, wl, wu, wcons = W[s]
                if wu or not wcons:
                    raise Violation("producer overwrites W of stage %d (chunk %d) still needed" % (s, wc))
                W[s] = (c, 0, 0, 0)
                tma_x = not gathered(c)
                arrive(bars, ("F", s), 1, 2 if tma_x else 1)
                ev = [("landW", c)]
                if tma_x:
                    begin_x_write(X, s, c, "TMA")
                    ev.append(("landX", c))
                pend2 = tuple(sorted(pend + tuple(ev)))
            elif k == "mma":
                c = op[1]; s = c % S; sl = c % ASLOTS
                if W[s][0] != c or W[s][1] != 1:
                    raise Violation("MMA of chunk %d reads W of stage %d = chunk %d, landed %d" % (c, s, W[s][0], W[s][1]))
                if A[sl][0] != c or A[sl][1] != wpg:
                    raise Violation("MMA of chunk %d reads A slot %d = chunk %d, %d/%d written" % (c, sl, A[sl][0], A[sl][1], wpg))
                W[s] = (c, 1, 1, 0)
                pend2 = tuple(sorted(pend + (("retire", c),)))
            elif k == "refill":
                c, g, w = op[1], op[2], op[3]; s = c % S
                if not gathered(c):
                    arrive(bars, ("F", s))
                else:
                    if X[s][0] != c:
                        begin_x_write(X, s, c, "refill")
                    pend2 = tuple(sorted(pend + (("landR", c, g, w),)))
            elif k == "read":
                c = op[1]; s = c % S
                if X[s][0] != c or X[s][1] != wpg:
                    raise Violation("splitter reads X of stage %d for chunk %d: holds chunk %d, %d/%d landed" % (s, c, X[s][0], X[s][1], wpg))
                X[s] = (X[s][0], X[s][1], X[s][2] + 1, X[s][3])
            elif k == "readdone":
                s = op[1] % S
                X[s] = (X[s][0], X[s][1], X[s][2] - 1, X[s][3] + 1)
            elif k == "check_retired":
                if retired <= op[1]:
                    raise Violation("A slot rewritten before the MMAs of chunk %d retired" % op[1])
            elif k == "writeA":
                c, g = op[1], op[2]
                if A[g][0] != c:
                    if not A[g][2]:
                        raise Violation("A slot %d (chunk %d) overwritten before it was consumed" % (g, A[g][0]))
                    A[g] = (c, 0, 0)
                A[g] = (c, A[g][1] + 1, 0)
                arrive(bars, ("Y", c % S))
            elif k == "bar_arrive":
                g = op[1]
                barc2 = (barc[0] + 1, barc[1]) if g == 0 else (barc[0], barc[1] + 1)
            elif k == "bar_wait":
                g, gen = op[1], op[2]
                if barc[g] < gen * wpg:
                    continue
            pcs2 = pcs[:a] + (pc + 1,) + pcs[a + 1:]
            out.append(("agent %d: %s" % (a, (op,)), (pcs2, tuple(sorted(bars.items())), tuple(X), tuple(W), tuple(A), pend2, retired2, barc2)))
        # asynchronous completions
        for i, ev in enumerate(pend):
            bars, X, W, A = dict(bars_t), list(X_t), list(W_t), list(A_t)
            retired2 = retired
            k = ev[0]
            if k == "retire":
                if ev[1] != retired:
                    continue                      # tcgen05.commit: MMAs retire in issue order
                c = ev[1]; s = c % S
                W[s] = (W[s][0], W[s][1], 0, 1)
                A[c % ASLOTS] = (A[c % ASLOTS][0], A[c % ASLOTS][1], 1)
                retired2 = retired + 1
                arrive(bars, ("E", s))
                if c == n - 1:
                    arrive(bars, ("A", 0))
            elif k == "landW":
                c = ev[1]; s = c % S
                if W[s][0] != c:
                    raise Violation("W data of chunk %d lands in stage %d that holds chunk %d" % (c, s, W[s][0]))
                W[s] = (c, 1, W[s][2], W[s][3])
                complete_tx(bars, ("F", s), 1)
            elif k == "landX":
                c = ev[1]; s = c % S
                if X[s][0] != c:
                    raise Violation("X data of chunk %d lands in stage %d that holds chunk %d" % (c, s, X[s][0]))
                X[s] = (c, wpg, X[s][2], X[s][3])
                complete_tx(bars, ("F", s), 1)
            elif k == "landR":
                c = ev[1]; s = c % S
                if X[s][0] != c:
                    raise Violation("refill data of chunk %d lands in stage %d that holds chunk %d" % (c, s, X[s][0]))
                X[s] = (c, X[s][1] + 1, X[s][2], X[s][3])
                arrive(bars, ("F", s))
            pend2 = pend[:i] + pend[i + 1:]
            out.append(("event %s" % (ev,), (pcs, tuple(sorted(bars.items())), tuple(X), tuple(W), tuple(A), pend2, retired2, barc)))
        return out

    seen = {init: None}
    queue = deque([init])

    def trace(st, last):
        steps = [last]
        while seen[st] is not None:
            st, label = seen[st]
            steps.append(label)
        return list(reversed(steps))

    while queue:
        st = queue.popleft()
        try:
            succ = successors(st)
        except Violation as v:
            return {"ok": False, "states": len(seen), "why": str(v), "trace": trace(st, "-> " + str(v))}
        if not succ:
            if any(pc < len(p) for pc, p in zip(st[0], progs)) or st[5]:
                stuck = [progs[a][pc] for a, pc in enumerate(st[0]) if pc < len(progs[a])]
                return {"ok": False, "states": len(seen), "why": "deadlock, waiting: %s" % (stuck,), "trace": trace(st, "-> deadlock")}
            continue
        for label, nx in succ:
            if nx not in seen:
                seen[nx] = (st, label)
                queue.append(nx)
                if len(seen) > limit:
                    return {"ok": None, "states": len(seen), "why": "state limit reached"}
    return {"ok": True, "states": len(seen)}


def mode_fn(mode, n, head=None):
    if mode == "none":
        return lambda c: False
    if mode == "all":
        return lambda c: True
    k = head if head is not None else max(1, n // 2)
    return lambda c: c < k


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rule", choices=["kernel", "own_only", "observe_all", "own_naive"], default="kernel",
                    help="kernel = what gemm_tc.cu does: own_only in CTAs on refill duty, observe_all in TMA-only CTAs")
    ap.add_argument("--mode", choices=["none", "all", "head"], default="all")
    ap.add_argument("--chunks", type=int, default=8)
    ap.add_argument("--head", type=int, default=None, help="mode head: number of gathered chunks")
    ap.add_argument("--wpg", type=int, default=1, help="warps per splitter group in the model")
    ap.add_argument("--trace", action="store_true")
    ap.add_argument("--stages", type=int, default=3, help="ring depth (the kernel is built with 3)")
    args = ap.parse_args()
    global S
    S = args.stages
    r = explore(args.chunks, mode_fn(args.mode, args.chunks, args.head), args.rule, args.wpg)
    print("rule %s, %d stages, X gathered: %s, %d chunks, %d warp(s)/group: %s after %d states%s"
          % (args.rule, S, args.mode, args.chunks, args.wpg, {True: "ALL SCHEDULES OK", False: "VIOLATION", None: "inconclusive"}[r["ok"]],
             r["states"], "" if r["ok"] else ": " + r["why"]))
    if args.trace and r.get("trace"):
        for t in r["trace"]:
            print("   ", t)


if __name__ == "__main__":
    main()
