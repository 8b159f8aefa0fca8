"""Exhaustive interleaving check of the flag protocol of csrc/peer.cu (compute-sanitizer's racecheck is closed on this pool, and it
does not see across GPUs anyway).  One CTA per rank is modelled — CTAs with different numbers never touch the same slots
(tests/test_host_peer.py) — as the sequence the kernel executes per call c = 1, 2, ...:

    gather   : write slice s of the OWN bucket, one step per slice                      (peer_copy_slots<true>)
    signal A : flagA[q][r] <- c for every rank q                                         (peer_barrier, phase 0)
    wait A   : until flagA[r][q] >= c for every q
    pull     : read slice r of EVERY bucket, one step per bucket                         (peer_reduce_slice, loads)
    push     : write the mean into slice r of EVERY bucket, one step per bucket          (peer_reduce_slice, stores)
    signal B : flagB[q][r] <- c                                                          (peer_barrier, phase 1)
    wait B   : until flagB[r][q] >= c
    scatter  : read slice s of the own bucket, one step per slice                        (peer_copy_slots<false>)

Every bucket slice carries a tag saying who wrote it in which call.  All interleavings of the ranks' steps are explored (depth-first
with memoisation); a pull must find the owner's gather of the SAME call, a scatter must find the slice owner's mean of the SAME call.
The mutants at the bottom (a barrier removed, flags compared for equality with a stale value) must be caught by the same checker."""
import sys

import pytest


def _program(world, calls, skip=()):
    """the step list of one rank: (kind, call, argument)"""
    prog = []
    for c in range(1, calls + 1):
        prog += [("gather", c, s) for s in range(world)]
        if "signal_a" not in skip:
            prog.append(("signal_a", c, None))
        if "wait_a" not in skip:
            prog.append(("wait_a", c, None))
        prog += [("pull", c, q) for q in range(world)]
        prog += [("push", c, q) for q in range(world)]
        if "signal_b" not in skip:
            prog.append(("signal_b", c, None))
        if "wait_b" not in skip:
            prog.append(("wait_b", c, None))
        prog += [("scatter", c, s) for s in range(world)]
    return prog


def _explore(world, calls, skip=()):
    """-> None when every interleaving is safe and terminates, else a description of the first violation found"""
    prog = _program(world, calls, skip)
    n = len(prog)
    # state: (pcs, flagA, flagB, buckets); buckets[q][s] = tag of slice s of rank q's bucket; flags[q][r] = last call rank r signalled to q
    init = (tuple([0] * world), tuple(tuple([0] * world) for _ in range(world)), tuple(tuple([0] * world) for _ in range(world)),
            tuple(tuple([("init", 0)] * world) for _ in range(world)))
    seen = set()
    stack = [init]
    sys.setrecursionlimit(10000)
    while stack:
        state = stack.pop()
        if state in seen:
            continue
        seen.add(state)
        pcs, fa, fb, buckets = state
        if all(pc == n for pc in pcs):
            continue
        progressed = False
        for r in range(world):
            pc = pcs[r]
            if pc == n:
                continue
            kind, c, arg = prog[pc]
            nfa, nfb, nb = fa, fb, buckets
            if kind == "gather":
                nb = _set(buckets, r, arg, ("grad", r, c))
            elif kind == "signal_a":
                nfa = tuple(_row_set(fa[q], r, c) for q in range(world))
            elif kind == "wait_a":
                if any(fa[r][q] < c for q in range(world)):
                    continue      # blocked
            elif kind == "pull":
                if buckets[arg][r] != ("grad", arg, c):
                    return "call %d: rank %d pulls slice %d of rank %d's bucket and finds %r" % (c, r, r, arg, buckets[arg][r])
            elif kind == "push":
                nb = _set(buckets, arg, r, ("mean", r, c))
            elif kind == "signal_b":
                nfb = tuple(_row_set(fb[q], r, c) for q in range(world))
            elif kind == "wait_b":
                if any(fb[r][q] < c for q in range(world)):
                    continue
            elif kind == "scatter":
                if buckets[r][arg] != ("mean", arg, c):
                    return "call %d: rank %d scatters slice %d of its bucket and finds %r" % (c, r, arg, buckets[r][arg])
            progressed = True
            stack.append((_row_set(pcs, r, pc + 1), nfa, nfb, nb))
        if not progressed:
            return "deadlock at program counters %r" % (pcs,)
    return None


def _row_set(row, i, v):
    return row[:i] + (v,) + row[i + 1:]


def _set(buckets, q, s, tag):
    return _row_set(buckets, q, _row_set(buckets[q], s, tag))


@pytest.mark.parametrize("world,calls", [(2, 3), (3, 2)])
def test_every_interleaving_of_the_flag_protocol_is_safe(world, calls):
    assert _explore(world, calls) is None


@pytest.mark.parametrize("skip", [("wait_a",), ("wait_b",), ("signal_a", "wait_a"), ("signal_b", "wait_b")])
def test_the_checker_catches_a_missing_barrier(skip):
    """without the first exchange a pull can overtake the peer's gather; without the second one a scatter can overtake a peer's push,
    and the next call's gather can overwrite a slice a slow peer has not pulled yet"""
    found = _explore(2, 2, skip)
    assert found is not None and "deadlock" not in found, found
