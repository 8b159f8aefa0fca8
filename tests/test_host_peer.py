"""Host-side model of the index arithmetic of csrc/peer.cu (hpvg_peer_allreduce_avg_tensors): the per-CTA flag exchange only orders
what CTA b of one rank and CTA b of another rank touch, so the kernel is correct only if
  * the slots CTA b of rank q gathers into rank q's bucket are exactly the slots CTA b of the other ranks pull from it,
  * the slots CTA b of rank r pushes into rank q's bucket are exactly the slots CTA b of rank q scatters back to its gradients,
  * every slot of every tensor is gathered and scattered exactly once.
The loops below restate the kernel's (peer_copy_slots, peer_reduce_slice, peer_grid) with the library's constants; the bucket layout
comes from the library itself (hpvg_peer_bucket_numel).  No GPU needed."""
import numpy as np
import pytest

from hpvg import lib

THREADS, UNROLL, MAX_BLOCKS = 256, 4, 64      # PEER_THREADS, PEER_UNROLL, PEER_MAX_BLOCKS of csrc/peer.cu


def _layout(numels, world):
    bucket = int(lib.load().hpvg_peer_bucket_numel(len(numels), lib.longlong_array(numels), world))
    start4 = np.concatenate([[0], np.cumsum([(n + 3) // 4 for n in numels])])
    n4 = bucket // (4 * world)
    grid = min(MAX_BLOCKS, max(1, -(-n4 // (THREADS * UNROLL))))
    return bucket, start4, n4, grid


def _copy_slots(b, grid, n4, world, total4):
    """slots CTA b gathers / scatters (peer_copy_slots): for every slice q, i = b * T + t, stepping by grid * T"""
    out = []
    for q in range(world):
        for t in range(THREADS):
            i = b * THREADS + t
            while i < n4:
                idx = q * n4 + i
                if idx >= total4:
                    break
                out.append(idx)
                i += grid * THREADS
    return out


def _reduce_slots(b, grid, n4, rank):
    """slots CTA b of `rank` pulls from every bucket and pushes into every bucket (peer_reduce_slice)"""
    out = []
    step = grid * THREADS
    for t in range(THREADS):
        i0 = b * THREADS + t
        while i0 < n4:
            for k in range(UNROLL):
                i = i0 + k * step
                if i < n4:
                    out.append(rank * n4 + i)
            i0 += step * UNROLL
    return out


@pytest.mark.parametrize("numels,world", [([110592, 64, 5184, 7, 5], 2), ([110592, 64, 5184, 7, 5], 8), ([3], 4), ([64] * 37 + [110592] * 6, 8),
                                          ([110592] * 12 + [3, 1728, 64, 64], 3), ([1_000_003], 8)])
def test_cta_b_gathers_and_scatters_what_cta_b_of_the_peers_pulls_and_pushes(numels, world):
    bucket, start4, n4, grid = _layout(numels, world)
    total4 = int(start4[-1])
    assert bucket == n4 * 4 * world and n4 * world >= total4 and bucket % (4 * world) == 0
    owner_copy = np.full(n4 * world, -1)
    for b in range(grid):
        slots = _copy_slots(b, grid, n4, world, total4)
        assert len(set(slots)) == len(slots)
        assert (owner_copy[slots] == -1).all()      # gathered / scattered by one CTA only
        owner_copy[slots] = b
    assert (owner_copy[:total4] >= 0).all() and (owner_copy[total4:] == -1).all()      # every tensor slot, no padding slot
    for rank in range(world):
        seen = np.zeros(n4 * world, dtype=int)
        for b in range(grid):
            slots = np.array(_reduce_slots(b, grid, n4, rank), dtype=int)
            if slots.size == 0:
                continue
            seen[slots] += 1
            real = slots[slots < total4]
            # the CTA that exchanges a slot with the peers is the CTA that gathered it before the first flag exchange and scatters it
            # after the second one — on every rank, because the mapping depends on the slot's position inside its slice only
            assert (owner_copy[real] == b).all()
        lo, hi = rank * n4, (rank + 1) * n4
        assert (seen[lo:hi] == 1).all() and seen[:lo].sum() == 0 and seen[hi:].sum() == 0      # slice `rank`, each slot once


def test_slot_lookup_of_the_gather_matches_the_layout():
    """the kernel's binary search (largest t with start4[t] <= slot) and its tail handling of tensors whose size is not a multiple of 4"""
    numels = [7, 5, 64, 3, 110592, 1]
    _, start4, _, _ = _layout(numels, 2)
    flat = np.concatenate([np.arange(n, dtype=np.float64) + 1000 * t for t, n in enumerate(numels)])
    offs = np.concatenate([[0], np.cumsum(numels)])
    bucket = np.zeros(int(start4[-1]) * 4)
    for idx in range(int(start4[-1])):
        lo, hi = 0, len(numels) - 1
        while lo < hi:
            mid = (lo + hi + 1) >> 1
            if start4[mid] <= idx:
                lo = mid
            else:
                hi = mid - 1
        off = (idx - int(start4[lo])) * 4
        left = numels[lo] - off
        assert left > 0
        take = min(4, left)
        bucket[idx * 4:idx * 4 + take] = flat[offs[lo] + off:offs[lo] + off + take]
    # scatter back with the same arithmetic: identity on every tensor
    for t, n in enumerate(numels):
        got = bucket[int(start4[t]) * 4:int(start4[t]) * 4 + n]
        assert np.array_equal(got, flat[offs[t]:offs[t] + n]), t
