#!/usr/bin/env python
"""Model of the z-slab Gauss-Seidel pipeline: P ranks, each with its own 296 persistent CTAs,
tiles handed out in tilted front order m*J + K, a tile waiting for its two predecessors' published progress -- the
predecessor in c of tile row 0 being the upstream rank's last tile row -- and every rank starting sweep n+1 when its
own sweep n (+ the boundary kernel) is complete.  Eight sweeps = one raster cycle with its two k-direction flips
(subs.f90:742-852: rasters 1,6,7,8 ascend in k, 2..5 descend).  All CTAs run at one step per time unit.

    python tools/sim_slab_pipeline.py --ranks 1 2 4 8 --tilt 1 4 8 16
"""
import argparse
import heapq

TILE, CHUNK, NCTA = 16, 8, 296
GAP = 230          # boundary kernel + loop control + launch between two sweeps, in steps (0.6 ms at 2.6 us per step)


def fill_order(ntb, ntc, m):
    out = []
    for s in range(m * (ntb - 1) + ntc):
        K = s % m
        while K < ntc and K <= s:
            J = (s - K) // m
            if J < ntb:
                out.append((J, K))
            K += m
    return out


def sweep(ntb, ntc, steps, m, t_start, upstream):
    """One sweep on one rank.  upstream: per J, the chunk end times of the upstream rank's last tile row (or None).
    Returns (end time, per-J chunk end times of this rank's last tile row)."""
    nchunks = (steps + CHUNK - 1) // CHUNK
    ends = {}
    free = [(t_start, c) for c in range(NCTA)]
    heapq.heapify(free)
    t_end = t_start
    for (J, K) in fill_order(ntb, ntc, m):
        t, cta = heapq.heappop(free)
        e = []
        for c in range(nchunks):
            pc = min(nchunks - 1, (c * CHUNK + CHUNK - 1 + TILE) // CHUNK)
            preds = []
            if J > 0:
                preds.append(ends[(J - 1, K)])
            if K > 0:
                preds.append(ends[(J, K - 1)])
            elif upstream is not None:
                preds.append(upstream[J])
            for pe in preds:
                if pe[pc] > t:
                    t = pe[pc]
            t += min(CHUNK, steps - c * CHUNK)
            e.append(t)
        ends[(J, K)] = e
        heapq.heappush(free, (t, cta))
        t_end = max(t_end, t)
    return t_end, {J: ends[(J, ntc - 1)] for J in range(ntb)}


def cycle(P, m, n=1024, cycles=2):
    ntb = (n - 2 + TILE - 1) // TILE
    ntc = (n + TILE - 1) // TILE               # planes per rank (weak scaling: n per rank)
    steps = (n - 2) + 2 * (TILE - 1) + 7
    up = [+1, -1, -1, -1, -1, +1, +1, +1]      # k direction of rasters 1..8
    t_rank = [0.0] * P
    marks = []
    for s in range(8 * cycles):
        d = up[s % 8]
        order = range(P) if d > 0 else range(P - 1, -1, -1)
        prev = None
        for r in order:
            t_end, last_row = sweep(ntb, ntc, steps, m, t_rank[r], prev)
            t_rank[r] = t_end + GAP
            prev = last_row
        if s % 8 == 7:
            marks.append(max(t_rank))
    return (marks[-1] - marks[-2]) if cycles > 1 else marks[-1]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", type=int, nargs="+", default=[1, 2, 4, 8])
    ap.add_argument("--tilt", type=int, nargs="+", default=[1, 8])
    ap.add_argument("--grid", type=int, default=1024)
    a = ap.parse_args()
    base = cycle(1, 1, a.grid)
    print("one raster cycle (8 sweeps) on 1 rank, tilt 1: %.0f steps" % base)
    for m in a.tilt:
        for P in a.ranks:
            t = cycle(P, m, a.grid)
            print("tilt %2d  ranks %d : cycle %8.0f steps, weak-scaling efficiency vs (1 rank, tilt 1) %.3f" % (m, P, t, base / t))
