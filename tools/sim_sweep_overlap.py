#!/usr/bin/env python
"""What would overlapping consecutive sweeps buy?  Model (same rules as tools/sim_tile_schedule.py) of one raster cycle
on one GPU where the CTAs go on to the tickets of sweep n+1 as soon as those of sweep n are handed out, a tile of sweep
n+1 additionally waiting until the 3x3 neighbourhood of tiles covering it has finished sweep n (its stencil reaches 3
cells into the neighbours; the boundary block and the RMS test are assumed to be folded into the tiles / decided late,
as the z-slab path already does).  Compared with the present one-launch-per-sweep schedule.

    python tools/sim_sweep_overlap.py 512 1024
"""
import heapq
import sys

TILE, CHUNK, NCTA, GAP = 16, 8, 296, 230
DIRS = [(+1, +1), (+1, -1), (-1, -1), (-1, -1), (+1, -1), (-1, +1), (+1, +1), (-1, +1)]   # (j, k) direction of rasters 1..8


def run(n, overlap):
    nt = (n - 2 + TILE - 1) // TILE
    steps = (n - 2) + 2 * (TILE - 1) + 7
    nchunks = (steps + CHUNK - 1) // CHUNK
    free = [(0.0, c) for c in range(NCTA)]
    heapq.heapify(free)
    done_prev = None                       # physical tile -> finish time in the previous sweep
    t_all = 0.0
    for s in range(8):
        dj, dk = DIRS[s]
        ends, done = {}, {}
        if not overlap:
            t0 = max(t for t, _ in free) + (GAP if s else 0)
            free = [(t0, c) for c in range(NCTA)]
            heapq.heapify(free)
        for f in range(2 * nt - 1):
            for J in range(nt):
                K = f - J
                if not 0 <= K < nt:
                    continue
                pj, pk = (J if dj > 0 else nt - 1 - J), (K if dk > 0 else nt - 1 - K)
                t, cta = heapq.heappop(free)
                if overlap and done_prev is not None:
                    for a in (-1, 0, 1):
                        for b in (-1, 0, 1):
                            q = (pj + a, pk + b)
                            if q in done_prev and done_prev[q] > t:
                                t = done_prev[q]
                e = []
                for c in range(nchunks):
                    pc = min(nchunks - 1, (c * CHUNK + CHUNK - 1 + TILE) // CHUNK)
                    for pred in ((J - 1, K), (J, K - 1)):
                        if pred in ends and ends[pred][pc] > t:
                            t = ends[pred][pc]
                    t += min(CHUNK, steps - c * CHUNK)
                    e.append(t)
                ends[(J, K)] = e
                done[(pj, pk)] = t
                heapq.heappush(free, (t, cta))
                t_all = max(t_all, t)
        done_prev = done
    return t_all, 8 * nt * nt * steps / NCTA


if __name__ == "__main__":
    for n in [int(a) for a in sys.argv[1:]] or [512, 1024]:
        a, w = run(n, False)
        b, _ = run(n, True)
        print("grid %4d^3: 8 sweeps, one launch each %8.0f steps (efficiency %.3f) | overlapped %8.0f steps (efficiency %.3f) | gain %.1f %%"
              % (n, a, w / a, b, w / b, 100 * (a / b - 1)))
