#!/usr/bin/env python
"""Model of the sweep kernel's tile schedule (lsf_march.cuh): ntb x ntc column tiles of `steps` steps each, handed out
by ticket in anti-diagonal front order to `ncta` persistent CTAs; a tile may execute step t only when both predecessors
(J-1,K), (J,K-1) have PUBLISHED step t + chunk - 1 + lag_tile (checked at chunk starts, published at chunk ends).
All CTAs run at one step per time unit.  Prints the makespan against the two lower bounds (work / CTAs, critical path).

    python tools/sim_tile_schedule.py 1024 512 --chunk 8 4 2
"""
import argparse
import heapq


def simulate(n, tile=16, ncta=296, chunk=8):
    nt = (n - 2 + tile - 1) // tile
    steps = (n - 2) + 2 * (tile - 1) + 3 + 4          # tend + look-ahead prologue (lsf_march.cuh: tend, M_LOOK)
    order = [(J, s - J) for s in range(2 * nt - 1) for J in range(nt) if 0 <= s - J < nt]
    # published progress of a tile as a function of time: it publishes step c*chunk-1 when it has finished that step
    start = {}      # tile -> list of (time at which chunk c begins)
    finish_chunk = {}   # tile -> list of times at which chunk c ends (published)
    nchunks = (steps + chunk - 1) // chunk
    free = [(0.0, c) for c in range(ncta)]
    heapq.heapify(free)
    makespan = 0.0
    wait_total = 0.0
    for (J, K) in order:
        t0, cta = heapq.heappop(free)
        t = t0
        ends = []
        for c in range(nchunks):
            need = c * chunk + chunk - 1 + tile            # predecessor step that must be published
            for pred in ((J - 1, K), (J, K - 1)):
                if pred[0] < 0 or pred[1] < 0:
                    continue
                pe = finish_chunk[pred]
                pc = min(nchunks - 1, need // chunk)       # the chunk whose end publishes >= need (last chunk publishes FIN)
                if pe[pc] > t:
                    wait_total += pe[pc] - t
                    t = pe[pc]
            t += min(chunk, steps - c * chunk)
            ends.append(t)
        finish_chunk[(J, K)] = ends
        heapq.heappush(free, (t, cta))
        makespan = max(makespan, t)
    work = nt * nt * steps / ncta
    crit = (2 * nt - 2) * (tile + chunk - 1) + steps
    return dict(n=n, tiles=nt * nt, steps=steps, chunk=chunk, makespan=makespan, work_bound=work, critical_path=crit,
                efficiency=work / makespan, wait_share=wait_total / (makespan * ncta))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("grids", type=int, nargs="+")
    ap.add_argument("--chunk", type=int, nargs="+", default=[8])
    ap.add_argument("--tile", type=int, default=16)
    ap.add_argument("--ncta", type=int, default=296)
    a = ap.parse_args()
    for n in a.grids:
        for ch in a.chunk:
            r = simulate(n, a.tile, a.ncta, ch)
            print("grid %5d^3 tile %2d chunk %2d: tiles %5d x %4d steps | makespan %8.0f  work/CTAs %8.0f  critical path %6d | "
                  "efficiency %.3f  CTA time spent waiting %.3f" % (r["n"], a.tile, ch, r["tiles"], r["steps"], r["makespan"],
                                                                     r["work_bound"], r["critical_path"], r["efficiency"], r["wait_share"]))
