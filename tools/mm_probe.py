"""probe: time lsf_grid_minmax on the bench geometry several times (host wall + library event time)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from levelsetfortran_b200 import DeviceGrid, _lib, stl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
DX = 0.05
L = _lib.lib(); _lib.check(L.lsf_init(0))
X, E = stl.dedup_nodes(stl.torus_cube_config((n, n, n), DX))
g = stl.grid_from_surface(X, DX)
G = DeviceGrid(g["nx"], g["ny"], g["nz"])
for rep in range(4):
    G.fill(1.0); G.signSearch(g["xLo"], DX, X, E, g["box"])
    G.reinit(7, DX, 0.1 * g["dxx"], tol=0.0)
    if rep % 2: G.checksum()
    t0 = time.perf_counter()
    rc, ne, h = G.minMaxFlow(64, DX, 0.01 * g["dxx"], tol=0.0)
    wall = time.perf_counter() - t0
    ms, nl = _lib.last_timing()
    print("rep %d: minmax wall %.1f ms, library event time %.1f ms, launches %d, active %d" % (rep, wall * 1e3, ms, nl, L.lsf_last_minmax_active()), flush=True)
