"""Sweep-kernel time on a constant field (every cell far from the surface: all differences 0, the sum of squares under the root
exactly 0) vs a smeared-sign sphere field, fp64 FAST arithmetic, one GPU.  usage: python tools/probe_farfield.py [n]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from levelsetfortran_b200 import DeviceGrid, set_subs as S, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
S.set_arith("fast")
G = DeviceGrid(n - 1, n - 1, n - 1)
x = np.arange(n, dtype=np.float64)
r = np.sqrt(((x[:, None, None] - n / 2.1) ** 2 + (x[None, :, None] - n / 1.9) ** 2 + (x[None, None, :] - n / 2.2) ** 2)) * 0.05 - 0.3 * n * 0.05
sph = np.asfortranarray(r / np.sqrt(r * r + 0.05 ** 2))
for name, f in (("constant +1", np.asfortranarray(np.ones((n, n, n)))), ("sphere sign", sph)):
    G.upload(f)
    G.reinit(8, 0.05, 1.0e-4, tol=0.0)
    t0 = time.perf_counter()
    rc, ne, hist = G.reinit(16, 0.05, 1.0e-4, tol=0.0)
    dt = time.perf_counter() - t0
    print(f"{name:12s} n={n}: {dt / 17 * 1e3:.3f} ms per sweep (wall, 17 sweeps), {(n - 2) ** 3 * 17 / dt / 1e9:.2f} Gcell/s, last rms {hist[-1]:.3e}, arith {S.last_arith()}")
G.close()
