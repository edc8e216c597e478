#!/usr/bin/env python
"""Smallest possible GPU check of the overlapped-sweeps path (no torch import): parity against the oracle on two small
grids (exact arithmetic, bit-identical; one with a tolerance EXIT inside a batch), then the sweep rate at 512^3 with
the option off / on."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
t00 = time.time()
from levelsetfortran_b200 import _lib, set_subs as S  # noqa: E402
from oracle import oracle as O  # noqa: E402
from conftest import dist_field, synth_field  # noqa: E402

DX = 0.05
S.set_arith(True)
S.set_overlap(True)
for shape, iters in (((40, 38, 36), 20), ((19, 50, 33), 12)):
    p0 = synth_field(shape, seed=7)
    a, b = p0.copy(order="F"), p0.copy(order="F")
    st, n, hist = O.reinit(a, iters, DX, 0.0014)
    n2, hist2 = S.reinit(b, None, None, shape[0] - 1, shape[1] - 1, shape[2] - 1, iters, DX, 0.0014)
    print("PARITY", shape, "n", n, n2, "bitwise", np.array_equal(a, b), "hist", np.allclose(hist, hist2, rtol=1e-12), flush=True)
p0 = dist_field((30, 28, 26), seed=3, noise=0.0005)
a, b = p0.copy(order="F"), p0.copy(order="F")
st, n, hist = O.reinit(a, 60, DX, 0.000345)
n2, hist2 = S.reinit(b, None, None, 29, 27, 25, 60, DX, 0.000345)
print("PARITY tolerance exit: n", n, n2, "bitwise", np.array_equal(a, b), flush=True)
S.set_arith(None)
n = int(os.environ.get("SHOT_GRID", "512"))
x = ((np.arange(n) - n / 2.0 + 0.37) * DX).astype(np.float64)
r = np.sqrt((x * x)[:, None, None] + (x * x)[None, :, None] + (x * x)[None, None, :]) - 0.3 * n * DX
phi0 = np.asfortranarray(r / np.sqrt(r * r + DX * DX))
del r
res = {}
for on in (False, True, False, True):
    S.set_overlap(on)
    g = S.DeviceGrid(n - 1, n - 1, n - 1)
    g.upload(phi0)
    g.reinit(7, DX, 1e-4, tol=0.0)
    ms = []
    for _ in range(3):
        rc, ne, h = g.reinit(7, DX, 1e-4, tol=0.0)
        ms.append(_lib.last_timing()[0])
    out = g.download() if on not in res else None
    g.close()
    if out is not None:
        res[on] = out
    print("RATE grid %d overlap %s: %.2f ms per 8 sweeps -> %.2f Gcell/s" % (n, on, min(ms), 8 * (n - 2) ** 3 / min(ms) / 1e6), flush=True)
print("SAME RESULT with / without overlap:", np.array_equal(res[False], res[True]), " total %.1f s" % (time.time() - t00))
