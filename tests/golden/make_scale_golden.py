"""Bench-scale parity fixtures: the CPU oracle (oracle/lsf_oracle.c, pinned bit for bit to the machine-translated
reference, see make_golden.py) run ONCE in the build container on grids that are too slow to sweep on the GPU box's
host cores inside a test: BASELINE config 3's geometry (sphere) and config 4's (torus + disjoint cube) at 256^3,
and a 512^3 reinit (config 3's size).  Stored per case: the input mesh, sha256 of every field (bit-exact checks need no
more), the RMS histories, and a strided sample of each field (for messages when a hash differs).

    python tests/golden/make_scale_golden.py [case ...]        # ~25 min of CPU in total

Cases
  sphere256  : 512-triangle sphere -> sign search (set3d.f90:176-268), 16 reinit sweeps = every raster twice
               (subs.f90:717-931), 10 min/max iterations (set3d.f90:394-462)
  torcube256 : torus + cube, 9 984 triangles (the bench geometry, stl.torus_cube_config) -> the same three stages
  reinit512  : analytic sphere sign field on 512^3, 8 sweeps = every raster once
A reduced part of each case (2 sweeps) is cross-checked against libref.so itself when it is available.
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from levelsetfortran_b200 import stl  # noqa: E402  (mesh generators and grid set-up only: host-side numpy)

OUT = os.path.dirname(os.path.abspath(__file__))
DX = 0.05


def sha(a):
    return hashlib.sha256(np.asfortranarray(a).tobytes(order="F")).hexdigest()


def sample(a):
    return np.ascontiguousarray(a[::8, ::8, ::8])


def analytic_sphere(n, dx=DX):
    """phiSign(gM = 1) of a sphere: + - * / sqrt only, so the field is reproducible bit for bit on any IEEE host."""
    c = (n - 1) * 0.5
    i = np.arange(n, dtype=np.float64)
    x, y, z = i[:, None, None] - c, i[None, :, None] - c * 1.03125, i[None, None, :] - c * 0.96875
    d = (np.sqrt(x * x + y * y + z * z) - 0.3125 * n) * dx
    return np.asfortranarray(d / np.sqrt(d * d + dx * dx))


def mesh_case(name, tris, sweeps=16, mm_iters=10):
    t0 = time.time()
    X, E = stl.dedup_nodes(tris)
    g = stl.grid_from_surface(X, DX)
    go = O.grid_from_surface(X, DX)
    assert (g["nx"], g["ny"], g["nz"], g["box"]) == (go["nx"], go["ny"], go["nz"], go["box"]) and np.array_equal(g["xLo"], go["xLo"])
    phi = np.ones((g["nx"] + 1, g["ny"] + 1, g["nz"] + 1), order="F")
    O.sign_init(phi, g["xLo"], DX, X, E, g["box"])
    sign = phi.copy(order="F")
    print(name, "sign search", round(time.time() - t0), "s", flush=True)
    h = 0.1 * g["dxx"]
    st, n1, h1 = O.reinit(phi, sweeps - 1, DX, h, tol=0.0)
    assert st == 2 and n1 == sweeps - 1
    reinit = phi.copy(order="F")
    print(name, "reinit", round(time.time() - t0), "s", flush=True)
    st, n2, h2, nb, sb = O.minmax(phi, mm_iters, DX, 0.01 * g["dxx"], tol=0.0)
    assert st == 2 and n2 == mm_iters, (st, n2)
    np.savez_compressed(f"{OUT}/scale_{name}.npz", surfX=X.astype(np.float32), surfElem=E, shape=np.array(sign.shape), h=h,
                        h1=0.01 * g["dxx"], sweeps=sweeps, mm_iters=mm_iters,
                        sign_sha=sha(sign), reinit_sha=sha(reinit), minmax_sha=sha(phi), nb_sha=sha(nb), sb_sha=sha(sb),
                        rms_reinit=h1, rms_minmax=h2, sign_s=sample(sign), reinit_s=sample(reinit), minmax_s=sample(phi),
                        n_neg=int((sign < 0).sum()), n_zero=int((sign == 0).sum()))
    print(name, "done", round(time.time() - t0), "s", flush=True)
    return sign, g


def reinit_case(name, n, sweeps=8):
    t0 = time.time()
    phi = analytic_sphere(n)
    in_sha = sha(phi)
    h = 0.1 * DX / (np.sqrt(3.0) * (n - 22.2) * DX)
    st, n1, h1 = O.reinit(phi, sweeps - 1, DX, h, tol=0.0)
    assert st == 2
    np.savez_compressed(f"{OUT}/scale_{name}.npz", n=n, h=h, sweeps=sweeps, input_sha=in_sha, reinit_sha=sha(phi), rms_reinit=h1,
                        reinit_s=sample(phi))
    print(name, "done", round(time.time() - t0), "s", flush=True)


def ref_crosscheck(sign, g):
    """2 sweeps of the translated reference itself on the same sign field == oracle (it writes gradPhi too: 4x the memory)"""
    from oracle import ref as R
    if not R.build():
        return
    a, b = sign.copy(order="F"), sign.copy(order="F")
    sa, na, ha = R.reinit(a, 1, DX, 0.1 * g["dxx"])
    sb, nb, hb = O.reinit(b, 1, DX, 0.1 * g["dxx"], tol=0.0)
    assert np.array_equal(a, b) and np.array_equal(ha[:2], hb[:2]), "libref != oracle at scale"
    print("libref == oracle on", sign.shape, flush=True)


def main():
    cases = sys.argv[1:] or ["sphere256", "torcube256", "reinit512"]
    if "sphere256" in cases:
        sign, g = mesh_case("sphere256", stl.sphere_config(256, DX, n_lat=17, n_lon=16))
        ref_crosscheck(sign, g)
    if "torcube256" in cases:
        mesh_case("torcube256", stl.torus_cube_config((256, 256, 256), DX))
    if "reinit512" in cases:
        reinit_case("reinit512", 512)


if __name__ == "__main__":
    main()
