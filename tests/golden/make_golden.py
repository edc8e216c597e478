"""Generates the golden fixtures in tests/golden/ from the CPU oracle (oracle/lsf_oracle.c) run on
the reference's own two inputs (/root/reference/cube40.stl, twoCube10.stl).  Run in the build
container (the reference tree is not present on the GPU box); takes ~2.5 minutes.

    python tests/golden/make_golden.py

Outputs (all numpy .npz, arrays in Fortran order):
  cube40_mesh.npz / twoCube10_mesh.npz : surfX (float32 is exact: STL data), surfElem -- the result
        of stlRead (subs.f90:17-121), so GPU-box tests do not need the STL files
  cube40_fields.npz  : phi after the sign search, after reinit #1 (2155 sweeps), after min/max flow
        (406 iterations), with the RMS histories and exit iterations
  twoCube10_fields.npz : sign field, RMS history up to the NaN at n=272, phi after sweep n=271
PARITY UNPINNED: these are outputs of the restatement, cross-checked against the independent
survey-time transcription, not of the gfortran binary.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DX = 0.05


def mesh(name):
    X, E = O.stl_read(f"{REF}/{name}.stl")
    assert np.array_equal(X.astype(np.float32).astype(np.float64), X)
    np.savez_compressed(f"{OUT}/{name}_mesh.npz", surfX=X.astype(np.float32), surfElem=E)
    return X, E


def main():
    # ---- cube40: the full default pipeline (BASELINE config 1) -------------------------------
    X, E = mesh("cube40")
    g = O.grid_from_surface(X, DX)
    phi = np.ones((g["nx"] + 1, g["ny"] + 1, g["nz"] + 1), order="F")
    O.sign_init(phi, g["xLo"], DX, X, E, g["box"])
    sign = phi.copy(order="F")
    st, n1, h1 = O.reinit(phi, 10000, DX, 0.1 * g["dxx"])
    assert st == 0 and n1 == 2154, (st, n1)
    reinit1 = phi.copy(order="F")
    st, n2, h2, nb, sb = O.minmax(phi, 10000, DX, 0.01 * g["dxx"])
    assert st == 0 and n2 == 406, (st, n2)
    minmax = phi.copy(order="F")
    st, n3, h3 = O.reinit(phi, 2000, DX, 0.001 * g["dxx"])
    assert st == 0 and n3 == 0
    np.savez_compressed(f"{OUT}/cube40_fields.npz", sign=sign, reinit1=reinit1, minmax=minmax, reinit2=phi,
                        rms_reinit1=h1, rms_minmax=h2, rms_reinit2=h3, n_exit=np.array([n1, n2, n3]),
                        phiNB=nb.astype(np.int8), phiSB=sb.astype(np.int8))
    # ---- twoCube10: NaN STOP at n = 272 (BASELINE config 2) ------------------------------------
    X, E = mesh("twoCube10")
    g = O.grid_from_surface(X, DX)
    phi = np.ones((g["nx"] + 1, g["ny"] + 1, g["nz"] + 1), order="F")
    O.sign_init(phi, g["xLo"], DX, X, E, g["box"])
    sign = phi.copy(order="F")
    st, n, h = O.reinit(phi.copy(order="F"), 10000, DX, 0.1 * g["dxx"])
    assert st == 1 and n == 272, (st, n)
    st2, n271, h271 = O.reinit(phi, 271, DX, 0.1 * g["dxx"])          # state after sweep n = 271
    assert st2 == 2 and np.array_equal(h271, h[:272])
    np.savez_compressed(f"{OUT}/twoCube10_fields.npz", sign=sign, rms_reinit1=h, n_nan=np.array([n]), phi_n271=phi)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
