"""Generates the golden fixtures in tests/golden/ by RUNNING THE REFERENCE PROGRAM -- set3d.f90 + subs.f90,
machine-translated to C by oracle/f90_to_c.py (oracle/_ref/libref.so; no Fortran compiler exists in the build
image) -- on the reference's own two inputs (/root/reference/cube40.stl, twoCube10.stl), stage by stage, and
checks at every stage that the hand-written C oracle (oracle/lsf_oracle.c) produces the same bits.  Run in the
build container (the reference tree is not present on the GPU box); takes ~8 minutes.

    python tests/golden/make_golden.py

Outputs (numpy .npz, arrays in Fortran order):
  cube40_mesh.npz / twoCube10_mesh.npz : surfX (float32 is exact: STL data), surfElem -- stlRead (subs.f90:17-121)
  cube40_fields.npz  : phi after the sign search (set3d.f90:176-268), after reinit #1 (:308, 2155 sweeps), after the
        min/max flow (:394-462, 406 iterations), after reinit #2 (:582); RMS histories as PRINTed; band masks;
        the node projection (:465-501): surfXX, phiSurf, gradPhiSurf
  twoCube10_fields.npz : sign field, RMS history up to the NaN at n=272 (STOP, subs.f90:926), phi after sweep n=271
  cube40_files.npz : sha256 + sizes of signedDistanceFunction.vti / smoothedDistanceFunction.vti / cube40.s3d as the
        translated program wrote them, the .vti header bytes, the first lines of the .s3d
  REF_PIN_REPORT.txt : what was compared and the verdict of each comparison
"""
import hashlib
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from oracle import ref as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DX = 0.05
P = R.Program
report = []


def check(what, ok):
    report.append(f"{'OK  ' if ok else 'FAIL'} {what}")
    print(report[-1], flush=True)
    assert ok, what


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    t0 = time.time()
    assert R.build(), "oracle/_ref/libref.so could not be built"
    tmp = tempfile.mkdtemp(prefix="lsf_ref_")
    # =========================================================== cube40: the whole default program (config 1)
    import shutil
    for nm in ("cube40", "twoCube10"):              # the program writes <stl basename>.s3d next to its input (set3d.f90:66)
        shutil.copy(f"{REF}/{nm}.stl", tmp)
    p = P(outdir=tmp)
    p.set_arg(f"{tmp}/cube40.stl")
    assert len(f"{tmp}/cube40.stl") <= 80           # CHARACTER filename*80
    assert p.run(*P.IMPORT) == 0
    X, E = p.get("surfX"), p.get("surfElem")
    Xo, Eo = O.stl_read(f"{REF}/cube40.stl")
    check("cube40 stlRead: surfX, surfElem (ref == oracle)", np.array_equal(X, Xo) and np.array_equal(E, Eo))
    assert np.array_equal(X.astype(np.float32).astype(np.float64), X)
    np.savez_compressed(f"{OUT}/cube40_mesh.npz", surfX=X.astype(np.float32), surfElem=E)

    assert p.run(P.BBOX_GRID[0], P.SIGN[1]) == 0
    g = O.grid_from_surface(Xo, DX)
    box = tuple(p.get(k) for k in ("im", "ip", "jm", "jp", "km", "kp"))
    check("cube40 grid: nx,ny,nz,xLo,sub-box (ref == oracle)",
          (p.get("nx"), p.get("ny"), p.get("nz")) == (g["nx"], g["ny"], g["nz"]) and np.array_equal(p.get("xLo"), g["xLo"])
          and box == g["box"])
    sign = p.get("phi")
    phio = np.ones(sign.shape, order="F")
    O.sign_init(phio, g["xLo"], DX, Xo, Eo, g["box"])
    check("cube40 sign search set3d.f90:176-268: phi bit-identical incl. -0.0 (ref == oracle)",
          np.array_equal(sign, phio) and np.array_equal(np.signbit(sign), np.signbit(phio)))

    p.prints()
    assert p.run(*P.REINIT1) == 0
    n1s, h1 = p.iteration_history()
    reinit1 = p.get("phi")
    check("cube40 dxx, h (ref == oracle)", p.get("dxx") == g["dxx"] and p.get("h") == 0.1 * g["dxx"])
    st, n1, h1o = O.reinit(phio, 10000, DX, 0.1 * g["dxx"])
    check(f"cube40 reinit #1 subs.f90:717-931: EXIT at n={len(n1s)} (ref) == {n1} (oracle), printed RMS history identical, "
          "phi bit-identical", st == 0 and len(n1s) == n1 == 2154 and np.array_equal(h1, h1o[:n1]) and np.array_equal(reinit1, phio))

    assert p.run(*P.VTI1) == 0
    assert p.run(*P.BAND_INIT) == 0
    nb0, sb0 = p.get("phiNB"), p.get("phiSB")
    nbo, sbo = O.narrowband(phio, DX)
    check("cube40 narrowBand subs.f90:178-207 (ref == oracle)", np.array_equal(nb0, nbo) and np.array_equal(sb0, sbo))
    p.prints()
    assert p.run(*P.MINMAX) == 0
    n2s, h2 = p.iteration_history()
    minmax = p.get("phi")
    nb, sb = p.get("phiNB"), p.get("phiSB")
    st, n2, h2o, nbo, sbo = O.minmax(phio, 10000, DX, 0.01 * g["dxx"])
    check(f"cube40 min/max loop set3d.f90:386-463: EXIT at n={len(n2s) + 1} (ref) == {n2} (oracle), printed RMS identical, "
          "phi, phiNB, phiSB bit-identical",
          st == 0 and len(n2s) + 1 == n2 == 406 and np.array_equal(h2, h2o[: n2 - 1]) and np.array_equal(minmax, phio)
          and np.array_equal(nb, nbo) and np.array_equal(sb, sbo))

    assert p.run(*P.NODES) == 0
    XX, ps, gs = p.get("surfXX"), p.get("phiSurf"), p.get("gradPhiSurf")
    sto, XXo, pso, gso, moves = O.advect_nodes(phio, sbo, g["xLo"], DX, Xo, iter=1000, literal=False)
    check(f"cube40 node projection set3d.f90:465-501 ({moves} moves): surfXX, phiSurf, gradPhiSurf bit-identical (ref == oracle)",
          sto == 0 and np.array_equal(XX, XXo) and np.array_equal(ps, pso) and np.array_equal(gs, gso))

    assert p.run(P.ASYMPTOTIC[0], P.VTI2[1]) == 0
    p.prints()
    assert p.run(*P.REINIT2) == 0
    n3s, h3 = p.iteration_history()
    reinit2 = p.get("phi")
    st, n3, h3o = O.reinit(phio, 2000, DX, 0.001 * g["dxx"])
    check("cube40 reinit #2 set3d.f90:571-582: EXIT at n=0, phi bit-identical (ref == oracle)",
          st == 0 and n3 == 0 and len(n3s) == 0 and np.array_equal(reinit2, phio))
    assert p.run(*P.S3D) == 0
    np.savez_compressed(f"{OUT}/cube40_fields.npz", sign=sign, reinit1=reinit1, minmax=minmax, reinit2=reinit2,
                        rms_reinit1=h1o, rms_minmax=h2o, rms_reinit2=h3o, n_exit=np.array([n1, n2, n3]),
                        phiNB=nb.astype(np.int8), phiSB=sb.astype(np.int8),
                        surfXX=XX, phiSurf=ps, gradPhiSurf=gs, n_moves=np.array([moves]))
    files = {fn: os.path.join(tmp, fn) for fn in ("signedDistanceFunction.vti", "smoothedDistanceFunction.vti", "cube40.s3d")}
    vti1 = open(files["signedDistanceFunction.vti"], "rb").read()
    vti2 = open(files["smoothedDistanceFunction.vti"], "rb").read()
    mark = b'<AppendedData encoding="raw">\n_'
    hdr_len = vti1.index(mark) + len(mark)
    check(".vti #1 payload == phi after reinit #1, #2 == phi after min/max (raw Float64 after the 4-byte nbytePhi)",
          vti1[hdr_len + 4: hdr_len + 4 + reinit1.nbytes] == reinit1.tobytes(order="F")
          and vti2[hdr_len + 4: hdr_len + 4 + minmax.nbytes] == minmax.tobytes(order="F"))
    s3d = open(files["cube40.s3d"], "rb").read()
    np.savez_compressed(f"{OUT}/cube40_files.npz",
                        vti1_sha256=sha(files["signedDistanceFunction.vti"]), vti2_sha256=sha(files["smoothedDistanceFunction.vti"]),
                        vti_header=np.frombuffer(vti1[:hdr_len + 4], dtype=np.uint8), vti_tail=np.frombuffer(vti1[-40:], dtype=np.uint8),
                        vti_size=len(vti1), s3d_sha256=hashlib.sha256(s3d).hexdigest(), s3d_size=len(s3d),
                        s3d_head=np.frombuffer(s3d[:4096], dtype=np.uint8))
    # =========================================================== twoCube10: NaN STOP at n = 272 (config 2)
    p = P(outdir=tmp)
    p.set_arg(f"{tmp}/twoCube10.stl")
    assert p.run(P.IMPORT[0], P.SIGN[1]) == 0
    X, E = p.get("surfX"), p.get("surfElem")
    Xo, Eo = O.stl_read(f"{REF}/twoCube10.stl")
    check("twoCube10 stlRead (ref == oracle)", np.array_equal(X, Xo) and np.array_equal(E, Eo))
    np.savez_compressed(f"{OUT}/twoCube10_mesh.npz", surfX=X.astype(np.float32), surfElem=E)
    g = O.grid_from_surface(Xo, DX)
    sign = p.get("phi")
    phio = np.ones(sign.shape, order="F")
    O.sign_init(phio, g["xLo"], DX, Xo, Eo, g["box"])
    check("twoCube10 sign search bit-identical (ref == oracle)", np.array_equal(sign, phio) and np.array_equal(np.signbit(sign), np.signbit(phio)))
    p.prints()
    st = p.run(*P.REINIT1)
    ns, h = p.iteration_history()
    phin = p.get("phi")
    sto, no, ho = O.reinit(phio, 10000, DX, 0.1 * g["dxx"])
    check(f"twoCube10 reinit #1: STOP (status {st}) after printing n={ns[-1]} with RMS NaN; history identical; phi identical NaN-aware",
          st == 1 and sto == 1 and ns[-1] == no == 272 and np.isnan(h[-1]) and np.array_equal(h[:-1], ho[:272])
          and np.array_equal(phin, phio, equal_nan=True))
    phi271 = sign.copy(order="F")
    st2, n271, h271 = R.reinit(phi271, 271, DX, 0.1 * g["dxx"])      # subs.f90 reinit called directly: state after sweep 271
    phio = sign.copy(order="F")
    st2o, _, h271o = O.reinit(phio, 271, DX, 0.1 * g["dxx"])
    check("twoCube10 reinit with iter=271 (all 272 sweeps, no exit): phi bit-identical (ref == oracle)",
          st2 == 2 and st2o == 2 and np.array_equal(phi271, phio) and np.array_equal(h271, h271o))
    np.savez_compressed(f"{OUT}/twoCube10_fields.npz", sign=sign, rms_reinit1=ho, n_nan=np.array([no]), phi_n271=phi271)
    report.append(f"total {time.time() - t0:.0f} s")
    with open(f"{OUT}/REF_PIN_REPORT.txt", "w") as f:
        f.write("Reference program (oracle/_ref/libref.so = f90_to_c.py translation of /root/reference/set3d.f90 + subs.f90)\n"
                "vs the hand-written oracle (oracle/lsf_oracle.c); generated by tests/golden/make_golden.py\n\n" + "\n".join(report) + "\n")
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
