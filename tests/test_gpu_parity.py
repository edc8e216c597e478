"""GPU parity tests (run on a B200 with `pytest -m gpu`).  Everything goes through the C ABI
(include/lsf_b200.h) via the set_subs mirror; the oracle / golden fixtures are only the checker.

Bars (BASELINE.json north_star): inside/outside sign field bit-exact; phi after init, reinit and
min/max flow within 1e-10 max-abs of the reference path in fp64; identical iteration counts.
The EXACT arithmetic mode and the whole min/max path are additionally required to be bit-identical.
"""
import numpy as np
import pytest

from conftest import GOLDEN, dist_field, load_mesh, synth_field

pytestmark = pytest.mark.gpu
DX = 0.05
TOL = 1.0e-10      # north_star: phi within 1e-10 max-abs in fp64


@pytest.fixture(scope="module")
def S(lsf):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from levelsetfortran_b200 import set_subs
    yield set_subs
    set_subs.set_arith(None)
    set_subs.set_sched(False)
    set_subs.set_minmax_algo(False)


def _mode(S, exact, plane, mm_march=False):
    S.set_arith(exact)
    S.set_sched(plane)
    S.set_minmax_algo(mm_march)


MM_ALGOS = {"list": (False, False), "march": (False, True), "plane": (True, False)}     # -> (plane, mm_march)


# ------------------------------------------------------------------------------------ sign search
@pytest.mark.parametrize("name", ["cube40", "twoCube10"])
def test_sign_search_bit_exact_on_reference_inputs(S, name):
    from levelsetfortran_b200 import stl
    X, E = load_mesh(name)
    g = stl.grid_from_surface(X, DX)
    gold = np.load(f"{GOLDEN}/{name}_fields.npz")["sign"]
    phi = np.ones(gold.shape, order="F")
    S.signSearch(phi, g["nx"], g["ny"], g["nz"], g["xLo"], DX, X, E, g["box"])
    assert np.array_equal(phi, gold)
    assert np.array_equal(np.signbit(phi), np.signbit(gold))          # includes -0.0 vs +0.0


def test_sign_search_bit_exact_synthetic(S, oracle):
    from levelsetfortran_b200 import stl
    X, E = stl.dedup_nodes(stl.torus_cube_config((48, 40, 56)))
    g = stl.grid_from_surface(X, DX)
    shape = (g["nx"] + 1, g["ny"] + 1, g["nz"] + 1)
    a = np.ones(shape, order="F")
    b = np.ones(shape, order="F")
    oracle.sign_init(a, g["xLo"], DX, X, E, g["box"])
    S.signSearch(b, g["nx"], g["ny"], g["nz"], g["xLo"], DX, X, E, g["box"])
    assert np.array_equal(a, b) and (a < 0).any() and (a == 1.0).any()


def test_sign_search_culling_is_exact_on_a_dense_mesh(S):
    """The tile-culled search (k_sign_search_tiled) against the brute-force kernel on the 20k-triangle sphere of
    BASELINE config 3 at reduced grid size: near the centre almost every centroid is a near-tie, so the
    candidate lists are long and the first-index rule is exercised; must be bit-identical."""
    import os
    from levelsetfortran_b200 import stl
    X, E = stl.dedup_nodes(stl.sphere_config(72, DX))
    g = stl.grid_from_surface(X, DX)
    shape = (g["nx"] + 1, g["ny"] + 1, g["nz"] + 1)
    a = np.ones(shape, order="F")
    b = np.ones(shape, order="F")
    S.signSearch(a, g["nx"], g["ny"], g["nz"], g["xLo"], DX, X, E, g["box"])
    os.environ["LSF_SIGN_BRUTE"] = "1"
    try:
        S.signSearch(b, g["nx"], g["ny"], g["nz"], g["xLo"], DX, X, E, g["box"])
    finally:
        del os.environ["LSF_SIGN_BRUTE"]
    assert np.array_equal(a, b) and np.array_equal(np.signbit(a), np.signbit(b))
    assert (a < 0).any() and (a > 0).any() and len(E) > 15000


# ------------------------------------------------------------------------------------ reinit
@pytest.mark.parametrize("plane", [False, True], ids=["march", "plane"])
@pytest.mark.parametrize("shape", [(22, 21, 23), (40, 38, 36), (19, 50, 33), (35, 18, 70), (6, 5, 7), (3, 3, 3)])
def test_reinit_sweeps_exact_mode_bitwise(S, oracle, shape, plane):
    """All 8 rasters + BC + RMS on small/ragged grids: exact arithmetic is bit-identical."""
    _mode(S, True, plane)
    p0 = synth_field(shape, seed=7)
    a, b = p0.copy(order="F"), p0.copy(order="F")
    nx, ny, nz = (s - 1 for s in shape)
    st, n, hist = oracle.reinit(a, 15, DX, 0.0014)                    # tol 1.E-5 as subs.f90:915
    n2, hist2 = S.reinit(b, None, None, nx, ny, nz, 15, DX, 0.0014)
    assert n2 == n
    assert np.array_equal(a, b)
    assert np.allclose(hist, hist2, rtol=1e-12, atol=0)


@pytest.mark.parametrize("plane", [False, True], ids=["march", "plane"])
def test_reinit_fast_mode_within_tolerance(S, oracle, plane):
    _mode(S, False, plane)
    shape = (40, 38, 36)
    p0 = synth_field(shape, seed=8)
    a, b = p0.copy(order="F"), p0.copy(order="F")
    st, n, hist = oracle.reinit(a, 31, DX, 0.0014)
    n2, hist2 = S.reinit(b, None, None, 39, 37, 35, 31, DX, 0.0014)
    assert n == n2 == 31
    assert np.abs(a - b).max() < 1e-13


def test_reinit_gradphi_outputs(S, oracle):
    _mode(S, True, False)
    shape = (20, 18, 19)
    p0 = synth_field(shape, seed=9)
    a, b = p0.copy(order="F"), p0.copy(order="F")
    st, n, hist, g, gm = oracle.reinit(a, 9, DX, 0.0014, want_grad=True)
    g2 = np.zeros(shape + (3,), order="F")
    gm2 = np.zeros(shape, order="F")
    S.reinit(b, g2, gm2, 19, 17, 18, 9, DX, 0.0014)
    assert np.array_equal(a, b) and np.array_equal(g, g2) and np.array_equal(gm, gm2)


@pytest.mark.parametrize("exact", [False, True], ids=["fast", "exact"])
def test_reinit_gradphi_outputs_after_a_tolerance_exit(S, oracle, exact):
    """The drop-in call passes gradPhi / gradPhiMag like the reference does.  On the march schedule they come from a
    replay of the last executed sweep on its snapshotted input; here the loop EXITs on tolerance at n = 18, in the
    middle of a batch of 8 enqueued sweeps."""
    _mode(S, exact, False)
    shape = (30, 28, 26)
    p0 = dist_field(shape, seed=3, noise=0.0005)
    a, b = p0.copy(order="F"), p0.copy(order="F")
    st, n, hist, g, gm = oracle.reinit(a, 60, DX, 0.000345, want_grad=True)
    assert st == 0 and n == 18
    g2 = np.zeros(shape + (3,), order="F")
    gm2 = np.zeros(shape, order="F")
    n2, hist2 = S.reinit(b, g2, gm2, 29, 27, 25, 60, DX, 0.000345)
    assert n2 == n
    if exact:
        assert np.array_equal(a, b) and np.array_equal(g, g2) and np.array_equal(gm, gm2)
    else:
        assert np.abs(a - b).max() < 1e-13 and np.abs(g - g2).max() < 1e-10 and np.abs(gm - gm2).max() < 1e-10


@pytest.mark.parametrize("exact", [False, True], ids=["fast", "exact"])
def test_cube40_reinit_full_parity(S, exact):
    """BASELINE config 1, reinit #1: 2155 sweeps to the reference's own exit."""
    _mode(S, exact, False)
    from levelsetfortran_b200 import stl
    X, E = load_mesh("cube40")
    g = stl.grid_from_surface(X, DX)
    gold = np.load(f"{GOLDEN}/cube40_fields.npz")
    phi = np.asfortranarray(gold["sign"].copy())
    n, hist = S.reinit(phi, None, None, 61, 61, 61, 10000, DX, 0.1 * g["dxx"])
    assert n == 2154 == int(gold["n_exit"][0])                        # same iteration count
    err = np.abs(phi - gold["reinit1"]).max()
    assert err <= (0.0 if exact else TOL), err
    assert np.allclose(hist, gold["rms_reinit1"], rtol=1e-9, atol=0)


def test_twocube10_nan_stop_parity(S):
    """BASELINE config 2: the reference STOPs with a NaN RMS at n = 272.  The NaN is
    0/0 in phiSign (subs.f90:169) at a cell whose frozen sign source is exactly -0.0, in the sweep where
    its Godunov gradient first evaluates to exactly 0 -- a bit-level event.  EXACT arithmetic reproduces it
    (and the field before it) bit for bit; the library default, AUTO, detects the ill-conditioned updates
    with its guard and finishes in EXACT, so the drop-in call gives the same answer."""
    from levelsetfortran_b200 import ReferenceStop, stl
    X, E = load_mesh("twoCube10")
    g = stl.grid_from_surface(X, DX)
    gold = np.load(f"{GOLDEN}/twoCube10_fields.npz")
    h = 0.1 * g["dxx"]
    for arith, plane in ((True, False), (True, True), (None, False)):
        _mode(S, arith, plane)
        phi = np.asfortranarray(gold["sign"].copy())
        with pytest.raises(ReferenceStop) as e:
            S.reinit(phi, None, None, 261, 41, 41, 10000, DX, h)
        assert S.last_arith() == "exact"
        assert e.value.n == 272 == int(gold["n_nan"][0])
        assert np.allclose(e.value.rms_hist[:272], gold["rms_reinit1"][:272], rtol=1e-9, atol=0)
        assert np.isnan(e.value.rms_hist[272])
        phi = np.asfortranarray(gold["sign"].copy())
        n, hist = S.reinit(phi, None, None, 261, 41, 41, 271, DX, h)
        assert n == 271 and np.array_equal(phi, gold["phi_n271"])
    # explicit FAST: no guarantee on this ill-conditioned input beyond "it also ends in the NaN STOP"
    _mode(S, False, False)
    phi = np.asfortranarray(gold["sign"].copy())
    with pytest.raises(ReferenceStop) as e:
        S.reinit(phi, None, None, 261, 41, 41, 10000, DX, h)
    assert e.value.n >= 200 and S.last_arith() == "fast"
    assert np.allclose(e.value.rms_hist[:100], gold["rms_reinit1"][:100], rtol=1e-6, atol=0)


def test_auto_mode_stays_fast_on_well_conditioned_input(S):
    """cube40 (BASELINE config 1) in the default AUTO mode: the guard does not fire, the call finishes in FAST
    arithmetic, within 1e-10 of the reference and with the reference's iteration count."""
    _mode(S, None, False)
    from levelsetfortran_b200 import stl
    X, E = load_mesh("cube40")
    g = stl.grid_from_surface(X, DX)
    gold = np.load(f"{GOLDEN}/cube40_fields.npz")
    phi = np.asfortranarray(gold["sign"].copy())
    n, hist = S.reinit(phi, None, None, 61, 61, 61, 10000, DX, 0.1 * g["dxx"])
    assert n == 2154
    assert np.abs(phi - gold["reinit1"]).max() <= TOL
    assert S.last_arith() == "fast"


# ------------------------------------------------------------------------------------ narrow band + min/max
def test_narrowband_exact(S, oracle):
    gold = np.load(f"{GOLDEN}/cube40_fields.npz")
    phi = np.asfortranarray(gold["reinit1"].copy())
    nb = np.zeros(phi.shape, dtype=np.int32, order="F")
    sb = np.zeros(phi.shape, dtype=np.int32, order="F")
    S.narrowBand(61, 61, 61, DX, phi, nb, sb)
    nbo, sbo = oracle.narrowband(phi, DX)
    assert np.array_equal(nb, nbo) and np.array_equal(sb, sbo) and nb.sum() == 84530


def test_cube40_minmax_full_parity_bit_exact(S):
    from levelsetfortran_b200 import stl
    X, E = load_mesh("cube40")
    g = stl.grid_from_surface(X, DX)
    gold = np.load(f"{GOLDEN}/cube40_fields.npz")
    phi = np.asfortranarray(gold["reinit1"].copy())
    phiN = phi.copy(order="F")
    nb = np.zeros(phi.shape, dtype=np.int32, order="F")
    sb = np.zeros(phi.shape, dtype=np.int32, order="F")
    S.narrowBand(61, 61, 61, DX, phi, nb, sb)
    n, hist = S.minMaxFlow(phi, phiN, nb, sb, 61, 61, 61, 10000, DX, 0.01 * g["dxx"])
    assert n == 406 == int(gold["n_exit"][1])
    assert np.array_equal(phi, gold["minmax"])
    assert np.allclose(hist, gold["rms_minmax"], rtol=1e-9, atol=0)
    assert np.array_equal(nb, gold["phiNB"].astype(np.int32)) and np.array_equal(sb, gold["phiSB"].astype(np.int32))


@pytest.mark.parametrize("algo", ["list", "march", "plane"])
@pytest.mark.parametrize("shape", [(30, 28, 26), (45, 20, 37), (19, 50, 33), (6, 5, 7)])
def test_minmax_small_vs_oracle_iteration_limit(S, oracle, shape, algo):
    _mode(S, False, *MM_ALGOS[algo])
    p0 = dist_field(shape, seed=11)
    a, b = p0.copy(order="F"), p0.copy(order="F")
    st, n, hist, nbo, sbo = oracle.minmax(a, 12, DX, 1.0e-4, tol=1e-30)
    assert st == 2 and n == 12
    phiN = b.copy(order="F")
    nb = np.zeros(shape, dtype=np.int32, order="F")
    sb = np.zeros(shape, dtype=np.int32, order="F")
    nx, ny, nz = (q - 1 for q in shape)
    S.narrowBand(nx, ny, nz, DX, b, nb, sb)
    n2, hist2 = S.minMaxFlow(b, phiN, nb, sb, nx, ny, nz, 12, DX, 1.0e-4, tol=1e-30)
    assert n2 == 12 and np.array_equal(a, b) and np.array_equal(phiN, b)
    assert np.array_equal(nb, nbo) and np.array_equal(sb, sbo)
    assert np.allclose(hist, hist2, rtol=1e-12, atol=0)


def test_minmax_given_mask_first_iteration(S, oracle):
    """Iteration 1 uses the caller's phiNB (set3d.f90:360), later ones narrowBand of the iterate (:460)."""
    import ctypes
    shape = (28, 26, 24)
    p0 = dist_field(shape, seed=14)
    nx, ny, nz = (q - 1 for q in shape)
    rng = np.random.default_rng(1)
    nb0 = np.zeros(shape, dtype=np.int32, order="F")
    nb0[2:-2, 2:-2, 2:-2] = rng.random((shape[0] - 4, shape[1] - 4, shape[2] - 4)) < 0.3
    dp = ctypes.POINTER(ctypes.c_double)
    for algo in MM_ALGOS.values():
        _mode(S, False, *algo)
        a, an, nba, sba = p0.copy(order="F"), p0.copy(order="F"), nb0.copy(order="F"), np.zeros(shape, dtype=np.int32, order="F")
        hist = np.zeros(3)
        ne = ctypes.c_int(0)
        st = oracle.lib().orc_minmax(a.ctypes.data_as(dp), an.ctypes.data_as(dp), nba.ctypes.data_as(oracle.c_i32_p),
                                     sba.ctypes.data_as(oracle.c_i32_p), nx, ny, nz, 3, DX, 1.0e-4, 1e-30,
                                     ctypes.byref(ne), hist.ctypes.data_as(dp))
        assert st == 2
        b, bn, nbb, sbb = p0.copy(order="F"), p0.copy(order="F"), nb0.copy(order="F"), np.zeros(shape, dtype=np.int32, order="F")
        n2, hist2 = S.minMaxFlow(b, bn, nbb, sbb, nx, ny, nz, 3, DX, 1.0e-4, tol=1e-30)
        assert n2 == 3 and np.array_equal(a, b) and np.array_equal(an, bn)
        assert np.array_equal(nba, nbb) and np.array_equal(sba, sbb)
        assert np.allclose(hist, hist2, rtol=1e-12, atol=0)


def test_minmax_band_on_boundary_is_an_error(S):
    from levelsetfortran_b200._lib import LsfError, LSF_ERR_BAND_ON_BOUNDARY
    shape = (12, 12, 12)
    phi = np.full(shape, 0.01, order="F")              # every cell, including the boundary, is in the band
    nb = np.ones(shape, dtype=np.int32, order="F")
    sb = np.ones(shape, dtype=np.int32, order="F")
    for algo in MM_ALGOS.values():
        _mode(S, False, *algo)
        with pytest.raises(LsfError) as e:
            S.minMaxFlow(phi.copy(order="F"), phi.copy(order="F"), nb, sb, 11, 11, 11, 2, DX, 1.0e-4)
        assert e.value.code == LSF_ERR_BAND_ON_BOUNDARY
    _mode(S, False, False)


@pytest.mark.parametrize("shape", [(128, 128, 128), (96, 160, 200)])
def test_minmax_march_equals_plane_schedule_at_size(S, shape):
    """Size-independent property: the fused march kernel and the per-hyperplane cross-check path are both
    exact re-orderings of the reference loop, so they agree bit for bit (phi, phiN, masks, RMS history)."""
    p0 = dist_field(shape, seed=15)
    nx, ny, nz = (q - 1 for q in shape)
    out = []
    for algo in ("plane", "march", "list"):
        _mode(S, False, *MM_ALGOS[algo])
        b, bn = p0.copy(order="F"), p0.copy(order="F")
        nb = np.zeros(shape, dtype=np.int32, order="F")
        sb = np.zeros(shape, dtype=np.int32, order="F")
        S.narrowBand(nx, ny, nz, DX, b, nb, sb)
        n, hist = S.minMaxFlow(b, bn, nb, sb, nx, ny, nz, 5, DX, 1.0e-4, tol=1e-30)
        out.append((n, b, bn, nb, sb, hist))
    assert out[0][0] == out[1][0] == out[2][0] == 5
    for other in (1, 2):
        for q in range(1, 5):
            assert np.array_equal(out[0][q], out[other][q])
        assert np.allclose(out[0][5], out[other][5], rtol=1e-11, atol=0)
    assert not np.array_equal(out[0][1], p0)
    _mode(S, False, False)


@pytest.mark.parametrize("h1,scale", [(2.0e-3, 1.0), (1.0e-3, 1.0e-2)])
def test_minmax_active_list_settle_path_on_gpu(S, oracle, h1, scale):
    """Large h1 / small amplitude: pAve hovers around zero, the 8-combination check and the settle queue of the
    active-list kernel really run (CPU twin: tests/test_march_emu.py); bit-exact against the oracle."""
    _mode(S, False, False)
    shape = (40, 36, 34)
    p0 = np.asfortranarray(dist_field(shape, seed=9) * scale)
    if scale != 1.0:
        edge = np.ones(shape, dtype=bool)
        edge[1:-1, 1:-1, 1:-1] = False
        p0[edge] = 1.0
    a, b = p0.copy(order="F"), p0.copy(order="F")
    st, n, hist, nbo, sbo = oracle.minmax(a, 12, DX, h1, tol=1e-300)
    assert n == 12
    nx, ny, nz = (q - 1 for q in shape)
    phiN = b.copy(order="F")
    nb = np.zeros(shape, dtype=np.int32, order="F")
    sb = np.zeros(shape, dtype=np.int32, order="F")
    S.narrowBand(nx, ny, nz, DX, b, nb, sb)
    n2, hist2 = S.minMaxFlow(b, phiN, nb, sb, nx, ny, nz, 12, DX, h1, tol=1e-300)
    assert n2 == 12 and np.array_equal(a, b)
    assert np.allclose(hist, hist2, rtol=1e-11, atol=0)


# ------------------------------------------------------------------------------------ device-resident pipeline + larger sizes
def test_device_pipeline_cube40_default_run(S):
    """sign search -> reinit -> min/max on the device without host round trips == golden."""
    _mode(S, False, False)
    from levelsetfortran_b200 import DeviceGrid, stl
    X, E = load_mesh("cube40")
    g = stl.grid_from_surface(X, DX)
    gold = np.load(f"{GOLDEN}/cube40_fields.npz")
    G = DeviceGrid(g["nx"], g["ny"], g["nz"])
    G.fill(1.0)
    G.signSearch(g["xLo"], DX, X, E, g["box"])
    assert np.array_equal(G.download(), gold["sign"])
    rc, n, hist = G.reinit(10000, DX, 0.1 * g["dxx"])
    assert (rc, n) == (0, 2154)
    assert np.abs(G.download() - gold["reinit1"]).max() <= TOL
    rc, n, hist = G.minMaxFlow(10000, DX, 0.01 * g["dxx"])
    assert (rc, n) == (0, 406)
    assert np.abs(G.download() - gold["minmax"]).max() <= TOL
    rc, n, hist = G.reinit(2000, DX, 0.001 * g["dxx"])               # reinit #2, set3d.f90:576-582
    assert (rc, n) == (0, 0)
    G.close()


@pytest.mark.parametrize("shape", [(128, 128, 128), (96, 160, 200)])
def test_march_equals_plane_schedule_at_size(S, shape):
    """Size-independent property: both schedules are exact re-orderings, so in exact arithmetic
    they agree bit for bit at sizes where the CPU oracle is too slow to run."""
    p0 = synth_field(shape, seed=12)
    nx, ny, nz = (s - 1 for s in shape)
    a, b = p0.copy(order="F"), p0.copy(order="F")
    _mode(S, True, True)
    S.reinit(a, None, None, nx, ny, nz, 7, DX, 0.0014)
    _mode(S, True, False)
    S.reinit(b, None, None, nx, ny, nz, 7, DX, 0.0014)
    assert np.array_equal(a, b)


def test_one_sweep_vs_oracle_128(S, oracle):
    """One full raster-1 sweep at 128^3 against the oracle (~0.4 s of CPU)."""
    shape = (128, 128, 128)
    p0 = synth_field(shape, seed=13)
    a, b = p0.copy(order="F"), p0.copy(order="F")
    oracle.reinit(a, 0, DX, 0.0014)
    _mode(S, True, False)
    S.reinit(b, None, None, 127, 127, 127, 0, DX, 0.0014)
    assert np.array_equal(a, b)
    c = p0.copy(order="F")
    _mode(S, False, False)
    S.reinit(c, None, None, 127, 127, 127, 0, DX, 0.0014)
    assert np.abs(a - c).max() < 1e-14


def test_grid_checksum_matches_numpy_and_is_partition_independent(S):
    """lsf_grid_checksum: sum of bits(phi(q))*(2q+1) mod 2^64 and xor of the bit patterns over the grid's points -- the digest
    bench.py prints for N = 1, 2, 4, 8 to show that the sharded fields are bit-identical."""
    shape = (23, 17, 19)
    phi = synth_field(shape, seed=31)
    G = S.DeviceGrid(shape[0] - 1, shape[1] - 1, shape[2] - 1)
    G.upload(phi)
    s, x = G.checksum()
    G.close()
    bits = phi.reshape(-1, order="F").view(np.uint64)
    q = np.arange(bits.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        want_s = int((bits * (np.uint64(2) * q + np.uint64(1))).sum(dtype=np.uint64))
    want_x = int(np.bitwise_xor.reduce(bits))
    assert (s, x) == (want_s, want_x)
    # the slabs of any cut along k add / xor up to the same digest (what the ranks of a sharded grid compute)
    sxy = shape[0] * shape[1]
    parts = [(0, 7), (7, 12), (12, 19)]
    with np.errstate(over="ignore"):
        tot = sum(int((bits[a * sxy:b * sxy] * (np.uint64(2) * q[a * sxy:b * sxy] + np.uint64(1))).sum(dtype=np.uint64)) for a, b in parts)
    assert tot % (1 << 64) == want_s


def test_host_register_roundtrip(S, lsf):
    """lsf_host_register / lsf_host_unregister: page-locking a caller-owned array does not change results"""
    import ctypes as C
    from levelsetfortran_b200 import _lib
    shape = (20, 18, 16)
    a = synth_field(shape, seed=8)
    b = a.copy(order="F")
    _lib.check(_lib.lib().lsf_host_register(a.ctypes.data, a.nbytes))
    try:
        S.set_arith(True)
        S.reinit(a, None, None, 19, 17, 15, 5, DX, 0.0014)
    finally:
        _lib.check(_lib.lib().lsf_host_unregister(a.ctypes.data))
    S.reinit(b, None, None, 19, 17, 15, 5, DX, 0.0014)
    S.set_arith(None)
    assert np.array_equal(a, b)
