"""The production sweep schedule (levelsetfortran_b200/csrc/lsf_march.cuh: skewed x-marching column
tiles with ticket/progress-flag synchronisation) compiled for the CPU -- one OS thread per CUDA
thread, several CTAs running concurrently -- and checked against the oracle's literal
lexicographic Gauss-Seidel sweep.  Exact arithmetic must be bit-identical for all 8 rasters."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, dist_field, synth_field

EMU_DIR = os.path.join(ROOT, "tests", "emu")
dp = C.POINTER(C.c_double)


def _rows_default():
    """The R (rows per thread) the GPU library is built with: LSF_ROWS in csrc/Makefile."""
    import re
    mk = open(os.path.join(ROOT, "levelsetfortran_b200", "csrc", "Makefile")).read()
    m = re.search(r"-DLSF_ROWS=(\d+)", mk)
    return int(m.group(1)) if m else 1


@pytest.fixture(scope="module", params=[_rows_default()], ids=lambda r: f"rows{r}")
def emu(request):
    so = os.path.join(EMU_DIR, f"libmarch_emu_r{request.param}.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-pthread",
                           "-std=c++17", "-Wno-unknown-pragmas", f"-DLSF_ROWS={request.param}", "-o", so,
                           os.path.join(EMU_DIR, "march_emu.cpp")])
    L = C.CDLL(so)
    L.emu_set_dynamic(1)          # the production scheduler: tiles picked dynamically (march_pick), not from a static ticket order
    L.emu_march_sweep.restype = C.c_double
    L.emu_march_sweep.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
    L.emu_march_sweep_slabs.restype = C.c_double
    L.emu_march_sweep_slabs.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                        C.c_int, C.c_int, C.c_int]
    L.emu_march_multi_slabs.restype = C.c_double
    L.emu_march_multi_slabs.argtypes = [dp, dp] + [C.c_int] * 6 + [C.c_double, C.c_double] + [C.c_int] * 4
    L.emu_mm_iteration_slabs.restype = C.c_double
    L.emu_mm_iteration_slabs.argtypes = [dp, dp] + [C.c_int] * 4 + [C.c_double, C.c_double, C.c_int, C.c_int]
    L.emu_mm_list_iteration.restype = C.c_double
    L.emu_mm_list_iteration.argtypes = [dp, dp, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_longlong)]
    L.emu_mm_iteration.restype = C.c_double
    L.emu_mm_iteration.argtypes = [dp, dp, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
    return L


# the last two shapes have tiles that lie inside the high-order window with all halo rows present, and enough steps along x
# for the steady-state copy of the step body (lsf_march.cuh, ST = true) to run: (60,40,39) on one tile, (47,56,55) on four
@pytest.mark.parametrize("shape,ncta", [((22, 21, 23), 1), ((24, 36, 22), 4), ((40, 38, 36), 9), ((12, 50, 20), 6),
                                        ((60, 40, 39), 4), ((47, 56, 55), 5)])
def test_march_schedule_is_an_exact_reordering(emu, oracle, shape, ncta):
    p0 = synth_field(shape, seed=1)
    pS = p0.copy(order="F")
    a, b, f = p0.copy(order="F"), p0.copy(order="F"), p0.copy(order="F")
    nx, ny, nz = (s - 1 for s in shape)
    for r in range(1, 9):
        before = a.copy(order="F")
        oracle.reinit_sweep(a, pS, 0.05, 0.0014, r)
        s = emu.emu_march_sweep(b.ctypes.data_as(dp), pS.ctypes.data_as(dp), nx, ny, nz, r, 0.05, 0.0014, 1, ncta)
        assert np.array_equal(a, b), f"raster {r}: exact arithmetic differs from the oracle"
        # fused RMS partial == sum over interior of (new-old)^2
        ref = float(((a - before)[1:-1, 1:-1, 1:-1] ** 2).sum())
        assert abs(s - ref) <= 1e-12 * max(ref, 1e-300)
        emu.emu_march_sweep(f.ctypes.data_as(dp), pS.ctypes.data_as(dp), nx, ny, nz, r, 0.05, 0.0014, 0, ncta)
        assert np.abs(a - f).max() < 1e-13, f"raster {r}: fast arithmetic drifted"


@pytest.mark.parametrize("shape,nranks,ncta,m", [((22, 21, 40), 2, 2, 1), ((20, 36, 51), 3, 3, 2), ((14, 70, 33), 4, 2, 3),
                                                 ((12, 20, 64), 8, 1, 1), ((50, 40, 80), 2, 3, 2)])
def test_march_slab_pipeline_is_an_exact_reordering(emu, oracle, shape, nranks, ncta, m):
    """z-slab sharding (lsf_slab.cuh): `nranks` slabs sweep concurrently, each on its own local array with
    ghost planes, coupled only through the streaming halo (peer stores + in_progress flags).  The gathered
    result must be bit-identical to the oracle's serial sweep of the whole grid, for all 8 rasters (both
    pipeline directions) and for tilted ticket orders."""
    p0 = synth_field(shape, seed=2)
    pS = p0.copy(order="F")
    a, b = p0.copy(order="F"), p0.copy(order="F")
    nx, ny, nz = (s - 1 for s in shape)
    for r in range(1, 9):
        before = a.copy(order="F")
        oracle.reinit_sweep(a, pS, 0.05, 0.0014, r)
        s = emu.emu_march_sweep_slabs(b.ctypes.data_as(dp), pS.ctypes.data_as(dp), nx, ny, nz, nranks, r, 0.05, 0.0014,
                                      1, ncta, m)
        assert s >= 0, f"emulator failed ({s})"
        assert np.array_equal(a, b), f"raster {r}: sharded sweep differs from the serial oracle"
        ref = float(((a - before)[1:-1, 1:-1, 1:-1] ** 2).sum())
        assert abs(s - ref) <= 1e-12 * max(ref, 1e-300)


@pytest.mark.parametrize("shape,nranks,ncta,m,skew", [((22, 21, 40), 2, 2, 1, 0), ((20, 41, 51), 3, 2, 2, 3000),
                                                      ((14, 70, 33), 4, 2, 8, 1000), ((12, 20, 64), 8, 1, 1, 2000),
                                                      ((20, 36, 37), 2, 3, 1, 3000), ((16, 25, 38), 2, 2, 2, 2000)])
def test_march_slab_pipeline_free_running_sweeps(emu, oracle, shape, nranks, ncta, m, skew):
    """Eleven consecutive sweeps (all 8 rasters, both k flips, a b-orientation change between neighbouring
    sweeps on a grid whose tile columns do not align) with the ranks running freely: no exchange or barrier
    between sweeps, only the per-tile flags.  Must equal the oracle's serial sweeps bit for bit."""
    p0 = synth_field(shape, seed=5)
    pS = p0.copy(order="F")
    a, b = p0.copy(order="F"), p0.copy(order="F")
    nx, ny, nz = (s - 1 for s in shape)
    nsweeps, first = 11, 7
    ref = 0.0
    for n in range(nsweeps):
        before = a.copy(order="F")
        oracle.reinit_sweep(a, pS, 0.05, 0.0014, (first - 1 + n) % 8 + 1)
        ref += float(((a - before)[1:-1, 1:-1, 1:-1] ** 2).sum())
    emu.emu_set_stall(20000 if skew else 0)      # every tile stalls 20 ms shortly before its end: ranks run ahead where allowed
    try:
        s = emu.emu_march_multi_slabs(b.ctypes.data_as(dp), pS.ctypes.data_as(dp), nx, ny, nz, nranks, nsweeps, first,
                                      0.05, 0.0014, 1, ncta, m, skew)
    finally:
        emu.emu_set_stall(0)
    assert s >= 0, f"emulator failed ({s})"
    assert np.array_equal(a, b), "free-running sharded sweeps differ from the serial oracle"
    assert abs(s - ref) <= 1e-12 * ref


@pytest.mark.parametrize("shape,ncta", [((22, 21, 23), 1), ((24, 36, 22), 4), ((40, 38, 36), 9), ((12, 50, 20), 6)])
def test_minmax_march_iteration_is_bit_exact(emu, oracle, shape, ncta):
    """The fused min/max iteration kernel (lsf_mm_march.cuh: out-of-place march, old/new rings) against
    the oracle's literal two-pass loop (set3d.f90:399-431), several iterations, ping-pong buffers."""
    p0 = dist_field(shape, seed=3)
    nx, ny, nz = (s - 1 for s in shape)
    a = p0.copy(order="F")
    st, n, hist, nbo, sbo = oracle.minmax(a, 6, 0.05, 1.0e-4, tol=1e-30)
    assert st == 2 and n == 6
    buf = [p0.copy(order="F"), p0.copy(order="F")]
    sums = []
    for it in range(1, 7):
        A, B = buf[(it - 1) & 1], buf[it & 1]
        s = emu.emu_mm_iteration(A.ctypes.data_as(dp), B.ctypes.data_as(dp), None, nx, ny, nz, 0.05, 1.0e-4, ncta)
        sums.append(np.sqrt(s / (nx * ny * nz)))
    assert np.array_equal(buf[0], a)                     # 6 iterations: result back in buffer 0
    assert np.allclose(sums, hist, rtol=1e-12, atol=0)
    assert (np.abs(p0) < 4.1 * 0.05).sum() > 100         # the band was not empty


@pytest.mark.parametrize("shape,nranks,ncta,m", [((22, 21, 40), 2, 2, 1), ((20, 36, 51), 3, 2, 2), ((12, 20, 64), 8, 1, 8)])
def test_minmax_march_slabs_bit_exact(emu, oracle, shape, nranks, ncta, m):
    """The fused min/max iteration on z-slabs (streaming NEW-value halo into the downstream rank's ghost
    plane) against the oracle's literal loop on the whole grid: bit-exact, several iterations."""
    p0 = dist_field(shape, seed=6)
    nx, ny, nz = (s - 1 for s in shape)
    a = p0.copy(order="F")
    st, n, hist, nbo, sbo = oracle.minmax(a, 4, 0.05, 1.0e-4, tol=1e-30)
    assert n == 4
    buf = [p0.copy(order="F"), p0.copy(order="F")]
    sums = []
    for it in range(1, 5):
        A, B = buf[(it - 1) & 1], buf[it & 1]
        s = emu.emu_mm_iteration_slabs(A.ctypes.data_as(dp), B.ctypes.data_as(dp), nx, ny, nz, nranks, 0.05, 1.0e-4, ncta, m)
        assert s >= 0
        sums.append(np.sqrt(s / (nx * ny * nz)))
    assert np.array_equal(buf[0], a)
    assert np.allclose(sums, hist, rtol=1e-12, atol=0)


@pytest.mark.parametrize("shape,h1,scale", [((24, 22, 23), 1.0e-4, 1.0), ((30, 26, 28), 2.0e-3, 1.0), ((26, 30, 24), 1.0e-3, 1.0e-2)])
def test_minmax_active_list_speculation_is_bit_exact(emu, oracle, shape, h1, scale):
    """The active-list formulation of the min/max iteration (lsf_mm_list.cuh): every band cell is decided
    from OLD values alone when all outcomes of its three upstream neighbours give the same sign of pAve,
    the rest are settled afterwards.  Against the oracle's literal in-place loop, 12 iterations, bit-exact;
    the large-h1 / small-amplitude cases force the 8-combination check and the settle path to run."""
    p0 = np.asfortranarray(dist_field(shape, seed=9) * scale) if scale != 1.0 else dist_field(shape, seed=9)
    if scale != 1.0:      # small amplitude: everything is in the band and pAve hovers around zero
        edge = np.ones(shape, dtype=bool); edge[1:-1, 1:-1, 1:-1] = False
        p0[edge] = 1.0
    nx, ny, nz = (s - 1 for s in shape)
    a = p0.copy(order="F")
    st, n, hist, nbo, sbo = oracle.minmax(a, 12, 0.05, h1, tol=1e-300)
    assert n == 12
    buf = [p0.copy(order="F"), p0.copy(order="F")]
    tot = np.zeros(3, dtype=np.int64)
    sums = []
    for it in range(1, 13):
        A, B = buf[(it - 1) & 1], buf[it & 1]
        stats = (C.c_longlong * 3)()
        s = emu.emu_mm_list_iteration(A.ctypes.data_as(dp), B.ctypes.data_as(dp), None, nx, ny, nz, 0.05, h1, stats)
        tot += np.array(list(stats))
        sums.append(np.sqrt(s / (nx * ny * nz)))
    assert np.array_equal(buf[0], a)
    assert np.allclose(sums, hist, rtol=1e-11, atol=0)
    assert tot[0] > 1000
    if h1 >= 2.0e-3:
        assert tot[1] > 0, "the exact combination check never ran"
    if scale != 1.0:
        assert tot[2] > 0, "no cell ever went through the settle path"


def test_minmax_march_given_mask(emu, oracle):
    """Iteration 1 honours a caller-provided band mask (phiNB, set3d.f90:360) instead of the abs test."""
    shape = (24, 22, 20)
    p0 = dist_field(shape, seed=4)
    nx, ny, nz = (s - 1 for s in shape)
    rng = np.random.default_rng(0)
    nb = np.zeros(shape, dtype=np.int32, order="F")
    nb[2:-2, 2:-2, 2:-2] = (rng.random((shape[0] - 4, shape[1] - 4, shape[2] - 4)) < 0.3)
    a = p0.copy(order="F")
    phiN = a.copy(order="F")
    sb = np.zeros(shape, dtype=np.int32, order="F")
    import ctypes
    from oracle import oracle as O
    hist = np.zeros(1)
    ne = ctypes.c_int(0)
    O.lib().orc_minmax(a.ctypes.data_as(dp), phiN.ctypes.data_as(dp), nb.copy(order="F").ctypes.data_as(O.c_i32_p),
                       sb.ctypes.data_as(O.c_i32_p), nx, ny, nz, 1, 0.05, 1.0e-4, 1e-30, ctypes.byref(ne),
                       hist.ctypes.data_as(dp))
    A, B = p0.copy(order="F"), p0.copy(order="F")
    m8 = np.asfortranarray(nb.astype(np.uint8))
    emu.emu_mm_iteration(A.ctypes.data_as(dp), B.ctypes.data_as(dp), m8.ctypes.data, nx, ny, nz, 0.05, 1.0e-4, 3)
    assert np.array_equal(B, a)


# ------------------------------------------------------------------------------------ fp32 mode
@pytest.mark.parametrize("shape,ncta", [((22, 21, 23), 1), ((24, 36, 22), 4), ((40, 38, 36), 6)])
def test_f32_march_schedule_is_an_exact_reordering(emu, oracle, shape, ncta):
    """The fp32 instantiation (F32Arith, float slot ring) of the production schedule against the literal in-place
    raster loop over the same cell arithmetic: bit-identical for all 8 rasters; and within the fp32 contract
    (1e-4 relative) of the fp64 oracle."""
    fp = C.POINTER(C.c_float)
    emu.emu_march_sweep_f32.restype = C.c_double
    emu.emu_march_sweep_f32.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
    emu.emu_lex_sweep_f32.restype = None
    emu.emu_lex_sweep_f32.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
    p0 = synth_field(shape, seed=1)
    d = p0.copy(order="F")
    pSd = p0.copy(order="F")
    a = p0.astype(np.float32, order="F")
    b = a.copy(order="F")
    pS = a.copy(order="F")
    nx, ny, nz = (s - 1 for s in shape)
    for r in range(1, 9):
        before = a.astype(np.float64)
        emu.emu_lex_sweep_f32(a.ctypes.data_as(fp), pS.ctypes.data_as(fp), nx, ny, nz, r, 0.05, 0.0014, 0)
        s = emu.emu_march_sweep_f32(b.ctypes.data_as(fp), pS.ctypes.data_as(fp), nx, ny, nz, r, 0.05, 0.0014, ncta)
        assert np.array_equal(a, b), f"raster {r}: fp32 march differs from the fp32 lexicographic loop"
        ref = float(((a.astype(np.float64) - before)[1:-1, 1:-1, 1:-1] ** 2).sum())
        assert abs(s - ref) <= 1e-4 * max(ref, 1e-300)
        oracle.reinit_sweep(d, pSd, 0.05, 0.0014, r)
        assert np.abs(a - d).max() <= 1e-4 * np.abs(d).max()


@pytest.mark.parametrize("shape,nranks,ncta,m", [((22, 21, 40), 2, 2, 1), ((20, 36, 51), 3, 3, 2)])
def test_f32_slab_pipeline_is_an_exact_reordering(emu, shape, nranks, ncta, m):
    """z-slab sharding of the fp32 mode: the slabs' concurrent sweeps (float peer stores + in_progress flags) gathered
    back must be bit-identical to the fp32 lexicographic sweep of the whole grid, all 8 rasters."""
    fp = C.POINTER(C.c_float)
    emu.emu_lex_sweep_f32.restype = None
    emu.emu_lex_sweep_f32.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
    emu.emu_march_sweep_slabs_f32.restype = C.c_double
    emu.emu_march_sweep_slabs_f32.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
    p0 = synth_field(shape, seed=2).astype(np.float32, order="F")
    pS = p0.copy(order="F")
    a, b = p0.copy(order="F"), p0.copy(order="F")
    nx, ny, nz = (s - 1 for s in shape)
    for r in range(1, 9):
        emu.emu_lex_sweep_f32(a.ctypes.data_as(fp), pS.ctypes.data_as(fp), nx, ny, nz, r, 0.05, 0.0014, 0)
        s = emu.emu_march_sweep_slabs_f32(b.ctypes.data_as(fp), pS.ctypes.data_as(fp), nx, ny, nz, nranks, r, 0.05, 0.0014, ncta, m)
        assert s >= 0, s
        assert np.array_equal(a, b), f"raster {r}: fp32 slab pipeline differs from the fp32 lexicographic sweep"


# ------------------------------------------------------------------------------------ overlapped sweeps
@pytest.mark.parametrize("shape,ncta,nsweeps,first,stall", [((22, 21, 23), 1, 3, 1, 0), ((40, 38, 36), 9, 10, 1, 20000),
                                                            ((19, 50, 33), 6, 11, 4, 5000), ((6, 5, 7), 2, 5, 7, 0)])
def test_overlapped_sweeps_are_an_exact_reordering(emu, oracle, shape, ncta, nsweeps, first, stall):
    """march_multi_cta: several consecutive sweeps in ONE pass over a shared ticket counter -- CTAs start the next
    sweep's tiles while the previous sweep drains, the boundary block is folded into the tiles and the boundary values
    alternate between phi and a shell array.  phi must equal the oracle's sweep + boundary block sequence bit for bit
    and the per-sweep RMS sums (interior + boundary parts) must match."""
    emu.emu_march_overlapped.restype = C.c_int
    emu.emu_march_overlapped.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, dp]
    p0 = synth_field(shape, seed=3)
    pS = p0.copy(order="F")
    a, b, f = p0.copy(order="F"), p0.copy(order="F"), p0.copy(order="F")
    nx, ny, nz = (s - 1 for s in shape)
    ref = np.zeros(nsweeps)
    for s in range(nsweeps):
        before = a.copy(order="F")
        oracle.reinit_sweep(a, pS, 0.05, 0.0014, (first - 1 + s) % 8 + 1)
        oracle.bc(a, 0.05)
        ref[s] = float(((a - before) ** 2).sum())
    rms = np.zeros(nsweeps)
    emu.emu_set_stall(stall)
    try:
        rc = emu.emu_march_overlapped(b.ctypes.data_as(dp), pS.ctypes.data_as(dp), nx, ny, nz, nsweeps, first, 0.05, 0.0014, 1, ncta,
                                      rms.ctypes.data_as(dp))
    finally:
        emu.emu_set_stall(0)
    assert rc == 0
    assert np.array_equal(a, b), "overlapped sweeps differ from the serial oracle"
    assert np.allclose(rms, ref, rtol=1e-12, atol=0)
    rc = emu.emu_march_overlapped(f.ctypes.data_as(dp), pS.ctypes.data_as(dp), nx, ny, nz, nsweeps, first, 0.05, 0.0014, 0, ncta, None)
    assert rc == 0 and np.abs(a - f).max() < 1e-13


@pytest.mark.parametrize("shape,ncta,nsweeps,first,stall", [((24, 36, 22), 2, 4, 1, 0), ((40, 38, 36), 3, 6, 3, 10000),
                                                            ((19, 50, 33), 2, 5, 6, 5000)])
def test_overlapped_sweeps_one_kernel_per_sweep(emu, oracle, shape, ncta, nsweeps, first, stall):
    """The second packaging (one kernel per sweep, chained by programmatic dependent launch on the GPU): here every sweep
    has its own CTAs and ticket counter and ALL sweeps run at once -- more overlap than the GPU ever admits.  Bit-identical."""
    emu.emu_march_overlapped_per_sweep.restype = C.c_int
    emu.emu_march_overlapped_per_sweep.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                                   C.c_int, C.c_int, dp]
    p0 = synth_field(shape, seed=3)
    pS = p0.copy(order="F")
    a, b = p0.copy(order="F"), p0.copy(order="F")
    nx, ny, nz = (s - 1 for s in shape)
    ref = np.zeros(nsweeps)
    for s in range(nsweeps):
        before = a.copy(order="F")
        oracle.reinit_sweep(a, pS, 0.05, 0.0014, (first - 1 + s) % 8 + 1)
        oracle.bc(a, 0.05)
        ref[s] = float(((a - before) ** 2).sum())
    rms = np.zeros(nsweeps)
    emu.emu_set_stall(stall)
    try:
        rc = emu.emu_march_overlapped_per_sweep(b.ctypes.data_as(dp), pS.ctypes.data_as(dp), nx, ny, nz, nsweeps, first, 0.05,
                                                0.0014, 1, ncta, rms.ctypes.data_as(dp))
    finally:
        emu.emu_set_stall(0)
    assert rc == 0 and np.array_equal(a, b)
    assert np.allclose(rms, ref, rtol=1e-12, atol=0)
