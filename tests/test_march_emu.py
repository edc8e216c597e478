"""The production sweep schedule (levelsetfortran_b200/csrc/lsf_march.cuh: skewed x-marching column
tiles with ticket/progress-flag synchronisation) compiled for the CPU -- one OS thread per CUDA
thread, several CTAs running concurrently -- and checked against the oracle's literal
lexicographic Gauss-Seidel sweep.  Exact arithmetic must be bit-identical for all 8 rasters."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, synth_field

EMU_DIR = os.path.join(ROOT, "tests", "emu")
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU_DIR, "libmarch_emu.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-pthread",
                           "-std=c++17", "-Wno-unknown-pragmas", "-o", so, os.path.join(EMU_DIR, "march_emu.cpp")])
    L = C.CDLL(so)
    L.emu_march_sweep.restype = C.c_double
    L.emu_march_sweep.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
    return L


@pytest.mark.parametrize("shape,ncta", [((22, 21, 23), 1), ((24, 36, 22), 4), ((40, 38, 36), 9), ((12, 50, 20), 6)])
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
