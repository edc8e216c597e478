"""ctypes front-end for the CPU oracle (oracle/lsf_oracle.c).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by the product
package `levelsetfortran_b200`.  Pinned bit for bit to the machine-translated reference (oracle/ref.py,
see lsf_oracle.c header).

All grid arrays are numpy float64/int32 in Fortran order with shape
(nx+1, ny+1, nz+1), i.e. the reference's phi(0:nx,0:ny,0:nz).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblsf_oracle.so")
_lib = None

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_i32_p = C.POINTER(C.c_int32)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lsf_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liblsf_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_phisign.restype = C.c_double
        L.orc_phisign.argtypes = [C.c_double] * 3
        L.orc_weno.restype = C.c_double
        L.orc_weno.argtypes = [C.c_int] * 6 + [C.c_double, c_double_p, c_double_p, c_double_p]
        L.orc_narrowband.restype = None
        L.orc_narrowband.argtypes = [C.c_int] * 3 + [C.c_double, c_double_p, c_i32_p, c_i32_p]
        for name in ("orc_bc_literal", "orc_bc_closed"):
            f = getattr(L, name)
            f.restype = None
            f.argtypes = [c_double_p] + [C.c_int] * 3 + [C.c_double]
        for name in ("orc_reinit_sweep", "orc_reinit_sweep_hyperplane"):
            f = getattr(L, name)
            f.restype = None
            f.argtypes = [c_double_p, c_double_p] + [C.c_int] * 3 + [C.c_double, C.c_double, C.c_int, c_double_p, c_double_p]
        L.orc_rms.restype = C.c_double
        L.orc_rms.argtypes = [c_double_p, c_double_p] + [C.c_int] * 3
        L.orc_reinit.restype = C.c_int
        L.orc_reinit.argtypes = ([c_double_p] * 3 + [C.c_int] * 4 + [C.c_double] * 3 + [C.c_int] * 2
                                 + [c_int_p, c_double_p])
        L.orc_secondderiv.restype = None
        L.orc_secondderiv.argtypes = [C.c_int] * 6 + [C.c_double, c_double_p, c_double_p]
        L.orc_minmax.restype = C.c_int
        L.orc_minmax.argtypes = ([c_double_p, c_double_p, c_i32_p, c_i32_p] + [C.c_int] * 4
                                 + [C.c_double] * 3 + [c_int_p, c_double_p])
        L.orc_grid_from_surface.restype = None
        L.orc_grid_from_surface.argtypes = [c_double_p, C.c_int, C.c_double, C.c_int, c_int_p, c_double_p, c_int_p, c_double_p]
        L.orc_sign_init.restype = None
        L.orc_sign_init.argtypes = ([c_double_p] + [C.c_int] * 3 + [c_double_p, C.c_double, c_double_p, C.c_int,
                                    c_i32_p, C.c_int] + [C.c_int] * 6)
        L.orc_stl_ntri.restype = C.c_int
        L.orc_stl_ntri.argtypes = [C.c_char_p]
        L.orc_stl_read.restype = C.c_int
        L.orc_stl_read.argtypes = [C.c_char_p, c_double_p, c_i32_p, c_int_p]
        L.orc_gradphi_band8.restype = C.c_int
        L.orc_gradphi_band8.argtypes = [C.c_int] * 3 + [C.c_double, c_double_p, c_i32_p, c_double_p]
        L.orc_advect_nodes.restype = C.c_int
        L.orc_advect_nodes.argtypes = ([c_double_p] + [C.c_int] * 3 + [C.c_double, c_double_p, c_double_p, C.c_int]
                                       + [c_double_p] * 3 + [C.c_int, C.c_int, C.POINTER(C.c_longlong)])
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_i32_p) if a is not None else None


def _chk(phi):
    assert phi.dtype == np.float64 and phi.flags.f_contiguous and phi.ndim == 3, "phi must be F-ordered float64 3-D"
    return phi.shape[0] - 1, phi.shape[1] - 1, phi.shape[2] - 1


# --------------------------------------------------------------------------- host-side helpers
def stl_read(path: str):
    """subs.f90:17-121.  Returns surfX (nSurfNode,3) F-order fp64, surfElem (ntri,3) F-order int32 1-based."""
    L = lib()
    ntri = L.orc_stl_ntri(path.encode())
    if ntri < 0:
        raise IOError(path)
    buf = np.zeros(9 * ntri, dtype=np.float64)
    elem = np.zeros((ntri, 3), dtype=np.int32, order="F")
    nn = C.c_int(0)
    if L.orc_stl_read(path.encode(), _dp(buf), _ip(elem), C.byref(nn)) < 0:
        raise IOError(path)
    n = nn.value
    surfX = np.asfortranarray(buf[: 3 * n].reshape((n, 3), order="F"))
    return surfX, elem


def grid_from_surface(surfX, dx=0.05, dd=10):
    """set3d.f90:90-186,301.  Returns dict(nx,ny,nz,xLo,box=(im,ip,jm,jp,km,kp),dxx)."""
    L = lib()
    surfX = np.asfortranarray(surfX, dtype=np.float64)
    n = (C.c_int * 3)()
    xlo = np.zeros(3)
    box = (C.c_int * 6)()
    dxx = C.c_double(0)
    L.orc_grid_from_surface(_dp(surfX), surfX.shape[0], dx, dd, n, _dp(xlo), box, C.byref(dxx))
    return dict(nx=n[0], ny=n[1], nz=n[2], xLo=xlo, box=tuple(box), dxx=dxx.value, dx=dx)


def sign_init(phi, xLo, dx, surfX, surfElem, box):
    nx, ny, nz = _chk(phi)
    surfX = np.asfortranarray(surfX, dtype=np.float64)
    surfElem = np.asfortranarray(surfElem, dtype=np.int32)
    xLo = np.ascontiguousarray(xLo, dtype=np.float64)
    lib().orc_sign_init(_dp(phi), nx, ny, nz, _dp(xLo), dx, _dp(surfX), surfX.shape[0],
                        _ip(surfElem), surfElem.shape[0], *[int(b) for b in box])
    return phi


def reinit(phi, iter, dx, h, tol=1.0e-5, order=0, bc=1, want_grad=False):
    """subs.f90:717-931.  Returns (status, n_exit, rms_hist[:n_exit+1]) and optionally gradPhi, gradPhiMag."""
    nx, ny, nz = _chk(phi)
    hist = np.zeros(iter + 1)
    nexit = C.c_int(-1)
    g = gm = None
    if want_grad:
        g = np.zeros(phi.shape + (3,), order="F")
        gm = np.zeros(phi.shape, order="F")
    st = lib().orc_reinit(_dp(phi), _dp(g), _dp(gm), nx, ny, nz, iter, dx, h, tol, order, bc, C.byref(nexit), _dp(hist))
    if st < 0:
        raise MemoryError
    out = (st, nexit.value, hist[: nexit.value + 1].copy())
    return out + (g, gm) if want_grad else out


def reinit_sweep(phi, phiS, dx, h, raster, hyperplane=False):
    nx, ny, nz = _chk(phi)
    f = lib().orc_reinit_sweep_hyperplane if hyperplane else lib().orc_reinit_sweep
    f(_dp(phi), _dp(phiS), nx, ny, nz, dx, h, raster, None, None)
    return phi


def bc(phi, dx, literal=False):
    nx, ny, nz = _chk(phi)
    (lib().orc_bc_literal if literal else lib().orc_bc_closed)(_dp(phi), nx, ny, nz, dx)
    return phi


def rms(phi, phiN):
    nx, ny, nz = _chk(phi)
    return lib().orc_rms(_dp(phi), _dp(phiN), nx, ny, nz)


def narrowband(phi, dx):
    nx, ny, nz = _chk(phi)
    nb = np.zeros(phi.shape, dtype=np.int32, order="F")
    sb = np.zeros(phi.shape, dtype=np.int32, order="F")
    lib().orc_narrowband(nx, ny, nz, dx, _dp(phi), _ip(nb), _ip(sb))
    return nb, sb


def minmax(phi, iter, dx, h1, tol=1.0e-7):
    """set3d.f90:357-462.  Returns (status, n_exit, rms_hist[:n_exit], phiNB, phiSB)."""
    nx, ny, nz = _chk(phi)
    nb, sb = narrowband(phi, dx)
    phiN = phi.copy(order="F")
    hist = np.zeros(max(iter, 1))
    nexit = C.c_int(-1)
    st = lib().orc_minmax(_dp(phi), _dp(phiN), _ip(nb), _ip(sb), nx, ny, nz, iter, dx, h1, tol, C.byref(nexit), _dp(hist))
    if st == -1:
        raise MemoryError
    return st, nexit.value, hist[: nexit.value].copy(), nb, sb


def weno_gm(phi, i, j, k, dx):
    nx, ny, nz = _chk(phi)
    return lib().orc_weno(i, j, k, nx, ny, nz, dx, _dp(phi), None, None)


def advect_nodes(phi, phiSB, xLo, dx, surfX, iter=1000, literal=False):
    """set3d.f90:465-501 (firstDeriv order 8 on the stencil band, setPhiSurf, the node loop).
    Returns (status, surfXX, phiSurf, gradPhiSurf, n_moves); status -3 / -5: the reference would read out of bounds."""
    nx, ny, nz = _chk(phi)
    phiSB = np.asfortranarray(phiSB, dtype=np.int32)
    grad = np.zeros(phi.shape + (3,), order="F")                      # gradPhi = 0., set3d.f90:372
    lib().orc_gradphi_band8(nx, ny, nz, dx, _dp(phi), _ip(phiSB), _dp(grad))      # band cells it cannot evaluate in bounds are poisoned
    st = 0
    XX = np.asfortranarray(surfX, dtype=np.float64).copy(order="F")
    n = XX.shape[0]
    ps = np.zeros(n)
    gs = np.zeros((n, 3), order="F")
    moves = C.c_longlong(0)
    if st == 0:
        xLo = np.ascontiguousarray(xLo, dtype=np.float64)
        st = lib().orc_advect_nodes(_dp(xLo), nx, ny, nz, dx, _dp(phi), _dp(grad), n, _dp(XX), _dp(ps), _dp(gs), int(iter),
                                    1 if literal else 0, C.byref(moves))
    return st, XX, ps, gs, moves.value
