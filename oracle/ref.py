"""ctypes front-end for oracle/_ref/libref.so -- the reference's own Fortran text, machine-translated to C by
oracle/f90_to_c.py and compiled here (see that script's docstring).  TEST INFRASTRUCTURE ONLY: imported by
tests/, tests/golden/make_golden.py, __graft_entry__ (build + smoke) and bench.py's CPU legs, never by the product.

libref.so is built in the BUILD container (where /root/reference exists) and travels to the GPU box as a
prebuilt, git-ignored file, like the product's own .so; `available()` says whether it is there.

Two ways in:
  * module procedures of subs.f90, called with Fortran's by-reference convention through their guarded
    entries `ref_call_<name>` (STOP comes back as status 1): `reinit`, `narrowband`, `weno_gm`, ...
  * the main program set3d.f90 as a line-addressable function: `Program.run(first, last)` executes the top-level
    statements that start on those source lines against the program's variables, which `Program.set/get` expose.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libref.so")
REFERENCE_DIR = "/root/reference"
_lib = None


class FDesc(C.Structure):
    _fields_ = [("base", C.c_void_p), ("rank", C.c_int), ("lb", C.c_long * 4), ("ext", C.c_long * 4),
                ("elsz", C.c_size_t), ("raw", C.c_void_p)]


class PrintItem(C.Structure):
    _fields_ = [("line", C.c_int), ("kind", C.c_int), ("i", C.c_long), ("r", C.c_double), ("s", C.c_char * 96)]


class VarEntry(C.Structure):
    _fields_ = [("name", C.c_char_p), ("kind", C.c_int), ("ptr", C.c_void_p), ("n", C.c_long), ("elt", C.c_int)]


_ELT = {0: np.int32, 1: np.float64, 2: np.float32, 3: np.int16}
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def build(force: bool = False) -> bool:
    """(Re)build libref.so when the reference tree is present; returns available()."""
    if os.path.isdir(REFERENCE_DIR):
        stale = force or not os.path.exists(_LIB_PATH)
        if not stale:
            t = os.path.getmtime(_LIB_PATH)
            deps = [os.path.join(_HERE, f) for f in ("f90_to_c.py", "ref_runtime.c", "ref_runtime.h", "Makefile")]
            deps += [os.path.join(REFERENCE_DIR, f) for f in ("subs.f90", "set3d.f90")]
            stale = any(os.path.getmtime(d) > t for d in deps)
        if stale:
            subprocess.check_call(["make", "-C", _HERE, "-B", "ref"], stdout=subprocess.DEVNULL)
    return available()


def available() -> bool:
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        build()
        if not available():
            raise RuntimeError("oracle/_ref/libref.so is missing (it is built where /root/reference exists)")
        L = C.CDLL(_LIB_PATH)
        L.ref_print_count.restype = C.c_long
        L.ref_print_get.restype = C.POINTER(PrintItem)
        L.ref_print_get.argtypes = [C.c_long]
        L.ref_var.restype = C.POINTER(VarEntry)
        L.ref_var.argtypes = [C.c_char_p]
        L.f_allocate.argtypes = [C.POINTER(FDesc), C.c_int, C.c_size_t, C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.f_deallocate.argtypes = [C.POINTER(FDesc)]
        L.ref_program_exec.argtypes = [C.c_int, C.c_int]
        L.ref_set_arg.argtypes = [C.c_int, C.c_char_p]
        L.f_set_outdir.argtypes = [C.c_char_p]
        _lib = L
    return _lib


# ------------------------------------------------------------------------------------------- PRINT capture
def prints(clear: bool = True):
    """-> list of records, each a list of python values, in the order the PRINT statements executed"""
    L = lib()
    recs, cur = [], []
    for q in range(L.ref_print_count()):
        it = L.ref_print_get(q).contents
        if it.kind == 3:
            recs.append((it.line, cur))
            cur = []
        else:
            cur.append(it.s.decode(errors="replace") if it.kind == 0 else (int(it.i) if it.kind == 1 else float(it.r)))
    if clear:
        L.ref_print_clear()
    return recs


def _iteration_history(recs):
    """(n, phiErr) pairs of the reference's `PRINT*, " Iteration: ",n," "," RMS Error: ",phiErr` lines"""
    ns, errs = [], []
    for _, r in recs:
        if len(r) == 5 and isinstance(r[0], str) and r[0].strip() == "Iteration:":
            ns.append(r[1])
            errs.append(r[4])
    return ns, np.array(errs)


def _chk(phi):
    assert phi.dtype == np.float64 and phi.flags.f_contiguous and phi.ndim == 3
    return [C.c_int32(s - 1) for s in phi.shape]


def _ref(x):
    return C.byref(x)


# ------------------------------------------------------------------------------------------- subs.f90 procedures
def phisign(pS, dxx, gM):
    """subs.f90:152-172"""
    s = C.c_double(0)
    lib().ref_call_phisign(_ref(C.c_double(pS)), _ref(s), _ref(C.c_double(dxx)), _ref(C.c_double(gM)))
    return s.value


def reinit(phi, iter, dx, h, want_grad=False):
    """subs.f90:717-931 on `phi` in place.  -> (status, n_exit, rms_hist[:n_exit]) where status 0 = steady state
    EXIT at sweep n_exit (its RMS is not printed by the reference), 1 = NaN STOP after printing sweep n_exit,
    2 = ran all iter+1 sweeps; rms_hist = the printed RMS errors."""
    L = lib()
    nx, ny, nz = _chk(phi)
    g = np.zeros(phi.shape + (3,), order="F")
    gm = np.zeros(phi.shape, order="F")
    L.ref_print_clear()
    st = L.ref_call_reinit(phi.ctypes.data_as(_dp), g.ctypes.data_as(_dp), gm.ctypes.data_as(_dp), _ref(nx), _ref(ny),
                           _ref(nz), _ref(C.c_int32(iter)), _ref(C.c_double(dx)), _ref(C.c_double(h)))
    recs = prints()
    ns, errs = _iteration_history(recs)
    steady = any(isinstance(r[0], str) and "steady state" in r[0] for _, r in recs if r)
    if st == 1:
        out = (1, ns[-1], errs)
    elif steady:
        out = (0, len(ns), errs)
    else:
        out = (2, iter, errs)
    return out + (g, gm) if want_grad else out


def narrowband(phi, dx):
    """subs.f90:178-207"""
    nx, ny, nz = _chk(phi)
    nb = np.full(phi.shape, -7, dtype=np.int32, order="F")
    sb = np.full(phi.shape, -7, dtype=np.int32, order="F")
    lib().ref_call_narrowband(_ref(nx), _ref(ny), _ref(nz), _ref(C.c_double(dx)), phi.ctypes.data_as(_dp),
                              nb.ctypes.data_as(_ip), sb.ctypes.data_as(_ip))
    return nb, sb


def weno_gm(phi, i, j, k, dx):
    """subs.f90:489-711 -> gM (also returns the gradPhi(i,j,k,1:3) it stored)"""
    nx, ny, nz = _chk(phi)
    g = np.zeros(phi.shape + (3,), order="F")
    gm = np.zeros(phi.shape, order="F")
    out = C.c_double(0)
    lib().ref_call_weno(_ref(out), _ref(C.c_int32(i)), _ref(C.c_int32(j)), _ref(C.c_int32(k)), _ref(nx), _ref(ny), _ref(nz),
                        _ref(C.c_double(dx)), phi.ctypes.data_as(_dp), g.ctypes.data_as(_dp), gm.ctypes.data_as(_dp))
    return out.value, g[i, j, k, :].copy()


def secondderiv(phi, i, j, k, dx):
    """subs.f90:370-407 -> (phiXX, phiYY, phiZZ, phiXY, phiXZ, phiYZ)"""
    nx, ny, nz = _chk(phi)
    o = [C.c_double(0) for _ in range(6)]
    lib().ref_call_secondderiv(_ref(C.c_int32(i)), _ref(C.c_int32(j)), _ref(C.c_int32(k)), _ref(nx), _ref(ny), _ref(nz),
                               _ref(C.c_double(dx)), phi.ctypes.data_as(_dp), *[_ref(x) for x in o], _ref(C.c_int32(2)))
    return tuple(x.value for x in o)


def firstderiv(phi, i, j, k, dx, order):
    """subs.f90:213-364 -> (phiX, phiY, phiZ, gMM)"""
    nx, ny, nz = _chk(phi)
    g = np.zeros(phi.shape + (3,), order="F")
    o = [C.c_double(0) for _ in range(4)]
    st = lib().ref_call_firstderiv(_ref(C.c_int32(i)), _ref(C.c_int32(j)), _ref(C.c_int32(k)), _ref(nx), _ref(ny), _ref(nz),
                                   _ref(C.c_double(dx)), phi.ctypes.data_as(_dp), _ref(o[0]), _ref(o[1]), _ref(o[2]),
                                   _ref(C.c_int32(order)), _ref(o[3]), g.ctypes.data_as(_dp))
    assert st == 0
    return tuple(x.value for x in o)


def setphisurf(xLo, phi, gradPhi, dx, surfX):
    """subs.f90:1057-1170 -> (phiSurf, gradPhiSurf)"""
    nx, ny, nz = _chk(phi)
    surfX = np.asfortranarray(surfX, dtype=np.float64)
    n = surfX.shape[0]
    ps = np.zeros(n)
    gs = np.zeros((n, 3), order="F")
    xLo = np.ascontiguousarray(xLo, dtype=np.float64)
    lib().ref_call_setphisurf(xLo.ctypes.data_as(_dp), _ref(nx), _ref(ny), _ref(nz), _ref(C.c_double(dx)),
                              ps.ctypes.data_as(_dp), phi.ctypes.data_as(_dp), _ref(C.c_int32(n)),
                              surfX.ctypes.data_as(_dp), gs.ctypes.data_as(_dp), gradPhi.ctypes.data_as(_dp))
    return ps, gs


# ------------------------------------------------------------------------------------------- set3d.f90, by line range
class Program:
    """The translated PROGRAM set3d.  One instance at a time (its variables are C globals)."""

    # top-level blocks of set3d.f90 (first line, last line)
    IMPORT = (52, 82)           # getarg, stlRead
    BBOX_GRID = (86, 173)       # bounding box, nx/ny/nz, xLo, phi = 1., gridX
    SIGN = (176, 268)           # sub-box, centroids, the inside/outside search
    REINIT1 = (276, 311)        # allocations, dxx, h, CALL reinit, phiO = phi
    VTI1 = (320, 351)
    BAND_INIT = (353, 384)      # narrowBand, allocations, zeroing, phiN = phi
    MINMAX = (386, 463)         # the min/max flow loop
    NODES = (465, 501)          # firstDeriv(8) on the stencil band, setPhiSurf, node loop
    ASYMPTOTIC = (504, 521)
    GRADMAG = (523, 536)        # firstDeriv(2) over ALL points (reads phi(-1,..): poisoned, result unused)
    VTI2 = (539, 569)
    REINIT2 = (571, 582)
    S3D = (584, 614)

    def __init__(self, outdir: str | None = None):
        self.L = lib()
        self.L.ref_program_reset()
        self.L.ref_print_clear()
        self.L.f_set_outdir((outdir or "").encode())

    def _var(self, name) -> VarEntry:
        e = self.L.ref_var(name.lower().encode())
        if not e:
            raise KeyError(name)
        return e.contents

    def run(self, first: int, last: int) -> int:
        """0 = ran through, 1 = the program executed STOP"""
        return self.L.ref_program_exec(first, last)

    def set_arg(self, filename: str):
        self.L.ref_set_arg(1, filename.encode())

    def get(self, name):
        e = self._var(name)
        if e.kind in (0, 1):
            return np.ctypeslib.as_array(C.cast(e.ptr, C.POINTER(np.ctypeslib.as_ctypes_type(_ELT[e.elt]))), (1,))[0].item()
        if e.kind == 3:
            return C.string_at(e.ptr, e.n).decode(errors="replace")
        if e.kind == 4:
            return np.ctypeslib.as_array(C.cast(e.ptr, C.POINTER(np.ctypeslib.as_ctypes_type(_ELT[e.elt]))), (e.n,)).copy()
        d = C.cast(e.ptr, C.POINTER(FDesc)).contents
        if not d.base:
            return None
        shape = tuple(d.ext[r] for r in range(d.rank))
        n = int(np.prod(shape))
        flat = np.ctypeslib.as_array(C.cast(d.base, C.POINTER(np.ctypeslib.as_ctypes_type(_ELT[e.elt]))), (max(n, 1),))[:n]
        return flat.reshape(shape, order="F").copy(order="F")

    def set(self, name, value, lower=None):
        e = self._var(name)
        if e.kind in (0, 1, 4):
            a = np.ctypeslib.as_array(C.cast(e.ptr, C.POINTER(np.ctypeslib.as_ctypes_type(_ELT[e.elt]))), (e.n,))
            a[...] = value
            return
        if e.kind == 3:
            b = str(value).encode().ljust(e.n)[: e.n]
            C.memmove(e.ptr, b, e.n)
            return
        v = np.asfortranarray(value, dtype=_ELT[e.elt])
        d = C.cast(e.ptr, C.POINTER(FDesc))
        self.L.f_deallocate(d)
        lb = list(lower) if lower is not None else [1] * v.ndim
        lbs = (C.c_long * 4)(*(lb + [0] * (4 - v.ndim)))
        ubs = (C.c_long * 4)(*([l + s - 1 for l, s in zip(lb, v.shape)] + [0] * (4 - v.ndim)))
        self.L.f_allocate(d, v.ndim, v.itemsize, lbs, ubs)
        C.memmove(d.contents.base, v.ctypes.data, v.nbytes)

    def prints(self):
        return prints()

    def iteration_history(self, recs=None):
        return _iteration_history(recs if recs is not None else self.prints())


def stl_read(path: str):
    """stlRead (subs.f90:17-121) through the program's own call at set3d.f90:68 -> surfX, surfElem (1-based)"""
    p = Program()
    p.set_arg(path)
    assert p.run(*Program.IMPORT) == 0
    return p.get("surfX"), p.get("surfElem")
