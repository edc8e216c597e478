"""Host-side mirror of the reference's `MODULE set_subs` interface for the grid hot path.

Same names, argument order and in-place (INTENT(INOUT)) behaviour as the Fortran procedures the
CUDA library replaces, so calls read like set3d.f90's:

    reinit(phi, gradPhi, gradPhiMag, nx, ny, nz, iter, dx, h)         # subs.f90:717
    narrowBand(nx, ny, nz, dx, phi, phiNB, phiSB)                     # subs.f90:178
    signSearch(phi, nx, ny, nz, xLo, dx, surfX, surfElem, box)        # inline loop set3d.f90:196-268
    minMaxFlow(phi, phiN, phiNB, phiSB, nx, ny, nz, iter, dx, h1)     # inline loop set3d.f90:394-462

Arrays are numpy, Fortran-ordered, shape (nx+1, ny+1, nz+1) == phi(0:nx,0:ny,0:nz); REAL -> float64,
INTEGER -> int32.  Everything runs on the GPU through the C ABI (include/lsf_b200.h); a NaN RMS is
reported the way the reference reports it -- by stopping: `ReferenceStop` is raised after the
arrays have been updated (subs.f90:926, set3d.f90:458).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import c_double_p, c_i32_p, check, lib


class ReferenceStop(RuntimeError):
    """The reference would execute STOP here (NaN RMS)."""

    def __init__(self, where, n, rms_hist):
        super().__init__(f"{where}: RMS error is NaN at iteration {n} (reference STOPs)")
        self.n = n
        self.rms_hist = rms_hist


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_i32_p)


def _grid_array(a, nx, ny, nz, dtype, name, extra=()):
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags.f_contiguous \
            or a.shape != (nx + 1, ny + 1, nz + 1) + tuple(extra):
        raise ValueError(f"{name} must be a Fortran-ordered {np.dtype(dtype).name} array of shape "
                         f"{(nx + 1, ny + 1, nz + 1) + tuple(extra)}")
    return a


def set_arith(exact):
    """True / "exact": bit-identical reference arithmetic; False / "fast": FMA + reciprocal-reduced form;
    None / "auto" (library default): fast with a conditioning guard that falls back to exact."""
    mode = {True: _lib.ARITH_EXACT, False: _lib.ARITH_FAST, None: _lib.ARITH_AUTO,
            "exact": _lib.ARITH_EXACT, "fast": _lib.ARITH_FAST, "auto": _lib.ARITH_AUTO}[exact]
    check(lib().lsf_set_arith(mode))


def last_arith() -> str:
    """Arithmetic the most recent reinit call finished in."""
    return "exact" if lib().lsf_last_arith() == _lib.ARITH_EXACT else "fast"


def set_sched(plane: bool):
    check(lib().lsf_set_sched(_lib.SCHED_PLANE if plane else _lib.SCHED_MARCH))


def set_minmax_algo(march: bool):
    """False (default): active-list min/max iteration; True: whole-grid march kernel (cross-check)."""
    check(lib().lsf_set_minmax_algo(_lib.MINMAX_MARCH if march else _lib.MINMAX_LIST))


def set_overlap(on: bool):
    """True: `reinit` (fp64, single GPU, march schedule) runs its sweeps in overlapped batches of 8 (one launch per batch;
    CTAs go on to the next sweep's tiles while the previous sweep drains).  Same results; opt-in."""
    check(lib().lsf_set_overlap(1 if on else 0))


def set_precision(f32: bool):
    """False (default): the reference's REAL(8) on the device; True: the optional fp32 mode of the host-buffer
    `reinit` (device fields and WENO5 arithmetic in single precision, host arrays stay float64; contract
    max|phi - phi_ref| <= 1e-4 max|phi_ref|; gradPhi / gradPhiMag are not written)."""
    check(lib().lsf_set_precision(_lib.PREC_F32 if f32 else _lib.PREC_F64))


def reinit(phi, gradPhi, gradPhiMag, nx, ny, nz, iter, dx, h, stop_on_nan=True):
    """SUBROUTINE reinit (subs.f90:717-931).  Returns (n_exit, rms_hist) -- the iteration index at
    which the loop left and the RMS errors the reference prints at subs.f90:923."""
    _grid_array(phi, nx, ny, nz, np.float64, "phi")
    if gradPhi is not None:
        _grid_array(gradPhi, nx, ny, nz, np.float64, "gradPhi", (3,))
    if gradPhiMag is not None:
        _grid_array(gradPhiMag, nx, ny, nz, np.float64, "gradPhiMag")
    hist = np.zeros(iter + 1)
    n_exit = C.c_int(-1)
    rc = check(lib().lsf_reinit(_dp(phi), _dp(gradPhi), _dp(gradPhiMag), nx, ny, nz, int(iter), float(dx), float(h),
                                C.byref(n_exit), _dp(hist)))
    hist = hist[: n_exit.value + 1]
    if rc == _lib.LSF_NAN and stop_on_nan:
        raise ReferenceStop("reinit", n_exit.value, hist)
    return n_exit.value, hist


def narrowBand(nx, ny, nz, dx, phi, phiNB, phiSB):
    """SUBROUTINE narrowBand (subs.f90:178-207)."""
    _grid_array(phi, nx, ny, nz, np.float64, "phi")
    _grid_array(phiNB, nx, ny, nz, np.int32, "phiNB")
    _grid_array(phiSB, nx, ny, nz, np.int32, "phiSB")
    check(lib().lsf_narrowband(nx, ny, nz, float(dx), _dp(phi), _ip(phiNB), _ip(phiSB)))


def signSearch(phi, nx, ny, nz, xLo, dx, surfX, surfElem, box):
    """The inside/outside search loop of set3d.f90:196-268.  box = (im, ip, jm, jp, km, kp)."""
    _grid_array(phi, nx, ny, nz, np.float64, "phi")
    surfX = np.asfortranarray(surfX, dtype=np.float64)
    surfElem = np.asfortranarray(surfElem, dtype=np.int32)
    xLo = np.ascontiguousarray(xLo, dtype=np.float64)
    check(lib().lsf_sign_init(_dp(phi), nx, ny, nz, _dp(xLo), float(dx), _dp(surfX), surfX.shape[0],
                              _ip(surfElem), surfElem.shape[0], *[int(v) for v in box]))


def minMaxFlow(phi, phiN, phiNB, phiSB, nx, ny, nz, iter, dx, h1, tol=1.0e-7, stop_on_nan=True):
    """The min/max time loop of set3d.f90:394-462.  Returns (n_exit, rms_hist)."""
    _grid_array(phi, nx, ny, nz, np.float64, "phi")
    _grid_array(phiN, nx, ny, nz, np.float64, "phiN")
    _grid_array(phiNB, nx, ny, nz, np.int32, "phiNB")
    _grid_array(phiSB, nx, ny, nz, np.int32, "phiSB")
    hist = np.zeros(max(int(iter), 1))
    n_exit = C.c_int(-1)
    rc = check(lib().lsf_minmax(_dp(phi), _dp(phiN), _ip(phiNB), _ip(phiSB), nx, ny, nz, int(iter), float(dx),
                                float(h1), float(tol), C.byref(n_exit), _dp(hist)))
    hist = hist[: n_exit.value]
    if rc == _lib.LSF_NAN and stop_on_nan:
        raise ReferenceStop("minMaxFlow", n_exit.value, hist)
    return n_exit.value, hist


def advectNodes(phi, phiSB, nx, ny, nz, xLo, dx, surfX, iter=1000):
    """The "Advect Nodes" block of set3d.f90:465-501: firstDeriv(order 8) on the stencil band, setPhiSurf, and the
    node loop (every node with phiSurf > 1E-13 moves by phiSurf*gradPhiSurf).  surfX (nSurfNode,3) is not modified;
    returns (surfXX, phiSurf, gradPhiSurf, n_moves) as the reference leaves them."""
    _grid_array(phi, nx, ny, nz, np.float64, "phi")
    _grid_array(phiSB, nx, ny, nz, np.int32, "phiSB")
    XX = np.asfortranarray(surfX, dtype=np.float64).copy(order="F")     # surfXX = surfX, set3d.f90:485
    n = XX.shape[0]
    ps = np.zeros(n)
    gs = np.zeros((n, 3), order="F")
    xLo = np.ascontiguousarray(xLo, dtype=np.float64)
    moves = C.c_longlong(0)
    check(lib().lsf_advect_nodes(_dp(phi), _ip(phiSB), nx, ny, nz, _dp(xLo), float(dx), _dp(XX), n, _dp(ps), _dp(gs),
                                 int(iter), C.byref(moves)))
    return XX, ps, gs, moves.value


class DeviceGrid:
    """A device-resident phi(0:nx,0:ny,0:nz): sign search -> reinit -> min/max without host round trips."""

    def __init__(self, nx, ny, nz, f32=False):
        """f32=True: the optional single-precision mode (phi stored as float on the device, reinit on the FP32
        pipe; upload / download still take float64 host arrays)."""
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.f32 = bool(f32)
        self._h = C.c_void_p()
        create = lib().lsf_grid_create_f32 if self.f32 else lib().lsf_grid_create
        check(create(C.byref(self._h), self.nx, self.ny, self.nz))

    @property
    def shape(self):
        return (self.nx + 1, self.ny + 1, self.nz + 1)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().lsf_grid_destroy(self._h)
            self._h = None

    __del__ = close

    def fill(self, value):
        check(lib().lsf_grid_fill(self._h, float(value)))

    def upload(self, phi):
        _grid_array(phi, self.nx, self.ny, self.nz, np.float64, "phi")
        check(lib().lsf_grid_upload(self._h, phi.ctypes.data))

    def upload_ptr(self, host_ptr: int):
        check(lib().lsf_grid_upload(self._h, host_ptr))

    def download_ptr(self, host_ptr: int):
        check(lib().lsf_grid_download(self._h, host_ptr))

    def download(self, out=None):
        if out is None:
            out = np.empty(self.shape, order="F")
        _grid_array(out, self.nx, self.ny, self.nz, np.float64, "out")
        check(lib().lsf_grid_download(self._h, out.ctypes.data))
        return out

    def device_ptr(self) -> int:
        return int(lib().lsf_grid_device_ptr(self._h))

    def checksum(self):
        """(sum, xor) digest of this rank's owned points of phi (lsf_grid_checksum): position-weighted wrapping sum and xor
        of the bit patterns; the ranks' digests of a sharded grid add / xor up to the single-GPU digest."""
        d = (C.c_uint64 * 2)()
        check(lib().lsf_grid_checksum(self._h, d))
        return int(d[0]), int(d[1])

    def signSearch(self, xLo, dx, surfX, surfElem, box):
        surfX = np.asfortranarray(surfX, dtype=np.float64)
        surfElem = np.asfortranarray(surfElem, dtype=np.int32)
        xLo = np.ascontiguousarray(xLo, dtype=np.float64)
        check(lib().lsf_grid_sign_init(self._h, _dp(xLo), float(dx), _dp(surfX), surfX.shape[0], _ip(surfElem),
                                       surfElem.shape[0], *[int(v) for v in box]))

    def reinit(self, iter, dx, h, tol=1.0e-5):
        hist = np.zeros(iter + 1)
        n_exit = C.c_int(-1)
        rc = check(lib().lsf_grid_reinit(self._h, int(iter), float(dx), float(h), float(tol), C.byref(n_exit), _dp(hist)))
        return rc, n_exit.value, hist[: n_exit.value + 1]

    def reinitRK3(self, steps, dx, dt, tol=1.0e-5):
        """K2' throughput mode (lsf_grid_reinit_rk3): Jacobi WENO5 + TVD-RK3 -- the north_star's literal scheme, NOT the
        reference's Gauss-Seidel algorithm.  Returns (rc, n_exit, rms_hist[:n_exit+1])."""
        hist = np.zeros(max(int(steps), 1))
        n_exit = C.c_int(-1)
        rc = check(lib().lsf_grid_reinit_rk3(self._h, int(steps), float(dx), float(dt), float(tol), C.byref(n_exit), _dp(hist)))
        return rc, n_exit.value, hist[: n_exit.value + 1]

    def narrowBand(self, dx):
        nb = np.empty(self.shape, dtype=np.int32, order="F")
        sb = np.empty(self.shape, dtype=np.int32, order="F")
        check(lib().lsf_grid_narrowband(self._h, float(dx), _ip(nb), _ip(sb)))
        return nb, sb

    def minMaxFlow(self, iter, dx, h1, tol=1.0e-7):
        hist = np.zeros(max(int(iter), 1))
        n_exit = C.c_int(-1)
        rc = check(lib().lsf_grid_minmax(self._h, int(iter), float(dx), float(h1), float(tol), C.byref(n_exit), _dp(hist)))
        return rc, n_exit.value, hist[: n_exit.value]


def _advect_nodes_method(self, xLo, dx, surfX, iter=1000):
    """set3d.f90:465-501 on the resident phi (see advectNodes); returns (surfXX, phiSurf, gradPhiSurf, n_moves)."""
    XX = np.asfortranarray(surfX, dtype=np.float64).copy(order="F")
    n = XX.shape[0]
    ps = np.zeros(n)
    gs = np.zeros((n, 3), order="F")
    xLo = np.ascontiguousarray(xLo, dtype=np.float64)
    moves = C.c_longlong(0)
    check(lib().lsf_grid_advect_nodes(self._h, _dp(xLo), float(dx), _dp(XX), n, _dp(ps), _dp(gs), int(iter), C.byref(moves)))
    return XX, ps, gs, moves.value


DeviceGrid.advectNodes = _advect_nodes_method


def slab_range(nz, nranks, rank):
    """Planes [k0, k1) of phi(0:nx,0:ny,0:nz) that `rank` of `nranks` owns (lsf_slab_range)."""
    k0, k1 = C.c_int(0), C.c_int(0)
    check(lib().lsf_slab_range(int(nz), int(nranks), int(rank), C.byref(k0), C.byref(k1)))
    return k0.value, k1.value


class ShardedGrid(DeviceGrid):
    """One z-slab of a global phi(0:nx,0:ny,0:nz), one per process / GPU (include/lsf_b200.h, z-slab
    sharding).  Same methods as DeviceGrid; every rank makes the same calls in the same order, host arrays
    hold the rank's OWNED planes k0..k1-1 -- shape (nx+1, ny+1, k1-k0) -- and n_exit / rms_hist are identical
    on all ranks.  `torch.distributed` must be initialised: it is used once, to all-gather the CUDA-IPC
    handles; the data path itself is peer stores over NVLink from inside the kernels."""

    def __init__(self, nx, ny, nz, rank=None, nranks=None, f32=False):
        import torch
        import torch.distributed as dist
        self.rank = dist.get_rank() if rank is None else int(rank)
        self.nranks = dist.get_world_size() if nranks is None else int(nranks)
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.f32 = bool(f32)          # optional fp32 mode (same calls)
        self.k0, self.k1 = slab_range(self.nz, self.nranks, self.rank)
        self._h = C.c_void_p()
        create = lib().lsf_sgrid_create_f32 if self.f32 else lib().lsf_sgrid_create
        check(create(C.byref(self._h), self.nx, self.ny, self.nz, self.rank, self.nranks))
        if self.nranks > 1:
            buf = (C.c_ubyte * _lib.IPC_HANDLE_BYTES)()
            check(lib().lsf_sgrid_ipc_handle(self._h, buf))
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
            mine = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
            every = [torch.empty_like(mine) for _ in range(self.nranks)]
            dist.all_gather(every, mine)
            blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in every)
            check(lib().lsf_sgrid_attach(self._h, blob))
            dist.barrier()

    @property
    def shape(self):
        return (self.nx + 1, self.ny + 1, self.k1 - self.k0)

    def _slab_array(self, a, dtype, name):
        if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags.f_contiguous or a.shape != self.shape:
            raise ValueError(f"{name} must be a Fortran-ordered {np.dtype(dtype).name} array of shape {self.shape} "
                             f"(planes {self.k0}..{self.k1 - 1} of the global grid)")
        return a

    def upload(self, phi):
        self._slab_array(phi, np.float64, "phi")
        check(lib().lsf_grid_upload(self._h, phi.ctypes.data))

    def download(self, out=None):
        if out is None:
            out = np.empty(self.shape, order="F")
        self._slab_array(out, np.float64, "out")
        check(lib().lsf_grid_download(self._h, out.ctypes.data))
        return out

    def sync_ghosts(self):
        check(lib().lsf_sgrid_sync_ghosts(self._h))

    def close(self):
        """Collective: no rank may free its slab while a neighbour can still store into it."""
        if getattr(self, "_h", None) is not None and self._h:
            if self.nranks > 1:
                import torch.distributed as dist
                if dist.is_initialized():
                    lib().lsf_sgrid_sync_ghosts(self._h)
                    dist.barrier()
            lib().lsf_grid_destroy(self._h)
            self._h = None

    __del__ = close
