#!/usr/bin/env python
"""bench.py -- WENO5 Gauss-Seidel reinitialisation throughput (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU

A step = one raster cycle (8 in-place sweeps, subs.f90:742-852, each followed by the boundary
block and the RMS test) of `reinit` over one synthetic grid: the torus+cube STL of BASELINE
configs 4/5 on a 1024 x 1024 x 1024 fp64 grid per GPU (weak scaling; `--grid` overrides).  The sign
field fed to reinit is produced by the library's own sign search from the synthetic STL.

  value : Gcell-updates/s, phi resident in HBM, CUDA-event time of the K steps, max over ranks
  e2e   : same metric through the host-buffer drop-in call lsf_reinit (pinned host phi in, phi out:
          H2D + 8 sweeps + D2H inside the timed region)
  roofline : the sweep kernel's algorithmic HBM bytes (24 B per cell update: read phi, read the
          frozen sign source, write phi) / its mean launch time, vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : the reference itself -- subs.f90 `reinit`, machine-translated to C (oracle/_ref/libref.so, built where
          /root/reference exists and shipped prebuilt; kind "reference") -- on a bounded slab sample of the same
          workload, one host core (the reference is serial); falls back to the hand-written C port (kind "port")
  parity : (a) the timed 8 sweeps repeated in EXACT arithmetic on a second 1024^3 grid: max|FAST - EXACT|; (b) the CPU
          sample slab swept by the GPU as a grid of its own: max|GPU - reference|, EXACT bit-identical
  strong : BASELINE config 4 -- ONE 1024^3 grid, reinit (8 sweeps) + 64 min/max iterations, cut into N z-slabs --
          with a partition-independent digest of phi after each stage (equal for N = 1, 2, 4, 8 <=> bit-identical)
  config3 : BASELINE config 3 -- 20k-triangle sphere on 512^3, reinit-only (N = 1)

One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DX = 0.05
SWEEPS_PER_STEP = 8
METRIC = "WENO5 reinit Gcell-updates/s"
UNIT = "Gcell-updates/s"
BYTES_PER_UPDATE = 24.0          # fp64: read phi + read phiS + write phi
FP64_PER_UPDATE = 272            # executed FP64-pipe instructions per cell update of the FAST sweep kernel (ncu source page of the
                                 # r2c capture, profiles/r2c_march_1024_full.txt: DFMA 102.9 + DMUL 88.4 + DADD 77.1 + DSETP 4.1; round 1: 307)
FP64_PIPE_PEAK = 148 * 64 * 1.965e9   # lane-ops/s


# ----------------------------------------------------------------------------------- helpers
def analytic_sign_field(shape, dx=DX):
    """Smeared sign (phiSign with gM=1, set3d.f90:260-264) of an analytic torus+cube placed like
    stl.torus_cube_config -- used for the CPU legs, which must not touch the GPU kernels."""
    nxp, nyp, nzp = shape
    ex, ey, ez = ((n - 22.2) * dx for n in shape)
    s = min(ex, ey)
    r = 0.12 * s
    R = 0.5 * s - r
    x = (np.arange(nxp) * dx - 10 * dx - 0.5 * ex)[:, None, None]
    y = (np.arange(nyp) * dx - 10 * dx - 0.5 * ey)[None, :, None]
    z = (np.arange(nzp) * dx - 10 * dx - 0.5 * ez)[None, None, :]
    zt = -0.5 * ez + r
    d_t = np.sqrt((np.sqrt(x * x + y * y) - R) ** 2 + (z - zt) ** 2) - r
    c = 0.18 * s
    q = np.stack(np.broadcast_arrays(np.abs(x - (0.5 * ex - 0.5 * c)) - 0.5 * c,
                                     np.abs(y - (0.5 * ey - 0.5 * c)) - 0.5 * c,
                                     np.abs(z - (0.5 * ez - 0.5 * c)) - 0.5 * c))
    d_c = np.linalg.norm(np.maximum(q, 0), axis=0) + np.minimum(q.max(axis=0), 0)
    d = np.minimum(d_t, d_c)
    return np.asfortranarray(d / np.sqrt(d * d + dx * dx))


def slab_field(nxy, nzs):
    """Smeared sign field on an nxy x nxy x nzs slab cut through the torus of the bench geometry (analytic, + - * / sqrt
    only: reproducible bit for bit).  Returns (phi, h)."""
    r_cells = int(0.12 * (nxy - 22.2))
    k0 = max(0, 10 + r_cells - nzs // 2)                      # the torus sits at the low-z end of the bbox
    ex = (nxy - 22.2) * DX
    r = 0.12 * ex
    R = 0.5 * ex - r
    x = (np.arange(nxy) * DX - 10 * DX - 0.5 * ex)[:, None, None]
    y = (np.arange(nxy) * DX - 10 * DX - 0.5 * ex)[None, :, None]
    z = ((np.arange(nzs) + k0) * DX - 10 * DX - r)[None, None, :]
    d = np.sqrt((np.sqrt(x * x + y * y) - R) ** 2 + z ** 2) - r
    phi = np.asfortranarray(d / np.sqrt(d * d + DX * DX))
    return phi, 0.1 * (DX / (np.sqrt(3.0) * ex))


def cpu_sample(nxy, nzs, sweeps, keep=False):
    """Times the reference's reinit (subs.f90:717-931) on the slab: libref.so (the reference's own text, machine-translated)
    when it is there, else the hand-written port.  Returns (Gcell-updates/s, description, kind, phi_in, phi_out)."""
    phi, h = slab_field(nxy, nzs)
    phi_in = phi.copy(order="F") if keep else None
    kind = "port"
    try:
        from oracle import ref as R
        if R.available():
            R.lib()
            kind = "reference"
    except Exception:
        kind = "port"
    t0 = time.perf_counter()
    if kind == "reference":
        R.reinit(phi, sweeps - 1, DX, h)
    else:
        from oracle import oracle as O
        O.build()
        O.reinit(phi, sweeps - 1, DX, h, tol=0.0)
    dt = time.perf_counter() - t0
    shape = phi.shape
    cells = (shape[0] - 2) * (shape[1] - 2) * (shape[2] - 2) * sweeps
    what = ("subs.f90 reinit machine-translated to C (oracle/_ref/libref.so: literal BC block, RMS pass, gradPhi stores)"
            if kind == "reference" else "hand-written C port of the reference (oracle/lsf_oracle.c)")
    return (cells / dt / 1e9, f"{sweeps} sweep(s) on a {shape[0]}x{shape[1]}x{shape[2]} slab through the torus, {dt:.1f} s, 1 core; {what}",
            kind, phi_in, phi if keep else None)


def dev_tensor(ptr, n, dtype="f8"):
    """torch view of n elements of device memory at ptr (no copy)."""
    import torch

    class _A:
        pass
    a = _A()
    a.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<" + dtype, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(a, device="cuda")


def dev_max_abs_diff(pa, pb, n, chunk=1 << 27):
    import torch
    m = 0.0
    for o in range(0, n, chunk):
        k = min(chunk, n - o)
        m = max(m, float((dev_tensor(pa + 8 * o, k) - dev_tensor(pb + 8 * o, k)).abs().max()))
    return m


class ClockSampler:
    """nvidia-smi SM clock / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.th.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for q, n in enumerate(names) if any(len(r) >= 6 and r[2 + q].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


NCU_TAG = {False: "r2c_march", True: "r2c_march_f32"}     # committed `ncu --set full` captures of the current sweep kernels


def ncu_traffic(grid, f32=False):
    """DRAM bytes (read + write) of ONE sweep-kernel launch from the committed `ncu --set full` capture of this
    kernel at the same grid size (profiles/r2c_march_<grid>_full.txt; fp32 kernel: r2c_march_f32_<grid>_full.txt),
    or None if there is no capture."""
    path = os.path.join(ROOT, "profiles", f"{NCU_TAG[bool(f32)]}_{grid}_full.txt")
    if not os.path.exists(path):
        return None
    tot = 0.0
    for line in open(path):
        f = line.split()
        if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(f[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[2]]
    return tot or None


# ----------------------------------------------------------------------------------- reference arm
def workload_config(n, world, f32, ntri=None):
    """The `config` object -- identical for this repo's arm and the reference arm."""
    return {"workload": f"BASELINE config {'4' if world == 1 else '5'}: synthetic torus+cube STL on ONE {n}x{n}x{n * world} "
                        f"{'fp32-mode' if f32 else 'fp64'} grid ({n}^3 points per GPU), reinit-only, one step = {SWEEPS_PER_STEP} "
                        "Gauss-Seidel raster sweeps (+BC+RMS each)",
            "global_grid": [n, n, n * world], "grid_per_gpu": [n, n, n], "sweeps_per_step": SWEEPS_PER_STEP, "dx": DX,
            "timed_region": "the K steps are 8K consecutive sweeps of one reinit call"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nxy = args.grid
    nzs = args.ref_slab
    rates = []
    desc, kind = "", "port"
    for q in range(args.warmup + args.steps):
        v, desc, kind, _, _ = cpu_sample(nxy, nzs, 1)
        if q >= args.warmup:
            rates.append(v)
    value = float(np.mean(rates)) if rates else 0.0
    cells = (nxy - 2) * (nxy - 2) * (nzs - 2)
    cfg = workload_config(nxy, max(args.gpus, 1), False)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cells / value / 1e6 if value else None,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
                             "sample": desc + f" per step: each step of this arm is ONE sweep (+BC+RMS) on that slab sample of the "
                                              f"{nxy}^3 grid (the serial reference needs ~15 min per full 1024^3 sweep)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from levelsetfortran_b200 import DeviceGrid, ShardedGrid, _lib, build, stl
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    L = _lib.lib()
    _lib.check(L.lsf_init(local))
    _lib.check(L.lsf_set_arith({"exact": _lib.ARITH_EXACT, "fast": _lib.ARITH_FAST, "auto": _lib.ARITH_AUTO}[args.arith]))
    _lib.check(L.lsf_set_sched(_lib.SCHED_PLANE if args.sched == "plane" else _lib.SCHED_MARCH))
    L.lsf_set_profile(1)
    if args.overlap:
        _lib.check(L.lsf_set_overlap(1))

    n = args.grid
    shape_pts = (n, n, n)                                      # per GPU (weak scaling)
    # N > 1: BASELINE config 5, ONE grid of n x n x (n*N) points cut into z-slabs, one per GPU
    tris = stl.torus_cube_config((n, n, n * world), DX)
    surfX, surfElem = stl.dedup_nodes(tris)
    g = stl.grid_from_surface(surfX, DX)
    nx, ny, nz = g["nx"], g["ny"], g["nz"]
    assert (nx + 1, ny + 1, nz + 1) == (n, n, n * world)
    h = 0.1 * g["dxx"]                                         # CFL = .1, set3d.f90:304-305
    cells_per_step = (nx - 1) * (ny - 1) * (nz - 1) * SWEEPS_PER_STEP     # whole job

    f32 = bool(args.f32)
    bytes_per_update = 12.0 if f32 else BYTES_PER_UPDATE
    if world > 1:
        G = ShardedGrid(nx, ny, nz, f32=f32)
        k_upd = min(G.k1 - 1, nz - 1) - max(G.k0, 1) + 1      # planes this rank's sweeps update
    else:
        G = DeviceGrid(nx, ny, nz, f32=f32)
        k_upd = nz - 1
    G.fill(1.0)
    t0 = time.perf_counter()
    G.signSearch(g["xLo"], DX, surfX, surfElem, g["box"])
    sign_ms, _ = _lib.last_timing()
    setup_s = time.perf_counter() - t0

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def steps(k):
        """k steps = k raster cycles = 8k consecutive sweeps of ONE reinit call (the reference's reinit runs thousands of
        sweeps per call, subs.f90:735; on z-slabs every call boundary drains the Gauss-Seidel pipeline once)"""
        rc, n_exit, hist = G.reinit(SWEEPS_PER_STEP * k - 1, DX, h, tol=0.0)     # tol 0: never EXITs early
        assert rc == 0 and n_exit == SWEEPS_PER_STEP * k - 1, (rc, n_exit)
        return hist

    def step():
        return steps(1)

    if args.warmup > 0:
        steps(args.warmup)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    hist = steps(args.steps)
    dev_ms, launches = _lib.last_timing()
    sweep_ms, n_sweeps = _lib.last_sweep_timing()
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop()
    main_arith = "exact" if L.lsf_last_arith() == _lib.ARITH_EXACT else "fast"

    # ---- BASELINE config 4 as a fixed protocol (not part of `value`): ONE n^3 grid, sign search -> 8 sweeps -> 64 min/max
    # iterations, with a partition-independent digest of phi after each stage.  N = 1: the bench grid itself; N > 1: a
    # second, strong-scaling grid cut into N z-slabs.  Equal digests for N = 1, 2, 4, 8 <=> the sharded results are
    # bit-identical to the single-GPU one.  The EXACT-arithmetic repeat of the 8 sweeps gives the FAST-vs-EXACT parity
    # number at full size (N = 1).
    def allreduce_digest(d):
        if dist is None:
            return d
        t = torch.tensor([d[0] - (1 << 64) if d[0] >= (1 << 63) else d[0]], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)                     # int64 addition wraps like the uint64 sum
        xs = [None] * world
        dist.all_gather_object(xs, d[1])
        x = 0
        for v in xs:
            x ^= v
        return int(t.item()) & ((1 << 64) - 1), x

    def hexd(d):
        return "%016x:%016x" % d

    mm = strong = parity = nodes = None
    if args.minmax_iters > 0 and not f32:
        if world == 1:
            SG, sg, sX, sE = G, g, surfX, surfElem
        else:
            stris = stl.torus_cube_config((n, n, n), DX)
            sX, sE = stl.dedup_nodes(stris)
            sg = stl.grid_from_surface(sX, DX)
            SG = ShardedGrid(sg["nx"], sg["ny"], sg["nz"])
        sh, sh1 = 0.1 * sg["dxx"], 0.01 * sg["dxx"]
        scells = (sg["nx"] - 1) * (sg["ny"] - 1) * (sg["nz"] - 1) * SWEEPS_PER_STEP
        spts = (sg["nx"] + 1) * (sg["ny"] + 1) * (sg["nz"] + 1)
        for rep in range(2):                                             # the first pass warms up (allocations, order tables)
            SG.fill(1.0)
            SG.signSearch(sg["xLo"], DX, sX, sE, sg["box"])
            barrier()
            rc, ne, shist = SG.reinit(SWEEPS_PER_STEP - 1, DX, sh, tol=0.0)
            s_ms, _ = _lib.last_timing()
            s_arith = "exact" if L.lsf_last_arith() == _lib.ARITH_EXACT else "fast"
            d_re = allreduce_digest(SG.checksum())
            barrier()
            rc2, n_mm, hist_mm = SG.minMaxFlow(args.minmax_iters, DX, sh1, tol=0.0)
            mm_ms, mm_launches = _lib.last_timing()
            d_mm = allreduce_digest(SG.checksum())
        active = int(L.lsf_last_minmax_active())
        tt = torch.tensor([s_ms, mm_ms], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        s_ms, mm_ms = (float(v) for v in tt.cpu())
        strong = {"workload": f"BASELINE config 4: torus+cube STL on ONE {n}^3 fp64 grid cut into {world} z-slab(s): sign search, "
                              f"{SWEEPS_PER_STEP} reinit sweeps, {n_mm} min/max iterations", "scaling": "strong", "n_gpus": world,
                  "reinit_ms": s_ms, "value": scells / (s_ms * 1e-3) / 1e9, "unit": UNIT, "arith_used": s_arith,
                  "minmax_ms": mm_ms, "minmax_iterations": n_mm,
                  "digest_after_reinit": hexd(d_re), "digest_after_minmax": hexd(d_mm),
                  "rms_reinit": [float(v) for v in shist], "rms_minmax_last": float(hist_mm[-1]) if len(hist_mm) else None,
                  "digest": "sum over the points of bits(phi)*(2q+1) mod 2^64 : xor of bits(phi), q = global linear index "
                            "(lsf_grid_checksum); identical across N <=> bit-identical fields"}
        mm_rate = spts * n_mm / (mm_ms * 1e-3) / 1e9
        mm = {"metric": "min/max flow Gpoint-iterations/s (all grid points per iteration)", "value": mm_rate,
              "iterations": n_mm, "ms_per_iteration": mm_ms / max(n_mm, 1), "launches": mm_launches,
              "active_cells_rank0": active, "active_fraction": active * world / spts,
              "bound": "launch/latency: the active-list algorithm touches only the narrow band (the cells that can still change, "
                       "%.2f %% of the grid), 3 launches per iteration; counted over all grid points (SURVEY 8d) that is %.0fx the "
                       "dense 16 B/point HBM roof, which therefore says nothing about this kernel"
                       % (100.0 * active * world / spts, 16.0 * mm_rate / measured_peak()[0]),
              "note": "ms_per_iteration includes building the active list once per call",
              "last_rms": float(hist_mm[-1]) if len(hist_mm) else None}
        # ---- companion: surface-node projection (set3d.f90:465-501) of the STL's own nodes on the resident field ----
        if world == 1 and not f32 and args.minmax_iters > 0:
            try:
                XX, ps_n, gs_n, n_moves = G.advectNodes(g["xLo"], DX, surfX, 1000)
                nd_ms, _nl = _lib.last_timing()
                nodes = {"metric": "surface-node projection (lsf_grid_advect_nodes), kernel ms", "ms": nd_ms, "nodes": int(len(surfX)),
                         "moves": int(n_moves), "max_displacement": float(np.abs(XX - surfX).max()),
                         "reference_cost": "O(moves x nodes) trilinear interpolations: %.3g" % (float(n_moves) * len(surfX))}
            except Exception as e:      # e.g. a band point too close to the boundary: reported, not fatal for the bench line
                nodes = {"error": str(e)[:200]}

        if world == 1:
            # FAST (as timed) vs EXACT on the full grid: the same sign field, the same 8 sweeps, a second grid
            G2 = DeviceGrid(nx, ny, nz)
            G2.fill(1.0)
            G2.signSearch(g["xLo"], DX, surfX, surfElem, g["box"])
            _lib.check(L.lsf_set_arith(_lib.ARITH_EXACT))
            rcx, nex, histx = G2.reinit(SWEEPS_PER_STEP - 1, DX, h, tol=0.0)
            x_ms, _ = _lib.last_timing()
            _lib.check(L.lsf_set_arith({"exact": _lib.ARITH_EXACT, "fast": _lib.ARITH_FAST, "auto": _lib.ARITH_AUTO}[args.arith]))
            G.fill(1.0)                                               # G: back to the state after the 8 FAST sweeps
            G.signSearch(g["xLo"], DX, surfX, surfElem, g["box"])
            rcf, nef, histf = G.reinit(SWEEPS_PER_STEP - 1, DX, h, tol=0.0)
            npts_all = (nx + 1) * (ny + 1) * (nz + 1)
            parity = {"full_grid": {"what": f"{SWEEPS_PER_STEP} sweeps from the sign field on the {n}^3 bench grid: arith '{s_arith}' (as timed) vs "
                                            "EXACT (bit-identical to the reference at every size tested against it)",
                                    "max_abs_fast_vs_exact": dev_max_abs_diff(G.device_ptr(), G2.device_ptr(), npts_all),
                                    "n_exit_equal": bool(nef == nex), "tolerance": 1e-10,
                                    "rms_hist_max_rel_diff": float(np.max(np.abs(histf - histx) / np.abs(histx))),
                                    "exact_ms_per_sweep": x_ms / SWEEPS_PER_STEP}}
            G2.close()
        if world > 1:
            SG.close()

    # ---- parity (b): the CPU sample slab swept by the GPU as a grid of its own, against the reference's result ----
    cpu = None
    if rank == 0 and not args.no_cpu:
        v, desc, kind, slab_in, slab_ref = cpu_sample(n, args.ref_slab, 1, keep=(not f32))
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": kind, "sample": desc}
        if not f32:
            Gs = DeviceGrid(slab_in.shape[0] - 1, slab_in.shape[1] - 1, slab_in.shape[2] - 1)
            _, hs = slab_field(n, args.ref_slab)
            out = {}
            for name, mode in (("exact", _lib.ARITH_EXACT), ("fast", _lib.ARITH_FAST)):
                _lib.check(L.lsf_set_arith(mode))
                Gs.upload(slab_in)
                Gs.reinit(0, DX, hs, tol=0.0)
                out[name] = float(np.abs(Gs.download() - slab_ref).max())
            _lib.check(L.lsf_set_arith({"exact": _lib.ARITH_EXACT, "fast": _lib.ARITH_FAST, "auto": _lib.ARITH_AUTO}[args.arith]))
            Gs.close()
            parity = parity or {}
            parity["cpu_sample_slab"] = {"what": f"one sweep (+BC) of the {slab_in.shape[0]}x{slab_in.shape[1]}x{slab_in.shape[2]} slab the "
                                                 f"cpu_baseline leg timed, GPU vs that leg's own result (kind '{kind}')",
                                         "max_abs_exact_vs_reference": out["exact"], "max_abs_fast_vs_reference": out["fast"],
                                         "bit_identical_exact": out["exact"] == 0.0, "tolerance": 1e-10}
        del slab_in, slab_ref

    # ---- companion: BASELINE config 3 -- 20k-triangle sphere on 512^3, reinit-only (single GPU) ----
    config3 = None
    if world == 1 and not f32 and not args.no_config3:
        tr3 = stl.sphere_config(512, DX)
        X3, E3 = stl.dedup_nodes(tr3)
        g3 = stl.grid_from_surface(X3, DX)
        G3 = DeviceGrid(g3["nx"], g3["ny"], g3["nz"])
        G3.fill(1.0)
        G3.signSearch(g3["xLo"], DX, X3, E3, g3["box"])
        sign3_ms, _ = _lib.last_timing()
        h3 = 0.1 * g3["dxx"]
        for _ in range(3):
            G3.reinit(SWEEPS_PER_STEP - 1, DX, h3, tol=0.0)
        barrier()
        ms3 = 0.0
        k3 = 5
        for _ in range(k3):
            rc3, ne3, hist3 = G3.reinit(SWEEPS_PER_STEP - 1, DX, h3, tol=0.0)
            ms, _nl = _lib.last_timing()
            ms3 += ms
        G3.close()
        c3 = (g3["nx"] - 1) * (g3["ny"] - 1) * (g3["nz"] - 1) * SWEEPS_PER_STEP
        config3 = {"workload": f"BASELINE config 3: synthetic sphere STL ({len(E3)} triangles) on a {g3['nx'] + 1}x{g3['ny'] + 1}x{g3['nz'] + 1} fp64 grid, "
                               "reinit-only (min/max iterations = 0), 1 B200", "value": c3 * k3 / (ms3 * 1e-3) / 1e9, "unit": UNIT,
                   "ms_per_step": ms3 / k3, "steps": k3, "sign_search_ms": sign3_ms, "last_rms": float(hist3[-1]),
                   "roofline_frac": 24.0 * c3 * k3 / (ms3 * 1e-3) / 1e9 / measured_peak()[0]}

    # ---- companion: the optional fp32 mode on the same geometry (single GPU; not part of `value`) ----
    fp32 = None
    if world == 1 and not f32 and not args.no_f32:
        G32 = DeviceGrid(nx, ny, nz, f32=True)
        G32.fill(1.0)
        G32.signSearch(g["xLo"], DX, surfX, surfElem, g["box"])
        for _ in range(2):
            G32.reinit(SWEEPS_PER_STEP - 1, DX, h, tol=0.0)
        barrier()
        ms32 = sw32 = 0.0
        ns32 = 0
        k32 = max(1, min(args.steps, 3))
        for _ in range(k32):
            rc, n_exit, hist32 = G32.reinit(SWEEPS_PER_STEP - 1, DX, h, tol=0.0)
            assert rc == 0 and n_exit == SWEEPS_PER_STEP - 1
            ms, _nl = _lib.last_timing()
            sm, ns = _lib.last_sweep_timing()
            ms32 += ms; sw32 += sm; ns32 += ns
        G32.close()
        l32 = sw32 / max(ns32, 1)
        a32 = 12.0 * (nx - 1) * (ny - 1) * (nz - 1) / (l32 * 1e-3) / 1e9 if l32 > 0 else None
        fp32 = {"metric": METRIC + " (fp32 mode)", "value": cells_per_step * k32 / (ms32 * 1e-3) / 1e9, "unit": UNIT, "dtype": "f32",
                "steps": k32, "ms_per_step": ms32 / k32,
                "roofline": {"bound": "hbm", "bytes_per_update": 12.0, "achieved": a32, "peak": measured_peak()[0], "unit": "GB/s",
                             "frac": a32 / measured_peak()[0] if a32 else None, "launch_ms": l32, "kernel": "k_reinit_march_f32",
                             "traffic": ncu_traffic(n, True)},
                "last_rms": float(hist32[-1]), "last_rms_fp64": float(hist[-1]) if hist is not None else None}

    # ---- companion: K2' throughput mode -- Jacobi WENO5 + TVD-RK3, the north_star's literal scheme, NOT the reference's algorithm
    # (lsf_grid_reinit_rk3; no reference parity is claimed for it, see DESIGN.md section 4 K2').  Same grid and geometry, single GPU.
    rk3 = None
    if world == 1 and not f32 and not args.no_rk3:
        GR = DeviceGrid(nx, ny, nz)
        GR.fill(1.0)
        GR.signSearch(g["xLo"], DX, surfX, surfElem, g["box"])
        GR.reinitRK3(1, DX, h, tol=0.0)
        barrier()
        krk = 3
        rc, ne, hist_rk = GR.reinitRK3(krk, DX, h, tol=0.0)
        ms_rk, nl_rk = _lib.last_timing()
        GR.close()
        ncell = (nx - 1) * (ny - 1) * (nz - 1)
        gbs = (24.0 + 32.0 + 32.0) * ncell * krk / (ms_rk * 1e-3) / 1e9
        rk3 = {"metric": "Jacobi WENO5 + TVD-RK3 reinit, Gcell-stage-updates/s (NOT the reference's Gauss-Seidel scheme; separate mode)",
               "value": 3.0 * ncell * krk / (ms_rk * 1e-3) / 1e9, "unit": "Gcell-stage-updates/s", "rk_steps": krk, "ms_per_rk_step": ms_rk / krk,
               "launches": nl_rk, "last_rms": float(hist_rk[-1]),
               "roofline": {"bound": "hbm", "bytes_per_cell_and_step": 88.0, "achieved": gbs, "peak": measured_peak()[0], "unit": "GB/s",
                            "frac": gbs / measured_peak()[0], "note": "stage 1 reads u, phiS and writes: 24 B; stages 2, 3 also read phi: 32 B each; "
                            "fp64 WENO5 stays FP64-pipe bound in this mode as well"}}

    # ---- companion at N > 1: the fp32-mode variant of config 5 (BASELINE configs[4] "plus fp32-mode variant") ----
    if world > 1 and not f32 and not args.no_f32:
        G32 = ShardedGrid(nx, ny, nz, f32=True)
        G32.fill(1.0)
        G32.signSearch(g["xLo"], DX, surfX, surfElem, g["box"])
        for _ in range(2):
            G32.reinit(SWEEPS_PER_STEP - 1, DX, h, tol=0.0)
        barrier()
        ms32 = 0.0
        k32 = max(1, min(args.steps, 3))
        for _ in range(k32):
            rc, n_exit32, hist32 = G32.reinit(SWEEPS_PER_STEP - 1, DX, h, tol=0.0)
            assert rc == 0 and n_exit32 == SWEEPS_PER_STEP - 1
            ms, _nl = _lib.last_timing()
            ms32 += ms
        d32 = allreduce_digest(G32.checksum()) if args.minmax_iters > 0 else None
        G32.close()
        t32 = torch.tensor([ms32], dtype=torch.float64, device="cuda")
        dist.all_reduce(t32, op=dist.ReduceOp.MAX)
        ms32 = float(t32.item())
        fp32 = {"metric": METRIC + " (fp32 mode, config 5 weak scaling)", "value": cells_per_step * k32 / (ms32 * 1e-3) / 1e9, "unit": UNIT,
                "dtype": "f32", "n_gpus": world, "steps": k32, "ms_per_step": ms32 / k32, "last_rms": float(hist32[-1]),
                "digest": hexd(d32) if d32 else None}

    # ---- companion at N > 1: K2' (Jacobi / TVD-RK3; not the reference's scheme) on the sharded config-5 grid -- a plain halo
    # problem (3-plane exchange after every stage, no pipeline), the mode whose weak scaling is NOT bound by the Gauss-Seidel order
    if world > 1 and not f32 and not args.no_rk3:
        GR = ShardedGrid(nx, ny, nz)
        GR.fill(1.0)
        GR.signSearch(g["xLo"], DX, surfX, surfElem, g["box"])
        GR.reinitRK3(1, DX, h, tol=0.0)
        barrier()
        krk = 3
        rc, ne, hist_rk = GR.reinitRK3(krk, DX, h, tol=0.0)
        ms_rk, nl_rk = _lib.last_timing()
        GR.close()
        trk = torch.tensor([ms_rk], dtype=torch.float64, device="cuda")
        dist.all_reduce(trk, op=dist.ReduceOp.MAX)
        ms_rk = float(trk.item())
        ncell = (nx - 1) * (ny - 1) * (nz - 1)
        rk3 = {"metric": "Jacobi WENO5 + TVD-RK3 reinit, Gcell-stage-updates/s (NOT the reference's Gauss-Seidel scheme; separate mode), config 5 weak scaling",
               "value": 3.0 * ncell * krk / (ms_rk * 1e-3) / 1e9, "unit": "Gcell-stage-updates/s", "n_gpus": world, "rk_steps": krk,
               "ms_per_rk_step": ms_rk / krk, "launches": nl_rk, "last_rms": float(hist_rk[-1])}

    # ---- e2e: the host-buffer drop-in call, pinned host memory, H2D + compute + D2H timed ------
    e2e = None
    if not args.no_e2e:
        npts = (nx + 1) * (ny + 1) * ((G.k1 - G.k0) if world > 1 else (nz + 1))     # this rank's owned points
        host = torch.empty(npts, dtype=torch.float64, pin_memory=True)
        G.download_ptr(host.data_ptr())                        # current phi as the e2e input
        hist_buf = np.zeros(SWEEPS_PER_STEP)
        import ctypes as C
        n_exit = C.c_int(0)
        if world == 1:
            if f32:
                _lib.check(L.lsf_set_precision(_lib.PREC_F32))

            def e2e_step():
                rc = L.lsf_reinit(C.cast(host.data_ptr(), _lib.c_double_p), None, None, nx, ny, nz, SWEEPS_PER_STEP - 1,
                                  DX, h, C.byref(n_exit), hist_buf.ctypes.data_as(_lib.c_double_p))
                _lib.check(rc)
            G.close()                                          # free the resident grid: lsf_reinit allocates its own
        else:
            def e2e_step():                                    # sharded: every rank moves its own slab (pinned host <-> its GPU)
                G.upload_ptr(host.data_ptr())
                step()
                G.download_ptr(host.data_ptr())
        e2e_step()                                             # warm-up
        barrier()
        t0 = time.perf_counter()
        k_e2e = max(1, min(args.steps, 3))
        for _ in range(k_e2e):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / k_e2e
        e2e = {"e2e_s": e2e_s, "bytes": npts * 8}
        # the same call as the reference's Fortran driver would make it through fortran/lsf_b200_mod.f90: ALLOCATEd
        # (pageable) arrays; (i) reinit_nograd_b200 -- phi only; (ii) reinit_b200 with the reference's full argument
        # list -- gradPhi(0:nx,0:ny,0:nz,3) and gradPhiMag shipped both ways (dead downstream, set3d.f90:372-375)
        if world == 1 and not f32 and not args.no_e2e_pageable:
            try:
                pg = np.empty(npts)                                 # pageable
                pg[:] = host.numpy()
                t0 = time.perf_counter()
                _lib.check(L.lsf_reinit(pg.ctypes.data_as(_lib.c_double_p), None, None, nx, ny, nz, SWEEPS_PER_STEP - 1, DX, h,
                                        C.byref(n_exit), hist_buf.ctypes.data_as(_lib.c_double_p)))
                e2e["pageable_nograd_s"] = time.perf_counter() - t0
                gp = np.zeros(3 * npts)
                gm = np.zeros(npts)
                pg[:] = host.numpy()
                t0 = time.perf_counter()
                _lib.check(L.lsf_reinit(pg.ctypes.data_as(_lib.c_double_p), gp.ctypes.data_as(_lib.c_double_p),
                                        gm.ctypes.data_as(_lib.c_double_p), nx, ny, nz, SWEEPS_PER_STEP - 1, DX, h,
                                        C.byref(n_exit), hist_buf.ctypes.data_as(_lib.c_double_p)))
                e2e["pageable_grad_s"] = time.perf_counter() - t0
                del pg, gp, gm
            except (MemoryError, _lib.LsfError) as ex:
                e2e["pageable_error"] = str(ex)[:200]
        if world > 1:
            G.close()
    else:
        G.close()

    # ---- reduce over ranks: MAX time -----------------------------------------------------------
    t = torch.tensor([dev_ms, wall_ms, e2e["e2e_s"] if e2e else 0.0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max, e2e_s_max = (float(v) for v in t.cpu())

    if rank == 0:
        value = cells_per_step * args.steps / (dev_ms_max * 1e-3) / 1e9
        peak, peak_src = measured_peak()
        launch_ms = sweep_ms / max(n_sweeps, 1)
        cells_per_launch = (nx - 1) * (ny - 1) * k_upd          # rank 0's sweep kernel
        achieved = bytes_per_update * cells_per_launch / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else None
        cfg = workload_config(n, world, f32)
        cfg.update({"triangles": int(len(surfElem)), "h": h,
                    "overlapped_sweeps": bool(args.overlap), "arith": args.arith,
                    "arith_used": main_arith, "sched": args.sched,
                    "parallelism": "single GPU" if world == 1 else
                    f"{world} z-slabs, Gauss-Seidel pipeline along k: streaming halo + ghost-plane exchange + RMS reduction "
                    "as peer stores over NVLink from inside the kernels (bit-identical to 1 GPU: see strong.digest_*)",
                    "l2": "inputs larger than L2 (%.1f GB per field)" % (8e-9 * n ** 3),
                    "wall_ms_per_step": wall_ms_max / args.steps, "sign_search_ms": sign_ms, "setup_s": setup_s,
                    "last_rms": float(hist[-1]) if hist is not None else None})
        e2e_line = None
        if e2e:
            e2e_line = {"value": cells_per_step / e2e_s_max / 1e9, "unit": UNIT,
                        "h2d_bytes_per_step": e2e["bytes"] * world, "d2h_bytes_per_step": e2e["bytes"] * world,
                        "api": "lsf_reinit(phi, NULL, NULL, ...) from page-locked host memory (lsf_host_register / reinit_nograd_b200)"
                               if world == 1 else "lsf_grid_upload + lsf_grid_reinit + lsf_grid_download on each rank's slab"}
            if "pageable_nograd_s" in e2e:
                e2e_line["as_fortran_driver"] = {
                    "pageable_phi_only": {"value": cells_per_step / e2e["pageable_nograd_s"] / 1e9, "s": e2e["pageable_nograd_s"],
                                          "api": "reinit_nograd_b200: ALLOCATEd (pageable) phi, no gradPhi"},
                    "pageable_full_argument_list": {"value": cells_per_step / e2e["pageable_grad_s"] / 1e9, "s": e2e["pageable_grad_s"],
                                                    "h2d_bytes": e2e["bytes"] * 5, "d2h_bytes": e2e["bytes"] * 5,
                                                    "api": "reinit_b200(phi,gradPhi,gradPhiMag,...): the reference's argument list, pageable "
                                                           "arrays, gradPhi/gradPhiMag shipped both ways + last sweep replayed"}}
            if "pageable_error" in e2e:
                e2e_line["as_fortran_driver"] = {"error": e2e["pageable_error"]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if f32 else "f64", "data": "synthetic",
                "config": cfg,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak if achieved else None, "traffic": ncu_traffic(n, f32),
                             "traffic_note": "DRAM read+write bytes of one launch of a whole %d^3 grid, ncu --set full (profiles/%s_%d_full.txt); "
                                             "algorithmic bytes per launch: %.4g" % (n, NCU_TAG[bool(f32)], n,
                                                                                    bytes_per_update * cells_per_launch),
                             "kernel": ("k_reinit_march_f32" if f32 else "k_reinit_march") if args.sched == "march" else "k_reinit_plane",
                             "launch_ms": launch_ms, "peak_source": peak_src,
                             "fp64_pipe_frac": (FP64_PER_UPDATE * cells_per_launch / (launch_ms * 1e-3)) / FP64_PIPE_PEAK if launch_ms > 0 and not f32 else None,
                             "fp64_pipe_note": "second roof: %d FP64-pipe instructions per cell update (SASS) vs %.1f T lane-ops/s "
                                               "(148 SM x 64 lanes x 1.965 GHz; micro-benchmarked 18.4 T, profiles/r1_fp64_pipe_ubench.txt)"
                                               % (FP64_PER_UPDATE, FP64_PIPE_PEAK / 1e12),
                             "note": "fp64 WENO5 is FP64-pipe bound (SURVEY.md fact 4); see DESIGN.md"},
                "cpu_baseline": cpu,
                "e2e": e2e_line,
                "parity": parity,
                "strong": strong,
                "config3": config3,
                "minmax_flow": mm,
                "fp32_mode": fp32,
                "rk3_mode": rk3,
                "node_projection": nodes,
                "sign_search": {"ms": sign_ms, "points": int(np.prod([g["box"][1] - g["box"][0] + 1, g["box"][3] - g["box"][2] + 1,
                                                                      g["box"][5] - g["box"][4] + 1])),
                                "triangles": int(len(surfElem))},
                "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=1024, help="grid points per axis per GPU")
    ap.add_argument("--arith", default="auto", choices=["auto", "fast", "exact"])
    ap.add_argument("--sched", default="march", choices=["march", "plane"])
    ap.add_argument("--ref-slab", type=int, default=16, help="z thickness of the CPU sample slab (16: ~10 s of serial CPU work per sweep "
                                                             "of the machine-translated reference)")
    ap.add_argument("--no-config3", action="store_true", help="skip the BASELINE config 3 companion (sphere, 512^3)")
    ap.add_argument("--no-e2e-pageable", action="store_true", help="skip the pageable-memory / full-argument-list e2e variants")
    ap.add_argument("--minmax-iters", type=int, default=64, help="min/max iterations of the companion measurement (0 = skip)")
    ap.add_argument("--overlap", action="store_true", help="run the sweeps in overlapped batches (lsf_set_overlap; opt-in)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--f32", action="store_true", help="measure the optional fp32 mode as the main line (single GPU)")
    ap.add_argument("--no-f32", action="store_true", help="skip the fp32-mode companion measurement")
    ap.add_argument("--no-rk3", action="store_true", help="skip the K2' (Jacobi / TVD-RK3) companion measurement")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
