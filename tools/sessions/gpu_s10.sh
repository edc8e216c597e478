#!/bin/bash
# GPU session 10: BVH sign search (parity + timing), min/max timing after the scratch fix
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_parity.py tests/test_gpu_f32.py tests/test_gpu_nodes.py -x -q > gpurun_out/s10_tests.txt 2>&1; tail -3 gpurun_out/s10_tests.txt
python tools/mm_probe.py 1024 > gpurun_out/s10_mm_probe.txt 2>&1; cat gpurun_out/s10_mm_probe.txt
python - <<'PY'
import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np
from levelsetfortran_b200 import DeviceGrid, _lib, stl
DX = 0.05
L = _lib.lib(); _lib.check(L.lsf_init(0))
def run(tag, tris, env=None):
    X, E = stl.dedup_nodes(tris)
    g = stl.grid_from_surface(X, DX)
    G = DeviceGrid(g["nx"], g["ny"], g["nz"])
    out = {}
    for mode in ("bvh", "tiled"):
        if mode == "tiled": os.environ["LSF_SIGN_TILED"] = "1"
        else: os.environ.pop("LSF_SIGN_TILED", None)
        for rep in range(2):
            G.fill(1.0)
            t0 = time.perf_counter(); G.signSearch(g["xLo"], DX, X, E, g["box"]); wall = time.perf_counter() - t0
            ms, _ = _lib.last_timing()
        out[mode] = (ms, wall * 1e3, G.checksum())
    G.close()
    print("%s: %d triangles, grid %d^3-class: bvh %.1f ms (wall %.1f), tiled %.1f ms (wall %.1f), identical field: %s" % (
        tag, len(E), g["nx"] + 1, out["bvh"][0], out["bvh"][1], out["tiled"][0], out["tiled"][1], out["bvh"][2] == out["tiled"][2]), flush=True)
run("torus+cube 1024", stl.torus_cube_config((1024, 1024, 1024), DX))
run("sphere 512 (config 3)", stl.sphere_config(512, DX))
run("sphere 512, 80k triangles", stl.sphere_config(512, DX, n_lat=201, n_lon=200))
PY
