// lsf_common.cuh -- shared declarations of the CUDA implementation (internal; the public
// surface is include/lsf_b200.h).
#pragma once
#if !defined(LSF_EMU)
#include <cuda_runtime.h>
#else
#define __host__
#define __device__
#endif
#include <stdint.h>
#include <stdio.h>

#include "lsf_cell.cuh"

namespace lsf {

// Device-side control block of an iteration loop (reinit: subs.f90:735-928; min/max:
// set3d.f90:394-462).  Kernels of later iterations turn into no-ops once `done` is set, so the
// host can enqueue iterations in batches without a host sync per iteration and still stop at
// exactly the reference's iteration.
struct Ctrl {
    int done;     // 1 once the loop has left (EXIT or NaN STOP)
    int status;   // 0 = EXIT on tolerance / still running, 1 = NaN, <0 = LSF_ERR_*
    int n;        // index of the iteration being executed (reinit: 0-based; min/max: 1-based)
    int n_exit;   // n at which the loop left
    int guard;    // set by the FAST sweep kernels when a cell update was ill-conditioned (lsf_cell.cuh)
};

struct Dims {
    int nx, ny, nz;          // reference extents: points are 0..nx etc.
    long long sx, sxy;       // strides of j and k in elements (i is contiguous)
};

// raster 1..8 -> sweep direction per axis, subs.f90:742-852
__host__ __device__ inline void raster_dirs(int raster, int d[3])
{
    const int tab[8][3] = {{+1, +1, +1}, {+1, +1, -1}, {+1, -1, -1}, {-1, -1, -1},
                           {-1, +1, -1}, {-1, -1, +1}, {-1, +1, +1}, {+1, -1, +1}};
    d[0] = tab[raster - 1][0];
    d[1] = tab[raster - 1][1];
    d[2] = tab[raster - 1][2];
}

}  // namespace lsf
