// lsf_slab.cuh -- z-slab sharding of phi(0:nx,0:ny,0:NZ) over the GPUs of one node.
//
// k is the slowest index of the reference layout, so the planes k0..k1-1 a rank owns are one contiguous
// range of the global array.  A rank stores its owned planes plus SLAB_GHOST = 3 ghost planes (the WENO5
// stencil half-width, subs.f90:509-552) towards each neighbour, in the same dense layout, so every kernel
// of the single-GPU path runs on the local array unchanged apart from the plane ranges it is given.
//
// The reference sweep is an in-place Gauss-Seidel (subs.f90:742-852): a slab needs the NEW values of the
// planes below it (in sweep direction) and the OLD values of the planes above it.  Hence
//   * between two sweeps every rank copies its 3 outermost owned planes into the neighbours' ghost planes
//     (k_slab_exchange: peer stores over NVLink + a system-scope flag) -- the OLD-value snapshot;
//   * during a sweep the upstream rank's last tile row streams its NEW values into the downstream rank's
//     ghost planes and publishes its progress there (lsf_march.cuh: push_delta / push_progress /
//     in_progress): the ranks form a software pipeline along k, with no host involvement and no
//     collective on the data path;
//   * the RMS exit test (subs.f90:902-918) is a sum over all ranks: every rank writes its partial sum into
//     every peer's SlabSync block and sums the P values in rank order (k_finalize), so all ranks take the
//     identical EXIT / NaN decision on the device.
// This header is host-only geometry + the peer-visible sync block; it is shared with tests/emu.
#pragma once
#include <stdint.h>

namespace lsf {

constexpr int SLAB_GHOST = 3;
constexpr int SLAB_MAX_RANKS = 16;
constexpr int SLAB_MIN_PLANES = 8;      // owned planes per rank (>= 2*ghost, keeps every exchange nearest-neighbour)
constexpr int SLAB_MAX_NTB = 4096;
constexpr int SLAB_BOX_BYTES = 128;    // host-side mailbox per source rank (SlabSync::box)
constexpr int SLAB_SUM_SLOTS = 16;      // reductions may be published this many sequence numbers ahead of their consumption

struct SlabGeom {
    int rank, nranks;
    int NZ;                 // global grid: planes 0..NZ
    int k0, k1;             // owned global planes [k0, k1)
    int g_lo, g_hi;         // ghost planes towards the low / high neighbour (0 or SLAB_GHOST)
    int kbase;              // global plane of local plane 0
    int nzl;                // local planes are 0..nzl
    int kupd_lo, kupd_hi;   // LOCAL plane range the sweeps update: owned planes minus the global faces k = 0, NZ
    int own_lo, own_hi;     // LOCAL plane range owned (inclusive)
};

// Balanced contiguous partition of the NZ+1 planes.  Returns false if a rank would own < SLAB_MIN_PLANES.
inline bool slab_geom(int NZ, int nranks, int rank, SlabGeom &s)
{
    if (nranks < 1 || nranks > SLAB_MAX_RANKS || rank < 0 || rank >= nranks) return false;
    const long long np = (long long)NZ + 1;
    if (np / nranks < SLAB_MIN_PLANES && nranks > 1) return false;
    s.rank = rank; s.nranks = nranks; s.NZ = NZ;
    s.k0 = (int)(np * rank / nranks);
    s.k1 = (int)(np * (rank + 1) / nranks);
    s.g_lo = rank > 0 ? SLAB_GHOST : 0;
    s.g_hi = rank < nranks - 1 ? SLAB_GHOST : 0;
    s.kbase = s.k0 - s.g_lo;
    s.nzl = (s.k1 - s.k0) + s.g_lo + s.g_hi - 1;
    s.own_lo = s.g_lo;
    s.own_hi = s.nzl - s.g_hi;
    s.kupd_lo = (s.k0 == 0) ? 1 : s.own_lo;
    s.kupd_hi = (s.k1 == NZ + 1) ? s.own_hi - 1 : s.own_hi;
    return true;
}

// Peer-visible synchronisation block, one per rank, at the start of the rank's shared allocation.
// Every field is written by exactly one remote rank (or one side), so there are no atomics across NVLink.
struct SlabSync {
    long long halo_seq[2];                   // [0] bumped by the low neighbour, [1] by the high one: "my planes of phase e are in your ghosts"
    long long phase_done[2];                 // same sides: "I have finished compute phase e" (I no longer read my ghost planes)
    long long sum_seq[SLAB_SUM_SLOTS][SLAB_MAX_RANKS];   // per slot and source rank: sequence number of the contribution held there
    double rank_sum[SLAB_SUM_SLOTS][SLAB_MAX_RANKS];     // per-rank partial sums of the RMS test, slot = sequence number mod SLOTS
    int rank_flag[SLAB_SUM_SLOTS][SLAB_MAX_RANKS];       // flags travelling with the sums (bit 0 guard, bit 1 band-on-boundary, bit 2 timeout, bit 3 other error)
    long long in_progress[SLAB_MAX_NTB];     // streaming-halo progress of the upstream rank's last tile row
    long long edge_done[2][SLAB_MAX_NTB];    // [side][J]: epoch of the last sweep in which that neighbour's tile J adjacent to this rank completed
    // host-side mailbox (lsf_slab.cu: slab_host_exchange): rank r copies up to SLAB_BOX_BYTES into box[r] of every rank and then raises
    // seq[r]; used for the rendezvous of collective calls that need one (IPC handles of a transient shadow grid, open / close barriers)
    long long box_seq[SLAB_MAX_RANKS];
    unsigned char box[SLAB_MAX_RANKS][SLAB_BOX_BYTES];
};

// Read-only view of a field of the GLOBAL grid phi(0:nx,0:ny,0:NZ) whose planes are spread over the ranks' slabs: element q
// (global linear index) lives on the rank owning plane q / sxy, at the same (i,j) in that rank's local array.  Used by kernels
// that gather a few values anywhere in the grid (surface-node projection): loads from a peer's slab travel over NVLink.
struct SlabView {
    const double *base[SLAB_MAX_RANKS];      // the ranks' local arrays (peer-mapped)
    long long shift[SLAB_MAX_RANKS];         // kbase * sxy: global index of the first element of the local array
    int kend[SLAB_MAX_RANKS];                // first global plane NOT owned by the rank
    int nranks;
    long long sxy;
#if defined(__CUDACC__)
    __host__ __device__
#endif
    double operator[](long long q) const
    {
        const int k = (int)(q / sxy);
        int r = 0;
        while (r < nranks - 1 && k >= kend[r]) ++r;
        return base[r][q - shift[r]];
    }
};

}  // namespace lsf
