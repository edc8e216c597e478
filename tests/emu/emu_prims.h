// emu_prims.h -- CPU stand-ins for the handful of device primitives lsf_march.cuh uses, so the
// very same tile/march code runs with one OS thread per CUDA thread (test infrastructure).
#pragma once
#include <pthread.h>
#include <sched.h>
#include <unistd.h>
#define LSF_DEV inline
namespace lsf {
struct EmuCta { pthread_barrier_t bar; };
extern thread_local EmuCta *emu_cta;
inline void p_sync() { pthread_barrier_wait(&emu_cta->bar); }
inline double p_ldcg(const double *p) { return *(const volatile double *)p; }
inline void p_stcg(double *p, double v) { *(volatile double *)p = v; }
inline float p_ldcg(const float *p) { return *(const volatile float *)p; }
inline double p_ldca(const double *p) { return *(const volatile double *)p; }
inline float p_ldca(const float *p) { return *(const volatile float *)p; }
inline void p_prefetch_l2(const void *) {}
template <class T> inline T p_lds(const T *p) { return *(const volatile T *)p; }   // ring gather (shared memory on the GPU)
inline void p_stcg(float *p, float v) { *(volatile float *)p = v; }
template <int VEC, class T> inline void p_ldcg_vec(const T *p, T *out) { for (int q = 0; q < VEC; ++q) out[q] = *(const volatile T *)(p + q); }
inline void p_fence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline unsigned p_ticket(unsigned *ctr) { return __atomic_fetch_add(ctr, 1u, __ATOMIC_SEQ_CST); }
inline long long p_ld_acquire(const long long *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline long long p_ld_relaxed(const long long *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline void p_fence_acquire() { __atomic_thread_fence(__ATOMIC_ACQ_REL); }
inline void p_st_release(long long *p, long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline void p_sleep() { sched_yield(); }
inline int p_ld_relaxed_i32(const int *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline bool p_cas_i32(int *p, int expect, int desired)
{
    return __atomic_compare_exchange_n(p, &expect, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
}
// test hook: stall a CTA close to the end of its tile, so that ranks / tiles that are allowed to run ahead
// really do (widens the hazard windows the flag protocol has to close)
extern int emu_stall_us;
inline void p_emu_hook(bool at) { if (at && emu_stall_us) usleep((useconds_t)emu_stall_us); }
inline long long p_ld_relaxed_sys(const long long *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline void p_fence_sys() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void p_st_release_sys(long long *p, long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline void p_st_peer(double *p, double v) { *(volatile double *)p = v; }
inline void p_st_peer(float *p, float v) { *(volatile float *)p = v; }
// split-phase barrier: one arrival per emulated thread, phases counted up (faithful: a waiter only waits for ARRIVALS);
// the word holds the thread count in its top 16 bits and the number of arrivals so far below
inline void p_bar_init(unsigned long long *bar, int nthreads) { __atomic_store_n(bar, (unsigned long long)nthreads << 48, __ATOMIC_SEQ_CST); }
inline void p_bar_arrive(unsigned long long *bar) { __atomic_fetch_add(bar, 1ull, __ATOMIC_RELEASE); }
inline void p_bar_wait(unsigned long long *bar, unsigned phase)
{
    for (;;) {
        const unsigned long long v = __atomic_load_n(bar, __ATOMIC_ACQUIRE);
        if ((v & 0xffffffffffffull) >= (v >> 48) * ((unsigned long long)phase + 1ull)) return;
        sched_yield();
    }
}
}  // namespace lsf
