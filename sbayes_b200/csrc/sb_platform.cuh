// sb_platform.cuh -- the only place the step kernel touches the machine.
//
// The kernels are written against a small SPMD vocabulary (thread / CTA ids, block barrier,
// 32-wide warp shuffles, global atomics, acquire loads, dynamic shared memory).  In the
// product build (nvcc, sm_100a) every wrapper is the CUDA intrinsic.  With -DSB_HOSTSIM
// (test builds only, see tests/hostsim/hostsim.h) the same wrappers bind to a fibre
// emulator so the control flow can be checked without a GPU.
#pragma once
#ifndef SB_BLOCK
#define SB_BLOCK 512
#endif
//   /* threads per CTA of the step kernel (product build) */
#include <stdint.h>
#include <math.h>

#ifdef SB_HOSTSIM
#include "hostsim.h"
#define SB_DEV inline
#define SB_HD inline
#define SB_KERNEL static void
#define SB_KERNEL_BIG static void
#define SB_KERNEL_LB(threads, blocks) static void
#define SB_GRID_CONSTANT
#define SB_RESTRICT
SB_DEV int sb_tid() { return hostsim::tid(); }
SB_DEV int sb_nthr() { return hostsim::nthr(); }
SB_DEV int sb_cta() { return hostsim::cta(); }
SB_DEV int sb_ncta() { return hostsim::ncta(); }
SB_DEV void sb_cta_sync() { hostsim::cta_sync(); }
SB_DEV void sb_warp_sync() { hostsim::warp_sync(); }
SB_DEV char *sb_smem() { return hostsim::smem(); }
SB_DEV double sb_shfl_f64(double v, int src) {
    uint64_t u; memcpy(&u, &v, 8); u = hostsim::shfl(u, src); memcpy(&v, &u, 8); return v;
}
SB_DEV int sb_shfl_i32(int v, int src) { return (int)(int64_t)hostsim::shfl((uint64_t)(int64_t)v, src); }
SB_DEV unsigned sb_ballot(bool p) { return hostsim::ballot(p); }
SB_DEV int sb_atomic_add_i32(int *p, int v) { int o = *p; *p = o + v; return o; }
SB_DEV unsigned long long sb_atomic_add_u64(unsigned long long *p, unsigned long long v) {
    unsigned long long o = *p; *p = o + v; return o;
}
SB_DEV unsigned long long sb_ld_acquire_u64(const unsigned long long *p) { return *(volatile const unsigned long long *)p; }
SB_DEV void sb_threadfence() {}
SB_DEV void sb_threadfence_system() {}
SB_DEV unsigned long long sb_ld_acquire_sys_u64(const unsigned long long *p) { return *(volatile const unsigned long long *)p; }
SB_DEV void sb_st_release_sys_u64(unsigned long long *p, unsigned long long v) { *(volatile unsigned long long *)p = v; }
SB_DEV int sb_ld_volatile_i32(const int *p) { return *(volatile const int *)p; }
SB_DEV unsigned long long sb_ld_volatile_u64(const unsigned long long *p) { return *(volatile const unsigned long long *)p; }
SB_DEV void sb_st_volatile_u64(unsigned long long *p, unsigned long long v) { *(volatile unsigned long long *)p = v; }
SB_DEV void sb_spin() { hostsim::yield(); }
SB_DEV long long sb_clock() { return 0; }
SB_DEV uint32_t sb_mulhi_u32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
SB_DEV int sb_popc(unsigned v) { return __builtin_popcount(v); }
SB_DEV int sb_ffs(unsigned v) { return __builtin_ffs((int)v); }
SB_DEV int sb_clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
template <class T> SB_DEV T sb_ldg(const T *p) { return *p; }
SB_DEV long long sb_d2ll(double v) { long long u; memcpy(&u, &v, 8); return u; }
SB_DEV double sb_ll2d(long long u) { double v; memcpy(&v, &u, 8); return v; }
#else
#include <cuda_runtime.h>
#define SB_DEV __device__ __forceinline__
#define SB_HD __host__ __device__ __forceinline__
#define SB_KERNEL __global__ void
#define SB_KERNEL_BIG __global__ void __launch_bounds__(SB_BLOCK, 1)
#define SB_KERNEL_LB(threads, blocks) __global__ void __launch_bounds__(threads, blocks)
#define SB_GRID_CONSTANT __grid_constant__
#define SB_RESTRICT __restrict__
SB_DEV int sb_tid() { return threadIdx.x; }
SB_DEV int sb_nthr() { return blockDim.x; }
SB_DEV int sb_cta() { return blockIdx.x; }
SB_DEV int sb_ncta() { return gridDim.x; }
SB_DEV void sb_cta_sync() { __syncthreads(); }
SB_DEV void sb_warp_sync() { __syncwarp(); }
extern __shared__ __align__(16) char sb_dyn_smem[];
SB_DEV char *sb_smem() { return sb_dyn_smem; }
SB_DEV double sb_shfl_f64(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
SB_DEV int sb_shfl_i32(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
SB_DEV unsigned sb_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
SB_DEV int sb_atomic_add_i32(int *p, int v) { return atomicAdd(p, v); }
SB_DEV unsigned long long sb_atomic_add_u64(unsigned long long *p, unsigned long long v) { return atomicAdd(p, v); }
SB_DEV unsigned long long sb_ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
SB_DEV void sb_threadfence() { __threadfence(); }
SB_DEV void sb_threadfence_system() { __threadfence_system(); }
SB_DEV unsigned long long sb_ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
SB_DEV void sb_st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
SB_DEV int sb_ld_volatile_i32(const int *p) { return *(volatile const int *)p; }
SB_DEV unsigned long long sb_ld_volatile_u64(const unsigned long long *p) { return *(volatile const unsigned long long *)p; }
SB_DEV void sb_st_volatile_u64(unsigned long long *p, unsigned long long v) { *(volatile unsigned long long *)p = v; }
SB_DEV void sb_spin() {}
SB_DEV long long sb_clock() { return clock64(); }
SB_DEV uint32_t sb_mulhi_u32(uint32_t a, uint32_t b) { return __umulhi(a, b); }
SB_DEV int sb_popc(unsigned v) { return __popc(v); }
SB_DEV int sb_ffs(unsigned v) { return __ffs((int)v); }
SB_DEV int sb_clz(unsigned v) { return __clz((int)v); }
template <class T> SB_DEV T sb_ldg(const T *p) { return __ldg(p); }
SB_DEV long long sb_d2ll(double v) { return __double_as_longlong(v); }
SB_DEV double sb_ll2d(long long u) { return __longlong_as_double(u); }
#endif

SB_DEV int sb_lane() { return sb_tid() & 31; }
SB_DEV int sb_warp() { return sb_tid() >> 5; }
SB_DEV int sb_nwarps() { return sb_nthr() >> 5; }
