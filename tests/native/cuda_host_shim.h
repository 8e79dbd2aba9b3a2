// cuda_host_shim.h -- TEST INFRASTRUCTURE: just enough of the CUDA execution model to run the library's lumped
// kernels (laml_b200/csrc/pmm_lumped.cuh, pmm_nni_lumped.cuh) on the host, unchanged, under g++.
//
// A CTA is emulated by one std::thread per CUDA thread (CTAs run one after another): __shared__ variables become
// function-local statics (one copy, shared by the CTA's threads), __syncthreads() is a barrier over the CTA,
// __syncwarp() / __shfl_xor_sync() a barrier / an exchange over the 32 threads of a warp.  All lanes of a warp must
// reach a warp primitive and all threads of the CTA a __syncthreads() -- true of the kernels run here.
// This checks the kernels' LOGIC (index arithmetic, closed forms, reductions, loop and convergence structure) on a
// machine without a GPU; it says nothing about memory ordering, registers or speed.
#pragma once
#define PMM_HOST_EMU 1
#include <cuda_runtime.h>      // vector types (double2, double4, int2, int4, uint3, dim3) and make_* for the host
#include <math.h>
#include <pthread.h>
#include <string.h>
#include <algorithm>
#include <thread>
#include <vector>

#undef __global__
#undef __device__
#undef __host__
#undef __shared__
#undef __forceinline__
#undef __noinline__
#undef __launch_bounds__
#undef __align__
#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define PMM_HD inline

namespace emu {
constexpr int MAX_CTA = 1024;
struct Cta {
    pthread_barrier_t cta_bar;
    pthread_barrier_t warp_bar[MAX_CTA / 32];
    double xchg[MAX_CTA];
};
inline Cta *&current() { static Cta *c = nullptr; return c; }
}  // namespace emu

static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;

inline void __syncthreads() { pthread_barrier_wait(&emu::current()->cta_bar); }
inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&emu::current()->warp_bar[threadIdx.x >> 5]); }
inline double __shfl_xor_sync(unsigned, double x, int lane_mask) {
    emu::Cta *c = emu::current();
    c->xchg[threadIdx.x] = x;
    pthread_barrier_wait(&c->warp_bar[threadIdx.x >> 5]);
    const double r = c->xchg[threadIdx.x ^ lane_mask];
    pthread_barrier_wait(&c->warp_bar[threadIdx.x >> 5]);
    return r;
}
inline double __shfl_sync(unsigned, double x, int src_lane) {
    emu::Cta *c = emu::current();
    c->xchg[threadIdx.x] = x;
    pthread_barrier_wait(&c->warp_bar[threadIdx.x >> 5]);
    const double r = c->xchg[(threadIdx.x & ~31u) + (unsigned)src_lane];
    pthread_barrier_wait(&c->warp_bar[threadIdx.x >> 5]);
    return r;
}
inline int __shfl_sync(unsigned m, int x, int src_lane) { return (int)__shfl_sync(m, (double)x, src_lane); }
template <class T> inline T __ldg(const T *p) { return *p; }
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int __double2hiint(double x) { long long b; memcpy(&b, &x, 8); return (int)(b >> 32); }
inline int __double2loint(double x) { long long b; memcpy(&b, &x, 8); return (int)(b & 0xffffffffll); }
inline double __hiloint2double(int hi, int lo) {
    const unsigned long long b = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
    double x; memcpy(&x, &b, 8); return x;
}
using std::max;
using std::min;

namespace emu {
// kernel<<<dim3(grid, grid_y), block>>>(args...) with a 1-D block; the kernel is any callable
template <class K, class... A>
void launch2(unsigned grid, unsigned grid_y, unsigned block, K kernel, A... args) {
    Cta cta;
    current() = &cta;
    pthread_barrier_init(&cta.cta_bar, nullptr, block);
    for (unsigned w = 0; w < (block + 31) / 32; w++) pthread_barrier_init(&cta.warp_bar[w], nullptr, std::min(32u, block - 32 * w));
    for (unsigned by = 0; by < grid_y; by++)
        for (unsigned b = 0; b < grid; b++) {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < block; t++)
                th.emplace_back([=] {
                    threadIdx = {t, 0, 0}; blockIdx = {b, by, 0};
                    blockDim = dim3(block); gridDim = dim3(grid, grid_y);
                    kernel(args...);
                });
            for (auto &x : th) x.join();
        }
    pthread_barrier_destroy(&cta.cta_bar);
    for (unsigned w = 0; w < (block + 31) / 32; w++) pthread_barrier_destroy(&cta.warp_bar[w]);
    current() = nullptr;
}
template <class K, class... A>
void launch(unsigned grid, unsigned block, K kernel, A... args) { launch2(grid, 1u, block, kernel, args...); }
}  // namespace emu
