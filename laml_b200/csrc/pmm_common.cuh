// pmm_common.cuh -- scalar helpers and small types shared by the full-state kernels (pmm_kernels.cuh, pmm_nni.cuh)
// and the lumped-state kernels (pmm_lumped.cuh, pmm_nni_lumped.cuh).  Nothing here needs more of CUDA than a few
// intrinsics, so that the lumped NNI kernel can also be compiled for the host under the test shim
// (tests/native/cuda_host_shim.h): these definitions are then the ones the emulated kernel runs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pmm {

// ---------------------------------------------------------------- exponent handling
__device__ __forceinline__ unsigned hi32(double x) { return (unsigned)__double2hiint(x); }
// exponent e such that max * 2^-e is in [1,2); 0 for zero / denormal maxima (dead site)
__device__ __forceinline__ int expo_of_hi(unsigned hi) {
    int be = (int)((hi >> 20) & 0x7ffu);
    return be ? be - 1023 : 0;
}
__device__ __forceinline__ double pow2i(int e) {          // exact 2^e for e in [-1022,1023]
    return __hiloint2double((1023 + e) << 20, 0);
}

// ---------------------------------------------------------------- per-branch M-step (EM_solver.py:351-357)
__device__ __forceinline__ double dobj_dd(double d, double nu, bool use_nu, double s0, double s1, double s24, double s3) {
    double g = -(nu + 1.0) * s0 - nu * (s1 + s3) + s1 / expm1(d);
    if (use_nu) g += nu * s24 / expm1(nu * d);
    return g;
}

// 1 / expm1 at the two bounds (the same for every branch): e[0] = 1/expm1(dmin), e[1] = 1/expm1(dmax),
// e[2] = nu/expm1(nu dmin), e[3] = nu/expm1(nu dmax) (0 when nu is switched off)
__device__ __forceinline__ double mstep_end_const(int i, double nu, bool use_nu, double dmin, double dmax) {
    const double t = (i & 1) ? dmax : dmin;
    return (i < 2) ? 1.0 / expm1(t) : (use_nu ? nu / expm1(nu * t) : 0.0);
}
// argmax over [dmin, dmax] of the per-branch program EM_solver.py:351-357 for normalised statistics s[0..4].
// f' is convex and strictly decreasing, so Newton from the left of the root increases monotonically towards
// it.  Start from the current length (EM moves it little) or, if that is a bound, from the root of the lower
// bound 1/expm1(t) >= 1/t - 1/2, which lies left of the true root.  The bracket is kept and bisection is
// the fallback; the two bound tests need no transcendental.
__device__ __forceinline__ double mstep_branch_length(const double *s, double nu, bool use_nu, double dmin, double dmax, double d_cur,
                                                      const double *__restrict__ e) {
    const double s24 = s[2] + s[4];
    const double lin = (nu + 1.0) * s[0] + nu * (s[1] + s[3]);
    double lo = dmin, hi = dmax;
    if (-lin + s[1] * e[0] + s24 * e[2] <= 0.0) return lo;            // f'(dmin) <= 0
    if (-lin + s[1] * e[1] + s24 * e[3] >= 0.0) return hi;            // f'(dmax) >= 0
    const double w24 = use_nu ? s24 : 0.0;
    double x = d_cur;
    if (!(x > lo && x < hi)) x = fmin(fmax((s[1] + w24) / (lin + 0.5 * s[1] + 0.5 * nu * w24), lo), hi);
    for (int it = 0; it < 100; it++) {
        const double em = expm1(x), r = 1.0 / em;
        double g = -lin + s[1] * r, gp = -s[1] * r * r * (em + 1.0);
        if (use_nu) {
            const double en = expm1(nu * x), rn = nu / en;
            g += s24 * rn;
            gp -= s24 * rn * rn * (en + 1.0);
        }
        if (g > 0.0) lo = x; else if (g < 0.0) hi = x; else break;
        double xn = x - g / gp;
        // converged: test the Newton step itself first -- a monotone approach from the left never moves `hi`, and
        // a step that rounds to x must end the iteration, not send it to the middle of a stale bracket
        if (fabs(xn - x) <= 1e-15 * x) { x = fmin(fmax(xn, lo), hi); break; }
        if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
        if (xn <= lo || xn >= hi) { x = xn; break; }
        x = xn;
    }
    return fmin(fmax(x, dmin), dmax);
}

// ---------------------------------------------------------------- NNI candidates (pmm_nni.cuh, pmm_nni_lumped.cuh)
struct NniCand { int g, s, A, B, C, v, u, root; };    // node ids (g = -1: v is a root), root index of the tree

// per-candidate state: d5 = lengths of (v, u, C, A, B)
struct NniState { double d[5]; double llh_prev, llh_cur; int iters, done; };

__device__ __forceinline__ double4 consts_of(double dv, double nu) {
    const double c = exp(-nu * dv);
    return make_double4(fabs(exp(-(1.0 + nu) * dv)), fabs(c * (-expm1(-dv))), fabs(c), fabs(-expm1(-nu * dv)));
}

}  // namespace pmm
