// pmm_nni_ultra.cuh -- the local M-step of an NNI candidate under the ultrametric constraint.
//
// Reference behaviour: Topology_search.apply_nni with strategy['ultra_constr'] (the runner's default,
// runner.py:91,287) frees the five branches around the move (Topology_search.py:259) and every M-step of the
// candidate's EM solves EM_solver.py:351-357 subject to the box AND the root-to-leaf equalities
// (EM_solver.py:53-89,359-361; cvxpy -> MOSEK).  With everything else fixed the heights of the three subtrees
// below the move (hA, hB, hC) and of v's parent (hTop) are constants, so in the height parametrisation of
// pmm_ultrametric.cpp the candidate has TWO unknowns, the heights of u' and v':
//     d_v = hTop - Hv   d_u = Hv - Hu   d_C = Hv - hC   d_A = Hu - hA   d_B = Hu - hB        (order of NniState.d)
// (when v is the root of the only tree its branch hangs from nothing: d_v is a third, separable unknown).
// ultra2_solve is pmm_ultrametric_solve specialised to this five-edge tree, statement by statement: same
// interior point and margin, same warm start, same barrier schedule, step rule, line search and stopping
// tests -- so a candidate scored in the batch kernel walks the same EM trajectory as one scored by the
// host M-step on the whole tree.  Plain fp64 arithmetic, __host__ __device__: the host instance is exported as
// pmm_debug_nni_ultra_mstep and tested against pmm_ultrametric_solve on CPU (tests/test_nni_ultra_host.py).
// On the device the routine is called by a WHOLE WARP with the same arguments: the transcendental calls of a Newton
// step (two expm1 per edge) and of a line-search trial (log(-expm1(.)) twice and two barrier logarithms per edge) are
// spread over lanes -- lane 4 e + j evaluates term j of edge e, the four lanes of an edge put its value / derivatives
// together (so the divisions stay parallel over the edges too) -- and handed to every lane by shuffles; every lane then
// forms the sums in the serial order e = 0..4, i.e. the same arithmetic in the same order as the host instance, with
// two dependent transcendental calls per trial instead of thirty.  Measured on config 5 (9 996 candidates, kernel time
// under ncu): thread 0 alone ~21 ms, one lane per edge 7.5 ms, this form 6.1 ms; the whole call 22.9 -> 7.0 ms.
#pragma once
#include <math.h>

#ifndef PMM_HD
#ifdef __CUDACC__
#define PMM_HD __host__ __device__ __forceinline__
#else
#define PMM_HD inline
#endif
#endif

namespace pmm {

struct NniUltra { double hA, hB, hC, hTop; int top_free, pad; };

// per-edge objective pieces: k = (nu+1) S0 + nu (S1+S3), S1, S24 = S2 + S4
struct Ultra2Stats { double k[5], S1[5], S24[5]; double nu; bool use_nu; };

// value alone (expm1 + log per term) and derivatives alone (one expm1 per term): neither caller pays for the other
// (the transcendental terms are arguments: A = log(-expm1(-x)), B = log(-expm1(-nu x)), em = expm1(x), emn = expm1(nu x))
PMM_HD double ultra2_value_of(const Ultra2Stats &T, int e, double x, double A, double B) {
    double val = -T.k[e] * x + T.S1[e] * A;
    if (T.use_nu) val += T.S24[e] * B;
    return val;
}
PMM_HD void ultra2_grad_hess_of(const Ultra2Stats &T, int e, double em, double emn, double &g, double &h) {
    g = -T.k[e] + T.S1[e] / em;
    h = -T.S1[e] * (em + 1.0) / (em * em);
    if (T.use_nu) {
        g += T.nu * T.S24[e] / emn;
        h -= T.nu * T.nu * T.S24[e] * (emn + 1.0) / (emn * emn);
    }
}
PMM_HD double ultra2_value(const Ultra2Stats &T, int e, double x) {
    return ultra2_value_of(T, e, x, log(-expm1(-x)), T.use_nu ? log(-expm1(-T.nu * x)) : 0.0);
}
PMM_HD void ultra2_grad_hess(const Ultra2Stats &T, int e, double x, double &g, double &h) {
    ultra2_grad_hess_of(T, e, expm1(x), T.use_nu ? expm1(T.nu * x) : 1.0, g, h);
}

#if (defined(__CUDA_ARCH__) && !defined(PMM_HOST_EMU)) || defined(PMM_ULTRA2_EMU_WARP)      // (the latter: the warp form on the host, tests)
#define PMM_ULTRA2_WARP 1
// lane 4 e + j owns term j of edge e < 5 (lanes 20-31 repeat edge 4 and are ignored)
PMM_HD int ultra2_my_edge() { return min((int)(threadIdx.x & 31u) >> 2, 4); }
PMM_HD int ultra2_my_term() { return (int)(threadIdx.x & 3u); }
PMM_HD double ultra2_pick(const double *x, int e) { return e == 0 ? x[0] : e == 1 ? x[1] : e == 2 ? x[2] : e == 3 ? x[3] : x[4]; }
#else
#define PMM_ULTRA2_WARP 0
#endif

// edge lengths (v, u, C, A, B) of heights H = (Hu, Hv, Ht)
PMM_HD void ultra2_edges(const NniUltra &c, const double *H, double *x) {
    x[0] = H[2] - H[1]; x[1] = H[1] - H[0]; x[2] = H[1] - c.hC; x[3] = H[0] - c.hA; x[4] = H[0] - c.hB;
}

// objective = fsum + mu * bsum, the sums kept apart so that the value at an accepted point is carried to the next
// Newton step (and re-weighted when mu changes) instead of being evaluated again -- as pmm_ultrametric.cpp does
struct Ultra2Obj { double fsum, bsum; bool inside; };
PMM_HD Ultra2Obj ultra2_objective(const NniUltra &c, const Ultra2Stats &T, const double *H, double lo, double hi) {
    double x[5];
    Ultra2Obj o;
    o.fsum = 0.0; o.bsum = 0.0; o.inside = true;
    ultra2_edges(c, H, x);
#if PMM_ULTRA2_WARP
    // term 0: log(-expm1(-x)), 1: log(-expm1(-nu x)), 2: log(x - lo), 3: log(hi - x) -- one logarithm per lane, behind an
    // expm1 for terms 0 and 1 (an edge outside the box yields NaNs that nobody reads)
    const int j = ultra2_my_term(), me = ultra2_my_edge(), base = (int)(threadIdx.x & 28u);
    const double xe = ultra2_pick(x, me);
    double y = (j == 2) ? xe - lo : hi - xe;
    if (j < 2) y = -expm1(-(j == 0 ? xe : T.nu * xe));
    const double term = log(y);
    // the four lanes of an edge put its value and barrier term together (everything with a division stays per edge) ...
    const double tA = __shfl_sync(0xffffffffu, term, base), tB = __shfl_sync(0xffffffffu, term, base + 1);
    const double tC = __shfl_sync(0xffffffffu, term, base + 2), tD = __shfl_sync(0xffffffffu, term, base + 3);
    const double val = ultra2_value_of(T, me, xe, tA, tB), bar = tC + tD;
    // ... and every lane adds the five edges up in order
    double v5[5], b5[5];
    for (int e = 0; e < 5; e++) { v5[e] = __shfl_sync(0xffffffffu, val, 4 * e); b5[e] = __shfl_sync(0xffffffffu, bar, 4 * e); }
    for (int e = 0; e < 5; e++) {
        if (x[e] <= lo || x[e] >= hi) { o.inside = false; return o; }
        o.fsum += v5[e];
        o.bsum += b5[e];
    }
#else
    for (int e = 0; e < 5; e++) {
        if (x[e] <= lo || x[e] >= hi) { o.inside = false; return o; }
        o.fsum += ultra2_value(T, e, x[e]);
        o.bsum += log(x[e] - lo) + log(hi - x[e]);
    }
#endif
    return o;
}
PMM_HD double ultra2_obj_at(const Ultra2Obj &o, double mu) { return o.inside ? o.fsum + mu * o.bsum : -INFINITY; }

// returns 0 = optimal, 1 = infeasible, 2 = failed (non-finite slope).  S: clamped / rescaled statistics of the
// five branches, [5][5] branch-major.  H (Hu, Hv, Ht): in = previous iteration's heights when `warm`, out = the
// optimum; d5 = the five lengths.
PMM_HD int ultra2_solve(const NniUltra &c, const double *S, double nu, double eps_nu, double lo, double hi, bool warm,
                        double *H, double *d5, int *newton_iters, double mu_final = 1e-11, double warm_theta = 1e-6,
                        double warm_mu = 1e-9) {
    Ultra2Stats T;
    T.nu = nu; T.use_nu = nu > eps_nu;
    for (int e = 0; e < 5; e++) {
        const double *s = S + 5 * e;
        T.k[e] = (nu + 1.0) * s[0] + nu * (s[1] + s[3]);
        T.S1[e] = s[1];
        T.S24[e] = s[2] + s[4];
    }
    // interior point: every parent as low as its children allow, plus the margin (interior_heights)
    const double margin = fmax(1e-3, 0.5 * lo);
    double Hint[3];
    Hint[0] = fmax(c.hA, c.hB) + lo + margin;
    Hint[1] = fmax(Hint[0], c.hC) + lo + margin;
    Hint[2] = c.top_free ? Hint[1] + lo + margin : c.hTop;
    // a fixed top that the interior point would have to lift is infeasible: the lift propagates to a leaf
    if (!c.top_free && Hint[1] + lo + margin > c.hTop) { if (newton_iters) *newton_iters = 0; return 1; }
    double Hc[3], Hn[3], x[5];
    if (warm) {
        ultra2_edges(c, H, x);
        for (int e = 0; e < 5; e++) if (!(x[e] >= lo - 1e-12 && x[e] <= hi + 1e-12)) warm = false;
    }
    for (int i = 0; i < 3; i++) Hc[i] = warm ? (1.0 - warm_theta) * H[i] + warm_theta * Hint[i] : Hint[i];
    if (!c.top_free) Hc[2] = c.hTop;
    ultra2_edges(c, Hc, x);
    int iters = 0, status = 0;
    if (newton_iters) *newton_iters = 0;
    for (int e = 0; e < 5; e++) if (!(x[e] > lo && x[e] < hi)) return 1;
    const double n_free = 5.0;
    double mu = warm ? warm_mu : 1.0;
    Ultra2Obj at_H = ultra2_objective(c, T, Hc, lo, hi);          // objective sums at the current point
    for (;;) {
        for (int it = 0; it < 60; it++) {
            double g[5], w[5];
            ultra2_edges(c, Hc, x);
#if PMM_ULTRA2_WARP
            {
                const int me = ultra2_my_edge(), base = (int)(threadIdx.x & 28u);
                const double xe = ultra2_pick(x, me);
                const double em = expm1((ultra2_my_term() & 1) ? T.nu * xe : xe);         // terms 0 / 1 (2 / 3 repeat them)
                const double e1 = __shfl_sync(0xffffffffu, em, base), en = __shfl_sync(0xffffffffu, em, base + 1);
                double ge, he;
                ultra2_grad_hess_of(T, me, e1, en, ge, he);                              // the divisions of an edge stay on its lanes
                const double gm = ge + mu * (1.0 / (xe - lo) - 1.0 / (hi - xe));
                const double wm = he - mu * (1.0 / ((xe - lo) * (xe - lo)) + 1.0 / ((hi - xe) * (hi - xe)));
                for (int e = 0; e < 5; e++) { g[e] = __shfl_sync(0xffffffffu, gm, 4 * e); w[e] = __shfl_sync(0xffffffffu, wm, 4 * e); }
            }
#else
            for (int e = 0; e < 5; e++) {
                double ge, he;
                ultra2_grad_hess(T, e, x[e], ge, he);
                g[e] = ge + mu * (1.0 / (x[e] - lo) - 1.0 / (hi - x[e]));
                w[e] = he - mu * (1.0 / ((x[e] - lo) * (x[e] - lo)) + 1.0 / ((hi - x[e]) * (hi - x[e])));
            }
#endif
            // Newton direction in (Hu, Hv[, Ht]).  A free top is separable: its edge drops out of the (u, v) system
            const double g0 = c.top_free ? 0.0 : g[0], w0 = c.top_free ? 0.0 : w[0];
            const double gu = -g[1] + g[3] + g[4], gv = -g0 + g[1] + g[2];
            const double a11 = w[1] + w[3] + w[4], a22 = w0 + w[1] + w[2], a12 = -w[1];
            const double det = a11 * a22 - a12 * a12;
            double dl[3];
            dl[0] = -(a22 * gu - a12 * gv) / det;
            dl[1] = -(a11 * gv - a12 * gu) / det;
            dl[2] = c.top_free ? dl[1] - g[0] / w[0] : 0.0;
            double dx[5];
            dx[0] = dl[2] - dl[1]; dx[1] = dl[1] - dl[0]; dx[2] = dl[1]; dx[3] = dl[0]; dx[4] = dl[0];
            double slope = 0.0;
            for (int e = 0; e < 5; e++) slope += g[e] * dx[e];
            iters++;
            if (!isfinite(slope)) { status = 2; break; }
            if (slope <= 1e-13 * n_free * fmax(mu, 1e-12) + 1e-15) break;
            double tmax = INFINITY;
            for (int e = 0; e < 5; e++) {
                if (dx[e] < 0) tmax = fmin(tmax, (x[e] - lo) / (-dx[e]));
                else if (dx[e] > 0) tmax = fmin(tmax, (hi - x[e]) / dx[e]);
            }
            double t = fmin(1.0, 0.99 * tmax);
            const double f0 = ultra2_obj_at(at_H, mu);
            bool found = false;
            Ultra2Obj trial = at_H;
            while (t > 1e-14) {
                for (int i = 0; i < 3; i++) Hn[i] = Hc[i] + t * dl[i];
                trial = ultra2_objective(c, T, Hn, lo, hi);
                if (ultra2_obj_at(trial, mu) >= f0 + 1e-4 * t * slope) { found = true; break; }
                t *= 0.5;
            }
            if (!found) break;
            for (int i = 0; i < 3; i++) Hc[i] = Hn[i];
            at_H = trial;
            if (t * slope < 1e-14 * n_free) break;
        }
        if (status != 0 || mu <= mu_final) break;
        mu = fmax(mu * 0.1, mu_final);
    }
    ultra2_edges(c, Hc, x);
    for (int e = 0; e < 5; e++) d5[e] = fmin(fmax(x[e], lo), hi);
#if PMM_ULTRA2_WARP || defined(PMM_HOST_EMU)
    __syncwarp();                      // (H may be shared by the lanes: every lane has read it before any lane writes it)
#endif
    for (int i = 0; i < 3; i++) H[i] = Hc[i];
    if (newton_iters) *newton_iters = iters;
    return status;
}

}  // namespace pmm
