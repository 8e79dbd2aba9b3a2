// pmm_ultrametric_dev.cuh -- the ultrametric-constrained branch-length M-step ON THE DEVICE.
//
// Same program and same algorithm as the host routine pmm_ultrametric_solve (pmm_ultrametric.cpp; the reference:
// laml_libs/EM_solver.py:53-89,349-366 through cvxpy -> MOSEK): node heights re-parametrise the equalities away, a
// log-barrier Newton method remains, and every Newton system is solved by one children-to-parents elimination and one
// parents-to-children back-substitution over the tree.  Here the WHOLE solve -- interior point, warm start, barrier
// schedule, Newton directions, step-to-the-boundary rule, Armijo line search, stopping tests -- is ONE cooperative
// launch: the statistics never leave the device and the host sees one status word per M-step.
//
//  * Tree sweeps are level-synchronous and PARENT-centric: the thread that owns internal node u combines the
//    contributions of its two child edges in a fixed order (no atomics: results are bit-reproducible), on the level
//    lists the engine already keeps (internal nodes by height for the upward sweep, by depth for the downward one).
//  * Levels with many nodes are spread over the whole grid and end in a grid barrier; runs of consecutive NARROW
//    levels (the top of the tree, the deep tail) are walked by CTA 0 alone with __syncthreads between them, so a
//    100k-leaf tree needs ~45 grid barriers per Newton step instead of 2 x height.
//  * Every decision (convergence, step length, barrier weight) is taken by every thread from the same fixed-order
//    reductions, so all CTAs follow the same control flow without a host in the loop.
//  * The per-edge derivative g of a step is kept (the slope needs it), everything else lives in O(n) scratch arrays.
//    Values another CTA wrote are read with ld.global.cg (L1 is not coherent across SMs inside one launch).
#pragma once
#include "pmm_common.cuh"

namespace pmm {

enum : int { UM_FREE = 0, UM_FIXED = 1, UM_ZERO = 2 };
constexpr int UM_THREADS = 256;
constexpr int UM_NARROW = 2 * UM_THREADS;      // a level with at most this many nodes is "narrow"

struct UmPhase { int lvl0, lvl1, wide, pad; }; // levels [lvl0, lvl1) of a sweep; wide: one level on the whole grid

struct UmArgs {
    int n, n_roots, n_int;
    const int *parent;                 // [n], -1 for roots
    const int *left;                   // [n], -1 for leaves
    const int *roots;                  // [n_roots]
    const int4 *up_items, *dn_items;   // internal nodes (u, a, b, _) by height / by depth
    const int *up_ptr, *dn_ptr;        // level -> first item
    const UmPhase *up_ph, *dn_ph;
    int n_up_ph, n_dn_ph;
    const signed char *kind;           // [n]
    const double *fixed_len;           // [n]
    const double *stats;               // [n][5] raw expected counts
    double *Snorm;                     // [n][5] clamped and rescaled (EM_solver.py:343-346): also what the nu program reads
    const double *params;              // [0] = nu
    double lo, hi, eps_s, eps_nu, mu_final, warm_theta, warm_mu;
    int K_total, warm;                 // warm: H_prev holds the previous M-step's heights
    double *H_prev;                    // [n+1] in: previous heights (warm), out: this M-step's
    // scratch
    double *H, *Hint, *g, *a, *b, *alpha, *beta, *delta, *k, *S1, *S24;
    unsigned char *grounded;           // [n+1]
    double *partial;                   // [2][G][4]
    unsigned *bar;                     // grid barrier counter (zeroed before the launch)
    double *d_out;                     // [n] lengths of every branch at the optimum
    int *result;                       // [0] status (0 optimal, 1 infeasible, 2 failed), [1] Newton steps
};

__device__ __forceinline__ double um_ld(const double *p) { return __ldcg(p); }

struct UmCtx {
    const UmArgs &A;
    unsigned epoch;                    // grid barriers passed so far
    double (*sh)[UM_THREADS];          // [4][UM_THREADS] reduction scratch
    __device__ UmCtx(const UmArgs &a, double (*s)[UM_THREADS]) : A(a), epoch(0), sh(s) {}

    __device__ void grid_barrier() {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(A.bar, 1u);
            const unsigned target = (epoch + 1u) * gridDim.x;
            while (*((volatile unsigned *)A.bar) < target) {}
            __threadfence();
        }
        epoch++;
        __syncthreads();
    }
    // fixed-order grid reduction of up to three sums and one minimum / flag; the result is the same on every thread
    __device__ void reduce(double &s0, double &s1, double &s2, double &mn) {
        const int tid = threadIdx.x, G = gridDim.x;
        sh[0][tid] = s0; sh[1][tid] = s1; sh[2][tid] = s2; sh[3][tid] = mn;
        __syncthreads();
        for (int o = UM_THREADS / 2; o > 0; o >>= 1) {
            if (tid < o) {
                sh[0][tid] += sh[0][tid + o]; sh[1][tid] += sh[1][tid + o]; sh[2][tid] += sh[2][tid + o];
                sh[3][tid] = fmin(sh[3][tid], sh[3][tid + o]);
            }
            __syncthreads();
        }
        double *pb = A.partial + (size_t)(epoch & 1u) * G * 4;
        if (tid < 4) pb[blockIdx.x * 4 + tid] = sh[tid][0];
        grid_barrier();
        double t0 = 0.0, t1 = 0.0, t2 = 0.0, tm = INFINITY;
        for (int c = tid; c < G; c += UM_THREADS) {       // G <= UM_THREADS in practice: one element per thread
            t0 += __ldcg(pb + c * 4); t1 += __ldcg(pb + c * 4 + 1); t2 += __ldcg(pb + c * 4 + 2); tm = fmin(tm, __ldcg(pb + c * 4 + 3));
        }
        sh[0][tid] = t0; sh[1][tid] = t1; sh[2][tid] = t2; sh[3][tid] = tm;
        __syncthreads();
        for (int o = UM_THREADS / 2; o > 0; o >>= 1) {
            if (tid < o) {
                sh[0][tid] += sh[0][tid + o]; sh[1][tid] += sh[1][tid + o]; sh[2][tid] += sh[2][tid + o];
                sh[3][tid] = fmin(sh[3][tid], sh[3][tid + o]);
            }
            __syncthreads();
        }
        s0 = sh[0][0]; s1 = sh[1][0]; s2 = sh[2][0]; mn = sh[3][0];
        __syncthreads();
    }
    // run `fn(item)` over the levels of a sweep: wide levels on the whole grid, runs of narrow ones on CTA 0
    template <class F> __device__ void sweep(const UmPhase *ph, int nph, const int *ptr, const int4 *items, F fn) {
        for (int p = 0; p < nph; p++) {
            const UmPhase q = ph[p];
            if (q.wide) {
                const int i0 = ptr[q.lvl0], i1 = ptr[q.lvl0 + 1];
                for (int i = i0 + blockIdx.x * UM_THREADS + threadIdx.x; i < i1; i += gridDim.x * UM_THREADS) fn(items[i]);
            } else if (blockIdx.x == 0) {
                for (int l = q.lvl0; l < q.lvl1; l++) {
                    const int i0 = ptr[l], i1 = ptr[l + 1];
                    for (int i = i0 + threadIdx.x; i < i1; i += UM_THREADS) fn(items[i]);
                    __threadfence();             // the next level reads these values past L1 (ld.global.cg)
                    __syncthreads();
                }
            }
            grid_barrier();
        }
    }
};

__device__ __forceinline__ double um_rigid_len(const UmArgs &A, int v) { return A.kind[v] == UM_FIXED ? A.fixed_len[v] : 0.0; }
__device__ __forceinline__ double um_f_value(const UmArgs &A, int v, double x, double nu, bool use_nu) {
    double val = -A.k[v] * x + A.S1[v] * log(-expm1(-x));
    if (use_nu) val += A.S24[v] * log(-expm1(-nu * x));
    return val;
}
__device__ __forceinline__ void um_f_grad_hess(const UmArgs &A, int v, double x, double nu, bool use_nu, double &g, double &h) {
    const double em = expm1(x);
    g = -A.k[v] + A.S1[v] / em;
    h = -A.S1[v] * (em + 1.0) / (em * em);
    if (use_nu) {
        const double emn = expm1(nu * x);
        g += nu * A.S24[v] / emn;
        h -= nu * nu * A.S24[v] * (emn + 1.0) / (emn * emn);
    }
}

// the contribution of child edge c -> parent (height Hp) to the parent's reduced model, pmm_ultrametric.cpp: newton_direction
__device__ __forceinline__ void um_eliminate_edge(const UmArgs &A, int c, double Hp, double mu, double nu, bool use_nu,
                                                  double &Aout, double &Bout, bool &grounds_parent) {
    const int kd = A.kind[c];
    const bool rg = kd != UM_FREE, gr = __ldcg(A.grounded + c) != 0;
    double al = 0.0, be = 0.0, gE = 0.0;
    Aout = 0.0; Bout = 0.0;
    if (!rg) {
        const double x = Hp - um_ld(A.H + c);
        double g, h;
        um_f_grad_hess(A, c, x, nu, use_nu, g, h);
        gE = g + mu * (1.0 / (x - A.lo) - 1.0 / (A.hi - x));
        const double w = h - mu * (1.0 / ((x - A.lo) * (x - A.lo)) + 1.0 / ((A.hi - x) * (A.hi - x)));
        if (gr) { Aout = gE; Bout = w; }                                  // delta_c = 0
        else {                                                            // series combination
            const double ac = um_ld(A.a + c), bc = um_ld(A.b + c), s = w + bc;
            al = (gE - ac) / s; be = w / s;
            Aout = (gE * bc + ac * w) / s; Bout = w * bc / s;
        }
    } else if (!gr) { Aout = um_ld(A.a + c); Bout = um_ld(A.b + c); be = 1.0; }   // rigid edge: delta_c = delta_p
    else grounds_parent = true;                                                  // rigid edge to a grounded child
    A.g[c] = gE; A.alpha[c] = al; A.beta[c] = be;
}

// delta of child c from its parent's, and the edge's share of the slope and of the step-to-the-boundary bound
__device__ __forceinline__ void um_back_edge(const UmArgs &A, int c, double Hp, double dp, double &slope, double &tmax) {
    const double dc = (__ldcg(A.grounded + c) != 0) ? 0.0 : um_ld(A.alpha + c) + um_ld(A.beta + c) * dp;
    A.delta[c] = dc;
    const double dx = dp - dc;
    slope += um_ld(A.g + c) * dx;
    if (A.kind[c] == UM_FREE) {
        const double x = Hp - um_ld(A.H + c);
        if (dx < 0) tmax = fmin(tmax, (x - A.lo) / (-dx));
        else if (dx > 0) tmax = fmin(tmax, (A.hi - x) / dx);
    }
}

__global__ void __launch_bounds__(UM_THREADS) pmm_um_solve(const UmArgs A) {
    __shared__ double sh[4][UM_THREADS];
    UmCtx C(A, sh);
    const int n = A.n, tid = threadIdx.x, gtid = blockIdx.x * UM_THREADS + tid, gsz = gridDim.x * UM_THREADS;
    const double nu = A.params[0], lo = A.lo, hi = A.hi;
    const bool use_nu = nu > A.eps_nu;
    const bool lead = blockIdx.x == 0 && tid == 0;

    // ---- statistics: clamp, rescale (EM_solver.py:343-346), and the three combinations the objective needs
    double nfree_loc = 0.0;
    for (int v = gtid; v < n; v += gsz) {
        double s[5], tot = 0.0;
#pragma unroll
        for (int j = 0; j < 5; j++) { s[j] = fmax(A.stats[(size_t)v * 5 + j], A.eps_s); tot += s[j]; }
#pragma unroll
        for (int j = 0; j < 5; j++) { s[j] = s[j] / tot * (double)A.K_total; A.Snorm[(size_t)v * 5 + j] = s[j]; }
        A.k[v] = (nu + 1.0) * s[0] + nu * (s[1] + s[3]);
        A.S1[v] = s[1];
        A.S24[v] = s[2] + s[4];
        A.grounded[v] = A.left[v] < 0;                   // leaves cannot move
        A.Hint[v] = 0.0;
        nfree_loc += (A.kind[v] == UM_FREE) ? 1.0 : 0.0;
    }
    if (lead) { A.Hint[n] = 0.0; A.grounded[n] = 0; }
    C.grid_barrier();

    // ---- a strictly feasible point (interior_heights): every free edge lo + margin long, then rigid edges made exact
    const double margin = fmax(1e-3, 0.5 * lo);
    auto need_of = [&](int c) { return um_ld(A.Hint + c) + (A.kind[c] == UM_FREE ? lo + margin : um_rigid_len(A, c)); };
    C.sweep(A.up_ph, A.n_up_ph, A.up_ptr, A.up_items, [&](const int4 it) { A.Hint[it.x] = fmax(need_of(it.y), need_of(it.z)); });
    if (lead) {
        double top = 0.0;
        for (int r = 0; r < A.n_roots; r++) top = fmax(top, need_of(A.roots[r]));
        A.Hint[n] = top;
        for (int r = 0; r < A.n_roots; r++) {
            const int v = A.roots[r];
            if (A.kind[v] != UM_FREE && A.left[v] >= 0) A.Hint[v] = top - um_rigid_len(A, v);
        }
    }
    C.grid_barrier();
    C.sweep(A.dn_ph, A.n_dn_ph, A.dn_ptr, A.dn_items, [&](const int4 it) {
        const double Hu = um_ld(A.Hint + it.x);
        if (A.kind[it.y] != UM_FREE && A.left[it.y] >= 0) A.Hint[it.y] = Hu - um_rigid_len(A, it.y);
        if (A.kind[it.z] != UM_FREE && A.left[it.z] >= 0) A.Hint[it.z] = Hu - um_rigid_len(A, it.z);
    });

    // feasibility of a height vector: count of violated edges (strict: inside the box; else with the host's tolerances)
    auto infeasible_edges = [&](const double *Hc, bool strict) {
        double bad = 0.0;
        for (int v = gtid; v < n; v += gsz) {
            const int p = A.parent[v] < 0 ? n : A.parent[v];
            const double hv = um_ld(Hc + v), x = um_ld(Hc + p) - hv;
            if (A.kind[v] == UM_FREE) { if (strict ? !(x > lo && x < hi) : !(x >= lo - 1e-12 && x <= hi + 1e-12)) bad += 1.0; }
            else { const double want = um_rigid_len(A, v); if (!(fabs(x - want) <= 1e-9 + 1e-5 * fabs(want))) bad += 1.0; }
            if (A.left[v] < 0 && hv != 0.0) bad += 1.0;
        }
        return bad;
    };
    // ---- warm start from the previous M-step's heights, pulled a little way towards the interior point
    double z1 = 0.0, z2 = 0.0, zm = INFINITY;
    double bad_prev = A.warm ? infeasible_edges(A.H_prev, false) : 1.0;
    C.reduce(bad_prev, nfree_loc, z2, zm);
    const bool warm = A.warm && bad_prev == 0.0;
    const double n_free = fmax(nfree_loc, 1.0);
    for (int v = gtid; v <= n; v += gsz)
        A.H[v] = warm ? (1.0 - A.warm_theta) * um_ld(A.H_prev + v) + A.warm_theta * um_ld(A.Hint + v) : um_ld(A.Hint + v);
    C.grid_barrier();
    double bad = infeasible_edges(A.H, true);
    z1 = 0.0; z2 = 0.0; zm = INFINITY;
    C.reduce(bad, z1, z2, zm);
    if (bad != 0.0) {                                                          // EM_solver's "d_infeasible"
        if (lead) { A.result[0] = 1; A.result[1] = 0; }
        return;
    }

    // objective = fsum + mu * bsum over the free edges at H + t * delta (t = 0: at H); outside the box: -inf
    auto objective_parts = [&](double t, double &fsum, double &bsum) -> bool {
        double f = 0.0, bsm = 0.0, out = 0.0, mn = INFINITY;
        for (int v = gtid; v < n; v += gsz) {
            if (A.kind[v] != UM_FREE) continue;
            const int p = A.parent[v] < 0 ? n : A.parent[v];
            double hp = um_ld(A.H + p), hv = um_ld(A.H + v);
            if (t != 0.0) { hp += t * um_ld(A.delta + p); hv += t * um_ld(A.delta + v); }
            const double x = hp - hv;
            if (x <= lo || x >= hi) { out += 1.0; continue; }
            f += um_f_value(A, v, x, nu, use_nu);
            bsm += log(x - lo) + log(hi - x);
        }
        C.reduce(f, bsm, out, mn);
        fsum = f; bsum = bsm;
        return out == 0.0;
    };

    double mu = warm ? A.warm_mu : 1.0;
    int iters = 0, status = 0;
    double atH_f = 0.0, atH_b = 0.0;
    bool have_at_H = false;
    for (;;) {
        for (int it = 0; it < 60; it++) {
            // ---- Newton direction: children-to-parents elimination ...
            C.sweep(A.up_ph, A.n_up_ph, A.up_ptr, A.up_items, [&](const int4 q) {
                const double Hu = um_ld(A.H + q.x);
                double A1, B1, A2, B2;
                bool gr = false;
                um_eliminate_edge(A, q.y, Hu, mu, nu, use_nu, A1, B1, gr);
                um_eliminate_edge(A, q.z, Hu, mu, nu, use_nu, A2, B2, gr);
                A.a[q.x] = A1 + A2; A.b[q.x] = B1 + B2; A.grounded[q.x] = gr;
            });
            double slope = 0.0, tmax = INFINITY;
            if (lead) {                                                     // the virtual super-root above the roots
                const double Hn_ = um_ld(A.H + n);
                double an = 0.0, bn = 0.0;
                bool gr = false;
                for (int r = 0; r < A.n_roots; r++) { double Ar, Br; um_eliminate_edge(A, A.roots[r], Hn_, mu, nu, use_nu, Ar, Br, gr); an += Ar; bn += Br; }
                const double dn = (!gr && bn < 0) ? -an / bn : 0.0;
                A.delta[n] = dn;
                __threadfence();
                for (int r = 0; r < A.n_roots; r++) um_back_edge(A, A.roots[r], Hn_, dn, slope, tmax);
            }
            C.grid_barrier();
            // ... and parents-to-children back-substitution, with the slope and the distance to the boundary on the way
            C.sweep(A.dn_ph, A.n_dn_ph, A.dn_ptr, A.dn_items, [&](const int4 q) {
                const double Hu = um_ld(A.H + q.x), du = um_ld(A.delta + q.x);
                um_back_edge(A, q.y, Hu, du, slope, tmax);
                um_back_edge(A, q.z, Hu, du, slope, tmax);
            });
            z1 = 0.0; z2 = 0.0;
            C.reduce(slope, z1, z2, tmax);
            iters++;
            if (!isfinite(slope)) { status = 2; break; }
            if (slope <= 1e-13 * n_free * fmax(mu, 1e-12) + 1e-15) break;
            double t = fmin(1.0, 0.99 * tmax);
            if (!have_at_H) { objective_parts(0.0, atH_f, atH_b); have_at_H = true; }
            const double f0 = atH_f + mu * atH_b;
            bool found = false;
            double tf = 0.0, tb = 0.0;
            while (t > 1e-14) {
                const bool inside = objective_parts(t, tf, tb);
                if (inside && tf + mu * tb >= f0 + 1e-4 * t * slope) { found = true; break; }
                t *= 0.5;
            }
            if (!found) break;
            for (int v = gtid; v <= n; v += gsz) A.H[v] = um_ld(A.H + v) + t * um_ld(A.delta + v);
            C.grid_barrier();
            atH_f = tf; atH_b = tb;
            if (t * slope < 1e-14 * n_free) break;
        }
        if (status != 0 || mu <= A.mu_final) break;
        mu = fmax(mu * 0.1, A.mu_final);
    }
    for (int v = gtid; v < n; v += gsz) {
        const int p = A.parent[v] < 0 ? n : A.parent[v];
        double d = um_ld(A.H + p) - um_ld(A.H + v);
        if (A.kind[v] == UM_FIXED) d = A.fixed_len[v];
        else if (A.kind[v] == UM_FREE) d = fmin(fmax(d, lo), hi);
        A.d_out[v] = d;
    }
    for (int v = gtid; v <= n; v += gsz) A.H_prev[v] = um_ld(A.H + v);
    if (lead) { A.result[0] = status; A.result[1] = iters; }
}

// the free branches take their new lengths -- unless the program was infeasible (EM_solver.py:425-426: nothing is updated)
__global__ void pmm_um_apply(int n, const signed char *__restrict__ kind, const double *__restrict__ d_new, const int *__restrict__ result,
                             double *__restrict__ d) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n && result[0] != 1 && kind[v] == UM_FREE) d[v] = d_new[v];
}

}  // namespace pmm
