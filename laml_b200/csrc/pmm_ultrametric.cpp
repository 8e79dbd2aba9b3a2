// Ultrametric-constrained branch-length M-step on the host, C ABI: pmm_ultrametric_solve.
//
// The reference maximises  sum_b -(nu+1) S0 d - nu (S1+S3) d + S1 log(1-e^-d) + (S2+S4) log(1-e^{-nu d})
// over dmin <= d <= dmax subject to the ultrametric equalities M d = b (laml_libs/EM_solver.py:53-89,349-366;
// cvxpy -> MOSEK).  Here node heights H re-parametrise the program (d_v = H_parent - H_v, leaves at 0, a fixed
// branch is an equality, a polytomy-marked branch counts as 0, a virtual super-root n turns root branches and
// the "all trees equally high" rows into ordinary edges), the equalities vanish, and a log-barrier Newton method
// remains whose linear systems have the tree's sparsity: one children-to-parents pass eliminates every node (the
// Schur complement of an edge is the series combination w b / (w + b) of its curvature with the child's), one
// parents-to-children pass back-substitutes.  O(n) memory and time per Newton step.
//
// Pure host code (no CUDA): it lives in libpmm_b200.so because the statistics it consumes come out of pmm_estep,
// and so that round 2 can swap the two passes for level kernels behind the same entry point.  The numpy model
// of the same algorithm, laml_b200/ultrametric.py: solve_numpy, is what tests/test_ultrametric_host.py checks it
// against.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <thread>
#include <vector>

namespace {

// The per-edge work (two expm1, up to four log per edge and evaluation) dominates at 10^5 nodes; it runs over
// fixed-size chunks of the free-edge list on a few host threads.  Partial sums are kept per chunk and added in
// chunk order, so results do not depend on the number of threads.
constexpr int CHUNK = 4096;

int host_threads() {
    static const int n = [] {
        const char *e = std::getenv("PMM_HOST_THREADS");
        int t = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
        return t < 1 ? 1 : (t > 16 ? 16 : t);
    }();
    return n;
}

template <class F>
void for_chunks(int nchunks, F fn) {
    const int nt = std::min(host_threads(), nchunks / 4);          // below ~16k edges threads cost more than they save
    if (nt <= 1) { for (int c = 0; c < nchunks; c++) fn(c); return; }
    std::atomic<int> next(0);
    auto work = [&] { for (int c = next++; c < nchunks; c = next++) fn(c); };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back(work);
    work();
    for (auto &th : pool) th.join();
}

enum : int8_t { FREE = 0, FIXED = 1, ZERO = 2 };

struct Problem {
    int n;                                   // real nodes; node n is the virtual super-root
    std::vector<int> par;                    // parent (roots -> n)
    const int8_t *kind;
    const double *fixed_len;
    double lo, hi;
    std::vector<char> is_leaf;
    std::vector<int> up;                     // children before parents (real nodes only)
    std::vector<int> fidx;                   // free edges
};

inline double rigid_len(const Problem &P, int v) { return P.kind[v] == FIXED ? P.fixed_len[v] : 0.0; }

void build(Problem &P, int n, const int32_t *parent, const int8_t *kind, const double *fixed_len, double lo, double hi) {
    P.n = n; P.kind = kind; P.fixed_len = fixed_len; P.lo = lo; P.hi = hi;
    P.par.resize(n);
    std::vector<int> deg(n + 1, 0);
    for (int v = 0; v < n; v++) { P.par[v] = parent[v] < 0 ? n : parent[v]; deg[P.par[v]]++; }
    P.is_leaf.assign(n, 0);
    P.up.clear(); P.up.reserve(n);
    for (int v = 0; v < n; v++) if (deg[v] == 0) { P.is_leaf[v] = 1; P.up.push_back(v); }
    for (size_t i = 0; i < P.up.size(); i++) {          // Kahn: a node enters once all its children have
        const int p = P.par[P.up[i]];
        if (p < n && --deg[p] == 0) P.up.push_back(p);
    }
    P.fidx.clear();
    for (int v = 0; v < n; v++) if (kind[v] == FREE) P.fidx.push_back(v);
}

// a strictly feasible height assignment: every free edge strictly inside its box
void interior_heights(const Problem &P, std::vector<double> &H) {
    const int n = P.n;
    H.assign(n + 1, 0.0);
    const double margin = std::fmax(1e-3, 0.5 * P.lo);
    for (int v : P.up) {
        const double need = H[v] + (P.kind[v] == FREE ? P.lo + margin : rigid_len(P, v));
        if (need > H[P.par[v]]) H[P.par[v]] = need;
    }
    // rigid edges need equality: a parent lifted by another child lifts the rigid child (and its subtree)
    for (size_t i = P.up.size(); i-- > 0;) {
        const int v = P.up[i];
        if (P.kind[v] != FREE && !P.is_leaf[v]) H[v] = H[P.par[v]] - rigid_len(P, v);
    }
}

bool feasible(const Problem &P, const std::vector<double> &H, bool strict) {
    const double tol = 1e-12;
    for (int v = 0; v < P.n; v++) {
        const double x = H[P.par[v]] - H[v];
        if (P.kind[v] == FREE) {
            if (strict ? !(x > P.lo && x < P.hi) : !(x >= P.lo - tol && x <= P.hi + tol)) return false;
        } else {
            const double want = rigid_len(P, v);
            if (!(std::fabs(x - want) <= 1e-9 + 1e-5 * std::fabs(want))) return false;      // numpy.allclose(atol=1e-9)
        }
        if (P.is_leaf[v] && H[v] != 0.0) return false;
    }
    return true;
}

struct Stats { std::vector<double> k, S1, S24; double nu; bool use_nu; };

// per-branch objective of EM_solver.py:351-357: the value alone (expm1 + log per term) and the derivatives alone
// (one expm1 per term), so that neither caller pays for what it does not use
inline double f_value(const Stats &T, int v, double x) {
    double val = -T.k[v] * x + T.S1[v] * std::log(-std::expm1(-x));
    if (T.use_nu) val += T.S24[v] * std::log(-std::expm1(-T.nu * x));
    return val;
}
inline void f_grad_hess(const Stats &T, int v, double x, double &g, double &h) {
    const double em = std::expm1(x);
    g = -T.k[v] + T.S1[v] / em;
    h = -T.S1[v] * (em + 1.0) / (em * em);
    if (T.use_nu) {
        const double emn = std::expm1(T.nu * x);
        g += T.nu * T.S24[v] / emn;
        h -= T.nu * T.nu * T.S24[v] * (emn + 1.0) / (emn * emn);
    }
}

// objective = fsum + mu * bsum; the two sums are returned separately so that the value at an accepted point can be
// carried to the next Newton step (and re-weighted when mu changes) instead of being evaluated again
struct ObjParts { double fsum, bsum; bool inside; };

ObjParts objective_parts(const Problem &P, const Stats &T, const std::vector<double> &H) {
    const int m = (int)P.fidx.size(), nchunks = (m + CHUNK - 1) / CHUNK;
    std::vector<double> sum(nchunks, 0.0), bar(nchunks, 0.0);
    std::atomic<bool> outside(false);
    for_chunks(nchunks, [&](int c) {
        double s = 0.0, b = 0.0;
        for (int i = c * CHUNK, e = std::min(m, (c + 1) * CHUNK); i < e; i++) {
            const int v = P.fidx[i];
            const double x = H[P.par[v]] - H[v];
            if (x <= P.lo || x >= P.hi) { outside = true; return; }
            s += f_value(T, v, x);
            b += std::log(x - P.lo) + std::log(P.hi - x);
        }
        sum[c] = s; bar[c] = b;
    });
    if (outside) return {0.0, 0.0, false};
    double s = 0.0, b = 0.0;
    for (int c = 0; c < nchunks; c++) { s += sum[c]; b += bar[c]; }
    return {s, b, true};
}

inline double objective_of(const ObjParts &o, double mu) {
    return o.inside ? o.fsum + mu * o.bsum : -std::numeric_limits<double>::infinity();
}

struct Work { std::vector<double> g, w, a, b, alpha, beta, delta; std::vector<char> grounded; };

// Newton direction of the barrier objective in H; returns the directional derivative along it
double newton_direction(const Problem &P, const Stats &T, const std::vector<double> &H, double mu, Work &W) {
    const int n = P.n;
    W.g.assign(n, 0.0); W.w.assign(n, 0.0);
    const int m = (int)P.fidx.size();
    for_chunks((m + CHUNK - 1) / CHUNK, [&](int c) {
        for (int i = c * CHUNK, e = std::min(m, (c + 1) * CHUNK); i < e; i++) {
            const int v = P.fidx[i];
            const double x = H[P.par[v]] - H[v];
            double g, h;
            f_grad_hess(T, v, x, g, h);
            W.g[v] = g + mu * (1.0 / (x - P.lo) - 1.0 / (P.hi - x));
            W.w[v] = h - mu * (1.0 / ((x - P.lo) * (x - P.lo)) + 1.0 / ((P.hi - x) * (P.hi - x)));
        }
    });
    // children-to-parents elimination: reduced model a*delta + b*delta^2/2 per node (b <= 0); a grounded node
    // cannot move (a leaf, or the parent of a grounded child across a rigid edge)
    W.a.assign(n + 1, 0.0); W.b.assign(n + 1, 0.0);
    W.alpha.assign(n, 0.0); W.beta.assign(n, 0.0);
    W.grounded.assign(n + 1, 0);
    for (int v = 0; v < n; v++) W.grounded[v] = P.is_leaf[v];
    for (int c : P.up) {
        const int p = P.par[c];
        const bool gr = W.grounded[c], rg = P.kind[c] != FREE;
        double A = 0.0, B = 0.0;
        if (!rg && gr) { A = W.g[c]; B = W.w[c]; }                              // delta_c = 0
        else if (!rg) {                                                          // series combination
            const double s = W.w[c] + W.b[c];
            W.alpha[c] = (W.g[c] - W.a[c]) / s;
            W.beta[c] = W.w[c] / s;
            A = (W.g[c] * W.b[c] + W.a[c] * W.w[c]) / s;
            B = W.w[c] * W.b[c] / s;
        } else if (!gr) { A = W.a[c]; B = W.b[c]; W.beta[c] = 1.0; }             // rigid edge: delta_c = delta_p
        else W.grounded[p] = 1;                                                  // rigid edge to a grounded child
        W.a[p] += A; W.b[p] += B;
    }
    W.delta.assign(n + 1, 0.0);
    if (!W.grounded[n] && W.b[n] < 0) W.delta[n] = -W.a[n] / W.b[n];
    for (size_t i = P.up.size(); i-- > 0;) {
        const int c = P.up[i];
        if (!W.grounded[c]) W.delta[c] = W.alpha[c] + W.beta[c] * W.delta[P.par[c]];
    }
    double slope = 0.0;
    for (int v = 0; v < n; v++) slope += W.g[v] * (W.delta[P.par[v]] - W.delta[v]);
    return slope;
}

}  // namespace

extern "C" int pmm_ultrametric_solve(int32_t n, const int32_t *parent, const int8_t *kind, const double *fixed_len,
                                     double lo, double hi, const double *S, double nu, double eps_nu, const double *H0,
                                     double mu_final, double warm_theta, double warm_mu, double *d_out, double *H_out,
                                     int32_t *newton_iters) {
    if (n <= 0 || !parent || !kind || !fixed_len || !S || !d_out || !H_out) return -1;
    Problem P;
    build(P, n, parent, kind, fixed_len, lo, hi);
    if ((int)P.up.size() != n) return -1;                                       // not a forest
    Stats T;
    T.nu = nu; T.use_nu = nu > eps_nu;
    T.k.resize(n); T.S1.resize(n); T.S24.resize(n);
    for (int v = 0; v < n; v++) {
        const double *s = S + (size_t)5 * v;
        T.k[v] = (nu + 1.0) * s[0] + nu * (s[1] + s[3]);
        T.S1[v] = s[1];
        T.S24[v] = s[2] + s[4];
    }
    std::vector<double> Hint, H, Hn(n + 1);
    interior_heights(P, Hint);
    bool warm = false;
    if (H0) {
        std::vector<double> h0(H0, H0 + n + 1);
        warm = feasible(P, h0, false);
        if (warm) {
            H.resize(n + 1);
            for (int v = 0; v <= n; v++) H[v] = (1.0 - warm_theta) * h0[v] + warm_theta * Hint[v];
        }
    }
    if (!warm) H = Hint;
    if (newton_iters) *newton_iters = 0;
    if (!feasible(P, H, true)) return 1;                                        // infeasible
    const double n_free = (double)(P.fidx.empty() ? 1 : P.fidx.size());
    Work W;
    double mu = warm ? warm_mu : 1.0;
    int iters = 0, status = 0;
    ObjParts at_H = {0.0, 0.0, false};                                          // objective sums at the current H
    bool have_at_H = false;
    for (;;) {
        for (int it = 0; it < 60; it++) {
            const double slope = newton_direction(P, T, H, mu, W);
            iters++;
            if (!std::isfinite(slope)) { status = 2; break; }
            if (slope <= 1e-13 * n_free * std::fmax(mu, 1e-12) + 1e-15) break;
            // largest step that keeps every free edge strictly inside its box
            double tmax = std::numeric_limits<double>::infinity();
            for (int v : P.fidx) {
                const double dx = W.delta[P.par[v]] - W.delta[v], x = H[P.par[v]] - H[v];
                if (dx < 0) tmax = std::fmin(tmax, (x - lo) / (-dx));
                else if (dx > 0) tmax = std::fmin(tmax, (hi - x) / dx);
            }
            double t = std::fmin(1.0, 0.99 * tmax);
            if (!have_at_H) { at_H = objective_parts(P, T, H); have_at_H = true; }
            const double f0 = objective_of(at_H, mu);
            bool found = false;
            ObjParts trial = at_H;
            while (t > 1e-14) {
                for (int v = 0; v <= n; v++) Hn[v] = H[v] + t * W.delta[v];
                trial = objective_parts(P, T, Hn);
                if (objective_of(trial, mu) >= f0 + 1e-4 * t * slope) { found = true; break; }
                t *= 0.5;
            }
            if (!found) break;
            H.swap(Hn);
            at_H = trial;                                                        // the accepted point's sums
            if (t * slope < 1e-14 * n_free) break;
        }
        if (status != 0 || mu <= mu_final) break;
        mu = std::fmax(mu * 0.1, mu_final);
    }
    for (int v = 0; v < n; v++) {
        double d = H[P.par[v]] - H[v];
        if (kind[v] == FIXED) d = fixed_len[v];
        else if (kind[v] == FREE) d = std::fmin(std::fmax(d, lo), hi);
        d_out[v] = d;
    }
    for (int v = 0; v <= n; v++) H_out[v] = H[v];
    if (newton_iters) *newton_iters = iters;
    return status;
}
