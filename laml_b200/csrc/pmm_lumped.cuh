// pmm_lumped.cuh -- sm_100a kernels of the LUMPED-state PMM engine (the default E-step path).
//
// What is computed: exactly the quantities of laml_libs/EM_solver.py:91-300 (inside pass, outside
// pass, expected transition counts, R~), in the full-state formulation of SURVEY.md section 3.5.
//
// What is stored: because edits are irreversible, the inside vector of ANY node is, at every site,
//       x_alpha = xb + [alpha == h] (xh - xb)              (at most one "hot" edited state h)
// and a hot state implies xb = 0 and x_m = 0 exactly (a subtree that shows an edited state cannot
// hang below a silenced or differently edited ancestor; a second, different hot state below the
// same node zeroes every edited slot).  The outside vector is
//       out_alpha = oB q_alpha + [alpha == k] oD .
// So a (node, site) is the 40-byte record (x0, xm, xz, qh | h, cum) -- xz is xb when there is no hot
// state and the hot value when there is one, qh = q_h, cum the accumulated power-of-two exponent --
// and (o0, om, oB, oD | k, F) for the outside pass.  Every sum over the alphabet has a closed form
// in qsum = sum_alpha q_alpha and q_h: SURVEY.md section 8's "exact lumping" (states no leaf of the
// subtree shows are indistinguishable) applied per (node, site) instead of per site.  The numpy
// model tests/lumped_model.py restates these closed forms and is pinned to the reference's dumps.
//
// Execution: sites are independent given the tree, so ONE THREAD OWNS ONE SITE and walks a whole
// CLADE (a connected piece of the tree with ~m internal nodes, cut out by the host) without any
// synchronisation: a CTA is (clade, 128 consecutive sites); its threads follow the clade's step
// list -- post-order for the inside pass, pre-order for the outside pass -- so that
//   * the record of the child visited last (inside) / the out-record of the child visited first
//     (outside) stays in registers: no memory round trip on the dependency chain;
//   * every other operand was written at least one step earlier (by this very thread, or by an
//     earlier launch for children in other clades) and is PREFETCHED one step ahead;
//   * all global accesses are coalesced across the warp's 32 sites (32 B + 8 B per thread).
// Clades that hang below other clades run in an earlier (inside) / later (outside) launch: a
// 100k-cell tree needs ~4 launches per pass instead of one per tree level.
// Expected counts: the 10 per-thread values of a step (5 per child branch) are summed over the
// warp's sites with a halving butterfly (16 shuffles) and written once per (warp, node) into
// Spart[site group][node][5]; pmm_reduce_S sums the site groups in a fixed order (bit-reproducible).
#pragma once
#include "pmm_kernels.cuh"

namespace pmm {

constexpr int LT = 128;                  // threads (= sites) per CTA

struct LParams {
    int K;                   // sites of this chunk
    int site0;               // first site of the chunk in the handle's site space
    int code_stride;         // row stride of `codes` (all sites of the handle)
    int n_leaves, n_nodes, n_clades;
    const int *codes;        // [n_leaves][code_stride] raw state codes
    const int *site_ptr;     // [K_handle + 1] CSR offsets into `prior`
    const double *prior;     // [nnz]
    const double *qsum;      // [K_handle] sum_alpha q_alpha
    const double *params;    // [0] = nu, [1] = phi
    const double *bc;        // [n][4] = a, b, c, g per branch (pmm_branch_consts)
    double4 *X; int2 *XI;    // [n_int][K] inside records (x0, xm, xz, qh), (h, cum)
    double4 *O; int2 *OI;    // [n_int][K] outside records (o0, om, oB, oD), (k, F)
    double *invL;            // [n_roots][K] 1 / (normalised site likelihood)
    double *root_site_llh;   // [n_roots][K]
    double *Spart;           // [G][n][5], G = ceil(K / 32) site groups (one per warp)
    double *Rpart;           // [G][n_clades]
    double *post0, *post1;   // optional [n][K_total] log posteriors (annotation path), or null
    int K_total;
    int store_all;           // outside pass: also store the out-records that are chained in registers
};

// inside step flags:  bits 0-1 kind of child v, bits 2-3 kind of child w, bit 4 = u is a root
// outside step flags: bits 0-1 where out[u] comes from (0 = u is a root, 1 = load, 2 = chained),
//                     bit 2 / 3: v / w is a leaf, bit 4 / 5: store out[v] / out[w],
//                     bits 6-7: 1 = out[v] is chained into the next step, 2 = out[w] is
enum { KIND_LEAF = 0, KIND_LOAD = 1, KIND_PREV = 2 };

struct Rec { double x0, xm, xz, qh; int h, cum; };
struct ORec { double o0, om, oB, oD; int k, F; };
struct Msg { double E0, Em, Eb, Eh, sq, xb; };

// what a step reads from memory for one child: a record (internal child) or a state code (leaf)
struct Pre { double4 r; int2 ri; };

// branch constants (a, b, c, g) of node v: uniform across the warp, read-only path
__device__ __forceinline__ double4 ld_bc(const double *__restrict__ bc, int v) {
    const double2 lo = __ldg(reinterpret_cast<const double2 *>(bc) + 2 * (size_t)v);
    const double2 hi = __ldg(reinterpret_cast<const double2 *>(bc) + 2 * (size_t)v + 1);
    return make_double4(lo.x, lo.y, hi.x, hi.y);
}

__device__ __forceinline__ Pre fetch_child(const LParams &P, int kind, int c, int s, int gs) {
    Pre p;
    p.r = make_double4(0.0, 0.0, 0.0, 0.0);
    p.ri = make_int2(0, 0);
    if (kind == KIND_LEAF) p.ri.x = __ldg(P.codes + (size_t)c * P.code_stride + gs);
    else if (kind == KIND_LOAD) {
        const size_t o = (size_t)(c - P.n_leaves) * P.K + s;
        p.r = P.X[o];
        p.ri = P.XI[o];
    }
    return p;
}

// leaf record from its state code (EM_solver.py:103-116 in full-state form):
//   '?' (-1) -> (phi, 1, phi..)   known silenced (-2) -> e_m   0 -> (1-phi) e_0   a -> (1-phi) e_a
__device__ __forceinline__ Rec leaf_rec(int code, double phi, const double *__restrict__ prior_site) {
    Rec r;
    r.x0 = (code == 0) ? 1.0 - phi : ((code == -1) ? phi : 0.0);
    r.xm = (code < 0) ? 1.0 : 0.0;
    r.xz = (code == -1) ? phi : ((code > 0) ? 1.0 - phi : 0.0);
    r.h = (code > 0) ? code - 1 : -1;
    r.qh = (code > 0) ? __ldg(prior_site + (code - 1)) : 0.0;
    r.cum = 0;
    return r;
}

__device__ __forceinline__ Rec child_rec(int kind, const Pre &p, const Rec &prev, double phi, const double *__restrict__ prior_site) {
    if (kind == KIND_LEAF) return leaf_rec(p.ri.x, phi, prior_site);
    if (kind == KIND_PREV) return prev;
    Rec r;
    r.x0 = p.r.x; r.xm = p.r.y; r.xz = p.r.z; r.qh = p.r.w; r.h = p.ri.x; r.cum = p.ri.y;
    return r;
}

// child -> parent message (SURVEY 3.5 "edge transfer") in lumped form:
//   E_0 = a x_0 + b sum(q x) + g x_m,  E_m = x_m,  E_alpha = c x_alpha + g x_m  (Eb off the hot slot, Eh on it)
__device__ __forceinline__ Msg message(const Rec &r, const double4 c, double qsum) {
    Msg m;
    const bool hot = r.h >= 0;
    m.xb = hot ? 0.0 : r.xz;
    m.sq = r.xz * (hot ? r.qh : qsum);
    const double gm = c.w * r.xm;
    m.E0 = fma(c.x, r.x0, fma(c.y, m.sq, gm));
    m.Em = r.xm;
    m.Eb = fma(c.z, m.xb, gm);
    m.Eh = fma(c.z, r.xz, gm);
    return m;
}

// ---------------------------------------------------------------- inside pass
__global__ void __launch_bounds__(LT) pmm_lumped_inside(LParams P, const int4 *__restrict__ clades, const int4 *__restrict__ steps) {
    const int4 cl = clades[blockIdx.x];                 // steps [x, y), root index z
    const int N = P.n_leaves;
    int s = blockIdx.y * LT + threadIdx.x;
    const bool valid = s < P.K;
    s = min(s, P.K - 1);
    const int gs = P.site0 + s;
    const double *prior_site = P.prior + __ldg(P.site_ptr + gs);
    const double qsum = __ldg(P.qsum + gs);
    const double phi = __ldg(P.params + 1);

    Rec prev;
    prev.x0 = prev.xm = prev.xz = prev.qh = 0.0; prev.h = -1; prev.cum = 0;
    int4 nst = __ldg(steps + cl.x);
    Pre nv = fetch_child(P, nst.w & 3, nst.y, s, gs), nw = fetch_child(P, (nst.w >> 2) & 3, nst.z, s, gs);
    for (int i = cl.x; i < cl.y; i++) {
        const int4 st = nst;
        const Pre pv = nv, pw = nw;
        if (i + 1 < cl.y) {                              // operands of the next step: in flight during this one
            nst = __ldg(steps + i + 1);
            nv = fetch_child(P, nst.w & 3, nst.y, s, gs);
            nw = fetch_child(P, (nst.w >> 2) & 3, nst.z, s, gs);
        }
        const int u = st.x, v = st.y, w = st.z;
        const Rec rv = child_rec(st.w & 3, pv, prev, phi, prior_site);
        const Rec rw = child_rec((st.w >> 2) & 3, pw, prev, phi, prior_site);
        const Msg mv = message(rv, ld_bc(P.bc, v), qsum), mw = message(rw, ld_bc(P.bc, w), qsum);
        // x_u = E^v .* E^w  (EM_solver.py:118-133)
        const bool bv = rv.h >= 0, bw = rw.h >= 0;
        const bool same = bv && bw && rv.h == rw.h;
        double x0 = mv.E0 * mw.E0, xm = mv.Em * mw.Em;
        const double xb = mv.Eb * mw.Eb;
        double xz = bv ? mv.Eh * (same ? mw.Eh : mw.Eb) : (bw ? mv.Eb * mw.Eh : xb);
        Rec ru;
        ru.h = bv ? ((bw && !same) ? -1 : rv.h) : rw.h;      // two different hot states: all edited slots are 0
        ru.qh = bv ? rv.qh : rw.qh;
        const int e = expo_of_hi(max(hi32(x0), max(hi32(xm), hi32(xz))));
        const double sc = pow2i(-e);
        ru.x0 = x0 * sc; ru.xm = xm * sc; ru.xz = xz * sc;
        ru.cum = rv.cum + rw.cum + e;
        prev = ru;
        if (valid) {
            const size_t o = (size_t)(u - N) * P.K + s;
            P.X[o] = make_double4(ru.x0, ru.xm, ru.xz, ru.qh);
            P.XI[o] = make_int2(ru.h, ru.cum);
            if (st.w & 16) {
                // root: site likelihood L = E[root]_0 (the root's own branch starts in state 0, EM_solver.py:140)
                const Msg mr = message(ru, ld_bc(P.bc, u), qsum);
                const double L = mr.E0;
                P.invL[(size_t)cl.z * P.K + s] = (L > 0.0) ? 1.0 / L : 0.0;
                P.root_site_llh[(size_t)cl.z * P.K + s] = (L > 0.0) ? log(L) + (double)ru.cum * 0.693147180559945309417232121458 : -INFINITY;
            }
        }
    }
}

// ---------------------------------------------------------------- outside pass + expected counts
// sum of 16 per-lane values over the warp with a halving butterfly: 8 + 4 + 2 + 1 + 1 shuffles;
// afterwards lane L (L even) holds the warp total of value  8 b4 + 4 b3 + 2 b2 + b1  (bits of L).
__device__ __forceinline__ double warp_sum16(double (&v)[16], int lane) {
    const unsigned full = 0xffffffffu;
    {
        const bool up = lane & 16;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double send = up ? v[j] : v[j + 8], keep = up ? v[j + 8] : v[j];
            v[j] = keep + __shfl_xor_sync(full, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const double send = up ? v[j] : v[j + 4], keep = up ? v[j + 4] : v[j];
            v[j] = keep + __shfl_xor_sync(full, send, 8);
        }
    }
    {
        const bool up = lane & 4;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const double send = up ? v[j] : v[j + 2], keep = up ? v[j + 2] : v[j];
            v[j] = keep + __shfl_xor_sync(full, send, 4);
        }
    }
    {
        const bool up = lane & 2;
        const double send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(full, send, 2);
    }
    return v[0] + __shfl_xor_sync(full, v[0], 1);
}

// For child v of u with sibling w (EM_solver.py:143-300 in lumped full-state form):
//   T^v = out[u] .* E^w;  out[v]_0 = a T_0, out[v]_a = b q_a T_0 + c T_a, out[v]_m = g (T_0 + sum_a T_a) + T_m
//   S0 = T_0 a x_0 / L, S1 = T_0 b sum(q x) / L, S2 = T_0 g x_m / L, S3 = c sum_a T_a x_a / L, S4 = g x_m sum_a T_a / L.
// Returns the unnormalised out-record of v; cnt[0..4] the five counts (already times fac).
__device__ __forceinline__ ORec child_outside(const ORec &o, const Rec &rv, const Msg &mv, const Rec &rw, const Msg &mw,
                                              const double4 cv, double qsum, double fac, double *__restrict__ cnt) {
    const bool hotw = rw.h >= 0, hotv = rv.h >= 0, hotk = o.k >= 0;
    const double T0 = o.o0 * mw.E0, Tm = o.om * mw.Em;
    const double sqE = fma(mw.Eb, qsum, hotw ? (mw.Eh - mw.Eb) * rw.qh : 0.0);        // sum_a q_a E^w_a
    const double Ewk = (hotk && o.k == rw.h) ? mw.Eh : mw.Eb;                          // E^w at slot k
    const double dk = hotk ? o.oD * Ewk : 0.0;
    const double sT = fma(o.oB, sqE, dk);                                              // sum_a T_a
    ORec n;
    n.o0 = cv.x * T0;
    n.om = fma(cv.w, T0 + sT, Tm);
    n.oB = fma(cv.y, T0, cv.z * o.oB * mw.Eb);
    const double bq = o.oB * rw.qh;
    n.oD = hotw ? cv.z * fma(bq + ((o.k == rw.h) ? o.oD : 0.0), mw.Eh, -bq * mw.Eb) : (hotk ? cv.z * o.oD * mw.Eb : 0.0);
    n.k = hotw ? rw.h : o.k;
    n.F = 0;
    // counts of the branch above v
    const double Ewhv = (hotv && rv.h == rw.h) ? mw.Eh : mw.Eb;                        // E^w at slot h_v
    const double sqEx = fma(mv.xb, sqE, hotv ? (rv.xz - mv.xb) * rv.qh * Ewhv : 0.0);  // sum_a q_a E^w_a x^v_a
    const double xvk = (hotk && o.k == rv.h) ? rv.xz : mv.xb;                           // x^v at slot k
    const double sTx = fma(o.oB, sqEx, dk * xvk);                                      // sum_a T_a x^v_a
    const double T0f = T0 * fac, gmf = cv.w * rv.xm * fac;
    cnt[0] = T0f * cv.x * rv.x0;
    cnt[1] = T0f * cv.y * mv.sq;
    cnt[2] = T0 * gmf;
    cnt[3] = cv.z * sTx * fac;
    cnt[4] = gmf * sT;
    return n;
}

__device__ __forceinline__ void normalise_out(ORec &n, int F) {
    const int e = expo_of_hi(max(max(hi32(n.o0), hi32(n.om)), max(hi32(n.oB), hi32(n.oD))));
    const double sc = pow2i(-e);
    n.o0 *= sc; n.om *= sc; n.oB *= sc; n.oD *= sc;
    n.F = F + e;
}

template <bool POST>
__global__ void __launch_bounds__(LT) pmm_lumped_outside(LParams P, const int4 *__restrict__ clades, const int4 *__restrict__ steps) {
    const int4 cl = clades[blockIdx.x];                 // steps [x, y), root index z, clade id w
    const int N = P.n_leaves;
    const int s_raw = blockIdx.y * LT + threadIdx.x;
    const bool valid = s_raw < P.K;
    const int s = min(s_raw, P.K - 1);
    const int gs = P.site0 + s;
    const int lane = threadIdx.x & 31;
    const int grp = s_raw >> 5;                          // site group of this warp
    if (grp * 32 >= P.K) return;                         // whole warp beyond the last site
    const double *prior_site = P.prior + __ldg(P.site_ptr + gs);
    const double qsum = __ldg(P.qsum + gs);
    const double phi = __ldg(P.params + 1);
    const double invL = valid ? P.invL[(size_t)cl.z * P.K + s] : 0.0;
    double *Sg = P.Spart + (size_t)grp * P.n_nodes * 5;
    // butterfly output: which of the 10 count values this lane ends up holding
    const int vidx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);

    Rec dummy;
    dummy.x0 = dummy.xm = dummy.xz = dummy.qh = 0.0; dummy.h = -1; dummy.cum = 0;
    ORec chain;
    chain.o0 = chain.om = chain.oB = chain.oD = 0.0; chain.k = -1; chain.F = 0;
    double rt = 0.0;

    int4 nst = __ldg(steps + cl.x);
    Pre nv = fetch_child(P, (nst.w & 4) ? KIND_LEAF : KIND_LOAD, nst.y, s, gs);
    Pre nw = fetch_child(P, (nst.w & 8) ? KIND_LEAF : KIND_LOAD, nst.z, s, gs);
    double4 no = make_double4(0.0, 0.0, 0.0, 0.0);
    int2 noi = make_int2(-1, 0);
    if ((nst.w & 3) == KIND_LOAD) { const size_t o = (size_t)(nst.x - N) * P.K + s; no = P.O[o]; noi = P.OI[o]; }
    for (int i = cl.x; i < cl.y; i++) {
        const int4 st = nst;
        const Pre pv = nv, pw = nw;
        const double4 po = no;
        const int2 poi = noi;
        if (i + 1 < cl.y) {
            nst = __ldg(steps + i + 1);
            nv = fetch_child(P, (nst.w & 4) ? KIND_LEAF : KIND_LOAD, nst.y, s, gs);
            nw = fetch_child(P, (nst.w & 8) ? KIND_LEAF : KIND_LOAD, nst.z, s, gs);
            if ((nst.w & 3) == KIND_LOAD) { const size_t o = (size_t)(nst.x - N) * P.K + s; no = P.O[o]; noi = P.OI[o]; }
        }
        const int u = st.x, v = st.y, w = st.z, f = st.w;
        const bool vleaf = f & 4, wleaf = f & 8;
        const double4 cv = ld_bc(P.bc, v), cw = ld_bc(P.bc, w);
        const Rec rv = child_rec(vleaf ? KIND_LEAF : KIND_LOAD, pv, dummy, phi, prior_site);
        const Rec rw = child_rec(wleaf ? KIND_LEAF : KIND_LOAD, pw, dummy, phi, prior_site);
        const Msg mv = message(rv, cv, qsum), mw = message(rw, cw, qsum);
        ORec o;
        const int how = f & 3;
        if (how == KIND_PREV) o = chain;
        else if (how == KIND_LOAD) { o.o0 = po.x; o.om = po.y; o.oB = po.z; o.oD = po.w; o.k = poi.x; o.F = poi.y; }
        else {
            // u is a root: out[root] = (a, g, b q) (EM_solver.py:151-158); its own branch's counts (:282-286)
            const size_t ou = (size_t)(u - N) * P.K + s;
            const double4 xr = P.X[ou];
            const int2 xri = P.XI[ou];
            const double4 cr = ld_bc(P.bc, u);
            Rec ru;
            ru.x0 = xr.x; ru.xm = xr.y; ru.xz = xr.z; ru.qh = xr.w; ru.h = xri.x; ru.cum = xri.y;
            const Msg mr = message(ru, cr, qsum);
            o.o0 = cr.x; o.om = cr.w; o.oB = cr.y; o.oD = 0.0; o.k = -1; o.F = -ru.cum;
            double r16[16];
#pragma unroll
            for (int q = 0; q < 16; q++) r16[q] = 0.0;
            const double t0 = cr.x * ru.x0 * invL, t2 = cr.w * ru.xm * invL;
            r16[0] = t0; r16[1] = cr.y * mr.sq * invL; r16[2] = t2;
            const double tot = warp_sum16(r16, lane);
            if (!(lane & 1) && vidx < 5) Sg[(size_t)u * 5 + vidx] = tot;
            if (POST && valid) {
                const size_t gp = (size_t)u * P.K_total + P.site0 + s;
                P.post0[gp] = (t0 > 0.0) ? log(t0) : -INFINITY;
                P.post1[gp] = (t2 > 0.0) ? log(t2) : -INFINITY;
            }
        }
        // everything is relative to the site likelihood: fac = 2^(F_u + cum_v + cum_w) / L
        const int delta = max(-1000, min(1000, o.F + rv.cum + rw.cum));
        const double fac = invL * pow2i(delta);
        double c16[16];
#pragma unroll
        for (int q = 10; q < 16; q++) c16[q] = 0.0;
        ORec ov = child_outside(o, rv, mv, rw, mw, cv, qsum, fac, c16);
        ORec ow = child_outside(o, rw, mw, rv, mv, cw, qsum, fac, c16 + 5);
        // phi statistic: 1 - P(leaf silenced | D) on '?' entries (EM_solver.py:323)
        if (vleaf && valid && pv.ri.x == -1) rt += 1.0 - ov.om * rv.xm * fac;
        if (wleaf && valid && pw.ri.x == -1) rt += 1.0 - ow.om * rw.xm * fac;
        if (POST && valid) {
            // post0 = log P(node in state 0 | D), post1 likewise for the silenced state (EM_solver.py:276-277)
            const size_t gv = (size_t)v * P.K_total + P.site0 + s, gw = (size_t)w * P.K_total + P.site0 + s;
            const double p0v = ov.o0 * rv.x0 * fac, p1v = ov.om * rv.xm * fac;
            const double p0w = ow.o0 * rw.x0 * fac, p1w = ow.om * rw.xm * fac;
            P.post0[gv] = (p0v > 0.0) ? log(p0v) : -INFINITY; P.post1[gv] = (p1v > 0.0) ? log(p1v) : -INFINITY;
            P.post0[gw] = (p0w > 0.0) ? log(p0w) : -INFINITY; P.post1[gw] = (p1w > 0.0) ? log(p1w) : -INFINITY;
        }
        // F_child = F_u + cum_sibling + (exponent taken out of the child's record)
        if (!vleaf) normalise_out(ov, o.F + rw.cum);
        if (!wleaf) normalise_out(ow, o.F + rv.cum);
        if (valid) {
            if (!vleaf && ((f & 16) || P.store_all)) {
                const size_t oo = (size_t)(v - N) * P.K + s;
                P.O[oo] = make_double4(ov.o0, ov.om, ov.oB, ov.oD); P.OI[oo] = make_int2(ov.k, ov.F);
            }
            if (!wleaf && ((f & 32) || P.store_all)) {
                const size_t oo = (size_t)(w - N) * P.K + s;
                P.O[oo] = make_double4(ow.o0, ow.om, ow.oB, ow.oD); P.OI[oo] = make_int2(ow.k, ow.F);
            }
        }
        const int keep = (f >> 6) & 3;
        if (keep == 1) chain = ov; else if (keep == 2) chain = ow;
        // counts: sum over the warp's sites, one 40-byte write per (warp, branch)
        const double tot = warp_sum16(c16, lane);
        if (!(lane & 1) && vidx < 10) {
            const int node = (vidx < 5) ? v : w;
            Sg[(size_t)node * 5 + (vidx < 5 ? vidx : vidx - 5)] = tot;
        }
    }
    // R~ of this (warp, clade)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rt += __shfl_xor_sync(0xffffffffu, rt, o);
    if (lane == 0) P.Rpart[(size_t)grp * P.n_clades + cl.w] = rt;
}

}  // namespace pmm
