// pmm_kernels.cuh -- sm_100a kernels of the PMM engine shared by both E-step engines (branch constants,
// fixed-order reductions, the M-step) and the full-state ROW engine (`pmm_level`, `pmm_root_llh`:
// PMM_ENGINE=fullstate and the PMM_NNI_ROWS cache).  The default E-step kernels are in pmm_lumped.cuh.
//
// Model and formulas: SURVEY.md section 3.5 (full-state restatement of
// laml_libs/EM_solver.py:91-300; verified against the reference in
// tests/test_fullstate_model.py).  Design notes: DESIGN.md.
//
// ---- row engine ----
// Data layout (per site chunk): every INTERNAL node owns one row of `Wp` fp64
// slots in X (inside vectors) and one in O (outside vectors).  A site s occupies
// `len_s = roundup2(A_s + 2)` consecutive slots starting at the even offset
// `site_off[s]`, ordered [x_0 (unedited), x_m (silenced), x_1 .. x_A, (pad)].
// Sites are grouped into tiles of at most TS slots; a tile of a row is one
// contiguous 16-byte-aligned byte range that a single 1-D TMA bulk copy
// (cp.async.bulk, SASS UBLKCP) moves from HBM into shared memory.  Leaves own
// no row: their messages are closed forms of a 4-byte state code.
// Pad slots need no masking: their prior q is 0 and their outside value is 0 at
// the root and stays 0 down the tree, so they drop out of every reduction.
//
// Scaling: rows are stored UNnormalised, and every row travels with a per-site SUMMARY that its
// writer computes in the same sweep that produces the row:
//   eX[u][s]  : exponent e of the largest slot of X[u] at site s (integer max over the high
//               words -- all values are non-negative); a reader multiplies by the exact power of
//               two 2^-e, which brings the maximum into [1,2)
//   sqX[u][s] : sum_a q_a X[u]_a (raw) -- the one reduction a parent needs before it can start
//   iexp[u][s]: accumulated exponent: x_true[u] = (X[u] 2^-e) 2^iexp
//   eO[u][s]  : exponent of the largest slot of O[u] at site s
//   dexp[u][s]: D_u = oexp[u] + iexp[u] - Le  (Le = exponent of the site likelihood), which obeys
//               D_child = D_u + eO(u) - eX(child), D_root = -eX(root)
// Because the summaries exist, both level kernels make ONE pass over a site's slots: every
// coefficient of the output is known before the sweep starts, and the sweep accumulates what the
// next reader (and the expected counts) need.  Power-of-two scaling is exact: mantissas are those
// of an unscaled computation.
//
// Execution: one launch per tree level.  grid = (P, ntiles); CTA (p, t) is
// PERSISTENT over the level's nodes p, p+P, ... for tile t.  Warp NCW is the
// producer: it stages the next nodes' child rows (TMA) and their per-site
// metadata into an S-stage shared-memory ring, signalling `full` mbarriers;
// NCW consumer warps (lane groups of G=4, one site per group, double2 per
// lane => 64 contiguous bytes per group and iteration) compute and write their
// results straight to global memory with 16-byte stores, then arrive on the
// stage's `done` mbarrier.  The tile's prior q and site table stay resident in
// shared memory for the whole launch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pmm_common.cuh"

namespace pmm {

constexpr int NCW = 7;                    // consumer warps per CTA
constexpr int NTHREADS = (NCW + 1) * 32;  // + 1 producer warp
#ifndef PMM_G
#define PMM_G 4
#endif
constexpr int G = PMM_G;                  // lanes per site group
constexpr int NG = NCW * (32 / G);        // site groups per CTA
constexpr int MAX_STAGES = 4;

struct KParams {
    // chunk layout
    const int *site_off;         // [K]   slot offset of each site in a row (even)
    const int *site_len;         // [K]   padded slot count (even)
    const int *tile_site;        // [ntiles+1]
    const int *tile_off;         // [ntiles+1]
    const int *tile_psite;       // [ntiles+1] first PADDED site index of each tile (multiple of 4)
    const int *site_pidx;        // [K]   padded site index of each site
    const int2 *ent;             // [K]   per tile, sites sorted by length: x = off_in_tile | len<<16, y = site index in tile
    const double *q;             // [Wp]  prior in slot layout (0 at x_0, x_m, pads)
    const double *qsum;          // [K]   sum_a q_a
    int K, Kp, Wp, ntiles, max_tile_sites;
    // data: per-(leaf,site) state codes in the padded site index space of this chunk
    const int *codes;            // [n_leaves][Kp]
    int site0;
    // tree
    const int *left, *right, *parent, *root_index;
    int n_leaves, n_nodes;
    // parameters (device resident so that an EM loop never needs the host)
    const double *params;        // [0]=nu, [1]=phi
    const double *bc;            // [n][4] = a, b, c, g per branch
    // state
    double *X, *O;               // [n_int][Wp] inside / outside rows
    int *iexp, *dexp, *eX, *eO;  // [n_int][Kp]   (padded site index space: every tile's sites
    double *sqX;                 // [n_int][Kp]    start at a multiple of 4, so a tile's segment of
    double *rootInvL;            // [n_roots][Kp]  any of these arrays is one aligned bulk copy)
    double *root_site_llh;       // [n_roots][K]
    double *Spart;               // [ntiles][n][5]
    double *Rpart;               // [ntiles][n_int]
    double *post0, *post1;       // optional [n][K_total] log posteriors (annotation path), or null
    int K_total;
    int TS, stages;
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx_arrive(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- scalar helpers (hi32, expo_of_hi, pow2i: pmm_common.cuh)
__device__ __forceinline__ unsigned group_mask() {
    unsigned lane = threadIdx.x & 31u;
    return ((1u << G) - 1u) << (lane & ~(unsigned)(G - 1));
}
__device__ __forceinline__ double gsum(double x, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) x += __shfl_xor_sync(m, x, o);
    return x;
}
__device__ __forceinline__ unsigned gmax(unsigned x, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) x = max(x, __shfl_xor_sync(m, x, o));
    return x;
}
__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }

// Leaf state vector as a closed form of its code (EM_solver.py:103-116 in full-state
// form): x_0, x_m, and x_alpha = xb + [alpha == hot] * dx.
//   '?' (-1) -> (phi, 1, phi..)   known silenced (-2) -> e_m   0 -> (1-phi) e_0   a -> (1-phi) e_a
struct Leaf { double x0, xm, xb, dx; int hot; };   // hot = slot index inside the site (1 + a) or -1
__device__ __forceinline__ Leaf leaf_of(int code, double phi) {
    Leaf L;
    L.x0 = (code == 0) ? 1.0 - phi : ((code == -1) ? phi : 0.0);
    L.xm = (code < 0) ? 1.0 : 0.0;
    L.xb = (code == -1) ? phi : 0.0;
    L.dx = (code > 0) ? 1.0 - phi : 0.0;
    L.hot = (code > 0) ? 1 + code : -1;
    return L;
}

// ---------------------------------------------------------------- K0: branch constants
// a = p^(1+nu), b = p^nu (1-p), c = p^nu, g = 1 - p^nu  (EM_solver.py:124-133), with
// expm1 so that d -> dmin and nu -> 0 keep full precision.  fabs() keeps zeros positive:
// the scaling logic orders values by their bit patterns.
__global__ void pmm_branch_consts(const double *__restrict__ d, const double *__restrict__ params, int n,
                                  double *__restrict__ bc) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    double nu = params[0], dv = d[v];
    double c = exp(-nu * dv);
    double4 r;
    r.x = fabs(exp(-(1.0 + nu) * dv));
    r.y = fabs(c * (-expm1(-dv)));
    r.z = fabs(c);
    r.w = fabs(-expm1(-nu * dv));
    reinterpret_cast<double4 *>(bc)[v] = r;
}

// ---------------------------------------------------------------- shared-memory plan
// [ q tile : TS doubles ][ qsum : MS doubles ][ ent : MS int2 ][ barriers ][ stage 0 ][ stage 1 ] ...
// stage = [ NARR row tiles : NARR*TS doubles ][ dA, dB, dL : 3*MS doubles ][ red : NCW*12 doubles (outside) ]
//         [ mA, mB, mC, mD : 4*MS ints ][ desc : 32 ints ]
// MS = max sites per tile, rounded up to a multiple of 4.
struct Desc { int u, v, w, cls; double4 cv, cw; };   // 80 bytes (stage reserves 128); cv/cw arrive by TMA
// node classes: both children internal / v internal, w leaf / both leaves / decided per item
enum { CLS_II = 0, CLS_IL = 1, CLS_LL = 2, CLS_ANY = 3 };

template <bool OUTSIDE, int CLS> struct Plan {
    // row tiles a stage holds: the internal children's X rows (+ the node's O row for the outside pass)
    static constexpr int NINT = (CLS == CLS_II || CLS == CLS_ANY) ? 2 : (CLS == CLS_IL ? 1 : 0);
    static constexpr int NARR = NINT + (OUTSIDE ? 1 : 0);
    __host__ __device__ static size_t stage_bytes(int TS, int MS) {
        size_t b = (size_t)NARR * TS * 8;
        b += (size_t)3 * MS * 8;                       // fp64 metadata: sq_v, sq_w, 1/L
        if (OUTSIDE) b += (size_t)NCW * 12 * 8;        // per-warp partial counts of the item in this stage
        b += (size_t)4 * MS * 4 + 32 * 4;              // int32 metadata x4, item descriptor
        return (b + 127) & ~(size_t)127;
    }
    __host__ __device__ static size_t fixed_bytes(int TS, int MS) {
        // q tile, qsum, site table, barriers, (outside) one 12-double accumulator slot per lane group
        return ((size_t)TS * 8 + (size_t)MS * 8 + (size_t)MS * 8 + 2 * MAX_STAGES * 8 + (OUTSIDE ? (size_t)NG * 12 * 8 : 0) + 127) & ~(size_t)127;
    }
    __host__ __device__ static size_t total_bytes(int TS, int MS, int S) { return fixed_bytes(TS, MS) + S * stage_bytes(TS, MS); }
};

// ---------------------------------------------------------------- per-site work: inside
// x_u = E_v .* E_w with E = child message (SURVEY 3.5 "edge transfer"):
//   E_0 = a x_0 + b sum(q x) + g x_m,  E_alpha = c x_alpha + g x_m,  E_m = x_m.
// iA/iB: leaf code, or eX of an internal child; cumA/cumB, sqA/sqB: its iexp and sqX.
template <bool VI, bool WI>
__device__ __forceinline__ void inside_site(const double *__restrict__ A, const double *__restrict__ B,
                                            const double *__restrict__ sq, double qsum, int off, int len, int iA, int iB,
                                            int cumA, int cumB, double sqA, double sqB, const double4 cv, const double4 cw,
                                            double phi, double *__restrict__ gout, int *__restrict__ eX_out,
                                            int *__restrict__ iexp_out, double *__restrict__ sqX_out, int gl, unsigned gm) {
    __builtin_assume(__isShared(A)); __builtin_assume(__isShared(B)); __builtin_assume(__isShared(sq));
    // message constants: E0 (k0), Em, and E_alpha = cs * x + gmc (internal)  or  Eb + [hot] dE (leaf)
    double k0v, Emv, csv = 0.0, gmv = 0.0, Ebv = 0.0, dEv = 0.0; int cumv = 0, hv = -1;
    if (VI) {
        const double2 h = ld2(A + off);
        const double sc = pow2i(-iA);
        Emv = h.y * sc;
        gmv = cv.w * Emv;
        k0v = (cv.x * h.x + cv.y * sqA) * sc + gmv;
        csv = cv.z * sc;
        cumv = cumA;
    } else {
        const Leaf L = leaf_of(iA, phi);
        hv = L.hot;
        const double qh = (hv >= 0) ? sq[off + hv] : 0.0;
        Emv = L.xm;
        k0v = cv.x * L.x0 + cv.y * (L.xb * qsum + qh * L.dx) + cv.w * L.xm;
        Ebv = cv.z * L.xb + cv.w * L.xm;
        dEv = cv.z * L.dx;
    }
    double k0w, Emw, csw = 0.0, gmw = 0.0, Ebw = 0.0, dEw = 0.0; int cumw = 0, hw = -1;
    if (WI) {
        const double2 h = ld2(B + off);
        const double sc = pow2i(-iB);
        Emw = h.y * sc;
        gmw = cw.w * Emw;
        k0w = (cw.x * h.x + cw.y * sqB) * sc + gmw;
        csw = cw.z * sc;
        cumw = cumB;
    } else {
        const Leaf L = leaf_of(iB, phi);
        hw = L.hot;
        const double qh = (hw >= 0) ? sq[off + hw] : 0.0;
        Emw = L.xm;
        k0w = cw.x * L.x0 + cw.y * (L.xb * qsum + qh * L.dx) + cw.w * L.xm;
        Ebw = cw.z * L.xb + cw.w * L.xm;
        dEw = cw.z * L.dx;
    }
    // the single sweep: write x_u and accumulate its summary (sum q x_u, exponent of its maximum)
    double su = 0.0;
    unsigned mu = 0u;
#pragma unroll 2
    for (int i = 2 + 2 * gl; i < len; i += 2 * G) {
        const double2 q2 = ld2(sq + off + i);
        double2 ev, ew;
        if (VI) { const double2 a = ld2(A + off + i); ev.x = fma(csv, a.x, gmv); ev.y = fma(csv, a.y, gmv); }
        else { ev.x = Ebv + ((i == hv) ? dEv : 0.0); ev.y = Ebv + ((i + 1 == hv) ? dEv : 0.0); }
        if (WI) { const double2 b = ld2(B + off + i); ew.x = fma(csw, b.x, gmw); ew.y = fma(csw, b.y, gmw); }
        else { ew.x = Ebw + ((i == hw) ? dEw : 0.0); ew.y = Ebw + ((i + 1 == hw) ? dEw : 0.0); }
        const double x0 = ev.x * ew.x, x1 = ev.y * ew.y;
        st2(gout + off + i, x0, x1);
        su = fma(q2.x, x0, su); su = fma(q2.y, x1, su);
        mu = max(mu, max(hi32(x0), hi32(x1)));
    }
    const double xu0 = k0v * k0w, xum = Emv * Emw;
    su = gsum(su, gm);
    mu = gmax(max(mu, max(hi32(xu0), hi32(xum))), gm);
    if (gl == 0) {
        const int e = expo_of_hi(mu);
        st2(gout + off, xu0, xum);
        *eX_out = e;
        *iexp_out = cumv + cumw + e;
        *sqX_out = su;
    }
}

// ---------------------------------------------------------------- per-site work: outside + counts
// For node u with children v, w (EM_solver.py:143-300 in full-state form):
//   T^v = out[u] .* E^w, out[v]_0 = a T_0, out[v]_a = b q_a T_0 + c T_a, out[v]_m = g (T_0 + sum_a T_a) + T_m
//   S0 = T_0 a x_0 / L, S1 = T_0 b sum(q x) / L, S2 = T_0 g x_m / L, S3 = c sum_a T_a x_a / L, S4 = g x_m sum_a T_a / L.
// Every coefficient of out[v]_a, out[w]_a is known from the row summaries before the sweep; the
// sweep writes them and accumulates  sum o, sum o xv, sum o xw, sum o xv xw  for the counts and
// the silenced slots, plus the exponents of the new rows' maxima.
// iA/iB: leaf code, or eX of an internal child; sqA/sqB: its sqX; Du, eOu, invL: of node u.
template <bool VI, bool WI, bool POST>
__device__ __forceinline__ void outside_site(const double *__restrict__ C, const double *__restrict__ A,
                                             const double *__restrict__ B, const double *__restrict__ sq, double qsum,
                                             int off, int len, int iA, int iB, double sqA, double sqB, int Du, int eOu,
                                             double invL, const double4 cv, const double4 cw, double phi,
                                             double *__restrict__ goutv, double *__restrict__ goutw, int *__restrict__ dexp_v,
                                             int *__restrict__ dexp_w, int *__restrict__ eO_v, int *__restrict__ eO_w,
                                             double *__restrict__ post, int gl, unsigned gm, double *__restrict__ acc) {
    __builtin_assume(__isShared(A)); __builtin_assume(__isShared(B)); __builtin_assume(__isShared(C)); __builtin_assume(__isShared(sq));
    __builtin_assume(__isShared(acc));
    const double2 oh = ld2(C + off);
    const double sco = pow2i(-eOu);
    const double no0 = oh.x * sco, nom = oh.y * sco;
    // normalised child summaries: x_0, x_m, S = sum q x; hot slot (leaves)
    double nx0v, nxmv, Sv, scv = 1.0; int ev = 0, hv = -1; double xbv = 0.0, dxv = 0.0;
    if (VI) {
        const double2 h = ld2(A + off);
        ev = iA; scv = pow2i(-ev);
        nx0v = h.x * scv; nxmv = h.y * scv; Sv = sqA * scv;
    } else {
        const Leaf L = leaf_of(iA, phi);
        hv = L.hot; xbv = L.xb; dxv = L.dx;
        nx0v = L.x0; nxmv = L.xm;
        Sv = xbv * qsum + ((hv >= 0) ? sq[off + hv] : 0.0) * dxv;
    }
    double nx0w, nxmw, Sw, scw = 1.0; int ew = 0, hw = -1; double xbw = 0.0, dxw = 0.0;
    if (WI) {
        const double2 h = ld2(B + off);
        ew = iB; scw = pow2i(-ew);
        nx0w = h.x * scw; nxmw = h.y * scw; Sw = sqB * scw;
    } else {
        const Leaf L = leaf_of(iB, phi);
        hw = L.hot; xbw = L.xb; dxw = L.dx;
        nx0w = L.x0; nxmw = L.xm;
        Sw = xbw * qsum + ((hw >= 0) ? sq[off + hw] : 0.0) * dxw;
    }
    // sibling messages: E_0, E_alpha = c x + g x_m, E_m = x_m;  T^v = out[u] .* E^w, T^w = out[u] .* E^v
    const double gmv = cv.w * nxmv, gmw = cw.w * nxmw;
    const double k0v = cv.x * nx0v + cv.y * Sv + gmv;
    const double k0w = cw.x * nx0w + cw.y * Sw + gmw;
    const double Tv0 = no0 * k0w, Tvm = nom * nxmw;
    const double Tw0 = no0 * k0v, Twm = nom * nxmv;

    double so = 0, sov = 0, sow = 0, sovw = 0;
    unsigned mov = 0u, mow = 0u;
    {
        const double bv = cv.y * Tv0, bw = cw.y * Tw0;
        const double cvo = cv.z * sco, cwo = cw.z * sco;
        const double csv = cv.z * scv, csw = cw.z * scw;
        const double Ebv = cv.z * xbv + gmv, dEv = cv.z * dxv;     // leaf forms of E^v, E^w
        const double Ebw = cw.z * xbw + gmw, dEw = cw.z * dxw;
#pragma unroll 1
        for (int i = 2 + 2 * gl; i < len; i += 2 * G) {
            const double2 o = ld2(C + off + i);
            so += o.x; so += o.y;
            double2 Ev, Ew, oa = make_double2(0.0, 0.0);
            if (VI) {
                const double2 a = ld2(A + off + i);
                Ev.x = fma(csv, a.x, gmv); Ev.y = fma(csv, a.y, gmv);
                oa.x = o.x * a.x; oa.y = o.y * a.y;
                sov += oa.x; sov += oa.y;
            } else { Ev.x = Ebv + ((i == hv) ? dEv : 0.0); Ev.y = Ebv + ((i + 1 == hv) ? dEv : 0.0); }
            if (WI) {
                const double2 b = ld2(B + off + i);
                Ew.x = fma(csw, b.x, gmw); Ew.y = fma(csw, b.y, gmw);
                sow = fma(o.x, b.x, sow); sow = fma(o.y, b.y, sow);
                if (VI) { sovw = fma(oa.x, b.x, sovw); sovw = fma(oa.y, b.y, sovw); }
            } else { Ew.x = Ebw + ((i == hw) ? dEw : 0.0); Ew.y = Ebw + ((i + 1 == hw) ? dEw : 0.0); }
            if (VI || WI) {
                const double2 q2 = ld2(sq + off + i);
                if (VI) {
                    const double y0 = fma(bv, q2.x, cvo * o.x * Ew.x), y1 = fma(bv, q2.y, cvo * o.y * Ew.y);
                    st2(goutv + off + i, y0, y1);
                    mov = max(mov, max(hi32(y0), hi32(y1)));
                }
                if (WI) {
                    const double y0 = fma(bw, q2.x, cwo * o.x * Ev.x), y1 = fma(bw, q2.y, cwo * o.y * Ev.y);
                    st2(goutw + off + i, y0, y1);
                    mow = max(mow, max(hi32(y0), hi32(y1)));
                }
            }
        }
    }
    so = gsum(so, gm);
    const double So = so * sco;
    double Sov, Sow, Sovw;                            // sum_a o_a xv_a, sum_a o_a xw_a, sum_a o_a xv_a xw_a (normalised)
    if (VI) Sov = gsum(sov, gm) * sco * scv;
    else Sov = xbv * So + ((hv >= 0) ? C[off + hv] * sco : 0.0) * dxv;
    if (WI) Sow = gsum(sow, gm) * sco * scw;
    else Sow = xbw * So + ((hw >= 0) ? C[off + hw] * sco : 0.0) * dxw;
    if (VI && WI) Sovw = gsum(sovw, gm) * sco * scv * scw;
    else if (VI) Sovw = xbw * Sov + ((hw >= 0) ? dxw * (C[off + hw] * sco) * (A[off + hw] * scv) : 0.0);
    else if (WI) Sovw = xbv * Sow + ((hv >= 0) ? dxv * (C[off + hv] * sco) * (B[off + hv] * scw) : 0.0);
    else {
        const double ohv = (hv >= 0) ? C[off + hv] * sco : 0.0, ohw = (hw >= 0) ? C[off + hw] * sco : 0.0;
        Sovw = xbv * xbw * So + xbw * dxv * ohv + xbv * dxw * ohw + ((hv >= 0 && hv == hw) ? dxv * dxw * ohv : 0.0);
    }
    const double sTv = cw.z * Sow + gmw * So, sTw = cv.z * Sov + gmv * So;      // sum_a T^v_a, sum_a T^w_a
    const double sTvx = cw.z * Sovw + gmw * Sov;     // sum_a T^v_a x^v_a
    const double sTwx = cv.z * Sovw + gmv * Sow;     // sum_a T^w_a x^w_a
    const double outv0 = cv.x * Tv0, outvm = cv.w * (Tv0 + sTv) + Tvm;
    const double outw0 = cw.x * Tw0, outwm = cw.w * (Tw0 + sTw) + Twm;
    if (VI) mov = gmax(max(mov, max(hi32(outv0), hi32(outvm))), gm);
    if (WI) mow = gmax(max(mow, max(hi32(outw0), hi32(outwm))), gm);
    if (gl == 0) {
        if (VI) { st2(goutv + off, outv0, outvm); *dexp_v = Du + eOu - ev; *eO_v = expo_of_hi(mov); }
        if (WI) { st2(goutw + off, outw0, outwm); *dexp_w = Du + eOu - ew; *eO_w = expo_of_hi(mow); }
        // expected counts (EM_solver.py:279-300): everything is relative to the site likelihood
        const int delta = max(-1000, min(1000, Du + eOu));
        const double fac = scalbn(invL, delta);
        acc[0] += Tv0 * cv.x * nx0v * fac;
        acc[1] += Tv0 * cv.y * Sv * fac;
        acc[2] += Tv0 * gmv * fac;
        acc[3] += cv.z * sTvx * fac;
        acc[4] += gmv * sTv * fac;
        acc[5] += Tw0 * cw.x * nx0w * fac;
        acc[6] += Tw0 * cw.y * Sw * fac;
        acc[7] += Tw0 * gmw * fac;
        acc[8] += cw.z * sTwx * fac;
        acc[9] += gmw * sTw * fac;
        // phi statistic: 1 - P(leaf silenced | D) on '?' entries (EM_solver.py:323)
        if (!VI && iA == -1) acc[10] += 1.0 - outvm * nxmv * fac;
        if (!WI && iB == -1) acc[10] += 1.0 - outwm * nxmw * fac;
        if (POST) {
            // post0 = log P(node in state 0 | D) = log(out_0 x_0 / L); post1 likewise for the
            // silenced state (EM_solver.py:276-277); [0],[1] child v, [2],[3] child w
            const double p0v = outv0 * nx0v * fac, p1v = outvm * nxmv * fac;
            const double p0w = outw0 * nx0w * fac, p1w = outwm * nxmw * fac;
            post[0] = (p0v > 0.0) ? log(p0v) : -INFINITY;
            post[1] = (p1v > 0.0) ? log(p1v) : -INFINITY;
            post[2] = (p0w > 0.0) ? log(p0w) : -INFINITY;
            post[3] = (p1w > 0.0) ? log(p1w) : -INFINITY;
        }
    }
}

// ---------------------------------------------------------------- K1 / K2: one tree level, persistent
// items[i] = (u, v, w) with the children ordered so that an internal child comes first
// (the recursions are symmetric in the children).  CLS fixes the node class at compile
// time for the big, class-sorted launches; CLS_ANY decides per item (small levels).
template <bool OUTSIDE, int CLS, bool POST = false>
__global__ void __launch_bounds__(NTHREADS, 2)
pmm_level(KParams P, const int4 *__restrict__ items, int n_level) {
    using PL = Plan<OUTSIDE, CLS>;
    constexpr int NARR = PL::NARR;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int TS = P.TS, MS = P.max_tile_sites, S = P.stages;
    double *s_q = reinterpret_cast<double *>(smem_raw);
    double *s_qsum = s_q + TS;
    int2 *s_ent = reinterpret_cast<int2 *>(s_qsum + MS);
    uint64_t *full = reinterpret_cast<uint64_t *>(s_ent + MS);
    uint64_t *done = full + MAX_STAGES;
    double *s_gacc = reinterpret_cast<double *>(done + MAX_STAGES);   // [NG][12], outside pass only
    unsigned char *stage0 = smem_raw + PL::fixed_bytes(TS, MS);       // pure pointer arithmetic keeps the
    const size_t stage_bytes = PL::stage_bytes(TS, MS);               // shared address space visible to the compiler

    const int t = blockIdx.y;
    const int s0 = P.tile_site[t], nsites = P.tile_site[t + 1] - s0;
    const int ps0 = P.tile_psite[t];                       // padded site index of the tile's first site
    const int nsp = (nsites + 3) & ~3;
    const int off0 = P.tile_off[t];
    const int tlen = P.tile_off[t + 1] - off0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = P.n_leaves;
    const int n_items = (n_level - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    // resident tile tables
    for (int i = tid; i < tlen; i += NTHREADS) s_q[i] = P.q[off0 + i];
    for (int i = tid; i < nsites; i += NTHREADS) { s_qsum[i] = P.qsum[s0 + i]; s_ent[i] = P.ent[s0 + i]; }
    if (tid == 0) {
        for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&done[s], NCW); }
        fence_mbar_init();
    }
    __syncthreads();

    auto stage_ptr = [&](int s) { return stage0 + (size_t)s * stage_bytes; };
    // stage field accessors
    auto f_rows = [&](unsigned char *st) { return reinterpret_cast<double *>(st); };
    auto f_dmeta = [&](unsigned char *st) { return reinterpret_cast<double *>(st) + (size_t)NARR * TS; };   // dA, dB, dL
    auto f_red = [&](unsigned char *st) { return f_dmeta(st) + 3 * MS; };
    auto f_meta = [&](unsigned char *st) { return reinterpret_cast<int *>(f_red(st) + (OUTSIDE ? NCW * 12 : 0)); };   // mA..mD
    auto f_desc = [&](unsigned char *st) { return reinterpret_cast<Desc *>(f_meta(st) + 4 * MS); };

    if (warp == NCW) {
        // ======================= producer warp: everything it moves is an async bulk copy =======================
        auto flush_counts = [&](unsigned char *st) {
            // deterministic cross-warp sum of the per-warp partial counts of a finished item
            const Desc *dsc = f_desc(st);
            const int du = dsc->u, dv = dsc->v, dw = dsc->w;
            const double *red = f_red(st);
            if (lane < 11) {
                double sum = 0.0;
                for (int wdx = 0; wdx < NCW; wdx++) sum += red[wdx * 12 + lane];
                if (lane < 5) P.Spart[((size_t)t * P.n_nodes + dv) * 5 + lane] = sum;
                else if (lane < 10) P.Spart[((size_t)t * P.n_nodes + dw) * 5 + (lane - 5)] = sum;
                else P.Rpart[(size_t)t * (P.n_nodes - N) + (du - N)] = sum;
            }
            __syncwarp();
        };
        int4 nxt = n_items > 0 ? items[blockIdx.x] : make_int4(0, 0, 0, 0);
        for (int k = 0; k < n_items; k++) {
            const int s = k % S;
            unsigned char *st = stage_ptr(s);
            const int4 it = nxt;
            if (k + 1 < n_items) nxt = items[blockIdx.x + (size_t)(k + 1) * gridDim.x];     // ahead of need
            if (k >= S) {
                mbar_wait(&done[s], ((k / S) - 1) & 1);          // consumers have finished item k-S
                if (OUTSIDE) flush_counts(st);
            }
            if (lane == 0) {
                const int u = it.x, v = it.y, w = it.z;
                const bool vi = v >= N, wi = w >= N;
                double *rows = f_rows(st);
                int *mA = f_meta(st), *mB = mA + MS, *mC = mB + MS, *mD = mC + MS;
                double *dA = f_dmeta(st), *dB = dA + MS, *dL = dB + MS;
                Desc *dsc = f_desc(st);
                dsc->u = u; dsc->v = v; dsc->w = w; dsc->cls = vi ? (wi ? CLS_II : CLS_IL) : CLS_LL;
                const uint32_t rb = (uint32_t)tlen * 8u, ib = (uint32_t)nsp * 4u, db = (uint32_t)nsp * 8u;
                const uint32_t nint = (vi ? 1u : 0u) + (wi ? 1u : 0u);
                uint32_t tx = 64u + rb * (nint + (OUTSIDE ? 1u : 0u));             // cv, cw + row tiles
                // internal child: eX (+ iexp for the inside pass) and sqX; leaf child: its codes
                tx += nint * (ib + db + (OUTSIDE ? 0u : ib)) + (2u - nint) * ib;
                if (OUTSIDE) tx += 2u * ib + db;                                   // dexp, eO, 1/L of node u
                mbar_expect_tx_arrive(&full[s], tx);                               // release: orders the id stores above
                bulk_load(&dsc->cv, P.bc + (size_t)4 * v, 32u, &full[s]);
                bulk_load(&dsc->cw, P.bc + (size_t)4 * w, 32u, &full[s]);
                double *rowA = rows + (OUTSIDE ? TS : 0), *rowB = rowA + TS;
                if (vi) {
                    const size_t mo = (size_t)(v - N) * P.Kp + ps0;
                    bulk_load(rowA, P.X + (size_t)(v - N) * P.Wp + off0, rb, &full[s]);
                    bulk_load(mA, P.eX + mo, ib, &full[s]);
                    bulk_load(dA, P.sqX + mo, db, &full[s]);
                    if (!OUTSIDE) bulk_load(mC, P.iexp + mo, ib, &full[s]);
                } else bulk_load(mA, P.codes + (size_t)v * P.Kp + ps0, ib, &full[s]);
                if (wi) {
                    const size_t mo = (size_t)(w - N) * P.Kp + ps0;
                    bulk_load(rowB, P.X + (size_t)(w - N) * P.Wp + off0, rb, &full[s]);
                    bulk_load(mB, P.eX + mo, ib, &full[s]);
                    bulk_load(dB, P.sqX + mo, db, &full[s]);
                    if (!OUTSIDE) bulk_load(mD, P.iexp + mo, ib, &full[s]);
                } else bulk_load(mB, P.codes + (size_t)w * P.Kp + ps0, ib, &full[s]);
                if (OUTSIDE) {
                    const size_t mo = (size_t)(u - N) * P.Kp + ps0;
                    bulk_load(rows, P.O + (size_t)(u - N) * P.Wp + off0, rb, &full[s]);
                    bulk_load(mC, P.dexp + mo, ib, &full[s]);
                    bulk_load(mD, P.eO + mo, ib, &full[s]);
                    bulk_load(dL, P.rootInvL + (size_t)it.w * P.Kp + ps0, db, &full[s]);
                }
            }
            __syncwarp();
        }
        if (OUTSIDE) {
            // drain: partial counts of the last min(S, n_items) items
            for (int k = max(0, n_items - S); k < n_items; k++) {
                const int s = k % S;
                mbar_wait(&done[s], (k / S) & 1);
                flush_counts(stage_ptr(s));
            }
        }
        return;
    }

    // ======================= consumer warps =======================
    const unsigned gm = group_mask();
    const int gl = lane & (G - 1);
    const int grp = warp * (32 / G) + lane / G;
    const double phi = P.params[1];
    for (int k = 0; k < n_items; k++) {
        const int s = k % S;
        unsigned char *st = stage_ptr(s);
        mbar_wait(&full[s], (k / S) & 1);
        const Desc *dsc = f_desc(st);
        const int u = dsc->u, v = dsc->v, w = dsc->w;
        const int cls = (CLS == CLS_ANY) ? dsc->cls : CLS;
        const double4 cv = dsc->cv, cw = dsc->cw;
        const double *C = f_rows(st), *A = C + (OUTSIDE ? TS : 0), *B = A + TS;
        const int *mA = f_meta(st), *mB = mA + MS, *mC = mB + MS, *mD = mC + MS;
        const double *dA = f_dmeta(st), *dB = dA + MS, *dL = dB + MS;
        if (!OUTSIDE) {
            const size_t mo = (size_t)(u - N) * P.Kp + ps0;
            double *gout = P.X + (size_t)(u - N) * P.Wp + off0;
            int *eX_u = P.eX + mo, *iexp_u = P.iexp + mo;
            double *sqX_u = P.sqX + mo;
            for (int j = grp; j < nsites; j += NG) {
                const int2 e = s_ent[j];
                const int off = e.x & 0xffff, len = e.x >> 16, ls = e.y;
                if (cls == CLS_II) inside_site<true, true>(A, B, s_q, s_qsum[ls], off, len, mA[ls], mB[ls], mC[ls], mD[ls], dA[ls], dB[ls], cv, cw, phi, gout, eX_u + ls, iexp_u + ls, sqX_u + ls, gl, gm);
                else if (cls == CLS_IL) inside_site<true, false>(A, B, s_q, s_qsum[ls], off, len, mA[ls], mB[ls], mC[ls], 0, dA[ls], 0.0, cv, cw, phi, gout, eX_u + ls, iexp_u + ls, sqX_u + ls, gl, gm);
                else inside_site<false, false>(A, B, s_q, s_qsum[ls], off, len, mA[ls], mB[ls], 0, 0, 0.0, 0.0, cv, cw, phi, gout, eX_u + ls, iexp_u + ls, sqX_u + ls, gl, gm);
            }
        } else {
            const bool vi = cls != CLS_LL, wi = cls == CLS_II;
            double *goutv = vi ? P.O + (size_t)(v - N) * P.Wp + off0 : nullptr;
            double *goutw = wi ? P.O + (size_t)(w - N) * P.Wp + off0 : nullptr;
            int *dv = vi ? P.dexp + (size_t)(v - N) * P.Kp + ps0 : nullptr;
            int *dw = wi ? P.dexp + (size_t)(w - N) * P.Kp + ps0 : nullptr;
            int *eov = vi ? P.eO + (size_t)(v - N) * P.Kp + ps0 : nullptr;
            int *eow = wi ? P.eO + (size_t)(w - N) * P.Kp + ps0 : nullptr;
            // expected-count accumulators live in shared memory (one 12-double slot per lane group):
            // keeping 11 fp64 accumulators in registers across the site sweep costs spills
            double *red = f_red(st);
            double *acc = s_gacc + (size_t)grp * 12;
            if (gl == 0) {
#pragma unroll
                for (int q = 0; q < 11; q++) acc[q] = 0.0;
            }
            for (int j = grp; j < nsites; j += NG) {
                const int2 e = s_ent[j];
                const int off = e.x & 0xffff, len = e.x >> 16, ls = e.y;
                double pst[4];
                double *post = pst;
                if (cls == CLS_II) outside_site<true, true, POST>(C, A, B, s_q, s_qsum[ls], off, len, mA[ls], mB[ls], dA[ls], dB[ls], mC[ls], mD[ls], dL[ls], cv, cw, phi, goutv, goutw, dv + ls, dw + ls, eov + ls, eow + ls, post, gl, gm, acc);
                else if (cls == CLS_IL) outside_site<true, false, POST>(C, A, B, s_q, s_qsum[ls], off, len, mA[ls], mB[ls], dA[ls], 0.0, mC[ls], mD[ls], dL[ls], cv, cw, phi, goutv, goutw, dv + ls, dw, eov + ls, eow, post, gl, gm, acc);
                else outside_site<false, false, POST>(C, A, B, s_q, s_qsum[ls], off, len, mA[ls], mB[ls], 0.0, 0.0, mC[ls], mD[ls], dL[ls], cv, cw, phi, goutv, goutw, dv, dw, eov, eow, post, gl, gm, acc);
                if (POST && gl == 0) {
                    const size_t gs = (size_t)P.site0 + s0 + ls;
                    P.post0[(size_t)v * P.K_total + gs] = pst[0]; P.post1[(size_t)v * P.K_total + gs] = pst[1];
                    P.post0[(size_t)w * P.K_total + gs] = pst[2]; P.post1[(size_t)w * P.K_total + gs] = pst[3];
                }
            }
            // warp-level sum over the warp's groups in a fixed order
            __syncwarp();
            if (lane < 11) {
                const double *g0 = s_gacc + (size_t)(warp * (32 / G)) * 12 + lane;
                double x = 0.0;
#pragma unroll
                for (int g = 0; g < 32 / G; g++) x += g0[g * 12];
                red[warp * 12 + lane] = x;
            }
            __syncwarp();      // the group slots are zeroed again at the start of the warp's next item
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[s]);
    }
}

// codes_p[r][site_pidx[s]] = raw[r][site0 + s]   (padded site index space of one chunk)
__global__ void pmm_pad_codes(const int *__restrict__ raw, int raw_stride, int site0, int K, const int *__restrict__ site_pidx,
                              int Kp, long long n_rows, int *__restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * K) return;
    long long r = i / K;
    int s = (int)(i - r * K);
    out[r * Kp + site_pidx[s]] = raw[r * raw_stride + site0 + s];
}

// ---------------------------------------------------------------- K1r: root likelihood
// One CTA per (root, tile): site likelihood L_s = E[root]_0 (the root's own branch
// starts in state 0, EM_solver.py:140), its log, the expected counts of the root
// branch (EM_solver.py:282-286), and the seeds of the outside pass: out[root]
// (EM_solver.py:151-158), D_root and 1/L.
__global__ void __launch_bounds__(256) pmm_root_llh(KParams P, const int *__restrict__ roots) {
    constexpr int RG = 256 / G;
    __shared__ double red[RG][3];
    const int t = blockIdx.x % P.ntiles;
    const int ri = blockIdx.x / P.ntiles;
    const int r = roots[ri];
    const int N = P.n_leaves;
    const int s0 = P.tile_site[t], s1 = P.tile_site[t + 1];
    const int tid = threadIdx.x;
    const unsigned gm = group_mask();
    const int gl = tid & (G - 1), grp = tid / G;
    const double4 cr = reinterpret_cast<const double4 *>(P.bc)[r];
    const size_t mo = (size_t)(r - N) * P.Kp;
    const double *x = P.X + (size_t)(r - N) * P.Wp;
    double *o = P.O + (size_t)(r - N) * P.Wp;
    double acc0 = 0, acc1 = 0, acc2 = 0;
    for (int s = s0 + grp; s < s1; s += RG) {
        const int goff = P.site_off[s], len = P.site_len[s];
        unsigned mx = 0u;
        for (int i = 2 + gl; i < len; i += G) {              // out[root]_alpha = b q_alpha
            const double y = cr.y * P.q[goff + i];
            o[goff + i] = y;
            mx = max(mx, hi32(y));
        }
        mx = gmax(max(mx, max(hi32(cr.x), hi32(cr.w))), gm);
        if (gl == 0) {
            const int ps = P.site_pidx[s];
            const int e = P.eX[mo + ps];
            const double sc = pow2i(-e);
            const double t0 = cr.x * x[goff] * sc, t1 = cr.y * P.sqX[mo + ps] * sc, t2 = cr.w * x[goff + 1] * sc;
            const double L = t0 + t1 + t2;
            const int Le = P.iexp[mo + ps];                  // accumulated exponent incl. e
            const double inv = (L > 0.0) ? 1.0 / L : 0.0;
            o[goff] = cr.x; o[goff + 1] = cr.w;
            P.eO[mo + ps] = expo_of_hi(mx);
            P.dexp[mo + ps] = -e;
            P.rootInvL[(size_t)ri * P.Kp + ps] = inv;
            P.root_site_llh[(size_t)ri * P.K + s] = (L > 0.0) ? log(L) + (double)Le * 0.693147180559945309417232121458 : -INFINITY;
            acc0 += t0 * inv; acc1 += t1 * inv; acc2 += t2 * inv;
            if (P.post0) {
                P.post0[(size_t)r * P.K_total + P.site0 + s] = (t0 * inv > 0.0) ? log(t0 * inv) : -INFINITY;
                P.post1[(size_t)r * P.K_total + P.site0 + s] = (t2 * inv > 0.0) ? log(t2 * inv) : -INFINITY;
            }
        }
    }
    if (gl == 0) { red[grp][0] = acc0; red[grp][1] = acc1; red[grp][2] = acc2; }
    __syncthreads();
    if (tid < 5) {
        double sum = 0.0;
        if (tid < 3) for (int g = 0; g < RG; g++) sum += red[g][tid];
        P.Spart[((size_t)t * P.n_nodes + r) * 5 + tid] = sum;
    }
}

// ---------------------------------------------------------------- reductions
// S[v][k] (+)= sum over tiles of Spart[t][v][k]; deterministic order.
__global__ void pmm_reduce_S(const double *__restrict__ Spart, int ntiles, int n5, double *__restrict__ S, int accumulate) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n5) return;
    double sum = 0.0;
    for (int t = 0; t < ntiles; t++) sum += Spart[(size_t)t * n5 + i];
    S[i] = accumulate ? S[i] + sum : sum;
}

// out[0] (+)= sum of in[0..n): single CTA, fixed order => bit-reproducible
__global__ void pmm_reduce_sum(const double *__restrict__ in, long long n, double *__restrict__ out, int accumulate) {
    __shared__ double sh[1024];
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) s += in[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = accumulate ? out[0] + sh[0] : sh[0];
}

// site_llh[site0+s] = sum over roots of root_site_llh[r][s]
__global__ void pmm_site_llh(const double *__restrict__ root_site_llh, int n_roots, int K, double *__restrict__ site_llh) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= K) return;
    double sum = 0.0;
    for (int r = 0; r < n_roots; r++) sum += root_site_llh[(size_t)r * K + s];
    site_llh[s] = sum;
}

// Everything that follows the passes of the lumped engine, in ONE launch (five before): blocks 0 .. nb-1 sum the count
// partials over the site groups (pmm_reduce_S), the last block forms the per-site log-likelihoods and their sum
// (pmm_site_llh + pmm_reduce_sum) and, after an outside pass, R~ from the per-(warp, clade) partials (pmm_reduce_S +
// pmm_reduce_sum).  Same summation orders as the separate kernels (1024-thread strided sums, then a tree): same bits.
__global__ void __launch_bounds__(1024) pmm_lumped_finish(const double *__restrict__ Spart, int G, int n5, double *__restrict__ S, int accumulate,
                                                          const double *__restrict__ root_site_llh, int n_roots, int K,
                                                          double *__restrict__ site_llh, double *__restrict__ llh_out,
                                                          const double *__restrict__ Rpart, int n_clades, double *__restrict__ Rnode,
                                                          double *__restrict__ Rt_out, int with_counts, const int *__restrict__ df_err) {
    __shared__ double sh[1024];
    const int tid = threadIdx.x;
    if (blockIdx.x + 1 < gridDim.x) {
        if (!with_counts) return;
        const int i = blockIdx.x * 1024 + tid;
        if (i >= n5) return;
        double sum = 0.0;
        for (int t = 0; t < G; t++) sum += Spart[(size_t)t * n5 + i];
        S[i] = accumulate ? S[i] + sum : sum;
        return;
    }
    auto block_sum = [&](double v) {
        sh[tid] = v;
        __syncthreads();
        for (int o = 512; o > 0; o >>= 1) {
            if (tid < o) sh[tid] += sh[tid + o];
            __syncthreads();
        }
        const double r = sh[0];
        __syncthreads();
        return r;
    };
    double acc = 0.0;
    for (int s = tid; s < K; s += 1024) {
        double sum = 0.0;
        for (int r = 0; r < n_roots; r++) sum += root_site_llh[(size_t)r * K + s];
        site_llh[s] = sum;
        acc += sum;
    }
    double llh = block_sum(acc);
    if (df_err && *df_err) llh = NAN;            // a dataflow launch gave up waiting (pmm_lumped.cuh: df_spin): no result, loudly
    if (tid == 0) llh_out[0] = accumulate ? llh_out[0] + llh : llh;
    if (!with_counts) return;
    acc = 0.0;
    for (int c = tid; c < n_clades; c += 1024) {
        double sum = 0.0;
        for (int t = 0; t < G; t++) sum += Rpart[(size_t)t * n_clades + c];
        Rnode[c] = sum;
        acc += sum;
    }
    const double rt = block_sum(acc);
    if (tid == 0) Rt_out[0] = accumulate ? Rt_out[0] + rt : rt;
}

// ---------------------------------------------------------------- K3: M-step
// stats = [S(n*5) | Rtilde | llh].  EM_solver.py:318-348 (phi, clamp, rescale),
// :351-357 (per-branch concave problem in d), :394-407 (1-D problem in nu).
// dobj_dd, mstep_end_const, mstep_branch_length: pmm_common.cuh

__global__ void __launch_bounds__(128) pmm_mstep_branches(const double *__restrict__ stats, const unsigned char *__restrict__ free_mask, int n,
                                                          int K_total, double *__restrict__ params, double *__restrict__ d,
                                                          double *__restrict__ Snorm, double dmin, double dmax, double eps_s, double eps_nu) {
    __shared__ double s_end[4];
    const double nu = params[0];
    const bool use_nu = nu > eps_nu;
    if (threadIdx.x < 4) s_end[threadIdx.x] = mstep_end_const(threadIdx.x, nu, use_nu, dmin, dmax);
    __syncthreads();
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n || !free_mask[v]) return;
    double s[5], tot = 0.0;
#pragma unroll
    for (int k = 0; k < 5; k++) { s[k] = fmax(stats[(size_t)v * 5 + k], eps_s); }
#pragma unroll
    for (int k = 0; k < 5; k++) tot += s[k];
#pragma unroll
    for (int k = 0; k < 5; k++) { s[k] = s[k] / tot * (double)K_total; Snorm[(size_t)v * 5 + k] = s[k]; }
    d[v] = mstep_branch_length(s, nu, use_nu, dmin, dmax, d[v], s_end);
}

// ---- the 1-D problem in nu (EM_solver.py:394-407) and phi (:318-326)
// g'(nu) = -sum (S0+S1+S3) d + sum (S2+S4) d / expm1(nu d) is convex and strictly decreasing on (0,1],
// +inf at 0: safeguarded Newton (bracket kept, bisection fallback).  Each evaluation is a grid-wide
// two-stage reduction in a fixed order (bit-reproducible, identical on every rank): NU_BLOCKS CTAs
// write partial sums, one CTA combines them and takes the Newton step.  The host enqueues a fixed
// number of (eval, step) pairs; once converged the state is a fixed point of the step.
constexpr int NU_BLOCKS = 256;
struct NuState { double x, lo, hi, lin, wsum; int done, iters; };

// partial[b] = { sum w d / expm1(x d),  d/dx of it,  sum (S0+S1+S3) d,  sum w }   over this CTA's branches
__global__ void __launch_bounds__(256) pmm_nu_eval(const unsigned char *__restrict__ free_mask, int n,
                                                    const double *__restrict__ Snorm, const double *__restrict__ d,
                                                    const NuState *__restrict__ st, int first, double *__restrict__ partial) {
    __shared__ double sh[4][256];
    const int tid = threadIdx.x;
    if (!first && st->done) return;
    const double x = first ? 1.0 : st->x;
    double f = 0.0, fp = 0.0, lin = 0.0, ws = 0.0;
    for (int v = blockIdx.x * 256 + tid; v < n; v += NU_BLOCKS * 256)
        if (free_mask[v]) {
            const double dv = d[v], w = Snorm[(size_t)v * 5 + 2] + Snorm[(size_t)v * 5 + 4];
            const double em = expm1(x * dv);
            const double r = dv / em;
            f += w * r;
            fp -= w * r * r * (em + 1.0);
            if (first) { lin += (Snorm[(size_t)v * 5] + Snorm[(size_t)v * 5 + 1] + Snorm[(size_t)v * 5 + 3]) * dv; ws += w; }
        }
    sh[0][tid] = f; sh[1][tid] = fp; sh[2][tid] = lin; sh[3][tid] = ws;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) {
#pragma unroll
            for (int k = 0; k < 4; k++) sh[k][tid] += sh[k][tid + o];
        }
        __syncthreads();
    }
    if (tid < 4) partial[blockIdx.x * 4 + tid] = sh[tid][0];
}

__global__ void __launch_bounds__(NU_BLOCKS) pmm_nu_step(const double *__restrict__ partial, NuState *__restrict__ st, int first,
                                                          double *__restrict__ params) {
    __shared__ double sh[4][NU_BLOCKS];
    const int tid = threadIdx.x;
    if (!first && st->done) return;
#pragma unroll
    for (int k = 0; k < 4; k++) sh[k][tid] = partial[tid * 4 + k];
    __syncthreads();
    for (int o = NU_BLOCKS / 2; o > 0; o >>= 1) {
        if (tid < o) {
#pragma unroll
            for (int k = 0; k < 4; k++) sh[k][tid] += sh[k][tid + o];
        }
        __syncthreads();
    }
    if (tid != 0) return;
    const double f = sh[0][0], fp = sh[1][0];
    NuState s = *st;
    if (first) {
        s.lin = sh[2][0]; s.wsum = sh[3][0]; s.iters = 0; s.done = 0;
        if (f - s.lin >= 0.0 || !(s.lin > 0.0)) { s.x = 1.0; s.done = 1; }        // upper bound active (EM_solver.py:402)
        else { s.lo = 0.0; s.hi = 1.0; s.x = fmin(0.5, s.wsum / s.lin); }           // 1/expm1(t) < 1/t  =>  root < wsum/lin
    } else {
        const double g = f - s.lin, x = s.x;
        if (g > 0.0) s.lo = x; else s.hi = x;
        s.iters++;
        if (g == 0.0) s.done = 1;
        else {
            double xn = x - g / fp;
            if (fabs(xn - x) <= 4e-16 * x) { s.done = 1; xn = fmin(fmax(xn, s.lo), s.hi); }
            else if (!(xn > s.lo && xn < s.hi)) xn = (s.lo > 0.0) ? 0.5 * (s.lo + s.hi) : 0.5 * s.hi;
            if (xn < 1e-300) { xn = 0.0; s.done = 1; }
            s.x = xn;
        }
    }
    *st = s;
    params[0] = s.x;
}

// The same Newton iteration in ONE cooperative launch: every CTA evaluates its branches, a grid
// barrier (monotonic counter) publishes the partial sums, and every CTA combines them in the same
// fixed order and takes the same step -- one barrier per iteration instead of two launches.
// partial: [2][gridDim.x][4] (double-buffered across iterations); bar: zeroed before the launch.
__global__ void __launch_bounds__(256) pmm_nu_solve(const unsigned char *__restrict__ free_mask, int n, const double *__restrict__ Snorm,
                                                     const double *__restrict__ d, double *__restrict__ partial, unsigned *bar,
                                                     NuState *__restrict__ st_out, double *__restrict__ params, int max_iter) {
    __shared__ double sh[4][256];
    const int tid = threadIdx.x, G = gridDim.x;
    NuState s;
    s.x = 1.0; s.lo = 0.0; s.hi = 1.0; s.lin = 0.0; s.wsum = 0.0; s.done = 0; s.iters = 0;
    const double nu_cur = params[0];         // read by every CTA before the first barrier; written after the last one
    for (int it = 0; it <= max_iter; it++) {
        const bool first = it == 0;
        const double x = s.x;
        double f = 0.0, fp = 0.0, lin = 0.0, ws = 0.0;
        for (int v = blockIdx.x * 256 + tid; v < n; v += G * 256)
            if (free_mask[v]) {
                const double dv = d[v], w = Snorm[(size_t)v * 5 + 2] + Snorm[(size_t)v * 5 + 4];
                const double em = expm1(x * dv);
                const double r = dv / em;
                f += w * r;
                fp -= w * r * r * (em + 1.0);
                if (first) { lin += (Snorm[(size_t)v * 5] + Snorm[(size_t)v * 5 + 1] + Snorm[(size_t)v * 5 + 3]) * dv; ws += w; }
            }
        sh[0][tid] = f; sh[1][tid] = fp; sh[2][tid] = lin; sh[3][tid] = ws;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (tid < o) {
#pragma unroll
                for (int k = 0; k < 4; k++) sh[k][tid] += sh[k][tid + o];
            }
            __syncthreads();
        }
        double *pb = partial + (size_t)(it & 1) * G * 4;
        if (tid < 4) pb[blockIdx.x * 4 + tid] = sh[tid][0];
        // ---- grid barrier
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(bar, 1u);
            const unsigned target = (unsigned)(it + 1) * (unsigned)G;
            while (*((volatile unsigned *)bar) < target) {}
            __threadfence();
        }
        __syncthreads();
        // ---- every CTA: fixed-order sum of the G partials (G <= 256), same step everywhere
#pragma unroll
        for (int k = 0; k < 4; k++) sh[k][tid] = (tid < G) ? __ldcg(pb + tid * 4 + k) : 0.0;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (tid < o) {
#pragma unroll
                for (int k = 0; k < 4; k++) sh[k][tid] += sh[k][tid + o];
            }
            __syncthreads();
        }
        const double F = sh[0][0], FP = sh[1][0];
        if (first) {
            s.lin = sh[2][0]; s.wsum = sh[3][0];
            if (F - s.lin >= 0.0 || !(s.lin > 0.0)) { s.x = 1.0; s.done = 1; }          // upper bound active (EM_solver.py:402)
            else {
                // start from the current nu (EM moves it little); else 1/expm1(t) < 1/t  =>  root < wsum/lin
                s.lo = 0.0; s.hi = 1.0;
                s.x = (nu_cur > 0.0 && nu_cur < 1.0) ? nu_cur : fmin(0.5, s.wsum / s.lin);
            }
        } else {
            const double g = F - s.lin, xo = s.x;
            if (g > 0.0) s.lo = xo; else s.hi = xo;
            s.iters++;
            if (g == 0.0) s.done = 1;
            else {
                double xn = xo - g / FP;
                if (fabs(xn - xo) <= 4e-16 * xo) { s.done = 1; xn = fmin(fmax(xn, s.lo), s.hi); }     // the step itself first (see mstep_branch_length)
                else if (!(xn > s.lo && xn < s.hi)) xn = (s.lo > 0.0) ? 0.5 * (s.lo + s.hi) : 0.5 * s.hi;
                if (xn < 1e-300) { xn = 0.0; s.done = 1; }
                s.x = xn;
            }
        }
        __syncthreads();
        if (s.done) break;          // identical on every CTA: all leave together
    }
    if (blockIdx.x == 0 && tid == 0) { *st_out = s; params[0] = s.x; }
}

// The whole unconstrained M-step in ONE cooperative launch (four launches and a memset before): phase 0 = pmm_mstep_branches
// grid-strided (every thread later re-reads, in the nu phase, exactly the branches it wrote), a grid barrier, the Newton
// iteration of pmm_nu_solve, phi by one thread.  The barrier counter is reset by the last CTA to leave, so the next launch
// needs no memset.  bar[0] = arrivals, bar[1] = CTAs that have left.
__global__ void __launch_bounds__(256) pmm_mstep_fused(const double *__restrict__ stats, const unsigned char *__restrict__ free_mask, int n,
                                                        int K_total, double *__restrict__ params, double *__restrict__ d,
                                                        double *__restrict__ Snorm, double dmin, double dmax, double eps_s, double eps_nu,
                                                        int do_branches, int do_nu, int do_phi, double R, double cells_times_sites,
                                                        double *__restrict__ partial, unsigned *bar, NuState *__restrict__ st_out, int max_iter) {
    __shared__ double sh[4][256];
    __shared__ double s_end[4];
    const int tid = threadIdx.x, G = gridDim.x;
    const double nu_cur = params[0];         // read by every CTA before the first barrier; written after the last one
    unsigned epoch = 0;
    auto grid_barrier = [&]() {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(bar, 1u);
            const unsigned target = (epoch + 1u) * (unsigned)G;
            while (*((volatile unsigned *)bar) < target) {}
            __threadfence();
        }
        epoch++;
        __syncthreads();
    };
    if (do_branches) {
        const bool use_nu = nu_cur > eps_nu;
        if (tid < 4) s_end[tid] = mstep_end_const(tid, nu_cur, use_nu, dmin, dmax);
        __syncthreads();
        for (int v = blockIdx.x * 256 + tid; v < n; v += G * 256) {
            if (!free_mask[v]) continue;
            double s[5], tot = 0.0;
#pragma unroll
            for (int k = 0; k < 5; k++) { s[k] = fmax(stats[(size_t)v * 5 + k], eps_s); }
#pragma unroll
            for (int k = 0; k < 5; k++) tot += s[k];
#pragma unroll
            for (int k = 0; k < 5; k++) { s[k] = s[k] / tot * (double)K_total; Snorm[(size_t)v * 5 + k] = s[k]; }
            d[v] = mstep_branch_length(s, nu_cur, use_nu, dmin, dmax, d[v], s_end);
        }
    }
    NuState s;
    s.x = 1.0; s.lo = 0.0; s.hi = 1.0; s.lin = 0.0; s.wsum = 0.0; s.done = 0; s.iters = 0;
    if (do_nu) {
        for (int it = 0; it <= max_iter; it++) {
            const bool first = it == 0;
            const double x = s.x;
            double f = 0.0, fp = 0.0, lin = 0.0, ws = 0.0;
            for (int v = blockIdx.x * 256 + tid; v < n; v += G * 256)
                if (free_mask[v]) {
                    const double dv = d[v], w = Snorm[(size_t)v * 5 + 2] + Snorm[(size_t)v * 5 + 4];
                    const double em = expm1(x * dv);
                    const double r = dv / em;
                    f += w * r;
                    fp -= w * r * r * (em + 1.0);
                    if (first) { lin += (Snorm[(size_t)v * 5] + Snorm[(size_t)v * 5 + 1] + Snorm[(size_t)v * 5 + 3]) * dv; ws += w; }
                }
            sh[0][tid] = f; sh[1][tid] = fp; sh[2][tid] = lin; sh[3][tid] = ws;
            __syncthreads();
            for (int o = 128; o > 0; o >>= 1) {
                if (tid < o) {
#pragma unroll
                    for (int k = 0; k < 4; k++) sh[k][tid] += sh[k][tid + o];
                }
                __syncthreads();
            }
            double *pb = partial + (size_t)(it & 1) * G * 4;
            if (tid < 4) pb[blockIdx.x * 4 + tid] = sh[tid][0];
            grid_barrier();
            // every CTA: fixed-order sum of the G partials (G <= 256), same step everywhere
#pragma unroll
            for (int k = 0; k < 4; k++) sh[k][tid] = (tid < G) ? __ldcg(pb + tid * 4 + k) : 0.0;
            __syncthreads();
            for (int o = 128; o > 0; o >>= 1) {
                if (tid < o) {
#pragma unroll
                    for (int k = 0; k < 4; k++) sh[k][tid] += sh[k][tid + o];
                }
                __syncthreads();
            }
            const double F = sh[0][0], FP = sh[1][0];
            if (first) {
                s.lin = sh[2][0]; s.wsum = sh[3][0];
                if (F - s.lin >= 0.0 || !(s.lin > 0.0)) { s.x = 1.0; s.done = 1; }          // upper bound active (EM_solver.py:402)
                else {
                    s.lo = 0.0; s.hi = 1.0;
                    s.x = (nu_cur > 0.0 && nu_cur < 1.0) ? nu_cur : fmin(0.5, s.wsum / s.lin);
                }
            } else {
                const double g = F - s.lin, xo = s.x;
                if (g > 0.0) s.lo = xo; else s.hi = xo;
                s.iters++;
                if (g == 0.0) s.done = 1;
                else {
                    double xn = xo - g / FP;
                    if (fabs(xn - xo) <= 4e-16 * xo) { s.done = 1; xn = fmin(fmax(xn, s.lo), s.hi); }
                    else if (!(xn > s.lo && xn < s.hi)) xn = (s.lo > 0.0) ? 0.5 * (s.lo + s.hi) : 0.5 * s.hi;
                    if (xn < 1e-300) { xn = 0.0; s.done = 1; }
                    s.x = xn;
                }
            }
            __syncthreads();
            if (s.done) break;          // identical on every CTA: all leave together
        }
    }
    if (blockIdx.x == 0 && tid == 0) {
        if (do_nu) { *st_out = s; params[0] = s.x; }
        if (do_phi) {
            const double Rt = stats[(size_t)n * 5];
            double phi_star = Rt / (R + Rt);
            if (fabs(phi_star) < 1.0 / cells_times_sites) phi_star = 0.0;
            params[1] = phi_star;
        }
    }
    // the last CTA to leave resets the counters for the next launch (every CTA has passed every barrier by then)
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(bar + 1, 1u) == (unsigned)G - 1u) { bar[0] = 0u; bar[1] = 0u; __threadfence(); }
    }
}

__global__ void pmm_phi_step(const double *__restrict__ stats, int n, double *__restrict__ params, double R, double cells_times_sites) {
    const double Rt = stats[(size_t)n * 5];
    double phi_star = Rt / (R + Rt);
    if (fabs(phi_star) < 1.0 / cells_times_sites) phi_star = 0.0;
    params[1] = phi_star;
}

}  // namespace pmm
