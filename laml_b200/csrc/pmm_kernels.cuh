// pmm_kernels.cuh -- sm_100a kernels of the PMM likelihood / E-step engine.
//
// Model and formulas: SURVEY.md section 3.5 (full-state restatement of
// laml_libs/EM_solver.py:91-300; verified against the reference in
// tests/test_fullstate_model.py).  Design: DESIGN.md.
//
// Data layout (per site chunk): every internal node owns one row of `Wp` fp64
// slots.  A site s occupies `site_len[s] = A_s + 2` consecutive slots starting
// at `site_off[s]`, ordered [x_0 (unedited), x_m (silenced), x_1 .. x_A].  Sites
// are grouped into tiles of at most TS slots; every tile starts on an even slot
// (16-byte aligned) so a tile of a row is one contiguous, aligned byte range
// that a single 1-D TMA bulk copy (cp.async.bulk) moves between HBM and shared
// memory.  Leaves own no row: their vectors are synthesised in shared memory
// from 4-byte state codes.
//
// Scaling: rows are stored UNnormalised; whoever reads a row first finds the
// largest exponent among a site's slots (integer max over the high words --
// all values are non-negative) and multiplies by the exact power of two that
// brings the maximum into [1,2).  The accumulated exponents live in int32
// arrays (iexp / oexp), one entry per (internal node, site).  Power-of-two
// scaling is exact, so the mantissas are those of an unscaled computation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pmm {

constexpr int NT = 256;          // threads per CTA

struct KParams {
    // chunk layout
    const int *site_off, *site_len, *tile_site, *tile_off;
    const double *q;             // [Wp] prior in slot layout (0 at x_0, x_m and pad slots)
    int K, Wp, ntiles;
    // data
    const int *codes;            // [n_leaves][codes_stride]
    int codes_stride, site0;
    // tree
    const int *left, *right, *parent, *root_index;
    int n_leaves, n_nodes;
    // parameters (device resident so that an EM loop never needs the host)
    const double *params;        // [0]=nu, [1]=phi
    const double *bc;            // [n][4] = a, b, c, g per branch
    // state
    double *X, *O;               // [n_int][Wp] inside / outside rows
    int *iexp, *oexp;            // [n_int][K] accumulated exponents of the stored rows
    double *rootLm; int *rootLe; // [n_roots][K] site likelihood mantissa / exponent
    double *root_site_llh;       // [n_roots][K]
    double *Spart;               // [ntiles][n][5]
    double *Rpart;               // [ntiles][n_int]
    double *post0, *post1;       // optional [n][K_total] log posteriors (annotation path), or null
    int K_total;
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
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
// 1-D TMA bulk copy shared -> global
__device__ __forceinline__ void bulk_store(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- scalar helpers
__device__ __forceinline__ unsigned hi32(double x) { return (unsigned)__double2hiint(x); }
// exponent e such that max * 2^-e is in [1,2); 0 for zero / denormal maxima (dead site)
__device__ __forceinline__ int expo_of_hi(unsigned hi) {
    int be = (int)((hi >> 20) & 0x7ffu);
    return be ? be - 1023 : 0;
}
__device__ __forceinline__ double pow2i(int e) {          // exact 2^e for e in [-1022,1023]
    return __hiloint2double((1023 + e) << 20, 0);
}

template <int G> __device__ __forceinline__ unsigned group_mask() {
    if (G == 32) return 0xffffffffu;
    unsigned lane = threadIdx.x & 31u;
    return ((1u << G) - 1u) << (lane & ~(unsigned)(G - 1));
}
template <int G> __device__ __forceinline__ double gsum(double x, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) x += __shfl_xor_sync(m, x, o);
    return x;
}
template <int G> __device__ __forceinline__ unsigned gmax(unsigned x, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) x = max(x, __shfl_xor_sync(m, x, o));
    return x;
}

// Leaf state vector written into a shared-memory tile (EM_solver.py:103-116 in
// full-state form): '?' -> (phi, 1, phi..), known silenced -> e_m,
// 0 -> (1-phi) e_0, alpha -> (1-phi) e_alpha.
template <int G>
__device__ __forceinline__ void synth_leaf(double *x, int off, int len, int code, double phi, int gl, unsigned gm) {
    double base = (code == -1) ? phi : 0.0;
    for (int i = 2 + gl; i < len; i += G) x[off + i] = base;
    __syncwarp(gm);
    if (gl == 0) {
        x[off] = (code == 0) ? 1.0 - phi : ((code == -1) ? phi : 0.0);
        x[off + 1] = (code < 0) ? 1.0 : 0.0;
        if (code > 0) x[off + 1 + code] = 1.0 - phi;
    }
    __syncwarp(gm);
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

// ---------------------------------------------------------------- K1: inside, one level
// One CTA per (internal node u of this level, tile).  Children rows are staged
// in shared memory by TMA bulk copies (leaf children are synthesised), groups
// of G lanes own one site each: pass 1 reduces sum(q x) and the exponent of the
// maximum per child, pass 2 forms x_u = E_v .* E_w in place; one bulk store
// writes the tile of row u.
template <int G>
__global__ void __launch_bounds__(NT) pmm_inside_level(KParams P, const int *__restrict__ nodes, int TS) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *xv = reinterpret_cast<double *>(smem_raw);
    double *xw = xv + TS;
    uint64_t *bar = reinterpret_cast<uint64_t *>(xw + TS);

    const int t = blockIdx.x % P.ntiles;
    const int u = nodes[blockIdx.x / P.ntiles];
    const int v = P.left[u], w = P.right[u];
    const int N = P.n_leaves;
    const bool v_int = v >= N, w_int = w >= N;
    const int s0 = P.tile_site[t], s1 = P.tile_site[t + 1];
    const int off0 = P.tile_off[t];
    const int tlen = P.tile_off[t + 1] - off0;
    const int tid = threadIdx.x;

    if (v_int || w_int) {
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) {
            uint32_t bytes = (uint32_t)tlen * 8u;
            mbar_expect_tx(bar, bytes * ((v_int ? 1u : 0u) + (w_int ? 1u : 0u)));
            if (v_int) bulk_load(xv, P.X + (size_t)(v - N) * P.Wp + off0, bytes, bar);
            if (w_int) bulk_load(xw, P.X + (size_t)(w - N) * P.Wp + off0, bytes, bar);
        }
    }
    const double phi = P.params[1];
    const double4 cv = reinterpret_cast<const double4 *>(P.bc)[v];
    const double4 cw = reinterpret_cast<const double4 *>(P.bc)[w];
    const unsigned gm = group_mask<G>();
    const int gl = tid & (G - 1);
    const int grp = tid / G;
    constexpr int NG = NT / G;

    // leaf children: synthesise their vectors while the bulk loads are in flight
    if (!v_int || !w_int) {
        for (int s = s0 + grp; s < s1; s += NG) {
            int off = P.site_off[s] - off0, len = P.site_len[s];
            if (!v_int) synth_leaf<G>(xv, off, len, P.codes[(size_t)v * P.codes_stride + P.site0 + s], phi, gl, gm);
            if (!w_int) synth_leaf<G>(xw, off, len, P.codes[(size_t)w * P.codes_stride + P.site0 + s], phi, gl, gm);
        }
        // the tile's tail pad slot (if any) travels with the bulk store of row u: keep it zero
        if (tid == 0 && !v_int) {
            const int real_end = P.site_off[s1 - 1] + P.site_len[s1 - 1] - off0;
            if (real_end < tlen) xv[real_end] = 0.0;
        }
    }
    if (v_int || w_int) mbar_wait(bar, 0);

    int *iexp_u = P.iexp + (size_t)(u - N) * P.K;
    const int *iexp_v = v_int ? P.iexp + (size_t)(v - N) * P.K : nullptr;
    const int *iexp_w = w_int ? P.iexp + (size_t)(w - N) * P.K : nullptr;

    for (int s = s0 + grp; s < s1; s += NG) {
        const int goff = P.site_off[s];
        const int off = goff - off0, len = P.site_len[s];
        double sv = 0.0, sw = 0.0;
        unsigned mv = 0u, mw = 0u;
        for (int i = 2 + gl; i < len; i += G) {
            double qi = __ldg(P.q + goff + i);
            double a = xv[off + i], b = xw[off + i];
            sv = fma(qi, a, sv);
            sw = fma(qi, b, sw);
            mv = max(mv, hi32(a));
            mw = max(mw, hi32(b));
        }
        const double x0v = xv[off], xmv = xv[off + 1], x0w = xw[off], xmw = xw[off + 1];
        mv = max(mv, max(hi32(x0v), hi32(xmv)));
        mw = max(mw, max(hi32(x0w), hi32(xmw)));
        sv = gsum<G>(sv, gm);
        sw = gsum<G>(sw, gm);
        mv = gmax<G>(mv, gm);
        mw = gmax<G>(mw, gm);
        const int ev = expo_of_hi(mv), ew = expo_of_hi(mw);
        const double scv = pow2i(-ev), scw = pow2i(-ew);
        // normalised child messages (SURVEY 3.5 "edge transfer")
        const double xmv_n = xmv * scv, xmw_n = xmw * scw;
        const double k0v = (cv.x * x0v + cv.y * sv) * scv + cv.w * xmv_n;
        const double k0w = (cw.x * x0w + cw.y * sw) * scw + cw.w * xmw_n;
        const double gmv = cv.w * xmv_n, gmw = cw.w * xmw_n;
        const double csv = cv.z * scv, csw = cw.z * scw;
        for (int i = 2 + gl; i < len; i += G)
            xv[off + i] = fma(csv, xv[off + i], gmv) * fma(csw, xw[off + i], gmw);
        if (gl == 0) {
            xv[off] = k0v * k0w;
            xv[off + 1] = xmv_n * xmw_n;
            iexp_u[s] = (v_int ? iexp_v[s] : 0) + ev + (w_int ? iexp_w[s] : 0) + ew;
        }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        bulk_store(P.X + (size_t)(u - N) * P.Wp + off0, xv, (uint32_t)tlen * 8u);
        bulk_commit();
        bulk_wait_read0();
    }
}

// ---------------------------------------------------------------- K1r: root likelihood
// One CTA per (root, tile): site likelihood L_s = E[root]_0 (the root's own branch
// starts in state 0, EM_solver.py:140), its log, and the expected counts of the
// root branch (EM_solver.py:282-286).
template <int G>
__global__ void __launch_bounds__(NT) pmm_root_llh(KParams P, const int *__restrict__ roots) {
    __shared__ double red[NT / G][3];
    const int t = blockIdx.x % P.ntiles;
    const int ri = blockIdx.x / P.ntiles;
    const int r = roots[ri];
    const int N = P.n_leaves;
    const int s0 = P.tile_site[t], s1 = P.tile_site[t + 1];
    const int tid = threadIdx.x;
    const unsigned gm = group_mask<G>();
    const int gl = tid & (G - 1), grp = tid / G;
    constexpr int NG = NT / G;
    const double4 cr = reinterpret_cast<const double4 *>(P.bc)[r];
    const double *x = P.X + (size_t)(r - N) * P.Wp;
    const int *iexp_r = P.iexp + (size_t)(r - N) * P.K;
    double acc0 = 0, acc1 = 0, acc2 = 0;
    for (int s = s0 + grp; s < s1; s += NG) {
        const int goff = P.site_off[s], len = P.site_len[s];
        double sq = 0.0;
        unsigned mx = 0u;
        for (int i = 2 + gl; i < len; i += G) {
            double a = x[goff + i];
            sq = fma(__ldg(P.q + goff + i), a, sq);
            mx = max(mx, hi32(a));
        }
        const double x0 = x[goff], xm = x[goff + 1];
        mx = max(mx, max(hi32(x0), hi32(xm)));
        sq = gsum<G>(sq, gm);
        mx = gmax<G>(mx, gm);
        if (gl == 0) {
            const int e = expo_of_hi(mx);
            const double sc = pow2i(-e);
            const double t0 = cr.x * x0 * sc, t1 = cr.y * sq * sc, t2 = cr.w * xm * sc;
            const double L = t0 + t1 + t2;
            const int Le = iexp_r[s] + e;
            P.rootLm[(size_t)ri * P.K + s] = L;
            P.rootLe[(size_t)ri * P.K + s] = Le;
            P.root_site_llh[(size_t)ri * P.K + s] = (L > 0.0) ? log(L) + (double)Le * 0.693147180559945309417232121458 : -INFINITY;
            if (L > 0.0) {
                const double inv = 1.0 / L;
                acc0 += t0 * inv; acc1 += t1 * inv; acc2 += t2 * inv;
            }
            if (P.post0) {
                // log P(root = 0 | D), log P(root silenced | D)  (EM_solver.py:276-277)
                P.post0[(size_t)r * P.K_total + P.site0 + s] = (L > 0.0 && t0 > 0.0) ? log(t0 / L) : -INFINITY;
                P.post1[(size_t)r * P.K_total + P.site0 + s] = (L > 0.0 && t2 > 0.0) ? log(t2 / L) : -INFINITY;
            }
        }
    }
    if (gl == 0) { red[grp][0] = acc0; red[grp][1] = acc1; red[grp][2] = acc2; }
    __syncthreads();
    if (tid < 5) {
        double sum = 0.0;
        if (tid < 3) for (int g = 0; g < NG; g++) sum += red[g][tid];
        P.Spart[((size_t)t * P.n_nodes + r) * 5 + tid] = sum;
    }
}

// ---------------------------------------------------------------- K2: outside + counts, one level
// One CTA per (internal node u of this depth, tile).  Stages out[u] (or builds
// the root's from the branch constants), x[v], x[w]; per site it forms the two
// sibling messages, the top-of-branch vectors T_v = out[u] .* E_w and
// T_w = out[u] .* E_v, writes out[v], out[w] (internal children only) and the
// expected transition counts S0..S4 of both child branches
// (EM_solver.py:143-300 in full-state form).  All six per-site
// reductions are taken on the raw rows in ONE pass:
//   sum q xv, sum q xw, sum o, sum o xv, sum o xw, sum o xv xw.
template <int G>
__global__ void __launch_bounds__(NT) pmm_outside_level(KParams P, const int *__restrict__ nodes, int TS) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *xo = reinterpret_cast<double *>(smem_raw);
    double *xv = xo + TS;
    double *xw = xv + TS;
    uint64_t *bar = reinterpret_cast<uint64_t *>(xw + TS);
    constexpr int NG = NT / G;
    double(*red)[11] = reinterpret_cast<double(*)[11]>(bar + 2);

    const int t = blockIdx.x % P.ntiles;
    const int u = nodes[blockIdx.x / P.ntiles];
    const int v = P.left[u], w = P.right[u];
    const int N = P.n_leaves;
    const bool v_int = v >= N, w_int = w >= N;
    const bool u_root = P.parent[u] < 0;
    const int s0 = P.tile_site[t], s1 = P.tile_site[t + 1];
    const int off0 = P.tile_off[t];
    const int tlen = P.tile_off[t + 1] - off0;
    const int tid = threadIdx.x;
    const int nload = (v_int ? 1 : 0) + (w_int ? 1 : 0) + (u_root ? 0 : 1);

    if (nload) {
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) {
            uint32_t bytes = (uint32_t)tlen * 8u;
            mbar_expect_tx(bar, bytes * (uint32_t)nload);
            if (!u_root) bulk_load(xo, P.O + (size_t)(u - N) * P.Wp + off0, bytes, bar);
            if (v_int) bulk_load(xv, P.X + (size_t)(v - N) * P.Wp + off0, bytes, bar);
            if (w_int) bulk_load(xw, P.X + (size_t)(w - N) * P.Wp + off0, bytes, bar);
        }
    }
    const double phi = P.params[1];
    const double4 cu = reinterpret_cast<const double4 *>(P.bc)[u];
    const double4 cv = reinterpret_cast<const double4 *>(P.bc)[v];
    const double4 cw = reinterpret_cast<const double4 *>(P.bc)[w];
    const unsigned gm = group_mask<G>();
    const int gl = tid & (G - 1), grp = tid / G;

    if (!v_int || !w_int || u_root) {
        for (int s = s0 + grp; s < s1; s += NG) {
            const int goff = P.site_off[s];
            const int off = goff - off0, len = P.site_len[s];
            if (!v_int) synth_leaf<G>(xv, off, len, P.codes[(size_t)v * P.codes_stride + P.site0 + s], phi, gl, gm);
            if (!w_int) synth_leaf<G>(xw, off, len, P.codes[(size_t)w * P.codes_stride + P.site0 + s], phi, gl, gm);
            if (u_root) {
                // out[root]: the root branch leaves a progenitor in state 0 (EM_solver.py:151-158)
                for (int i = 2 + gl; i < len; i += G) xo[off + i] = cu.y * __ldg(P.q + goff + i);
                if (gl == 0) { xo[off] = cu.x; xo[off + 1] = cu.w; }
                __syncwarp(gm);
            }
        }
    }
    if (nload) mbar_wait(bar, 0);

    const int ri = P.root_index[u];
    const double *Lm = P.rootLm + (size_t)ri * P.K;
    const int *Le = P.rootLe + (size_t)ri * P.K;
    const int *iexp_u = P.iexp + (size_t)(u - N) * P.K;
    const int *iexp_v = v_int ? P.iexp + (size_t)(v - N) * P.K : nullptr;
    const int *iexp_w = w_int ? P.iexp + (size_t)(w - N) * P.K : nullptr;
    const int *oexp_u = u_root ? nullptr : P.oexp + (size_t)(u - N) * P.K;
    int *oexp_v = v_int ? P.oexp + (size_t)(v - N) * P.K : nullptr;
    int *oexp_w = w_int ? P.oexp + (size_t)(w - N) * P.K : nullptr;

    double acc[11];
#pragma unroll
    for (int k = 0; k < 11; k++) acc[k] = 0.0;

    for (int s = s0 + grp; s < s1; s += NG) {
        const int goff = P.site_off[s];
        const int off = goff - off0, len = P.site_len[s];
        double sv = 0, sw = 0, so = 0, sov = 0, sow = 0, sovw = 0;
        unsigned mv = 0u, mw = 0u, mo = 0u;
        for (int i = 2 + gl; i < len; i += G) {
            const double qi = __ldg(P.q + goff + i);
            const double a = xv[off + i], b = xw[off + i], o = xo[off + i];
            const double oa = o * a;
            sv = fma(qi, a, sv);
            sw = fma(qi, b, sw);
            so += o;
            sov += oa;
            sow = fma(o, b, sow);
            sovw = fma(oa, b, sovw);
            mv = max(mv, hi32(a));
            mw = max(mw, hi32(b));
            mo = max(mo, hi32(o));
        }
        const double x0v = xv[off], xmv = xv[off + 1], x0w = xw[off], xmw = xw[off + 1];
        const double o0 = xo[off], om = xo[off + 1];
        mv = max(mv, max(hi32(x0v), hi32(xmv)));
        mw = max(mw, max(hi32(x0w), hi32(xmw)));
        mo = max(mo, max(hi32(o0), hi32(om)));
        sv = gsum<G>(sv, gm); sw = gsum<G>(sw, gm); so = gsum<G>(so, gm);
        sov = gsum<G>(sov, gm); sow = gsum<G>(sow, gm); sovw = gsum<G>(sovw, gm);
        mv = gmax<G>(mv, gm); mw = gmax<G>(mw, gm); mo = gmax<G>(mo, gm);
        const int ev = expo_of_hi(mv), ew = expo_of_hi(mw), eo = expo_of_hi(mo);
        const double scv = pow2i(-ev), scw = pow2i(-ew), sco = pow2i(-eo);
        // normalised heads and sums
        const double nx0v = x0v * scv, nxmv = xmv * scv, nx0w = x0w * scw, nxmw = xmw * scw;
        const double no0 = o0 * sco, nom = om * sco;
        const double Sv = sv * scv, Sw = sw * scw, So = so * sco;
        const double Sov = sov * sco * scv, Sow = sow * sco * scw, Sovw = sovw * sco * scv * scw;
        // sibling messages E: E_0, E_alpha = c x + g x_m, E_m = x_m
        const double k0v = cv.x * nx0v + cv.y * Sv + cv.w * nxmv, gmv = cv.w * nxmv;
        const double k0w = cw.x * nx0w + cw.y * Sw + cw.w * nxmw, gmw = cw.w * nxmw;
        // top-of-branch vectors: T^v = out[u] .* E^w,  T^w = out[u] .* E^v
        const double Tv0 = no0 * k0w, Tvm = nom * nxmw, sTv = cw.z * Sow + gmw * So;
        const double Tw0 = no0 * k0v, Twm = nom * nxmv, sTw = cv.z * Sov + gmv * So;
        const double sTvx = cw.z * Sovw + gmw * Sov;     // sum_a T^v_a x^v_a
        const double sTwx = cv.z * Sovw + gmv * Sow;     // sum_a T^w_a x^w_a
        // out rows of the children, in place: out[v] over xv, out[w] over xw
        {
            const double bv = cv.y * Tv0, bw = cw.y * Tw0;
            const double cvo = cv.z * sco, cwo = cw.z * sco;
            const double csv = cv.z * scv, csw = cw.z * scw;
            for (int i = 2 + gl; i < len; i += G) {
                const double qi = __ldg(P.q + goff + i);
                const double a = xv[off + i], b = xw[off + i], o = xo[off + i];
                const double Ew = fma(csw, b, gmw), Ev = fma(csv, a, gmv);
                xv[off + i] = fma(bv, qi, cvo * o * Ew);
                xw[off + i] = fma(bw, qi, cwo * o * Ev);
            }
        }
        const double outvm = cv.w * (Tv0 + sTv) + Tvm;
        const double outwm = cw.w * (Tw0 + sTw) + Twm;
        if (gl == 0) {
            xv[off] = cv.x * Tv0; xv[off + 1] = outvm;
            xw[off] = cw.x * Tw0; xw[off + 1] = outwm;
            const int cum_v = v_int ? iexp_v[s] : 0, cum_w = w_int ? iexp_w[s] : 0;
            const int oe = (u_root ? 0 : oexp_u[s]) + eo;
            if (v_int) oexp_v[s] = oe + cum_w + ew;
            if (w_int) oexp_w[s] = oe + cum_v + ev;
            // expected counts (EM_solver.py:279-300): everything is relative to the site likelihood
            const double L = Lm[s];
            int delta = oe + iexp_u[s] - Le[s];
            delta = max(-1000, min(1000, delta));
            const double fac = (L > 0.0) ? scalbn(1.0 / L, delta) : 0.0;
            acc[0] += Tv0 * cv.x * nx0v * fac;
            acc[1] += Tv0 * cv.y * Sv * fac;
            acc[2] += Tv0 * cv.w * nxmv * fac;
            acc[3] += cv.z * sTvx * fac;
            acc[4] += cv.w * nxmv * sTv * fac;
            acc[5] += Tw0 * cw.x * nx0w * fac;
            acc[6] += Tw0 * cw.y * Sw * fac;
            acc[7] += Tw0 * cw.w * nxmw * fac;
            acc[8] += cw.z * sTwx * fac;
            acc[9] += cw.w * nxmw * sTw * fac;
            // phi statistic: 1 - P(leaf silenced | D) on '?' entries (EM_solver.py:323)
            if (!v_int && P.codes[(size_t)v * P.codes_stride + P.site0 + s] == -1) acc[10] += 1.0 - outvm * nxmv * fac;
            if (!w_int && P.codes[(size_t)w * P.codes_stride + P.site0 + s] == -1) acc[10] += 1.0 - outwm * nxmw * fac;
            if (P.post0) {
                // post0 = log P(node in state 0 | D) = log(out_0 x_0 / L); post1 likewise for the
                // silenced state (EM_solver.py:276-277)
                const size_t gs = (size_t)P.site0 + s;
                const double p0v = cv.x * Tv0 * nx0v * fac, p1v = outvm * nxmv * fac;
                const double p0w = cw.x * Tw0 * nx0w * fac, p1w = outwm * nxmw * fac;
                P.post0[(size_t)v * P.K_total + gs] = (p0v > 0.0) ? log(p0v) : -INFINITY;
                P.post1[(size_t)v * P.K_total + gs] = (p1v > 0.0) ? log(p1v) : -INFINITY;
                P.post0[(size_t)w * P.K_total + gs] = (p0w > 0.0) ? log(p0w) : -INFINITY;
                P.post1[(size_t)w * P.K_total + gs] = (p1w > 0.0) ? log(p1w) : -INFINITY;
            }
        }
    }
    if (gl == 0) {
#pragma unroll
        for (int k = 0; k < 11; k++) red[grp][k] = acc[k];
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0 && (v_int || w_int)) {
        if (v_int) bulk_store(P.O + (size_t)(v - N) * P.Wp + off0, xv, (uint32_t)tlen * 8u);
        if (w_int) bulk_store(P.O + (size_t)(w - N) * P.Wp + off0, xw, (uint32_t)tlen * 8u);
        bulk_commit();
    }
    if (tid >= 32 && tid < 32 + 11) {
        const int k = tid - 32;
        double sum = 0.0;
        for (int g = 0; g < NG; g++) sum += red[g][k];
        if (k < 5) P.Spart[((size_t)t * P.n_nodes + v) * 5 + k] = sum;
        else if (k < 10) P.Spart[((size_t)t * P.n_nodes + w) * 5 + (k - 5)] = sum;
        else P.Rpart[(size_t)t * (P.n_nodes - N) + (u - N)] = sum;
    }
    if (tid == 0 && (v_int || w_int)) bulk_wait_read0();
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

// ---------------------------------------------------------------- K3: M-step
// stats = [S(n*5) | Rtilde | llh].  EM_solver.py:318-348 (phi, clamp, rescale),
// :351-357 (per-branch concave problem in d), :394-407 (1-D problem in nu).
__device__ __forceinline__ double dobj_dd(double d, double nu, bool use_nu, double s0, double s1, double s24, double s3) {
    double g = -(nu + 1.0) * s0 - nu * (s1 + s3) + s1 / expm1(d);
    if (use_nu) g += nu * s24 / expm1(nu * d);
    return g;
}

__global__ void pmm_mstep_branches(const double *__restrict__ stats, const unsigned char *__restrict__ free_mask, int n,
                                   int K_total, double *__restrict__ params, double *__restrict__ d,
                                   double *__restrict__ Snorm, double dmin, double dmax, double eps_s, double eps_nu) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n || !free_mask[v]) return;
    const double nu = params[0];
    double s[5], tot = 0.0;
#pragma unroll
    for (int k = 0; k < 5; k++) { s[k] = fmax(stats[(size_t)v * 5 + k], eps_s); }
#pragma unroll
    for (int k = 0; k < 5; k++) tot += s[k];
#pragma unroll
    for (int k = 0; k < 5; k++) { s[k] = s[k] / tot * (double)K_total; Snorm[(size_t)v * 5 + k] = s[k]; }
    const bool use_nu = nu > eps_nu;
    const double s24 = s[2] + s[4];
    double lo = dmin, hi = dmax, res;
    if (dobj_dd(lo, nu, use_nu, s[0], s[1], s24, s[3]) <= 0.0) res = lo;
    else if (dobj_dd(hi, nu, use_nu, s[0], s[1], s24, s[3]) >= 0.0) res = hi;
    else {
        for (int it = 0; it < 200; it++) {
            double mid = 0.5 * (lo + hi);
            if (mid <= lo || mid >= hi) break;
            if (dobj_dd(mid, nu, use_nu, s[0], s[1], s24, s[3]) > 0.0) lo = mid; else hi = mid;
        }
        res = 0.5 * (lo + hi);
    }
    d[v] = res;
}

// single CTA: phi closed form and the nu bisection with deterministic block reductions
__global__ void __launch_bounds__(1024) pmm_mstep_globals(const double *__restrict__ stats, const unsigned char *__restrict__ free_mask,
                                                           int n, const double *__restrict__ Snorm, const double *__restrict__ d,
                                                           double *__restrict__ params, double R, double cells_times_sites,
                                                           int optimize_nu, int optimize_phi) {
    __shared__ double sh[1024];
    __shared__ double bcast;
    const int tid = threadIdx.x;
    auto block_sum = [&](double x) -> double {
        sh[tid] = x;
        __syncthreads();
        for (int o = 512; o > 0; o >>= 1) {
            if (tid < o) sh[tid] += sh[tid + o];
            __syncthreads();
        }
        double r = sh[0];
        __syncthreads();
        return r;
    };
    if (optimize_nu) {
        double lin_p = 0.0;
        for (int v = tid; v < n; v += 1024)
            if (free_mask[v]) lin_p += (Snorm[(size_t)v * 5] + Snorm[(size_t)v * 5 + 1] + Snorm[(size_t)v * 5 + 3]) * d[v];
        const double lin = block_sum(lin_p);
        auto grad = [&](double x) -> double {
            double g = 0.0;
            for (int v = tid; v < n; v += 1024)
                if (free_mask[v]) g += (Snorm[(size_t)v * 5 + 2] + Snorm[(size_t)v * 5 + 4]) * d[v] / expm1(x * d[v]);
            return block_sum(g) - lin;
        };
        double lo, hi = 1.0, res;
        if (grad(hi) >= 0.0) res = 1.0;
        else {
            lo = 0.5;
            bool zero = false;
            for (;;) {
                if (grad(lo) > 0.0) break;
                hi = lo; lo *= 0.5;
                if (lo < 1e-300) { zero = true; break; }
            }
            if (zero) res = 0.0;
            else {
                for (int it = 0; it < 200; it++) {
                    double mid = 0.5 * (lo + hi);
                    if (mid <= lo || mid >= hi) break;
                    if (grad(mid) > 0.0) lo = mid; else hi = mid;
                }
                res = 0.5 * (lo + hi);
            }
        }
        if (tid == 0) bcast = res;
        __syncthreads();
        if (tid == 0) params[0] = bcast;
    }
    if (optimize_phi && tid == 0) {
        const double Rt = stats[(size_t)n * 5];
        double phi_star = Rt / (R + Rt);
        if (fabs(phi_star) < 1.0 / cells_times_sites) phi_star = 0.0;
        params[1] = phi_star;
    }
}

}  // namespace pmm
