// pmm_nni.cuh -- batched scoring of NNI candidate trees.
//
// Reference behaviour (laml_libs/Topology_search.py:242-322): for an internal, non-root node u with
// parent v and sibling w, each of u's children c_k in turn is swapped with w; a brand-new solver is
// built from newick strings and `score_tree` runs a full-tree EM in which only the five branches
// around the move are free -- those above u's children, v's children and v (:259) -- with nu and phi
// fixed (:257-258).  Every EM iteration of every candidate walks the whole tree.
//
// Here: after ONE full E-step of the current tree (inside rows X, outside rows O and their per-site
// summaries are cached on the device), a candidate only needs the rows of the three subtrees that
// hang below the move (A = c_{1-k}, B = w, C = c_k), the sibling s of v and the outside row of v's
// parent g:
//     new topology   v' -> (u', C),  u' -> (A, B)          free lengths  d_v, d_u, d_C, d_A, d_B
//     T^v = out[g] .* E^s  (fixed),  x_u' = E^A .* E^B,  x_v' = E^u' .* E^C,  L' = sum_j T^v_j E^v'_j
// and the local outside recursion gives the expected counts of the five branches.  One sweep over
// a site's slots evaluates likelihood and counts: O(5 W) bytes per candidate and EM iteration
// instead of O(n W).  Thousands of candidates advance together: kernel `pmm_nni_sweep` has one CTA
// per (candidate, tile), kernel `pmm_nni_mstep` one thread per candidate (statistics clamp/rescale,
// five 1-D concave problems EM_solver.py:351-357, convergence test :503).
#pragma once
#include "pmm_kernels.cuh"

namespace pmm {

constexpr int NNI_THREADS = 128;
constexpr int NNI_G = 4;
constexpr int NNI_NG = NNI_THREADS / NNI_G;

// NniCand, NniState, consts_of: pmm_common.cuh

__device__ __forceinline__ double gsum4(double x, unsigned m) {
    x += __shfl_xor_sync(m, x, 2);
    x += __shfl_xor_sync(m, x, 1);
    return x;
}

// a child subtree as the sweep sees it: normalised heads, sum q x, and E_alpha = cs * raw + gm
struct Sub { double x0, xm, Sq, E0, Em, cs, gm, sc; };

// partial[cand][tile][26]: 25 counts (branch-major: v,u,C,A,B x S0..S4) + sum log(L'/L)
__global__ void __launch_bounds__(NNI_THREADS) pmm_nni_sweep(KParams P, const NniCand *__restrict__ cands,
                                                              const NniState *__restrict__ state, double *__restrict__ partial) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int TS = P.TS;
    double *tO = reinterpret_cast<double *>(smem_raw);          // out[g] tile, overwritten by T^v
    double *tS = tO + TS, *tA = tS + TS, *tB = tA + TS, *tC = tB + TS;
    __shared__ double4 bc5[5];
    __shared__ double red[NNI_NG][26];
    const int ci = blockIdx.x, t = blockIdx.y;
    const NniState st = state[ci];
    const int tid = threadIdx.x;
    double *out_part = partial + ((size_t)ci * P.ntiles + t) * 26;
    if (st.done) return;                                       // partials of a finished candidate are not read again
    const NniCand cd = cands[ci];
    const int N = P.n_leaves;
    const int s0 = P.tile_site[t], s1 = P.tile_site[t + 1];
    const int ps0 = P.tile_psite[t];
    const int off0 = P.tile_off[t], tlen = P.tile_off[t + 1] - off0;
    const double nu = P.params[0], phi = P.params[1];
    if (tid < 5) bc5[tid] = consts_of(st.d[tid], nu);
    // stage the row tiles of the internal nodes involved (leaves are closed forms)
    auto stage = [&](double *dst, const double *rows, int node) {
        if (node < N) return;
        const double2 *src = reinterpret_cast<const double2 *>(rows + (size_t)(node - N) * P.Wp + off0);
        for (int i = tid; i < tlen / 2; i += NNI_THREADS) reinterpret_cast<double2 *>(dst)[i] = src[i];
    };
    if (cd.g >= 0) { stage(tO, P.O, cd.g); stage(tS, P.X, cd.s); }
    stage(tA, P.X, cd.A); stage(tB, P.X, cd.B); stage(tC, P.X, cd.C);
    __syncthreads();

    const double4 cv = bc5[0], cu = bc5[1], cC = bc5[2], cA = bc5[3], cB = bc5[4];
    const double4 cs_ = (cd.g >= 0) ? reinterpret_cast<const double4 *>(P.bc)[cd.s] : make_double4(0, 0, 0, 0);
    const int gl = tid & (NNI_G - 1), grp = tid / NNI_G;
    const unsigned gm = 0xFu << ((tid & 31) & ~3);
    double acc[26];
#pragma unroll
    for (int k = 0; k < 26; k++) acc[k] = 0.0;

    // summary of a subtree root at site (ls): internal -> from the row + its writer's summary; leaf -> closed form
    auto sub_of = [&](int node, const double *tile, int off, int ls, const double4 c, double qsum, const double *q, Leaf &L) -> Sub {
        Sub r;
        if (node >= N) {
            const size_t mo = (size_t)(node - N) * P.Kp + ps0 + ls;
            r.sc = pow2i(-P.eX[mo]);
            const double2 h = ld2(tile + off);
            r.x0 = h.x * r.sc; r.xm = h.y * r.sc; r.Sq = P.sqX[mo] * r.sc;
            L.hot = -1; L.xb = 0.0; L.dx = 0.0;
        } else {
            L = leaf_of(P.codes[(size_t)node * P.Kp + ps0 + ls], phi);
            r.sc = 1.0; r.x0 = L.x0; r.xm = L.xm;
            r.Sq = L.xb * qsum + ((L.hot >= 0) ? q[off + L.hot] : 0.0) * L.dx;
        }
        r.gm = c.w * r.xm;
        r.E0 = c.x * r.x0 + c.y * r.Sq + r.gm;
        r.Em = r.xm;
        r.cs = c.z * r.sc;
        return r;
    };

    for (int s = s0 + grp; s < s1; s += NNI_NG) {
        const int goff = P.site_off[s], len = P.site_len[s];
        const int off = goff - off0, ls = s - s0;
        const double *q = P.q + off0;                          // q in tile coordinates
        const double qsum = P.qsum[s];
        Leaf lS, lA, lB, lC;
        const Sub A = sub_of(cd.A, tA, off, ls, cA, qsum, q, lA);
        const Sub B = sub_of(cd.B, tB, off, ls, cB, qsum, q, lB);
        const Sub C = sub_of(cd.C, tC, off, ls, cC, qsum, q, lC);
        // T^v = out[g] .* E^s (normalised), or e_0 when v is a root
        double Tv0 = 1.0, Tvm = 0.0, sco = 0.0; int shift = 0;
        Sub S; S.cs = 0.0; S.gm = 0.0;
        lS.hot = -1; lS.xb = 0.0; lS.dx = 0.0;
        double invL = 1.0;
        if (cd.g >= 0) {
            const size_t mo = (size_t)(cd.g - N) * P.Kp + ps0 + ls;
            const int eOg = P.eO[mo];
            sco = pow2i(-eOg);
            S = sub_of(cd.s, tS, off, ls, cs_, qsum, q, lS);
            const double2 oh = ld2(tO + off);
            Tv0 = oh.x * sco * S.E0; Tvm = oh.y * sco * S.Em;
            invL = P.rootInvL[(size_t)cd.root * P.Kp + ps0 + ls];
            // exponent of L'/L: D_g + eO_g + (iexp A + iexp B + iexp C - iexp v)
            const int iv = P.iexp[(size_t)(cd.v - N) * P.Kp + ps0 + ls];
            auto ie = [&](int node) { return node >= N ? P.iexp[(size_t)(node - N) * P.Kp + ps0 + ls] : 0; };
            shift = P.dexp[mo] + eOg + ie(cd.A) + ie(cd.B) + ie(cd.C) - iv;
        } else {
            auto ie = [&](int node) { return node >= N ? P.iexp[(size_t)(node - N) * P.Kp + ps0 + ls] : 0; };
            shift = ie(cd.A) + ie(cd.B) + ie(cd.C);
        }
        const double xum = A.Em * B.Em;                          // x_u'_m
        const double Tu0 = cv.x * Tv0 * C.E0;                    // T^u'_0 = out[v']_0 E^C_0
        const double bvT = cv.y * Tv0, buT = cu.y * Tu0, gux = cu.w * xum;
        double s_qxu = 0, s_qxv = 0, s_tv = 0, s_tvxv = 0, s_tu = 0, s_tuxu = 0, s_tc = 0, s_tcxc = 0, s_ta = 0, s_taxa = 0,
               s_tb = 0, s_tbxb = 0;
        for (int i = 2 + gl; i < len; i += NNI_G) {
            const double qi = q[off + i];
            const double xa = (cd.A >= N) ? tA[off + i] * A.sc : lA.xb + ((i == lA.hot) ? lA.dx : 0.0);
            const double xb = (cd.B >= N) ? tB[off + i] * B.sc : lB.xb + ((i == lB.hot) ? lB.dx : 0.0);
            const double xc = (cd.C >= N) ? tC[off + i] * C.sc : lC.xb + ((i == lC.hot) ? lC.dx : 0.0);
            double tva = 0.0;
            if (cd.g >= 0) {
                const double xs = (cd.s >= N) ? tS[off + i] * S.sc : lS.xb + ((i == lS.hot) ? lS.dx : 0.0);
                tva = tO[off + i] * sco * fma(cs_.z, xs, S.gm);
            }
            const double ea = fma(cA.z, xa, A.gm), eb = fma(cB.z, xb, B.gm), ec = fma(cC.z, xc, C.gm);
            const double xu = ea * eb;
            const double eu = fma(cu.z, xu, gux);
            const double xv = eu * ec;
            s_qxu = fma(qi, xu, s_qxu); s_qxv = fma(qi, xv, s_qxv);
            s_tv += tva; s_tvxv = fma(tva, xv, s_tvxv);
            const double ov = fma(bvT, qi, cv.z * tva);          // out[v']_alpha
            const double tu = ov * ec, tc = ov * eu;
            s_tu += tu; s_tuxu = fma(tu, xu, s_tuxu);
            s_tc += tc; s_tcxc = fma(tc, xc, s_tcxc);
            const double ou = fma(buT, qi, cu.z * tu);           // out[u']_alpha
            const double ta = ou * eb, tb = ou * ea;
            s_ta += ta; s_taxa = fma(ta, xa, s_taxa);
            s_tb += tb; s_tbxb = fma(tb, xb, s_tbxb);
        }
        s_qxu = gsum4(s_qxu, gm); s_qxv = gsum4(s_qxv, gm); s_tv = gsum4(s_tv, gm); s_tvxv = gsum4(s_tvxv, gm);
        s_tu = gsum4(s_tu, gm); s_tuxu = gsum4(s_tuxu, gm); s_tc = gsum4(s_tc, gm); s_tcxc = gsum4(s_tcxc, gm);
        s_ta = gsum4(s_ta, gm); s_taxa = gsum4(s_taxa, gm); s_tb = gsum4(s_tb, gm); s_tbxb = gsum4(s_tbxb, gm);
        if (gl == 0) {
            const double xu0 = A.E0 * B.E0;
            const double Eu0 = cu.x * xu0 + cu.y * s_qxu + gux;
            const double xv0 = Eu0 * C.E0, xvm = xum * C.Em;
            const double gvx = cv.w * xvm;
            const double Ev0 = cv.x * xv0 + cv.y * s_qxv + gvx;
            const double Lloc = Tv0 * Ev0 + (cv.z * s_tvxv + gvx * s_tv) + Tvm * xvm;
            if (Lloc > 0.0) {
                const double iL = 1.0 / Lloc;
                // branch v'
                acc[0] += Tv0 * cv.x * xv0 * iL; acc[1] += Tv0 * cv.y * s_qxv * iL; acc[2] += Tv0 * gvx * iL;
                acc[3] += cv.z * s_tvxv * iL; acc[4] += gvx * s_tv * iL;
                const double ov0 = cv.x * Tv0, ovm = cv.w * (Tv0 + s_tv) + Tvm;
                // branch u'
                acc[5] += Tu0 * cu.x * xu0 * iL; acc[6] += Tu0 * cu.y * s_qxu * iL; acc[7] += Tu0 * gux * iL;
                acc[8] += cu.z * s_tuxu * iL; acc[9] += gux * s_tu * iL;
                const double Tum = ovm * C.Em;
                // branch C: T^C = out[v'] .* E^u'
                const double Tc0 = ov0 * Eu0, gcx = cC.w * C.xm;
                acc[10] += Tc0 * cC.x * C.x0 * iL; acc[11] += Tc0 * cC.y * C.Sq * iL; acc[12] += Tc0 * gcx * iL;
                acc[13] += cC.z * s_tcxc * iL; acc[14] += gcx * s_tc * iL;
                // branches A, B: T^A = out[u'] .* E^B, T^B = out[u'] .* E^A
                const double ou0 = cu.x * Tu0;
                (void)Tum;
                const double Ta0 = ou0 * B.E0, Tb0 = ou0 * A.E0, gax = cA.w * A.xm, gbx = cB.w * B.xm;
                acc[15] += Ta0 * cA.x * A.x0 * iL; acc[16] += Ta0 * cA.y * A.Sq * iL; acc[17] += Ta0 * gax * iL;
                acc[18] += cA.z * s_taxa * iL; acc[19] += gax * s_ta * iL;
                acc[20] += Tb0 * cB.x * B.x0 * iL; acc[21] += Tb0 * cB.y * B.Sq * iL; acc[22] += Tb0 * gbx * iL;
                acc[23] += cB.z * s_tbxb * iL; acc[24] += gbx * s_tb * iL;
                // log L' of this tree at this site: relative to the cached L when v is not a root
                const int sh = max(-1000, min(1000, shift));
                acc[25] += (cd.g >= 0) ? log(scalbn(Lloc * invL, sh)) : log(Lloc) + (double)shift * 0.693147180559945309417232121458;
            } else acc[25] += -INFINITY;
        }
    }
    if (gl == 0) {
#pragma unroll
        for (int k = 0; k < 26; k++) red[grp][k] = acc[k];
    }
    __syncthreads();
    if (tid < 26) {
        double sum = 0.0;
        for (int g = 0; g < NNI_NG; g++) sum += red[g][tid];
        out_part[tid] = sum;
    }
}

// one thread per candidate: sum the tile partials in order, convergence test, M-step of the 5 branches
__global__ void pmm_nni_mstep(const double *__restrict__ partial, int n_cand, int ntiles, const NniCand *__restrict__ cands,
                              const double *__restrict__ root_site_llh_sum, double llh_total, NniState *__restrict__ state,
                              const double *__restrict__ params, int K_total, double dmin, double dmax, double eps_s, double eps_nu,
                              double conv_eps, int max_iter, int *__restrict__ n_active) {
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= n_cand) return;
    NniState st = state[ci];
    if (st.done) return;
    double s[26];
    for (int k = 0; k < 26; k++) s[k] = 0.0;
    for (int t = 0; t < ntiles; t++)
        for (int k = 0; k < 26; k++) s[k] += partial[((size_t)ci * ntiles + t) * 26 + k];
    const NniCand cd = cands[ci];
    // llh of the forest with this tree replaced: rooted moves recompute the whole tree's likelihood,
    // others move it by sum log(L'/L)
    const double llh = (cd.g >= 0) ? llh_total + s[25] : llh_total - root_site_llh_sum[cd.root] + s[25];
    st.llh_prev = st.llh_cur;
    st.llh_cur = llh;
    if (st.iters > 0 && (fabs((llh - st.llh_prev) / st.llh_prev) < conv_eps || st.iters >= max_iter || !(llh > -INFINITY))) {
        st.done = 1;                                        // EM_solver.py:503; st.d is the point llh was evaluated at
        state[ci] = st;
        return;
    }
    const double nu = params[0];
    const bool use_nu = nu > eps_nu;
    for (int b = 0; b < 5; b++) {
        double x[5], tot = 0.0;
        for (int k = 0; k < 5; k++) { x[k] = fmax(s[b * 5 + k], eps_s); tot += x[k]; }
        for (int k = 0; k < 5; k++) x[k] = x[k] / tot * (double)K_total;
        const double s24 = x[2] + x[4];
        double lo = dmin, hi = dmax, res;
        if (dobj_dd(lo, nu, use_nu, x[0], x[1], s24, x[3]) <= 0.0) res = lo;
        else if (dobj_dd(hi, nu, use_nu, x[0], x[1], s24, x[3]) >= 0.0) res = hi;
        else {
            for (int it = 0; it < 200; it++) {
                const double mid = 0.5 * (lo + hi);
                if (mid <= lo || mid >= hi) break;
                if (dobj_dd(mid, nu, use_nu, x[0], x[1], s24, x[3]) > 0.0) lo = mid; else hi = mid;
            }
            res = 0.5 * (lo + hi);
        }
        st.d[b] = res;
    }
    st.iters++;
    state[ci] = st;
    atomicAdd(n_active, 1);
}

}  // namespace pmm
