// pmm_nni_lumped.cuh -- batched scoring of NNI candidate trees on the lumped-state records
// (reference behaviour: laml_libs/Topology_search.py:242-322, see pmm_nni.cuh).
//
// After ONE full E-step of the current tree with every out-record stored, a candidate needs, per site,
// five 40-byte records: the inside records of the three subtrees below the move (A = c_{1-k}, B = w,
// C = c_k) and of the sibling s of v, and the out-record of v's parent g:
//     new topology   v' -> (u', C),  u' -> (A, B)          free lengths  d_v, d_u, d_C, d_A, d_B
//     T^v = out[g] .* E^s  (fixed),  x_u' = E^A .* E^B,  x_v' = E^u' .* E^C,  L' = sum_j T^v_j E^v'_j
// and the local outside recursion gives the expected counts of the five branches -- all O(1) per site in
// lumped form.  ONE CTA owns a candidate for its whole local EM (EM_solver.py:461-512 restricted to five
// free branches): threads stride over the sites, a fixed-order block reduction yields the 25 counts and
// sum log L', five threads solve the five 1-D concave problems (EM_solver.py:351-357), thread 0 applies
// the convergence test (:503).  One launch scores every candidate; nothing but the records is read.
#pragma once
#include "pmm_lumped.cuh"
#ifndef PMM_HOST_EMU
#include "pmm_nni.cuh"
#endif
#include "pmm_nni_ultra.cuh"

namespace pmm {

// out of line: the barrier Newton loop must not compete for registers with the site loop of the kernel.
// Called by all 32 lanes of a warp with the same arguments (pmm_nni_ultra.cuh: lanes 0-4 evaluate one edge each).
__device__ __noinline__ int ultra2_solve_dev(const NniUltra *c, const double *S, double nu, double eps_nu, double lo, double hi,
                                             bool warm, double *H, double *d5) {
    return ultra2_solve(*c, S, nu, eps_nu, lo, hi, warm, H, d5, nullptr);
}

constexpr int NNIL_THREADS = 128;

// x_u = E^v .* E^w of two decoded children as an (unnormalised) operand
__device__ __forceinline__ Opd merge_children(const Child &v, const Child &w) {
    Opd r;
    const bool bv = v.h >= 0, bw = w.h >= 0;
    const bool clash = bv && bw && v.h != w.h;
    const int hu = bv ? (clash ? -1 : v.h) : w.h;
    r.r.x = v.E0 * w.E0;
    r.r.y = v.Eh * (clash ? w.Eb : w.Eh);
    r.r.z = (hu >= 0) ? (bv ? v.qh : w.qh) : v.xm * w.xm;
    r.r.w = pack_hc(hu, 0);
    r.ri = make_int2(0, 0);
    return r;
}

// ULTRA: the M-step of the five branches under the ultrametric equalities (pmm_nni_ultra.cuh) instead of five
// independent 1-D problems; `ultra` holds the fixed heights around each candidate.  A candidate whose M-step is
// infeasible (or fails) ends with done = 2 and is left to the caller.
template <bool ULTRA>
__global__ void __launch_bounds__(NNIL_THREADS, 3) pmm_nni_lumped(LParams P, const NniCand *__restrict__ cands, NniState *__restrict__ state,
                                                                const double *__restrict__ root_llh_sum, double llh_total, int K_total,
                                                                double dmin, double dmax, double eps_s, double eps_nu, double conv_eps,
                                                                int max_iter, const NniUltra *__restrict__ ultra) {
    __shared__ double s_d[5];
    __shared__ double4 s_bc[5];
    __shared__ double s_red[NNIL_THREADS / 32][26];
    __shared__ double s_tot[26];
    __shared__ double s_end[4];
    __shared__ double4 s_lut[4];
    __shared__ int s_done;
    __shared__ double s_S[ULTRA ? 25 : 1], s_H[3];
    const int ci = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const NniCand cd = cands[ci];
    NniState st = state[ci];
    const int N = P.n_leaves;
    const double nu = P.params[0], phi = P.params[1];
    const bool use_nu = nu > eps_nu;
    if (tid < 5) s_d[tid] = st.d[tid];
    if (tid < 4) {
        s_end[tid] = mstep_end_const(tid, nu, use_nu, dmin, dmax);
        const int c = tid;
        s_lut[c] = make_double4(c == 1 ? phi : (c == 2 ? 1.0 - phi : 0.0), c < 2 ? 1.0 : 0.0, c == 1 ? phi : (c == 3 ? 1.0 - phi : 0.0),
                                c == 1 ? phi : 0.0);
    }
    if (tid == 0) s_done = 0;
    const double4 bcs = (cd.g >= 0) ? ld_bc(P.bc, cd.s) : make_double4(0.0, 0.0, 0.0, 0.0);
    __syncthreads();

    // a subtree root as an operand: its inside record, or the leaf table + the prior of the leaf's state
    auto operand = [&](int node, int s, int gs, const double *prior_site) {
        Opd p;
        if (node >= N) { const unsigned o = rec_index(P, node, s); ld_xrec(p, P.X, o); }
        else {
            const int code = __ldg(P.codes + (size_t)node * P.code_stride + gs);
            p.r.z = (code > 0) ? __ldg(prior_site + (code - 1)) : 0.0;
            leaf_operand(p, code, s_lut);
        }
        return p;
    };

    for (;;) {
        if (tid < 5) s_bc[tid] = consts_of(s_d[tid], nu);
        __syncthreads();
        const double4 cv = s_bc[0], cu = s_bc[1], cC = s_bc[2], cA = s_bc[3], cB = s_bc[4];
        double acc[26];
#pragma unroll
        for (int k = 0; k < 26; k++) acc[k] = 0.0;
        for (int s = tid; s < P.K; s += NNIL_THREADS) {
            const int gs = P.site0 + s;
            const double *prior_site = P.prior + __ldg(P.site_ptr + gs);
            const double qsum = __ldg(P.qsum + gs);
            const Opd oA = operand(cd.A, s, gs, prior_site), oB = operand(cd.B, s, gs, prior_site), oC = operand(cd.C, s, gs, prior_site);
            int shift = __double2hiint(oA.r.w) + __double2hiint(oB.r.w) + __double2hiint(oC.r.w);   // exponent of L' relative to the records' scales
            // T^v = out[g] .* E^s, or e_0 when v is a root
            TRec T;
            T.T0 = 1.0; T.Tm = 0.0; T.tB = 0.0; T.tD = 0.0; T.tk = -1;
            double invL = 1.0;
            if (cd.g >= 0) {
                const unsigned og = rec_index(P, cd.g, s);
                Opd o;
                o.r = P.O[og]; o.ri = P.OI[og];
                const Opd oS = operand(cd.s, s, gs, prior_site);
                Child cs;
                decode(cs, oS, bcs, qsum);
                T = make_T(o, cs);
                shift += o.ri.y + __double2hiint(oS.r.w);        // F_g + cum_s
                invL = P.invL[(size_t)cd.root * P.K + s];
            }
            Child a, b, c, u, v;
            decode(a, oA, cA, qsum); decode(b, oB, cB, qsum); decode(c, oC, cC, qsum);
            const Opd xu = merge_children(a, b);
            decode(u, xu, cu, qsum);
            const Opd xv = merge_children(u, c);
            decode(v, xv, cv, qsum);
            // L' = sum_j T^v_j E^v'_j
            const double sqE = fma(v.Eb, qsum, (v.Eh - v.Eb) * v.qh);
            const double Lloc = T.T0 * v.E0 + T.Tm * v.xm + fma(T.tB, sqE, T.tD * ((T.tk == v.h) ? v.Eh : v.Eb));
            if (Lloc > 0.0) {
                const double iL = 1.0 / Lloc;
                Opd ov, ou, unused;
                double cnt[25];
                finish_T(T, v, cv, qsum, iL, ov, cnt);                        // branch v'
                child_outside(ov, u, c, cu, qsum, iL, ou, cnt + 5);           // branch u'
                child_outside(ov, c, u, cC, qsum, iL, unused, cnt + 10);      // branch C
                child_outside(ou, a, b, cA, qsum, iL, unused, cnt + 15);      // branch A
                child_outside(ou, b, a, cB, qsum, iL, unused, cnt + 20);      // branch B
#pragma unroll
                for (int k = 0; k < 25; k++) acc[k] += cnt[k];
                // log L' of this tree at this site: relative to the cached L when v is not a root
                const double ln2 = 0.693147180559945309417232121458;
                acc[25] += (cd.g >= 0) ? ((invL > 0.0) ? log(Lloc * invL) + (double)shift * ln2 : -INFINITY)
                                       : log(Lloc) + (double)shift * ln2;
            } else acc[25] += -INFINITY;
        }
        // fixed-order block sum: butterfly inside each warp, then the warps in order
#pragma unroll
        for (int k = 0; k < 26; k++) {
            double x = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) s_red[warp][k] = x;
        }
        __syncthreads();
        if (tid < 26) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < NNIL_THREADS / 32; w++) sum += s_red[w][tid];
            s_tot[tid] = sum;
        }
        __syncthreads();
        if (tid == 0) {
            // llh of the forest with this tree replaced: rooted moves recompute the whole tree's likelihood,
            // others move it by sum log(L'/L)
            const double llh = (cd.g >= 0) ? llh_total + s_tot[25] : llh_total - root_llh_sum[cd.root] + s_tot[25];
            st.llh_prev = st.llh_cur;
            st.llh_cur = llh;
            if (st.iters > 0 && (fabs((llh - st.llh_prev) / st.llh_prev) < conv_eps || st.iters >= max_iter || !(llh > -INFINITY)))
                s_done = 1;                                 // EM_solver.py:503; s_d is the point llh was evaluated at
            else st.iters++;
        }
        __syncthreads();
        if (s_done) break;
        if (ULTRA) {
            if (tid < 5) {                                  // statistics clamp / rescale, EM_solver.py:343-346
                double tot = 0.0;
#pragma unroll
                for (int k = 0; k < 5; k++) { s_S[tid * 5 + k] = fmax(s_tot[tid * 5 + k], eps_s); tot += s_S[tid * 5 + k]; }
#pragma unroll
                for (int k = 0; k < 5; k++) s_S[tid * 5 + k] = s_S[tid * 5 + k] / tot * (double)K_total;
            }
            __syncthreads();
            if (tid < 32) {                                 // the first warp: lanes 0-4 take one edge each inside the solve
                double d5[5];
                const bool warm = __shfl_sync(0xffffffffu, (int)(st.iters > 1), 0) != 0;       // first M-step cold (st lives in thread 0)
                const int rc = ultra2_solve_dev(ultra + ci, s_S, nu, eps_nu, dmin, dmax, warm, s_H, d5);
                if (tid == 0) {
                    if (rc != 0) s_done = 2;
                    else {
#pragma unroll
                        for (int b = 0; b < 5; b++) s_d[b] = d5[b];
                    }
                }
            }
            __syncthreads();
            if (s_done) break;
        } else {
            if (tid < 5) {
                double x[5], tot = 0.0;
#pragma unroll
                for (int k = 0; k < 5; k++) { x[k] = fmax(s_tot[tid * 5 + k], eps_s); tot += x[k]; }
#pragma unroll
                for (int k = 0; k < 5; k++) x[k] = x[k] / tot * (double)K_total;
                s_d[tid] = mstep_branch_length(x, nu, use_nu, dmin, dmax, s_d[tid], s_end);
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < 5; b++) st.d[b] = s_d[b];
        st.done = s_done;
        state[ci] = st;
    }
}

}  // namespace pmm
