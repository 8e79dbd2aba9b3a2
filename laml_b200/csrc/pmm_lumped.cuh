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
// So a (node, site) is the record (x0, xm, xz, qh | h, cum) -- xz is xb when there is no hot state and the hot
// value when there is one, qh = q_h, cum the accumulated power-of-two exponent; 32 bytes, because x_m (zero
// with a hot state) and q_h (absent without one) share a slot -- and the 40-byte record (o0, om, oB, oD | k, F)
// for the outside pass.  Every sum over the alphabet has a closed form in qsum = sum_alpha q_alpha and q_h:
// SURVEY.md section 8's "exact lumping" (states no leaf of the subtree shows are indistinguishable) applied
// per (node, site) instead of per site.  The numpy model tests/lumped_model.py restates these closed forms
// and is pinned to the reference's dumps.
//
// Execution: sites are independent given the tree, so ONE THREAD OWNS ONE SITE and walks a whole
// CLADE (a connected piece of the tree with < 128 internal nodes, cut out by the host) without any
// synchronisation: a CTA is (clade, 128 consecutive sites); its threads follow the clade's step
// list -- post-order for the inside pass, pre-order for the outside pass -- so that
//   * the record of the child visited last (inside) / the out-record of the child visited first
//     (outside) stays in registers: no memory round trip on the dependency chain;
//   * every other operand was written at least one step earlier (by this very thread, or by an
//     earlier launch for children in other clades) and is loaded one step ahead into a second
//     register set (the loop is unrolled by two so that the sets swap roles without moves); leaf
//     codes are loaded two steps ahead because the prior of a leaf's state is a dependent gather;
//   * the step list and the branch constants of the clade sit in shared memory;
//   * all global accesses are coalesced across the warp's 32 sites (32 B, or 32 B + 8 B, per thread);
//   * the outside step exists in one straight-line variant per shape (which children are leaves).
// The host cuts the forest into tiers (pmm_engine.cu: build_lumped_schedule): tier 0 = every maximal
// subtree below a size threshold (one launch, ~97 % of the nodes), then the small top tree peeled the
// same way: a 100k-cell tree needs 5 launches per pass instead of one per tree level.
// Expected counts: the 10 per-thread values of a step (5 per child branch) are summed over the
// warp's sites through a transposed shared-memory tile and written once per (warp, node) into
// Spart[site group][node][5]; pmm_reduce_S sums the site groups in a fixed order (bit-reproducible).
#pragma once
#ifdef PMM_HOST_EMU                      // host build of the kernels under tests/native/cuda_host_shim.h
#include "pmm_common.cuh"
#else
#include "pmm_kernels.cuh"
#endif
#include <type_traits>

namespace pmm {

#ifndef PMM_LT
#define PMM_LT 128
#endif
constexpr int LT = PMM_LT;               // threads (= sites) per CTA
constexpr int STEP_CAP = 128;            // a clade has fewer steps than this (the host caps the clade size at 64)
constexpr int STEP_PAD = 8;              // entries past the end of the staged step list (copies of the last step: look-ahead needs no guard)
#ifndef PMM_LIN_MINB
#define PMM_LIN_MINB 5
#endif
#ifndef PMM_LOUT_MINB
#define PMM_LOUT_MINB 3
#endif
#ifndef PMM_PF_DIST
#define PMM_PF_DIST 6                    // L2 prefetch distance in steps for operands of earlier launches (0 = off)
#endif
constexpr int PF_DIST = PMM_PF_DIST;
#ifndef PMM_PF_DIST_OUT
#define PMM_PF_DIST_OUT 3
#endif
constexpr int PF_DIST_OUT = PMM_PF_DIST_OUT;
#ifndef PMM_LD256
#define PMM_LD256 1                      // records move as one 256-bit access (LDG.E.ENL2.256 / STG.E.ENL2.256): -3 % on both passes
#endif

struct LParams {
    int K;                   // sites of this chunk
    int Kp;                  // K rounded up to a multiple of 32: row stride of the record arrays, so that the idle
                             // lanes of the last warp own (and privately read / write) padding records
    int site0;               // first site of the chunk in the handle's site space
    int code_stride;         // row stride of `codes` (all sites of the handle)
    int n_leaves, n_nodes, n_clades;
    const int *codes;        // [n_leaves][code_stride] raw state codes
    const int *site_ptr;     // [K_handle + 1] CSR offsets into `prior`
    const double *prior;     // [nnz]
    const double *qsum;      // [K_handle] sum_alpha q_alpha
    const double *params;    // [0] = nu, [1] = phi
    const double *bc;        // [n][4] = a, b, c, g per branch (pmm_branch_consts)
    double2 *X;              // [n_int][K][2] inside records, 32 bytes: (x0, xz), (t, h | cum) -- t = q_h with a hot state
                             // (then x_m = 0 exactly), x_m without one
    double4 *O; int2 *OI;    // [n_int][K] outside records (o0, om, oB, oD), (k, F)
    double *invL;            // [n_roots][K] 1 / (normalised site likelihood)
    double *root_site_llh;   // [n_roots][K]
    double *Spart;           // [G][n][5], G = ceil(K / 32) site groups (one per warp)
    double *Rpart;           // [G][n_clades]
    double *post0, *post1;   // optional [n][K_total] log posteriors (annotation path), or null
    int K_total;
    int store_all;           // outside pass: also store the out-records that are chained in registers
    int sblk0;               // first site block (of LT sites) this launch covers: site slices run on separate streams
    // ---- dataflow launches (several tiers of the clade schedule in ONE launch; "dataflow" below).  Unused otherwise.
    unsigned *df_ticket;     // ticket counter of this launch site ...
    unsigned df_base;        // ... and its value when the launch starts: a CTA's item is its ticket minus this
    int *df_flags;           // [clade position][df_stride] completion marks of (clade, site block): == df_epoch once done
    const int *df_dep_ptr;   // CSR over clade positions: the clades whose records (inside) / out-records (outside) a clade loads
    const int *df_deps;
    int df_lo;               // clade position of the launch's first clade; producers below it ran in earlier launches
    int df_n1, df_n2;        // clades in the launch's first / second segment (inside: tier 0 | top tiers, outside: top tiers | tier 0)
    int df_nblk;             // site blocks the launch covers (from sblk0)
    int df_stride;           // row stride of df_flags = site blocks of a chunk
    int df_epoch;            // value that marks "done" for this launch (the host counts launches)
    int *df_err;             // set when a wait gives up (a schedule bug must not hang the GPU)
};

enum { KIND_LEAF = 0, KIND_LOAD = 1, KIND_PREV = 2 };

// index of (internal node, site) in the record arrays.  32 bits: a chunk never holds 2^31 records (they would
// need 172 GB; the host caps the chunk accordingly), and 32-bit indices halve the address arithmetic
__device__ __forceinline__ unsigned rec_index(const LParams &P, int node, int s) {
    return (unsigned)(node - P.n_leaves) * (unsigned)P.Kp + (unsigned)s;
}

// one prefetched operand: an inside / outside record, or (leaves) the prior of the leaf's state in r.w
struct Opd { double4 r; int2 ri; };

// An INSIDE operand is kept in the packed form of the 32-byte record: r = (x0, xz, t, h | cum) -- a record with a
// hot state has x_m = 0 exactly and one without has no q_h, so the two share the slot t; h and cum ride in the
// last 8 bytes.  Nothing is unpacked at load time (that would wait for the load): `decode` does it at the use.
#if PMM_LD256 && !defined(PMM_HOST_EMU)
// 32-byte records are 32-byte aligned (cudaMalloc base, rows of whole warps): one 256-bit access each
__device__ __forceinline__ double4 ld256(const void *p) {
    double4 r;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st256(void *p, const double4 r) {
    asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" ::"d"(r.x), "d"(r.y), "d"(r.z), "d"(r.w), "l"(p) : "memory");
}
#else
__device__ __forceinline__ double4 ld256(const void *p) {
    const double2 a = reinterpret_cast<const double2 *>(p)[0], b = reinterpret_cast<const double2 *>(p)[1];
    return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st256(void *p, const double4 r) {
    reinterpret_cast<double2 *>(p)[0] = make_double2(r.x, r.y);
    reinterpret_cast<double2 *>(p)[1] = make_double2(r.z, r.w);
}
#endif
__device__ __forceinline__ void ld_xrec(Opd &p, const double2 *__restrict__ X, unsigned o) { p.r = ld256(X + 2 * (size_t)o); }
__device__ __forceinline__ void st_xrec(double2 *__restrict__ X, unsigned o, const Opd &p) { st256(X + 2 * (size_t)o, p.r); }
__device__ __forceinline__ void ld_orec(Opd &p, const double4 *__restrict__ O, const int2 *__restrict__ OI, unsigned o) {
    p.r = ld256(O + o); p.ri = OI[o];
}
__device__ __forceinline__ void st_orec(double4 *__restrict__ O, int2 *__restrict__ OI, unsigned o, const Opd &p) {
    st256(O + o, p.r); OI[o] = p.ri;
}
__device__ __forceinline__ double pack_hc(int h, int cum) { return __hiloint2double(cum, h); }

__device__ __forceinline__ double4 ld_bc(const double *__restrict__ bc, int v) {
    const double2 lo = __ldg(reinterpret_cast<const double2 *>(bc) + 2 * (size_t)v);
    const double2 hi = __ldg(reinterpret_cast<const double2 *>(bc) + 2 * (size_t)v + 1);
    return make_double4(lo.x, lo.y, hi.x, hi.y);
}
#ifdef PMM_HOST_EMU
__device__ __forceinline__ void prefetch_l2(const void *) {}
__device__ __forceinline__ void grid_dependency_wait() {}
__device__ __forceinline__ void grid_launch_dependents() {}
#else
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// programmatic dependent launch: the next launch of the stream may start (and stage its clade) while this grid
// drains; it must not touch records before grid_dependency_wait() (no-ops without the launch attribute)
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ---------------------------------------------------------------- dataflow launches
// The tiers of the clade schedule depend on each other only clade by clade: a clade of the inside pass needs the records
// of the lower-tier clades it loads from, a clade of the outside pass the out-records its parent clade stored.  One
// launch per tier turns that into grid-wide barriers -- a handful of latency-bound launches whose chains of <= 16 steps
// do not shrink with the site shard (8 GPUs: a fifth of the E-step) or with the tree (1k cells: all of it).  A DATAFLOW
// launch covers several tiers at once: its work items are (clade, site block) pairs in an order in which every producer
// precedes its consumers; a CTA takes the next item with an atomic ticket (so the order holds whatever order the
// hardware starts CTAs in: every earlier item is held by a CTA that is running or done, hence no deadlock), stages its
// clade, waits until the marks of its producer clades show this launch's epoch, walks the clade, and publishes its own
// mark (bar.sync, then fence + release store by one thread).  Item order, for n1 clades in the launch's first segment
// and n2 in its second (inside: tier 0 | top tiers; outside: top tiers | tier 0) over B site blocks:
//     first(0), [first(1), second(0)], [first(2), second(1)], ..., second(B-1)
// so that the latency-bound top tiers of one site block run beside the bulk of the next.
template <bool DF> __device__ __forceinline__ void df_item(const LParams &P, int &cl, int &blk) {
    if (!DF) { cl = blockIdx.x; blk = blockIdx.y; return; }
    __shared__ int s_item;
    if (threadIdx.x == 0) s_item = (int)(atomicAdd(P.df_ticket, 1u) - P.df_base);
    __syncthreads();
    const int i = s_item, n1 = P.df_n1, n12 = P.df_n1 + P.df_n2;
    if (i < n1) { cl = i; blk = 0; return; }
    const int j = i - n1, g = j / n12, r = j - g * n12;
    if (g < P.df_nblk - 1) { cl = r; blk = (r < n1) ? g + 1 : g; }
    else { cl = n1 + r; blk = P.df_nblk - 1; }
}
__device__ __forceinline__ void df_spin(const int *flag, int epoch, int *err) {
#ifdef PMM_HOST_EMU
    if (*flag != epoch) *err = 1;            // CTAs run one after another in ticket order: a producer that has not run is a schedule bug
#else
    const long long t0 = clock64();
    for (unsigned polls = 1;; polls++) {
        int v;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v == epoch) break;
        // ~1 s without the mark: give up and flag the error (the finishing kernel turns it into a NaN likelihood); once
        // one wait has given up the others follow at once
        if ((polls & 1023u) == 0 && (clock64() - t0 > (1ll << 31) || *(volatile int *)err)) { *err = 1; break; }
        __nanosleep(40);
    }
#endif
}
// wait for the producers of the clade at position `pos` (site block `blk` of the chunk); all threads of the CTA call it
template <bool DF> __device__ __forceinline__ void df_wait(const LParams &P, int pos, int blk) {
    if (!DF) return;
    const int b = P.df_dep_ptr[pos], e = P.df_dep_ptr[pos + 1];
    for (int t = b + (int)threadIdx.x; t < e; t += LT) {
        const int dp = P.df_deps[t];
        if (dp >= P.df_lo) df_spin(P.df_flags + (size_t)dp * P.df_stride + blk, P.df_epoch, P.df_err);
    }
    __syncthreads();
}
// every thread's records are written: publish the mark of (pos, blk)
template <bool DF> __device__ __forceinline__ void df_done(const LParams &P, int pos, int blk) {
    if (!DF) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        int *flag = P.df_flags + (size_t)pos * P.df_stride + blk;
#ifdef PMM_HOST_EMU
        *flag = P.df_epoch;
#else
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(P.df_epoch) : "memory");
#endif
    }
}

// the clade's steps, the branch constants of every step's two children and the four leaf vectors
// (EM_solver.py:103-116 in full-state form, indexed by min(code, 1) + 2:
//   known silenced (-2) -> e_m   '?' (-1) -> (phi, 1, phi..)   0 -> (1-phi) e_0   a -> (1-phi) e_a)
__device__ __forceinline__ void stage_clade(int4 *s_steps, double4 *s_bc, double4 *s_lut, const LParams &P,
                                            const int4 *__restrict__ steps, int begin, int nsteps) {
    for (int i = threadIdx.x; i < nsteps + STEP_PAD; i += LT) s_steps[i] = __ldg(steps + begin + min(i, nsteps - 1));
    for (int i = threadIdx.x; i < 2 * nsteps; i += LT) {
        const int4 d = __ldg(steps + begin + (i >> 1));
        s_bc[i] = ld_bc(P.bc, (i & 1) ? d.z : d.y);
    }
    if (threadIdx.x < 4) {
        const double phi = __ldg(P.params + 1);
        const int c = threadIdx.x;                  // (x0, xm, xz)
        s_lut[c] = make_double4(c == 1 ? phi : (c == 2 ? 1.0 - phi : 0.0), c < 2 ? 1.0 : 0.0, c == 1 ? phi : (c == 3 ? 1.0 - phi : 0.0),
                                c == 1 ? phi : 0.0);        // (x0, xm, xz, xb)
    }
    __syncthreads();
}

// decoded child: record fields plus the message to the parent (SURVEY 3.5 "edge transfer"):
//   E_0 = a x_0 + b sum(q x) + g x_m,  E_m = x_m,  E_alpha = c x_alpha + g x_m  (Eb off the hot slot, Eh on it;
//   Eh == Eb when there is no hot state)
struct Child {
    double x0, xm, xz, qh, xb, sq;      // sq = sum_a q_a x_a
    int h, cum;
    double t0, t1, gm, E0, Eb, Eh;      // t0 = a x0, t1 = b sq, gm = g xm
};
__device__ __forceinline__ void decode(Child &c, const Opd &p, const double4 bc, double qsum) {
    c.x0 = p.r.x; c.xz = p.r.y; c.qh = p.r.z;                 // qh is meaningful only with a hot state
    c.h = __double2loint(p.r.w); c.cum = __double2hiint(p.r.w);
    const bool hot = c.h >= 0;
    c.xm = hot ? 0.0 : p.r.z;
    c.xb = hot ? 0.0 : c.xz;
    c.sq = c.xz * (hot ? c.qh : qsum);
    c.gm = bc.w * c.xm;
    c.t0 = bc.x * c.x0;
    c.t1 = bc.y * c.sq;
    c.E0 = c.t0 + c.t1 + c.gm;
    c.Eb = fma(bc.z, c.xb, c.gm);
    c.Eh = fma(bc.z, c.xz, c.gm);
}
// a leaf decoded straight from its code (no packing round trip): (x0, xm, xz, xb) from the table, qh = the gathered prior
__device__ __forceinline__ void decode_leaf(Child &c, int code, double qh, const double4 bc, double qsum, const double4 *s_lut) {
    const double4 t = s_lut[min(code, 1) + 2];
    c.x0 = t.x; c.xm = t.y; c.xz = t.z; c.xb = t.w; c.qh = qh;
    c.h = max(code, 0) - 1; c.cum = 0;
    c.sq = c.xz * ((code > 0) ? qh : qsum);
    c.gm = bc.w * c.xm;
    c.t0 = bc.x * c.x0;
    c.t1 = bc.y * c.sq;
    c.E0 = c.t0 + c.t1 + c.gm;
    c.Eb = fma(bc.z, c.xb, c.gm);
    c.Eh = fma(bc.z, c.xz, c.gm);
}
// a leaf's operand from its code: (x0, xm, xz) from the table; q of its state was gathered into p.r.z one step ahead
__device__ __forceinline__ void leaf_operand(Opd &p, int code, const double4 *s_lut) {
    const double4 t = s_lut[min(code, 1) + 2];
    p.r.x = t.x; p.r.y = t.z;
    p.r.z = (code > 0) ? p.r.z : t.y;
    p.r.w = pack_hc(max(code, 0) - 1, 0);
}

// ---------------------------------------------------------------- inside pass
// step flags: bits 0-1 kind of child v (leaf / load), bits 2-3 kind of child w (leaf / load / chained:
//             the host orders the children so that a chained child is w), bit 4 = u is a root,
//             bit 5 / 6: the record of v / w was written by an earlier launch
struct InSlot { Opd v, w; };

__device__ __forceinline__ void inside_fetch(const LParams &P, const int4 d, int cv, int cw, InSlot &n, const double *__restrict__ prior_site,
                                             int s, int N) {
    const int kv = d.w & 3, kw = (d.w >> 2) & 3;
    if (kv == KIND_LEAF) n.v.r.z = (cv > 0) ? __ldg(prior_site + (cv - 1)) : 0.0;
    else { const unsigned o = rec_index(P, d.y, s); ld_xrec(n.v, P.X, o); }
    if (kw == KIND_LEAF) n.w.r.z = (cw > 0) ? __ldg(prior_site + (cw - 1)) : 0.0;
    else if (kw == KIND_LOAD) { const unsigned o = rec_index(P, d.z, s); ld_xrec(n.w, P.X, o); }
}

// one thread's walk through the clade `cl` (steps [x, y), root index z) at its site of block `blk`
__device__ __forceinline__ void inside_walk(const LParams &P, const int4 cl, int blk, const int4 *s_steps, const double4 *s_bc,
                                            const double4 *s_lut) {
    const int nsteps = cl.y - cl.x;
    const int N = P.n_leaves;
    const int sr = (blk + P.sblk0) * LT + threadIdx.x;            // record column (idle lanes own padding columns)
    if ((sr & ~31) >= P.K) return;                                   // whole warp beyond the last site
    const bool valid = sr < P.K;
    const int s = min(sr, P.K - 1);                                   // inputs of an idle lane: those of the last site
    const int gs = P.site0 + s;
    const double *prior_site = P.prior + __ldg(P.site_ptr + gs);
    const double qsum = __ldg(P.qsum + gs);
    const int *codes_s = P.codes + gs;

    auto codes_of = [&](const int4 d, int &cv, int &cw) {
        if ((d.w & 3) == KIND_LEAF) cv = __ldg(codes_s + (size_t)d.y * P.code_stride);
        if (((d.w >> 2) & 3) == KIND_LEAF) cw = __ldg(codes_s + (size_t)d.z * P.code_stride);
    };
    Opd prev;                                            // the record computed by the previous step
    prev.r = make_double4(0.0, 0.0, 0.0, pack_hc(-1, 0)); prev.ri = make_int2(-1, 0);
    InSlot A, B;
    A.v = A.w = B.v = B.w = prev;
    int4 d0 = s_steps[0], d1 = (nsteps > 1) ? s_steps[1] : d0;
    int c0v = 0, c0w = 0, c1v = 0, c1w = 0;
    codes_of(d0, c0v, c0w);
    if (nsteps > 1) codes_of(d1, c1v, c1w);
    // the first PF_DIST steps have nobody to prefetch for them: pull their far operands into L2 now.  Everything up to
    // here reads inputs only, so it runs before the dependency wait, while the previous launch drains (a top-tier clade
    // has a dozen steps: ncu shows its warps waiting for DRAM at the first use of nearly every loaded record)
    for (int i = 0; i < PF_DIST && i < nsteps; i++) {
        const int4 dp = s_steps[i];
        if ((dp.w & 3) == KIND_LEAF) prefetch_l2(codes_s + (size_t)dp.y * P.code_stride);
        else if (dp.w & 32) { const unsigned o = rec_index(P, dp.y, sr); prefetch_l2(P.X + 2 * (size_t)o); }
        if (((dp.w >> 2) & 3) == KIND_LEAF) prefetch_l2(codes_s + (size_t)dp.z * P.code_stride);
        else if (dp.w & 64) { const unsigned o = rec_index(P, dp.z, sr); prefetch_l2(P.X + 2 * (size_t)o); }
    }
    grid_dependency_wait();
    inside_fetch(P, d0, c0v, c0w, A, prior_site, sr, N);

    auto step = [&](int i, InSlot &cur, InSlot &nxt) {
        int4 d2 = d1;
        int c2v = 0, c2w = 0;
        if (i + 2 < nsteps) { d2 = s_steps[i + 2]; codes_of(d2, c2v, c2w); }
        if (i + 1 < nsteps) inside_fetch(P, d1, c1v, c1w, nxt, prior_site, sr, N);
        if (PF_DIST > 0 && i + PF_DIST < nsteps) {
            const int4 dp = s_steps[i + PF_DIST];
            if ((dp.w & 3) == KIND_LEAF) prefetch_l2(codes_s + (size_t)dp.y * P.code_stride);
            else if (dp.w & 32) { const unsigned o = rec_index(P, dp.y, sr); prefetch_l2(P.X + 2 * (size_t)o); }
            if (((dp.w >> 2) & 3) == KIND_LEAF) prefetch_l2(codes_s + (size_t)dp.z * P.code_stride);
            else if (dp.w & 64) { const unsigned o = rec_index(P, dp.z, sr); prefetch_l2(P.X + 2 * (size_t)o); }
        }
        const int kv = d0.w & 3, kw = (d0.w >> 2) & 3;
        Child v, w;
        if (kv == KIND_LEAF) decode_leaf(v, c0v, cur.v.r.z, s_bc[2 * i], qsum, s_lut); else decode(v, cur.v, s_bc[2 * i], qsum);
        if (kw == KIND_LEAF) decode_leaf(w, c0w, cur.w.r.z, s_bc[2 * i + 1], qsum, s_lut);
        else decode(w, (kw == KIND_PREV) ? prev : cur.w, s_bc[2 * i + 1], qsum);
        // x_u = E^v .* E^w  (EM_solver.py:118-133); Eh == Eb for a child without a hot state, and two
        // different hot states leave no edited slot alive (E^w_b = 0 then)
        const bool bv = v.h >= 0, bw = w.h >= 0;
        const bool clash = bv && bw && v.h != w.h;
        const double x0 = v.E0 * w.E0, xm = v.xm * w.xm;
        const double xz = v.Eh * (clash ? w.Eb : w.Eh);
        const int e = expo_of_hi(max(hi32(x0), max(hi32(xm), hi32(xz))));
        const double sc = pow2i(-e);
        const int hu = bv ? (clash ? -1 : v.h) : w.h;
        prev.r.x = x0 * sc; prev.r.y = xz * sc;
        prev.r.z = (hu >= 0) ? (bv ? v.qh : w.qh) : xm * sc;
        prev.r.w = pack_hc(hu, v.cum + w.cum + e);
        st_xrec(P.X, rec_index(P, d0.x, sr), prev);               // idle lanes write their own padding column
        if (d0.w & 16) {
            // root: site likelihood L = E[root]_0 (the root's own branch starts in state 0, EM_solver.py:140) ...
            const double4 cr = ld_bc(P.bc, d0.x);
            Child r;
            decode(r, prev, cr, qsum);
            const double L = r.E0, iL = (valid && L > 0.0) ? 1.0 / L : 0.0;
            if (valid) {
                P.invL[(size_t)cl.z * P.K + s] = iL;
                P.root_site_llh[(size_t)cl.z * P.K + s] = (L > 0.0) ? log(L) + (double)r.cum * 0.693147180559945309417232121458 : -INFINITY;
            }
            // ... and everything the outside pass needs to know about a root, so that a root is an ordinary step there:
            // out[root] = (a, g, b q) (EM_solver.py:151-158) as a stored out-record, and the counts of the root's own
            // branch (:282-286), summed over the warp's sites (a rare step: plain shuffles)
            Opd seed;
            seed.r = make_double4(cr.x, cr.w, cr.y, 0.0); seed.ri = make_int2(-1, -r.cum);
            st_orec(P.O, P.OI, rec_index(P, d0.x, sr), seed);
            double t0 = r.t0 * iL, t1 = r.t1 * iL, t2 = r.gm * iL;
            if (P.post0 && valid) {
                const size_t gp = (size_t)d0.x * P.K_total + P.site0 + s;
                P.post0[gp] = (t0 > 0.0) ? log(t0) : -INFINITY;
                P.post1[gp] = (t2 > 0.0) ? log(t2) : -INFINITY;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                t0 += __shfl_xor_sync(0xffffffffu, t0, o); t1 += __shfl_xor_sync(0xffffffffu, t1, o); t2 += __shfl_xor_sync(0xffffffffu, t2, o);
            }
            if ((threadIdx.x & 31) < 5) {
                const int j = threadIdx.x & 31;
                P.Spart[((size_t)(sr >> 5) * P.n_nodes + d0.x) * 5 + j] = (j == 0) ? t0 : (j == 1) ? t1 : (j == 2) ? t2 : 0.0;
            }
        }
        d0 = d1; d1 = d2; c0v = c1v; c0w = c1w; c1v = c2v; c1w = c2w;
    };
    int i = 0;
    for (; i + 1 < nsteps; i += 2) { step(i, A, B); step(i + 1, B, A); }
    if (i < nsteps) step(i, A, B);
    grid_launch_dependents();
}

// grid = (clades of the tier, site blocks); DF: a dataflow launch over several tiers, grid = clades x site blocks in one dimension
template <bool DF = false>
__global__ void __launch_bounds__(LT, PMM_LIN_MINB) pmm_lumped_inside(LParams P, const int4 *__restrict__ clades, const int4 *__restrict__ steps) {
    __shared__ int4 s_steps[STEP_CAP + STEP_PAD];
    __shared__ double4 s_bc[2 * STEP_CAP];
    __shared__ double4 s_lut[4];
    int ci, blk;
    df_item<DF>(P, ci, blk);
    const int4 cl = clades[ci];
    stage_clade(s_steps, s_bc, s_lut, P, steps, cl.x, cl.y - cl.x);
    if (DF) grid_dependency_wait();                     // (otherwise inside the walk, after its look-ahead on the inputs)
    df_wait<DF>(P, P.df_lo + ci, blk + P.sblk0);
    inside_walk(P, cl, blk, s_steps, s_bc, s_lut);
    df_done<DF>(P, P.df_lo + ci, blk + P.sblk0);
}

// ---------------------------------------------------------------- outside pass + expected counts
// For child v of u with sibling w (EM_solver.py:143-300 in lumped full-state form):
//   T^v = out[u] .* E^w;  out[v]_0 = a T_0, out[v]_a = b q_a T_0 + c T_a, out[v]_m = g (T_0 + sum_a T_a) + T_m
//   S0 = T_0 a x_0 / L, S1 = T_0 b sum(q x) / L, S2 = T_0 g x_m / L, S3 = c sum_a T_a x_a / L, S4 = g x_m sum_a T_a / L.
// T has the same lumped shape as an out-record: T_a = tB q_a + [a == tk] tD  (if w has a hot state E^w vanishes
// off it, so tB = 0 and the hot slot is w's; otherwise E^w is flat and u's hot slot, if any, survives).
struct TRec { double T0, Tm, tB, tD; int tk; };
__device__ __forceinline__ TRec make_T(const Opd &o, const Child &w) {
    TRec t;
    const double oB = o.r.z, oD = o.r.w;
    const bool hotw = w.h >= 0;
    t.T0 = o.r.x * w.E0;
    t.Tm = o.r.y * w.xm;
    t.tB = oB * w.Eb;
    t.tD = w.Eh * (hotw ? fma(oB, w.qh, (o.ri.x == w.h) ? oD : 0.0) : oD);
    t.tk = hotw ? w.h : o.ri.x;
    return t;
}
// out-record of v (unnormalised) and the five counts of the branch above v (times fac) from T^v
template <bool REC = true>                                               // REC = false: v is a leaf, only out[v]_0 and out[v]_m are formed
__device__ __forceinline__ void finish_T(const TRec &t, const Child &v, const double4 bcv, double qsum, double fac,
                                         Opd &n, double *__restrict__ cnt) {
    const double sT = fma(t.tB, qsum, t.tD);                               // sum_a T_a   (tD = 0 without a hot slot)
    n.r.x = bcv.x * t.T0;
    n.r.y = fma(bcv.w, t.T0 + sT, t.Tm);
    if (REC) {
        n.r.z = fma(bcv.y, t.T0, bcv.z * t.tB);
        n.r.w = bcv.z * t.tD;
        n.ri.x = t.tk;
    }
    const double xvk = (t.tk == v.h) ? v.xz : v.xb;                        // x^v at slot tk
    const double sTx = fma(t.tB, v.sq, t.tD * xvk);                        // sum_a T_a x^v_a
    const double T0f = t.T0 * fac, gmf = v.gm * fac;
    cnt[0] = T0f * v.t0;
    cnt[1] = T0f * v.t1;
    cnt[2] = t.T0 * gmf;
    cnt[3] = bcv.z * sTx * fac;
    cnt[4] = gmf * sT;
}
template <bool REC = true>
__device__ __forceinline__ void child_outside(const Opd &o, const Child &v, const Child &w, const double4 bcv, double qsum, double fac,
                                              Opd &n, double *__restrict__ cnt) {
    finish_T<REC>(make_T(o, w), v, bcv, qsum, fac, n, cnt);
}

__device__ __forceinline__ void normalise_out(Opd &n, int F) {
    const int e = expo_of_hi(max(max(hi32(n.r.x), hi32(n.r.y)), max(hi32(n.r.z), hi32(n.r.w))));
    const double sc = pow2i(-e);
    n.r.x *= sc; n.r.y *= sc; n.r.z *= sc; n.r.w *= sc;
    n.ri.y = F + e;
}

// Expected counts: a step yields 10 values per lane (5 per child branch) that must be summed over the warp's 32 sites.
// They are buffered for RED_STEPS steps in a shared-memory tile, one ROW per (step, value), lane l in column l -- the
// stores of a warp are 256 contiguous bytes, conflict-free -- and then every lane that has a row adds it up (16 LDS.128 at a
// row stride of 34 doubles = 4 banks mod 32: conflict-free per quarter-warp; four independent chains of eight adds) and
// writes its sum straight to Spart.  ncu on the previous form (ten rows reduced at the end of EVERY step by twenty lanes,
// two chains of eight adds behind the loads): a fifth of the kernel's stall samples sat on those dependent adds.
#ifndef PMM_RED_STEPS
#define PMM_RED_STEPS 2                 // measured on config 4: 1 -> 3.29 ms, 2 -> 2.83 ms, 3 -> 2.86 ms (outside pass)
#endif
constexpr int RED_STEPS = PMM_RED_STEPS;
constexpr int RED_ROWS = 10 * RED_STEPS;             // <= 32: one lane per row
constexpr int RED_ROW = 34;
__device__ __forceinline__ void red_put5(const double *__restrict__ v, double *__restrict__ red, int lane, int first) {
#pragma unroll
    for (int j = 0; j < 5; j++) red[(first + j) * RED_ROW + lane] = v[j];
}
// sums of the rows of the `cnt` buffered steps (the last of them is step i_last of the staged list) -> Spart
__device__ __forceinline__ void red_flush(const double *__restrict__ red, int lane, int cnt, int i_last, const int4 *__restrict__ s_steps,
                                          double *__restrict__ Sg) {
    __syncwarp();
    const int r = min(lane, 10 * cnt - 1);
    // lanes without a row sit the flush out: a quarter fewer shared-memory wavefronts (ncu: the L1 data pipe is this kernel's
    // busiest unit, 87 %, two thirds of it shared memory).  Measured against it, both slower: pairs of lanes added with a
    // shuffle before a 16-column tile (2.93 ms), three steps per flush (2.80); this form 2.78, all lanes loading 2.81
    if (lane < 10 * cnt) {
        const double2 *row = reinterpret_cast<const double2 *>(red + r * RED_ROW);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int t = 0; t < 16; t += 2) {
            const double2 x = row[t], y = row[t + 1];
            a0 += x.x; a1 += x.y; a2 += y.x; a3 += y.y;
        }
        const double tot = (a0 + a1) + (a2 + a3);
        const int a = r / 10, j = r - 10 * a;          // buffered step, value (0-4: branch above v, 5-9: above w)
        const int4 d = s_steps[i_last - (cnt - 1) + a];
        Sg[(size_t)((j < 5) ? d.y : d.z) * 5 + ((j < 5) ? j : j - 5)] = tot;
    }
    __syncwarp();                                    // readers done before the next step's writers
}

// step flags: bits 0-1 where out[u] comes from (1 = load, 2 = chained in registers; a root loads the record the inside
//             pass' root step stored), bit 2 / 3: v / w is a leaf, bit 4 / 5: store out[v] / out[w],
//             bit 6: out[v] is chained into the next step (the host orders the children accordingly)
struct OutSlot { Opd v, w, o; double qv, qw; };      // qv / qw: the prior of a leaf child's state (own registers: see ld_child_pred)

// ---- predicated accesses.  The per-step prologue (operands of step i + 1, leaf codes of step i + 2, L2 prefetch of
// step i + PF) is straight-line code: which access a child needs is a warp-uniform predicate (a descriptor flag) or a
// per-lane one (a leaf's code), so every access is issued under a predicate and its address is formed unconditionally.
// ncu on the branching version of this prologue: 18 branches per step, 40 % of the stall samples on the control flow.
// The predicates are formed inside the asm from the raw flag BITS (and a leaf's code), so that no 0/1 integer is ever
// materialised: `bit` arguments are masked flag words, non-zero = true.
#ifdef PMM_HOST_EMU
__device__ __forceinline__ void ld_code_pred(int &c, unsigned bit, const int *a) { if (bit) c = *a; }
__device__ __forceinline__ void ld_child_pred(Opd &n, double &ql, unsigned leafbit, int code, const double *q, const double2 *x) {
    if (leafbit) { if (code > 0) ql = *q; }
    else n.r = ld256(x);
}
__device__ __forceinline__ void ld_orec_pred(Opd &n, unsigned bit, const double4 *o, const int2 *oi) { if (bit) { n.r = ld256(o); n.ri = *oi; } }
__device__ __forceinline__ void prefetch_l2_either(unsigned, const void *, const void *) {}
#else
__device__ __forceinline__ void ld_code_pred(int &c, unsigned bit, const int *a) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p ld.global.nc.s32 %0, [%2];\n\t}" : "+r"(c) : "r"(bit), "l"(a));
}
// leaf: the prior of the leaf's state (lanes with an edited state only) into ql; internal child: its 32-byte inside record.
// The two loads write DIFFERENT registers: predicated loads into overlapping registers serialise on the scoreboard even
// though only one of them ever executes (ncu: the record load waited for the gather, 12 % of the stall samples)
__device__ __forceinline__ void ld_child_pred(Opd &n, double &ql, unsigned leafbit, int code, const double *q, const double2 *x) {
    asm volatile("{\n\t.reg .pred pleaf, pq;\n\tsetp.ne.u32 pleaf, %5, 0;\n\tsetp.gt.and.s32 pq, %6, 0, pleaf;\n\t"
                 "@pq ld.global.nc.f64 %4, [%7];\n\t"
                 "@!pleaf ld.global.v4.f64 {%0,%1,%2,%3}, [%8];\n\t}"
                 : "+d"(n.r.x), "+d"(n.r.y), "+d"(n.r.z), "+d"(n.r.w), "+d"(ql) : "r"(leafbit), "r"(code), "l"(q), "l"(x));
}
__device__ __forceinline__ void ld_orec_pred(Opd &n, unsigned bit, const double4 *o, const int2 *oi) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %6, 0;\n\t"
                 "@p ld.global.v4.f64 {%0,%1,%2,%3}, [%7];\n\t"
                 "@p ld.global.v2.s32 {%4,%5}, [%8];\n\t}"
                 : "+d"(n.r.x), "+d"(n.r.y), "+d"(n.r.z), "+d"(n.r.w), "+r"(n.ri.x), "+r"(n.ri.y) : "r"(bit), "l"(o), "l"(oi));
}
__device__ __forceinline__ void prefetch_l2_either(unsigned bit, const void *a, const void *b) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p prefetch.global.L2 [%1];\n\t@!p prefetch.global.L2 [%2];\n\t}"
                 ::"r"(bit), "l"(a), "l"(b));
}
#endif

// Look-ahead form of the clade's steps, staged once per CTA next to the descriptors: record / code ROWS premultiplied
// (32-bit, unsigned), so that an address in the prologue is one add and one multiply-add:
//   x = (u - N) Kp: row of u's records;  y, z = per child: leaf -> leaf id * code_stride (its code row),
//   internal -> (c - N) Kp (its record row);  w = the step's flags
__device__ __forceinline__ void stage_rows(uint4 *s_rows, const int4 *s_steps, const LParams &P, int nsteps) {
    const unsigned N = (unsigned)P.n_leaves, Kp = (unsigned)P.Kp, cs = (unsigned)P.code_stride;
    for (int i = threadIdx.x; i < nsteps + STEP_PAD; i += LT) {
        const int4 d = s_steps[i];
        auto row = [&](int c) { return ((unsigned)c < N) ? (unsigned)c * cs : ((unsigned)c - N) * Kp; };
        s_rows[i] = make_uint4(((unsigned)d.x - N) * Kp, row(d.y), row(d.z), (unsigned)d.w);
    }
    __syncthreads();
}

// one thread's walk through the clade `cl` (steps [x, y), root index z, clade id w) at its site of block `blk`
template <bool POST>
__device__ __forceinline__ void outside_walk(const LParams &P, const int4 cl, int blk, const int4 *s_steps, const uint4 *s_rows,
                                             const double4 *s_bc, const double4 *s_lut, double *s_red) {
    const int nsteps = cl.y - cl.x;
    const unsigned s_raw = (blk + P.sblk0) * LT + threadIdx.x;           // record column (idle lanes of the last warp own padding columns)
    const bool valid = (int)s_raw < P.K;
    const int s = min((int)s_raw, P.K - 1);
    const int gs = P.site0 + s;
    const int lane = threadIdx.x & 31;
    const int grp = s_raw >> 5;                          // site group of this warp
    if (grp * 32 >= P.K) return;                         // whole warp beyond the last site
    double *red = s_red + (threadIdx.x >> 5) * RED_ROWS * RED_ROW;
    int nbuf = 0;                                        // steps whose counts sit in the tile
    const double *prior_m1 = P.prior + __ldg(P.site_ptr + gs) - 1;       // prior of state a at [a]
    const double qsum = __ldg(P.qsum + gs);
    const int *codes_s = P.codes + gs;
    double *Sg = P.Spart + (size_t)grp * P.n_nodes * 5;

    // leaf codes of a step (two steps ahead: the prior of a leaf's state is a dependent gather)
    auto codes_of = [&](const uint4 r, int &cv, int &cw) {
        ld_code_pred(cv, r.w & 4, codes_s + r.y);
        ld_code_pred(cw, r.w & 8, codes_s + r.z);
    };
    // operands of a step (one step ahead, into the other register set)
    auto fetch = [&](const uint4 r, int cv, int cw, OutSlot &n) {
        ld_child_pred(n.v, n.qv, r.w & 4, cv, prior_m1 + cv, P.X + 2 * (size_t)(r.y + s_raw));
        ld_child_pred(n.w, n.qw, r.w & 8, cw, prior_m1 + cw, P.X + 2 * (size_t)(r.z + s_raw));
        const unsigned o = r.x + s_raw;
        ld_orec_pred(n.o, r.w & 1, P.O + o, P.OI + o);       // how is 1 (load) or 2 (chained): bit 0
    };
    Opd chain;                                           // out-record kept in registers for the next step
    chain.r = make_double4(0.0, 0.0, 0.0, 0.0); chain.ri = make_int2(-1, 0);
    OutSlot A, B;
    A.v = A.w = A.o = B.v = B.w = B.o = chain;
    A.qv = A.qw = B.qv = B.qw = 0.0;
    double rt = 0.0;
    int c0v = 0, c0w = 0, c1v = 0, c1w = 0;
    codes_of(s_rows[0], c0v, c0w);
    codes_of(s_rows[1], c1v, c1w);                       // the staged lists are padded with copies of the last step
    // operands of the first PF_DIST_OUT steps into L2 (nobody prefetches for them), the out-record a step LOADS included;
    // like everything above this reads inputs or only warms L2, so it precedes the dependency wait
    for (int i = 0; i < PF_DIST_OUT; i++) {
        const uint4 rp = s_rows[i];
        prefetch_l2_either(rp.w & 4, codes_s + rp.y, P.X + 2 * (size_t)(rp.y + s_raw));
        prefetch_l2_either(rp.w & 8, codes_s + rp.z, P.X + 2 * (size_t)(rp.z + s_raw));
        if (rp.w & 1) { prefetch_l2(P.O + (rp.x + s_raw)); prefetch_l2(P.OI + (rp.x + s_raw)); }
    }
    grid_dependency_wait();
    const double invL = valid ? P.invL[(size_t)cl.z * P.K + s] : 0.0;
    fetch(s_rows[0], c0v, c0w, A);

    auto step = [&](int i, OutSlot &cur, OutSlot &nxt) {
        // ---- prologue: look-ahead past the end of the clade re-reads the last step's operands (padding), no guards
        int c2v = 0, c2w = 0;
        codes_of(s_rows[i + 2], c2v, c2w);
        fetch(s_rows[i + 1], c1v, c1w, nxt);
        if (PF_DIST_OUT > 0) {
            // operands written by earlier launches (every inside record, leaf codes) are pulled into L2 a few steps ahead
            const uint4 rp = s_rows[i + PF_DIST_OUT];
            prefetch_l2_either(rp.w & 4, codes_s + rp.y, P.X + 2 * (size_t)(rp.y + s_raw));
            prefetch_l2_either(rp.w & 8, codes_s + rp.z, P.X + 2 * (size_t)(rp.z + s_raw));
            // (the out-record of a step that LOADS it is not prefetched: ncu shows 10 % of the stall samples on its first
            // use, but two more accesses per step on the L1 data pipe cost more than they save: 2.78 -> 2.85 ms)
        }
        const int4 d0 = s_steps[i];                      // node ids of this step: count rows, posteriors
        const uint4 r0 = s_rows[i];
        const int f = d0.w;
        const double4 bcv = s_bc[2 * i], bcw = s_bc[2 * i + 1];
        // one straight-line body per step shape (which children are leaves): big basic blocks for the scheduler,
        // and everything a leaf does not need (out-record, normalisation, store) is compiled out
        auto core = [&](auto VL, auto WL) {
        constexpr bool vleaf = decltype(VL)::value, wleaf = decltype(WL)::value;
        Child v, w;
        if (vleaf) decode_leaf(v, c0v, cur.qv, bcv, qsum, s_lut); else decode(v, cur.v, bcv, qsum);
        if (wleaf) decode_leaf(w, c0w, cur.qw, bcw, qsum, s_lut); else decode(w, cur.w, bcw, qsum);
        if (f & 1) chain = cur.o;                       // how = 1: loaded; otherwise out[u] is what the previous step left in the registers
        // everything is relative to the site likelihood: fac = 2^(F_u + cum_v + cum_w) / L
        const int F = chain.ri.y;
        const double fac = invL * pow2i(max(-1000, min(1000, F + v.cum + w.cum)));
        double c5[5];
        Opd ov, ow;
        child_outside<!wleaf>(chain, w, v, bcw, qsum, fac, ow, c5);
        red_put5(c5, red, lane, 10 * nbuf + 5);          // counts leave the registers at once
        child_outside<!vleaf>(chain, v, w, bcv, qsum, fac, ov, c5);
        red_put5(c5, red, lane, 10 * nbuf);
        // phi statistic: 1 - P(leaf silenced | D) on '?' entries (EM_solver.py:323)
        if (vleaf && valid && c0v == -1) rt += 1.0 - ov.r.y * fac;
        if (wleaf && valid && c0w == -1) rt += 1.0 - ow.r.y * fac;
        if (POST && valid) {
            // post0 = log P(node in state 0 | D), post1 likewise for the silenced state (EM_solver.py:276-277)
            const size_t gv = (size_t)d0.y * P.K_total + P.site0 + s, gw = (size_t)d0.z * P.K_total + P.site0 + s;
            const double p0v = ov.r.x * v.x0 * fac, p1v = ov.r.y * v.xm * fac;
            const double p0w = ow.r.x * w.x0 * fac, p1w = ow.r.y * w.xm * fac;
            P.post0[gv] = (p0v > 0.0) ? log(p0v) : -INFINITY; P.post1[gv] = (p1v > 0.0) ? log(p1v) : -INFINITY;
            P.post0[gw] = (p0w > 0.0) ? log(p0w) : -INFINITY; P.post1[gw] = (p1w > 0.0) ? log(p1w) : -INFINITY;
        }
        // F_child = F_u + cum_sibling + (exponent taken out of the child's record)
        if (!wleaf) {
            normalise_out(ow, F + v.cum);
            st_orec(P.O, P.OI, r0.z + s_raw, ow);        // an internal w is never the chained child: always stored
        }
        if (!vleaf) {
            normalise_out(ov, F + w.cum);
            if ((f & 16) || P.store_all) st_orec(P.O, P.OI, r0.y + s_raw, ov);
            chain = ov;
        }
        };
        switch ((f >> 2) & 3) {
            case 0: core(std::false_type{}, std::false_type{}); break;
            case 1: core(std::true_type{}, std::false_type{}); break;
            case 2: core(std::false_type{}, std::true_type{}); break;
            default: core(std::true_type{}, std::true_type{}); break;
        }
        // counts: summed over the warp's sites once RED_STEPS steps are buffered
        if (++nbuf == RED_STEPS) { red_flush(red, lane, RED_STEPS, i, s_steps, Sg); nbuf = 0; }
        c0v = c1v; c0w = c1w; c1v = c2v; c1w = c2w;
    };
    int i = 0;
    for (; i + 1 < nsteps; i += 2) { step(i, A, B); step(i + 1, B, A); }      // two register sets swap roles: no moves
    if (i < nsteps) step(i, A, B);
    if (nbuf) red_flush(red, lane, nbuf, nsteps - 1, s_steps, Sg);
    grid_launch_dependents();
    // R~ of this (warp, clade)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rt += __shfl_xor_sync(0xffffffffu, rt, o);
    if (lane == 0) P.Rpart[(size_t)grp * P.n_clades + cl.w] = rt;
}

template <bool POST, bool DF = false>
__global__ void __launch_bounds__(LT, PMM_LOUT_MINB) pmm_lumped_outside(LParams P, const int4 *__restrict__ clades, const int4 *__restrict__ steps) {
    __shared__ int4 s_steps[STEP_CAP + STEP_PAD];
    __shared__ uint4 s_rows[STEP_CAP + STEP_PAD];
    __shared__ double4 s_bc[2 * STEP_CAP];
    __shared__ double4 s_lut[4];
    __shared__ __align__(16) double s_red[(LT / 32) * RED_ROWS * RED_ROW];
    int ci, blk;
    df_item<DF>(P, ci, blk);
    const int4 cl = clades[ci];
    stage_clade(s_steps, s_bc, s_lut, P, steps, cl.x, cl.y - cl.x);
    stage_rows(s_rows, s_steps, P, cl.y - cl.x);
    if (DF) grid_dependency_wait();                     // (otherwise inside the walk, after its look-ahead on the inputs)
    df_wait<DF>(P, P.df_lo + ci, blk + P.sblk0);
    outside_walk<POST>(P, cl, blk, s_steps, s_rows, s_bc, s_lut, s_red);
    df_done<DF>(P, P.df_lo + ci, blk + P.sblk0);
}

}  // namespace pmm
