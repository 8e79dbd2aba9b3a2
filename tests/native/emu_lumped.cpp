// emu_lumped.cpp -- TEST INFRASTRUCTURE: pmm_nni_lumped<ULTRA> (laml_b200/csrc/pmm_nni_lumped.cuh), the kernel
// behind pmm_score_nni_batch / pmm_score_nni_batch_ultra, compiled for the host under cuda_host_shim.h and fed with
// records built by the numpy model of the lumped engine (tests/lumped_model.py).  Built and driven by
// tests/test_kernels_emulated.py.
#include "cuda_host_shim.h"
#include "../../laml_b200/csrc/pmm_nni_lumped.cuh"

using namespace pmm;

extern "C" int emu_nni_lumped(int ultra, int n_cand, const int *cand8 /* [n_cand][8] = g,s,A,B,C,v,u,root */, const double *d5_in,
                              const double *ultra5 /* [n_cand][5] = hA,hB,hC,hTop,top_free or null */,
                              int K, int Kp, int n_leaves, int n_nodes, const int *codes /* [n_leaves][K] */, const int *site_ptr,
                              const double *prior, const double *qsum, double nu, double phi, const double *bc /* [n][4] */,
                              const double *X /* [n_int][Kp][4] */, const double *O /* [n_int][Kp][4] */, const int *OI /* [n_int][Kp][2] */,
                              const double *invL /* [n_roots][K] */, const double *root_llh_sum, double llh_total, int K_total,
                              double dmin, double dmax, double conv_eps, int max_iter,
                              double *out_llh, double *out_d5, int *out_iters, int *out_done) {
    LParams P;
    memset(&P, 0, sizeof(P));
    P.K = K; P.Kp = Kp; P.site0 = 0; P.code_stride = K; P.n_leaves = n_leaves; P.n_nodes = n_nodes; P.n_clades = 0;
    P.codes = codes; P.site_ptr = site_ptr; P.prior = prior; P.qsum = qsum;
    const double params[2] = {nu, phi};
    P.params = params; P.bc = bc;
    P.X = reinterpret_cast<double2 *>(const_cast<double *>(X));
    P.O = reinterpret_cast<double4 *>(const_cast<double *>(O));
    P.OI = reinterpret_cast<int2 *>(const_cast<int *>(OI));
    P.invL = const_cast<double *>(invL);
    P.K_total = K_total;
    std::vector<NniCand> cands(n_cand);
    std::vector<NniState> st(n_cand);
    std::vector<NniUltra> ul(n_cand);
    for (int i = 0; i < n_cand; i++) {
        const int *c = cand8 + 8 * i;
        cands[i] = NniCand{c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7]};
        for (int b = 0; b < 5; b++) st[i].d[b] = d5_in[5 * i + b];
        st[i].llh_prev = st[i].llh_cur = 0.0; st[i].iters = 0; st[i].done = 0;
        if (ultra) { const double *u = ultra5 + 5 * i; ul[i] = NniUltra{u[0], u[1], u[2], u[3], (int)u[4], 0}; }
    }
    if (ultra)
        emu::launch((unsigned)n_cand, NNIL_THREADS, pmm_nni_lumped<true>, P, cands.data(), st.data(), root_llh_sum, llh_total, K_total,
                    dmin, dmax, 1e-6, 1e-5, conv_eps, max_iter, ul.data());
    else
        emu::launch((unsigned)n_cand, NNIL_THREADS, pmm_nni_lumped<false>, P, cands.data(), st.data(), root_llh_sum, llh_total, K_total,
                    dmin, dmax, 1e-6, 1e-5, conv_eps, max_iter, (const NniUltra *)nullptr);
    for (int i = 0; i < n_cand; i++) {
        out_llh[i] = st[i].llh_cur; out_iters[i] = st[i].iters; out_done[i] = st[i].done;
        for (int b = 0; b < 5; b++) out_d5[5 * i + b] = st[i].d[b];
    }
    return 0;
}

// The E-step kernels themselves: pmm_lumped_inside, then pmm_lumped_outside<false>, phase by phase as
// pmm_engine.cu: run_lumped launches them (one launch per tier of the clade schedule, grid = clades x site blocks; or the
// dataflow launches over several tiers, df_mode 1 / 2).
extern "C" int emu_lumped_estep(int K, int Kp, int n_leaves, int n_nodes, int n_clades, const int *codes, const int *site_ptr,
                                const double *prior, const double *qsum, double nu, double phi, const double *d /* [n] */,
                                const int *in_steps, const int *in_clades, const int *in_phase_ptr, int n_in_phases,
                                const int *out_steps, const int *out_clades, const int *out_phase_ptr, int n_out_phases, int store_all,
                                double *X, double *O, int *OI, double *invL, double *root_site_llh, double *Spart, double *Rpart,
                                double *bc_out /* [n][4] */, double *post0 /* [n][K] or null: the annotation variant */, double *post1,
                                int df_mode /* 0: a launch per tier; 1: tier 0 | top tiers as one dataflow launch; 2: one dataflow launch per pass */,
                                const int *in_dep_ptr, const int *in_deps, const int *out_dep_ptr, const int *out_deps) {
    LParams P;
    memset(&P, 0, sizeof(P));
    P.K = K; P.Kp = Kp; P.site0 = 0; P.code_stride = K; P.n_leaves = n_leaves; P.n_nodes = n_nodes; P.n_clades = n_clades;
    P.codes = codes; P.site_ptr = site_ptr; P.prior = prior; P.qsum = qsum;
    const double params[2] = {nu, phi};
    P.params = params;
    for (int v = 0; v < n_nodes; v++) {                       // pmm_branch_consts
        const double4 c = consts_of(d[v], nu);
        bc_out[4 * v] = c.x; bc_out[4 * v + 1] = c.y; bc_out[4 * v + 2] = c.z; bc_out[4 * v + 3] = c.w;
    }
    P.bc = bc_out;
    P.X = reinterpret_cast<double2 *>(X); P.O = reinterpret_cast<double4 *>(O); P.OI = reinterpret_cast<int2 *>(OI);
    P.invL = invL; P.root_site_llh = root_site_llh; P.Spart = Spart; P.Rpart = Rpart;
    P.K_total = K; P.store_all = store_all; P.sblk0 = 0;
    P.post0 = post0; P.post1 = post1;
    const unsigned gy = (unsigned)((K + LT - 1) / LT);
    if (df_mode) {
        // as pmm_engine.cu: run_lumped groups the tiers; the CTAs of a launch run one after another here, in ticket order
        const int ncl = in_phase_ptr[n_in_phases], nt0 = in_phase_ptr[1];
        std::vector<int> flags((size_t)2 * ncl * gy, 0);
        unsigned ticket = 0;
        int err = 0, epoch = 0;
        struct G { int c0, c1, n1, n2, df; };
        std::vector<G> gin, gout;
        if (df_mode == 1) { gin = {{0, nt0, 0, 0, 0}, {nt0, ncl, 0, ncl - nt0, 1}}; gout = {{0, ncl - nt0, ncl - nt0, 0, 1}, {ncl - nt0, ncl, 0, 0, 0}}; }
        else { gin = {{0, ncl, nt0, ncl - nt0, 1}}; gout = {{0, ncl, ncl - nt0, nt0, 1}}; }
        auto run = [&](const G &g, int pass) {
            if (g.c1 <= g.c0) return;
            LParams Q = P;
            const int4 *cl = reinterpret_cast<const int4 *>(pass ? out_clades : in_clades) + g.c0;
            const int4 *st = reinterpret_cast<const int4 *>(pass ? out_steps : in_steps);
            if (g.df) {
                Q.df_ticket = &ticket; Q.df_base = ticket; Q.df_flags = flags.data() + (size_t)pass * ncl * gy;
                Q.df_dep_ptr = pass ? out_dep_ptr : in_dep_ptr; Q.df_deps = pass ? out_deps : in_deps;
                Q.df_lo = g.c0; Q.df_n1 = g.n1; Q.df_n2 = g.n2; Q.df_nblk = (int)gy; Q.df_stride = (int)gy; Q.df_epoch = ++epoch; Q.df_err = &err;
                const unsigned total = (unsigned)(g.c1 - g.c0) * gy;
                if (!pass) emu::launch(total, LT, pmm_lumped_inside<true>, Q, cl, st);
                else if (post0) emu::launch(total, LT, pmm_lumped_outside<true, true>, Q, cl, st);
                else emu::launch(total, LT, pmm_lumped_outside<false, true>, Q, cl, st);
            } else {
                if (!pass) emu::launch2((unsigned)(g.c1 - g.c0), gy, LT, pmm_lumped_inside<false>, Q, cl, st);
                else if (post0) emu::launch2((unsigned)(g.c1 - g.c0), gy, LT, pmm_lumped_outside<true, false>, Q, cl, st);
                else emu::launch2((unsigned)(g.c1 - g.c0), gy, LT, pmm_lumped_outside<false, false>, Q, cl, st);
            }
        };
        for (const G &g : gin) run(g, 0);
        for (const G &g : gout) run(g, 1);
        return err ? 2 : 0;              // 2: a CTA found a producer's mark missing -- the item order is wrong
    }
    for (int p = 0; p < n_in_phases; p++) {
        const int c0 = in_phase_ptr[p], c1 = in_phase_ptr[p + 1];
        if (c1 > c0) emu::launch2((unsigned)(c1 - c0), gy, LT, pmm_lumped_inside<false>, P, reinterpret_cast<const int4 *>(in_clades) + c0,
                                  reinterpret_cast<const int4 *>(in_steps));
    }
    for (int p = 0; p < n_out_phases; p++) {
        const int c0 = out_phase_ptr[p], c1 = out_phase_ptr[p + 1];
        if (c1 <= c0) continue;
        if (post0) emu::launch2((unsigned)(c1 - c0), gy, LT, pmm_lumped_outside<true, false>, P, reinterpret_cast<const int4 *>(out_clades) + c0,
                                reinterpret_cast<const int4 *>(out_steps));
        else emu::launch2((unsigned)(c1 - c0), gy, LT, pmm_lumped_outside<false, false>, P, reinterpret_cast<const int4 *>(out_clades) + c0,
                          reinterpret_cast<const int4 *>(out_steps));
    }
    return 0;
}
