// TEST INFRASTRUCTURE: the device ultrametric M-step kernel (laml_b200/csrc/pmm_ultrametric_dev.cuh), unchanged, run
// on the host under tests/native/cuda_host_shim.h as ONE CTA (its grid barrier degenerates for a one-CTA grid, the
// level sweeps, reductions, line search and barrier schedule are exactly the code the GPU runs).
#include "cuda_host_shim.h"
#include <atomic>
template <class T> inline T __ldcg(const T *p) { return *p; }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
#include "../../laml_b200/csrc/pmm_ultrametric_dev.cuh"
using namespace pmm;

extern "C" int emu_um_solve(int n, int n_roots, const int *parent, const int *left, const int *roots, const int *up_items, const int *dn_items,
                            const int *up_ptr, const int *dn_ptr, const int *up_ph, int n_up_ph, const int *dn_ph, int n_dn_ph,
                            const signed char *kind, const double *fixed_len, const double *stats, double *Snorm, const double *params,
                            double lo, double hi, double eps_s, double eps_nu, int K_total, int warm, double *H_prev, double *d_out,
                            int *result, int threads) {
    const size_t n1 = (size_t)n + 1;
    std::vector<double> sc(11 * n1, 0.0), partial(2 * 1024 * 4, 0.0);
    std::vector<unsigned char> grounded(n1, 0);
    unsigned bar = 0;
    UmArgs A;
    A.n = n; A.n_roots = n_roots; A.n_int = 0;
    A.parent = parent; A.left = left; A.roots = roots;
    A.up_items = (const int4 *)up_items; A.dn_items = (const int4 *)dn_items; A.up_ptr = up_ptr; A.dn_ptr = dn_ptr;
    A.up_ph = (const UmPhase *)up_ph; A.dn_ph = (const UmPhase *)dn_ph; A.n_up_ph = n_up_ph; A.n_dn_ph = n_dn_ph;
    A.kind = kind; A.fixed_len = fixed_len; A.stats = stats; A.Snorm = Snorm; A.params = params;
    A.lo = lo; A.hi = hi; A.eps_s = eps_s; A.eps_nu = eps_nu; A.mu_final = 1e-11; A.warm_theta = 1e-6; A.warm_mu = 1e-9;
    A.K_total = K_total; A.warm = warm; A.H_prev = H_prev;
    double *s = sc.data();
    A.H = s; A.Hint = s + n1; A.g = s + 2 * n1; A.a = s + 3 * n1; A.b = s + 4 * n1; A.alpha = s + 5 * n1; A.beta = s + 6 * n1;
    A.delta = s + 7 * n1; A.k = s + 8 * n1; A.S1 = s + 9 * n1; A.S24 = s + 10 * n1;
    A.grounded = grounded.data(); A.partial = partial.data(); A.bar = &bar; A.d_out = d_out; A.result = result;
    if (threads != UM_THREADS) return -1;
    emu::launch(1u, (unsigned)UM_THREADS, pmm_um_solve, A);
    return 0;
}
