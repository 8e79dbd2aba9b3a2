/*
 * pmm_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, fp64, CPU restatement of the reference's PMM likelihood / E-step /
 * M-step arithmetic (raphael-group/LAML v1.0.5, pure Python).  It exists so the
 * CUDA engine in laml_b200/csrc can be checked against the reference's own
 * algorithm on machines where the Python reference cannot travel.  It follows
 * the reference's *compressed z/alpha/?* formulation in log space -- i.e. it is
 * deliberately NOT the full-state linear-space recursion the GPU engine uses,
 * so agreement between the two is a real check and not a tautology.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product path never does.
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against
 *   (1) every known-answer likelihood in laml_unit_tests/unit_tests_EMSolver.py
 *       (tests 1-20, 46) and the example1 / case_tests stored optima, and
 *   (2) per-(node,site) L0/L1/out0/out1/post0/post1/S0..S4 dumped from the
 *       reference's own EM_solver imported verbatim (tests/golden/make_golden.py).
 *
 * Reference citations are to /root/reference/laml_libs/<file>:<line>.
 *
 * Data model (flattened by the caller):
 *   nodes 0..n-1, binary, parent[]/left[]/right[] (-1 = none), leaves are the
 *   nodes with left==-1; order[] is any children-before-parent sequence
 *   (the reference uses treeswift postorder, EM_solver.py:97); several roots
 *   (a forest, ML_solver.py:26-32) are allowed.
 *   codes[leaf_row*K + s]: 0 = unedited, a>=1 = edited state with 1-based rank a
 *   in the site's alphabet, -1 = '?' (missing), -2 = literal -1 (known silenced,
 *   ML_solver.py:139-140).  leaf_row[v] = row of codes for leaf v.
 *   qptr[K+1], qval[]: qval[qptr[s]+a-1] = normalised prior of state a at site s
 *   (ML_solver.py:70-74).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NEG_INF (-INFINITY)
#define TAG_Z (-1)   /* 'z' */
#define TAG_Q (-2)   /* '?' */

typedef struct {
    int n, K, nleaf_rows;
    const int32_t *parent, *left, *right, *order, *leaf_row;
    const int32_t *codes;
    const int32_t *qptr;
    const double *qval;
} oracle_problem;

/* EM_solver.py:8-13 log_sum_exp.  The reference returns NaN when every entry
 * is -inf (inf-inf); this restatement defines that case as -inf. */
static double lse(const double *v, int k) {
    double m = NEG_INF;
    for (int i = 0; i < k; i++) if (v[i] > m) m = v[i];
    if (k == 0 || m == NEG_INF) return NEG_INF;
    double s = 0.0;
    for (int i = 0; i < k; i++) s += exp(v[i] - m);
    return m + log(s);
}
/* EM_solver.py:15-16 */
static double pseudo_log(double x) { return x > 0 ? log(x) : NEG_INF; }

static inline double qof(const oracle_problem *P, int s, int a) { return P->qval[P->qptr[s] + a - 1]; }

/* ---- az_partition, ML_solver.py:125-154 (one site) ------------------------ */
static void az_site(const oracle_problem *P, int s, int32_t *tag) {
    for (int i = 0; i < P->n; i++) {
        int v = P->order[i];
        if (P->left[v] < 0) {
            int c = P->codes[(size_t)P->leaf_row[v] * P->K + s];
            if (c == 0) tag[v] = TAG_Z;                 /* :137-138 */
            else if (c == -2) tag[v] = TAG_Q;           /* :139-140 literal -1 */
            else if (c == -1) tag[v] = TAG_Q;           /* '?' is its own tag, :141-142 + :148 */
            else tag[v] = c;
        } else {
            int t1 = tag[P->left[v]], t2 = tag[P->right[v]];
            /* :147-154: S = set of child tags, R = S - {'z','?',-1} */
            if (t1 == TAG_Z || t2 == TAG_Z) tag[v] = TAG_Z;
            else if (t1 > 0 && t2 > 0 && t1 != t2) tag[v] = TAG_Z;
            else if (t1 > 0) tag[v] = t1;
            else if (t2 > 0) tag[v] = t2;
            else tag[v] = TAG_Q;
        }
    }
}

/* ---- Estep_in_llh, EM_solver.py:91-133 (one site) -------------------------- */
static void inside_site(const oracle_problem *P, int s, const int32_t *tag, const double *d,
                        double nu, double phi, double *L0, double *L1) {
    for (int i = 0; i < P->n; i++) {
        int v = P->order[i];
        double dv = d[v];
        double p = exp(-dv);                                           /* :98 */
        int t = tag[v];
        double q = (t > 0) ? qof(P, s, t) : 1.0;                        /* :102 */
        if (P->left[v] < 0) {
            int c = P->codes[(size_t)P->leaf_row[v] * P->K + s];
            if (t == TAG_Q) {                                           /* :104-110 */
                double m;
                if (c == -1) { double x = 1 - (1 - phi) * pow(p, nu); m = x > 0 ? log(x) : NEG_INF; }
                else { double x = 1 - pow(p, nu); m = x > 0 ? log(x) : NEG_INF; }
                L0[v] = L1[v] = m;
            } else if (t == TAG_Z) {                                    /* :111-113 */
                L0[v] = (1 - phi > 0) ? (nu + 1) * (-dv) + log(1 - phi) : NEG_INF;
                L1[v] = NEG_INF;
            } else {                                                    /* :114-116 */
                L0[v] = ((1 - p) * q * (1 - phi) > 0) ? nu * (-dv) + log(1 - p) + log(q) + log(1 - phi) : NEG_INF;
                L1[v] = (1 - phi > 0) ? nu * (-dv) + log(1 - phi) : NEG_INF;
            }
        } else {                                                        /* :117-133 */
            int c1 = P->left[v], c2 = P->right[v];
            double l0 = 0, l1 = 0;
            l0 += L0[c1]; l1 += L1[c1];
            l0 += L0[c2]; l1 += L1[c2];
            double lst[3]; int k = 0;
            lst[k++] = l0 + (nu + 1) * (-dv);
            if (t != TAG_Z && q * (1 - p) > 0) lst[k++] = l1 + log(1 - p) + log(q) + nu * (-dv);
            if (t == TAG_Q && (1 - pow(p, nu)) > 0) lst[k++] = log(1 - pow(p, nu));
            L0[v] = lse(lst, k);
            if (t == TAG_Z) L1[v] = NEG_INF;
            else if (t != TAG_Q || nu == 0 || p == 1) L1[v] = l1 + nu * (-dv);
            else { double b[2] = { l1 + nu * (-dv), log(1 - pow(p, nu)) }; L1[v] = lse(b, 2); }
        }
    }
}

/* ---- out_alpha dictionaries: per (node) a short list of (alpha -> value) ---- */
typedef struct { int alpha; int next; double val; } oa_entry;
typedef struct { oa_entry *e; int used, cap; int32_t *head; } oa_pool;

static double *oa_find(oa_pool *D, int v, int alpha) {
    for (int i = D->head[v]; i >= 0; i = D->e[i].next) if (D->e[i].alpha == alpha) return &D->e[i].val;
    return NULL;
}
static void oa_set(oa_pool *D, int v, int alpha, double val) {
    double *p = oa_find(D, v, alpha);
    if (p) { *p = val; return; }
    if (D->used == D->cap) { D->cap = D->cap ? 2 * D->cap : 1024; D->e = (oa_entry *)realloc(D->e, sizeof(oa_entry) * D->cap); }
    D->e[D->used].alpha = alpha; D->e[D->used].val = val; D->e[D->used].next = D->head[v];
    D->head[v] = D->used++;
}

static inline int sister(const oracle_problem *P, int v) {
    int u = P->parent[v];
    return P->left[u] == v ? P->right[u] : P->left[u];
}

/* ---- __out_alpha_up__, EM_solver.py:206-245 -------------------------------- */
static void out_alpha_up(const oracle_problem *P, int s, const int32_t *tag, const double *d, double nu,
                         const double *L1, const double *A, oa_pool *D, int node, int alpha0,
                         int32_t *path) {
    int v = node, np = 0;
    double qa = qof(P, s, alpha0);
    while (P->parent[v] >= 0) {                                         /* :211 */
        int u = P->parent[v], w = sister(P, v);
        if (tag[w] != TAG_Q && tag[w] != alpha0) {                      /* :217 */
            oa_set(D, v, alpha0, A[v] - nu * d[v] + pseudo_log(1 - exp(-d[v])) + pseudo_log(qa)); /* :222 */
            break;
        }
        path[np++] = v;
        v = u;
    }
    if (P->parent[v] < 0)                                               /* :226-231 */
        oa_set(D, v, alpha0, -nu * d[v] + pseudo_log(1 - exp(-d[v])) + pseudo_log(qa));
    while (np > 0) {                                                    /* :233-245 */
        v = path[--np];
        int u = P->parent[v], w = sister(P, v);
        double B = *oa_find(D, u, alpha0) + nu * (-d[v]) + L1[w];
        double C = A[v] - nu * d[v] + pseudo_log(1 - exp(-d[v])) + pseudo_log(qa);
        double bc[2] = { B, C };
        oa_set(D, v, alpha0, lse(bc, 2));
    }
}

/* ---- Estep_out_llh, EM_solver.py:143-204 (one site) ------------------------ */
static void outside_site(const oracle_problem *P, int s, const int32_t *tag, const double *d, double nu,
                         const double *L0, const double *L1, double *out0, double *out1,
                         double *A, double *X, oa_pool *D, int32_t *path) {
    D->used = 0;
    for (int i = 0; i < P->n; i++) D->head[i] = -1;
    for (int i = P->n - 1; i >= 0; i--) {                               /* parents before children */
        int v = P->order[i];
        double dv = d[v];
        if (P->parent[v] < 0) {                                         /* :151-158 */
            A[v] = 0;
            X[v] = (nu * dv > 0) ? -nu * dv + log(1 - exp(-dv)) : NEG_INF;
            out0[v] = -(1 + nu) * dv;
            out1[v] = (nu * dv > 0) ? log(1 - exp(-dv * nu)) : NEG_INF;
            continue;
        }
        int u = P->parent[v], w = sister(P, v);
        out0[v] = out0[u] + L0[w] - (1 + nu) * dv;                      /* :179 */
        A[v] = out0[u] + L0[w];                                         /* :180 */
        X[v] = (dv > 0) ? -nu * dv + log(1 - exp(-dv)) + A[v] : NEG_INF; /* :181 */
        if (tag[w] == TAG_Z) {                                          /* :183-184 */
            out1[v] = (nu * dv > 0) ? log(1 - exp(-dv * nu)) + A[v] : NEG_INF;
        } else if (tag[w] == TAG_Q) {                                   /* :185-189 */
            double x2[2] = { X[v], X[u] + L1[w] - nu * dv };
            X[v] = lse(x2, 2);
            double p = 1 - exp(-dv * nu);
            double pl = p > 0 ? log(p) : NEG_INF;
            if (nu == 0) out1[v] = out1[u];
            else { double t3[3] = { pl + A[v], pl + X[u] + L1[w], out1[u] }; out1[v] = lse(t3, 3); }
        } else {                                                        /* :190-204 */
            int alpha0 = tag[w];
            if (!oa_find(D, u, alpha0)) out_alpha_up(P, s, tag, d, nu, L1, A, D, u, alpha0, path);
            double ua = *oa_find(D, u, alpha0);
            double qa = qof(P, s, alpha0);
            double B = ua + nu * (-dv) + L1[w];
            double C = (dv > 0 && qa > 0) ? A[v] - nu * dv + log(1 - exp(-dv)) + log(qa) : NEG_INF;
            double bc[2] = { B, C };
            oa_set(D, v, alpha0, lse(bc, 2));
            double xb[2] = { X[v], B };
            X[v] = lse(xb, 2);
            if (nu * dv > 0) { double t2[2] = { A[v], L1[w] + ua }; out1[v] = log(1 - exp(-dv * nu)) + lse(t2, 2); }
            else out1[v] = NEG_INF;
        }
    }
}

/* ---- Estep_posterior, EM_solver.py:247-300 (one site) ---------------------- */
static void posterior_site(const oracle_problem *P, int s, const int32_t *tag, const double *d,
                           double nu, double phi, const double *L0, const double *L1,
                           const double *out0, const double *out1, const int32_t *root_of,
                           double *post0, double *post1, double *S /* 5 arrays of n, S[k*n+v] */) {
    int n = P->n;
    for (int i = n - 1; i >= 0; i--) {
        int v = P->order[i];
        double dv = d[v];
        double full = L0[root_of[v]];                                   /* :254 */
        int t = tag[v];
        int has_in1 = (t == TAG_Q);                                     /* :265 v_in1 = 0 or None */
        double v_in0;
        if (P->left[v] < 0) {                                           /* :266-271 */
            int c = P->codes[(size_t)P->leaf_row[v] * P->K + s];
            if (c == 0) v_in0 = (1 - phi > 0) ? log(1 - phi) : NEG_INF;
            else v_in0 = (c == -1 && phi > 0) ? log(phi) : NEG_INF;
        } else v_in0 = L0[P->left[v]] + L0[P->right[v]];                /* :273-274 */
        post0[v] = v_in0 + out0[v] - full;                              /* :276 */
        post1[v] = has_in1 ? 0 + out1[v] - full : NEG_INF;              /* :277 */
        double *S0 = S, *S1 = S + n, *S2 = S + 2 * n, *S3 = S + 3 * n, *S4 = S + 4 * n;
        if (t == TAG_Z) {                                               /* :279-281 */
            S0[v] = 1.0; S1[v] = S2[v] = S3[v] = S4[v] = 0.0;
        } else if (P->parent[v] < 0) {                                  /* :282-286 */
            S0[v] = exp(v_in0 + (1.0 + nu) * (-dv) - L0[v]);
            S2[v] = (t != TAG_Q) ? 0.0 : (1.0 - exp(-nu * dv)) / exp(L0[v]);
            S1[v] = 1.0 - S0[v] - S2[v];
            S3[v] = S4[v] = 0.0;
        } else {                                                        /* :287-300 */
            int u = P->parent[v];
            S0[v] = exp(post0[u] + v_in0 + (1.0 + nu) * (-dv) - L0[v]);
            if (t != TAG_Q) { S2[v] = 0.0; S4[v] = 0.0; }
            else {
                S2[v] = exp(post0[u] - L0[v]) * (1.0 - exp(-nu * dv));
                double r = 1.0 - exp(post0[u]) - exp(post1[u]);
                double g = 1.0 - exp(-nu * dv);
                S4[v] = (r == 0 || g == 0) ? 0 : r * g / exp(L1[v]);
            }
            S1[v] = exp(post0[u]) - S0[v] - S2[v];
            S3[v] = 1.0 - S0[v] - S1[v] - exp(post1[v]);
        }
    }
}

/* =============================== public API ================================= */

typedef struct {
    int32_t *tag; double *L0, *L1, *out0, *out1, *A, *X, *post0, *post1, *S;
    oa_pool D; int32_t *path;
} scratch;

static void scratch_init(scratch *w, int n) {
    w->tag = (int32_t *)malloc(sizeof(int32_t) * n);
    w->L0 = (double *)malloc(sizeof(double) * n * 13);
    w->L1 = w->L0 + n; w->out0 = w->L1 + n; w->out1 = w->out0 + n; w->A = w->out1 + n; w->X = w->A + n;
    w->post0 = w->X + n; w->post1 = w->post0 + n; w->S = w->post1 + n;
    w->D.e = NULL; w->D.used = w->D.cap = 0; w->D.head = (int32_t *)malloc(sizeof(int32_t) * n);
    w->path = (int32_t *)malloc(sizeof(int32_t) * n);
}
static void scratch_free(scratch *w) { free(w->tag); free(w->L0); free(w->D.e); free(w->D.head); free(w->path); }

static void fill_root_of(const oracle_problem *P, int32_t *root_of) {
    for (int i = P->n - 1; i >= 0; i--) {
        int v = P->order[i];
        root_of[v] = P->parent[v] < 0 ? v : root_of[P->parent[v]];
    }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* az_partition for all sites: tag_out[v*K+s]  (ML_solver.py:125-154) */
void oracle_az_partition(int n, int K, const int32_t *parent, const int32_t *left, const int32_t *right,
                         const int32_t *order, const int32_t *leaf_row, const int32_t *codes,
                         int32_t *tag_out) {
    oracle_problem P = { n, K, 0, parent, left, right, order, leaf_row, codes, NULL, NULL };
    int32_t *tag = (int32_t *)malloc(sizeof(int32_t) * n);
    for (int s = 0; s < K; s++) {
        az_site(&P, s, tag);
        for (int v = 0; v < n; v++) tag_out[(size_t)v * K + s] = tag[v];
    }
    free(tag);
}

/*
 * lineage_llh (EM_solver.py:135-141) over sites [s_lo, s_hi): returns sum over
 * trees and sites of root.L0; optional per-(node,site) dumps (NULL to skip),
 * all laid out [v*K + s].
 */
double oracle_inside(int n, int K, const int32_t *parent, const int32_t *left, const int32_t *right,
                     const int32_t *order, const int32_t *leaf_row, const int32_t *codes,
                     const int32_t *qptr, const double *qval, const double *d, double nu, double phi,
                     int s_lo, int s_hi, int nthreads, double *site_llh, double *L0_out, double *L1_out) {
    oracle_problem P = { n, K, 0, parent, left, right, order, leaf_row, codes, qptr, qval };
    double total = 0.0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads) reduction(+:total)
    {
        scratch w; scratch_init(&w, n);
#pragma omp for schedule(dynamic, 1)
        for (int s = s_lo; s < s_hi; s++) {
            az_site(&P, s, w.tag);
            inside_site(&P, s, w.tag, d, nu, phi, w.L0, w.L1);
            double l = 0;
            for (int v = 0; v < n; v++) if (parent[v] < 0) l += w.L0[v];
            if (site_llh) site_llh[s] = l;
            total += l;
            if (L0_out) for (int v = 0; v < n; v++) { L0_out[(size_t)v * K + s] = w.L0[v]; L1_out[(size_t)v * K + s] = w.L1[v]; }
        }
        scratch_free(&w);
    }
    if (site_llh) { total = 0; for (int s = s_lo; s < s_hi; s++) total += site_llh[s]; }
    return total;
}

/*
 * Estep (EM_solver.py:302-305) over sites [s_lo, s_hi).
 * Ssum[k*n+v] += sum over sites of S_k[v][site]   (what Mstep sums at :341)
 * Rtilde      += sum over leaves, '?' entries of 1-exp(post1)   (:323)
 * returns the log-likelihood of those sites.
 * dump[] (optional, may be NULL): 13 arrays [q*n*K + v*K + s] in the order
 *   L0,L1,out0,out1,A,X,post0,post1,S0,S1,S2,S3,S4.
 */
double oracle_estep(int n, int K, const int32_t *parent, const int32_t *left, const int32_t *right,
                    const int32_t *order, const int32_t *leaf_row, const int32_t *codes,
                    const int32_t *qptr, const double *qval, const double *d, double nu, double phi,
                    int s_lo, int s_hi, int nthreads, double *Ssum, double *Rtilde, double *site_llh,
                    double *dump) {
    oracle_problem P = { n, K, 0, parent, left, right, order, leaf_row, codes, qptr, qval };
    int32_t *root_of = (int32_t *)malloc(sizeof(int32_t) * n);
    fill_root_of(&P, root_of);
    if (nthreads < 1) nthreads = 1;
    double total = 0.0, rt = 0.0;
    /* per-site partial sums are accumulated in site order per thread and then
       combined in thread order; callers needing bit-reproducibility use nthreads=1 */
    double *Sacc = (double *)calloc((size_t)nthreads * 5 * n, sizeof(double));
#pragma omp parallel num_threads(nthreads) reduction(+:total, rt)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        double *Sa = Sacc + (size_t)tid * 5 * n;
        scratch w; scratch_init(&w, n);
#pragma omp for schedule(dynamic, 1)
        for (int s = s_lo; s < s_hi; s++) {
            az_site(&P, s, w.tag);
            inside_site(&P, s, w.tag, d, nu, phi, w.L0, w.L1);
            outside_site(&P, s, w.tag, d, nu, w.L0, w.L1, w.out0, w.out1, w.A, w.X, &w.D, w.path);
            posterior_site(&P, s, w.tag, d, nu, phi, w.L0, w.L1, w.out0, w.out1, root_of, w.post0, w.post1, w.S);
            double l = 0;
            for (int v = 0; v < n; v++) if (parent[v] < 0) l += w.L0[v];
            if (site_llh) site_llh[s] = l;
            total += l;
            for (int j = 0; j < 5 * n; j++) Sa[j] += w.S[j];
            for (int v = 0; v < n; v++)
                if (left[v] < 0 && codes[(size_t)leaf_row[v] * K + s] == -1) rt += 1 - exp(w.post1[v]);
            if (dump) {
                const double *src[13] = { w.L0, w.L1, w.out0, w.out1, w.A, w.X, w.post0, w.post1,
                                          w.S, w.S + n, w.S + 2 * n, w.S + 3 * n, w.S + 4 * n };
                for (int q = 0; q < 13; q++)
                    for (int v = 0; v < n; v++) dump[((size_t)q * n + v) * K + s] = src[q][v];
            }
        }
        scratch_free(&w);
    }
    for (int t = 0; t < nthreads; t++)
        for (int j = 0; j < 5 * n; j++) Ssum[j] += Sacc[(size_t)t * 5 * n + j];
    *Rtilde += rt;
    free(Sacc); free(root_of);
    return total;
}

/* ---- M-step ------------------------------------------------------------------
 * The reference hands the objective at EM_solver.py:351-357 (branch lengths, nu
 * fixed) and :394-399 (nu, branch lengths fixed) to MOSEK through cvxpy; neither
 * package is under /root/reference nor installable here (Mosek>=10.1.16,
 * cvxpy>=1.4, setup.py:36-37).  Without the ultrametric equalities both programs
 * are strictly concave and separable, so their unique optimum is the root of the
 * derivative; this restatement finds it by bisection to the last bit.
 */
static double dobj_dd(double dd, double nu, int use_nu, const double *s) {
    /* d/dd of -(nu+1)S0 d - nu S1 d + S1 log(1-e^-d) + S2 log(1-e^-nu d) - nu S3 d + S4 log(1-e^-nu d) */
    double g = -(nu + 1) * s[0] - nu * (s[1] + s[3]) + s[1] / expm1(dd);
    if (use_nu) g += nu * (s[2] + s[4]) / expm1(nu * dd);
    return g;
}

/* phi_star and normalised statistics, EM_solver.py:318-348.
 * Ssum[k*n+v] raw site-sums; free_mask[v]!=0 marks branches that enter the
 * program (not polytomy-marked, not mark_fixed); Snorm[k*n+v] receives the
 * clamped / rescaled values for those branches (others untouched). */
double oracle_mstep_stats(int n, int K, int n_cells, const double *Ssum, const uint8_t *free_mask,
                          double R, double Rtilde, double eps_s, double *Snorm) {
    double phi_star = Rtilde / (R + Rtilde);                            /* :324 */
    if (fabs(phi_star) < 1.0 / ((double)K * n_cells)) phi_star = 0;     /* :325-326 */
    for (int v = 0; v < n; v++) {
        if (!free_mask[v]) continue;
        double s[5], tot = 0;
        for (int k = 0; k < 5; k++) { s[k] = Ssum[(size_t)k * n + v]; if (s[k] < eps_s) s[k] = eps_s; }  /* :343 */
        for (int k = 0; k < 5; k++) tot += s[k];                        /* python sum(): left to right */
        for (int k = 0; k < 5; k++) Snorm[(size_t)k * n + v] = s[k] / tot * K;                         /* :345 */
    }
    return phi_star;
}

/* argmax over d in [dmin,dmax]^B of the objective at EM_solver.py:351-357 with nu fixed */
void oracle_mstep_brlen(int n, const double *Snorm, const uint8_t *free_mask, double nu, double eps_nu,
                        double dmin, double dmax, double *d) {
    int use_nu = nu > eps_nu;                                           /* :353,355 (sum(S2)>0 always after the clamp) */
    for (int v = 0; v < n; v++) {
        if (!free_mask[v]) continue;
        double s[5];
        for (int k = 0; k < 5; k++) s[k] = Snorm[(size_t)k * n + v];
        double lo = dmin, hi = dmax;
        if (dobj_dd(lo, nu, use_nu, s) <= 0) { d[v] = lo; continue; }
        if (dobj_dd(hi, nu, use_nu, s) >= 0) { d[v] = hi; continue; }
        for (int it = 0; it < 200; it++) {
            double mid = 0.5 * (lo + hi);
            if (mid <= lo || mid >= hi) break;
            if (dobj_dd(mid, nu, use_nu, s) > 0) lo = mid; else hi = mid;
        }
        d[v] = 0.5 * (lo + hi);
    }
}

/* argmax over nu in [0,1] of the objective at EM_solver.py:394-399 with d fixed */
double oracle_mstep_nu(int n, const double *Snorm, const uint8_t *free_mask, const double *d) {
    double lin = 0;
    for (int v = 0; v < n; v++) if (free_mask[v]) lin += (Snorm[v] + Snorm[(size_t)n + v] + Snorm[(size_t)3 * n + v]) * d[v];
    /* g'(nu) = -lin + sum (S2+S4) d / expm1(nu d): +inf at nu -> 0+, strictly decreasing */
#define NU_GRAD(x, out) do { double g_ = -lin; \
        for (int v = 0; v < n; v++) if (free_mask[v]) \
            g_ += (Snorm[(size_t)2 * n + v] + Snorm[(size_t)4 * n + v]) * d[v] / expm1((x) * d[v]); \
        (out) = g_; } while (0)
    double g, lo, hi = 1.0;
    NU_GRAD(hi, g);
    if (g >= 0) return 1.0;                                             /* upper bound active, :402 */
    lo = 0.5;
    for (;;) {                                                          /* bracket the root from below */
        NU_GRAD(lo, g);
        if (g > 0) break;
        hi = lo; lo *= 0.5;
        if (lo < 1e-300) return 0.0;
    }
    for (int it = 0; it < 200; it++) {
        double mid = 0.5 * (lo + hi);
        if (mid <= lo || mid >= hi) break;
        NU_GRAD(mid, g);
        if (g > 0) lo = mid; else hi = mid;
    }
#undef NU_GRAD
    return 0.5 * (lo + hi);
}

/*
 * EM_optimization, EM_solver.py:461-512, without the ultrametric equalities
 * (ultra_constr=False).  d, nu, phi are updated in place.  Returns the number of
 * EM iterations; *nllh receives -curr_llh (:512).  R = number of non-'?' leaf
 * entries (:322).
 */
int oracle_em(int n, int K, int n_cells, const int32_t *parent, const int32_t *left, const int32_t *right,
              const int32_t *order, const int32_t *leaf_row, const int32_t *codes,
              const int32_t *qptr, const double *qval, const uint8_t *free_mask,
              int optimize_phi, int optimize_nu, double dmin, double dmax, double conv_eps, int max_iter,
              int nthreads, double *d, double *nu, double *phi, double *nllh) {
    double R = 0;
    for (int v = 0; v < n; v++) if (left[v] < 0)
        for (int s = 0; s < K; s++) if (codes[(size_t)leaf_row[v] * K + s] != -1) R += 1;
    double *Ssum = (double *)malloc(sizeof(double) * 5 * n);
    double *Snorm = (double *)calloc((size_t)5 * n, sizeof(double));
    double pre = oracle_inside(n, K, parent, left, right, order, leaf_row, codes, qptr, qval, d, *nu, *phi, 0, K, nthreads, NULL, NULL, NULL);
    double cur = pre;
    int it = 1;
    while (it <= max_iter) {                                            /* :475 */
        memset(Ssum, 0, sizeof(double) * 5 * n);
        double rt = 0;
        oracle_estep(n, K, parent, left, right, order, leaf_row, codes, qptr, qval, d, *nu, *phi, 0, K, nthreads, Ssum, &rt, NULL, NULL);
        double phi_star = oracle_mstep_stats(n, K, n_cells, Ssum, free_mask, R, rt, 1e-6, Snorm);
        if (!optimize_phi) phi_star = *phi;                             /* :311-314 */
        oracle_mstep_brlen(n, Snorm, free_mask, *nu, 1e-5, dmin, dmax, d);   /* :415-418, nu_star = old nu */
        if (optimize_nu) *nu = oracle_mstep_nu(n, Snorm, free_mask, d); /* :427-433 */
        *phi = phi_star;                                                /* :437 */
        cur = oracle_inside(n, K, parent, left, right, order, leaf_row, codes, qptr, qval, d, *nu, *phi, 0, K, nthreads, NULL, NULL, NULL); /* :500 */
        if (fabs((cur - pre) / pre) < conv_eps) break;                  /* :503 */
        pre = cur;
        it++;
    }
    *nllh = -cur;
    free(Ssum); free(Snorm);
    return it;
}
