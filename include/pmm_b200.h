/*
 * pmm_b200.h -- C ABI of the B200-native PMM likelihood / EM engine.
 *
 * The reference (raphael-group/LAML v1.0.5) is pure Python and has no FFI; its
 * extension point for this path is the solver-class protocol
 * (laml_libs/Virtual_solver.py:5-17) implemented by EM_solver
 * (laml_libs/EM_solver.py:18-527).  Each entry point below replaces the named
 * method(s) of that class at array level; laml_b200/solver.py is the Python
 * class that binds them behind the reference's method names, and
 * INTEGRATION.md shows the ctypes stub a maintainer would add to the reference.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success, non-zero on
 *     error, and pmm_last_error() returns a human-readable message.
 *   - host pointers unless the name ends in _dev; the caller owns all host
 *     buffers, the handle owns all device memory.
 *   - a handle is bound to one CUDA device and one stream; it is not
 *     thread-safe.  One handle per GPU (one process per GPU under torchrun).
 *   - node ids: leaves 0..n_leaves-1, internal nodes n_leaves..n_nodes-1
 *     (the reference's own array convention, laml_libs/fastEM_solver.py:82-134);
 *     trees are binary (EM_solver binarises in its constructor, EM_solver.py:26-43);
 *     several roots = a forest (ML_solver.py:26-32).
 *   - state codes: 0 unedited; a >= 1 the 1-based rank of the edited state in the
 *     site's prior; -1 '?' (missing); -2 a literal -1 = known silenced
 *     (ML_solver.py:137-142).
 *   - there is NO CPU fallback: every compute entry point fails if no CUDA
 *     device is usable.
 */
#ifndef PMM_B200_H
#define PMM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pmm_handle pmm_handle;

/* Library / ABI version (bumped when a signature changes). */
int pmm_abi_version(void);

/* Message for the last failure on this handle (or for a failed pmm_create when
 * h == NULL).  Valid until the next call on the same handle. */
const char *pmm_last_error(const pmm_handle *h);

/*
 * Replaces ML_solver.__init__'s data model (ML_solver.py:19-86): uploads the
 * integer-coded character matrix and the normalised per-site prior.
 *   site_ptr[n_sites+1] : CSR offsets of the sites' edited alphabets
 *   prior[nnz]          : normalised prior q (ML_solver.py:70-74), by state rank
 *   codes[n_leaves*n_sites] : row-major by leaf id
 *   device              : CUDA device ordinal
 *   mem_budget_bytes    : device memory the handle may use for the per-node
 *                         state vectors (0 = 85% of what is free); if the problem
 *                         needs more, sites are processed in several chunks.
 */
int pmm_create(int32_t n_leaves, int32_t n_sites, const int32_t *site_ptr, const double *prior,
               const int32_t *codes, int32_t device, size_t mem_budget_bytes, pmm_handle **out);

void pmm_destroy(pmm_handle *h);

/* Run the handle's kernels on this cudaStream_t (default: a stream it creates). */
int pmm_set_stream(pmm_handle *h, void *cuda_stream);

/*
 * Replaces the tree the solver walks (treeswift objects, ML_solver.py:26-32):
 * the topology as arrays; the level schedules (height-ordered for the inside
 * pass, depth-ordered for the outside pass) are derived here.
 */
int pmm_set_tree(pmm_handle *h, int32_t n_nodes, const int32_t *parent, const int32_t *left,
                 const int32_t *right);

/*
 * EM_solver.lineage_llh / Estep_in_llh (EM_solver.py:91-141): inside pass.
 *   d[n_nodes] branch lengths (node.edge_length), nu, phi (self.params)
 *   site_llh[n_sites] (optional, may be NULL): per-site log-likelihood summed over trees
 *   llh: sum over trees and sites of log P(D)
 */
int pmm_inside(pmm_handle *h, const double *d, double nu, double phi, double *site_llh, double *llh);

/*
 * EM_solver.Estep (EM_solver.py:302-305) reduced to what Mstep consumes
 * (EM_solver.py:318-348): S[v*5+k] = sum over sites of S_k for the branch above
 * node v; Rtilde = sum over leaves and '?' entries of 1-exp(post1) (:323);
 * llh as pmm_inside.  site_llh optional.
 */
int pmm_estep(pmm_handle *h, const double *d, double nu, double phi, double *S, double *Rtilde,
              double *llh, double *site_llh);

/*
 * Per-(node,site) posteriors for the annotation writer (runner.py:391-435),
 * the only caller that needs them: post0/post1 = log P(node state 0 / silenced | D)
 * (EM_solver.py:276-277), each [n_nodes*n_sites] row-major by node, NULL to skip.
 */
int pmm_posteriors(pmm_handle *h, const double *d, double nu, double phi, double *post0, double *post1);

/*
 * EM_solver.Mstep without the ultrametric equalities (EM_solver.py:307-459 with
 * ultra_constr_cache=None): phi closed form (:318-326), statistics clamp and
 * rescale (:342-346), per-branch maximisation of (:351-357) over [dmin,dmax]
 * with nu fixed, then the 1-D problem in nu (:394-407).
 *   S[n*5], Rtilde: from pmm_estep;  R = number of non-'?' leaf entries (:322)
 *   free_mask[n]: 1 = branch enters the program (not polytomy-marked, not mark_fixed);
 *                 NULL = every branch is free (never "the mask of an earlier call": the handle
 *                 resets it, and pmm_set_tree always does).  The same holds for pmm_em and
 *                 pmm_mstep_dev; a mask equal to the previous one is not uploaded again.
 *   d in/out; nu, phi in/out.
 */
int pmm_mstep(pmm_handle *h, const double *S, double Rtilde, double R, int32_t n_cells,
              const uint8_t *free_mask, int32_t optimize_nu, int32_t optimize_phi, double dmin,
              double dmax, double eps_s, double eps_nu, double *d, double *nu, double *phi);

/*
 * EM_solver.EM_optimization (EM_solver.py:461-512), ultra_constr=False, fully on
 * the device: E-step, M-step and the convergence test |dllh/llh| < conv_eps.
 * status: 0 = "optimal".  iters = em_iter as the reference counts it.
 */
int pmm_em(pmm_handle *h, double *d, double *nu, double *phi, const uint8_t *free_mask,
           int32_t optimize_nu, int32_t optimize_phi, double R, int32_t n_cells, double dmin, double dmax,
           double conv_eps, int32_t max_iter, double *nllh, int32_t *iters, int32_t *status);

/*
 * ONE iteration of that loop (EM_solver.py:483-499: Estep, then Mstep) from host parameters to host parameters:
 * d / nu / phi in, the updated d / nu / phi out, llh = log-likelihood AT THE INPUT point (what the E-step's inside
 * pass evaluates, i.e. the value the reference prints for the previous iteration).  The statistics never leave the
 * device; one synchronisation per call.  This is the per-iteration protocol of a caller that wants to watch or
 * steer the EM (GPU_EM_solver.EM_optimization with verbose output; bench.py's end-to-end step).
 */
int pmm_em_step(pmm_handle *h, double *d, double *nu, double *phi, const uint8_t *free_mask, int32_t optimize_nu,
                int32_t optimize_phi, double R, int32_t n_cells, double dmin, double dmax, double *llh);

/* ---- device-resident variants (multi-GPU site sharding; see laml_b200/dist.py) ----
 * The statistics buffer is [S(n*5) | Rtilde | llh] fp64 on the device; the host
 * layer all-reduces it in place over NCCL between pmm_estep_dev and pmm_mstep_dev. */
/* When the handle holds a site shard, the M-step rescales statistics to the site count
 * of the WHOLE problem (self.numsites, EM_solver.py:345); default = the handle's own. */
int pmm_set_total_sites(pmm_handle *h, int32_t n_sites_total);
int pmm_upload_params(pmm_handle *h, const double *d, double nu, double phi);
int pmm_download_params(pmm_handle *h, double *d, double *nu, double *phi);
int pmm_estep_dev(pmm_handle *h);                 /* params on device -> stats on device */
int pmm_inside_dev(pmm_handle *h);                /* params on device -> llh slot of stats */
void *pmm_stats_dev_ptr(pmm_handle *h);           /* device pointer, (5*n_nodes+2) doubles */
int64_t pmm_stats_len(const pmm_handle *h);
int pmm_mstep_dev(pmm_handle *h, double R, int32_t n_cells, const uint8_t *free_mask, int32_t optimize_nu,
                  int32_t optimize_phi, double dmin, double dmax, double eps_s, double eps_nu);
/* ---- the exchange over peer memory (site shards on the GPUs of one NVSwitch domain; laml_b200/csrc/pmm_peer.cuh) ----
 * The reference sums the expected counts over all sites in one loop (EM_solver.py:247-305); site shards need that sum
 * across ranks.  After pmm_peer_setup the kernel that finishes an E-step / inside pass also reduces the statistics over
 * the ranks -- loads from and stores to the peers' buffers over NVLink -- so that pmm_estep_dev / pmm_inside_dev /
 * pmm_estep / pmm_em / pmm_em_step leave the SAME all-reduced statistics (to the bit) on every rank, with no NCCL call.
 * Every rank must then make the same sequence of those calls (as with any collective).  Lumped-state engine only.
 *   pmm_peer_export      64 opaque bytes (a CUDA IPC handle of this rank's exchange buffer); call after pmm_set_tree
 *   pmm_peer_setup       handles = world x 64 bytes, indexed by rank (own entry ignored); world = 1 drops the peers
 *   pmm_peer_setup_local the same for ranks that are handles of ONE process (peers[rank] may be NULL)
 *   pmm_peer_close       unmap the peers (call on every rank before any of them destroys its handle)
 * A pmm_set_tree that changes the node count frees the exchange buffer: export / set up again afterwards.
 * A rank that waits for a peer in vain (PMM_PEER_TIMEOUT_MS, read at set-up, default 30 s: ranks may drift apart between two
 * calls) gives up instead of hanging the GPU: its statistics are then incomplete, and
 * the next entry point that synchronises with the host (pmm_sync, pmm_read_llh, pmm_download_params, pmm_estep, pmm_inside,
 * pmm_em, pmm_em_step, ...) fails with an error that names the rank.  One rank per GPU: kernels of several ranks that wait
 * for each other on ONE GPU can block each other in the device's work queues (tests only; CUDA_DEVICE_MAX_CONNECTIONS=32).
 */
#define PMM_PEER_HANDLE_BYTES 64
int pmm_peer_export(pmm_handle *h, void *handle_out);
int pmm_peer_setup(pmm_handle *h, int32_t rank, int32_t world, const void *handles);
int pmm_peer_setup_local(pmm_handle *h, int32_t rank, int32_t world, pmm_handle *const *peers);
int pmm_peer_close(pmm_handle *h);
int pmm_read_llh(pmm_handle *h, double *llh);      /* the llh slot of the statistics (after pmm_estep_dev / pmm_inside_dev) */
int pmm_sync(pmm_handle *h);

/*
 * Batched scoring of NNI candidate trees: replaces, for many candidates at
 * once, what Topology_search.apply_nni does one candidate at a time (Topology_search.py:242-322:
 * rebuild a solver from newick strings, score_tree -> optimize with every branch fixed except the
 * five around the move, nu and phi fixed).  Candidate i swaps child cand_k[i] (0 = left, 1 = right)
 * of the internal non-root node cand_u[i] with u's sibling.  d[n] are the current branch lengths
 * (the caller applies ini_brlens' max(2 dmin, .) clamp, ML_solver.py:195-203, if it wants the
 * reference's starting point).  Outputs per candidate: nllh of the whole forest after the local
 * EM (EM_solver.py:461-512 restricted to 5 free branches), the five optimised lengths in the order
 * (v, u, moved child, other child, former sibling), the EM iteration count, and the node ids those
 * lengths belong to.  Needs all sites resident on this handle (1 site chunk).
 */
int pmm_score_nni_batch(pmm_handle *h, int32_t n_cand, const int32_t *cand_u, const int32_t *cand_k,
                        const double *d, double nu, double phi, double dmin, double dmax, double conv_eps,
                        int32_t max_iter, double *out_nllh, double *out_d5, int32_t *out_iters,
                        int32_t *out_nodes);

/*
 * The same under the ultrametric constraint -- strategy['ultra_constr'] of Topology_search.apply_nni, the
 * runner's default (runner.py:91,287): every M-step of a candidate's local EM solves EM_solver.py:351-357 for the
 * five free branches subject to the box and the root-to-leaf equalities (EM_solver.py:53-89,359-361).  With the
 * rest of the tree fixed a candidate has two unknown node heights (three when v is the root of the only tree);
 * the kernel's M-step is pmm_ultrametric_solve specialised to that five-edge tree
 * (laml_b200/csrc/pmm_nni_ultra.cuh).  d[n] must be ultrametric already (leaf paths below every node equal to
 * 1e-9, trees of a forest equally high): otherwise the call fails.  A candidate whose constrained M-step is
 * infeasible or fails (the reference's "d_infeasible" / "failed_d" -> its retry with every branch free,
 * Topology_search.py:301-306) gets out_nllh = NaN and out_iters = -1 and is left to the caller.
 * Lumped-state engine only.
 */
int pmm_score_nni_batch_ultra(pmm_handle *h, int32_t n_cand, const int32_t *cand_u, const int32_t *cand_k,
                              const double *d, double nu, double phi, double dmin, double dmax, double conv_eps,
                              int32_t max_iter, double *out_nllh, double *out_d5, int32_t *out_iters,
                              int32_t *out_nodes);

/*
 * Test hook, pure host code (works without a GPU): the host instance of the constrained M-step the kernel behind
 * pmm_score_nni_batch_ultra runs per candidate and EM iteration.  S25: clamped / rescaled statistics of the
 * branches (v, u, moved child, other child, former sibling), [5][5]; hA / hB / hC: heights of the other child,
 * the former sibling and the moved child; hTop: height of v's parent (ignored when top_free).  H3 (heights of u,
 * v and the top): previous iteration's on input when warm, the optimum on output; d5: the five lengths.
 * Returns 0 = optimal, 1 = infeasible, 2 = failed, -1 = bad arguments.
 */
int pmm_debug_nni_ultra_mstep(const double *S25, double nu, double eps_nu, double lo, double hi, double hA,
                              double hB, double hC, double hTop, int32_t top_free, int32_t warm, double *H3,
                              double *d5, int32_t *newton_iters);

/*
 * Runtime options.  "slices" (1..4, default 2 or PMM_SLICES): the sites of a handle are processed as this many
 * slices on separate CUDA streams so that the latency-bound top-of-the-tree launches of one slice overlap the
 * big launch of another; with 1 the passes run back to back on the handle's stream and pmm_last_timing can
 * split the E-step into its inside and outside pass.
 */
int pmm_set_option(pmm_handle *h, const char *name, int64_t value);

/*
 * Ultrametric-constrained branch-length M-step, pure host code (works without a GPU): replaces the cvxpy -> MOSEK
 * program of EM_solver.Mstep under ultra_constr (EM_solver.py:349-366) and the dense equality rows M d = b of
 * EM_solver.ultrametric_constr (EM_solver.py:53-89; O(N B) Python lists) by a log-barrier Newton method on node
 * heights whose linear systems have the tree's sparsity (laml_b200/csrc/pmm_ultrametric.cpp; O(n) per Newton
 * step).  parent[n]: -1 for roots (all trees of a forest end at the same height, EM_solver.py:80-83).
 * kind[n]: 0 = free branch in [lo, hi], 1 = fixed at fixed_len (mark_fixed), 2 = counts as 0 (polytomy mark,
 * EM_solver.py:60-79).  S[n*5]: the clamped and rescaled statistics of EM_solver.py:343-346 (rows of non-free
 * branches are ignored).  H0[n+1] or NULL: heights returned by the previous EM iteration -- the barrier path then
 * starts from them (pulled warm_theta of the way to an interior point) at weight warm_mu instead of 1.
 * Outputs: d_out[n] branch lengths, H_out[n+1] heights (H_out[n] = the common height), newton_iters.
 * Returns 0 = optimal, 1 = infeasible (the reference's "d_infeasible"), 2 = failed ("failed_d"), -1 = bad arguments.
 */
int pmm_ultrametric_solve(int32_t n, const int32_t *parent, const int8_t *kind, const double *fixed_len, double lo,
                          double hi, const double *S, double nu, double eps_nu, const double *H0, double mu_final,
                          double warm_theta, double warm_mu, double *d_out, double *H_out, int32_t *newton_iters);

/*
 * The same program ON THE DEVICE, fed by the statistics the E-step left there (pmm_estep_dev, or an all-reduce of them):
 * EM_solver.Mstep with ultra_constr_cache (EM_solver.py:307-459) -- lengths under the equalities by one cooperative
 * launch that runs the whole barrier Newton method (laml_b200/csrc/pmm_ultrametric_dev.cuh: level-synchronous tree
 * sweeps, ~45 grid barriers per Newton step at 100k cells), then nu given the new lengths and phi in closed form.
 * kind[n] / fixed_len[n] as for pmm_ultrametric_solve.  warm != 0: start the barrier path from the heights of the
 * previous call on this handle (same problem).  *status: 0 optimal, 1 infeasible ("d_infeasible": nothing is updated),
 * 2 failed ("failed_d").  Results equal the host routine's to its stopping tolerance (summation orders differ).
 */
int pmm_mstep_ultra_dev(pmm_handle *h, double R, int32_t n_cells, const int8_t *kind, const double *fixed_len,
                        int32_t optimize_nu, int32_t optimize_phi, double dmin, double dmax, double eps_s, double eps_nu,
                        int32_t warm, int32_t *status, int32_t *newton_iters);

/*
 * EM_solver.EM_optimization with ultra_constr=True (EM_solver.py:461-512) with every step on the device: E-step,
 * constrained M-step (above, warm-started from the previous iteration), one log-likelihood read per iteration for the
 * convergence test.  status: 0 "optimal", 1 "d_infeasible" (nllh / parameters of the last feasible iterate), 2 "failed_d".
 */
int pmm_em_ultra(pmm_handle *h, double *d, double *nu, double *phi, const int8_t *kind, const double *fixed_len,
                 int32_t optimize_nu, int32_t optimize_phi, double R, int32_t n_cells, double dmin, double dmax,
                 double conv_eps, int32_t max_iter, double *nllh, int32_t *iters, int32_t *status);

/*
 * Test hook, pure host code (works without a GPU): the clade schedule the lumped-state engine derives for a
 * forest -- what pmm_set_tree computes, with the clade capacities given explicitly (1..127).  Outputs (caller
 * allocated; n_int = n_nodes - n_leaves): in_steps / out_steps [n_int][4] = (u, v, w, flags) in walking order
 * (flags: laml_b200/csrc/pmm_lumped.cuh), in_clades / out_clades [n_clades][4] = (first step, end step, tree
 * index, clade id), *_phase_ptr [n_phases + 1] = first clade of every launch.
 */
int pmm_debug_lumped_schedule(int32_t n_leaves, int32_t n_nodes, const int32_t *parent, const int32_t *left,
                              const int32_t *right, int32_t clade_cap, int32_t clade_cap_top, int32_t *in_steps,
                              int32_t *out_steps, int32_t *in_clades, int32_t *out_clades, int32_t *in_phase_ptr,
                              int32_t *out_phase_ptr, int32_t *n_clades, int32_t *n_phases);

/* The waits of the dataflow launches of the same schedule (laml_b200/csrc/pmm_lumped.cuh, "dataflow launches"): CSR
 * lists over clade POSITIONS in in_clades / out_clades -- for the inside pass the clades whose records a clade loads,
 * for the outside pass the clades that store the out-records of its subtree roots.  in_dep_ptr / out_dep_ptr hold
 * n_clades + 1 entries, in_deps at most 2 (n_nodes - n_leaves), out_deps at most n_nodes - n_leaves.  Pure host code.
 */
int pmm_debug_lumped_deps(int32_t n_leaves, int32_t n_nodes, const int32_t *parent, const int32_t *left, const int32_t *right,
                          int32_t clade_cap, int32_t clade_cap_top, int32_t *in_dep_ptr, int32_t *in_deps,
                          int32_t *out_dep_ptr, int32_t *out_deps);

/* ---- introspection for benchmarks ------------------------------------------------- */
typedef struct {
    int64_t n_nodes, n_leaves, n_sites;
    int64_t slots_per_node;      /* W = sum over sites of (A_s + 2): the algorithmic state slots */
    int64_t row_stride;          /* slots per row (W plus even-length padding), largest chunk */
    int64_t padded_slots_per_node; /* the same, summed over all site chunks */
    int64_t n_chunks, n_tiles, tile_slots;
    int64_t inside_levels, outside_levels;
    int64_t device_bytes;        /* allocated by the handle */
    int64_t kernel_launches;     /* launched by this handle so far */
    /* lumped-state engine (the default E-step path; PMM_ENGINE=fullstate selects the row engine) */
    int64_t engine;              /* 0 = lumped records, 1 = full-state rows */
    int64_t clades, clade_target, launches_per_pass;
    int64_t inside_bytes_per_site, outside_bytes_per_site;   /* bytes the lumped passes must move per site */
} pmm_info;
int pmm_get_info(const pmm_handle *h, pmm_info *info);

/* Timing of the last pmm_estep / pmm_inside on the handle's stream (CUDA events), ms.  With site slices
 * (pmm_set_option "slices" > 1) the two passes interleave: inside_ms then covers both and outside_ms is 0. */
int pmm_last_timing(const pmm_handle *h, double *inside_ms, double *outside_ms, double *total_ms);

#ifdef __cplusplus
}
#endif
#endif
