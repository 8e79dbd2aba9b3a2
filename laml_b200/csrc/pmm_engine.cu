// pmm_engine.cu -- host side of the C ABI declared in include/pmm_b200.h.
//
// Owns device memory, derives the level schedules, builds the tiled slot
// layout, and launches the kernels in pmm_kernels.cuh.  No torch types, no CPU
// fallback: if CUDA is unusable every entry point fails with an error string.
#include "../../include/pmm_b200.h"
#include "pmm_kernels.cuh"
#include "pmm_nni.cuh"
#include "pmm_lumped.cuh"
#include "pmm_nni_lumped.cuh"
#include "pmm_ultrametric_dev.cuh"
#include "pmm_peer.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace pmm;

namespace {

thread_local std::string g_create_error;

struct Chunk {
    int site0 = 0, K = 0, Kp = 0, W = 0, Wp = 0, ntiles = 0, max_tile_sites = 0;
    int *d_site_off = nullptr, *d_site_len = nullptr, *d_tile_site = nullptr, *d_tile_off = nullptr;
    int *d_tile_psite = nullptr, *d_site_pidx = nullptr, *d_codes = nullptr;
    int2 *d_ent = nullptr;
    double *d_q = nullptr, *d_qsum = nullptr;
};

}  // namespace

struct pmm_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    // data
    int n_leaves = 0, K = 0;
    int K_total = 0;                // sites of the whole problem when this handle holds a shard
    std::vector<int> site_ptr;      // host copy
    std::vector<double> prior;
    int *d_codes = nullptr;
    int64_t W = 0;                  // algorithmic slots per node
    int TS = 1024;
    int stages_in[4] = {2, 2, 2, 2}, stages_out[4] = {2, 2, 2, 2}, ctas_in[4] = {1, 1, 1, 1}, ctas_out[4] = {1, 1, 1, 1}, n_sms = 148;
    size_t smem_in[4] = {0, 0, 0, 0}, smem_out[4] = {0, 0, 0, 0};
    int ctas_post = 1;
    size_t budget = 0;
    std::vector<Chunk> chunks;
    int maxWp = 0, maxK = 0, maxKp = 0, max_tiles = 0, max_tile_sites = 0;
    // tree
    int n = 0, n_int = 0, n_roots = 0;
    std::vector<int> parent, left, right;
    int *d_parent = nullptr, *d_left = nullptr, *d_right = nullptr, *d_root_index = nullptr, *d_roots = nullptr;
    // level schedules: items (u, v, w, root index) per level, sorted by node class
    int4 *d_in_items = nullptr, *d_out_items = nullptr;
    std::vector<int> in_ptr, out_ptr;              // level -> first item
    std::vector<int> in_cls, out_cls;              // 3 class counts per level (II, IL, LL)
    int spec_threshold = 2000;
    // state
    double *d_X = nullptr, *d_O = nullptr;
    int *d_iexp = nullptr, *d_dexp = nullptr, *d_eX = nullptr, *d_eO = nullptr;
    double *d_sqX = nullptr;
    double *d_rootInvL = nullptr, *d_root_site_llh = nullptr;
    double *d_Spart = nullptr, *d_Rpart = nullptr, *d_Rnode = nullptr;
    double *d_site_llh = nullptr;   // [K]
    double *d_bc = nullptr, *d_d = nullptr, *d_params = nullptr;
    double *d_stats = nullptr;      // [5n | Rtilde | llh]: the `red` part of the exchange block d_xb (not a separate allocation)
    // ---- exchange over peer memory (site shards on several GPUs; pmm_peer.cuh).  d_xb = [loc | red | flags], one allocation
    double *d_xb = nullptr, *d_xloc = nullptr;
    int *d_xflags = nullptr;
    size_t xL = 0;                  // doubles per part (5n + 2 rounded up to even)
    int peer_world = 1, peer_rank = 0, peer_epoch = 0;
    bool peer_lost = false;         // a pmm_set_tree with another node count freed the exchange block the peers had mapped
    PeerArgs peer = {};
    void *peer_mapped[PEER_MAX] = {};    // bases opened through CUDA IPC (to be closed again)
    unsigned *d_peer_bar = nullptr;
    int *d_peer_err = nullptr;
    // scratch of the batch NNI scorer, kept from call to call (a search scores a neighbourhood per iteration; cudaMalloc /
    // cudaFree per call cost up to 100 ms on a virtualised box)
    NniCand *d_nni_c = nullptr; NniState *d_nni_s = nullptr; NniUltra *d_nni_u = nullptr; double *d_nni_rsum = nullptr;
    size_t nni_cap = 0, nni_roots_cap = 0;
    double *d_Snorm = nullptr;
    unsigned char *d_free = nullptr;
    std::vector<unsigned char> free_host;   // the mask d_free holds (empty = all branches free)
    // ---- ultrametric M-step on the device (pmm_ultrametric_dev.cuh): scratch and level phases, built on first use
    bool um_ready = false, um_warm = false;
    std::vector<signed char> um_kind_host;
    std::vector<double> um_fixed_host;
    signed char *um_kind = nullptr;
    double *um_fixed = nullptr, *um_scratch = nullptr, *um_Hprev = nullptr, *um_dnew = nullptr, *um_partial = nullptr;
    unsigned char *um_grounded = nullptr;
    int *um_up_ptr = nullptr, *um_dn_ptr = nullptr, *um_result = nullptr;
    UmPhase *um_up_ph = nullptr, *um_dn_ph = nullptr;
    int um_n_up_ph = 0, um_n_dn_ph = 0;
    unsigned *um_bar = nullptr;
    // pinned staging for the parameter vectors (d | nu | phi): uploads and downloads of the per-iteration protocol
    // (pmm_upload_params / pmm_download_params, i.e. ShardedEngine.em_step) are one DMA each, no pageable bounce
    double *pin_up = nullptr, *pin_down = nullptr;
    size_t pin_cap = 0;
    cudaEvent_t ev_up = nullptr;            // the last upload has left pin_up
    NuState *d_nu_state = nullptr;
    unsigned *d_nu_bar = nullptr;
    double *d_nu_partial = nullptr;
    double *d_post0 = nullptr, *d_post1 = nullptr;
    // ---- lumped-state engine (default): clade schedules and 40-byte records, see pmm_lumped.cuh
    int engine = 0;                 // 0 = lumped (default), 1 = full-state rows (PMM_ENGINE=fullstate)
    bool force_store_all = false;   // the NNI batch scorer needs every out-record of the cached E-step
    bool last_sliced = false;       // the last lumped pass ran site slices on several streams (no inside/outside split)
    bool last_full = false;         // which engine ran the last pass (for pmm_last_timing)
    bool full_ready = false;        // full-state buffers are allocated on first use (NNI batch, PMM_ENGINE=fullstate)
    int clade_target = 0, n_clades = 0;
    std::vector<int> lin_phase_ptr, lout_phase_ptr;     // tier -> first clade (inside: tier 0 first; outside: top tier first)
    // launches of a pass: groups of consecutive tiers; a group of several tiers is ONE dataflow launch (pmm_lumped.cuh)
    struct LGroup { int c0, c1, n1, n2, df; };
    std::vector<LGroup> lin_groups, lout_groups;
    int df_mode = 0;                // 0: one launch per tier; 1: tier 0 | all top tiers (dataflow); 2: all tiers in one dataflow launch
    int *d_df_dep_ptr_in = nullptr, *d_df_deps_in = nullptr, *d_df_dep_ptr_out = nullptr, *d_df_deps_out = nullptr;
    int *d_df_flags = nullptr;      // [2][n_clades][df_stride]
    unsigned *d_df_tickets = nullptr;
    int *d_df_err = nullptr;
    unsigned df_ticket_host[64] = {};
    int df_epoch = 0, df_stride = 0;
    size_t df_deps_cap_in = 0, df_deps_cap_out = 0;
    int4 *d_lin_steps = nullptr, *d_lout_steps = nullptr, *d_lin_clades = nullptr, *d_lout_clades = nullptr;
    int lKc = 0, lG = 0;            // sites per chunk, site groups (warps) per chunk
    // site slices: the sites of a chunk are split into `slices` ranges whose launches go to separate streams,
    // so that one slice's latency-bound top-tier launches overlap another slice's big tier-0 launch
    int slices = 2;
    int stagger = 0;
    int pdl = 1;                    // programmatic dependent launch between consecutive lumped launches of a stream (PMM_PDL)
    cudaStream_t aux[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // per slice: normal, high priority
    cudaEvent_t ev_fork = nullptr, ev_s[4][3] = {};
    double2 *d_LX = nullptr;        // packed 32-byte inside records (two double2 each)
    double4 *d_LO = nullptr;
    int2 *d_LOI = nullptr;
    double *d_LinvL = nullptr, *d_Lroot_llh = nullptr, *d_LSpart = nullptr, *d_LRpart = nullptr, *d_qsum = nullptr, *d_prior = nullptr;
    int *d_site_ptr = nullptr;
    int64_t l_in_bytes_per_site = 0, l_out_bytes_per_site = 0;   // algorithmic bytes the lumped passes move per site
    size_t device_bytes = 0;
    int64_t launches = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float t_inside = 0, t_outside = 0, t_total = 0;
    bool timing_valid = false;
};

namespace {

int fail(pmm_handle *h, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return 1;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            (void)cudaGetLastError(); /* do not leave a sticky error behind for the next caller */   \
            return fail(h, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                          \
    } while (0)

template <class T> int dalloc(pmm_handle *h, T **p, size_t count) {
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    cudaError_t e = cudaMalloc((void **)p, bytes);
    if (e != cudaSuccess) return fail(h, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    h->device_bytes += bytes;
    return 0;
}
template <class T> void dfree(T *&p) { if (p) cudaFree(p); p = nullptr; }

int env_int(const char *name, int dflt) {
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// forget the peers (their buffers are unmapped; ours stays): back to a single-GPU handle
void peer_drop(pmm_handle *h) {
    for (auto &m : h->peer_mapped) if (m) { cudaIpcCloseMemHandle(m); m = nullptr; }
    h->peer_world = 1; h->peer_rank = 0; h->peer_epoch = 0;
    h->peer = PeerArgs{};
}

void free_chunks(pmm_handle *h) {
    for (auto &c : h->chunks) {
        dfree(c.d_site_off); dfree(c.d_site_len); dfree(c.d_tile_site); dfree(c.d_tile_off); dfree(c.d_ent); dfree(c.d_q); dfree(c.d_qsum);
        dfree(c.d_tile_psite); dfree(c.d_site_pidx); dfree(c.d_codes);
    }
    h->chunks.clear();
}

void free_tree_state(pmm_handle *h) {
    dfree(h->d_parent); dfree(h->d_left); dfree(h->d_right); dfree(h->d_root_index); dfree(h->d_roots);
    dfree(h->d_in_items); dfree(h->d_out_items);
    dfree(h->d_X); dfree(h->d_O); dfree(h->d_iexp); dfree(h->d_dexp); dfree(h->d_eX); dfree(h->d_eO); dfree(h->d_sqX);
    dfree(h->d_rootInvL); dfree(h->d_root_site_llh);
    dfree(h->d_Spart); dfree(h->d_Rpart); dfree(h->d_Rnode); dfree(h->d_site_llh);
    if (h->peer_world > 1) h->peer_lost = true;
    peer_drop(h);
    dfree(h->d_bc); dfree(h->d_d); dfree(h->d_xb); h->d_stats = h->d_xloc = nullptr; h->d_xflags = nullptr; dfree(h->d_peer_bar); dfree(h->d_peer_err);
    dfree(h->d_Snorm); dfree(h->d_free); dfree(h->d_nu_state); dfree(h->d_nu_partial); dfree(h->d_nu_bar);
    dfree(h->d_post0); dfree(h->d_post1);
    dfree(h->d_lin_steps); dfree(h->d_lout_steps); dfree(h->d_lin_clades); dfree(h->d_lout_clades);
    dfree(h->d_df_dep_ptr_in); dfree(h->d_df_deps_in); dfree(h->d_df_dep_ptr_out); dfree(h->d_df_deps_out);
    dfree(h->d_df_flags); dfree(h->d_df_tickets); dfree(h->d_df_err);
    h->df_deps_cap_in = h->df_deps_cap_out = 0;
    dfree(h->d_LX); dfree(h->d_LO); dfree(h->d_LOI); dfree(h->d_LinvL); dfree(h->d_Lroot_llh);
    dfree(h->d_LSpart); dfree(h->d_LRpart);
    dfree(h->um_kind); dfree(h->um_fixed); dfree(h->um_scratch); dfree(h->um_Hprev); dfree(h->um_dnew); dfree(h->um_partial);
    dfree(h->um_grounded); dfree(h->um_up_ptr); dfree(h->um_dn_ptr); dfree(h->um_result); dfree(h->um_up_ph); dfree(h->um_dn_ph); dfree(h->um_bar);
    h->um_ready = false; h->um_warm = false; h->um_kind_host.clear(); h->um_fixed_host.clear();
    h->full_ready = false;
    free_chunks(h);
}

// Tiled slot layout of the sites [s_lo, s_hi): see the header of pmm_kernels.cuh.
int build_chunk(pmm_handle *h, int s_lo, int s_hi, Chunk &c) {
    const int K = s_hi - s_lo;
    std::vector<int> site_off(K), site_len(K), tile_site{0}, tile_off{0};
    int off = 0, cur = 0;
    for (int s = 0; s < K; s++) {
        int len = (h->site_ptr[s_lo + s + 1] - h->site_ptr[s_lo + s] + 2 + 1) & ~1;   // padded to even
        if (cur > 0 && cur + len > h->TS) {
            tile_site.push_back(s);
            tile_off.push_back(off);
            cur = 0;
        }
        site_off[s] = off;
        site_len[s] = len;
        off += len;
        cur += len;
    }
    tile_site.push_back(K);
    tile_off.push_back(off);
    c.site0 = s_lo; c.K = K; c.Wp = off; c.ntiles = (int)tile_site.size() - 1;
    c.W = h->site_ptr[s_hi] - h->site_ptr[s_lo] + 2 * K;
    std::vector<double> q((size_t)std::max(off, 1), 0.0), qsum((size_t)K, 0.0);
    for (int s = 0; s < K; s++) {
        int A = h->site_ptr[s_lo + s + 1] - h->site_ptr[s_lo + s];
        double acc = 0.0;
        for (int a = 0; a < A; a++) { double x = h->prior[h->site_ptr[s_lo + s] + a]; q[site_off[s] + 2 + a] = x; acc += x; }
        qsum[s] = acc;
    }
    // per tile: sites sorted by decreasing length, so that the lane groups of one warp (which take
    // consecutive entries) run similar trip counts and the round-robin assignment is balanced
    std::vector<int2> ent((size_t)K);
    c.max_tile_sites = 0;
    for (int t = 0; t < c.ntiles; t++) {
        const int a = tile_site[t], b = tile_site[t + 1];
        std::vector<int> idx(b - a);
        for (int i = 0; i < b - a; i++) idx[i] = i;
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return site_len[a + x] > site_len[a + y]; });
        for (int i = 0; i < b - a; i++) {
            const int ls = idx[i];
            ent[a + i].x = (site_off[a + ls] - tile_off[t]) | (site_len[a + ls] << 16);
            ent[a + i].y = ls;
        }
        c.max_tile_sites = std::max(c.max_tile_sites, b - a);
    }
    c.max_tile_sites = (c.max_tile_sites + 3) & ~3;
    // padded site index space: every tile's sites start at a multiple of 4, so the tile's segment
    // of a per-(node,site) int32/fp64 array is one aligned bulk copy
    std::vector<int> tile_psite(c.ntiles + 1), site_pidx(K);
    int pp = 0;
    for (int t = 0; t < c.ntiles; t++) {
        tile_psite[t] = pp;
        for (int s2 = tile_site[t]; s2 < tile_site[t + 1]; s2++) site_pidx[s2] = pp + (s2 - tile_site[t]);
        pp += (tile_site[t + 1] - tile_site[t] + 3) & ~3;
    }
    tile_psite[c.ntiles] = pp;
    c.Kp = pp;
    if (dalloc(h, &c.d_tile_psite, tile_psite.size()) || dalloc(h, &c.d_site_pidx, K) || dalloc(h, &c.d_codes, (size_t)h->n_leaves * c.Kp))
        return 1;
    CK(cudaMemcpy(c.d_tile_psite, tile_psite.data(), sizeof(int) * tile_psite.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.d_site_pidx, site_pidx.data(), sizeof(int) * K, cudaMemcpyHostToDevice));
    CK(cudaMemset(c.d_codes, 0, sizeof(int) * (size_t)h->n_leaves * c.Kp));
    {
        const long long tot = (long long)h->n_leaves * K;
        pmm_pad_codes<<<(unsigned)((tot + 255) / 256), 256>>>(h->d_codes, h->K, s_lo, K, c.d_site_pidx, c.Kp, h->n_leaves, c.d_codes);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
    }
    if (dalloc(h, &c.d_site_off, K) || dalloc(h, &c.d_site_len, K) || dalloc(h, &c.d_tile_site, tile_site.size()) ||
        dalloc(h, &c.d_tile_off, tile_off.size()) || dalloc(h, &c.d_q, q.size()) || dalloc(h, &c.d_qsum, K) || dalloc(h, &c.d_ent, K))
        return 1;
    CK(cudaMemcpy(c.d_site_off, site_off.data(), sizeof(int) * K, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.d_site_len, site_len.data(), sizeof(int) * K, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.d_tile_site, tile_site.data(), sizeof(int) * tile_site.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.d_tile_off, tile_off.data(), sizeof(int) * tile_off.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.d_q, q.data(), sizeof(double) * q.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.d_qsum, qsum.data(), sizeof(double) * K, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.d_ent, ent.data(), sizeof(int2) * K, cudaMemcpyHostToDevice));
    return 0;
}

// Split the sites into chunks whose rows fit the memory budget, balanced by slots.
int plan_chunks(pmm_handle *h) {
    free_chunks(h);
    const int K = h->K;
    const double per_slot = 16.0 * std::max(h->n_int, 1);            // X + O rows, bytes per slot column
    const double per_site = 24.0 * std::max(h->n_int, 1);            // iexp, dexp, eX, eO (int32) + sqX (fp64)
    double total = per_slot * (double)(h->W + 2 * K / std::max(h->TS / 64, 1)) + per_site * K;
    size_t budget = h->budget;
    if (budget == 0) {
        size_t fr = 0, tot = 0;
        cudaError_t e = cudaMemGetInfo(&fr, &tot);
        if (e != cudaSuccess) return fail(h, "cudaMemGetInfo failed: %s", cudaGetErrorString(e));
        budget = (size_t)(0.85 * (double)fr);
    }
    int forced = env_int("PMM_CHUNKS", 0);
    int nch = forced > 0 ? forced : (int)std::ceil(total / (double)budget);
    nch = std::max(1, std::min(nch, std::max(K, 1)));
    // contiguous ranges with ~equal slot counts
    std::vector<int> bounds{0};
    int64_t acc = 0;
    for (int s = 0; s < K; s++) {
        acc += h->site_ptr[s + 1] - h->site_ptr[s] + 2;
        int want = (int)bounds.size();
        if (want < nch && acc * nch >= (int64_t)want * h->W) bounds.push_back(s + 1);
    }
    if (bounds.back() != K) bounds.push_back(K);
    h->chunks.resize(bounds.size() - 1);
    h->maxWp = h->maxK = h->maxKp = h->max_tiles = h->max_tile_sites = 0;
    for (size_t c = 0; c + 1 < bounds.size(); c++) {
        if (build_chunk(h, bounds[c], bounds[c + 1], h->chunks[c])) return 1;
        h->maxWp = std::max(h->maxWp, h->chunks[c].Wp);
        h->maxK = std::max(h->maxK, h->chunks[c].K);
        h->maxKp = std::max(h->maxKp, h->chunks[c].Kp);
        h->max_tiles = std::max(h->max_tiles, h->chunks[c].ntiles);
        h->max_tile_sites = std::max(h->max_tile_sites, h->chunks[c].max_tile_sites);
    }
    return 0;
}

KParams make_params(pmm_handle *h, const Chunk &c, bool with_post) {
    KParams P;
    P.site_off = c.d_site_off; P.site_len = c.d_site_len; P.tile_site = c.d_tile_site; P.tile_off = c.d_tile_off;
    P.tile_psite = c.d_tile_psite; P.site_pidx = c.d_site_pidx;
    P.ent = c.d_ent; P.q = c.d_q; P.qsum = c.d_qsum;
    P.K = c.K; P.Kp = c.Kp; P.Wp = c.Wp; P.ntiles = c.ntiles; P.max_tile_sites = h->max_tile_sites;
    P.codes = c.d_codes; P.site0 = c.site0;
    P.left = h->d_left; P.right = h->d_right; P.parent = h->d_parent; P.root_index = h->d_root_index;
    P.n_leaves = h->n_leaves; P.n_nodes = h->n;
    P.params = h->d_params; P.bc = h->d_bc;
    P.X = h->d_X; P.O = h->d_O; P.iexp = h->d_iexp; P.dexp = h->d_dexp; P.eX = h->d_eX; P.eO = h->d_eO; P.sqX = h->d_sqX;
    P.rootInvL = h->d_rootInvL; P.root_site_llh = h->d_root_site_llh;
    P.Spart = h->d_Spart; P.Rpart = h->d_Rpart;
    P.post0 = with_post ? h->d_post0 : nullptr; P.post1 = with_post ? h->d_post1 : nullptr;
    P.K_total = h->K;
    P.TS = h->TS; P.stages = 0;
    return P;
}

// Shared-memory ring depth and resident CTAs per SM for both level kernels, given the tile
// geometry.  Called once the chunk plan (hence max_tile_sites) is known.
int configure_kernels(pmm_handle *h) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->device));
    h->n_sms = prop.multiProcessorCount;
    const size_t smem_cap = prop.sharedMemPerBlockOptin;       // 227 KB on sm_100
    const size_t smem_sm = prop.sharedMemPerMultiprocessor;
    const int MS = h->max_tile_sites;
    const int want_in = env_int("PMM_STAGES_IN", 3), want_out = env_int("PMM_STAGES_OUT", 2);
    const int want_ctas = env_int("PMM_CTAS", 2);
    // per (pass, class): the deepest ring (<= wanted) that still leaves `want_ctas` CTAs resident per SM
#define SETUP(OUT, CLS, IDX)                                                                                                  \
    {                                                                                                                         \
        int S = std::min(OUT ? want_out : want_in, MAX_STAGES);                                                               \
        size_t b = 0;                                                                                                         \
        for (; S >= 1; S--) {                                                                                                 \
            b = Plan<OUT, CLS>::total_bytes(h->TS, MS, S);                                                                    \
            if (b <= smem_cap && (S == 1 || (b + 1024) * want_ctas <= smem_sm)) break;                                        \
        }                                                                                                                     \
        if (S < 1) return fail(h, "tile of %d slots does not fit in shared memory", h->TS);                                   \
        int ctas = 0;                                                                                                         \
        CK(cudaFuncSetAttribute(pmm_level<OUT, CLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b));                   \
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, pmm_level<OUT, CLS>, NTHREADS, b));                           \
        if (ctas < 1) return fail(h, "level kernel cannot be resident (smem %zu bytes)", b);                                  \
        (OUT ? h->stages_out : h->stages_in)[IDX] = S;                                                                        \
        (OUT ? h->smem_out : h->smem_in)[IDX] = b;                                                                            \
        (OUT ? h->ctas_out : h->ctas_in)[IDX] = ctas;                                                                         \
        if (env_int("PMM_VERBOSE", 0))                                                                                        \
            fprintf(stderr, "[pmm] %s class %d: TS=%d stages=%d smem=%zu B, %d CTA/SM\n", OUT ? "outside" : "inside", IDX, h->TS, S, b, ctas); \
    }
    SETUP(false, CLS_II, 0) SETUP(false, CLS_IL, 1) SETUP(false, CLS_LL, 2) SETUP(false, CLS_ANY, 3)
    SETUP(true, CLS_II, 0) SETUP(true, CLS_IL, 1) SETUP(true, CLS_LL, 2) SETUP(true, CLS_ANY, 3)
#undef SETUP
    CK(cudaFuncSetAttribute(pmm_level<true, CLS_ANY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_out[CLS_ANY]));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->ctas_post, pmm_level<true, CLS_ANY, true>, NTHREADS, h->smem_out[CLS_ANY]));
    if (h->ctas_post < 1) return fail(h, "posterior kernel cannot be resident");
    h->spec_threshold = env_int("PMM_SPEC_THRESHOLD", 2000);   // levels with fewer nodes take one class-agnostic launch
    return 0;
}

// grid = (P, ntiles): P persistent CTAs per tile, sized to fill the SMs once
template <bool OUTSIDE, int CLS> void launch_one(pmm_handle *h, KParams P, const int4 *d_items, int cnt) {
    if (cnt <= 0) return;
    const int resident = h->n_sms * (OUTSIDE ? h->ctas_out[CLS] : h->ctas_in[CLS]);
    int per_tile = std::max(1, resident / std::max(P.ntiles, 1));
    per_tile = std::min(per_tile, cnt);
    P.stages = OUTSIDE ? h->stages_out[CLS] : h->stages_in[CLS];
    dim3 grid((unsigned)per_tile, (unsigned)P.ntiles);
    pmm_level<OUTSIDE, CLS><<<grid, NTHREADS, OUTSIDE ? h->smem_out[CLS] : h->smem_in[CLS], h->stream>>>(P, d_items, cnt);
    h->launches++;
}

// one tree level: big levels run one class-specialised launch per node class (the items are
// sorted by class), small ones a single launch that reads the class per item
template <bool OUTSIDE> void launch_level(pmm_handle *h, const KParams &P, const int4 *d_items, const int *cls_cnt) {
    const int cnt = cls_cnt[0] + cls_cnt[1] + cls_cnt[2];
    if (OUTSIDE && P.post0) {
        // annotation path (runner.py:391-435): per-(node,site) posteriors; one class-agnostic launch
        if (cnt <= 0) return;
        KParams Q = P;
        Q.stages = h->stages_out[CLS_ANY];
        const int resident = h->n_sms * h->ctas_post;
        int per_tile = std::min(std::max(1, resident / std::max(P.ntiles, 1)), cnt);
        dim3 grid((unsigned)per_tile, (unsigned)P.ntiles);
        pmm_level<true, CLS_ANY, true><<<grid, NTHREADS, h->smem_out[CLS_ANY], h->stream>>>(Q, d_items, cnt);
        h->launches++;
        return;
    }
    if (cnt < h->spec_threshold) { launch_one<OUTSIDE, CLS_ANY>(h, P, d_items, cnt); return; }
    launch_one<OUTSIDE, CLS_II>(h, P, d_items, cls_cnt[0]);
    launch_one<OUTSIDE, CLS_IL>(h, P, d_items + cls_cnt[0], cls_cnt[1]);
    launch_one<OUTSIDE, CLS_LL>(h, P, d_items + cls_cnt[0] + cls_cnt[1], cls_cnt[2]);
}

int launch_inside(pmm_handle *h, const KParams &P) {
    const int nl = (int)h->in_ptr.size() - 1;
    for (int l = 0; l < nl; l++) launch_level<false>(h, P, h->d_in_items + h->in_ptr[l], &h->in_cls[3 * l]);
    pmm_root_llh<<<(unsigned)(h->n_roots * P.ntiles), 256, 0, h->stream>>>(P, h->d_roots);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

int launch_outside(pmm_handle *h, const KParams &P) {
    const int nl = (int)h->out_ptr.size() - 1;
    for (int l = 0; l < nl; l++) launch_level<true>(h, P, h->d_out_items + h->out_ptr[l], &h->out_cls[3 * l]);
    CK(cudaGetLastError());
    return 0;
}


// ======================================================== lumped-state engine: host side
// Partition the forest into TIERS and write, per clade (= one CTA's work), the step lists the kernels
// in pmm_lumped.cuh walk.  Tier 0 holds every maximal subtree with fewer than M0 internal nodes: all of
// them are independent, so one launch covers ~97 % of the tree; what is left (the nodes with >= M0
// internal descendants, ~2 n / M0 of them) is a small "top tree" that is peeled the same way with a
// small threshold, because its launches are latency-bound chains.  The subtrees of a tier are packed
// (best fit, decreasing) into clades of at most C steps.  `order` lists parents before children.
struct LumpedSchedule {
    std::vector<int4> in_steps, out_steps, in_clades, out_clades;
    std::vector<int> in_phase_ptr, out_phase_ptr;       // launch -> first clade
    std::vector<size_t> tier_clades;
    // dataflow launches: per clade POSITION in in_clades / out_clades, the positions of the clades it waits for (CSR)
    std::vector<int> in_dep_ptr, in_deps, out_dep_ptr, out_deps;
    int n_clades = 0, n_tiers = 0;
    int64_t in_loads = 0, out_loads = 0, out_stores = 0;
};

// pure host function (no CUDA): also reachable through pmm_debug_lumped_schedule for the CPU tests
void make_lumped_schedule(int N, int n, const int *L, const int *R, const int *Pa, const std::vector<int> &order,
                          const std::vector<int> &root_index, int C0, int Ctop, LumpedSchedule &S) {
    const int n_int = n - N;
    // ---- tiers
    std::vector<int> tier(n, -1), tsz(n, 0), sz(n, 0);
    int remaining = n_int, n_tiers = 0;
    while (remaining > 0) {
        const int M = (n_tiers == 0 ? C0 : Ctop) + 1;
        for (int i = n - 1; i >= 0; i--) {
            const int u = order[i];
            if (u < N || tier[u] >= 0) { sz[u] = 0; continue; }
            sz[u] = 1 + sz[L[u]] + sz[R[u]];
        }
        for (int u = N; u < n; u++)
            if (tier[u] < 0 && sz[u] < M) { tier[u] = n_tiers; tsz[u] = sz[u]; remaining--; }
        for (int u = N; u < n; u++) if (tier[u] == n_tiers) sz[u] = 0;
        n_tiers++;
    }
    auto open_internal = [&](int u, int c) { return c >= N && tier[c] == tier[u]; };

    // ---- clades: subtree roots of each tier packed into bins of <= C steps, one tree of the forest per bin
    struct Bin { std::vector<int> roots; int root_index; };
    std::vector<std::vector<Bin>> bins(n_tiers);
    {
        std::vector<std::vector<int>> roots_of(n_tiers);
        for (int i = 0; i < n; i++) {
            const int u = order[i];
            if (u >= N && (Pa[u] < 0 || tier[Pa[u]] != tier[u])) roots_of[tier[u]].push_back(u);
        }
        for (int t = 0; t < n_tiers; t++) {
            const int C = (t == 0) ? C0 : Ctop;
            std::vector<int> &rs = roots_of[t];
            std::stable_sort(rs.begin(), rs.end(), [&](int a, int b) {
                return root_index[a] != root_index[b] ? root_index[a] < root_index[b] : tsz[a] > tsz[b]; });
            std::multimap<int, int> open_bins;          // remaining capacity -> bin (of the current tree)
            int cur_tree = -1;
            for (int r : rs) {
                if (root_index[r] != cur_tree) { open_bins.clear(); cur_tree = root_index[r]; }
                auto it = open_bins.lower_bound(tsz[r]);
                int b;
                if (it == open_bins.end()) { b = (int)bins[t].size(); bins[t].push_back(Bin{{}, cur_tree}); it = open_bins.emplace(C, b); }
                b = it->second;
                const int rem = it->first - tsz[r];
                open_bins.erase(it);
                if (rem > 0) open_bins.emplace(rem, b);
                bins[t][b].roots.push_back(r);
            }
        }
    }

    std::vector<int4> &in_steps = S.in_steps, &out_steps = S.out_steps, &in_clades = S.in_clades, &out_clades = S.out_clades;
    in_steps.clear(); out_steps.clear(); in_clades.clear(); out_clades.clear();
    in_steps.reserve(n_int); out_steps.reserve(n_int);
    S.in_phase_ptr.assign(1, 0); S.out_phase_ptr.assign(1, 0);
    int64_t &in_loads = S.in_loads, &out_loads = S.out_loads, &out_stores = S.out_stores;
    in_loads = out_loads = out_stores = 0;
    int ncl = 0;
    std::vector<std::vector<int>> bin_id(n_tiers);
    std::vector<std::pair<int, int>> stk;
    std::vector<int> clade_of(n, -1), prod;             // internal node -> clade id (= position in in_clades)
    S.in_dep_ptr.assign(1, 0); S.in_deps.clear(); S.out_dep_ptr.assign(1, 0); S.out_deps.clear();
    auto close_deps = [&](std::vector<int> &ptr, std::vector<int> &deps) {
        std::sort(prod.begin(), prod.end());
        prod.erase(std::unique(prod.begin(), prod.end()), prod.end());
        deps.insert(deps.end(), prod.begin(), prod.end());
        ptr.push_back((int)deps.size());
        prod.clear();
    };
    // ---- inside: post-order; of two open internal children the bigger subtree goes first, the one
    // visited last stays in registers (KIND_PREV), everything else is loaded
    for (int t = 0; t < n_tiers; t++) {
        for (const Bin &bin : bins[t]) {
            bin_id[t].push_back(ncl);
            const int begin = (int)in_steps.size();
            for (int r : bin.roots) {
                stk.clear(); stk.push_back({r, 0});
                while (!stk.empty()) {
                    const int u = stk.back().first, stt = stk.back().second;
                    const int a = L[u], b = R[u];
                    int first = -1, second = -1;
                    const bool oa = open_internal(u, a), ob = open_internal(u, b);
                    if (oa && ob) { if (tsz[a] >= tsz[b]) { first = a; second = b; } else { first = b; second = a; } }
                    else if (oa) second = a;
                    else if (ob) second = b;
                    if (stt == 0) { stk.back().second = 1; if (first >= 0) stk.push_back({first, 0}); continue; }
                    if (stt == 1) { stk.back().second = 2; if (second >= 0) stk.push_back({second, 0}); continue; }
                    stk.pop_back();
                    // the recursion is symmetric in the children: list the chained one (if any) second
                    const int cv = (second == a) ? b : a, cw = (second == a) ? a : b;
                    auto kind = [&](int c) { return c < N ? KIND_LEAF : (c == second ? KIND_PREV : KIND_LOAD); };
                    const int kv = kind(cv), kw = kind(cw);
                    in_loads += (kv == KIND_LOAD) + (kw == KIND_LOAD);
                    // bits 5 / 6: the loaded record belongs to a lower tier, i.e. an earlier launch wrote it
                    const int far = ((kv == KIND_LOAD && tier[cv] != t) ? 32 : 0) | ((kw == KIND_LOAD && tier[cw] != t) ? 64 : 0);
                    in_steps.push_back(make_int4(u, cv, cw, kv | (kw << 2) | (Pa[u] < 0 ? 16 : 0) | far));
                    clade_of[u] = ncl;
                }
            }
            // producers: the clades (of lower tiers, numbered earlier) that hold the children marked "far"
            for (int i = begin; i < (int)in_steps.size(); i++) {
                const int4 d = in_steps[i];
                if (d.w & 32) prod.push_back(clade_of[d.y]);
                if (d.w & 64) prod.push_back(clade_of[d.z]);
            }
            close_deps(S.in_dep_ptr, S.in_deps);
            in_clades.push_back(make_int4(begin, (int)in_steps.size(), bin.root_index, ncl));
            ncl++;
        }
        S.in_phase_ptr.push_back((int)in_clades.size());
    }
    // ---- outside: pre-order, top tiers first; of two open internal children the smaller subtree
    // goes first and its out-record stays in registers, the other one is stored and loaded later
    std::vector<std::pair<int, int>> stk2;
    std::vector<int> out_pos(ncl, -1);                 // clade id -> position in out_clades (top tiers first)
    for (int t = n_tiers - 1; t >= 0; t--) {
        for (size_t bi = 0; bi < bins[t].size(); bi++) {
            const Bin &bin = bins[t][bi];
            const int begin = (int)out_steps.size();
            out_pos[bin_id[t][bi]] = (int)out_clades.size();
            // producers: the (higher-tier, hence earlier) clades whose steps store the out-records of this clade's subtree roots
            for (int r : bin.roots) if (Pa[r] >= 0) prod.push_back(out_pos[clade_of[Pa[r]]]);
            close_deps(S.out_dep_ptr, S.out_deps);
            for (int r : bin.roots) {
                stk2.clear(); stk2.push_back({r, KIND_LOAD});      // a root's out-record is stored by the inside pass
                while (!stk2.empty()) {
                    const int u = stk2.back().first, how = stk2.back().second;
                    stk2.pop_back();
                    const int a = L[u], b = R[u];
                    int first = -1, second = -1;
                    const bool oa = open_internal(u, a), ob = open_internal(u, b);
                    if (oa && ob) { if (tsz[a] <= tsz[b]) { first = a; second = b; } else { first = b; second = a; } }
                    else if (oa) first = a;
                    else if (ob) first = b;
                    // the child whose out-record stays in registers (if any) is listed first
                    const int cv = (first == b) ? b : a, cw = (first == b) ? a : b;
                    int f = how | (cv < N ? 4 : 0) | (cw < N ? 8 : 0);
                    if (cv >= N && cv != first) { f |= 16; out_stores++; }
                    if (cw >= N) { f |= 32; out_stores++; }
                    if (first >= 0) f |= 64;
                    out_loads += (how == KIND_LOAD);
                    out_steps.push_back(make_int4(u, cv, cw, f));
                    if (second >= 0) stk2.push_back({second, KIND_LOAD});
                    if (first >= 0) stk2.push_back({first, KIND_PREV});
                }
            }
            out_clades.push_back(make_int4(begin, (int)out_steps.size(), bin.root_index, bin_id[t][bi]));
        }
        S.out_phase_ptr.push_back((int)out_clades.size());
    }
    S.n_clades = ncl; S.n_tiers = n_tiers;
    S.tier_clades.clear();
    for (int t = 0; t < n_tiers; t++) S.tier_clades.push_back(bins[t].size());
}

int build_lumped_schedule(pmm_handle *h, const std::vector<int> &order, const std::vector<int> &root_index) {
    const int N = h->n_leaves, n = h->n_int + N;
    int C0 = env_int("PMM_CLADE", 0);
    if (C0 <= 0) {
        // enough CTAs for ~5.5 waves over the SMs (at four resident CTAs per SM), chains of 16..120 steps; measured on a 250-site
        // shard of the 100k-cell tree: capacity 42 (8 waves) 0.663 ms per E-step, 52: 0.653, 60: 0.649, 68: 0.642-0.654
        // (profiles/r2_clade_sweeps.log)
        const double site_blocks = (double)((h->K + LT - 1) / LT);
        const double want_clades = 5.5 * h->n_sms * 4.0 / site_blocks;
        C0 = (int)((double)h->n_int / std::max(want_clades, 1.0));
        C0 = std::max(16, std::min(120, C0));
    }
    C0 = std::max(1, std::min(C0, STEP_CAP - 1));      // the kernels stage a clade's steps in shared memory
    // dataflow launches (pmm_lumped.cuh), PMM_DF = 1 / 2: several tiers per launch.  Off by default: measured on a B200
    // (profiles/r2_dataflow.md) a hand-over between clades through a mark costs about as much as six steps, and with
    // programmatic dependent launch a tier boundary costs no more -- 1k x 30: 70 us either way, 10k x 500: 218 -> 256 us,
    // a 250-site shard of the 100k-cell tree: 667 -> 720 us
    h->df_mode = std::max(0, std::min(2, env_int("PMM_DF", 0)));
    const int Ctop = std::max(1, std::min(env_int("PMM_CLADE_TOP", 16), STEP_CAP - 1));
    h->clade_target = C0;
    LumpedSchedule S;
    make_lumped_schedule(N, n, h->left.data(), h->right.data(), h->parent.data(), order, root_index, C0, Ctop, S);
    const std::vector<int4> &in_steps = S.in_steps, &out_steps = S.out_steps, &in_clades = S.in_clades, &out_clades = S.out_clades;
    h->lin_phase_ptr = S.in_phase_ptr; h->lout_phase_ptr = S.out_phase_ptr;
    const int64_t in_loads = S.in_loads, out_loads = S.out_loads, out_stores = S.out_stores;
    const int ncl = S.n_clades, n_tiers = S.n_tiers;
    h->n_clades = ncl;
    if ((int)in_steps.size() != h->n_int || (int)out_steps.size() != h->n_int) return fail(h, "internal error: clade schedule does not cover the tree");
    // algorithmic bytes per site: 32-byte inside / 40-byte outside records written once and read where not chained, 4-byte leaf codes,
    // 40 bytes of expected counts per (branch, 32 sites)
    h->l_in_bytes_per_site = 32 * (int64_t)h->n_int + 32 * in_loads + 4 * (int64_t)N + (16 + 40) * (int64_t)h->n_roots;
    h->l_out_bytes_per_site = 32 * (int64_t)(h->n_int - h->n_roots) /* inside records of internal children */ + 4 * (int64_t)N +
                              40 * (out_loads + out_stores) + 8 * (int64_t)h->n_roots + (40 * (int64_t)n) / 32;
    // device copies (capacity n_int for every array, so that a new topology over the same counts reuses them)
    const size_t cap = (size_t)std::max(h->n_int, 1);
    if (!h->d_lin_steps &&
        (dalloc(h, &h->d_lin_steps, cap) || dalloc(h, &h->d_lout_steps, cap) || dalloc(h, &h->d_lin_clades, cap) || dalloc(h, &h->d_lout_clades, cap)))
        return 1;
    CK(cudaMemcpy(h->d_lin_steps, in_steps.data(), sizeof(int4) * in_steps.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_lout_steps, out_steps.data(), sizeof(int4) * out_steps.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_lin_clades, in_clades.data(), sizeof(int4) * in_clades.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_lout_clades, out_clades.data(), sizeof(int4) * out_clades.size(), cudaMemcpyHostToDevice));
    // ---- launch groups: consecutive tiers per launch
    h->lin_groups.clear(); h->lout_groups.clear();
    {
        const int nt0 = (int)S.tier_clades[0];
        if (h->df_mode == 0 || n_tiers == 1) {
            for (int t = 0; t < n_tiers; t++) {
                h->lin_groups.push_back({S.in_phase_ptr[t], S.in_phase_ptr[t + 1], 0, 0, 0});
                h->lout_groups.push_back({S.out_phase_ptr[t], S.out_phase_ptr[t + 1], 0, 0, 0});
            }
        } else if (h->df_mode == 1) {
            h->lin_groups.push_back({0, nt0, 0, 0, 0});
            h->lin_groups.push_back({nt0, ncl, 0, ncl - nt0, 1});
            h->lout_groups.push_back({0, ncl - nt0, ncl - nt0, 0, 1});
            h->lout_groups.push_back({ncl - nt0, ncl, 0, 0, 0});
        } else {
            h->lin_groups.push_back({0, ncl, nt0, ncl - nt0, 1});
            h->lout_groups.push_back({0, ncl, ncl - nt0, nt0, 1});
        }
    }
    if (h->df_mode) {
        if (!h->d_df_dep_ptr_in) {
            if (dalloc(h, &h->d_df_dep_ptr_in, cap + 1) || dalloc(h, &h->d_df_dep_ptr_out, cap + 1) || dalloc(h, &h->d_df_tickets, 64) ||
                dalloc(h, &h->d_df_err, 1))
                return 1;
            CK(cudaMemset(h->d_df_tickets, 0, 64 * sizeof(unsigned)));
            memset(h->df_ticket_host, 0, sizeof h->df_ticket_host);
        }
        if (S.in_deps.size() > h->df_deps_cap_in) {
            dfree(h->d_df_deps_in);
            h->df_deps_cap_in = S.in_deps.size() + S.in_deps.size() / 4 + 64;
            if (dalloc(h, &h->d_df_deps_in, h->df_deps_cap_in)) return 1;
        }
        if (S.out_deps.size() > h->df_deps_cap_out) {
            dfree(h->d_df_deps_out);
            h->df_deps_cap_out = S.out_deps.size() + S.out_deps.size() / 4 + 64;
            if (dalloc(h, &h->d_df_deps_out, h->df_deps_cap_out)) return 1;
        }
        CK(cudaMemcpy(h->d_df_dep_ptr_in, S.in_dep_ptr.data(), sizeof(int) * S.in_dep_ptr.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->d_df_dep_ptr_out, S.out_dep_ptr.data(), sizeof(int) * S.out_dep_ptr.size(), cudaMemcpyHostToDevice));
        if (!S.in_deps.empty()) CK(cudaMemcpy(h->d_df_deps_in, S.in_deps.data(), sizeof(int) * S.in_deps.size(), cudaMemcpyHostToDevice));
        if (!S.out_deps.empty()) CK(cudaMemcpy(h->d_df_deps_out, S.out_deps.data(), sizeof(int) * S.out_deps.size(), cudaMemcpyHostToDevice));
        // tickets count up from wherever they are, and the marks of an old topology can never equal a future epoch
        CK(cudaMemset(h->d_df_err, 0, sizeof(int)));
    }
    if (env_int("PMM_VERBOSE", 0)) {
        fprintf(stderr, "[pmm] lumped schedule: clade capacity %d (top %d), %d clades in %d tiers:", C0, Ctop, ncl, n_tiers);
        for (int t = 0; t < n_tiers; t++) fprintf(stderr, " %zu", S.tier_clades[t]);
        fprintf(stderr, "; %lld inside loads, %lld outside loads, %lld outside stores; dataflow mode %d (%zu + %zu waits)\n", (long long)in_loads,
                (long long)out_loads, (long long)out_stores, h->df_mode, S.in_deps.size(), S.out_deps.size());
    }
    return 0;
}

// records and partial sums of the lumped engine; sites are processed in chunks of lKc if they do not fit
int alloc_lumped(pmm_handle *h) {
    if (!h->ev_fork) {
        int lo_pri = 0, hi_pri = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
        for (int j = 0; j < 8; j++) CK(cudaStreamCreateWithPriority(&h->aux[j], cudaStreamNonBlocking, (j & 1) ? hi_pri : lo_pri));
        CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        for (auto &row : h->ev_s) for (auto &e : row) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->slices = std::max(1, std::min(4, env_int("PMM_SLICES", 2)));
        h->stagger = env_int("PMM_STAGGER", 0);
        h->pdl = env_int("PMM_PDL", 1);
    }
    const int n = h->n_int + h->n_leaves, K = h->K;     // h->n is set only once pmm_set_tree has succeeded
    size_t budget = h->budget;
    if (budget == 0) {
        size_t fr = 0, tot = 0;
        cudaError_t e = cudaMemGetInfo(&fr, &tot);
        if (e != cudaSuccess) return fail(h, "cudaMemGetInfo failed: %s", cudaGetErrorString(e));
        budget = (size_t)(0.85 * (double)fr);
    }
    const double per_site = 72.0 * h->n_int + 16.0 * h->n_roots + (40.0 * n + 8.0 * h->n_int) / 32.0;
    int Kc = K;
    const int forced = env_int("PMM_CHUNKS", 0);
    if (forced > 0) Kc = std::max(1, (K + forced - 1) / forced);
    else if (per_site * K > (double)budget) Kc = std::min(K, std::max(32, (int)((double)budget / per_site) / 32 * 32));
    // the kernels index records with 32 bits
    Kc = (int)std::min<int64_t>(Kc, std::max<int64_t>(1, ((int64_t)1 << 31) / std::max(h->n_int, 1) - 32));
    h->lKc = Kc; h->lG = (Kc + 31) / 32;
    const size_t rows = (size_t)h->n_int * (size_t)((Kc + 31) & ~31);      // record rows are padded to whole warps
    if (dalloc(h, &h->d_LX, 2 * rows) || dalloc(h, &h->d_LO, rows) || dalloc(h, &h->d_LOI, rows) ||
        dalloc(h, &h->d_LinvL, (size_t)h->n_roots * Kc) || dalloc(h, &h->d_Lroot_llh, (size_t)h->n_roots * Kc) ||
        dalloc(h, &h->d_LSpart, (size_t)h->lG * n * 5) || dalloc(h, &h->d_LRpart, (size_t)h->lG * std::max(h->n_int, 1)))
        return 1;
    // completion marks of the dataflow launches: (pass, clade position, site block of the chunk)
    h->df_stride = (Kc + LT - 1) / LT;
    const size_t nflags = (size_t)2 * std::max(h->n_int, 1) * h->df_stride;
    if (dalloc(h, &h->d_df_flags, nflags)) return 1;
    CK(cudaMemset(h->d_df_flags, 0, nflags * sizeof(int)));
    return 0;
}

LParams make_lparams(pmm_handle *h, int ci, bool with_post) {
    LParams P{};
    P.site0 = ci * h->lKc; P.K = std::min(h->lKc, h->K - P.site0); P.Kp = (h->lKc + 31) & ~31; P.code_stride = h->K;
    P.n_leaves = h->n_leaves; P.n_nodes = h->n; P.n_clades = h->n_clades;
    P.codes = h->d_codes; P.site_ptr = h->d_site_ptr; P.prior = h->d_prior; P.qsum = h->d_qsum;
    P.params = h->d_params; P.bc = h->d_bc;
    P.X = h->d_LX; P.O = h->d_LO; P.OI = h->d_LOI;
    P.invL = h->d_LinvL; P.root_site_llh = h->d_Lroot_llh; P.Spart = h->d_LSpart; P.Rpart = h->d_LRpart;
    P.post0 = with_post ? h->d_post0 : nullptr; P.post1 = with_post ? h->d_post1 : nullptr;
    P.K_total = h->K; P.store_all = (env_int("PMM_STORE_ALL", 0) || h->force_store_all) ? 1 : 0;
    P.sblk0 = 0;
    return P;
}

int run_lumped(pmm_handle *h, bool outside, bool with_post) {
    const int n = h->n;
    if (h->peer_lost)
        return fail(h, "this handle exchanged statistics with peers until pmm_set_tree changed the node count and freed the exchange "
                       "buffer: call pmm_peer_export / pmm_peer_setup again on every rank (or pmm_peer_close to go on alone)");
    CK(cudaEventRecord(h->ev[0], h->stream));
    pmm_branch_consts<<<(n + 255) / 256, 256, 0, h->stream>>>(h->d_d, h->d_params, n, h->d_bc);
    h->launches++;
    // with peers the rank's own sums go to the `loc` part of the exchange block; the last chunk's finishing kernel is
    // then pmm_finish_peer, which also reduces over the ranks into every rank's statistics
    const bool peers = h->peer_world > 1;
    double *stats_out = peers ? h->d_xloc : h->d_stats;
    double *d_llh = stats_out + (size_t)5 * n + 1;
    double *d_Rt = stats_out + (size_t)5 * n;
    const int nchunks = (h->K + h->lKc - 1) / h->lKc;
    float acc_in = 0, acc_out = 0;
    for (int ci = 0; ci < nchunks; ci++) {
        LParams P = make_lparams(h, ci, with_post);
        const unsigned gy = (unsigned)((P.K + LT - 1) / LT);
        const int G = (P.K + 31) / 32;
        P.sblk0 = 0;
        CK(cudaEventRecord(h->ev[1], h->stream));
        // pdl: programmatic dependent launch -- the kernel may start while its predecessor in the stream drains
        // (it stages its clade, then waits for the predecessor's records); only between two lumped launches
        const bool use_pdl = h->pdl != 0;
        auto launch_k = [&](auto kern, cudaStream_t st, dim3 grid, bool pdl, LParams Q, const int4 *cl, const int4 *steps) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = grid; cfg.blockDim = dim3(LT); cfg.dynamicSmemBytes = 0; cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = (pdl && use_pdl) ? 1 : 0;
            cudaLaunchKernelEx(&cfg, kern, Q, cl, steps);
            h->launches++;
        };
        // a dataflow group: one 1-D launch over (clades of the group) x (site blocks), with its own ticket counter
        int cur_slice = 0;
        auto df_setup = [&](LParams &Q, const pmm_handle::LGroup &g, int pass, size_t p, unsigned gyq) {
            const int slot = pass * 32 + cur_slice * 8 + (int)std::min<size_t>(p, 7);
            Q.df_ticket = h->d_df_tickets + slot; Q.df_base = h->df_ticket_host[slot];
            h->df_ticket_host[slot] += (unsigned)(g.c1 - g.c0) * gyq;
            Q.df_flags = h->d_df_flags + (size_t)pass * std::max(h->n_int, 1) * h->df_stride;
            Q.df_dep_ptr = pass ? h->d_df_dep_ptr_out : h->d_df_dep_ptr_in; Q.df_deps = pass ? h->d_df_deps_out : h->d_df_deps_in;
            Q.df_lo = g.c0; Q.df_n1 = g.n1; Q.df_n2 = g.n2; Q.df_nblk = (int)gyq; Q.df_stride = h->df_stride;
            Q.df_epoch = ++h->df_epoch; Q.df_err = h->d_df_err;
            if (h->df_epoch == 0x7fffffff) h->df_epoch = 0;     // (marks are compared for equality; a wrap after 2^31 launches is harmless)
        };
        auto launch_in = [&](cudaStream_t st, size_t p, LParams Q, unsigned gyq, bool pdl) {
            const pmm_handle::LGroup &g = h->lin_groups[p];
            if (g.c1 <= g.c0) return;
            const int4 *cl = (const int4 *)(h->d_lin_clades + g.c0), *steps = (const int4 *)h->d_lin_steps;
            if (g.df) { df_setup(Q, g, 0, p, gyq); launch_k(pmm_lumped_inside<true>, st, dim3((unsigned)(g.c1 - g.c0) * gyq), pdl, Q, cl, steps); }
            else launch_k(pmm_lumped_inside<false>, st, dim3((unsigned)(g.c1 - g.c0), gyq), pdl, Q, cl, steps);
        };
        auto launch_out = [&](cudaStream_t st, size_t p, LParams Q, unsigned gyq, bool pdl) {
            const pmm_handle::LGroup &g = h->lout_groups[p];
            if (g.c1 <= g.c0) return;
            const int4 *cl = (const int4 *)(h->d_lout_clades + g.c0), *steps = (const int4 *)h->d_lout_steps;
            const dim3 grid = g.df ? dim3((unsigned)(g.c1 - g.c0) * gyq) : dim3((unsigned)(g.c1 - g.c0), gyq);
            if (g.df) df_setup(Q, g, 1, p, gyq);
            if (g.df) {
                if (with_post) launch_k(pmm_lumped_outside<true, true>, st, grid, pdl, Q, cl, steps);
                else launch_k(pmm_lumped_outside<false, true>, st, grid, pdl, Q, cl, steps);
            } else {
                if (with_post) launch_k(pmm_lumped_outside<true, false>, st, grid, pdl, Q, cl, steps);
                else launch_k(pmm_lumped_outside<false, false>, st, grid, pdl, Q, cl, steps);
            }
        };
        const size_t n_in = h->lin_groups.size(), n_out = h->lout_groups.size();
        // a dataflow launch over all tiers interleaves the site blocks itself: no site slices
        const int nslice = (h->df_mode == 2) ? 1 : std::max(1, std::min({h->slices, (int)gy, 4}));
        h->last_sliced = nslice > 1;
        if (nslice == 1) {
            for (size_t p = 0; p < n_in; p++) launch_in(h->stream, p, P, gy, p > 0);      // the first follows pmm_branch_consts
            CK(cudaEventRecord(h->ev[2], h->stream));
            if (outside) for (size_t p = 0; p < n_out; p++) launch_out(h->stream, p, P, gy, p > 0);   // the first follows the timing event
        } else {
            // per slice: tier 0 (inside) on a normal-priority stream, every top-tier launch of both passes on a
            // high-priority stream, tier 0 (outside) on the normal stream again
            CK(cudaEventRecord(h->ev_fork, h->stream));
            for (int j = 0; j < nslice; j++) {
                cudaStream_t sn = h->aux[2 * j], sh = h->aux[2 * j + 1];
                LParams Q = P;
                cur_slice = j;
                Q.sblk0 = (int)((uint64_t)gy * j / nslice);
                const unsigned gyq = (unsigned)((uint64_t)gy * (j + 1) / nslice) - (unsigned)Q.sblk0;
                CK(cudaStreamWaitEvent(sn, h->ev_fork, 0));
                // stagger: a slice's big launch starts when the previous slice's has finished, i.e. while that
                // slice walks its latency-bound top tiers
                if (j > 0 && h->stagger) CK(cudaStreamWaitEvent(sn, h->ev_s[j - 1][0], 0));
                launch_in(sn, 0, Q, gyq, false);
                CK(cudaEventRecord(h->ev_s[j][0], sn));
                CK(cudaStreamWaitEvent(sh, h->ev_s[j][0], 0));
                for (size_t p = 1; p < n_in; p++) launch_in(sh, p, Q, gyq, p > 1);
                if (outside) for (size_t p = 0; p + 1 < n_out; p++) launch_out(sh, p, Q, gyq, n_in > 1 || p > 0);
                CK(cudaEventRecord(h->ev_s[j][1], sh));
                CK(cudaStreamWaitEvent(sn, h->ev_s[j][1], 0));
                if (outside) launch_out(sn, n_out - 1, Q, gyq, false);
                CK(cudaEventRecord(h->ev_s[j][2], sn));
            }
            for (int j = 0; j < nslice; j++) CK(cudaStreamWaitEvent(h->stream, h->ev_s[j][2], 0));
            CK(cudaEventRecord(h->ev[2], h->stream));
        }
        if (peers && ci == nchunks - 1) {
            // the same sums, then the exchange over peer memory: ONE cooperative launch (at most one CTA per SM)
            const double *sp = h->d_LSpart, *rl = h->d_Lroot_llh, *rp = h->d_LRpart;
            int gg = G, n5 = 5 * n, acc = ci > 0, nr = h->n_roots, kk = P.K, ncl = h->n_clades, wc = outside ? 1 : 0;
            double *sl = h->d_site_llh + P.site0, *rn = h->d_Rnode;
            const int *de = h->df_mode ? h->d_df_err : nullptr;
            PeerArgs A = h->peer;
            A.epoch = ++h->peer_epoch;
            void *args[] = {&sp, &gg, &n5, &acc, &rl, &nr, &kk, &sl, &rp, &ncl, &rn, &wc, &de, &A};
            const int nblk = std::max(1, std::min(h->n_sms, (n5 + 2 * PEER_THREADS - 1) / (2 * PEER_THREADS)));
            CK(cudaLaunchCooperativeKernel((void *)pmm_finish_peer, dim3((unsigned)nblk), dim3(PEER_THREADS), args, 0, h->stream));
            h->launches += 1;
        } else {   // per-site and total log-likelihood, counts summed over the site groups, R~: one launch
            const unsigned nb = outside ? (unsigned)((5 * n + 1023) / 1024) : 0u;
            pmm_lumped_finish<<<nb + 1, 1024, 0, h->stream>>>(h->d_LSpart, G, 5 * n, stats_out, ci > 0, h->d_Lroot_llh, h->n_roots, P.K,
                                                             h->d_site_llh + P.site0, d_llh, h->d_LRpart, h->n_clades, h->d_Rnode, d_Rt,
                                                             outside ? 1 : 0, h->df_mode ? h->d_df_err : nullptr);
            h->launches += 1;
        }
        CK(cudaEventRecord(h->ev[3], h->stream));
        if (nchunks > 1) {
            CK(cudaEventSynchronize(h->ev[3]));
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, h->ev[1], h->ev[2]);
            cudaEventElapsedTime(&b, h->ev[2], h->ev[3]);
            acc_in += a; acc_out += b;
        }
    }
    CK(cudaGetLastError());
    h->t_inside = acc_in; h->t_outside = acc_out;
    h->timing_valid = false;
    return 0;
}

// full-state rows (pmm_kernels.cuh): allocated on first use -- the NNI batch scorer and PMM_ENGINE=fullstate
int ensure_fullstate(pmm_handle *h) {
    if (h->full_ready) return 0;
    const int n = h->n_int + h->n_leaves;      // h->n is set only once pmm_set_tree has succeeded
    if (plan_chunks(h)) return 1;
    if (configure_kernels(h)) return 1;
    const size_t rows = (size_t)h->n_int;
    if (dalloc(h, &h->d_X, rows * h->maxWp) || dalloc(h, &h->d_O, rows * h->maxWp) || dalloc(h, &h->d_iexp, rows * h->maxKp) ||
        dalloc(h, &h->d_dexp, rows * h->maxKp) || dalloc(h, &h->d_eX, rows * h->maxKp) || dalloc(h, &h->d_eO, rows * h->maxKp) ||
        dalloc(h, &h->d_sqX, rows * h->maxKp) || dalloc(h, &h->d_rootInvL, (size_t)h->n_roots * h->maxKp) ||
        dalloc(h, &h->d_root_site_llh, (size_t)h->n_roots * h->maxK) ||
        dalloc(h, &h->d_Spart, (size_t)h->max_tiles * n * 5) || dalloc(h, &h->d_Rpart, (size_t)h->max_tiles * h->n_int))
        return 1;
    h->full_ready = true;
    return 0;
}

int check_ready(pmm_handle *h) {
    if (!h) return 1;
    if (h->n == 0) return fail(h, "pmm_set_tree has not been called");
    CK(cudaSetDevice(h->device));
    return 0;
}

// After a synchronisation of the handle's stream: did a wait of the exchange over peer memory (pmm_peer.cuh) or of a
// dataflow launch give up?  The statistics are then incomplete (a late peer's push may even have replaced the NaN the
// kernel left in the likelihood slot), so every entry point that hands results to the host asks here first.
int waits_ok(pmm_handle *h) {
    int e[2] = {0, 0};
    if (h->peer_world <= 1 && !h->df_mode) return 0;
    // (on the handle's own stream, which the caller has just synchronised: no implicit join with other streams of the process)
    if (h->peer_world > 1) CK(cudaMemcpyAsync(&e[0], h->d_peer_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (h->df_mode) CK(cudaMemcpyAsync(&e[1], h->d_df_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (e[0])
        return fail(h, "the exchange over peer memory gave up waiting for a peer (rank %d of %d; PMM_PEER_TIMEOUT_MS, default 30 s): a rank "
                       "that did not make the same call, or kernels of several ranks sharing one GPU's work queues; set up the peers again",
                    h->peer_rank, h->peer_world);
    if (e[1]) return fail(h, "a dataflow launch gave up waiting for a producer clade (PMM_DF): please report");
    return 0;
}

// free-branch mask of the M-step: NULL means "every branch is free" (never "whatever the last call left"); the
// handle remembers what the device holds, so repeating a mask costs a memcmp and no copy
int set_free_mask(pmm_handle *h, const uint8_t *free_mask) {
    const size_t n = (size_t)h->n;
    if (!free_mask) {
        if (!h->free_host.empty()) {
            CK(cudaMemsetAsync(h->d_free, 1, n, h->stream));
            h->free_host.clear();
        }
        return 0;
    }
    if (h->free_host.size() == n && memcmp(h->free_host.data(), free_mask, n) == 0) return 0;
    // pageable source: the copy is staged before the call returns, the caller's buffer is free afterwards
    h->free_host.assign(free_mask, free_mask + n);
    CK(cudaMemcpyAsync(h->d_free, h->free_host.data(), n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// inside (+ optionally outside/counts) over all chunks, parameters already on the device
int run_passes(pmm_handle *h, bool outside, bool with_post, bool force_full = false) {
    if (h->engine == 0 && !force_full) { h->last_full = false; return run_lumped(h, outside, with_post); }
    if (h->peer_world > 1) return fail(h, "the exchange over peer memory belongs to the lumped-state engine (PMM_ENGINE=fullstate: all-reduce the statistics yourself)");
    if (ensure_fullstate(h)) return 1;
    h->last_full = true;
    const int n = h->n;
    CK(cudaEventRecord(h->ev[0], h->stream));
    pmm_branch_consts<<<(n + 255) / 256, 256, 0, h->stream>>>(h->d_d, h->d_params, n, h->d_bc);
    h->launches++;
    double *d_llh = h->d_stats + (size_t)5 * n + 1;
    double *d_Rt = h->d_stats + (size_t)5 * n;
    float acc_in = 0, acc_out = 0;
    for (size_t ci = 0; ci < h->chunks.size(); ci++) {
        const Chunk &c = h->chunks[ci];
        KParams P = make_params(h, c, with_post);
        CK(cudaEventRecord(h->ev[1], h->stream));
        if (launch_inside(h, P)) return 1;
        pmm_site_llh<<<(c.K + 255) / 256, 256, 0, h->stream>>>(h->d_root_site_llh, h->n_roots, c.K, h->d_site_llh + c.site0);
        pmm_reduce_sum<<<1, 1024, 0, h->stream>>>(h->d_site_llh + c.site0, c.K, d_llh, ci > 0);
        h->launches += 2;
        CK(cudaEventRecord(h->ev[2], h->stream));
        if (outside) {
            if (launch_outside(h, P)) return 1;
            pmm_reduce_S<<<(5 * n + 255) / 256, 256, 0, h->stream>>>(h->d_Spart, c.ntiles, 5 * n, h->d_stats, ci > 0);
            pmm_reduce_S<<<(h->n_int + 255) / 256, 256, 0, h->stream>>>(h->d_Rpart, c.ntiles, h->n_int, h->d_Rnode, 0);
            pmm_reduce_sum<<<1, 1024, 0, h->stream>>>(h->d_Rnode, h->n_int, d_Rt, ci > 0);
            h->launches += 3;
        }
        CK(cudaEventRecord(h->ev[3], h->stream));
        if (h->chunks.size() > 1) {
            // several chunks: accumulate phase times chunk by chunk (costs a sync per chunk)
            CK(cudaEventSynchronize(h->ev[3]));
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, h->ev[1], h->ev[2]);
            cudaEventElapsedTime(&b, h->ev[2], h->ev[3]);
            acc_in += a; acc_out += b;
        }
    }
    CK(cudaGetLastError());
    h->t_inside = acc_in; h->t_outside = acc_out;
    h->timing_valid = false;     // resolved lazily in pmm_last_timing after a sync
    return 0;
}

}  // namespace

// =============================================================== C ABI
extern "C" {

int pmm_abi_version(void) { return 2; }

const char *pmm_last_error(const pmm_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int pmm_create(int32_t n_leaves, int32_t n_sites, const int32_t *site_ptr, const double *prior, const int32_t *codes,
               int32_t device, size_t mem_budget_bytes, pmm_handle **out) {
    if (!out) return fail(nullptr, "pmm_create: out is NULL");
    *out = nullptr;
    if (n_leaves < 2 || n_sites < 1 || !site_ptr || !codes) return fail(nullptr, "pmm_create: need >=2 leaves, >=1 site, non-NULL arrays");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, "pmm_create: no usable CUDA device (%s); this engine has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, "pmm_create: device %d out of range (%d visible)", device, ndev);
    pmm_handle *h = new pmm_handle();
    h->device = device;
    auto bail = [&](int) { g_create_error = h->err; pmm_destroy(h); return 1; };
    if (cudaSetDevice(device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return bail(1); }
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { h->err = "cudaStreamCreate failed"; return bail(1); }
    h->own_stream = true;
    for (auto &ev : h->ev) cudaEventCreate(&ev);
    h->n_leaves = n_leaves; h->K = n_sites; h->budget = mem_budget_bytes;
    h->site_ptr.assign(site_ptr, site_ptr + n_sites + 1);
    if (h->site_ptr[0] != 0) { h->err = "site_ptr[0] must be 0"; return bail(1); }
    int maxlen = 0;
    for (int s = 0; s < n_sites; s++) {
        int A = site_ptr[s + 1] - site_ptr[s];
        if (A < 0) { h->err = "site_ptr must be non-decreasing"; return bail(1); }
        maxlen = std::max(maxlen, A + 2);
    }
    const int nnz = site_ptr[n_sites];
    h->prior.assign(prior, prior + nnz);
    for (int i = 0; i < nnz; i++)
        if (!(h->prior[i] >= 0.0)) { h->err = "prior entries must be non-negative numbers"; return bail(1); }
    h->W = (int64_t)nnz + 2 * (int64_t)n_sites;
    h->TS = env_int("PMM_TILE_SLOTS", 1792);
    h->TS = std::max((h->TS + 1) & ~1, (maxlen + 1) & ~1);
    if (h->TS > 8192) { h->err = "an alphabet is too large for one shared-memory tile (limit ~8000 states)"; return bail(1); }
    // validate codes against the alphabets, then upload
    for (int64_t r = 0; r < n_leaves; r++)
        for (int s = 0; s < n_sites; s++) {
            int c = codes[r * n_sites + s];
            if (c < -2 || c > site_ptr[s + 1] - site_ptr[s]) {
                char b[160];
                snprintf(b, sizeof b, "code %d of leaf %lld at site %d is outside the site's alphabet (%d states)", c, (long long)r, s,
                         site_ptr[s + 1] - site_ptr[s]);
                h->err = b;
                return bail(1);
            }
        }
    {
        const char *eng = getenv("PMM_ENGINE");
        h->engine = (eng && std::string(eng) == "fullstate") ? 1 : 0;
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->n_sms = prop.multiProcessorCount;
    }
    std::vector<double> qsum((size_t)n_sites, 0.0);
    for (int s = 0; s < n_sites; s++) {
        double acc = 0.0;
        for (int a = site_ptr[s]; a < site_ptr[s + 1]; a++) acc += prior[a];
        qsum[s] = acc;
    }
    if (dalloc(h, &h->d_codes, (size_t)n_leaves * n_sites) || dalloc(h, &h->d_params, 2) || dalloc(h, &h->d_site_ptr, (size_t)n_sites + 1) ||
        dalloc(h, &h->d_prior, (size_t)nnz) || dalloc(h, &h->d_qsum, (size_t)n_sites))
        return bail(1);
    if (cudaMemcpy(h->d_site_ptr, site_ptr, sizeof(int) * ((size_t)n_sites + 1), cudaMemcpyHostToDevice) != cudaSuccess ||
        (nnz > 0 && cudaMemcpy(h->d_prior, prior, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice) != cudaSuccess) ||
        cudaMemcpy(h->d_qsum, qsum.data(), sizeof(double) * (size_t)n_sites, cudaMemcpyHostToDevice) != cudaSuccess) {
        h->err = "uploading the prior failed";
        return bail(1);
    }
    if (cudaMemcpy(h->d_codes, codes, sizeof(int) * (size_t)n_leaves * n_sites, cudaMemcpyHostToDevice) != cudaSuccess) {
        h->err = "uploading codes failed";
        return bail(1);
    }
    *out = h;
    return 0;
}

void pmm_destroy(pmm_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_tree_state(h);
    dfree(h->d_codes); dfree(h->d_params); dfree(h->d_site_ptr); dfree(h->d_prior); dfree(h->d_qsum);
    dfree(h->d_nni_c); dfree(h->d_nni_s); dfree(h->d_nni_u); dfree(h->d_nni_rsum);
    for (auto &ev : h->ev) if (ev) cudaEventDestroy(ev);
    if (h->pin_up) cudaFreeHost(h->pin_up);
    if (h->pin_down) cudaFreeHost(h->pin_down);
    if (h->ev_up) cudaEventDestroy(h->ev_up);
    for (auto &st : h->aux) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    for (auto &row : h->ev_s) for (auto &e : row) if (e) cudaEventDestroy(e);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int pmm_set_stream(pmm_handle *h, void *cuda_stream) {
    if (!h) return 1;
    CK(cudaSetDevice(h->device));
    if (h->stream) CK(cudaStreamSynchronize(h->stream));
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)cuda_stream;
    h->own_stream = false;
    return 0;
}

int pmm_set_tree(pmm_handle *h, int32_t n_nodes, const int32_t *parent, const int32_t *left, const int32_t *right) {
    if (!h) return 1;
    CK(cudaSetDevice(h->device));
    const int N = h->n_leaves, n = n_nodes;
    if (n <= N || n > 2 * N - 1 || !parent || !left || !right)
        return fail(h, "pmm_set_tree: a binary forest over %d leaf nodes has between %d and %d nodes (got %d)", N, N + 1, 2 * N - 1, n);
    // validate: leaves first, binary, consistent parent links
    int n_roots = 0;
    for (int v = 0; v < n; v++) {
        bool leaf = left[v] < 0;
        if (leaf != (right[v] < 0)) return fail(h, "node %d has exactly one child; trees must be binary", v);
        if (leaf != (v < N)) return fail(h, "node ids must list the %d leaves first (node %d violates this)", N, v);
        if (!leaf) {
            if (left[v] >= n || right[v] >= n || left[v] == right[v]) return fail(h, "node %d has invalid children", v);
            if (parent[left[v]] != v || parent[right[v]] != v) return fail(h, "parent/child arrays disagree at node %d", v);
        }
        if (parent[v] < 0) { n_roots++; if (leaf) return fail(h, "a single-leaf tree has no likelihood recursion"); }
        else if (parent[v] >= n) return fail(h, "node %d has an invalid parent", v);
    }
    if (h->stream) CK(cudaStreamSynchronize(h->stream));
    // a new topology over the same node / root counts (topology search: Topology_search.py:289-299 builds
    // a fresh solver per candidate) keeps every device buffer and the chunk plan: only the tree arrays and
    // the level schedules are replaced
    const bool reuse = (h->n == n && h->n_roots == n_roots && h->d_LX != nullptr);
    (void)cudaGetLastError();      // a stale error of another library in this process is not ours
    if (!reuse) {
        free_tree_state(h);
        h->device_bytes = sizeof(int) * (size_t)h->n_leaves * h->K + 16 + sizeof(int) * ((size_t)h->K + 1) +
                          sizeof(double) * (h->prior.size() + (size_t)h->K);   // what pmm_create holds
        h->n = 0;                  // "no tree" until this call has succeeded
    }
    // depth (root = 0) and height (leaf = 0) without recursion
    std::vector<int> depth(n, -1), height(n, 0), order;
    order.reserve(n);
    std::vector<int> roots;
    for (int v = 0; v < n; v++) if (parent[v] < 0) { roots.push_back(v); depth[v] = 0; order.push_back(v); }
    for (size_t i = 0; i < order.size(); i++) {
        int v = order[i];
        if (left[v] >= 0) {
            for (int c : {left[v], right[v]}) {
                if (depth[c] >= 0) return fail(h, "node %d is reachable twice; not a forest", c);
                depth[c] = depth[v] + 1;
                order.push_back(c);
            }
        }
    }
    if ((int)order.size() != n) return fail(h, "tree arrays contain a cycle or unreachable nodes");
    h->n_int = n - N; h->n_roots = n_roots;
    h->parent.assign(parent, parent + n); h->left.assign(left, left + n); h->right.assign(right, right + n);
    std::vector<int> root_index(n, 0);
    for (size_t i = 0; i < roots.size(); i++) root_index[roots[i]] = (int)i;
    for (int v : order) if (parent[v] >= 0) root_index[v] = root_index[parent[v]];
    for (int i = n - 1; i >= 0; i--) {
        int v = order[i];
        if (left[v] >= 0) height[v] = 1 + std::max(height[left[v]], height[right[v]]);
    }
    int H = 0, D = 0;
    for (int v = N; v < n; v++) { H = std::max(H, height[v]); D = std::max(D, depth[v]); }
    // inside schedule: internal nodes by height 1..H; outside: internal nodes by depth 0..D.
    // Within a level the items are sorted by node class (both children internal / one / none) and
    // the children are ordered internal-first: the recursions are symmetric in the two children.
    std::vector<std::vector<int>> inl(H), outl(D + 1);
    for (int v = N; v < n; v++) { inl[height[v] - 1].push_back(v); outl[depth[v]].push_back(v); }
    auto build_items = [&](std::vector<std::vector<int>> &levels, std::vector<int4> &items, std::vector<int> &ptr, std::vector<int> &cls) {
        ptr.assign(1, 0); cls.clear(); items.clear();
        for (auto &L : levels) {
            int cnt[3] = {0, 0, 0};
            for (int pass = 0; pass < 3; pass++)
                for (int u : L) {
                    int a = left[u], b = right[u];
                    if (a < N && b >= N) std::swap(a, b);
                    const int c = (a >= N) ? ((b >= N) ? 0 : 1) : 2;
                    if (c != pass) continue;
                    items.push_back(make_int4(u, a, b, root_index[u]));
                    cnt[c]++;
                }
            cls.insert(cls.end(), cnt, cnt + 3);
            ptr.push_back((int)items.size());
        }
    };
    std::vector<int4> in_items, out_items;
    build_items(inl, in_items, h->in_ptr, h->in_cls);
    build_items(outl, out_items, h->out_ptr, h->out_cls);

    if (!reuse &&
        (dalloc(h, &h->d_parent, n) || dalloc(h, &h->d_left, n) || dalloc(h, &h->d_right, n) || dalloc(h, &h->d_root_index, n) ||
         dalloc(h, &h->d_roots, roots.size()) || dalloc(h, &h->d_in_items, in_items.size()) || dalloc(h, &h->d_out_items, out_items.size())))
        return 1;
    CK(cudaMemcpy(h->d_parent, parent, sizeof(int) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_left, left, sizeof(int) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_right, right, sizeof(int) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_root_index, root_index.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_roots, roots.data(), sizeof(int) * roots.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_in_items, in_items.data(), sizeof(int4) * in_items.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_out_items, out_items.data(), sizeof(int4) * out_items.size(), cudaMemcpyHostToDevice));

    if (build_lumped_schedule(h, order, root_index)) return 1;
    if (reuse) {
        // the level phases of the device ultrametric solver belong to the old topology, and so do its warm-start heights
        dfree(h->um_up_ptr); dfree(h->um_dn_ptr); dfree(h->um_up_ph); dfree(h->um_dn_ph);
        h->um_ready = false; h->um_warm = false; h->um_kind_host.clear(); h->um_fixed_host.clear();
        // node ids mean something else under the new topology: a mask of the old one must not survive
        CK(cudaMemsetAsync(h->d_free, 1, n, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        h->free_host.clear();
        return 0;
    }
    // small per-node buffers first, then the records against what is left
    h->xL = ((size_t)5 * n + 2 + 1) & ~(size_t)1;
    if (dalloc(h, &h->d_xb, 2 * h->xL + PEER_MAX) || dalloc(h, &h->d_peer_bar, 2) || dalloc(h, &h->d_peer_err, 1)) return 1;
    CK(cudaMemset(h->d_xb, 0, sizeof(double) * (2 * h->xL + PEER_MAX)));
    CK(cudaMemset(h->d_peer_bar, 0, 2 * sizeof(unsigned)));
    CK(cudaMemset(h->d_peer_err, 0, sizeof(int)));
    h->d_xloc = h->d_xb; h->d_stats = h->d_xb + h->xL; h->d_xflags = reinterpret_cast<int *>(h->d_xb + 2 * h->xL);
    if (dalloc(h, &h->d_bc, (size_t)4 * n) || dalloc(h, &h->d_d, n) ||
        dalloc(h, &h->d_Snorm, (size_t)5 * n) || dalloc(h, &h->d_free, n) || dalloc(h, &h->d_nu_state, 1) || dalloc(h, &h->d_nu_partial, NU_BLOCKS * 8) || dalloc(h, &h->d_nu_bar, 2) || dalloc(h, &h->d_Rnode, h->n_int) ||
        dalloc(h, &h->d_site_llh, h->K))
        return 1;
    if (alloc_lumped(h)) return 1;
    if (h->engine == 1 && ensure_fullstate(h)) return 1;
    CK(cudaMemsetAsync(h->d_free, 1, n, h->stream));
    CK(cudaMemsetAsync(h->d_nu_bar, 0, 2 * sizeof(unsigned), h->stream));    // pmm_mstep_fused leaves its barrier counters at zero
    CK(cudaStreamSynchronize(h->stream));
    h->free_host.clear();
    h->n = n;
    return 0;
}

int pmm_debug_lumped_schedule(int32_t n_leaves, int32_t n_nodes, const int32_t *parent, const int32_t *left, const int32_t *right,
                              int32_t clade_cap, int32_t clade_cap_top, int32_t *in_steps, int32_t *out_steps, int32_t *in_clades,
                              int32_t *out_clades, int32_t *in_phase_ptr, int32_t *out_phase_ptr, int32_t *n_clades, int32_t *n_phases) {
    const int N = n_leaves, n = n_nodes;
    if (N < 2 || n <= N || n > 2 * N - 1 || !parent || !left || !right) return 1;
    std::vector<int> order, root_index(n, 0);
    order.reserve(n);
    int nr = 0;
    for (int v = 0; v < n; v++) if (parent[v] < 0) { root_index[v] = nr++; order.push_back(v); }
    for (size_t i = 0; i < order.size(); i++) {
        const int v = order[i];
        if (left[v] >= 0) for (int c : {left[v], right[v]}) { if (c < 0 || c >= n) return 1; root_index[c] = root_index[v]; order.push_back(c); }
        if ((int)order.size() > n) return 1;
    }
    if ((int)order.size() != n) return 1;
    LumpedSchedule S;
    make_lumped_schedule(N, n, left, right, parent, order, root_index, std::max(1, std::min((int)clade_cap, STEP_CAP - 1)),
                         std::max(1, std::min((int)clade_cap_top, STEP_CAP - 1)), S);
    auto put = [](const std::vector<int4> &v, int32_t *o) { if (o) memcpy(o, v.data(), sizeof(int4) * v.size()); };
    put(S.in_steps, in_steps); put(S.out_steps, out_steps); put(S.in_clades, in_clades); put(S.out_clades, out_clades);
    if (in_phase_ptr) memcpy(in_phase_ptr, S.in_phase_ptr.data(), sizeof(int) * S.in_phase_ptr.size());
    if (out_phase_ptr) memcpy(out_phase_ptr, S.out_phase_ptr.data(), sizeof(int) * S.out_phase_ptr.size());
    if (n_clades) *n_clades = S.n_clades;
    if (n_phases) *n_phases = S.n_tiers;
    return 0;
}

int pmm_debug_lumped_deps(int32_t n_leaves, int32_t n_nodes, const int32_t *parent, const int32_t *left, const int32_t *right,
                          int32_t clade_cap, int32_t clade_cap_top, int32_t *in_dep_ptr, int32_t *in_deps, int32_t *out_dep_ptr,
                          int32_t *out_deps) {
    const int N = n_leaves, n = n_nodes;
    if (N < 2 || n <= N || n > 2 * N - 1 || !parent || !left || !right || !in_dep_ptr || !in_deps || !out_dep_ptr || !out_deps) return 1;
    std::vector<int> order, root_index(n, 0);
    order.reserve(n);
    int nr = 0;
    for (int v = 0; v < n; v++) if (parent[v] < 0) { root_index[v] = nr++; order.push_back(v); }
    for (size_t i = 0; i < order.size(); i++) {
        const int v = order[i];
        if (left[v] >= 0) for (int c : {left[v], right[v]}) { if (c < 0 || c >= n) return 1; root_index[c] = root_index[v]; order.push_back(c); }
        if ((int)order.size() > n) return 1;
    }
    if ((int)order.size() != n) return 1;
    LumpedSchedule S;
    make_lumped_schedule(N, n, left, right, parent, order, root_index, std::max(1, std::min((int)clade_cap, STEP_CAP - 1)),
                         std::max(1, std::min((int)clade_cap_top, STEP_CAP - 1)), S);
    memcpy(in_dep_ptr, S.in_dep_ptr.data(), sizeof(int) * S.in_dep_ptr.size());
    memcpy(out_dep_ptr, S.out_dep_ptr.data(), sizeof(int) * S.out_dep_ptr.size());
    if (!S.in_deps.empty()) memcpy(in_deps, S.in_deps.data(), sizeof(int) * S.in_deps.size());
    if (!S.out_deps.empty()) memcpy(out_deps, S.out_deps.data(), sizeof(int) * S.out_deps.size());
    return 0;
}

int pmm_set_option(pmm_handle *h, const char *name, int64_t value) {
    if (!h || !name) return 1;
    const std::string k(name);
    if (k == "slices") { h->slices = (int)std::max<int64_t>(1, std::min<int64_t>(4, value)); return 0; }
    return fail(h, "pmm_set_option: unknown option '%s'", name);
}

// ---- exchange over peer memory (pmm_peer.cuh)
static int peer_finish_setup(pmm_handle *h, int rank, int world, double *const *bases) {
    PeerArgs A = {};
    A.world = world; A.rank = rank;
    for (int p = 0; p < world; p++) {
        A.loc[p] = bases[p]; A.red[p] = bases[p] + h->xL; A.flags[p] = reinterpret_cast<int *>(bases[p] + 2 * h->xL);
    }
    A.err = h->d_peer_err; A.bar = h->d_peer_bar;
    {   // how long a rank waits for its peers before it gives up, in SM cycles
        int khz = 1900000;
        (void)cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device);
        A.patience = (long long)std::max(1, env_int("PMM_PEER_TIMEOUT_MS", 30000)) * (long long)khz;
    }
    CK(cudaMemsetAsync(h->d_xflags, 0, sizeof(int) * 2 * PEER_MAX, h->stream));
    CK(cudaMemsetAsync(h->d_peer_err, 0, sizeof(int), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->peer = A; h->peer_world = world; h->peer_rank = rank; h->peer_epoch = 0;
    h->peer_lost = false;
    return 0;
}

int pmm_peer_export(pmm_handle *h, void *handle_out) {
    if (check_ready(h)) return 1;
    if (!handle_out) return fail(h, "pmm_peer_export: handle_out is NULL");
    static_assert(sizeof(cudaIpcMemHandle_t) == PMM_PEER_HANDLE_BYTES, "PMM_PEER_HANDLE_BYTES");
    cudaIpcMemHandle_t mh;
    CK(cudaIpcGetMemHandle(&mh, h->d_xb));
    memcpy(handle_out, &mh, sizeof mh);
    return 0;
}

int pmm_peer_setup(pmm_handle *h, int32_t rank, int32_t world, const void *handles) {
    if (check_ready(h)) return 1;
    if (world < 1 || world > PEER_MAX || rank < 0 || rank >= world || !handles) return fail(h, "pmm_peer_setup: need 1 <= world <= %d, 0 <= rank < world", PEER_MAX);
    if (h->engine != 0) return fail(h, "pmm_peer_setup: the exchange over peer memory belongs to the lumped-state engine");
    if (h->stream) CK(cudaStreamSynchronize(h->stream));
    peer_drop(h);
    h->peer_lost = false;
    if (world == 1) return 0;
    double *bases[PEER_MAX] = {};
    for (int p = 0; p < world; p++) {
        if (p == rank) { bases[p] = h->d_xb; continue; }
        cudaIpcMemHandle_t mh;
        memcpy(&mh, (const char *)handles + (size_t)p * sizeof mh, sizeof mh);
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            peer_drop(h);
            return fail(h, "pmm_peer_setup: cannot map the exchange buffer of rank %d (%s); is there a peer-to-peer path?", p, cudaGetErrorString(e));
        }
        h->peer_mapped[p] = ptr;
        bases[p] = (double *)ptr;
    }
    return peer_finish_setup(h, rank, world, bases);
}

int pmm_peer_setup_local(pmm_handle *h, int32_t rank, int32_t world, pmm_handle *const *peers) {
    if (check_ready(h)) return 1;
    if (world < 1 || world > PEER_MAX || rank < 0 || rank >= world || !peers) return fail(h, "pmm_peer_setup_local: need 1 <= world <= %d, 0 <= rank < world", PEER_MAX);
    if (h->engine != 0) return fail(h, "pmm_peer_setup_local: the exchange over peer memory belongs to the lumped-state engine");
    if (h->stream) CK(cudaStreamSynchronize(h->stream));
    peer_drop(h);
    h->peer_lost = false;
    if (world == 1) return 0;
    double *bases[PEER_MAX] = {};
    for (int p = 0; p < world; p++) {
        const pmm_handle *q = (p == rank) ? h : peers[p];
        if (!q || q->n != h->n || !q->d_xb) return fail(h, "pmm_peer_setup_local: rank %d has no tree of %d nodes", p, h->n);
        if (q->device != h->device) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, h->device, q->device));
            if (!can) return fail(h, "pmm_peer_setup_local: device %d cannot access device %d", h->device, q->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(q->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(h, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            (void)cudaGetLastError();
        }
        bases[p] = q->d_xb;
    }
    return peer_finish_setup(h, rank, world, bases);
}

int pmm_peer_close(pmm_handle *h) {
    if (!h) return 1;
    CK(cudaSetDevice(h->device));
    if (h->stream) CK(cudaStreamSynchronize(h->stream));
    peer_drop(h);
    h->peer_lost = false;
    return 0;
}

int pmm_set_total_sites(pmm_handle *h, int32_t n_sites_total) {
    if (!h) return 1;
    if (n_sites_total < h->K) return fail(h, "total sites (%d) cannot be below this handle's sites (%d)", n_sites_total, h->K);
    h->K_total = n_sites_total;
    return 0;
}

static int ensure_pinned(pmm_handle *h) {
    const size_t need = (size_t)h->n + 4;
    if (h->pin_cap >= need) return 0;
    if (h->pin_up && h->stream) CK(cudaStreamSynchronize(h->stream));      // a copy may still be reading the old buffer
    if (h->pin_up) { cudaFreeHost(h->pin_up); h->pin_up = nullptr; }
    if (h->pin_down) { cudaFreeHost(h->pin_down); h->pin_down = nullptr; }
    h->pin_cap = 0;
    CK(cudaMallocHost((void **)&h->pin_up, sizeof(double) * need));
    CK(cudaMallocHost((void **)&h->pin_down, sizeof(double) * need));
    if (!h->ev_up) CK(cudaEventCreateWithFlags(&h->ev_up, cudaEventDisableTiming));
    h->pin_cap = need;
    return 0;
}

// is `p` page-locked host memory the copy engines can read and write directly?
static bool host_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// direct: the caller's array is page-locked and stays untouched until the stream has been synchronised (the synchronous
// entry points): the DMA reads it in place, no staging copy
static int upload_impl(pmm_handle *h, const double *d, double nu, double phi, bool direct) {
    if (check_ready(h)) return 1;
    if (!d) return fail(h, "pmm_upload_params: d is NULL");
    if (!(nu >= 0.0) || !(phi >= 0.0) || phi > 1.0) return fail(h, "need nu >= 0 and 0 <= phi <= 1 (got nu=%g phi=%g)", nu, phi);
    const bool had = h->pin_cap >= (size_t)h->n + 4;
    if (ensure_pinned(h)) return 1;
    if (had) CK(cudaEventSynchronize(h->ev_up));            // the previous upload has been read out of the staging buffer
    // validate and stage in one pass; the call returns once the staging buffer holds the values (the caller's
    // array is free), the DMA itself is ordered on the handle's stream like every kernel
    const int n = h->n;
    double *st = h->pin_up;
    int bad = 0;                                             // branch-free (vectorised): 200k lengths are staged in ~30 us
    if (direct)
        for (int v = 0; v < n; v++) bad |= !(d[v] >= 0.0 && d[v] <= 1.79769313486231570815e308);
    else
        for (int v = 0; v < n; v++) {
            const double x = d[v];
            bad |= !(x >= 0.0 && x <= 1.79769313486231570815e308);
            st[v] = x;
        }
    if (bad)
        for (int v = 0; v < n; v++)
            if (!(d[v] >= 0.0) || !std::isfinite(d[v])) return fail(h, "branch length of node %d is %g; need finite, non-negative", v, d[v]);
    st[n] = nu; st[n + 1] = phi;
    CK(cudaMemcpyAsync(h->d_d, direct ? d : st, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_params, st + n, 2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(h->ev_up, h->stream));
    return 0;
}

int pmm_upload_params(pmm_handle *h, const double *d, double nu, double phi) { return upload_impl(h, d, nu, phi, false); }

int pmm_download_params(pmm_handle *h, double *d, double *nu, double *phi) {
    if (check_ready(h)) return 1;
    if (ensure_pinned(h)) return 1;
    const int n = h->n;
    double *st = h->pin_down;
    if (d) CK(cudaMemcpyAsync(st, h->d_d, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(st + n, h->d_params, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (waits_ok(h)) return 1;
    if (d) memcpy(d, st, sizeof(double) * n);
    if (nu) *nu = st[n];
    if (phi) *phi = st[n + 1];
    return 0;
}

int pmm_inside_dev(pmm_handle *h) {
    if (check_ready(h)) return 1;
    return run_passes(h, false, false);
}

int pmm_estep_dev(pmm_handle *h) {
    if (check_ready(h)) return 1;
    return run_passes(h, true, false);
}

void *pmm_stats_dev_ptr(pmm_handle *h) { return h ? (void *)h->d_stats : nullptr; }
int64_t pmm_stats_len(const pmm_handle *h) { return h ? (int64_t)5 * h->n + 2 : 0; }

int pmm_read_llh(pmm_handle *h, double *llh) {
    if (check_ready(h)) return 1;
    if (!llh) return fail(h, "pmm_read_llh: llh is NULL");
    CK(cudaMemcpyAsync(llh, h->d_stats + (size_t)5 * h->n + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return waits_ok(h);
}

int pmm_sync(pmm_handle *h) {
    if (!h) return 1;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    return h->n ? waits_ok(h) : 0;
}

int pmm_inside(pmm_handle *h, const double *d, double nu, double phi, double *site_llh, double *llh) {
    if (pmm_upload_params(h, d, nu, phi)) return 1;
    if (run_passes(h, false, false)) return 1;
    if (llh) CK(cudaMemcpyAsync(llh, h->d_stats + (size_t)5 * h->n + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (site_llh) CK(cudaMemcpyAsync(site_llh, h->d_site_llh, sizeof(double) * h->K, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return waits_ok(h);
}

int pmm_estep(pmm_handle *h, const double *d, double nu, double phi, double *S, double *Rtilde, double *llh, double *site_llh) {
    if (pmm_upload_params(h, d, nu, phi)) return 1;
    if (run_passes(h, true, false)) return 1;
    const size_t n5 = (size_t)5 * h->n;
    if (S) CK(cudaMemcpyAsync(S, h->d_stats, sizeof(double) * n5, cudaMemcpyDeviceToHost, h->stream));
    if (Rtilde) CK(cudaMemcpyAsync(Rtilde, h->d_stats + n5, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (llh) CK(cudaMemcpyAsync(llh, h->d_stats + n5 + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (site_llh) CK(cudaMemcpyAsync(site_llh, h->d_site_llh, sizeof(double) * h->K, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return waits_ok(h);
}

int pmm_posteriors(pmm_handle *h, const double *d, double nu, double phi, double *post0, double *post1) {
    if (pmm_upload_params(h, d, nu, phi)) return 1;
    const size_t cnt = (size_t)h->n * h->K;
    if (!h->d_post0) { if (dalloc(h, &h->d_post0, cnt) || dalloc(h, &h->d_post1, cnt)) return 1; }
    if (run_passes(h, true, true)) return 1;
    if (post0) CK(cudaMemcpyAsync(post0, h->d_post0, sizeof(double) * cnt, cudaMemcpyDeviceToHost, h->stream));
    if (post1) CK(cudaMemcpyAsync(post1, h->d_post1, sizeof(double) * cnt, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return waits_ok(h);
}

// the M-step on the device with the mask d_free holds
static int mstep_dev_impl(pmm_handle *h, double R, int32_t n_cells, int32_t optimize_nu, int32_t optimize_phi,
                          double dmin, double dmax, double eps_s, double eps_nu, bool lengths_done = false) {
    const int n = h->n;
    const int K_total = h->K_total > 0 ? h->K_total : h->K;
    if (!env_int("PMM_NU_MULTI_LAUNCH", 0) && !env_int("PMM_MSTEP_SPLIT", 0)) {
        // ONE cooperative launch: per-branch Newton, grid barrier, the nu iteration, phi (pmm_mstep_fused)
        int G = std::max(1, std::min({2 * h->n_sms, NU_BLOCKS, (n + 255) / 256}));
        const double *stats = h->d_stats; const unsigned char *fm = h->d_free; int nn = n, kt = K_total; double *prm = h->d_params, *dd = h->d_d;
        double *sn = h->d_Snorm; int db = lengths_done ? 0 : 1, dn = optimize_nu ? 1 : 0, dp = optimize_phi ? 1 : 0;
        double cs = (double)K_total * (double)n_cells; double *part = h->d_nu_partial; unsigned *bar = h->d_nu_bar; NuState *st = h->d_nu_state;
        int mi = env_int("PMM_NU_ITERS", 60);
        void *args[] = {&stats, &fm, &nn, &kt, &prm, &dd, &sn, &dmin, &dmax, &eps_s, &eps_nu, &db, &dn, &dp, &R, &cs, &part, &bar, &st, &mi};
        CK(cudaLaunchCooperativeKernel((void *)pmm_mstep_fused, dim3((unsigned)G), dim3(256), args, 0, h->stream));
        h->launches += 1;
        CK(cudaGetLastError());
        return 0;
    }
    if (!lengths_done) {      // (under the ultrametric constraint the lengths and the normalised statistics exist already)
        pmm_mstep_branches<<<(n + 127) / 128, 128, 0, h->stream>>>(h->d_stats, h->d_free, n, K_total, h->d_params, h->d_d, h->d_Snorm,
                                                                   dmin, dmax, eps_s, eps_nu);
        h->launches += 1;
    }
    if (optimize_nu) {
        const bool multi = env_int("PMM_NU_MULTI_LAUNCH", 0) != 0;
        const int iters = env_int("PMM_NU_ITERS", multi ? 24 : 60);
        if (multi) {
            // two launches per Newton iteration, fixed launch count: the state stops changing once `done` is set
            for (int it = 0; it <= iters; it++) {
                pmm_nu_eval<<<NU_BLOCKS, 256, 0, h->stream>>>(h->d_free, n, h->d_Snorm, h->d_d, h->d_nu_state, it == 0, h->d_nu_partial);
                pmm_nu_step<<<1, NU_BLOCKS, 0, h->stream>>>(h->d_nu_partial, h->d_nu_state, it == 0, h->d_params);
            }
            h->launches += 2 * (iters + 1);
        } else {
            // one cooperative launch: a grid barrier per Newton iteration (all CTAs co-resident: <= one per SM)
            int G = std::min(h->n_sms, NU_BLOCKS);
            G = std::max(1, std::min(G, (n + 255) / 256));
            CK(cudaMemsetAsync(h->d_nu_bar, 0, sizeof(unsigned), h->stream));
            const unsigned char *fm = h->d_free; int nn = n; const double *sn = h->d_Snorm, *dd = h->d_d;
            double *part = h->d_nu_partial; unsigned *bar = h->d_nu_bar; NuState *st = h->d_nu_state; double *prm = h->d_params; int mi = iters;
            void *args[] = {&fm, &nn, &sn, &dd, &part, &bar, &st, &prm, &mi};
            CK(cudaLaunchCooperativeKernel((void *)pmm_nu_solve, dim3((unsigned)G), dim3(256), args, 0, h->stream));
            CK(cudaMemsetAsync(h->d_nu_bar, 0, 2 * sizeof(unsigned), h->stream));     // the fused kernel expects zeroed counters
            h->launches += 1;
        }
    }
    if (optimize_phi) {
        pmm_phi_step<<<1, 1, 0, h->stream>>>(h->d_stats, n, h->d_params, R, (double)K_total * (double)n_cells);
        h->launches += 1;
    }
    CK(cudaGetLastError());
    return 0;
}

int pmm_mstep_dev(pmm_handle *h, double R, int32_t n_cells, const uint8_t *free_mask, int32_t optimize_nu, int32_t optimize_phi,
                  double dmin, double dmax, double eps_s, double eps_nu) {
    if (check_ready(h)) return 1;
    if (set_free_mask(h, free_mask)) return 1;
    return mstep_dev_impl(h, R, n_cells, optimize_nu, optimize_phi, dmin, dmax, eps_s, eps_nu);
}

int pmm_mstep(pmm_handle *h, const double *S, double Rtilde, double R, int32_t n_cells, const uint8_t *free_mask, int32_t optimize_nu,
              int32_t optimize_phi, double dmin, double dmax, double eps_s, double eps_nu, double *d, double *nu, double *phi) {
    if (pmm_upload_params(h, d, *nu, *phi)) return 1;
    const size_t n5 = (size_t)5 * h->n;
    CK(cudaMemcpyAsync(h->d_stats, S, sizeof(double) * n5, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_stats + n5, &Rtilde, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (pmm_mstep_dev(h, R, n_cells, free_mask, optimize_nu, optimize_phi, dmin, dmax, eps_s, eps_nu)) return 1;
    return pmm_download_params(h, d, nu, phi);
}

// ---- ultrametric-constrained M-step on the device (pmm_ultrametric_dev.cuh) -------------------------------------
// level phases: a wide level is one phase on the whole grid; a run of narrow levels is one phase on CTA 0
static void um_phases(const std::vector<int> &ptr, std::vector<UmPhase> &ph) {
    ph.clear();
    const int nl = (int)ptr.size() - 1;
    for (int l = 0; l < nl;) {
        if (ptr[l + 1] - ptr[l] > UM_NARROW) { ph.push_back(UmPhase{l, l + 1, 1, 0}); l++; continue; }
        int e = l;
        while (e < nl && ptr[e + 1] - ptr[e] <= UM_NARROW) e++;
        ph.push_back(UmPhase{l, e, 0, 0});
        l = e;
    }
}

static int ensure_um(pmm_handle *h) {
    if (h->um_ready) return 0;
    const size_t n1 = (size_t)h->n + 1;
    if (!h->um_scratch) {
        if (dalloc(h, &h->um_kind, n1) || dalloc(h, &h->um_fixed, n1) || dalloc(h, &h->um_scratch, 11 * n1) || dalloc(h, &h->um_Hprev, n1) ||
            dalloc(h, &h->um_dnew, n1) || dalloc(h, &h->um_partial, (size_t)2 * 1024 * 4) || dalloc(h, &h->um_grounded, n1) ||
            dalloc(h, &h->um_result, 2) || dalloc(h, &h->um_bar, 1))
            return 1;
    }
    std::vector<UmPhase> up, dn;
    um_phases(h->in_ptr, up);
    um_phases(h->out_ptr, dn);
    dfree(h->um_up_ptr); dfree(h->um_dn_ptr); dfree(h->um_up_ph); dfree(h->um_dn_ph);
    if (dalloc(h, &h->um_up_ptr, h->in_ptr.size()) || dalloc(h, &h->um_dn_ptr, h->out_ptr.size()) || dalloc(h, &h->um_up_ph, up.size()) ||
        dalloc(h, &h->um_dn_ph, dn.size()))
        return 1;
    CK(cudaMemcpy(h->um_up_ptr, h->in_ptr.data(), sizeof(int) * h->in_ptr.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->um_dn_ptr, h->out_ptr.data(), sizeof(int) * h->out_ptr.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->um_up_ph, up.data(), sizeof(UmPhase) * up.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->um_dn_ph, dn.data(), sizeof(UmPhase) * dn.size(), cudaMemcpyHostToDevice));
    h->um_n_up_ph = (int)up.size(); h->um_n_dn_ph = (int)dn.size();
    h->um_ready = true; h->um_warm = false;
    return 0;
}

// kinds / fixed lengths of the branches: uploaded when they change (the EM loop repeats them every iteration)
static int um_set_problem(pmm_handle *h, const int8_t *kind, const double *fixed_len) {
    const size_t n = (size_t)h->n;
    if (h->um_kind_host.size() == n && memcmp(h->um_kind_host.data(), kind, n) == 0 &&
        memcmp(h->um_fixed_host.data(), fixed_len, n * sizeof(double)) == 0)
        return 0;
    for (size_t v = 0; v < n; v++)
        if (kind[v] < 0 || kind[v] > 2) return fail(h, "branch kind of node %zu is %d; need 0 (free), 1 (fixed) or 2 (counts as zero)", v, (int)kind[v]);
    h->um_kind_host.assign(kind, kind + n);
    h->um_fixed_host.assign(fixed_len, fixed_len + n);
    CK(cudaMemcpyAsync(h->um_kind, h->um_kind_host.data(), n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->um_fixed, h->um_fixed_host.data(), n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->um_warm = false;              // heights of another problem are no warm start
    return 0;
}

// one M-step under the constraint from the statistics on the device; *status: 0 optimal, 1 infeasible, 2 failed
static int mstep_ultra_impl(pmm_handle *h, double R, int32_t n_cells, int32_t optimize_nu, int32_t optimize_phi, double dmin, double dmax,
                            double eps_s, double eps_nu, int32_t *status, int32_t *newton_iters) {
    const int n = h->n;
    const size_t n1 = (size_t)n + 1;
    UmArgs A;
    A.n = n; A.n_roots = h->n_roots; A.n_int = h->n_int;
    A.parent = h->d_parent; A.left = h->d_left; A.roots = h->d_roots;
    A.up_items = h->d_in_items; A.dn_items = h->d_out_items; A.up_ptr = h->um_up_ptr; A.dn_ptr = h->um_dn_ptr;
    A.up_ph = h->um_up_ph; A.dn_ph = h->um_dn_ph; A.n_up_ph = h->um_n_up_ph; A.n_dn_ph = h->um_n_dn_ph;
    A.kind = h->um_kind; A.fixed_len = h->um_fixed; A.stats = h->d_stats; A.Snorm = h->d_Snorm; A.params = h->d_params;
    A.lo = dmin; A.hi = dmax; A.eps_s = eps_s; A.eps_nu = eps_nu; A.mu_final = 1e-11; A.warm_theta = 1e-6; A.warm_mu = 1e-9;
    A.K_total = h->K_total > 0 ? h->K_total : h->K; A.warm = h->um_warm ? 1 : 0;
    A.H_prev = h->um_Hprev;
    double *sc = h->um_scratch;
    A.H = sc; A.Hint = sc + n1; A.g = sc + 2 * n1; A.a = sc + 3 * n1; A.b = sc + 4 * n1; A.alpha = sc + 5 * n1; A.beta = sc + 6 * n1;
    A.delta = sc + 7 * n1; A.k = sc + 8 * n1; A.S1 = sc + 9 * n1; A.S24 = sc + 10 * n1;
    A.grounded = h->um_grounded; A.partial = h->um_partial; A.bar = h->um_bar; A.d_out = h->um_dnew; A.result = h->um_result;
    CK(cudaMemsetAsync(h->um_bar, 0, sizeof(unsigned), h->stream));
    // one CTA per SM (all co-resident: the kernel synchronises the grid itself), fewer for small trees
    int G = std::max(1, std::min(h->n_sms, (n + UM_THREADS - 1) / UM_THREADS));
    void *args[] = {&A};
    CK(cudaLaunchCooperativeKernel((void *)pmm_um_solve, dim3((unsigned)G), dim3(UM_THREADS), args, 0, h->stream));
    pmm_um_apply<<<(n + 255) / 256, 256, 0, h->stream>>>(n, h->um_kind, h->um_dnew, h->um_result, h->d_d);
    h->launches += 2;
    int res[2] = {0, 0};
    CK(cudaMemcpyAsync(res, h->um_result, sizeof res, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (status) *status = res[0];
    if (newton_iters) *newton_iters = res[1];
    if (res[0] == 1) return 0;                       // infeasible: nothing is updated (EM_solver.py:425-426)
    h->um_warm = true;
    // nu with the new lengths fixed (:427-433), phi in closed form (:318-326); the free branches are those of kind 0
    std::vector<uint8_t> fm(n);
    for (int v = 0; v < n; v++) fm[v] = h->um_kind_host[v] == UM_FREE;
    if (set_free_mask(h, fm.data())) return 1;
    return mstep_dev_impl(h, R, n_cells, optimize_nu, optimize_phi, dmin, dmax, eps_s, eps_nu, true);
}

int pmm_mstep_ultra_dev(pmm_handle *h, double R, int32_t n_cells, const int8_t *kind, const double *fixed_len, int32_t optimize_nu,
                        int32_t optimize_phi, double dmin, double dmax, double eps_s, double eps_nu, int32_t warm, int32_t *status,
                        int32_t *newton_iters) {
    if (check_ready(h)) return 1;
    if (!kind || !fixed_len) return fail(h, "pmm_mstep_ultra_dev: kind / fixed_len are NULL");
    if (ensure_um(h) || um_set_problem(h, kind, fixed_len)) return 1;
    if (!warm) h->um_warm = false;
    return mstep_ultra_impl(h, R, n_cells, optimize_nu, optimize_phi, dmin, dmax, eps_s, eps_nu, status, newton_iters);
}

int pmm_em_ultra(pmm_handle *h, double *d, double *nu, double *phi, const int8_t *kind, const double *fixed_len, int32_t optimize_nu,
                 int32_t optimize_phi, double R, int32_t n_cells, double dmin, double dmax, double conv_eps, int32_t max_iter,
                 double *nllh, int32_t *iters, int32_t *status) {
    if (!nu || !phi) return fail(h, "pmm_em_ultra: nu / phi are NULL");
    if (pmm_upload_params(h, d, *nu, *phi)) return 1;
    if (!kind || !fixed_len) return fail(h, "pmm_em_ultra: kind / fixed_len are NULL");
    if (ensure_um(h) || um_set_problem(h, kind, fixed_len)) return 1;
    h->um_warm = false;                                  // every EM run starts its barrier path from the interior point
    const size_t n5 = (size_t)5 * h->n;
    double pre = 0, cur = 0;
    // EM_solver.py:467 pre_llh, then per iteration E, M, llh (:475-505); as in pmm_em the inside pass that yields
    // llh(p_k) is the one the next E-step needs, so one E-step per iteration serves both
    if (run_passes(h, true, false)) return 1;
    CK(cudaMemcpyAsync(&pre, h->d_stats + n5 + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (waits_ok(h)) return 1;
    cur = pre;
    int it = 1, st = 0, last = 0;
    while (it <= max_iter) {
        if (mstep_ultra_impl(h, R, n_cells, optimize_nu, optimize_phi, dmin, dmax, 1e-6, 1e-5, &st, nullptr)) return 1;
        if (st == 1) break;                              // "d_infeasible": the parameters of the previous iterate stand (:497-499)
        last = st;
        if (run_passes(h, true, false)) return 1;
        int gave_up = 0;                                 // (rides along with the likelihood: one synchronisation)
        if (h->peer_world > 1) CK(cudaMemcpyAsync(&gave_up, h->d_peer_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(&cur, h->d_stats + n5 + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (gave_up && waits_ok(h)) return 1;
        if (std::fabs((cur - pre) / pre) < conv_eps) break;
        pre = cur;
        it++;
    }
    if (nllh) *nllh = (st == 1) ? -pre : -cur;
    if (iters) *iters = it;
    if (status) *status = (st == 1) ? 1 : last;
    return pmm_download_params(h, d, nu, phi);
}

int pmm_em_step(pmm_handle *h, double *d, double *nu, double *phi, const uint8_t *free_mask, int32_t optimize_nu,
                int32_t optimize_phi, double R, int32_t n_cells, double dmin, double dmax, double *llh) {
    if (!nu || !phi) return fail(h, "pmm_em_step: nu / phi are NULL");
    // a page-locked d (cudaHostAlloc / cudaHostRegister, e.g. torch's pin_memory) is read and written in place by the copy engines
    const bool direct = d && host_pinned(d);
    if (upload_impl(h, d, *nu, *phi, direct)) return 1;
    if (set_free_mask(h, free_mask)) return 1;
    if (run_passes(h, true, false)) return 1;
    if (mstep_dev_impl(h, R, n_cells, optimize_nu, optimize_phi, dmin, dmax, 1e-6, 1e-5)) return 1;
    // one synchronisation for everything that comes back: d | nu | phi | llh (at the INPUT point)
    const int n = h->n;
    double *st = h->pin_down;
    CK(cudaMemcpyAsync(direct ? d : st, h->d_d, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(st + n, h->d_params, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(st + n + 2, h->d_stats + (size_t)5 * n + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (waits_ok(h)) return 1;
    if (!direct) memcpy(d, st, sizeof(double) * n);
    *nu = st[n]; *phi = st[n + 1];
    if (llh) *llh = st[n + 2];
    return 0;
}

int pmm_em(pmm_handle *h, double *d, double *nu, double *phi, const uint8_t *free_mask, int32_t optimize_nu, int32_t optimize_phi,
           double R, int32_t n_cells, double dmin, double dmax, double conv_eps, int32_t max_iter, double *nllh, int32_t *iters,
           int32_t *status) {
    if (pmm_upload_params(h, d, *nu, *phi)) return 1;
    const int n = h->n;
    const size_t n5 = (size_t)5 * n;
    if (set_free_mask(h, free_mask)) return 1;
    double pre = 0, cur = 0;
    // EM_solver.py:467 pre_llh, then per iteration E, M, llh (:475-505).  The inside pass that
    // yields llh(p_k) is the same one the next E-step needs, so it is run once per iteration.
    if (run_passes(h, true, false)) return 1;
    CK(cudaMemcpyAsync(&pre, h->d_stats + n5 + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (waits_ok(h)) return 1;
    cur = pre;
    int it = 1;
    while (it <= max_iter) {
        if (mstep_dev_impl(h, R, n_cells, optimize_nu, optimize_phi, dmin, dmax, 1e-6, 1e-5)) return 1;
        if (run_passes(h, true, false)) return 1;
        int gave_up = 0;                                 // (rides along with the likelihood: one synchronisation)
        if (h->peer_world > 1) CK(cudaMemcpyAsync(&gave_up, h->d_peer_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(&cur, h->d_stats + n5 + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (gave_up && waits_ok(h)) return 1;
        if (std::isnan(cur)) return fail(h, "pmm_em: the E-step gave no likelihood (a wait of a dataflow launch or of the exchange over peer memory gave up)");
        if (std::fabs((cur - pre) / pre) < conv_eps) break;
        pre = cur;
        it++;
    }
    if (nllh) *nllh = -cur;
    if (iters) *iters = it;
    if (status) *status = 0;
    return pmm_download_params(h, d, nu, phi);
}


}  // extern "C"

// ultra: every candidate's local M-step under the ultrametric equalities (pmm_nni_ultra.cuh); needs lengths that
// are ultrametric already, because everything outside the five free branches is a constant of the candidate
static int score_nni_impl(pmm_handle *h, bool ultra, int32_t n_cand, const int32_t *cand_u, const int32_t *cand_k, const double *d,
                          double nu, double phi, double dmin, double dmax, double conv_eps, int32_t max_iter, double *out_nllh,
                          double *out_d5, int32_t *out_iters, int32_t *out_nodes) {
    if (n_cand <= 0) return 0;
    if (h && h->peer_world > 1)
        return fail(h, "pmm_score_nni_batch scores candidates on ALL sites of a handle: shard the candidates over the ranks "
                       "(dist.score_nni_sharded), not the sites (this handle exchanges statistics with %d peers)", h->peer_world - 1);
    if (pmm_upload_params(h, d, nu, phi)) return 1;
    const bool lumped_nni = h->engine == 0 && !env_int("PMM_NNI_ROWS", 0);
    if (ultra && !lumped_nni) return fail(h, "pmm_score_nni_batch_ultra runs on the lumped-state engine only");
    if (lumped_nni && h->lKc < h->K) return fail(h, "pmm_score_nni_batch needs all sites resident (1 site chunk)");
    if (!lumped_nni && ensure_fullstate(h)) return 1;       // PMM_NNI_ROWS=1 / PMM_ENGINE=fullstate: cached full-state rows (pmm_nni.cuh)
    if (!lumped_nni && h->chunks.size() != 1) return fail(h, "pmm_score_nni_batch needs all sites resident (1 site chunk, have %zu)", h->chunks.size());
    const int N = h->n_leaves, n = h->n;
    std::vector<int> root_index(n, 0);
    std::vector<int> order; order.reserve(n);                    // parents before children
    {   // root index per node (host copy of the tree)
        int nr = 0;
        for (int v = 0; v < n; v++) if (h->parent[v] < 0) { root_index[v] = nr++; order.push_back(v); }
        for (size_t i = 0; i < order.size(); i++) {
            const int v = order[i];
            if (h->left[v] >= 0) for (int c : {h->left[v], h->right[v]}) { root_index[c] = root_index[v]; order.push_back(c); }
        }
    }
    // heights above the leaves (ultra): longest and shortest leaf path below every node must agree
    std::vector<double> height;
    double super_height = 0.0;
    if (ultra) {
        const double tol = 1e-9;
        std::vector<double> hmin(n, 0.0);
        height.assign(n, 0.0);
        for (size_t i = order.size(); i-- > 0;) {
            const int v = order[i];
            if (h->left[v] < 0) continue;
            const int a = h->left[v], b = h->right[v];
            height[v] = std::fmax(height[a] + d[a], height[b] + d[b]);
            hmin[v] = std::fmin(hmin[a] + d[a], hmin[b] + d[b]);
            if (!(height[v] - hmin[v] <= tol))
                return fail(h, "pmm_score_nni_batch_ultra: the current lengths are not ultrametric (node %d: leaf paths differ by %.3g)",
                            v, height[v] - hmin[v]);
        }
        double lo_s = INFINITY;
        for (int v = 0; v < n; v++) if (h->parent[v] < 0) {
            super_height = std::fmax(super_height, height[v] + d[v]);
            lo_s = std::fmin(lo_s, height[v] + d[v]);
        }
        if (h->n_roots > 1 && !(super_height - lo_s <= tol))
            return fail(h, "pmm_score_nni_batch_ultra: the trees of the forest are not equally high (%.3g)", super_height - lo_s);
    }
    std::vector<NniUltra> ultras(ultra ? n_cand : 0);
    std::vector<NniCand> cands(n_cand);
    std::vector<NniState> st(n_cand);
    for (int i = 0; i < n_cand; i++) {
        const int u = cand_u[i], k = cand_k[i];
        if (u < N || u >= n || h->parent[u] < 0 || (k != 0 && k != 1))
            return fail(h, "NNI candidate %d: node %d must be an internal, non-root node and k in {0,1}", i, u);
        const int v = h->parent[u];
        const int w = (h->left[v] == u) ? h->right[v] : h->left[v];
        const int ck = k ? h->right[u] : h->left[u], co = k ? h->left[u] : h->right[u];
        NniCand c;
        c.v = v; c.u = u; c.A = co; c.B = w; c.C = ck;
        c.g = h->parent[v];
        c.s = (c.g >= 0) ? ((h->left[c.g] == v) ? h->right[c.g] : h->left[c.g]) : -1;
        c.root = root_index[u];
        cands[i] = c;
        NniState z;
        z.d[0] = d[v]; z.d[1] = d[u]; z.d[2] = d[ck]; z.d[3] = d[co]; z.d[4] = d[w];
        z.llh_prev = z.llh_cur = 0.0; z.iters = 0; z.done = 0;
        st[i] = z;
        if (out_nodes) { int32_t *o = out_nodes + (size_t)5 * i; o[0] = v; o[1] = u; o[2] = ck; o[3] = co; o[4] = w; }
        if (ultra) {
            NniUltra q;
            q.hA = height[co]; q.hB = height[w]; q.hC = height[ck];
            q.top_free = (c.g < 0 && h->n_roots == 1) ? 1 : 0;
            q.hTop = (c.g >= 0) ? height[c.g] : super_height;     // a root of a forest hangs from the common top
            q.pad = 0;
            ultras[i] = q;
        }
    }
    if (lumped_nni) {
        // ---- lumped records: one E-step that stores every out-record, then ONE launch, one CTA per candidate
        h->force_store_all = true;
        const int rc = run_passes(h, true, false);
        h->force_store_all = false;
        if (rc) return 1;
        double llh_total = 0.0;
        CK(cudaMemcpyAsync(&llh_total, h->d_stats + (size_t)5 * n + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if ((size_t)n_cand > h->nni_cap || (size_t)h->n_roots > h->nni_roots_cap) {
            dfree(h->d_nni_c); dfree(h->d_nni_s); dfree(h->d_nni_u); dfree(h->d_nni_rsum);
            h->nni_cap = h->nni_roots_cap = 0;
            const size_t cap = (size_t)n_cand + (size_t)n_cand / 8, rcap = (size_t)h->n_roots;
            if (dalloc(h, &h->d_nni_c, cap) || dalloc(h, &h->d_nni_s, cap) || dalloc(h, &h->d_nni_u, cap) || dalloc(h, &h->d_nni_rsum, rcap)) return 1;
            h->nni_cap = cap; h->nni_roots_cap = rcap;
        }
        NniCand *d_c = h->d_nni_c; NniState *d_s = h->d_nni_s; double *d_rsum = h->d_nni_rsum; NniUltra *d_u = h->d_nni_u;
        auto cleanup = [&]() {};
        cudaError_t e = cudaMemcpyAsync(d_c, cands.data(), sizeof(NniCand) * n_cand, cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess && ultra) e = cudaMemcpyAsync(d_u, ultras.data(), sizeof(NniUltra) * n_cand, cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_s, st.data(), sizeof(NniState) * n_cand, cudaMemcpyHostToDevice, h->stream);
        for (int r = 0; r < h->n_roots && e == cudaSuccess; r++)
            pmm_reduce_sum<<<1, 1024, 0, h->stream>>>(h->d_Lroot_llh + (size_t)r * h->K, h->K, d_rsum + r, 0);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);          // llh_total is on the host now
        if (e == cudaSuccess) {
            const LParams P = make_lparams(h, 0, false);
            const int K_total = h->K_total > 0 ? h->K_total : h->K;
            if (ultra)
                pmm_nni_lumped<true><<<(unsigned)n_cand, NNIL_THREADS, 0, h->stream>>>(P, d_c, d_s, d_rsum, llh_total, K_total, dmin, dmax, 1e-6,
                                                                                      1e-5, conv_eps, max_iter, d_u);
            else
                pmm_nni_lumped<false><<<(unsigned)n_cand, NNIL_THREADS, 0, h->stream>>>(P, d_c, d_s, d_rsum, llh_total, K_total, dmin, dmax, 1e-6,
                                                                                       1e-5, conv_eps, max_iter, nullptr);
            h->launches += 1 + h->n_roots;
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(st.data(), d_s, sizeof(NniState) * n_cand, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        cleanup();
        if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(h, "pmm_score_nni_batch: %s", cudaGetErrorString(e)); }
        for (int i = 0; i < n_cand; i++) {
            const bool bad = st[i].done == 2;               // constrained M-step infeasible / failed: no score
            if (out_nllh) out_nllh[i] = bad ? NAN : -st[i].llh_cur;
            if (out_iters) out_iters[i] = bad ? -1 : st[i].iters;
            if (out_d5) for (int b = 0; b < 5; b++) out_d5[(size_t)5 * i + b] = st[i].d[b];
        }
        return 0;
    }
    // ---- full-state rows: caches of the current tree: one full E-step
    if (run_passes(h, true, false, true)) return 1;
    const Chunk &c0 = h->chunks[0];
    KParams P = make_params(h, c0, false);
    double llh_total = 0.0;
    CK(cudaMemcpyAsync(&llh_total, h->d_stats + (size_t)5 * n + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    NniCand *d_c = nullptr; NniState *d_s = nullptr; double *d_part = nullptr, *d_rsum = nullptr; int *d_active = nullptr;
    auto cleanup = [&]() { dfree(d_c); dfree(d_s); dfree(d_part); dfree(d_rsum); dfree(d_active); };
    const size_t keep = h->device_bytes;
    if (dalloc(h, &d_c, n_cand) || dalloc(h, &d_s, n_cand) || dalloc(h, &d_part, (size_t)n_cand * c0.ntiles * 26) ||
        dalloc(h, &d_rsum, h->n_roots) || dalloc(h, &d_active, 1)) { cleanup(); h->device_bytes = keep; return 1; }
    h->device_bytes = keep;     // scratch of this call only
    cudaError_t e = cudaMemcpyAsync(d_c, cands.data(), sizeof(NniCand) * n_cand, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_s, st.data(), sizeof(NniState) * n_cand, cudaMemcpyHostToDevice, h->stream);
    for (int r = 0; r < h->n_roots && e == cudaSuccess; r++)
        pmm_reduce_sum<<<1, 1024, 0, h->stream>>>(h->d_root_site_llh + (size_t)r * c0.K, c0.K, d_rsum + r, 0);
    const size_t smem = (size_t)5 * h->TS * 8;
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pmm_nni_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);          // llh_total is on the host now
    const int K_total = h->K_total > 0 ? h->K_total : h->K;
    const int check_every = std::max(1, env_int("PMM_NNI_CHECK_EVERY", 4));
    int active = n_cand;
    for (int it = 0; it <= max_iter + 1 && active > 0 && e == cudaSuccess; it++) {
        if (it % check_every == 0) e = cudaMemsetAsync(d_active, 0, sizeof(int), h->stream);
        dim3 grid((unsigned)n_cand, (unsigned)c0.ntiles);
        pmm_nni_sweep<<<grid, NNI_THREADS, smem, h->stream>>>(P, d_c, d_s, d_part);
        pmm_nni_mstep<<<(n_cand + 127) / 128, 128, 0, h->stream>>>(d_part, n_cand, c0.ntiles, d_c, d_rsum, llh_total, d_s, h->d_params,
                                                                   K_total, dmin, dmax, 1e-6, 1e-5, conv_eps, max_iter, d_active);
        h->launches += 2;
        if (it % check_every == check_every - 1) {
            if (e == cudaSuccess) e = cudaMemcpyAsync(&active, d_active, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        }
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(st.data(), d_s, sizeof(NniState) * n_cand, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cleanup();
    if (e != cudaSuccess) return fail(h, "pmm_score_nni_batch: %s", cudaGetErrorString(e));
    for (int i = 0; i < n_cand; i++) {
        if (out_nllh) out_nllh[i] = -st[i].llh_cur;
        if (out_iters) out_iters[i] = st[i].iters;
        if (out_d5) for (int b = 0; b < 5; b++) out_d5[(size_t)5 * i + b] = st[i].d[b];
    }
    return 0;
}

extern "C" {

int pmm_score_nni_batch(pmm_handle *h, int32_t n_cand, const int32_t *cand_u, const int32_t *cand_k, const double *d, double nu,
                        double phi, double dmin, double dmax, double conv_eps, int32_t max_iter, double *out_nllh, double *out_d5,
                        int32_t *out_iters, int32_t *out_nodes) {
    return score_nni_impl(h, false, n_cand, cand_u, cand_k, d, nu, phi, dmin, dmax, conv_eps, max_iter, out_nllh, out_d5, out_iters, out_nodes);
}

int pmm_score_nni_batch_ultra(pmm_handle *h, int32_t n_cand, const int32_t *cand_u, const int32_t *cand_k, const double *d, double nu,
                              double phi, double dmin, double dmax, double conv_eps, int32_t max_iter, double *out_nllh,
                              double *out_d5, int32_t *out_iters, int32_t *out_nodes) {
    return score_nni_impl(h, true, n_cand, cand_u, cand_k, d, nu, phi, dmin, dmax, conv_eps, max_iter, out_nllh, out_d5, out_iters, out_nodes);
}

// host instance of the batch kernel's constrained M-step (pmm_nni_ultra.cuh), for tests on machines without a GPU
int pmm_debug_nni_ultra_mstep(const double *S25, double nu, double eps_nu, double lo, double hi, double hA, double hB, double hC,
                              double hTop, int32_t top_free, int32_t warm, double *H3, double *d5, int32_t *newton_iters) {
    if (!S25 || !H3 || !d5) return -1;
    NniUltra c;
    c.hA = hA; c.hB = hB; c.hC = hC; c.hTop = hTop; c.top_free = top_free; c.pad = 0;
    int it = 0;
    const int rc = ultra2_solve(c, S25, nu, eps_nu, lo, hi, warm != 0, H3, d5, &it);
    if (newton_iters) *newton_iters = it;
    return rc;
}

int pmm_get_info(const pmm_handle *h, pmm_info *info) {
    if (!h || !info) return 1;
    info->n_nodes = h->n; info->n_leaves = h->n_leaves; info->n_sites = h->K;
    info->slots_per_node = h->W; info->row_stride = h->maxWp;
    info->padded_slots_per_node = 0;
    for (auto &c : h->chunks) info->padded_slots_per_node += c.Wp;
    info->n_chunks = h->engine == 0 ? (int64_t)((h->K + std::max(h->lKc, 1) - 1) / std::max(h->lKc, 1)) : (int64_t)h->chunks.size();
    info->n_tiles = h->max_tiles; info->tile_slots = h->TS;
    info->engine = h->engine; info->clades = h->n_clades; info->clade_target = h->clade_target;
    info->launches_per_pass = (int64_t)h->lin_groups.size();
    info->inside_bytes_per_site = h->l_in_bytes_per_site; info->outside_bytes_per_site = h->l_out_bytes_per_site;
    info->inside_levels = (int64_t)h->in_ptr.size() - 1; info->outside_levels = (int64_t)h->out_ptr.size() - 1;
    info->device_bytes = (int64_t)h->device_bytes; info->kernel_launches = h->launches;
    return 0;
}

int pmm_last_timing(const pmm_handle *hc, double *inside_ms, double *outside_ms, double *total_ms) {
    pmm_handle *h = const_cast<pmm_handle *>(hc);
    if (!h || h->n == 0) return 1;
    CK(cudaSetDevice(h->device));
    CK(cudaEventSynchronize(h->ev[3]));
    float a = h->t_inside, b = h->t_outside, tot = 0;
    const bool one_chunk = h->last_full ? h->chunks.size() == 1 : h->lKc >= h->K;
    if (one_chunk) {
        cudaEventElapsedTime(&a, h->ev[1], h->ev[2]);
        cudaEventElapsedTime(&b, h->ev[2], h->ev[3]);
    }
    if (!h->last_full && h->last_sliced) b = 0.0f;      // site slices interleave the passes: `a` covers both
    cudaEventElapsedTime(&tot, h->ev[0], h->ev[3]);
    if (inside_ms) *inside_ms = a;
    if (outside_ms) *outside_ms = b;
    if (total_ms) *total_ms = tot;
    return 0;
}

}  // extern "C"
