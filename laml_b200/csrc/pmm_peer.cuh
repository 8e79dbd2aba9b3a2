// pmm_peer.cuh -- the one exchange of the site-sharded E-step as part of the kernel that produces its input.
//
// Site shards (one process per GPU) each hold partial expected counts; the M-step needs their sum
// (EM_solver.py:247-305 is a sum over sites).  Instead of "finish kernel, then NCCL all_reduce" this is ONE cooperative
// kernel per rank that
//   A  sums the count partials of the rank's site groups (what pmm_lumped_finish does) into the rank's exchange buffer,
//   B  tells every peer "my sums are there" and waits for theirs (flags in peer memory, release / acquire at system scope),
//   C  reduces ITS slice of the vector over all ranks -- loads straight from the peers' buffers over NVLink, always in rank
//      order 0..W-1, so the result does not depend on who computes it -- and pushes the reduced slice into EVERY rank's
//      statistics buffer (reduce-scatter and all-gather in one sweep: each element crosses NVLink once in, W-1 times out),
//   D  tells every peer "my slice is in your buffer" and waits for theirs.
// Every rank ends up with the same bits.  8 MB of statistics (100k cells): each GPU pulls 7 x 1 MB and pushes 7 x 1 MB.
// Buffers: one allocation per rank, shared by CUDA IPC (or plain pointers when the ranks live in one process):
//   [ loc: L doubles | red: L doubles | flags: 2 x PEER_MAX ints ]      L = 5 n + 2 padded to an even count
// `red` IS the handle's statistics buffer (pmm_stats_dev_ptr).  Flags hold the call count (epoch): nothing is reset.
// A wait that sees nothing for PMM_PEER_TIMEOUT_MS (default 30 s: ranks may drift apart between two calls -- one writes a
// checkpoint, another is descheduled) gives up and raises the handle's error word instead of hanging the GPU; the host side
// (pmm_engine.cu: waits_ok) turns it into a failed call at the next synchronisation -- the NaN this kernel leaves in the
// likelihood slot is only a second line of defence: a late peer's push can replace it.
#pragma once
#include "pmm_kernels.cuh"

namespace pmm {

constexpr int PEER_MAX = 8;               // ranks of one NVSwitch domain
constexpr int PEER_THREADS = 512;

struct PeerArgs {
    int world, rank;
    double *loc[PEER_MAX];                // every rank's local sums   (index = rank; own entry = local pointer)
    double *red[PEER_MAX];                // every rank's statistics buffer
    int *flags[PEER_MAX];                 // every rank's flags: [2][PEER_MAX]
    int epoch;                            // this call's number (the host counts; ranks call in lockstep, as with any collective)
    long long patience;                   // SM cycles a wait may last before it gives up (PMM_PEER_TIMEOUT_MS, default 30 s)
    int *err;
    unsigned *bar;                        // grid barrier of this rank: [0] arrivals, [1] CTAs that have left (both zero between launches)
};

__device__ __forceinline__ void peer_signal(int *flag, int epoch) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}
__device__ __forceinline__ void peer_wait(const int *flag, int epoch, long long patience, int *err) {
    const long long t0 = clock64();
    for (unsigned polls = 1;; polls++) {
        int v;
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v - epoch >= 0) break;        // (a peer that is already one call ahead has passed this point too)
        if ((polls & 255u) == 0 && (clock64() - t0 > patience || *(volatile int *)err)) { *err = 2; break; }
        __nanosleep(100);
    }
}

// Same arguments as pmm_lumped_finish (whose sums it forms in the same orders: same bits), plus the exchange.
// Cooperative launch: every CTA is resident.  with_counts = 0 (inside pass only): just the likelihood is exchanged.
__global__ void __launch_bounds__(PEER_THREADS) pmm_finish_peer(const double *__restrict__ Spart, int G, int n5, int accumulate,
                                                                const double *__restrict__ root_site_llh, int n_roots, int K,
                                                                double *__restrict__ site_llh, const double *__restrict__ Rpart, int n_clades,
                                                                double *__restrict__ Rnode, int with_counts, const int *__restrict__ df_err,
                                                                PeerArgs A) {
    __shared__ double sh[PEER_THREADS];
    const int tid = threadIdx.x, nblk = gridDim.x;
    double *loc = A.loc[A.rank];
    unsigned phase = 0;
    auto grid_barrier = [&]() {
        __threadfence_system();            // this thread's stores (local sums / pushed slices) before the arrival
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(A.bar, 1u);
            const unsigned target = (phase + 1u) * (unsigned)nblk;
            while (*((volatile unsigned *)A.bar) < target) {}
            __threadfence();
        }
        phase++;
        __syncthreads();
    };
    auto block_sum = [&](double v) {
        sh[tid] = v;
        __syncthreads();
        for (int o = PEER_THREADS / 2; o > 0; o >>= 1) {
            if (tid < o) sh[tid] += sh[tid + o];
            __syncthreads();
        }
        const double r = sh[0];
        __syncthreads();
        return r;
    };
    // ---- A: local sums.  Counts: grid-strided; likelihood and R~: the last CTA (fixed orders: a strided sum per thread, then a tree)
    if (with_counts)
        for (int i = blockIdx.x * PEER_THREADS + tid; i < n5; i += nblk * PEER_THREADS) {
            double sum = 0.0;
            for (int t = 0; t < G; t++) sum += Spart[(size_t)t * n5 + i];
            loc[i] = accumulate ? loc[i] + sum : sum;
        }
    if (blockIdx.x == nblk - 1) {
        double acc = 0.0;
        for (int s = tid; s < K; s += PEER_THREADS) {
            double sum = 0.0;
            for (int r = 0; r < n_roots; r++) sum += root_site_llh[(size_t)r * K + s];
            site_llh[s] = sum;
            acc += sum;
        }
        double llh = block_sum(acc);
        if (df_err && *df_err) llh = NAN;
        if (tid == 0) loc[n5 + 1] = accumulate ? loc[n5 + 1] + llh : llh;
        if (with_counts) {
            acc = 0.0;
            for (int c = tid; c < n_clades; c += PEER_THREADS) {
                double sum = 0.0;
                for (int t = 0; t < G; t++) sum += Rpart[(size_t)t * n_clades + c];
                Rnode[c] = sum;
                acc += sum;
            }
            const double rt = block_sum(acc);
            if (tid == 0) loc[n5] = accumulate ? loc[n5] + rt : rt;
        }
    }
    grid_barrier();
    // ---- B: "my sums are there" to every peer; wait for every peer's
    const int W = A.world, me = A.rank;
    if (blockIdx.x == 0 && tid < W && tid != me) peer_signal(A.flags[tid] + me, A.epoch);
    if (tid < W && tid != me) peer_wait(A.flags[me] + tid, A.epoch, A.patience, A.err);
    __syncthreads();
    // ---- C: my slice, summed over the ranks in rank order, pushed into every rank's statistics
    const int first = with_counts ? 0 : n5;                      // inside pass only: [R~ | llh] (R~ keeps its value: same local parts)
    const int len = n5 + 2 - first;
    const int per = (((len + W - 1) / W) + 1) & ~1;              // even slice length: with an even `first` every slice starts 16-byte aligned
    const int lo = first + min(len, me * per), hi = first + min(len, (me + 1) * per);
    const bool vec = ((first & 1) == 0);                         // slices start on even elements of 16-byte aligned buffers
    if (vec) {
        const int lo2 = lo >> 1, hi2 = hi >> 1;                  // whole double2 of the slice; an odd last element is handled below
        for (int i = lo2 + blockIdx.x * PEER_THREADS + tid; i < hi2; i += nblk * PEER_THREADS) {
            double2 v[PEER_MAX];
#pragma unroll
            for (int p = 0; p < PEER_MAX; p++)
                if (p < W) v[p] = __ldcg(reinterpret_cast<const double2 *>(A.loc[p]) + i);
            double2 s = v[0];
#pragma unroll
            for (int p = 1; p < PEER_MAX; p++)
                if (p < W) { s.x += v[p].x; s.y += v[p].y; }
#pragma unroll
            for (int p = 0; p < PEER_MAX; p++)
                if (p < W) reinterpret_cast<double2 *>(A.red[p])[i] = s;
        }
    }
    {   // scalar leftovers: the odd last element of the slice (or the whole slice when it starts on an odd element)
        const int a = vec ? (hi & ~1) : lo;
        for (int i = a + blockIdx.x * PEER_THREADS + tid; i < hi; i += nblk * PEER_THREADS) {
            double s = 0.0;
            for (int p = 0; p < W; p++) s += __ldcg(A.loc[p] + i);
            for (int p = 0; p < W; p++) A.red[p][i] = s;
        }
    }
    grid_barrier();
    // ---- D: "my slice is in your buffer"; this rank's kernel ends once every peer's slice has arrived
    if (blockIdx.x == 0) {
        if (tid < W && tid != me) {
            peer_signal(A.flags[tid] + PEER_MAX + me, A.epoch);
            peer_wait(A.flags[me] + PEER_MAX + tid, A.epoch, A.patience, A.err);
        }
        __syncthreads();
        if (tid == 0 && *(volatile int *)A.err) A.red[me][n5 + 1] = NAN;      // an exchange that gave up must not look like a result
    }
    // the last CTA to leave resets the barrier counters for the next launch
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(A.bar + 1, 1u) == (unsigned)nblk - 1u) { A.bar[0] = 0u; A.bar[1] = 0u; __threadfence(); }
    }
}

}  // namespace pmm
