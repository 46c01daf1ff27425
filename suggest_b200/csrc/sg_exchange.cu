// sg_exchange.cu — record-id-range shards, one process per GPU (BASELINE.json config #4): the exchange of
// the per-shard top-k rows and their merge as ONE kernel over NVLink peer memory.
//
// The reference is a single process and has no counterpart; the reduction is FuzzyCollectorManager.Collect merging
// per-segment queues (pkg/suggest/collector.go:165-178) across shards, under Candidate.Less (collector.go:20-26).
//
// Every rank owns one region of HBM [flags | its shard's packed rows | the merged rows of the whole batch], allocated
// with cudaMalloc and opened by the other ranks through CUDA IPC.  A step on rank r:
//   1. sg_tokens_kernel + sg_bitmap_search_kernel write the shard's rows into the region (local HBM);
//   2. sg_exchange_merge_kernel
//      a. start barrier: block 0 stores the step number into flag[r] of every peer (st.release.sys over NVLink), every
//         block spins on the local flags until all peers have published theirs (ld.acquire.sys) - the peers' rows are
//         then complete and visible;
//      b. rank r merges queries [n_q * r / N, n_q * (r + 1) / N): a warp takes 32 consecutive queries, one per lane; the
//         counts of a part for them are one 128-byte load from that part's HBM (peer load), only the valid entries of
//         the rows are fetched (about 0.7 per query on config #2 instead of k), and the k best are stored - valid
//         entries and counts only - into the merged-rows block of EVERY rank (peer stores): the all-gather of the
//         result is these stores;
//      c. end barrier: the last block to finish publishes the step in the peers' end flags and waits for theirs, so when
//         the kernel exits every rank holds the merged rows of the whole batch and nobody reads this rank's shard rows
//         any more (the next step may overwrite them).
// The work and the bytes of the merge are 1/N per rank (the NCCL path all-gathers N full fixed-stride blocks to every rank
// and merges all queries on every rank: suggest_b200/sharding.py, SG_SHARD_EXCHANGE=nccl).
// A rank that never arrives would make the others spin: the waits give up after ~4 s (sg_exchange_status reports it).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "../../include/suggest_b200.h"
#include "sg_common.cuh"
#include "sg_exchange.h"

namespace sg {

namespace {

constexpr uint32_t kExWarps = 4;                     // warps per CTA
constexpr long long kSpinLimitCycles = 8000000000ll;  // ~4 s at 2 GHz

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// wait until every local flag [0, n) has reached `step`; false on timeout
__device__ bool wait_flags(const unsigned long long *flags, uint32_t n, unsigned long long step) {
    const long long t0 = clock64();
    for (uint32_t p = 0; p < n; p++) {
        while (ld_acquire_sys(flags + p) < step) {
            if (clock64() - t0 > kSpinLimitCycles) return false;
            __nanosleep(64);
        }
    }
    return true;
}

struct PartHeads {  // per warp: head of every part's list for the 32 queries of the chunk (lane = query)
    double score[kMaxExchangeRanks][32];
    uint32_t id[kMaxExchangeRanks][32];
    uint32_t left[kMaxExchangeRanks][32];   // entries of the list not yet consumed, the head included
    uint32_t next[kMaxExchangeRanks][32];   // position of the entry behind the head
};

}  // namespace

__global__ void __launch_bounds__(kExWarps * 32) sg_exchange_merge_kernel(ExchangeParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ int s_ok;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t N = p.world;

    // ---- a. start barrier ----
    if (blockIdx.x == 0 && threadIdx.x < N) st_release_sys(p.regions[threadIdx.x] + p.rank, p.step);  // start flag [rank] of peer
    if (threadIdx.x == 0) {
        const bool ok = wait_flags(p.regions[p.rank], N, p.step);
        s_ok = ok ? 1 : 0;
        if (!ok) atomicExch(p.status, 1u);
    }
    __syncthreads();
    if (!s_ok) return;  // (no end barrier either: every rank times out on its own)

    // ---- b. merge this rank's queries ----
    PartHeads *h = (PartHeads *)smem_raw + warp;
    const uint32_t q_lo = (uint32_t)((uint64_t)p.n_q * p.rank / N), q_hi = (uint32_t)((uint64_t)p.n_q * (p.rank + 1) / N);
    const size_t nk = (size_t)p.n_q * p.k;
    const uint32_t n_chunks = (q_hi - q_lo + 31) >> 5;
    const uint32_t warps = gridDim.x * kExWarps;
    for (uint32_t chunk = blockIdx.x * kExWarps + warp; chunk < n_chunks; chunk += warps) {
        const uint32_t q = q_lo + (chunk << 5) + lane;
        const bool live = q < q_hi;
        const size_t row = (size_t)q * p.k;
        bool unsupported = false;
        // counts of every part for the 32 queries (one 128-byte load per part), then the head of every non-empty list:
        // all independent, one NVLink round trip each way
        for (uint32_t s = 0; s < N; s++) {
            const uint32_t *cnts = (const uint32_t *)((const double *)p.shard_rows[s] + nk) + nk;
            uint32_t c = live ? cnts[q] : 0u;
            if (c == kCountUnsupported) { unsupported = true; c = 0; }
            h->left[s][lane] = c < p.k ? c : p.k;
            h->next[s][lane] = 1u;
        }
        for (uint32_t s = 0; s < N; s++) {
            if (h->left[s][lane] == 0u) continue;
            const double *sc = (const double *)p.shard_rows[s];
            h->score[s][lane] = sc[row];
            h->id[s][lane] = ((const uint32_t *)(sc + nk))[row];
        }
        // lane-serial k-way merge of the sorted per-part lists of this lane's query (Candidate.Less: score desc, id asc)
        uint32_t n_out = 0;
        while (live && n_out < p.k) {
            int best = -1;
            double bs = 0.0;
            uint32_t bi = 0;
            for (uint32_t s = 0; s < N; s++) {
                if (h->left[s][lane] == 0u) continue;
                const double s_ = h->score[s][lane];
                const uint32_t i_ = h->id[s][lane];
                if (best < 0 || s_ > bs || (s_ == bs && i_ < bi)) { best = (int)s; bs = s_; bi = i_; }
            }
            if (best < 0) break;
            for (uint32_t d = 0; d < N; d++) {  // the all-gather: the entry goes into every rank's merged rows
                double *osc = (double *)p.merged_rows[d];
                osc[row + n_out] = bs;
                ((uint32_t *)(osc + nk))[row + n_out] = bi;
            }
            n_out++;
            if (--h->left[best][lane] != 0u) {  // advance the part the entry came from (a dependent peer load; lists are short)
                const uint32_t t = h->next[best][lane]++;
                const double *sc = (const double *)p.shard_rows[best];
                h->score[best][lane] = sc[row + t];
                h->id[best][lane] = ((const uint32_t *)(sc + nk))[row + t];
            }
        }
        if (live) {
            const uint32_t c = unsupported ? kCountUnsupported : n_out;
            for (uint32_t d = 0; d < N; d++) ((uint32_t *)((double *)p.merged_rows[d] + nk) + nk)[q] = c;
        }
        __syncwarp();
    }

    // ---- c. end barrier: last block out publishes and waits ----
    __threadfence_system();  // this thread's peer stores are ordered before the counter
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(p.done_counter, 1u);
        if (done == gridDim.x - 1) {
            *p.done_counter = 0u;  // for the next launch (stream-ordered behind this kernel)
            __threadfence_system();
            for (uint32_t d = 0; d < N; d++) st_release_sys(p.regions[d] + kMaxExchangeRanks + p.rank, p.step);
            if (!wait_flags(p.regions[p.rank] + kMaxExchangeRanks, N, p.step)) atomicExch(p.status, 2u);
        }
    }
}

cudaError_t launch_exchange_merge(const ExchangeParams &p, int sm_count, cudaStream_t stream) {
    const size_t smem = sizeof(PartHeads) * kExWarps;
    static_assert(sizeof(PartHeads) * kExWarps <= 48 * 1024, "stays under the shared memory a kernel gets without opting in");
    const uint32_t mine = (uint32_t)((uint64_t)p.n_q * (p.rank + 1) / p.world) - (uint32_t)((uint64_t)p.n_q * p.rank / p.world);
    int blocks = (int)(((mine + 31) / 32 + kExWarps - 1) / kExWarps);
    if (blocks < 1) blocks = 1;
    if (blocks > sm_count * 4) blocks = sm_count * 4;
    sg_exchange_merge_kernel<<<blocks, kExWarps * 32, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace sg

// ---------------------------------------------------------------------------------------------------------------
// C ABI (include/suggest_b200.h: sg_exchange_*)
// ---------------------------------------------------------------------------------------------------------------
namespace {

int ex_fail(int code, const std::string &msg) { return sg_internal_fail(code, msg); }

#define EX_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            cudaGetLastError();                                                                    \
            return ex_fail(e__ == cudaErrorMemoryAllocation ? SG_ERR_NOMEM : SG_ERR_CUDA,          \
                           std::string(#expr) + ": " + cudaGetErrorString(e__));                   \
        }                                                                                          \
    } while (0)

struct DeviceScope {
    int prev = -1;
    cudaError_t set(int dev) {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) return e;
        return dev == prev ? cudaSuccess : cudaSetDevice(dev);
    }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};

constexpr size_t kFlagBytes = 2 * sg::kMaxExchangeRanks * sizeof(unsigned long long);  // 256

}  // namespace

struct sg_exchange {
    int device = 0, sm_count = 0;
    uint32_t rank = 0, world = 1, max_q = 0, max_k = 0;
    size_t block_bytes = 0;            // sg_packed_rows_bytes(max_q, max_k), 256-byte aligned
    uint8_t *region = nullptr;         // [flags | shard rows | merged rows]
    uint8_t *peer[sg::kMaxExchangeRanks] = {};
    bool opened[sg::kMaxExchangeRanks] = {};
    bool connected = false;
    unsigned long long step = 0;
    unsigned int *d_words = nullptr;   // [done counter, status]
};

extern "C" {

int sg_exchange_create(int device, uint32_t rank, uint32_t world, uint32_t max_queries, uint32_t max_k, sg_exchange **out) {
    if (!out) return ex_fail(SG_ERR_INVALID, "null out");
    *out = nullptr;
    if (world < 1 || world > sg::kMaxExchangeRanks || rank >= world) return ex_fail(SG_ERR_INVALID, "world must be in 1..16 and rank below it");
    if (max_queries < 1 || max_k < 1 || max_k > SG_MAX_TOPK_SHARED) return ex_fail(SG_ERR_INVALID, "max_queries / max_k out of range");
    sg_exchange *ex = new (std::nothrow) sg_exchange();
    if (!ex) return ex_fail(SG_ERR_NOMEM, "out of host memory");
    ex->device = device;
    ex->rank = rank;
    ex->world = world;
    ex->max_q = max_queries;
    ex->max_k = max_k;
    ex->block_bytes = ((size_t)sg_packed_rows_bytes(max_queries, max_k) + 255) & ~(size_t)255;
    DeviceScope scope;
    cudaError_t e = scope.set(device);
    cudaDeviceProp prop{};
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    ex->sm_count = prop.multiProcessorCount;
    if (e == cudaSuccess) e = cudaMalloc((void **)&ex->region, kFlagBytes + 2 * ex->block_bytes);
    if (e == cudaSuccess) e = cudaMemset(ex->region, 0, kFlagBytes + 2 * ex->block_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void **)&ex->d_words, 2 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(ex->d_words, 0, 2 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        if (ex->region) cudaFree(ex->region);
        if (ex->d_words) cudaFree(ex->d_words);
        delete ex;
        cudaGetLastError();
        return ex_fail(e == cudaErrorMemoryAllocation ? SG_ERR_NOMEM : SG_ERR_CUDA, cudaGetErrorString(e));
    }
    ex->peer[rank] = ex->region;
    ex->connected = world == 1;
    *out = ex;
    return SG_OK;
}

int sg_exchange_handle(sg_exchange *ex, void *handle_out) {
    if (!ex || !handle_out) return ex_fail(SG_ERR_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == SG_EXCHANGE_HANDLE_BYTES, "handle size");
    DeviceScope scope;
    EX_CUDA(scope.set(ex->device));
    cudaIpcMemHandle_t h;
    EX_CUDA(cudaIpcGetMemHandle(&h, ex->region));
    std::memcpy(handle_out, &h, sizeof(h));
    return SG_OK;
}

int sg_exchange_connect(sg_exchange *ex, const void *handles) {
    if (!ex || !handles) return ex_fail(SG_ERR_INVALID, "null argument");
    if (ex->connected) return SG_OK;
    DeviceScope scope;
    EX_CUDA(scope.set(ex->device));
    for (uint32_t r = 0; r < ex->world; r++) {
        if (r == ex->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const uint8_t *)handles + (size_t)r * sizeof(h), sizeof(h));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return ex_fail(SG_ERR_UNSUPPORTED, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
        }
        ex->peer[r] = (uint8_t *)p;
        ex->opened[r] = true;
    }
    ex->connected = true;
    return SG_OK;
}

int sg_exchange_search(sg_exchange *ex, sg_index *ix, const char *d_q_bytes, const uint32_t *d_q_off, uint32_t n_q, int metric,
                       double alpha, uint32_t k, uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts, void *stream) {
    if (!ex || !ix) return ex_fail(SG_ERR_INVALID, "null argument");
    if (!ex->connected) return ex_fail(SG_ERR_INVALID, "sg_exchange_connect has not been called");
    if (n_q == 0) return SG_OK;
    if (n_q > ex->max_q || k > ex->max_k || k < 1) return ex_fail(SG_ERR_INVALID, "batch or k larger than the exchange was created for");
    DeviceScope scope;
    EX_CUDA(scope.set(ex->device));
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *rows = ex->region + kFlagBytes;
    uint8_t *merged = rows + ex->block_bytes;
    const size_t nk = (size_t)n_q * k;
    if (ex->world == 1) {  // nothing to exchange: the shard's rows are the result
        int rc = sg_search_batch_packed_device(ix, d_q_bytes, d_q_off, n_q, metric, alpha, k, merged, st);
        if (rc != SG_OK) return rc;
    } else {
        int rc = sg_search_batch_packed_device(ix, d_q_bytes, d_q_off, n_q, metric, alpha, k, rows, st);
        if (rc != SG_OK) return rc;
        sg::ExchangeParams p{};
        p.rank = ex->rank;
        p.world = ex->world;
        p.n_q = n_q;
        p.k = k;
        p.step = ++ex->step;
        for (uint32_t r = 0; r < ex->world; r++) {
            p.regions[r] = (unsigned long long *)ex->peer[r];
            p.shard_rows[r] = ex->peer[r] + kFlagBytes;
            p.merged_rows[r] = ex->peer[r] + kFlagBytes + ex->block_bytes;
        }
        p.done_counter = ex->d_words;
        p.status = ex->d_words + 1;
        EX_CUDA(sg::launch_exchange_merge(p, ex->sm_count, st));
    }
    if (d_out_ids) EX_CUDA(cudaMemcpyAsync(d_out_ids, (double *)merged + nk, nk * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    if (d_out_scores) EX_CUDA(cudaMemcpyAsync(d_out_scores, merged, nk * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (d_out_counts) EX_CUDA(cudaMemcpyAsync(d_out_counts, (uint32_t *)((double *)merged + nk) + nk, (size_t)n_q * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    return SG_OK;
}

int sg_exchange_result(sg_exchange *ex, uint32_t n_q, uint32_t k, const uint32_t **d_ids, const double **d_scores, const uint32_t **d_counts) {
    if (!ex) return ex_fail(SG_ERR_INVALID, "null argument");
    if (n_q > ex->max_q || k > ex->max_k) return ex_fail(SG_ERR_INVALID, "batch or k larger than the exchange was created for");
    const uint8_t *merged = ex->region + kFlagBytes + ex->block_bytes;
    const size_t nk = (size_t)n_q * k;
    if (d_scores) *d_scores = (const double *)merged;
    if (d_ids) *d_ids = (const uint32_t *)((const double *)merged + nk);
    if (d_counts) *d_counts = (const uint32_t *)((const double *)merged + nk) + nk;
    return SG_OK;
}

int sg_exchange_status(sg_exchange *ex, void *stream) {
    if (!ex) return ex_fail(SG_ERR_INVALID, "null argument");
    DeviceScope scope;
    EX_CUDA(scope.set(ex->device));
    unsigned int status = 0;
    EX_CUDA(cudaMemcpyAsync(&status, ex->d_words + 1, sizeof(status), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    EX_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (status != 0) return ex_fail(SG_ERR_CUDA, status == 1 ? "shard exchange: a rank did not reach the start barrier" : "shard exchange: a rank did not reach the end barrier");
    return SG_OK;
}

void sg_exchange_free(sg_exchange *ex) {
    if (!ex) return;
    DeviceScope scope;
    scope.set(ex->device);
    cudaDeviceSynchronize();
    for (uint32_t r = 0; r < ex->world; r++)
        if (ex->opened[r]) cudaIpcCloseMemHandle(ex->peer[r]);
    if (ex->region) cudaFree(ex->region);
    if (ex->d_words) cudaFree(ex->d_words);
    cudaGetLastError();
    delete ex;
}

}  // extern "C"
