// sg_api.cu — the C ABI of libsuggest_b200 (include/suggest_b200.h): index handles in HBM, host <-> device
// staging for sg_search_batch, per-call streams and scratch so one handle can be searched from many
// host threads at once (pkg/suggest/service_test.go:19-80 is the reference's concurrency test).
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <shared_mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/suggest_b200.h"
#include "sg_host.h"
#include "sg_kernels.h"

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
constexpr uint32_t kWorkRing = 1024;
constexpr uint32_t kMaxSlices = 16;        // sg_search_batch pipelines a batch in at most this many slices
constexpr uint32_t kMaxChunks = 16;        // ... or, with page-locked rows, lets the queries arrive in at most this many chunks under one launch
constexpr int kMaxWarps = sg::kMaxSearchThreads / 32;
constexpr int kDefaultTblBytes = 8192;     // 16 warps per SM at k = 10; one pass covers 1M documents at 128 per bucket

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

}  // namespace

// the other translation units of the library report through the same thread-local message (sg_last_error)
int sg_internal_fail(int code, const std::string &msg) { return fail(code, msg); }

namespace {

#define SG_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            std::string m__ = std::string(#expr) + ": " + cudaGetErrorString(e__);             \
            cudaGetLastError();                                                                \
            return fail(e__ == cudaErrorMemoryAllocation ? SG_ERR_NOMEM : SG_ERR_CUDA, m__);   \
        }                                                                                      \
    } while (0)

struct DeviceGuard {  // the library never leaves the caller on another device
    int prev = -1;
    cudaError_t set(int dev) {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) return e;
        return dev == prev ? cudaSuccess : cudaSetDevice(dev);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// one in-flight sg_search_batch: its stream and device staging
struct CallCtx {
    cudaStream_t stream = nullptr;   // slices of one call alternate between the two streams so that the copies of
    cudaStream_t stream2 = nullptr;  // one slice overlap the kernel of the other
    cudaStream_t copy_stream = nullptr;  // every H2D copy of a call: a slice's queries never queue behind another slice's kernels
    cudaEvent_t h2d_done[kMaxSlices] = {};
    uint32_t *too_long = nullptr;        // page-locked words: [0] set by the search kernel if a query has too many n-grams (direct result
                                         // path); [8 .. 8 + kMaxChunks) the arrival counter values copied behind every chunk of queries
    DevBuf<uint32_t> arrived;            // the arrival counter in HBM (SearchParams::arrived); only ever grows
    uint32_t arrive_epoch = 0;
    DevBuf<char> q_bytes;
    DevBuf<uint32_t> q_off, ids, counts, work;
    DevBuf<uint8_t> plans;   // per-query plans, sg_plan_kernel -> sg_search_kernel or sg_tokens_kernel -> sg_bitmap_search_kernel
    DevBuf<uint8_t> wtab;    // window tables of the bitmap engine, one set per slice
    DevBuf<double> scores;
    DevBuf<uint4> packed;                // staged rows of sg_candidate entries (sg_search_batch_candidates, pageable rows)
    DevBuf<uint32_t> cand;               // sg_candidates_batch: [query | id | overlap | segment] x cap, then the thresholds
    DevBuf<unsigned long long> cand_total;
};

}  // namespace

// Window tables (sg_device.h: WindowTables) depend on (metric, similarity, mode) only: the first few combinations an
// index sees are kept in HBM and reused by every later call; `ready` orders other streams behind the kernel that fills them.
struct WtabEntry {
    bool used = false;
    int metric = 0, mode = 0;
    double alpha = 0.0;
    uint8_t *tab = nullptr;
    cudaEvent_t ready = nullptr;
};
constexpr int kWtabCache = 4;

struct sg_index {
    sg::HostIndex host;  // posting arrays are dropped after the upload; the description stays
    sg::DevIndex dev{};
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    uint64_t device_bytes = 0;
    std::vector<void *> allocations;
    std::mutex mu;
    std::condition_variable cv;
    int inflight = 0;
    bool closing = false;
    std::vector<CallCtx *> pool;
    uint32_t *work_ring = nullptr;       // kWorkRing query counters for sg_search_batch_device
    std::atomic<uint32_t> work_rr{0};
    uint32_t tbl_bytes = kDefaultTblBytes;
    uint32_t slice_queries = 16384;      // queries per pipelined slice of sg_search_batch
    uint32_t direct_slice_queries = 0;   // ... when the result rows are written straight into page-locked caller buffers
                                         // (SG_DIRECT_SLICE_QUERIES; 0: cut at direct_split)
    std::vector<uint32_t> direct_split{25};  // percent of the batch where the slices end (SG_DIRECT_SPLIT="25"): a small first
                                         // slice so that the kernels start early; the rest travels under its kernels.
                                         // Every further slice costs more in launch gaps and kernel tails than it hides
                                         // (measured: "25" 186 M q/s, "10,40" 184, "6,20,50" 187 / Cosine 140, 137, 132)
    bool direct_out = true;              // SG_DIRECT_OUT=0: always stage the rows in HBM and copy them back
    uint32_t lean_flags_per_query = sg::kFlagsPerQuery, lean_nodes_per_query = sg::kNodesPerQuery;  // SG_LEAN_FLAGS_PER_QUERY / SG_LEAN_NODES_PER_QUERY: scratch a launch
                                         // may use per query (pooled; at most what is allocated) - the tests shrink it to force the fallback
    int direct_chunks = 0;               // SG_DIRECT_CHUNKS: chunks the queries of a call with page-locked rows arrive in under one launch
                                         // (0, the default: slices on two streams - calls of several host threads then overlap; a chunked
                                         // call has the device to itself)
                                         // (search_batch_chunked); 0: such calls are cut into slices like the others
    size_t l2_persist_bytes = 0;         // persisting-L2 carve-out used for the posting array (0: none)
    size_t l2_window_bytes = 0;
    float l2_hit_ratio = 1.0f;
    int force_shift = -1;
    int max_warps = kMaxWarps;
    bool bitmap_engine = false;          // searches run sg_bitmap_search_kernel (default when the bitmaps fit their budget)
    size_t plan_stride = sg::kPlanStride;
    size_t wtab_bytes = 0;               // window tables per launch
    WtabEntry wtab_cache[kWtabCache];    // guarded by mu
    uint32_t n_terms = 0;
    size_t persist_l2_max = 0, policy_window_max = 0;
    bool built_on_device = false;
    bool lean_pipeline = false;          // Suggest top-k runs sg_count_kernel -> sg_resolve_kernel (the exact level is built, or one bit per document)
    std::string fine_note;               // why the exact level was not built, if it was not
};

namespace {

template <typename T>
int upload(sg_index *ix, const std::vector<T> &v, const T **out, size_t min_elems = 1) {
    size_t n = v.size() > min_elems ? v.size() : min_elems;
    void *d = nullptr;
    SG_CUDA(cudaMalloc(&d, n * sizeof(T)));
    ix->allocations.push_back(d);
    ix->device_bytes += n * sizeof(T);
    if (!v.empty()) SG_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    if (v.size() < n) SG_CUDA(cudaMemset((char *)d + v.size() * sizeof(T), 0, (n - v.size()) * sizeof(T)));
    *out = (const T *)d;
    return SG_OK;
}

void decode_runes(const std::string &s, uint32_t *out, int32_t *n) {
    std::string low;
    sg::to_lower((const uint8_t *)s.data(), s.size(), &low);
    *n = 0;
    for (size_t i = 0; i < low.size();) {
        uint32_t r;
        i += (size_t)sg::utf8_decode((const uint8_t *)low.data() + i, low.size() - i, &r);
        out[(*n)++] = r;
    }
}

int env_int(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}

int finish_setup(sg_index *ix);

// device properties and the tokenizer part of the DevIndex view (alphabet ranges go to HBM)
int setup_text(sg_index *ix) {
    DeviceGuard guard;
    SG_CUDA(guard.set(ix->device));
    cudaDeviceProp prop;
    SG_CUDA(cudaGetDeviceProperties(&prop, ix->device));
    if (prop.major < 10) return fail(SG_ERR_UNSUPPORTED, "libsuggest_b200 needs an sm_100 device, found " + std::string(prop.name));
    ix->sm_count = prop.multiProcessorCount;
    ix->smem_optin = prop.sharedMemPerBlockOptin;
    ix->persist_l2_max = (size_t)prop.persistingL2CacheMaxSize;
    ix->policy_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    sg::HostIndex &h = ix->host;
    sg::DevIndex &d = ix->dev;
    d.n = h.text.n;
    d.bits = h.text.bits;
    d.pad_code = h.text.pad_code;
    decode_runes(h.text.wrap_start, d.wrap_start, &d.n_wrap_start);
    decode_runes(h.text.wrap_end, d.wrap_end, &d.n_wrap_end);
    d.wrap_ascii = 1;
    for (int i = 0; i < d.n_wrap_start; i++) d.wrap_ascii &= d.wrap_start[i] < 128;
    for (int i = 0; i < d.n_wrap_end; i++) d.wrap_ascii &= d.wrap_end[i] < 128;
    std::memcpy(d.ascii_code, h.text.ascii_code, sizeof(d.ascii_code));
    d.n_ranges = (int32_t)h.text.ranges.size();
    d.id_base = h.id_base;
    return upload(ix, h.text.ranges, &d.ranges);
}

// host arrays -> HBM, fill the DevIndex view
int finalize(sg_index *ix) {
    int rc = setup_text(ix);
    if (rc != SG_OK) return rc;
    DeviceGuard guard;
    SG_CUDA(guard.set(ix->device));
    sg::HostIndex &h = ix->host;
    sg::DevIndex &d = ix->dev;
    d.n_terms = (uint32_t)h.term_keys.size();
    d.term_mask = (uint32_t)h.ht_keys.size() - 1;
    d.n_segments = h.n_segments;
    d.n_docs = h.n_docs;
    ix->n_terms = d.n_terms;
    {
        std::vector<uint4> table(h.ht_keys.size());
        for (size_t i = 0; i < table.size(); i++)
            table[i] = uint4{(unsigned)h.ht_keys[i], (unsigned)(h.ht_keys[i] >> 32), h.ht_vals[i], 0u};
        if ((rc = upload(ix, table, &d.term_table)) != SG_OK) return rc;
    }
    if ((rc = upload(ix, h.seg_start, &d.seg_start)) != SG_OK) return rc;
    if ((rc = upload(ix, h.list_off, &d.list_off)) != SG_OK) return rc;
    if ((rc = upload(ix, h.postings, &d.postings, 8)) != SG_OK) return rc;
    if ((rc = upload(ix, h.perm, &d.perm)) != SG_OK) return rc;
    d.n_ids = h.n_ids;
    d.bshift = h.bshift;
    d.row_words = h.row_words;
    d.bitmaps = nullptr;
    if (h.row_words) {
        if ((rc = upload(ix, h.bitmaps, &d.bitmaps)) != SG_OK) return rc;
        std::vector<uint32_t>().swap(h.bitmaps);
    }
    std::vector<uint32_t>().swap(h.postings);
    std::vector<uint32_t>().swap(h.list_off);
    std::vector<uint32_t>().swap(h.perm);
    std::vector<uint64_t>().swap(h.ht_keys);
    std::vector<uint32_t>().swap(h.ht_vals);
    return finish_setup(ix);
}

// engine choice, per-index scratch and the tuning knobs; the DevIndex view is complete
int finish_setup(sg_index *ix) {
    DeviceGuard guard;
    SG_CUDA(guard.set(ix->device));
    sg::HostIndex &h = ix->host;
    {
        const char *eng = std::getenv("SG_ENGINE");
        const bool want_scan = eng && std::strcmp(eng, "scancount") == 0;
        if (eng && *eng && !want_scan && std::strcmp(eng, "bitmap") != 0) return fail(SG_ERR_INVALID, "SG_ENGINE must be bitmap or scancount");
        ix->bitmap_engine = h.row_words != 0 && !want_scan;
        if (eng && std::strcmp(eng, "bitmap") == 0 && !ix->bitmap_engine) return fail(SG_ERR_NOMEM, "the bucket bitmaps do not fit their memory budget");
        ix->plan_stride = ix->bitmap_engine ? sg::kTokStride : sg::kPlanStride;
        if (ix->bitmap_engine) {
            SG_CUDA(sg::preload_bitmap_kernels());
            {   // ... and the driver's own memset path, which the pipeline uses on its launching stream
                void *scratch = nullptr;
                SG_CUDA(cudaMalloc(&scratch, 64));
                cudaMemsetAsync(scratch, 0, 64, nullptr);
                cudaStreamSynchronize(nullptr);
                cudaFree(scratch);
            }
            // the exact level under the bitmaps (sg_fine.cu) and with it the count -> resolve pipeline; SG_PIPELINE=classic keeps
            // every search on sg_bitmap_search_kernel
            const char *pl = std::getenv("SG_PIPELINE");
            const bool classic = pl && std::strcmp(pl, "classic") == 0;
            if (pl && *pl && !classic && std::strcmp(pl, "lean") != 0) return fail(SG_ERR_INVALID, "SG_PIPELINE must be lean or classic");
            if (!classic) {
                size_t free_b = 0, total_b = 0;
                uint64_t budget = 1ull << 62;
                if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) budget = free_b / 2;
                const int mb = env_int("SG_FINE_MAX_MB", -1);
                if (mb >= 0) budget = (uint64_t)mb << 20;
                const std::string err = sg::build_fine_level(&ix->dev, h.n_postings, budget, &ix->allocations, &ix->device_bytes);
                if (err.rfind("skip:", 0) == 0) ix->fine_note = err;
                else if (!err.empty()) return fail(SG_ERR_CUDA, "exact level: " + err);
                // One bit per document (small dictionaries): sg_bitmap_search_kernel has every overlap in its planes for free,
                // while the resolve kernel would re-read the lists of every flagged word; measured on the reference's words.dict
                // the pipeline is 1.4x faster for Jaccard 0.5 but 3x slower for Cosine / Dice 0.5.  SG_PIPELINE=lean forces it.
                const bool forced = pl && std::strcmp(pl, "lean") == 0;
                ix->lean_pipeline = ix->dev.fine != nullptr || (forced && ix->dev.bshift == 0);
                if (forced && !ix->lean_pipeline) return fail(SG_ERR_NOMEM, ix->fine_note);
                g_launches.fetch_add(ix->dev.fine ? 2 : 0, std::memory_order_relaxed);
            }
            if (ix->lean_pipeline) ix->plan_stride = (sg::kTokStride + sg::kLeanScratchPerQuery + 16 + 15) & ~(size_t)15;  // + the alignment gap behind the plans
        }
        const size_t rows = sg::kWindowRows;
        ix->wtab_bytes = ((rows * h.n_segments + 15) & ~(size_t)15) + ((rows * h.row_words + 15) & ~(size_t)15) + rows * sizeof(sg::WordRange);
    }
    {
        void *ring = nullptr;
        SG_CUDA(cudaMalloc(&ring, (size_t)kWorkRing * sg::kWorkWords * sizeof(uint32_t)));
        ix->allocations.push_back(ring);
        ix->work_ring = (uint32_t *)ring;
    }
    // tuning knobs (documented in DESIGN.md); defaults are what bench.py measures
    int tb = env_int("SG_TBL_BYTES", kDefaultTblBytes);
    if (tb < 2048) tb = 2048;
    if (tb > 200000) tb = 200000;
    ix->tbl_bytes = ((uint32_t)tb + 15u) & ~15u;
    ix->force_shift = env_int("SG_FORCE_SHIFT", -1);
    {   // sg_search_batch_device takes its query-plan scratch from the stream-ordered pool: keep freed blocks cached
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, ix->device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    if (env_int("SG_L2_PERSIST", 0)) {
        size_t want = ix->persist_l2_max;
        const size_t bytes = ((size_t)h.n_postings + 8) * sizeof(uint32_t);
        if (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
            ix->l2_persist_bytes = want;
            ix->l2_window_bytes = bytes < ix->policy_window_max ? bytes : ix->policy_window_max;
            ix->l2_hit_ratio = bytes <= want ? 1.0f : (float)want / (float)bytes;
        }
    }
    int sq = env_int("SG_SLICE_QUERIES", 16384);
    ix->slice_queries = sq < 256 ? 256u : (uint32_t)sq;
    sq = env_int("SG_DIRECT_SLICE_QUERIES", 0);
    ix->direct_slice_queries = sq <= 0 ? 0u : sq < 256 ? 256u : (uint32_t)sq;
    if (const char *sp = std::getenv("SG_DIRECT_SPLIT")) {
        ix->direct_split.clear();
        for (const char *c = sp; *c;) {
            char *end = nullptr;
            const long v = std::strtol(c, &end, 10);
            if (end == c) break;
            if (v > 0 && v < 100 && (ix->direct_split.empty() || (uint32_t)v > ix->direct_split.back()) && ix->direct_split.size() + 2 < kMaxSlices)
                ix->direct_split.push_back((uint32_t)v);
            c = *end == ',' ? end + 1 : end;
            if (*end != ',' ) break;
        }
    }
    ix->direct_out = env_int("SG_DIRECT_OUT", 1) != 0;
    ix->direct_chunks = env_int("SG_DIRECT_CHUNKS", 0);
    ix->lean_flags_per_query = (uint32_t)std::min<long long>(std::max<long long>(env_int("SG_LEAN_FLAGS_PER_QUERY", (int)sg::kFlagsPerQuery), 0), sg::kFlagsPerQuery);
    ix->lean_nodes_per_query = (uint32_t)std::min<long long>(std::max<long long>(env_int("SG_LEAN_NODES_PER_QUERY", (int)sg::kNodesPerQuery), 0), sg::kNodesPerQuery);
    ix->max_warps = env_int("SG_WARPS", kMaxWarps);
    if (ix->max_warps < 1) ix->max_warps = 1;
    if (ix->max_warps > kMaxWarps) ix->max_warps = kMaxWarps;
    return SG_OK;
}

void destroy(sg_index *ix) {
    if (!ix) return;
    DeviceGuard guard;
    guard.set(ix->device);
    for (CallCtx *c : ix->pool) {
        c->arrived.release(); c->q_bytes.release(); c->q_off.release(); c->ids.release(); c->counts.release(); c->work.release(); c->scores.release(); c->packed.release(); c->plans.release(); c->wtab.release(); c->cand.release(); c->cand_total.release();
        if (c->stream) cudaStreamDestroy(c->stream);
        if (c->stream2) cudaStreamDestroy(c->stream2);
        if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
        for (cudaEvent_t e : c->h2d_done) if (e) cudaEventDestroy(e);
        if (c->too_long) cudaFreeHost(c->too_long);
        delete c;
    }
    for (WtabEntry &w : ix->wtab_cache) {
        if (w.tab) cudaFree(w.tab);
        if (w.ready) cudaEventDestroy(w.ready);
    }
    for (void *p : ix->allocations) cudaFree(p);
    delete ix;
}

int check_config(const sg_config *cfg, sg_index **out) {
    if (!cfg || !out) return fail(SG_ERR_INVALID, "null argument");
    int n_dev = 0;
    SG_CUDA(cudaGetDeviceCount(&n_dev));
    if (cfg->device < 0 || cfg->device >= n_dev) return fail(SG_ERR_INVALID, "no such CUDA device");
    return SG_OK;
}

int make_index(const sg_config *cfg, sg_index **out, sg_index **ixp) {
    int rc = check_config(cfg, out);
    if (rc != SG_OK) return rc;
    sg_index *ix = new (std::nothrow) sg_index();
    if (!ix) return fail(SG_ERR_NOMEM, "out of host memory");
    ix->device = cfg->device;
    std::string err = ix->host.text.init(cfg->ngram_size, cfg->wrap_start, cfg->wrap_end, cfg->pad, cfg->alphabet, cfg->n_alphabet);
    if (!err.empty()) { delete ix; return fail(SG_ERR_UNSUPPORTED, err); }
    {   // bucket bitmaps: at most half of the free HBM, or SG_BITMAP_MAX_MB
        DeviceGuard guard;
        size_t free_b = 0, total_b = 0;
        if (guard.set(cfg->device) == cudaSuccess && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) ix->host.bitmap_budget = free_b / 2;
        const int mb = env_int("SG_BITMAP_MAX_MB", -1);
        if (mb >= 0) ix->host.bitmap_budget = (uint64_t)mb << 20;
        ix->host.want_bshift = env_int("SG_BUCKET_SHIFT", -1);
    }
    *ixp = ix;
    return SG_OK;
}

struct Geometry { int blocks, warps; size_t smem; uint32_t warp_smem; };

// sg_search_batch, page-locked rows: the queries arrive in chunks while the kernels run (SearchParams::arrived)
struct ArriveArgs {
    const uint32_t *d_arrived;
    uint32_t base, chunk_queries;
};

// sg_candidates_batch: device side of SearchParams::cand_* / custom_thr
struct CollectArgs {
    const uint8_t *d_thr;             // nullptr: thresholds of the built-in metric
    unsigned long long *d_total;
    unsigned long long cap;
    uint32_t *d_query, *d_ids, *d_overlap, *d_segment;
};

int geometry(const sg_index *ix, uint32_t n_q, uint32_t k, Geometry *g) {
    uint32_t warp_smem = ix->tbl_bytes + sg::kWarpFixedSmem + k * 12u;  // layout: sg_search_kernel
    warp_smem = (warp_smem + 15u) & ~15u;
    int warps = (int)(ix->smem_optin / warp_smem);
    if (warps > ix->max_warps) warps = ix->max_warps;
    if (warps < 1) return fail(SG_ERR_INVALID, "k too large for the shared-memory top-k at this table size");
    int blocks = ix->sm_count;
    if ((uint64_t)blocks * warps > n_q) {
        blocks = (int)((n_q + warps - 1) / warps);
        if (blocks < 1) blocks = 1;
        if (blocks == 1) warps = (int)(n_q < (uint32_t)warps ? (n_q ? n_q : 1) : warps);
    }
    g->blocks = blocks;
    g->warps = warps;
    g->warp_smem = warp_smem;
    g->smem = (size_t)warps * warp_smem;
    return SG_OK;
}

int validate_search(const sg_index *ix, uint32_t n_q, int metric, double alpha, uint32_t k) {
    if (!ix) return fail(SG_ERR_INVALID, "null index");
    if (k < 1 || k > SG_MAX_TOPK) return fail(SG_ERR_INVALID, "topK is invalid");                    // search.go:18-21
    if (!(alpha > 0.0) || alpha > 1.0) return fail(SG_ERR_INVALID, "similarity shoud be in (0.0, 1.0]");  // search.go:23-25
    if (metric < SG_JACCARD || metric > SG_EXACT) return fail(SG_ERR_INVALID, "unknown metric");
    (void)n_q;
    return SG_OK;
}

int enqueue_search(sg_index *ix, const char *d_q_bytes, const uint32_t *d_q_off, uint32_t n_q, int metric, double alpha,
                   uint32_t k, uint32_t *d_ids, double *d_scores, uint32_t *d_counts, uint32_t *d_stats, uint32_t *d_work,
                   uint8_t *d_plans, uint8_t *d_wtab, cudaStream_t stream, int mode = 0, cudaEvent_t *stage_events = nullptr,
                   const sg::LmContext *d_lm_ctx = nullptr, int sparse_rows = 0, uint32_t *too_long_flag = nullptr,
                   const CollectArgs *collect = nullptr, const ArriveArgs *arrive = nullptr, uint4 *d_packed = nullptr) {
    Geometry g{};
    int rc = SG_OK;
    if (!ix->bitmap_engine && (rc = geometry(ix, n_q, k, &g)) != SG_OK) return rc;
    sg::SearchParams p{};
    p.q_bytes = d_q_bytes;
    p.q_off = d_q_off;
    p.n_q = n_q;
    p.metric = metric;
    p.alpha = alpha;
    p.k = k;
    p.out_ids = d_ids;
    p.out_scores = d_scores;
    p.out_counts = d_counts;
    p.out_packed = d_packed;  // rows of sg_candidate entries instead of d_ids / d_scores (bitmap engine)
    p.stats = d_stats;
    p.work_counter = d_work;
    p.plans = d_plans;
    p.tbl_bytes = ix->tbl_bytes;
    p.warp_smem = g.warp_smem;
    p.force_shift = ix->force_shift;
    p.mode = mode;
    p.lm_ctx = d_lm_ctx;
    p.sparse_rows = sparse_rows;
    p.too_long_flag = too_long_flag;
    if (arrive) {
        p.arrived = arrive->d_arrived;
        p.arrive_base = arrive->base;
        p.chunk_queries = arrive->chunk_queries;
    }
    static const int resolve_debug = env_int("SG_RESOLVE_DEBUG", 0);
    p.debug = (uint32_t)resolve_debug;
    if (collect) {
        if (!ix->bitmap_engine) return fail(SG_ERR_UNSUPPORTED, "sg_candidates_batch needs an index with bitmaps (engine 1)");
        p.custom_thr = collect->d_thr;
        p.cand_total = collect->d_total;
        p.cand_cap = collect->cap;
        p.cand_query = collect->d_query;
        p.cand_ids = collect->d_ids;
        p.cand_overlap = collect->d_overlap;
        p.cand_segment = collect->d_segment;
    }
    if (ix->l2_persist_bytes) {
        // keep the posting array resident in L2: query plans and result rows stream through the same cache
        cudaStreamAttrValue attr{};
        attr.accessPolicyWindow.base_ptr = (void *)ix->dev.postings;
        attr.accessPolicyWindow.num_bytes = ix->l2_window_bytes;
        attr.accessPolicyWindow.hitRatio = ix->l2_hit_ratio;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    }
    if (ix->bitmap_engine) {
        const size_t rows = sg::kWindowRows;
        auto point_tables = [&](uint8_t *tab) {
            p.wt.seg_thr = tab;
            p.wt.word_thr = tab + ((rows * ix->dev.n_segments + 15) & ~(size_t)15);
            p.wt.win = (sg::WordRange *)(p.wt.word_thr + ((rows * ix->dev.row_words + 15) & ~(size_t)15));
        };
        p.warp_smem = (uint32_t)sg::bitmap_warp_smem(k);
        if ((size_t)p.warp_smem * 8 > ix->smem_optin) return fail(SG_ERR_INVALID, "k too large for the shared-memory top-k");
        int per_sm = 0;
        SG_CUDA(sg::bitmap_search_occupancy(ix->device, k, &per_sm));
        // k above kSmemTopK: the per-warp sorted top-k lives in HBM, one slice per warp of the launch (stream-ordered scratch,
        // freed behind the kernels); such calls run sg_bitmap_search_kernel alone
        void *tk_scratch = nullptr;
        if (k > sg::kSmemTopK) {
            int blocks = ix->sm_count * per_sm;
            const int need = (int)((n_q + 7) / 8);
            if (blocks > need) blocks = need;
            SG_CUDA(cudaMallocAsync(&tk_scratch, (size_t)blocks * 8 * (((size_t)k * 12 + 15) & ~(size_t)15), stream));  // 16-byte aligned slices
            p.tk_global = (uint8_t *)tk_scratch;
        }
        struct TkFree {
            void *p; cudaStream_t st;
            ~TkFree() { if (p) cudaFreeAsync(p, st); }
        } tk_free{tk_scratch, stream};
        // window tables: cached per (metric, similarity, mode); a measurement launch (stage_events) always computes them
        bool run_window = true;
        point_tables(d_wtab);
        if (!stage_events && !(collect && collect->d_thr)) {  // a caller-tabulated metric has no key to cache under
            std::lock_guard<std::mutex> lk(ix->mu);
            WtabEntry *hit = nullptr, *vacant = nullptr;
            for (WtabEntry &w : ix->wtab_cache) {
                if (w.used && w.metric == metric && w.mode == mode && w.alpha == alpha) { hit = &w; break; }
                if (!w.used && !vacant) vacant = &w;
            }
            if (hit) {
                point_tables(hit->tab);
                SG_CUDA(cudaStreamWaitEvent(stream, hit->ready, 0));
                run_window = false;
            } else if (vacant) {
                if (!vacant->tab) SG_CUDA(cudaMalloc((void **)&vacant->tab, ix->wtab_bytes));
                if (!vacant->ready) SG_CUDA(cudaEventCreateWithFlags(&vacant->ready, cudaEventDisableTiming));
                point_tables(vacant->tab);
                SG_CUDA(sg::launch_window(ix->dev, p, stream));
                SG_CUDA(cudaEventRecord(vacant->ready, stream));
                g_launches.fetch_add(1, std::memory_order_relaxed);
                vacant->metric = metric;
                vacant->mode = mode;
                vacant->alpha = alpha;
                vacant->used = true;
                run_window = false;
            }
        }
        if (ix->lean_pipeline && mode == 0 && !collect && !d_lm_ctx && k <= sg::kSmemTopK) {
            // scratch of the pipeline behind the plans: [plans n_q x kTokStride | flags | nodes | pending | head]
            uint8_t *at = d_plans + (((size_t)n_q * sg::kTokStride + 15) & ~(size_t)15);
            p.lean_flags = (uint4 *)at;
            at += (size_t)n_q * sg::kFlagsPerQuery * sizeof(uint4);
            p.lean_nodes = (uint4 *)at;
            at += (size_t)n_q * sg::kNodesPerQuery * sizeof(uint4);
            p.lean_pending = (uint32_t *)at;
            p.flag_cap = n_q * ix->lean_flags_per_query;
            p.node_cap = n_q * ix->lean_nodes_per_query;
            p.lean_head = p.lean_pending + n_q;
            int count_per_sm = 0, resolve_per_sm = 0;
            SG_CUDA(sg::lean_occupancy(ix->device, k, &count_per_sm, &resolve_per_sm));
            static const bool want_fused = env_int("SG_FUSED_TOKENS", 1) != 0;
            const bool fused = want_fused && d_stats == nullptr;  // the stats pass lives in sg_tokens_kernel
            if (fused) SG_CUDA(cudaMemsetAsync(d_work, 0, sg::kWorkWords * sizeof(uint32_t), stream));
            SG_CUDA(sg::launch_lean_search(ix->dev, p, ix->sm_count, count_per_sm, resolve_per_sm, per_sm, run_window, fused, stream, stage_events));
            g_launches.fetch_add((run_window ? 4 : 3) + (fused ? 0 : 1), std::memory_order_relaxed);  // [window +] [tokens +] count + resolve + fallback search
            return SG_OK;
        }
        SG_CUDA(sg::launch_bitmap_search(ix->dev, p, ix->sm_count, per_sm, run_window, stream, stage_events));
        g_launches.fetch_add(run_window ? 3 : 2, std::memory_order_relaxed);  // [sg_window_kernel +] sg_tokens_kernel + sg_bitmap_search_kernel
        return SG_OK;
    }
    SG_CUDA(cudaMemsetAsync(d_work, 0, sizeof(uint32_t), stream));
    SG_CUDA(sg::launch_search(ix->dev, p, g.blocks, g.warps, g.smem, stream, stage_events));
    g_launches.fetch_add(2, std::memory_order_relaxed);  // sg_plan_kernel + sg_search_kernel
    return SG_OK;
}

// A whole batch as the device wants it: the caller's bytes if they are all ASCII (the kernels lower A-Z themselves), else
// every query through strings.ToLower on the host (Go's tables, sg_text.cpp) with rebuilt offsets.  sg_search_batch does
// the same per slice.
struct LoweredQueries {
    std::string low;
    std::vector<uint32_t> low_off;
    const char *bytes;
    const uint32_t *off;
    size_t n_bytes;
    LoweredQueries(const char *q_bytes, const uint32_t *q_off, uint32_t n_q) : bytes(q_bytes), off(q_off), n_bytes(q_off[n_q]) {
        unsigned char high = 0;  // no early exit: the loop vectorises
        for (size_t i = 0; i < n_bytes; i++) high |= (unsigned char)q_bytes[i];
        if (!(high & 0x80)) return;
        low_off.resize((size_t)n_q + 1);
        for (uint32_t q = 0; q < n_q; q++) {
            low_off[q] = (uint32_t)low.size();
            sg::to_lower((const uint8_t *)q_bytes + q_off[q], q_off[q + 1] - q_off[q], &low);
        }
        low_off[n_q] = (uint32_t)low.size();
        bytes = low.data();
        off = low_off.data();
        n_bytes = low.size();
    }
    LoweredQueries(const LoweredQueries &) = delete;
    LoweredQueries &operator=(const LoweredQueries &) = delete;
};

struct CtxLease {
    sg_index *ix;
    CallCtx *ctx = nullptr;
    bool counted = false;
    explicit CtxLease(sg_index *i) : ix(i) {}
    int acquire() {
        std::unique_lock<std::mutex> lk(ix->mu);
        if (ix->closing) return fail(SG_ERR_INVALID, "index is being freed");
        ix->inflight++;
        counted = true;
        if (!ix->pool.empty()) { ctx = ix->pool.back(); ix->pool.pop_back(); return SG_OK; }
        lk.unlock();
        ctx = new (std::nothrow) CallCtx();
        if (!ctx) return fail(SG_ERR_NOMEM, "out of host memory");
        cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
        for (cudaEvent_t &ev : ctx->h2d_done)
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaHostAlloc((void **)&ctx->too_long, 64 * sizeof(uint32_t), cudaHostAllocMapped);
        if (e != cudaSuccess) {
            if (ctx->stream) cudaStreamDestroy(ctx->stream);
            if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
            if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
            for (cudaEvent_t ev : ctx->h2d_done) if (ev) cudaEventDestroy(ev);
            delete ctx;
            ctx = nullptr;
            return fail(SG_ERR_CUDA, cudaGetErrorString(e));
        }
        return SG_OK;
    }
    ~CtxLease() {
        std::lock_guard<std::mutex> lk(ix->mu);
        if (ctx) ix->pool.push_back(ctx);
        if (counted) ix->inflight--;
        ix->cv.notify_all();
    }
};

}  // namespace

extern "C" {

int sg_index_build(const sg_config *cfg, const char *doc_bytes, const uint64_t *doc_off, uint32_t n_docs,
                   uint32_t id_base, sg_index **out) {
    sg_index *ix = nullptr;
    int rc = make_index(cfg, out, &ix);
    if (rc != SG_OK) return rc;
    if (n_docs && (!doc_bytes || !doc_off)) { delete ix; return fail(SG_ERR_INVALID, "null documents"); }
    ix->host.id_base = id_base;
    static const uint64_t zero_off[1] = {0};
    // SG_BUILD: "gpu" (device build or an error), "host", default: device build when the dictionary is eligible
    const char *where = std::getenv("SG_BUILD");
    const bool want_host = where && std::strcmp(where, "host") == 0, want_gpu = where && std::strcmp(where, "gpu") == 0;
    if (!want_host && n_docs > 0) {
        rc = setup_text(ix);
        if (rc != SG_OK) { destroy(ix); return rc; }
        sg::GpuBuilt b;
        std::string err;
        {
            DeviceGuard guard;
            cudaError_t e = guard.set(ix->device);
            err = e != cudaSuccess ? std::string("cuda: ") + cudaGetErrorString(e)
                                   : sg::gpu_build(ix->dev, doc_bytes, doc_off, n_docs, ix->host.want_bshift, ix->host.bitmap_budget, &b, &ix->allocations);
        }
        if (err.empty()) {
            sg::DevIndex &d = ix->dev;
            sg::HostIndex &h = ix->host;
            d.term_table = b.term_table; d.term_mask = b.term_mask; d.n_terms = b.n_terms;
            d.n_segments = b.n_segments; d.n_docs = n_docs; d.seg_start = b.seg_start; d.list_off = b.list_off;
            d.postings = b.postings; d.perm = b.perm; d.n_ids = b.n_ids; d.bshift = b.bshift; d.row_words = b.row_words; d.bitmaps = b.bitmaps;
            h.n_docs = n_docs; h.n_segments = b.n_segments; h.n_lists = b.n_lists; h.n_postings = b.n_postings;
            h.n_ids = b.n_ids; h.bshift = b.bshift; h.row_words = b.row_words;
            ix->n_terms = b.n_terms;
            ix->device_bytes += b.device_bytes;
            ix->built_on_device = true;
            g_launches.fetch_add((uint64_t)b.kernel_launches, std::memory_order_relaxed);
            rc = finish_setup(ix);
            if (rc != SG_OK) { destroy(ix); return rc; }
            *out = ix;
            return SG_OK;
        }
        const bool fallback = err.rfind("fallback:", 0) == 0;
        if (!fallback || want_gpu) { destroy(ix); cudaGetLastError(); return fail(fallback ? SG_ERR_UNSUPPORTED : SG_ERR_CUDA, err); }
        // not eligible: release what the attempt left on the device and build on the host
        {
            DeviceGuard guard;
            guard.set(ix->device);
            for (void *p : ix->allocations) cudaFree(p);
            ix->allocations.clear();
            ix->device_bytes = 0;
        }
    }
    std::string err = sg::build_from_docs(&ix->host, doc_bytes, n_docs ? doc_off : zero_off, n_docs);
    if (!err.empty()) { delete ix; return fail(SG_ERR_UNSUPPORTED, err); }
    rc = finalize(ix);
    if (rc != SG_OK) { destroy(ix); return rc; }
    *out = ix;
    return SG_OK;
}

int sg_index_from_lists(const sg_config *cfg, uint32_t n_segments, uint64_t n_lists, const uint32_t *list_segment,
                        const char *term_bytes, const uint64_t *list_term_off, const uint32_t *ids,
                        const uint64_t *list_off, sg_index **out) {
    sg_index *ix = nullptr;
    int rc = make_index(cfg, out, &ix);
    if (rc != SG_OK) return rc;
    if (n_segments == 0) n_segments = 1;
    if (n_lists && (!list_segment || !term_bytes || !list_term_off || !ids || !list_off)) {
        delete ix;
        return fail(SG_ERR_INVALID, "null list arrays");
    }
    std::string err = sg::build_from_lists(&ix->host, n_segments, n_lists, list_segment, term_bytes, list_term_off, ids, list_off);
    if (!err.empty()) { delete ix; return fail(SG_ERR_FORMAT, err); }
    rc = finalize(ix);
    if (rc != SG_OK) { destroy(ix); return rc; }
    *out = ix;
    return SG_OK;
}

int sg_index_open_disk(const sg_config *cfg, const char *hd_path, const char *dl_path, sg_index **out) {
    sg_index *ix = nullptr;
    int rc = make_index(cfg, out, &ix);
    if (rc != SG_OK) return rc;
    if (!hd_path || !dl_path) { delete ix; return fail(SG_ERR_INVALID, "null path"); }
    std::string err = sg::build_from_disk(&ix->host, hd_path, dl_path);
    if (!err.empty()) {
        delete ix;
        return fail(err.rfind("io:", 0) == 0 ? SG_ERR_IO : SG_ERR_FORMAT, err);
    }
    rc = finalize(ix);
    if (rc != SG_OK) { destroy(ix); return rc; }
    *out = ix;
    return SG_OK;
}

void sg_internal_drop_submit_pool(sg_index *ix);  // sg_submit.cpp

void sg_index_free(sg_index *ix) {
    if (!ix) return;
    sg_internal_drop_submit_pool(ix);  // calls submitted and not yet waited for are served first
    {
        std::unique_lock<std::mutex> lk(ix->mu);
        ix->closing = true;
        ix->cv.wait(lk, [ix] { return ix->inflight == 0; });
    }
    destroy(ix);
}

int sg_index_get_info(const sg_index *ix, sg_index_info *info) {
    if (!ix || !info) return fail(SG_ERR_INVALID, "null argument");
    info->n_docs = ix->host.n_docs;
    info->n_segments = ix->host.n_segments;
    info->n_terms = ix->n_terms;
    info->n_lists = ix->host.n_lists;
    info->n_postings = ix->host.n_postings;
    info->device_bytes = ix->device_bytes;
    info->id_base = ix->host.id_base;
    info->device = ix->device;
    return SG_OK;
}

int sg_index_get_layout(const sg_index *ix, sg_index_layout *layout) {
    if (!ix || !layout) return fail(SG_ERR_INVALID, "null argument");
    layout->n_slots = ix->host.n_ids;
    layout->bucket_shift = ix->host.bshift;
    layout->row_words = ix->host.row_words;
    layout->engine = ix->bitmap_engine ? 1u : 0u;
    layout->built_on_device = ix->built_on_device ? 1u : 0u;
    layout->pipeline = ix->lean_pipeline ? 1u : 0u;
    layout->bitmap_bytes = ix->host.row_words ? (uint64_t)(ix->n_terms + 1) * ix->host.row_words * sizeof(uint32_t) : 0;
    return SG_OK;
}

// Device address of [p, p + bytes) if that is page-locked host memory the device can address (cudaHostAlloc,
// cudaHostRegister, sg_pinned_alloc); nullptr for pageable memory.
static void *mapped_host_range(const void *p, size_t bytes) {
    if (!p || !bytes) return nullptr;
    cudaPointerAttributes a{}, b{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess || cudaPointerGetAttributes(&b, (const char *)p + bytes - 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (a.type != cudaMemoryTypeHost || b.type != cudaMemoryTypeHost || !a.devicePointer || !b.devicePointer) return nullptr;
    if ((const char *)b.devicePointer - (const char *)a.devicePointer != (ptrdiff_t)(bytes - 1)) return nullptr;  // two allocations
    return a.devicePointer;
}

int sg_pinned_alloc(uint64_t bytes, void **out) {
    if (!out) return fail(SG_ERR_INVALID, "null out");
    *out = nullptr;
    if (bytes == 0) return SG_OK;
    SG_CUDA(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    return SG_OK;
}

void sg_pinned_free(void *p) {
    if (p) cudaFreeHost(p);
}

int sg_is_pinned(const void *p, uint64_t bytes) { return mapped_host_range(p, (size_t)bytes) != nullptr ? 1 : 0; }

// where the rows of a host-buffer call go: separate id / score arrays (sg_search_batch) or rows of 16-byte sg_candidate
// entries, suggest.Candidate's own layout (sg_search_batch_candidates)
struct OutRows {
    uint32_t *ids;
    double *scores;
    sg_candidate *rows;
};

// Queries the batched kernels refused (more than 128 n-grams: count SG_COUNT_UNSUPPORTED) are answered here, after the
// batch: host tokenization (the index build's chain, sg_text.cpp), one warp per query on the device (sg_long.cu).  The
// reference has no such limit (pkg/merger/list_merger.go:9 saturates overlaps at 0xFFFF).  Synchronous, rare.
static int answer_long_queries(sg_index *ix, const char *q_bytes, const uint32_t *q_off, uint32_t n_q, int metric, double alpha, uint32_t k,
                               const OutRows &out, uint32_t *out_counts, int mode) {
    std::vector<uint32_t> which;
    for (uint32_t q = 0; q < n_q; q++)
        if (out_counts[q] == SG_COUNT_UNSUPPORTED) which.push_back(q);
    if (which.empty()) return SG_OK;
    if (k > sg::kSmemTopK) return fail(SG_ERR_QUERY_TOO_LONG, "query " + std::to_string(which[0]) + " has more than 128 n-grams (not served with topK above 1024)");
    sg::TextConfig text = ix->host.text;
    if (mode == 1) text.wrap_end.clear();  // NewAutocompleteTokenizer: no tail wrap (pkg/suggest/tokenizer.go:23-34)
    std::vector<uint64_t> keys, all_keys;
    std::vector<uint32_t> key_off{0u};
    sg::TokenScratch scratch;
    for (uint32_t q : which) {
        sg::tokenize_keys(text, (const uint8_t *)q_bytes + q_off[q], q_off[q + 1] - q_off[q], &keys, &scratch);
        if (keys.size() > 0xFFFFu) return fail(SG_ERR_QUERY_TOO_LONG, "query " + std::to_string(q) + " has more than 65535 n-grams");
        all_keys.insert(all_keys.end(), keys.begin(), keys.end());
        key_off.push_back((uint32_t)all_keys.size());
    }
    const uint32_t n_long = (uint32_t)which.size();
    const int blocks = n_long < 4 ? (int)n_long : 4;
    struct Dev {
        std::vector<void *> ptrs;
        cudaError_t get(void **p, size_t bytes) {
            cudaError_t e = cudaMalloc(p, bytes ? bytes : 4);
            if (e == cudaSuccess) ptrs.push_back(*p);
            return e;
        }
        ~Dev() { for (void *p : ptrs) cudaFree(p); }
    } dev;
    sg::LongParams lp{};
    uint64_t *d_keys = nullptr;
    uint32_t *d_off = nullptr;
    SG_CUDA(dev.get((void **)&d_keys, all_keys.size() * 8));
    SG_CUDA(dev.get((void **)&d_off, key_off.size() * 4));
    SG_CUDA(dev.get((void **)&lp.terms, all_keys.size() * 4));
    SG_CUDA(dev.get((void **)&lp.counters, (size_t)blocks * ix->dev.n_ids * 4));
    SG_CUDA(dev.get((void **)&lp.out_ids, (size_t)n_long * k * 4));
    SG_CUDA(dev.get((void **)&lp.out_scores, (size_t)n_long * k * 8));
    SG_CUDA(dev.get((void **)&lp.out_counts, (size_t)n_long * 4));
    if (!all_keys.empty()) SG_CUDA(cudaMemcpy(d_keys, all_keys.data(), all_keys.size() * 8, cudaMemcpyHostToDevice));
    SG_CUDA(cudaMemcpy(d_off, key_off.data(), key_off.size() * 4, cudaMemcpyHostToDevice));
    lp.keys = d_keys;
    lp.key_off = d_off;
    lp.n_long = n_long;
    lp.metric = metric;
    lp.mode = mode;
    lp.alpha = alpha;
    lp.k = k;
    SG_CUDA(sg::launch_long_queries(ix->dev, lp, blocks, nullptr));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    std::vector<uint32_t> ids((size_t)n_long * k), counts(n_long);
    std::vector<double> scores((size_t)n_long * k);
    SG_CUDA(cudaMemcpy(ids.data(), lp.out_ids, ids.size() * 4, cudaMemcpyDeviceToHost));
    SG_CUDA(cudaMemcpy(scores.data(), lp.out_scores, scores.size() * 8, cudaMemcpyDeviceToHost));
    SG_CUDA(cudaMemcpy(counts.data(), lp.out_counts, counts.size() * 4, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < n_long; i++) {
        const uint32_t q = which[i];
        out_counts[q] = counts[i];
        if (out.rows) {
            for (uint32_t j = 0; j < counts[i]; j++) out.rows[(size_t)q * k + j] = sg_candidate{ids[(size_t)i * k + j], 0u, scores[(size_t)i * k + j]};
            continue;
        }
        std::memcpy(out.ids + (size_t)q * k, ids.data() + (size_t)i * k, (size_t)counts[i] * 4);
        std::memcpy(out.scores + (size_t)q * k, scores.data() + (size_t)i * k, (size_t)counts[i] * 8);
    }
    return SG_OK;
}

// sg_search_batch with page-locked result rows, 16,384 queries or more: ONE set of kernels for the whole batch, started at
// once; the queries travel to the device in chunks on the copy stream while the kernels run, a counter copied behind every
// chunk tells the tokenizing kernel how far they are (wait_for_query, sg_common.cuh).  Only the first chunk's copy (~12 us)
// is exposed, and every kernel's start-up and tail is paid once per call instead of once per slice.
// A chunk that holds non-ASCII bytes is lower-cased on the host (strings.ToLower) into an area of its own behind the raw
// bytes, with offsets of its own; that is why every chunk carries its n + 1 offsets (kArriveOffPad apart).
static int search_batch_chunked(sg_index *ix, CallCtx *c, const char *q_bytes, const uint32_t *q_off, uint32_t n_q, int metric, double alpha,
                                uint32_t k, uint32_t *m_ids, double *m_scores, uint4 *m_rows, uint32_t *out_counts, int mode, uint32_t n_chunks) {
    const uint32_t total = q_off[n_q];
    uint32_t chunk_q = (n_q + n_chunks - 1) / n_chunks;
    chunk_q = (chunk_q + 31u) & ~31u;  // whole 128-byte lines of offsets per chunk
    n_chunks = (n_q + chunk_q - 1) / chunk_q;
    const size_t raw_area = ((size_t)total + 127) & ~(size_t)127;
    SG_CUDA(c->q_bytes.reserve(raw_area + (size_t)total * 3 + 128 * (size_t)n_chunks + 128));  // raw bytes | lowered chunks (worst case 3x)
    SG_CUDA(c->q_off.reserve((size_t)n_chunks * (chunk_q + sg::kArriveOffPad) + 1));
    SG_CUDA(c->counts.reserve(n_q));
    SG_CUDA(c->work.reserve((size_t)kMaxSlices * sg::kWorkWords));
    SG_CUDA(c->plans.reserve((size_t)n_q * ix->plan_stride));
    SG_CUDA(c->wtab.reserve(ix->wtab_bytes * kMaxSlices));
    if (!c->arrived.p) {
        SG_CUDA(c->arrived.reserve(1));
        SG_CUDA(cudaMemset(c->arrived.p, 0, sizeof(uint32_t)));
        c->arrive_epoch = 0;
    }
    if (c->arrive_epoch > 0x7FFF0000u) {  // (the kernels compare with a signed difference; start over long before it wraps)
        SG_CUDA(cudaMemset(c->arrived.p, 0, sizeof(uint32_t)));
        c->arrive_epoch = 0;
    }
    static const int trace = env_int("SG_TRACE", 0);
    const auto t_begin = std::chrono::steady_clock::now();
    cudaStream_t st = c->stream, cs = c->copy_stream;
    uint32_t *d_too_long = nullptr;
    c->too_long[0] = c->too_long[1] = 0u;  // [1]: a kernel gave up waiting for a chunk
    SG_CUDA(cudaHostGetDevicePointer((void **)&d_too_long, c->too_long, 0));
    uint32_t *arrive_vals = c->too_long + 8;
    const ArriveArgs arrive{c->arrived.p, c->arrive_epoch, chunk_q};
    std::vector<std::string> low_bytes(n_chunks);
    std::vector<std::vector<uint32_t>> low_off(n_chunks);
    size_t low_cursor = raw_area;
    int rc = SG_OK;
    auto copy_chunk = [&](uint32_t ch) -> int {
        const uint32_t lo = ch * chunk_q, hi = lo + chunk_q < n_q ? lo + chunk_q : n_q;
        const uint32_t b0 = q_off[lo], b1 = q_off[hi];
        uint32_t *d_off = c->q_off.p + (size_t)ch * (chunk_q + sg::kArriveOffPad);
        // the raw bytes go first: nearly every chunk is plain ASCII and is used as it is (the device lowers A-Z itself)
        if (b1 > b0) SG_CUDA(cudaMemcpyAsync(c->q_bytes.p + b0, q_bytes + b0, b1 - b0, cudaMemcpyHostToDevice, cs));
        unsigned char high = 0;  // no early exit: the loop vectorises
        for (uint32_t i = b0; i < b1; i++) high |= (unsigned char)q_bytes[i];
        if (high & 0x80) {
            std::string &lb = low_bytes[ch];
            std::vector<uint32_t> &lof = low_off[ch];
            lof.resize((size_t)(hi - lo) + 1);
            lb.reserve((size_t)(b1 - b0) * 3 / 2 + 16);
            for (uint32_t q = lo; q < hi; q++) {
                lof[q - lo] = (uint32_t)(low_cursor + lb.size());
                sg::to_lower((const uint8_t *)q_bytes + q_off[q], q_off[q + 1] - q_off[q], &lb);
            }
            lof[hi - lo] = (uint32_t)(low_cursor + lb.size());
            if (!lb.empty()) SG_CUDA(cudaMemcpyAsync(c->q_bytes.p + low_cursor, lb.data(), lb.size(), cudaMemcpyHostToDevice, cs));
            SG_CUDA(cudaMemcpyAsync(d_off, lof.data(), lof.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, cs));
            low_cursor += (lb.size() + 127) & ~(size_t)127;
        } else {
            SG_CUDA(cudaMemcpyAsync(d_off, q_off + lo, ((size_t)(hi - lo) + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, cs));
        }
        arrive_vals[ch] = arrive.base + ch + 1u;
        SG_CUDA(cudaMemcpyAsync(c->arrived.p, arrive_vals + ch, sizeof(uint32_t), cudaMemcpyHostToDevice, cs));
        return SG_OK;
    };
    auto enqueue_all = [&]() -> int {
        int r = copy_chunk(0);
        if (r != SG_OK) return r;
        r = enqueue_search(ix, c->q_bytes.p, c->q_off.p, n_q, metric, alpha, k, m_ids, m_scores, c->counts.p, nullptr, c->work.p, c->plans.p,
                           c->wtab.p, st, mode, nullptr, nullptr, 1, d_too_long, nullptr, &arrive, m_rows);
        if (r != SG_OK) return r;
        for (uint32_t ch = 1; ch < n_chunks; ch++)
            if ((r = copy_chunk(ch)) != SG_OK) return r;
        // The copy of the counts is enqueued LAST.  It waits for the kernels, the kernels wait for the chunks: a copy engine
        // serves its queue in order, and were this copy queued ahead of the chunks on the same engine, nothing would ever move.
        SG_CUDA(cudaMemcpyAsync(out_counts, c->counts.p, (size_t)n_q * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        return SG_OK;
    };
    rc = enqueue_all();
    const auto t_enqueued = std::chrono::steady_clock::now();
    if (rc != SG_OK) {
        // a kernel may be waiting for chunks that will never come: make every query it has not seen yet empty (all offsets
        // zero), release it, then drain
        cudaMemsetAsync(c->q_off.p, 0, ((size_t)n_chunks * (chunk_q + sg::kArriveOffPad) + 1) * sizeof(uint32_t), cs);
        arrive_vals[kMaxChunks] = arrive.base + n_chunks;
        cudaMemcpyAsync(c->arrived.p, arrive_vals + kMaxChunks, sizeof(uint32_t), cudaMemcpyHostToDevice, cs);
    }
    c->arrive_epoch += n_chunks;
    {
        const cudaError_t e1 = cudaStreamSynchronize(cs), e2 = cudaStreamSynchronize(st);
        const cudaError_t e = e1 != cudaSuccess ? e1 : e2;
        if (rc == SG_OK && e != cudaSuccess) { cudaGetLastError(); rc = fail(SG_ERR_CUDA, std::string("sg_search_batch: ") + cudaGetErrorString(e)); }
    }
    if (rc != SG_OK) return rc;
    if (((volatile uint32_t *)c->too_long)[1] != 0u) return -101;  // a chunk did not reach the device in time: the caller runs the sliced path
    if (trace)
        std::fprintf(stderr, "sg_search_batch: %u queries in %u chunks under one launch (rows stored into page-locked host memory): enqueue %.1f us, wait %.1f us\n",
                     n_q, n_chunks, std::chrono::duration<double, std::micro>(t_enqueued - t_begin).count(),
                     std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_enqueued).count());
    if (*(volatile uint32_t *)c->too_long == 0u) return SG_OK;
    return -100;  // some query was refused: the caller answers it (answer_long_queries)
}

static int search_batch_impl(sg_index *ix, const char *q_bytes, const uint32_t *q_off, uint32_t n_q, int metric, double alpha,
                             uint32_t k, const OutRows &out, uint32_t *out_counts, int mode) {
    int rc = validate_search(ix, n_q, metric, alpha, k);
    if (rc != SG_OK) return rc;
    if (n_q == 0) return SG_OK;
    uint32_t *const out_ids = out.ids;
    double *const out_scores = out.scores;
    sg_candidate *const out_rows = out.rows;
    if (!q_off || !out_counts || (out_rows ? false : (!out_ids || !out_scores))) return fail(SG_ERR_INVALID, "null buffer");
    if (out_rows && !ix->bitmap_engine) return fail(SG_ERR_UNSUPPORTED, "sg_search_batch_candidates needs an index with bitmaps (engine 1)");
    const uint32_t total = q_off[n_q];
    if (total && !q_bytes) return fail(SG_ERR_INVALID, "null query bytes");

    DeviceGuard guard;
    SG_CUDA(guard.set(ix->device));
    CtxLease lease(ix);
    rc = lease.acquire();
    if (rc != SG_OK) return rc;
    CallCtx *c = lease.ctx;
    // Result rows in page-locked caller buffers: the search kernel stores the valid entries of every row (and its count)
    // straight into them over PCIe while it runs - no staging in HBM, no D2H copy behind the kernel, and therefore no
    // need to cut the batch into small slices to overlap that copy.  Entries at and behind out_counts[q] are not written.
    uint32_t *m_ids = nullptr, *m_counts = nullptr;
    double *m_scores = nullptr;
    uint4 *m_rows = nullptr;
    static_assert(sizeof(sg_candidate) == sizeof(uint4), "sg_candidate is 16 bytes");
    if (ix->direct_out && ix->bitmap_engine) {
        if (out_rows) {
            m_rows = (uint4 *)mapped_host_range(out_rows, (size_t)n_q * k * sizeof(sg_candidate));
        } else {
            m_ids = (uint32_t *)mapped_host_range(out_ids, (size_t)n_q * k * sizeof(uint32_t));
            m_scores = (double *)mapped_host_range(out_scores, (size_t)n_q * k * sizeof(double));
        }
        m_counts = (uint32_t *)mapped_host_range(out_counts, (size_t)n_q * sizeof(uint32_t));
    }
    const bool direct = (out_rows ? m_rows != nullptr : (m_ids && m_scores)) && m_counts;
    // A chunked call has kernels on the device that wait for copies the host has yet to enqueue.  Nothing else of this
    // library is enqueued on the device meanwhile (exclusive; the sliced path holds the lock shared): a copy of another
    // call that waits for ITS kernel could sit in front of our chunks in a copy engine's queue while that kernel waits for
    // the SMs ours holds.  Should the chunks still not arrive (work of the host application in the way), the kernels give
    // up after ~2 s and the batch is answered again by the sliced path, which never waits on the device.
    static std::shared_mutex device_gate[64];
    std::shared_mutex &gate = device_gate[ix->device & 63];
    const int want_chunks = ix->direct_chunks;
    if (direct && want_chunks > 0 && n_q >= 16384) {
        const uint32_t n_chunks = (uint32_t)want_chunks > kMaxChunks ? kMaxChunks : (uint32_t)want_chunks;
        {
            std::unique_lock<std::shared_mutex> exclusive(gate);
            rc = search_batch_chunked(ix, c, q_bytes, q_off, n_q, metric, alpha, k, m_ids, m_scores, m_rows, out_counts, mode, n_chunks);
        }
        if (rc == -100) return answer_long_queries(ix, q_bytes, q_off, n_q, metric, alpha, k, out, out_counts, mode);
        if (rc != -101) return rc;
        // -101: the chunks did not arrive in time; fall through to the sliced path
    }
    std::shared_lock<std::shared_mutex> shared(gate);
    std::vector<uint32_t> bounds{0u};  // slice sl = queries [bounds[sl], bounds[sl + 1])
    if (direct && ix->direct_slice_queries == 0) {
        if (n_q >= 16384)
            for (uint32_t pc : ix->direct_split) bounds.push_back((uint32_t)((uint64_t)n_q * pc / 100));
        bounds.push_back(n_q);
    } else {
        const uint32_t slice_q = direct ? ix->direct_slice_queries : ix->slice_queries;
        uint32_t n = (n_q + slice_q - 1) / slice_q;
        if (n > kMaxSlices) n = kMaxSlices;
        if (n < 1) n = 1;
        for (uint32_t sl = 1; sl <= n; sl++) bounds.push_back((uint32_t)((uint64_t)n_q * sl / n));
    }
    const uint32_t n_slices = (uint32_t)bounds.size() - 1;
    // strings.ToLower maps every invalid UTF-8 byte to U+FFFD (1 -> 3 bytes), so a slice can triple (sg_text.cpp: to_lower)
    SG_CUDA(c->q_bytes.reserve((size_t)total * 3 + 64 * (size_t)n_slices + 64));
    SG_CUDA(c->q_off.reserve((size_t)n_q + n_slices + 1));
    if (!direct) {
        if (out_rows) {
            SG_CUDA(c->packed.reserve((size_t)n_q * k));
        } else {
            SG_CUDA(c->ids.reserve((size_t)n_q * k));
            SG_CUDA(c->scores.reserve((size_t)n_q * k));
        }
    }
    SG_CUDA(c->counts.reserve(n_q));  // direct rows too: counts are staged (one copy per slice instead of a PCIe write per query)
    SG_CUDA(c->work.reserve((size_t)kMaxSlices * sg::kWorkWords));
    SG_CUDA(c->plans.reserve((size_t)n_q * ix->plan_stride));
    SG_CUDA(c->wtab.reserve(ix->wtab_bytes * kMaxSlices));
    // Slices of the batch go down two streams: the H2D / D2H copies of one slice overlap the kernels of the other, and
    // the host-side look at the next slice's bytes overlaps both.  A slice without non-ASCII bytes is copied straight
    // from the caller's buffer (the device lowers A-Z itself); otherwise its queries go through strings.ToLower first.
    std::vector<std::string> low_bytes(n_slices);
    std::vector<std::vector<uint32_t>> low_off(n_slices);
    size_t dev_cursor = 0;
    static const int trace = env_int("SG_TRACE", 0);
    const auto t_begin = std::chrono::steady_clock::now();
    // SG_TRACE=2: device timeline of the slices (events before the H2D copies, before and behind the kernels, behind the D2H)
    std::vector<cudaEvent_t> tev;
    std::vector<double> t_host;
    if (trace >= 2) {
        tev.resize((size_t)n_slices * 4 + 1);
        for (auto &e : tev) SG_CUDA(cudaEventCreate(&e));
        SG_CUDA(cudaEventRecord(tev[(size_t)n_slices * 4], c->copy_stream));
    }
    cudaStream_t cs = c->copy_stream;
    uint32_t *d_too_long = nullptr;
    if (direct) {
        *c->too_long = 0u;
        SG_CUDA(cudaHostGetDevicePointer((void **)&d_too_long, c->too_long, 0));
    }
    // Everything enqueued for the call; any failure leaves through the one cleanup path behind it (the streams are drained
    // before the caller's buffers are handed back: kernels and copies of earlier slices may still be writing them).
    auto enqueue_slices = [&]() -> int {
    for (uint32_t sl = 0; sl < n_slices; sl++) {
        const uint32_t lo = bounds[sl], hi = bounds[sl + 1];
        if (lo == hi) continue;
        cudaStream_t st = (sl & 1) ? c->stream2 : c->stream;
        const uint32_t b0 = q_off[lo], b1 = q_off[hi];
        const char *src_bytes = q_bytes + b0;
        const uint32_t *src_off = q_off + lo;
        size_t n_bytes = b1 - b0;
        char *region = c->q_bytes.p + dev_cursor;
        const char *d_q_bytes = region - b0;  // offsets stay absolute; only [b0, b1) is ever addressed
        if (trace >= 2) {
            t_host.push_back(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count());
            cudaEventRecord(tev[(size_t)sl * 4], cs);
        }
        // Queries travel on the copy stream, the kernels of the slice wait for its event: the copy of slice i + 1 runs
        // under the kernels of slice i instead of queueing behind the kernels of slice i - 1 on its own stream.
        // The copy starts before the host has looked at the bytes: nearly every slice is plain ASCII and goes as it is.
        if (n_bytes) SG_CUDA(cudaMemcpyAsync(region, src_bytes, n_bytes, cudaMemcpyHostToDevice, cs));
        const size_t n_raw = n_bytes;
        unsigned char high = 0;  // no early exit: the loop vectorises
        for (uint32_t i = b0; i < b1; i++) high |= (unsigned char)q_bytes[i];
        if (high & 0x80) {
            std::string &lb = low_bytes[sl];
            std::vector<uint32_t> &lof = low_off[sl];
            lof.resize((size_t)(hi - lo) + 1);
            lb.reserve(n_bytes + n_bytes / 2 + 16);
            for (uint32_t q = lo; q < hi; q++) {
                lof[q - lo] = (uint32_t)lb.size();
                sg::to_lower((const uint8_t *)q_bytes + q_off[q], q_off[q + 1] - q_off[q], &lb);
            }
            lof[hi - lo] = (uint32_t)lb.size();
            src_bytes = lb.data();
            src_off = lof.data();
            n_bytes = lb.size();
            d_q_bytes = region;
            if (n_bytes) SG_CUDA(cudaMemcpyAsync(region, src_bytes, n_bytes, cudaMemcpyHostToDevice, cs));  // replaces the raw bytes
        }
        dev_cursor += ((n_bytes > n_raw ? n_bytes : n_raw) + 63) & ~(size_t)63;
        uint32_t *d_off = c->q_off.p + lo + sl;
        SG_CUDA(cudaMemcpyAsync(d_off, src_off, ((size_t)(hi - lo) + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, cs));
        if (trace >= 2) cudaEventRecord(tev[(size_t)sl * 4 + 1], cs);
        SG_CUDA(cudaEventRecord(c->h2d_done[sl], cs));
        SG_CUDA(cudaStreamWaitEvent(st, c->h2d_done[sl], 0));
        uint4 *d_rows = out_rows ? (direct ? m_rows : c->packed.p) + (size_t)lo * k : nullptr;
        rc = enqueue_search(ix, d_q_bytes, d_off, hi - lo, metric, alpha, k, out_rows ? nullptr : (direct ? m_ids : c->ids.p) + (size_t)lo * k,
                            out_rows ? nullptr : (direct ? m_scores : c->scores.p) + (size_t)lo * k, c->counts.p + lo, nullptr,
                            c->work.p + (size_t)sl * sg::kWorkWords, c->plans.p + (size_t)lo * ix->plan_stride, c->wtab.p + (size_t)sl * ix->wtab_bytes, st,
                            mode, nullptr, nullptr, direct ? 1 : 0, direct ? d_too_long : nullptr, nullptr, nullptr, d_rows);
        if (rc != SG_OK) return rc;
        if (trace >= 2) cudaEventRecord(tev[(size_t)sl * 4 + 2], st);
        if (direct) {
            // The valid entries of the rows were stored by the kernels while they ran.  The counts are not: 65,536 four-byte
            // stores over PCIe cost more than the kernels of the batch take (the link moves ~0.75 G writes per second whatever
            // their size), one copy of the slice's counts is ~10 us.
            SG_CUDA(cudaMemcpyAsync(out_counts + lo, c->counts.p + lo, (size_t)(hi - lo) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            if (trace >= 2) cudaEventRecord(tev[(size_t)sl * 4 + 3], st);
            continue;
        }
        if (out_rows) {
            SG_CUDA(cudaMemcpyAsync(out_rows + (size_t)lo * k, c->packed.p + (size_t)lo * k, (size_t)(hi - lo) * k * sizeof(sg_candidate),
                                    cudaMemcpyDeviceToHost, st));
        } else {
            SG_CUDA(cudaMemcpyAsync(out_ids + (size_t)lo * k, c->ids.p + (size_t)lo * k, (size_t)(hi - lo) * k * sizeof(uint32_t),
                                    cudaMemcpyDeviceToHost, st));
            SG_CUDA(cudaMemcpyAsync(out_scores + (size_t)lo * k, c->scores.p + (size_t)lo * k, (size_t)(hi - lo) * k * sizeof(double),
                                    cudaMemcpyDeviceToHost, st));
        }
        SG_CUDA(cudaMemcpyAsync(out_counts + lo, c->counts.p + lo, (size_t)(hi - lo) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        if (trace >= 2) cudaEventRecord(tev[(size_t)sl * 4 + 3], st);
    }
    return SG_OK;
    };
    rc = enqueue_slices();
    const auto t_enqueued = std::chrono::steady_clock::now();
    {
        const cudaError_t e1 = cudaStreamSynchronize(c->stream), e2 = cudaStreamSynchronize(c->stream2), e3 = cudaStreamSynchronize(cs);
        const cudaError_t e = e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3;
        if (rc == SG_OK && e != cudaSuccess) { cudaGetLastError(); rc = fail(SG_ERR_CUDA, std::string("sg_search_batch: ") + cudaGetErrorString(e)); }
    }
    if (rc != SG_OK) {
        for (auto &e : tev) cudaEventDestroy(e);
        return rc;
    }
    if (trace) {
        const auto t_done = std::chrono::steady_clock::now();
        std::fprintf(stderr, "sg_search_batch: %u queries, %u slices%s: enqueue %.1f us, wait %.1f us\n", n_q, n_slices,
                     direct ? " (rows stored into page-locked host memory)" : "",
                     std::chrono::duration<double, std::micro>(t_enqueued - t_begin).count(),
                     std::chrono::duration<double, std::micro>(t_done - t_enqueued).count());
    }
    if (trace >= 2) {
        const cudaEvent_t e0 = tev[(size_t)n_slices * 4];
        for (uint32_t sl = 0; sl < n_slices && sl < t_host.size(); sl++) {
            float t[4] = {0, 0, 0, 0};
            for (int i = 0; i < 4; i++) cudaEventElapsedTime(&t[i], e0, tev[(size_t)sl * 4 + i]);
            std::fprintf(stderr, "  slice %u: host enqueue at %.1f us; device: h2d %.1f-%.1f us, kernels until %.1f us, rows on the host at %.1f us\n",
                         sl, t_host[sl], t[0] * 1e3, t[1] * 1e3, t[2] * 1e3, t[3] * 1e3);
        }
        for (auto &e : tev) cudaEventDestroy(e);
    }
    if (direct && *(volatile uint32_t *)c->too_long == 0u) return SG_OK;  // the kernel saw no such query: nothing to look for
    return answer_long_queries(ix, q_bytes, q_off, n_q, metric, alpha, k, out, out_counts, mode);
}

int sg_search_batch(sg_index *ix, const char *q_bytes, const uint32_t *q_off, uint32_t n_q, int metric, double alpha,
                    uint32_t k, uint32_t *out_ids, double *out_scores, uint32_t *out_counts) {
    return search_batch_impl(ix, q_bytes, q_off, n_q, metric, alpha, k, OutRows{out_ids, out_scores, nullptr}, out_counts, 0);
}

int sg_search_batch_candidates(sg_index *ix, const char *q_bytes, const uint32_t *q_off, uint32_t n_q, int metric, double alpha,
                               uint32_t k, sg_candidate *out_rows, uint32_t *out_counts) {
    if (!out_rows && n_q) return fail(SG_ERR_INVALID, "null buffer");
    return search_batch_impl(ix, q_bytes, q_off, n_q, metric, alpha, k, OutRows{nullptr, nullptr, out_rows}, out_counts, 0);
}

int sg_autocomplete_batch(sg_index *ix, const char *q_bytes, const uint32_t *q_off, uint32_t n_q, uint32_t limit,
                          uint32_t *out_ids, double *out_scores, uint32_t *out_counts) {
    return search_batch_impl(ix, q_bytes, q_off, n_q, SG_EXACT, 1.0, limit, OutRows{out_ids, out_scores, nullptr}, out_counts, 1);
}

int sg_candidates_batch(sg_index *ix, const char *q_bytes, const uint32_t *q_off, uint32_t n_q, int metric, double alpha,
                        const uint8_t *thresholds, uint64_t cap, uint32_t *out_query, uint32_t *out_ids, uint32_t *out_overlap,
                        uint32_t *out_segment, uint64_t *out_total, uint32_t *out_size_a) {
    if (!ix) return fail(SG_ERR_INVALID, "null index");
    if (!out_total) return fail(SG_ERR_INVALID, "null out_total");
    *out_total = 0;
    if (!thresholds) {
        int rc0 = validate_search(ix, n_q, metric, alpha, 1);
        if (rc0 != SG_OK) return rc0;
    } else {
        metric = SG_JACCARD;  // unused: every threshold comes from the table
        alpha = 1.0;
    }
    if (n_q == 0) return SG_OK;
    if (!q_off || !out_size_a || (cap && (!out_query || !out_ids || !out_overlap || !out_segment))) return fail(SG_ERR_INVALID, "null buffer");
    const uint32_t total_bytes = q_off[n_q];
    if (total_bytes && !q_bytes) return fail(SG_ERR_INVALID, "null query bytes");

    DeviceGuard guard;
    SG_CUDA(guard.set(ix->device));
    CtxLease lease(ix);
    int rc = lease.acquire();
    if (rc != SG_OK) return rc;
    CallCtx *c = lease.ctx;
    LoweredQueries lq(q_bytes, q_off, n_q);
    const char *src_bytes = lq.bytes;
    const uint32_t *src_off = lq.off;
    const size_t n_bytes = lq.n_bytes;
    const size_t thr_bytes = thresholds ? (size_t)sg::kWindowRows * ix->dev.n_segments : 0;
    const size_t thr_words = (thr_bytes + 3) / 4;
    SG_CUDA(c->q_bytes.reserve(n_bytes + 64));
    SG_CUDA(c->q_off.reserve((size_t)n_q + 1));
    SG_CUDA(c->counts.reserve(n_q));
    SG_CUDA(c->work.reserve((size_t)kMaxSlices * sg::kWorkWords));
    SG_CUDA(c->plans.reserve((size_t)n_q * ix->plan_stride));
    SG_CUDA(c->wtab.reserve(ix->wtab_bytes * kMaxSlices));
    SG_CUDA(c->cand.reserve((size_t)cap * 4 + thr_words + 4));
    SG_CUDA(c->cand_total.reserve(1));
    cudaStream_t st = c->stream;
    if (n_bytes) SG_CUDA(cudaMemcpyAsync(c->q_bytes.p, src_bytes, n_bytes, cudaMemcpyHostToDevice, st));
    SG_CUDA(cudaMemcpyAsync(c->q_off.p, src_off, ((size_t)n_q + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SG_CUDA(cudaMemsetAsync(c->cand_total.p, 0, sizeof(unsigned long long), st));
    CollectArgs ca{};
    ca.d_total = c->cand_total.p;
    ca.cap = cap;
    ca.d_query = c->cand.p;
    ca.d_ids = c->cand.p + cap;
    ca.d_overlap = c->cand.p + 2 * cap;
    ca.d_segment = c->cand.p + 3 * cap;
    if (thresholds) {
        uint8_t *d_thr = (uint8_t *)(c->cand.p + 4 * cap);
        SG_CUDA(cudaMemcpyAsync(d_thr, thresholds, thr_bytes, cudaMemcpyHostToDevice, st));
        ca.d_thr = d_thr;
    }
    // k = 1: the top-k of the search kernel stays empty in collect mode; rows are "sparse" so nothing of them is written
    rc = enqueue_search(ix, c->q_bytes.p, c->q_off.p, n_q, metric, alpha, 1, nullptr, nullptr, c->counts.p, nullptr, c->work.p,
                        c->plans.p, c->wtab.p, st, 0, nullptr, nullptr, 1, nullptr, &ca);
    if (rc != SG_OK) { cudaStreamSynchronize(st); return rc; }
    unsigned long long found = 0;
    SG_CUDA(cudaMemcpyAsync(&found, c->cand_total.p, sizeof(found), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(out_size_a, c->counts.p, (size_t)n_q * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    *out_total = found;
    const size_t n_out = (size_t)(found < cap ? found : cap);
    if (n_out) {
        SG_CUDA(cudaMemcpyAsync(out_query, ca.d_query, n_out * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaMemcpyAsync(out_ids, ca.d_ids, n_out * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaMemcpyAsync(out_overlap, ca.d_overlap, n_out * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaMemcpyAsync(out_segment, ca.d_segment, n_out * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaStreamSynchronize(st));
    }
    for (uint32_t q = 0; q < n_q; q++)
        if (out_size_a[q] == SG_COUNT_UNSUPPORTED)
            return fail(SG_ERR_QUERY_TOO_LONG, "query " + std::to_string(q) + " has more than 128 n-grams");
    return SG_OK;
}

static int search_batch_device_impl(sg_index *ix, const char *d_q_bytes, const uint32_t *d_q_off, uint32_t n_q, int metric, double alpha,
                                    uint32_t k, uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts, uint32_t *d_stats,
                                    void *stream, int mode);

int sg_search_batch_device(sg_index *ix, const char *d_q_bytes, const uint32_t *d_q_off, uint32_t n_q, int metric,
                           double alpha, uint32_t k, uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts,
                           uint32_t *d_stats, void *stream) {
    return search_batch_device_impl(ix, d_q_bytes, d_q_off, n_q, metric, alpha, k, d_out_ids, d_out_scores, d_out_counts, d_stats, stream, 0);
}

static int search_batch_device_impl(sg_index *ix, const char *d_q_bytes, const uint32_t *d_q_off, uint32_t n_q, int metric, double alpha,
                                    uint32_t k, uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts, uint32_t *d_stats,
                                    void *stream, int mode) {
    int rc = validate_search(ix, n_q, metric, alpha, k);
    if (rc != SG_OK) return rc;
    if (n_q == 0) return SG_OK;
    if (!d_q_off || !d_out_ids || !d_out_scores || !d_out_counts) return fail(SG_ERR_INVALID, "null buffer");
    DeviceGuard guard;
    SG_CUDA(guard.set(ix->device));
    // work counters for caller-owned streams come from a ring owned by the index: the memset and the
    // kernel are ordered on the caller's stream, and a slot is only reused 1024 launches later
    const uint32_t slot = ix->work_rr.fetch_add(1, std::memory_order_relaxed) & (kWorkRing - 1);
    // the query plans live in stream-ordered scratch: allocated, used by the two kernels and freed on the caller's stream
    void *plans = nullptr;
    const size_t plan_bytes = ((size_t)n_q * ix->plan_stride + 255) & ~(size_t)255;
    SG_CUDA(cudaMallocAsync(&plans, plan_bytes + ix->wtab_bytes, (cudaStream_t)stream));
    rc = enqueue_search(ix, d_q_bytes, d_q_off, n_q, metric, alpha, k, d_out_ids, d_out_scores, d_out_counts, d_stats,
                        ix->work_ring + (size_t)slot * sg::kWorkWords, (uint8_t *)plans, (uint8_t *)plans + plan_bytes, (cudaStream_t)stream, mode);
    cudaError_t fe = cudaFreeAsync(plans, (cudaStream_t)stream);
    if (rc == SG_OK && fe != cudaSuccess) return fail(SG_ERR_CUDA, cudaGetErrorString(fe));
    return rc;
}

int sg_autocomplete_batch_device(sg_index *ix, const char *d_q_bytes, const uint32_t *d_q_off, uint32_t n_q, uint32_t limit,
                                 uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts, uint32_t *d_stats, void *stream) {
    return search_batch_device_impl(ix, d_q_bytes, d_q_off, n_q, SG_EXACT, 1.0, limit, d_out_ids, d_out_scores, d_out_counts, d_stats, stream, 1);
}

static int stage_times_impl(sg_index *ix, const char *d_q_bytes, const uint32_t *d_q_off, uint32_t n_q, int metric, double alpha,
                            uint32_t k, uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts, void *stream,
                            float *ms_out, char *names_out, uint32_t names_cap, int mode);

int sg_search_stage_times(sg_index *ix, const char *d_q_bytes, const uint32_t *d_q_off, uint32_t n_q, int metric, double alpha,
                          uint32_t k, uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts, void *stream,
                          float *ms_out, char *names_out, uint32_t names_cap) {
    return stage_times_impl(ix, d_q_bytes, d_q_off, n_q, metric, alpha, k, d_out_ids, d_out_scores, d_out_counts, stream, ms_out, names_out, names_cap, 0);
}

int sg_autocomplete_stage_times(sg_index *ix, const char *d_q_bytes, const uint32_t *d_q_off, uint32_t n_q, uint32_t limit,
                                uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts, void *stream, float *ms_out,
                                char *names_out, uint32_t names_cap) {
    return stage_times_impl(ix, d_q_bytes, d_q_off, n_q, SG_EXACT, 1.0, limit, d_out_ids, d_out_scores, d_out_counts, stream, ms_out, names_out, names_cap, 1);
}

static int stage_times_impl(sg_index *ix, const char *d_q_bytes, const uint32_t *d_q_off, uint32_t n_q, int metric, double alpha,
                            uint32_t k, uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts, void *stream,
                            float *ms_out, char *names_out, uint32_t names_cap, int mode) {
    int rc = validate_search(ix, n_q, metric, alpha, k);
    if (rc != SG_OK) return rc;
    if (n_q == 0 || !ms_out) return fail(SG_ERR_INVALID, "empty batch or null output");
    if (!d_q_off || !d_out_ids || !d_out_scores || !d_out_counts) return fail(SG_ERR_INVALID, "null buffer");
    DeviceGuard guard;
    SG_CUDA(guard.set(ix->device));
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    for (auto &e : ev) SG_CUDA(cudaEventCreate(&e));
    const uint32_t slot = ix->work_rr.fetch_add(1, std::memory_order_relaxed) & (kWorkRing - 1);
    void *plans = nullptr;
    const size_t plan_bytes = ((size_t)n_q * ix->plan_stride + 255) & ~(size_t)255;
    SG_CUDA(cudaMalloc(&plans, plan_bytes + ix->wtab_bytes));
    rc = enqueue_search(ix, d_q_bytes, d_q_off, n_q, metric, alpha, k, d_out_ids, d_out_scores, d_out_counts, nullptr,
                        ix->work_ring + (size_t)slot * sg::kWorkWords, (uint8_t *)plans, (uint8_t *)plans + plan_bytes, st, mode, ev);
    cudaError_t se = cudaStreamSynchronize(st);
    if (env_int("SG_TRACE", 0) >= 1 && ix->bitmap_engine && se == cudaSuccess) {  // counters of the launch (sg_device.h: kWork*)
        uint32_t w[sg::kWorkWords] = {0};
        cudaMemcpy(w, ix->work_ring + (size_t)slot * sg::kWorkWords, sizeof(w), cudaMemcpyDeviceToHost);
        std::fprintf(stderr, "sg_search_stage_times: %u queries: flagged bitmap words %u, survivor nodes %u, dirty %u, fallback queries taken %u\n", n_q,
                     w[sg::kWorkFlagCursor], w[sg::kWorkNodeCursor], w[sg::kWorkDirtyAny], w[sg::kWorkFallbackQuery]);
    }
    const bool lean = ix->lean_pipeline && mode == 0 && k <= sg::kSmemTopK;  // (what enqueue_search runs the pipeline for)
    const int n_stages = !ix->bitmap_engine ? 2 : lean ? 5 : 3;
    if (rc == SG_OK && se == cudaSuccess)
        for (int i = 0; i < n_stages; i++) cudaEventElapsedTime(ms_out + i, ev[i], ev[i + 1]);
    for (auto &e : ev) cudaEventDestroy(e);
    cudaFree(plans);
    if (rc != SG_OK) return rc;
    if (se != cudaSuccess) return fail(SG_ERR_CUDA, cudaGetErrorString(se));
    if (names_out && names_cap) {
        const char *names = !ix->bitmap_engine ? "sg_plan_kernel,sg_search_kernel"
                            : lean ? (env_int("SG_FUSED_TOKENS", 1) != 0
                                                       ? "sg_window_kernel,(tokenizer fused into the next kernel),sg_tokens_count_kernel,sg_resolve_kernel,sg_bitmap_search_kernel"
                                                       : "sg_window_kernel,sg_tokens_kernel,sg_count_kernel,sg_resolve_kernel,sg_bitmap_search_kernel")
                                                : "sg_window_kernel,sg_tokens_kernel,sg_bitmap_search_kernel";
        std::strncpy(names_out, names, names_cap - 1);
        names_out[names_cap - 1] = 0;
    }
    return n_stages;
}

int sg_merge_topk_device(int device, uint32_t n_parts, uint32_t n_q, uint32_t k, const uint32_t *d_part_ids,
                         const double *d_part_scores, const uint32_t *d_part_counts, uint32_t *d_out_ids,
                         double *d_out_scores, uint32_t *d_out_counts, void *stream) {
    if (n_parts < 1 || n_parts > 32) return fail(SG_ERR_INVALID, "n_parts must be in 1..32");
    if (k < 1 || k > SG_MAX_TOPK_SHARED) return fail(SG_ERR_INVALID, "topK is invalid");
    if (n_q == 0) return SG_OK;
    if (!d_part_ids || !d_part_scores || !d_part_counts || !d_out_ids || !d_out_scores || !d_out_counts)
        return fail(SG_ERR_INVALID, "null buffer");
    DeviceGuard guard;
    SG_CUDA(guard.set(device));
    int blocks = (int)((n_q + 7) / 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    SG_CUDA(sg::launch_merge_topk(n_parts, n_q, k, d_part_ids, d_part_scores, d_part_counts, (size_t)n_q * k, (size_t)n_q * k, n_q,
                                  d_out_ids, d_out_scores, d_out_counts, blocks, (cudaStream_t)stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return SG_OK;
}

uint64_t sg_packed_rows_bytes(uint32_t n_q, uint32_t k) {
    const uint64_t b = (uint64_t)n_q * k * 12 + (uint64_t)n_q * 4;
    return (b + 15) & ~(uint64_t)15;
}

int sg_search_batch_packed_device(sg_index *ix, const char *d_q_bytes, const uint32_t *d_q_off, uint32_t n_q, int metric,
                                  double alpha, uint32_t k, void *d_packed, void *stream) {
    if (!d_packed) return fail(SG_ERR_INVALID, "null buffer");
    double *sc = (double *)d_packed;
    uint32_t *ids = (uint32_t *)(sc + (size_t)n_q * k);
    return sg_search_batch_device(ix, d_q_bytes, d_q_off, n_q, metric, alpha, k, ids, sc, ids + (size_t)n_q * k, nullptr, stream);
}

int sg_merge_topk_packed_device(int device, uint32_t n_parts, uint32_t n_q, uint32_t k, const void *d_parts, uint32_t *d_out_ids,
                                double *d_out_scores, uint32_t *d_out_counts, void *stream) {
    if (n_parts < 1 || n_parts > 32) return fail(SG_ERR_INVALID, "n_parts must be in 1..32");
    if (k < 1 || k > SG_MAX_TOPK_SHARED) return fail(SG_ERR_INVALID, "topK is invalid");
    if (n_q == 0) return SG_OK;
    if (!d_parts || !d_out_ids || !d_out_scores || !d_out_counts) return fail(SG_ERR_INVALID, "null buffer");
    DeviceGuard guard;
    SG_CUDA(guard.set(device));
    const uint64_t stride = sg_packed_rows_bytes(n_q, k);
    const double *sc = (const double *)d_parts;
    const uint32_t *ids = (const uint32_t *)(sc + (size_t)n_q * k);
    const uint32_t *cnt = ids + (size_t)n_q * k;
    int blocks = (int)((n_q + 7) / 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    SG_CUDA(sg::launch_merge_topk(n_parts, n_q, k, ids, sc, cnt, stride / 4, stride / 8, stride / 4, d_out_ids, d_out_scores,
                                  d_out_counts, blocks, (cudaStream_t)stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return SG_OK;
}

// ---------------- record-id-range shards over the GPUs of one box, one host process ----------------
// What suggest_b200/sharding.py does with one process per GPU and an NCCL all-gather, for a host that is a single
// process (the Go service): shard s is an sg_index on devices[s] with id_base = its first document; a call uploads the
// queries once, hands them to the other GPUs over NVLink (peer copies), runs every shard's search concurrently on its
// own stream, and the merge kernel on the first GPU reads the per-shard rows straight out of the other GPUs' HBM
// (peer access) - the exchange is those loads.  Without peer access the blocks are peer-copied and merged locally.
// A batch is cut in two slices so that the merge of the first (its result rows cross PCIe entry by entry when they
// are page-locked host memory: ~160 us for 65,536 queries) and the upload of the second run under the searches.
constexpr uint32_t kShardSlices = 2;

struct ShardCtx {
    sg_index *ix = nullptr;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ready[kShardSlices] = {};  // rows of this shard written, per slice of the batch
    DevBuf<char> q_bytes;
    DevBuf<uint32_t> q_off;
    DevBuf<uint8_t> rows;            // packed blocks of this shard (sg_packed_rows_bytes), one per slice
    // shards 1.. are enqueued by a worker thread each: ~12 runtime calls per shard, serial on one host thread they cost more
    // than the search of a small shard takes (8 shards: the last one's kernels would start ~0.4 ms late)
    std::thread worker;
    std::mutex m;
    std::condition_variable cv;
    int state = 0;                   // 0 idle, 1 job posted, 2 job done, 3 exit
    int rc = SG_OK;
    std::string err;
};

struct ShardJob {                    // one slice of one sg_sharded_search_batch, as the shards see it
    uint32_t k = 0;
    int metric = 0;
    double alpha = 0.0;
    uint32_t slice = 0;
    uint32_t lo = 0, hi = 0;         // queries [lo, hi)
    size_t b0 = 0, b1 = 0;           // their bytes (offsets stay absolute)
    size_t rows_off = 0, block = 0;  // where the slice's packed block starts in ShardCtx::rows, and its size
    size_t parts_off = 0;            // copy path: where the slice's blocks start in sg_sharded::parts
};

struct sg_sharded {
    std::deque<ShardCtx> shards;     // (deque: ShardCtx holds a mutex and a thread and is never moved)
    ShardJob job;
    std::mutex mu;                   // one search at a time per handle
    bool peer_reads = false;         // every shard's HBM is addressable from shards[0].device
    DevBuf<uint8_t> parts;           // copy path: the blocks of all shards, contiguous, on shards[0].device
    DevBuf<uint32_t> out_ids, out_counts;
    DevBuf<double> out_scores;
    const void **d_ptrs = nullptr;   // block pointers for the merge kernel, 32 per slice, on shards[0].device
    cudaEvent_t queries_up[kShardSlices] = {};
    cudaStream_t copy_stream = nullptr, merge_stream = nullptr;  // on shards[0].device
    uint32_t n_docs = 0;
};

static void sharded_destroy(sg_sharded *sx) {
    if (!sx) return;
    for (ShardCtx &sh : sx->shards) {
        if (!sh.worker.joinable()) continue;
        { std::lock_guard<std::mutex> lk(sh.m); sh.state = 3; }
        sh.cv.notify_all();
        sh.worker.join();
    }
    DeviceGuard guard;
    if (!sx->shards.empty()) guard.set(sx->shards[0].device);  // remembers the caller's device; cudaSetDevice from here on
    for (ShardCtx &sh : sx->shards) {
        if (cudaSetDevice(sh.device) == cudaSuccess) {
            if (sh.stream) { cudaStreamSynchronize(sh.stream); cudaStreamDestroy(sh.stream); }
            for (cudaEvent_t ev : sh.ready) if (ev) cudaEventDestroy(ev);
            sh.q_bytes.release(); sh.q_off.release(); sh.rows.release();
        }
        if (sh.ix) sg_index_free(sh.ix);
    }
    if (!sx->shards.empty() && cudaSetDevice(sx->shards[0].device) == cudaSuccess) {
        sx->parts.release(); sx->out_ids.release(); sx->out_counts.release(); sx->out_scores.release();
        if (sx->d_ptrs) cudaFree((void *)sx->d_ptrs);
        for (cudaEvent_t ev : sx->queries_up) if (ev) cudaEventDestroy(ev);
        if (sx->copy_stream) { cudaStreamSynchronize(sx->copy_stream); cudaStreamDestroy(sx->copy_stream); }
        if (sx->merge_stream) { cudaStreamSynchronize(sx->merge_stream); cudaStreamDestroy(sx->merge_stream); }
    }
    cudaGetLastError();  // a device that does not exist (failed build) must not leave its error for the next launch check
    delete sx;
}

// Everything shard s does for one call, enqueued on its stream: wait for the queries on the first GPU, fetch them over
// NVLink, search, (copy path: push the rows to the first GPU,) record `ready`.  Runs on the shard's worker thread
// (shard 0: on the caller's).  Errors come back through sh.rc / sh.err (sg_last_error is thread-local).
static int shard_enqueue(sg_sharded *sx, uint32_t s) {
    ShardCtx &sh = sx->shards[s];
    ShardCtx &s0 = sx->shards[0];
    const ShardJob &j = sx->job;
    const uint32_t n = j.hi - j.lo;
    SG_CUDA(cudaSetDevice(sh.device));
    const char *d_q = s0.q_bytes.p;                      // base pointer: the offsets of a slice stay absolute
    const uint32_t *d_off = s0.q_off.p + j.lo + j.slice; // slice sl keeps its n + 1 offsets at lo + sl
    SG_CUDA(cudaStreamWaitEvent(sh.stream, sx->queries_up[j.slice], 0));
    if (sh.device != s0.device) {
        if (j.b1 > j.b0) SG_CUDA(cudaMemcpyPeerAsync(sh.q_bytes.p + j.b0, sh.device, s0.q_bytes.p + j.b0, s0.device, j.b1 - j.b0, sh.stream));
        SG_CUDA(cudaMemcpyPeerAsync(sh.q_off.p + j.lo + j.slice, sh.device, d_off, s0.device, ((size_t)n + 1) * sizeof(uint32_t), sh.stream));
        d_q = sh.q_bytes.p;
        d_off = sh.q_off.p + j.lo + j.slice;
    }
    uint8_t *rows = sh.rows.p + j.rows_off;
    int rc = sg_search_batch_packed_device(sh.ix, d_q, d_off, n, j.metric, j.alpha, j.k, rows, sh.stream);
    if (rc != SG_OK) return rc;
    if (!sx->peer_reads)
        SG_CUDA(cudaMemcpyPeerAsync(sx->parts.p + j.parts_off + (size_t)s * j.block, s0.device, rows, sh.device, j.block, sh.stream));
    SG_CUDA(cudaEventRecord(sh.ready[j.slice], sh.stream));
    return SG_OK;
}

static void shard_worker(sg_sharded *sx, uint32_t s) {
    ShardCtx &sh = sx->shards[s];
    cudaSetDevice(sh.device);
    for (;;) {
        std::unique_lock<std::mutex> lk(sh.m);
        sh.cv.wait(lk, [&] { return sh.state == 1 || sh.state == 3; });
        if (sh.state == 3) return;
        lk.unlock();
        const int rc = shard_enqueue(sx, s);
        lk.lock();
        sh.rc = rc;
        sh.err = rc == SG_OK ? std::string() : g_err;
        sh.state = 2;
        lk.unlock();
        sh.cv.notify_all();
    }
}

int sg_sharded_build(const sg_config *cfg, const char *doc_bytes, const uint64_t *doc_off, uint32_t n_docs, const int32_t *devices,
                     uint32_t n_shards, sg_sharded **out) {
    if (!out) return fail(SG_ERR_INVALID, "null out");
    *out = nullptr;
    if (!cfg || !devices || (n_docs && !doc_off)) return fail(SG_ERR_INVALID, "null argument");
    if (n_shards < 1 || n_shards > 32) return fail(SG_ERR_INVALID, "n_shards must be in 1..32");
    sg_sharded *sx = new (std::nothrow) sg_sharded();
    if (!sx) return fail(SG_ERR_NOMEM, "out of host memory");
    for (uint32_t s = 0; s < n_shards; s++) sx->shards.emplace_back();
    sx->n_docs = n_docs;
    DeviceGuard guard;
    if (guard.set(devices[0]) != cudaSuccess) cudaGetLastError();  // remembers the caller's device; sg_index_build reports a bad ordinal
    std::vector<uint64_t> sub_off;
    for (uint32_t s = 0; s < n_shards; s++) {
        const uint32_t lo = (uint32_t)((uint64_t)n_docs * s / n_shards), hi = (uint32_t)((uint64_t)n_docs * (s + 1) / n_shards);
        sub_off.resize((size_t)(hi - lo) + 1);
        const uint64_t base = n_docs ? doc_off[lo] : 0;
        for (uint32_t i = lo; i <= hi && n_docs; i++) sub_off[i - lo] = doc_off[i] - base;
        if (!n_docs) sub_off[0] = 0;
        sg_config c = *cfg;
        c.device = devices[s];
        ShardCtx &sh = sx->shards[s];
        sh.device = devices[s];
        int rc = sg_index_build(&c, doc_bytes ? doc_bytes + base : nullptr, sub_off.data(), hi - lo, lo, &sh.ix);
        if (rc != SG_OK) { const std::string msg = g_err; sharded_destroy(sx); return fail(rc, "shard " + std::to_string(s) + ": " + msg); }
        cudaError_t e = cudaSetDevice(sh.device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&sh.stream, cudaStreamNonBlocking);
        for (cudaEvent_t &ev : sh.ready)
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (e != cudaSuccess) { sharded_destroy(sx); return fail(SG_ERR_CUDA, cudaGetErrorString(e)); }
    }
    // peer access from the merging GPU to every other shard's GPU (NVLink / NVSwitch on a B200 box)
    const int dev0 = sx->shards[0].device;
    bool peer = env_int("SG_SHARD_GATHER_COPY", 0) == 0;
    cudaError_t e = cudaSetDevice(dev0);
    for (uint32_t s = 1; s < n_shards && peer && e == cudaSuccess; s++) {
        const int d = sx->shards[s].device;
        if (d == dev0) continue;
        int can = 0;
        e = cudaDeviceCanAccessPeer(&can, dev0, d);
        if (e != cudaSuccess || !can) { peer = false; break; }
        cudaError_t pe = cudaDeviceEnablePeerAccess(d, 0);
        if (pe == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); pe = cudaSuccess; }
        if (pe != cudaSuccess) { cudaGetLastError(); peer = false; }
    }
    for (cudaEvent_t &ev : sx->queries_up)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&sx->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&sx->merge_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void **)&sx->d_ptrs, 32 * kShardSlices * sizeof(void *));
    if (e != cudaSuccess) { sharded_destroy(sx); return fail(SG_ERR_CUDA, cudaGetErrorString(e)); }
    sx->peer_reads = peer;
    for (uint32_t s = 1; s < n_shards; s++) sx->shards[s].worker = std::thread(shard_worker, sx, s);
    *out = sx;
    return SG_OK;
}

void sg_sharded_free(sg_sharded *sx) { sharded_destroy(sx); }

int sg_sharded_get_info(const sg_sharded *sx, uint32_t *n_shards, uint32_t *n_docs, int32_t *peer_reads) {
    if (!sx) return fail(SG_ERR_INVALID, "null handle");
    if (n_shards) *n_shards = (uint32_t)sx->shards.size();
    if (n_docs) *n_docs = sx->n_docs;
    if (peer_reads) *peer_reads = sx->peer_reads ? 1 : 0;
    return SG_OK;
}

sg_index *sg_sharded_shard(const sg_sharded *sx, uint32_t s) {
    return sx && s < sx->shards.size() ? sx->shards[s].ix : nullptr;
}

int sg_sharded_search_batch(sg_sharded *sx, const char *q_bytes, const uint32_t *q_off, uint32_t n_q, int metric, double alpha,
                            uint32_t k, uint32_t *out_ids, double *out_scores, uint32_t *out_counts) {
    if (!sx) return fail(SG_ERR_INVALID, "null handle");
    int rc = validate_search(sx->shards[0].ix, n_q, metric, alpha, k);
    if (rc != SG_OK) return rc;
    if (k > SG_MAX_TOPK_SHARED) return fail(SG_ERR_INVALID, "topK above 1024 is not served by the shard merge");
    if (n_q == 0) return SG_OK;
    if (!q_off || !out_ids || !out_scores || !out_counts) return fail(SG_ERR_INVALID, "null buffer");
    const uint32_t total_bytes = q_off[n_q];
    if (total_bytes && !q_bytes) return fail(SG_ERR_INVALID, "null query bytes");
    std::lock_guard<std::mutex> lock(sx->mu);
    LoweredQueries lq(q_bytes, q_off, n_q);
    const char *src_bytes = lq.bytes;
    const uint32_t *src_off = lq.off;
    const size_t n_bytes = lq.n_bytes;
    const uint32_t n = (uint32_t)sx->shards.size();
    DeviceGuard guard;
    ShardCtx &s0 = sx->shards[0];
    SG_CUDA(guard.set(s0.device));
    // slices of the batch: [0, split) and [split, n_q)
    static const int split_pc = env_int("SG_SHARD_SPLIT", 50);
    uint32_t bounds[kShardSlices + 1] = {0, n_q, n_q};
    if (n_q >= 16384 && split_pc > 0 && split_pc < 100) bounds[1] = (uint32_t)((uint64_t)n_q * split_pc / 100);
    size_t rows_off[kShardSlices + 1] = {0}, blocks_[kShardSlices] = {0};
    for (uint32_t sl = 0; sl < kShardSlices; sl++) {
        blocks_[sl] = bounds[sl + 1] > bounds[sl] ? (size_t)sg_packed_rows_bytes(bounds[sl + 1] - bounds[sl], k) : 0;
        rows_off[sl + 1] = rows_off[sl] + blocks_[sl];
    }
    // page-locked result buffers (sg_pinned_alloc): the merge kernel stores the valid entries straight into them
    uint32_t *m_ids = (uint32_t *)mapped_host_range(out_ids, (size_t)n_q * k * sizeof(uint32_t));
    double *m_scores = (double *)mapped_host_range(out_scores, (size_t)n_q * k * sizeof(double));
    uint32_t *m_counts = (uint32_t *)mapped_host_range(out_counts, (size_t)n_q * sizeof(uint32_t));
    const bool direct = m_ids && m_scores && m_counts;
    if (!direct) {
        SG_CUDA(sx->out_ids.reserve((size_t)n_q * k));
        SG_CUDA(sx->out_scores.reserve((size_t)n_q * k));
        SG_CUDA(sx->out_counts.reserve(n_q));
        m_ids = sx->out_ids.p;
        m_scores = sx->out_scores.p;
        m_counts = sx->out_counts.p;
    }
    // every buffer a worker touches is sized here, before any job is posted
    for (uint32_t s = 0; s < n; s++) {
        ShardCtx &sh = sx->shards[s];
        SG_CUDA(cudaSetDevice(sh.device));
        SG_CUDA(sh.rows.reserve(rows_off[kShardSlices]));
        if (s == 0 || sh.device != s0.device) {
            SG_CUDA(sh.q_bytes.reserve(n_bytes + 64));
            SG_CUDA(sh.q_off.reserve((size_t)n_q + kShardSlices + 1));
        }
    }
    SG_CUDA(cudaSetDevice(s0.device));
    if (!sx->peer_reads) SG_CUDA(sx->parts.reserve(rows_off[kShardSlices] * n));
    // SG_TRACE: timeline of the call on the first GPU and the host's enqueue time
    static const int trace = env_int("SG_TRACE", 0);
    cudaEvent_t tev[2 + 2 * kShardSlices] = {};
    const auto t_begin = std::chrono::steady_clock::now();
    if (trace) {
        for (auto &e : tev) SG_CUDA(cudaEventCreate(&e));
        cudaEventRecord(tev[0], sx->copy_stream);
    }
    const void *ptrs[32 * kShardSlices] = {nullptr};
    std::string msg;
    // one cleanup path: a failure anywhere in here (CUDA call or shard) drains every stream below before returning
    auto enqueue_slices = [&]() -> int {
    for (uint32_t sl = 0; sl < kShardSlices && rc == SG_OK; sl++) {
        const uint32_t lo = bounds[sl], hi = bounds[sl + 1];
        if (hi == lo) continue;
        const size_t b0 = src_off[lo], b1 = src_off[hi];
        // queries of the slice: host -> first GPU on its copy stream (the second slice travels under the first one's search)
        if (b1 > b0) SG_CUDA(cudaMemcpyAsync(s0.q_bytes.p + b0, src_bytes + b0, b1 - b0, cudaMemcpyHostToDevice, sx->copy_stream));
        SG_CUDA(cudaMemcpyAsync(s0.q_off.p + lo + sl, src_off + lo, ((size_t)(hi - lo) + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                sx->copy_stream));
        SG_CUDA(cudaEventRecord(sx->queries_up[sl], sx->copy_stream));
        // every shard enqueues its own work: shards 1.. on their worker threads, shard 0 here
        ShardJob &j = sx->job;
        j.k = k; j.metric = metric; j.alpha = alpha; j.slice = sl; j.lo = lo; j.hi = hi; j.b0 = b0; j.b1 = b1;
        j.rows_off = rows_off[sl]; j.block = blocks_[sl]; j.parts_off = rows_off[sl] * n;
        for (uint32_t s = 1; s < n; s++) {
            ShardCtx &sh = sx->shards[s];
            { std::lock_guard<std::mutex> lk(sh.m); sh.state = 1; }
            sh.cv.notify_all();
        }
        rc = shard_enqueue(sx, 0);
        if (rc != SG_OK) msg = g_err;
        for (uint32_t s = 1; s < n; s++) {
            ShardCtx &sh = sx->shards[s];
            std::unique_lock<std::mutex> lk(sh.m);
            sh.cv.wait(lk, [&] { return sh.state == 2; });
            sh.state = 0;
            if (sh.rc != SG_OK && rc == SG_OK) { rc = sh.rc; msg = sh.err; }
        }
        if (rc != SG_OK) break;
        // merge of the slice on the first GPU's merge stream, behind every shard's search of it; the pointers name the
        // shards' own rows (peer reads) or their copies on this GPU
        SG_CUDA(cudaSetDevice(s0.device));
        for (uint32_t s = 0; s < n; s++) {
            SG_CUDA(cudaStreamWaitEvent(sx->merge_stream, sx->shards[s].ready[sl], 0));
            ptrs[32 * sl + s] = sx->peer_reads ? (const void *)(sx->shards[s].rows.p + rows_off[sl])
                                               : (const void *)(sx->parts.p + rows_off[sl] * n + (size_t)s * blocks_[sl]);
        }
        if (trace) cudaEventRecord(tev[1 + 2 * sl], sx->merge_stream);
        SG_CUDA(cudaMemcpyAsync((void *)(sx->d_ptrs + 32 * sl), ptrs + 32 * sl, n * sizeof(void *), cudaMemcpyHostToDevice, sx->merge_stream));
        int blocks = (int)(((hi - lo + 31) / 32 + 7) / 8);
        if (blocks > 148 * 8) blocks = 148 * 8;
        SG_CUDA(sg::launch_merge_topk_peer(n, hi - lo, k, sx->d_ptrs + 32 * sl, m_ids + (size_t)lo * k, m_scores + (size_t)lo * k, m_counts + lo,
                                           blocks, sx->merge_stream, direct ? 1 : 0));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (!direct) {
            SG_CUDA(cudaMemcpyAsync(out_ids + (size_t)lo * k, m_ids + (size_t)lo * k, (size_t)(hi - lo) * k * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                    sx->merge_stream));
            SG_CUDA(cudaMemcpyAsync(out_scores + (size_t)lo * k, m_scores + (size_t)lo * k, (size_t)(hi - lo) * k * sizeof(double),
                                    cudaMemcpyDeviceToHost, sx->merge_stream));
            SG_CUDA(cudaMemcpyAsync(out_counts + lo, m_counts + lo, (size_t)(hi - lo) * sizeof(uint32_t), cudaMemcpyDeviceToHost, sx->merge_stream));
        }
        if (trace) cudaEventRecord(tev[2 + 2 * sl], sx->merge_stream);
    }
    return rc;
    };
    {
        const int rc2 = enqueue_slices();
        if (rc == SG_OK && rc2 != SG_OK) { rc = rc2; msg = g_err; }  // a CUDA call of the loop itself
    }
    if (rc != SG_OK) {
        for (ShardCtx &sh : sx->shards) { cudaSetDevice(sh.device); cudaStreamSynchronize(sh.stream); }
        cudaSetDevice(s0.device);
        cudaStreamSynchronize(sx->copy_stream);
        cudaStreamSynchronize(sx->merge_stream);
        if (trace) for (auto &e : tev) if (e) cudaEventDestroy(e);
        return fail(rc, msg);
    }
    const auto t_enqueued = std::chrono::steady_clock::now();
    SG_CUDA(cudaStreamSynchronize(sx->merge_stream));   // behind every shard's stream (events) and the copy stream
    if (trace) {
        std::fprintf(stderr, "sg_sharded_search_batch: %u queries, %u shards (%s), rows %s: host enqueue %.1f us, wait %.1f us; first GPU, us from the start:",
                     n_q, n, sx->peer_reads ? "peer reads" : "peer copies", direct ? "page-locked" : "staged + D2H",
                     std::chrono::duration<double, std::micro>(t_enqueued - t_begin).count(),
                     std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_enqueued).count());
        for (uint32_t sl = 0; sl < kShardSlices; sl++) {
            if (bounds[sl + 1] == bounds[sl]) continue;
            float t0 = 0, t1 = 0;
            cudaEventElapsedTime(&t0, tev[0], tev[1 + 2 * sl]);
            cudaEventElapsedTime(&t1, tev[0], tev[2 + 2 * sl]);
            std::fprintf(stderr, " slice %u searched by all at %.1f, merged at %.1f;", sl, t0 * 1e3, t1 * 1e3);
        }
        std::fprintf(stderr, "\n");
        for (auto &e : tev) cudaEventDestroy(e);
    }
    for (uint32_t q = 0; q < n_q; q++)
        if (out_counts[q] == SG_COUNT_UNSUPPORTED)
            return fail(SG_ERR_QUERY_TOO_LONG, "query " + std::to_string(q) + " has more than 128 n-grams");
    return SG_OK;
}

// ---------------- language model and spellchecker (pkg/lm, pkg/spellchecker) ----------------
struct sg_lm {
    sg::DevLm dev{};
    int device = 0;
    std::vector<void *> allocations;
};

static void lm_destroy(sg_lm *lm) {
    if (!lm) return;
    DeviceGuard guard;
    guard.set(lm->device);
    for (void *p : lm->allocations) cudaFree(p);
    delete lm;
}

int sg_lm_create(uint32_t order, const uint64_t *const *containers, const uint64_t *n_containers, const uint64_t *const *values,
                 const uint64_t *n_values, const uint32_t *totals, int device, sg_lm **out) {
    if (!out || !containers || !n_containers || !values || !n_values || !totals) return fail(SG_ERR_INVALID, "null argument");
    if (order < 1 || order > (uint32_t)sg::kMaxLmOrder) return fail(SG_ERR_UNSUPPORTED, "nGramOrder must be in 1..8");
    int n_dev = 0;
    SG_CUDA(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) return fail(SG_ERR_INVALID, "no such CUDA device");
    sg_lm *lm = new (std::nothrow) sg_lm();
    if (!lm) return fail(SG_ERR_NOMEM, "out of host memory");
    lm->device = device;
    lm->dev.order = order;
    DeviceGuard guard;
    cudaError_t e = guard.set(device);
    for (uint32_t i = 0; i < order && e == cudaSuccess; i++) {
        if (n_containers[i] > 0xFFFFFFF0ull || n_values[i] > 0xFFFFFFF0ull) { lm_destroy(lm); return fail(SG_ERR_UNSUPPORTED, "level exceeds 2^32 entries"); }
        lm->dev.n_containers[i] = (uint32_t)n_containers[i];
        lm->dev.n_values[i] = (uint32_t)n_values[i];
        lm->dev.totals[i] = totals[i];
        for (int which = 0; which < 2 && e == cudaSuccess; which++) {
            const uint64_t *src = which ? values[i] : containers[i];
            const size_t n = which ? n_values[i] : n_containers[i];
            void *d = nullptr;
            e = cudaMalloc(&d, (n ? n : 1) * sizeof(uint64_t));
            if (e != cudaSuccess) break;
            lm->allocations.push_back(d);
            if (n) e = cudaMemcpy(d, src, n * sizeof(uint64_t), cudaMemcpyHostToDevice);
            (which ? lm->dev.values[i] : lm->dev.containers[i]) = (const uint64_t *)d;
        }
    }
    if (e != cudaSuccess) {
        lm_destroy(lm);
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? SG_ERR_NOMEM : SG_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = lm;
    return SG_OK;
}

int sg_lm_open(const char *path, int device, sg_lm **out) {
    // nGramModel.Load, pkg/lm/ngram_model.go:126-160 + packedArray.Load, packed_array.go:124-160; whatever follows the model
    // in the file (the MPH table, binary.go:47-49) is not read
    if (!path || !out) return fail(SG_ERR_INVALID, "null argument");
    FILE *f = std::fopen(path, "rb");
    if (!f) return fail(SG_ERR_IO, std::string("io: cannot open ") + path);
    std::vector<unsigned char> data;
    unsigned char buf[1 << 16];
    for (size_t n; (n = std::fread(buf, 1, sizeof(buf), f)) > 0;) data.insert(data.end(), buf, buf + n);
    std::fclose(f);
    if (data.size() < 6 || std::memcmp(data.data(), "0.0.2", 5) != 0) return fail(SG_ERR_FORMAT, "Version mismatch, expected 0.0.2");
    const uint32_t order = data[5];
    if (order < 1 || order > (uint32_t)sg::kMaxLmOrder) return fail(SG_ERR_UNSUPPORTED, "nGramOrder must be in 1..8");
    size_t p = 6;
    std::vector<std::vector<uint64_t>> cont(order), vals(order);
    std::vector<const uint64_t *> cp(order), vp(order);
    std::vector<uint64_t> nc(order), nv(order);
    std::vector<uint32_t> totals(order);
    for (uint32_t i = 0; i < order; i++) {
        unsigned long long cs = 0, vs = 0, total = 0;
        size_t nl = p;
        while (nl < data.size() && data[nl] != '\n') nl++;
        if (nl >= data.size()) return fail(SG_ERR_FORMAT, "language model file is truncated");
        std::string line((const char *)data.data() + p, nl - p);
        if (std::sscanf(line.c_str(), "%llu %llu %llu", &cs, &vs, &total) != 3 || cs % 8 || vs % 8) return fail(SG_ERR_FORMAT, "bad level header");
        p = nl + 1;
        if (p > data.size() || cs > data.size() - p || vs > data.size() - p - cs) return fail(SG_ERR_FORMAT, "language model file is truncated");
        cont[i].resize(cs / 8);
        vals[i].resize(vs / 8);
        if (cs) std::memcpy(cont[i].data(), data.data() + p, cs);
        if (vs) std::memcpy(vals[i].data(), data.data() + p + cs, vs);
        p += cs + vs;
        cp[i] = cont[i].data(); vp[i] = vals[i].data(); nc[i] = cont[i].size(); nv[i] = vals[i].size(); totals[i] = (uint32_t)total;
    }
    return sg_lm_create(order, cp.data(), nc.data(), vp.data(), nv.data(), totals.data(), device, out);
}

void sg_lm_free(sg_lm *lm) { lm_destroy(lm); }

namespace {
// small RAII device scratch for the LM entry points (they are host-buffer calls: copy in, run, copy out); stream-ordered
// allocations from the device's default pool, whose release threshold finish_setup() raises, so repeated calls reuse them
struct Scratch {
    cudaStream_t st = nullptr;
    std::vector<void *> ptrs;
    explicit Scratch(cudaStream_t s = nullptr) : st(s) {}
    cudaError_t get(void **p, size_t bytes) {
        cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 1, st);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
    ~Scratch() { for (void *p : ptrs) cudaFreeAsync(p, st); }
};
}  // namespace

int sg_lm_score_batch(sg_lm *lm, const uint32_t *ids, const uint32_t *off, uint32_t n, double *out_scores) {
    if (!lm || !off || !out_scores) return fail(SG_ERR_INVALID, "null argument");
    if (n == 0) return SG_OK;
    if (off[n] && !ids) return fail(SG_ERR_INVALID, "null ids");
    DeviceGuard guard;
    SG_CUDA(guard.set(lm->device));
    Scratch s;
    uint32_t *d_ids, *d_off;
    double *d_out;
    SG_CUDA(s.get((void **)&d_ids, (size_t)off[n] * 4));
    SG_CUDA(s.get((void **)&d_off, ((size_t)n + 1) * 4));
    SG_CUDA(s.get((void **)&d_out, (size_t)n * 8));
    if (off[n]) SG_CUDA(cudaMemcpy(d_ids, ids, (size_t)off[n] * 4, cudaMemcpyHostToDevice));
    SG_CUDA(cudaMemcpy(d_off, off, ((size_t)n + 1) * 4, cudaMemcpyHostToDevice));
    SG_CUDA(sg::launch_lm_score(lm->dev, d_ids, d_off, n, d_out, nullptr));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    SG_CUDA(cudaMemcpy(out_scores, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return SG_OK;
}

int sg_lm_score_next_batch(sg_lm *lm, const uint32_t *ctx_ids, const uint32_t *ctx_off, uint32_t n_q, const uint32_t *cand_ids,
                           const uint32_t *cand_off, double *out_scores, uint8_t *out_has_scorer) {
    if (!lm || !ctx_off || !cand_off || !out_scores) return fail(SG_ERR_INVALID, "null argument");
    if (n_q == 0) return SG_OK;
    DeviceGuard guard;
    SG_CUDA(guard.set(lm->device));
    Scratch s;
    uint32_t *d_ctx, *d_ctx_off, *d_cand, *d_cand_off;
    double *d_out;
    sg::LmContext *d_lc;
    const size_t n_ctx = ctx_off[n_q], n_cand = cand_off[n_q];
    SG_CUDA(s.get((void **)&d_ctx, n_ctx * 4));
    SG_CUDA(s.get((void **)&d_ctx_off, ((size_t)n_q + 1) * 4));
    SG_CUDA(s.get((void **)&d_cand, n_cand * 4));
    SG_CUDA(s.get((void **)&d_cand_off, ((size_t)n_q + 1) * 4));
    SG_CUDA(s.get((void **)&d_out, n_cand * 8));
    SG_CUDA(s.get((void **)&d_lc, (size_t)n_q * sizeof(sg::LmContext)));
    if (n_ctx) SG_CUDA(cudaMemcpy(d_ctx, ctx_ids, n_ctx * 4, cudaMemcpyHostToDevice));
    SG_CUDA(cudaMemcpy(d_ctx_off, ctx_off, ((size_t)n_q + 1) * 4, cudaMemcpyHostToDevice));
    if (n_cand) SG_CUDA(cudaMemcpy(d_cand, cand_ids, n_cand * 4, cudaMemcpyHostToDevice));
    SG_CUDA(cudaMemcpy(d_cand_off, cand_off, ((size_t)n_q + 1) * 4, cudaMemcpyHostToDevice));
    SG_CUDA(sg::launch_lm_context(lm->dev, d_ctx, d_ctx_off, n_q, d_lc, nullptr));
    SG_CUDA(sg::launch_lm_score_next(d_lc, d_cand, d_cand_off, n_q, d_out, nullptr));
    g_launches.fetch_add(2, std::memory_order_relaxed);
    if (n_cand) SG_CUDA(cudaMemcpy(out_scores, d_out, n_cand * 8, cudaMemcpyDeviceToHost));
    if (out_has_scorer) {
        std::vector<sg::LmContext> lc(n_q);
        SG_CUDA(cudaMemcpy(lc.data(), d_lc, (size_t)n_q * sizeof(sg::LmContext), cudaMemcpyDeviceToHost));
        for (uint32_t q = 0; q < n_q; q++) out_has_scorer[q] = (uint8_t)lc[q].valid;
    }
    return SG_OK;
}

int sg_predict_batch(sg_index *ix, sg_lm *lm, const char *w_bytes, const uint32_t *w_off, const uint32_t *ctx_ids,
                     const uint32_t *ctx_off, uint32_t n_q, double similarity, uint32_t k, uint32_t *out_ids, uint32_t *out_counts) {
    int rc = validate_search(ix, n_q, SG_COSINE, similarity, k);
    if (rc != SG_OK) return rc;
    if (k > SG_MAX_TOPK_SHARED) return fail(SG_ERR_INVALID, "topK above 1024 is not served by sg_predict_batch");
    if (!lm || !w_off || !ctx_off || !out_ids || !out_counts) return fail(SG_ERR_INVALID, "null argument");
    if (n_q == 0) return SG_OK;
    if (!ix->bitmap_engine) return fail(SG_ERR_UNSUPPORTED, "sg_predict_batch needs the bitmap engine");
    if (lm->device != ix->device) return fail(SG_ERR_INVALID, "index and language model live on different devices");
    // the last words: strings.ToLower for non-ASCII input, as sg_search_batch does
    std::string low;
    std::vector<uint32_t> low_off;
    const char *src = w_bytes;
    const uint32_t *src_off = w_off;
    unsigned char high = 0;
    for (uint32_t i = 0; i < w_off[n_q]; i++) high |= (unsigned char)w_bytes[i];
    if (high & 0x80) {
        low_off.resize((size_t)n_q + 1);
        for (uint32_t q = 0; q < n_q; q++) {
            low_off[q] = (uint32_t)low.size();
            sg::to_lower((const uint8_t *)w_bytes + w_off[q], w_off[q + 1] - w_off[q], &low);
        }
        low_off[n_q] = (uint32_t)low.size();
        src = low.data();
        src_off = low_off.data();
    }
    DeviceGuard guard;
    SG_CUDA(guard.set(ix->device));
    CtxLease lease(ix);
    rc = lease.acquire();
    if (rc != SG_OK) return rc;
    cudaStream_t st = lease.ctx->stream;
    Scratch s(st);
    char *d_w;
    uint32_t *d_w_off, *d_ctx, *d_ctx_off, *d_ac_ids, *d_ac_cnt, *d_fz_ids, *d_fz_cnt, *d_out_ids, *d_out_cnt, *d_tmp, *d_work;
    double *d_sc;
    uint8_t *d_plans, *d_wtab;
    sg::LmContext *d_lc;
    const size_t n_ctx = ctx_off[n_q];
    SG_CUDA(s.get((void **)&d_w, (size_t)src_off[n_q] + 16));
    SG_CUDA(s.get((void **)&d_w_off, ((size_t)n_q + 1) * 4));
    SG_CUDA(s.get((void **)&d_ctx, n_ctx * 4));
    SG_CUDA(s.get((void **)&d_ctx_off, ((size_t)n_q + 1) * 4));
    SG_CUDA(s.get((void **)&d_ac_ids, (size_t)n_q * k * 4));
    SG_CUDA(s.get((void **)&d_ac_cnt, (size_t)n_q * 4));
    SG_CUDA(s.get((void **)&d_fz_ids, (size_t)n_q * k * 4));
    SG_CUDA(s.get((void **)&d_fz_cnt, (size_t)n_q * 4));
    SG_CUDA(s.get((void **)&d_sc, (size_t)n_q * k * 8));
    SG_CUDA(s.get((void **)&d_out_ids, (size_t)n_q * (k + 1) * 4));
    SG_CUDA(s.get((void **)&d_out_cnt, (size_t)n_q * 4));
    SG_CUDA(s.get((void **)&d_tmp, (size_t)n_q * 4 * k * 4));
    SG_CUDA(s.get((void **)&d_work, 2 * sg::kWorkWords * sizeof(uint32_t)));
    SG_CUDA(s.get((void **)&d_plans, (size_t)n_q * ix->plan_stride));
    SG_CUDA(s.get((void **)&d_wtab, ix->wtab_bytes));
    SG_CUDA(s.get((void **)&d_lc, (size_t)n_q * sizeof(sg::LmContext)));
    if (src_off[n_q]) SG_CUDA(cudaMemcpyAsync(d_w, src, src_off[n_q], cudaMemcpyHostToDevice, st));
    SG_CUDA(cudaMemcpyAsync(d_w_off, src_off, ((size_t)n_q + 1) * 4, cudaMemcpyHostToDevice, st));
    if (n_ctx) SG_CUDA(cudaMemcpyAsync(d_ctx, ctx_ids, n_ctx * 4, cudaMemcpyHostToDevice, st));
    SG_CUDA(cudaMemcpyAsync(d_ctx_off, ctx_off, ((size_t)n_q + 1) * 4, cudaMemcpyHostToDevice, st));
    SG_CUDA(sg::launch_lm_context(lm->dev, d_ctx, d_ctx_off, n_q, d_lc, st));
    // completions of the last word, the k best by the model (index.Autocomplete with the lm collector, spellchecker.go:58-60)
    rc = enqueue_search(ix, d_w, d_w_off, n_q, SG_EXACT, 1.0, k, d_ac_ids, d_sc, d_ac_cnt, nullptr, d_work, d_plans, d_wtab, st, 1, nullptr, d_lc);
    // fuzzy candidates (index.Suggest with CosineMetric, spellchecker.go:67-74); searched for every query, used where needed
    if (rc == SG_OK)
        rc = enqueue_search(ix, d_w, d_w_off, n_q, SG_COSINE, similarity, k, d_fz_ids, d_sc, d_fz_cnt, nullptr, d_work + sg::kWorkWords, d_plans, d_wtab, st, 0);
    if (rc != SG_OK) { cudaStreamSynchronize(st); return rc; }
    SG_CUDA(sg::launch_predict_merge(d_lc, n_q, k, d_ac_ids, d_ac_cnt, d_fz_ids, d_fz_cnt, d_out_ids, d_out_cnt, d_tmp, st));
    g_launches.fetch_add(2, std::memory_order_relaxed);
    SG_CUDA(cudaMemcpyAsync(out_ids, d_out_ids, (size_t)n_q * (k + 1) * 4, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(out_counts, d_out_cnt, (size_t)n_q * 4, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    return SG_OK;
}

uint64_t sg_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

const char *sg_last_error(void) { return g_err.c_str(); }

const char *sg_version(void) { return "suggest_b200 0.1 (sm_100a)"; }

}  // extern "C"
