// sg_bitmap.cu — the bitmap engine of libsuggest_b200 (sm_100a): T-occurrence counting as bit-sliced addition.
//
// What the reference does per query with one lazily decoded posting-list iterator per (token, segment) and CPMerge
//   (pkg/suggest/suggester.go:46-131, pkg/index/searcher.go:28-78, pkg/merger/cp_merge.go:19-120)
// is done here on a second representation of the same posting lists: for every term one row of bits, one bit per bucket
// of 2^bshift consecutive documents (documents are numbered by cardinality segment and segments start on bucket
// boundaries, so a bucket belongs to one segment).  A query's segment window [MinY, MaxY] is a contiguous range of
// words of each of its terms' rows.  A warp walks that range 32 words at a time (one word per lane, coalesced 128-byte
// loads per list), adds the words of all lists with carry-save adders into bit planes (bit j of plane i = bit i of
// "how many of the query's lists hit bucket j"), and compares the planes with the threshold of the segment the word
// belongs to.  That count is an upper bound of the overlap of every document in the bucket (ScanCount semantics,
// pkg/merger/scan_count.go:14-88, at bucket granularity), so a bucket below its threshold holds no candidate.
//   * bshift = 0 (small dictionaries): one bit per document, the planes hold the overlap itself;
//   * bshift > 0: the rare bucket that reaches its threshold is resolved exactly from the posting lists of its segment
//     (shared-memory counters, one per document of the bucket).
// Survivors are scored in float64 in the reference's operation order and kept in a per-warp sorted top-k
// (pkg/metric/*.go, pkg/suggest/scorer.go:29-31, collector.go:20-26).  No atomics and no shared-memory table on the
// counting path: every lane owns its words.
//
// Three launches per batch:
//   sg_window_kernel         per len(tokens) = 0..128: segment window, T(segment), T(bitmap word)   (tiny)
//   sg_tokens_kernel         tokenise every query -> len(tokens), term ids
//   sg_bitmap_search_kernel  count, compare, resolve, score, top-k
#include "sg_common.cuh"
#include "sg_kernels.h"

namespace sg {

namespace {

constexpr int kWindowThreads = 128;
constexpr int kBitmapWarps = 8;            // warps per CTA of sg_bitmap_search_kernel
constexpr int kResolveSlots = 1 << kMaxBucketShift;
// per-warp shared memory: [row offset of every list (136 x 4) | term id of every list (128 x 4) | resolve counters (256 x 4) | top-k]
constexpr uint32_t kRowSlots = kMaxQueryTokens + 8;  // padded to a multiple of 8 with the all-zero row
constexpr uint32_t kBitmapWarpFixedSmem = kRowSlots * 4 + kMaxQueryTokens * 4 + kResolveSlots * 4;

// carry-save adder: (h, l) = a + b + c per bit position; two LOP3
__device__ __forceinline__ void csa(uint32_t &h, uint32_t &l, uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t u = a ^ b;
    h = (a & b) | (u & c);
    l = u ^ c;
}

// planes += x (one more list), rippling the carry up
template <int M>
__device__ __forceinline__ void add_one(uint32_t (&c)[M], uint32_t x) {
#pragma unroll
    for (int j = 0; j < M; j++) {
        const uint32_t t = c[j] & x;
        c[j] ^= x;
        x = t;
    }
}

// bits whose M-plane count is >= T (T in 1..255, per lane)
template <int M>
__device__ __forceinline__ uint32_t planes_ge(const uint32_t (&c)[M], uint32_t T) {
    uint32_t gt = 0u, eq = 0xFFFFFFFFu;
#pragma unroll
    for (int j = M - 1; j >= 0; j--) {
        const uint32_t tj = 0u - ((T >> j) & 1u);
        gt |= eq & c[j] & ~tj;
        eq &= ~(c[j] ^ tj);
    }
    return (T >> M) ? 0u : (gt | eq);
}

template <int M>
__device__ __forceinline__ int planes_count(const uint32_t (&c)[M], int bit) {
    int n = 0;
#pragma unroll
    for (int j = 0; j < M; j++) n |= (int)((c[j] >> bit) & 1u) << j;
    return n;
}

// Per-query, warp-uniform state of the search kernel.
struct BitmapQuery {
    QueryCtx c;
    int n_lists;
    int cur_seg;       // segment cursor: flagged buckets arrive in ascending order
    const uint8_t *seg_thr;  // row len(tokens) of WindowTables::seg_thr
};

// segment that owns `bucket` (ascending calls within a query): seg_start[B] <= bucket << bshift < seg_start[B + 1]
__device__ __forceinline__ int segment_of(const DevIndex &ix, BitmapQuery &bq, uint32_t bucket) {
    const uint32_t id = bucket << ix.bshift;
    int B = bq.cur_seg;
    if (__ldg(ix.seg_start + B + 1) <= id) {
        int lo = B + 1, hi = (int)ix.n_segments - 1;  // first segment whose end is above id
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(ix.seg_start + mid + 1) <= id) lo = mid + 1; else hi = mid;
        }
        B = lo;
        bq.cur_seg = B;
    }
    return B;
}

// A bucket of bshift > 0 whose list count reached its segment's threshold: count every document of the bucket exactly.
// Lane l takes lists l, l + 32, ...: the (term, segment) posting list is searched for the bucket's id range and every
// posting inside adds one to its document's counter; then the counters are compared with T and the survivors offered.
__device__ __noinline__ void resolve_bucket_exact(const DevIndex &ix, BitmapQuery &bq, const uint32_t *s_term, uint32_t *s_cnt,
                                                  uint32_t bucket, int B, int T, int lane) {
    const uint32_t width = 1u << ix.bshift;
    const uint32_t id_lo = bucket << ix.bshift, id_hi = id_lo + width;
    for (uint32_t i = lane; i < width; i += 32) s_cnt[i] = 0u;
    __syncwarp();
    const uint32_t *__restrict__ postings = ix.postings;
    const size_t stride = (size_t)ix.n_segments + 1;
    for (int j = lane; j < bq.n_lists; j += 32) {
        const uint32_t *o = ix.list_off + (size_t)s_term[j] * stride + B;
        const uint32_t b = __ldg(o + 1);
        uint32_t pos = lower_bound(postings, __ldg(o), b, id_lo);
        for (; pos < b; pos++) {
            const uint32_t x = __ldg(postings + pos);
            if (x >= id_hi) break;
            atomicAdd(s_cnt + (x - id_lo), 1u);
        }
    }
    __syncwarp();
    for (uint32_t base = 0; base < width; base += 32) {
        const uint32_t v = base + lane < width ? s_cnt[base + lane] : 0u;
        unsigned m = __ballot_sync(kFull, v >= (uint32_t)T);
        while (m) {
            const int i = __ffs(m) - 1;
            m &= m - 1;
            emit_candidate(ix, bq.c, id_lo + base + (uint32_t)i, (int)__shfl_sync(kFull, v, i), B, T, lane);
        }
    }
    __syncwarp();
}

// One tile of 32 bitmap words (lane l owns word w0 + l): add the words of all lists, compare, handle the hits.
template <int M>
__device__ __forceinline__ void search_tile(const DevIndex &ix, BitmapQuery &bq, const uint32_t *s_row, const uint32_t *s_term,
                                            uint32_t *s_cnt, const uint8_t *__restrict__ word_thr, uint32_t w0, WordRange win, int lane) {
    const uint32_t W = w0 + (uint32_t)lane;
    const uint32_t *__restrict__ bm = ix.bitmaps + W;
    const uint32_t T_w = __ldg(word_thr + W);
    uint32_t c[M];
#pragma unroll
    for (int j = 0; j < M; j++) c[j] = 0u;
    // lists in blocks of 8 (the tail is padded with the all-zero row): seven carry-save adders turn eight words into one
    // carry of weight 8, which ripples into the planes above
    for (int j0 = 0; j0 < bq.n_lists; j0 += 8) {
        uint32_t x[8];
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = __ldg(bm + s_row[j0 + i]);
        uint32_t tA, tB, tC, tD, fA, fB, e;
        csa(tA, c[0], c[0], x[0], x[1]);
        csa(tB, c[0], c[0], x[2], x[3]);
        csa(fA, c[1], c[1], tA, tB);
        csa(tC, c[0], c[0], x[4], x[5]);
        csa(tD, c[0], c[0], x[6], x[7]);
        csa(fB, c[1], c[1], tC, tD);
        csa(e, c[2], c[2], fA, fB);
#pragma unroll
        for (int j = 3; j < M; j++) {
            const uint32_t t = c[j] & e;
            c[j] ^= e;
            e = t;
        }
    }
    uint32_t flag = planes_ge<M>(c, T_w);
    if (W < win.x || W >= win.y) flag = 0u;
    unsigned bal;
    while ((bal = __ballot_sync(kFull, flag != 0u)) != 0u) {
        const int src = __ffs(bal) - 1;
        uint32_t f = __shfl_sync(kFull, flag, src);
        if (lane == src) flag = 0u;
        uint32_t cs[M];
        if (ix.bshift == 0) {
#pragma unroll
            for (int j = 0; j < M; j++) cs[j] = __shfl_sync(kFull, c[j], src);
        }
        while (f) {
            const int bit = __ffs(f) - 1;
            f &= f - 1;
            const uint32_t bucket = (w0 + (uint32_t)src) * 32u + (uint32_t)bit;
            const int B = segment_of(ix, bq, bucket);
            const int T = (int)__ldg(bq.seg_thr + B);
            if (T == 0) continue;  // the word's threshold came from a neighbouring segment
            if (ix.bshift == 0) emit_candidate(ix, bq.c, bucket, planes_count<M>(cs, bit), B, T, lane);
            else resolve_bucket_exact(ix, bq, s_term, s_cnt, bucket, B, T, lane);
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// sg_window_kernel: block a handles len(tokens) = a.  suggester.go:53-59 (window), :73-78 (threshold, admissibility).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWindowThreads) sg_window_kernel(const DevIndex ix, const SearchParams p) {
    const int a = (int)blockIdx.x;
    const int S = (int)ix.n_segments;
    const int metric = p.mode == 1 ? (int)kAutocomplete : p.metric;
    uint8_t *seg_thr = p.wt.seg_thr + (size_t)a * S;
    uint8_t *word_thr = p.wt.word_thr + (size_t)a * ix.row_words;
    __shared__ uint32_t s_lo, s_hi;
    if (threadIdx.x == 0) { s_lo = 0xFFFFFFFFu; s_hi = 0u; }
    int b_min = 0, b_max = -1;
    if (a > 0) {
        b_min = max(metric_min_y(metric, p.alpha, a), 0);
        b_max = metric_max_y(metric, p.alpha, a);
        if (b_max >= S) b_max = S - 1;
    }
    for (int B = threadIdx.x; B < S; B += kWindowThreads) {
        int T = 0;
        if (B >= b_min && B <= b_max) {
            T = metric_threshold(metric, p.alpha, a, B);
            if (!threshold_admits(T, a, B) || ix.seg_start[B + 1] <= ix.seg_start[B]) T = 0;
        }
        seg_thr[B] = (uint8_t)T;  // T <= a <= 128
    }
    __syncthreads();
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (uint32_t w = threadIdx.x; w < ix.row_words; w += kWindowThreads) {
        // segments owning new ids of [w * 32 << bshift, (w + 1) * 32 << bshift)
        const uint64_t id0 = (uint64_t)w * 32u << ix.bshift, id1 = (uint64_t)(w + 1) * 32u << ix.bshift;
        int l = 0, h = S;  // first segment whose end is above id0
        while (l < h) {
            const int mid = (l + h) >> 1;
            if ((uint64_t)ix.seg_start[mid + 1] <= id0) l = mid + 1; else h = mid;
        }
        uint32_t m = 255u;
        for (int B = l; B < S && (uint64_t)ix.seg_start[B] < id1; B++) {
            const uint32_t t = seg_thr[B];
            if (t != 0u && t < m) m = t;
        }
        word_thr[w] = (uint8_t)m;
        if (m != 255u) { lo = min(lo, w); hi = max(hi, w + 1); }
    }
    atomicMin(&s_lo, lo);
    atomicMax(&s_hi, hi);
    __syncthreads();
    if (threadIdx.x == 0) p.wt.win[a] = s_hi > s_lo ? WordRange{s_lo, s_hi} : WordRange{0u, 0u};
}

// ---------------------------------------------------------------------------------------------------------------
// sg_tokens_kernel: the tokenizer chain for every query; one warp per query.  With p.stats it also counts the
// admissible postings / lists of SURVEY.md section 8(d) (the algorithmic bytes of the roofline).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPlanThreads) sg_tokens_kernel(const DevIndex ix, const SearchParams p) {
    __shared__ __align__(16) uint32_t s_scratch[kPlanThreads / 32][kMaxRunes + 2 * kMaxQueryTokens];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint32_t *s_runes = s_scratch[warp];
    uint32_t *s_lterm = s_runes + kMaxRunes;
    uint32_t *s_hash = s_lterm + kMaxQueryTokens;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const size_t stride = (size_t)ix.n_segments + 1;
    for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < p.n_q; q += n_warps) {
        int size_a = 0, n_lists = 0;
        const bool unsupported = tokenize_query(ix, p, q, s_runes, s_lterm, s_hash, lane, &size_a, &n_lists);
        if (p.mode == 1 && n_lists < size_a) n_lists = 0;  // a query token that is in no list: nothing can hold them all
        if (unsupported) { size_a = 0; n_lists = 0; }
        uint8_t *plan_base = p.plans + (size_t)q * kTokStride;
        uint32_t *plan_terms = (uint32_t *)(plan_base + kTokTermsOffset);
        for (int j = lane; j < n_lists; j += 32) plan_terms[j] = s_lterm[j];
        if (lane == 0) *(uint4 *)plan_base = make_uint4(unsupported ? 1u : 0u, (uint32_t)size_a, (uint32_t)n_lists, 0u);
        if (p.stats != nullptr) {
            uint32_t st_postings = 0, st_lists = 0;
            const uint8_t *seg_thr = p.wt.seg_thr + (size_t)size_a * ix.n_segments;
            // admissible segments regardless of emptiness: an empty segment has no lists
            const int metric = p.mode == 1 ? (int)kAutocomplete : p.metric;
            int b_min = 0, b_max = -1;
            if (size_a > 0) {
                b_min = max(metric_min_y(metric, p.alpha, size_a), 0);
                b_max = min(metric_max_y(metric, p.alpha, size_a), (int)ix.n_segments - 1);
            }
            for (int B = b_min; B <= b_max; B++) {
                if (seg_thr[B] == 0) continue;
                for (int j = lane; j < n_lists; j += 32) {
                    const uint32_t *o = ix.list_off + (size_t)s_lterm[j] * stride + B;
                    const uint32_t len = __ldg(o + 1) - __ldg(o);
                    st_postings += len;
                    st_lists += len != 0;
                }
            }
            st_postings = __reduce_add_sync(kFull, st_postings);
            st_lists = __reduce_add_sync(kFull, st_lists);
            if (lane == 0) { p.stats[2 * q] = st_postings; p.stats[2 * q + 1] = st_lists; }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// sg_bitmap_search_kernel: one warp per query, query numbers from a global counter.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBitmapWarps * 32) sg_bitmap_search_kernel(const DevIndex ix, const SearchParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t *wsm = smem + (size_t)warp * p.warp_smem;
    uint32_t *s_row = (uint32_t *)wsm;                 // [136] word offset of the bitmap row of every list
    uint32_t *s_term = s_row + kRowSlots;              // [128] term id of every list
    uint32_t *s_cnt = s_term + kMaxQueryTokens;        // [256] per-document counters of the bucket being resolved
    double *tk_score = (double *)(s_cnt + kResolveSlots);  // [k]
    uint32_t *tk_id = (uint32_t *)(tk_score + p.k);    // [k]
    const uint32_t zero_row = ix.n_terms * ix.row_words;

    for (;;) {
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(p.work_counter, 1u);
        q = __shfl_sync(kFull, q, 0);
        if (q >= p.n_q) break;

        const uint8_t *plan_base = p.plans + (size_t)q * kTokStride;
        const uint4 h0 = __ldg((const uint4 *)plan_base);  // TokenPlan
        BitmapQuery bq;
        bq.c.metric = p.mode == 1 ? (int)kAutocomplete : p.metric;
        bq.c.alpha = p.alpha;
        bq.c.k = p.k;
        bq.c.tk_len = 0;
        bq.c.tk_score = tk_score;
        bq.c.tk_id = tk_id;
        bq.c.size_a = (int)h0.y;
        bq.c.b_lo = 0;
        bq.c.b_hi = -1;
        bq.n_lists = (int)h0.z;
        bq.cur_seg = 0;
        bq.seg_thr = p.wt.seg_thr + (size_t)bq.c.size_a * ix.n_segments;
        const bool unsupported = h0.x != 0u;
        const WordRange win = p.wt.win[bq.c.size_a];

        if (bq.n_lists > 0 && win.y > win.x) {
            const int n_pad = (bq.n_lists + 7) & ~7;
            for (int j = lane; j < n_pad; j += 32) {
                uint32_t row = zero_row;
                if (j < bq.n_lists) {
                    const uint32_t t = __ldg((const uint32_t *)(plan_base + kTokTermsOffset) + j);
                    s_term[j] = t;
                    row = t * ix.row_words;
                }
                s_row[j] = row;
            }
            __syncwarp();
            const uint8_t *word_thr = p.wt.word_thr + (size_t)bq.c.size_a * ix.row_words;
            if (bq.n_lists < 32) {
                for (uint32_t w0 = win.x & ~31u; w0 < win.y; w0 += 32) search_tile<5>(ix, bq, s_row, s_term, s_cnt, word_thr, w0, win, lane);
            } else {
                for (uint32_t w0 = win.x & ~31u; w0 < win.y; w0 += 32) search_tile<8>(ix, bq, s_row, s_term, s_cnt, word_thr, w0, win, lane);
            }
        }

        // results: GetCandidates order, fixed stride k
        const size_t row = (size_t)q * p.k;
        for (uint32_t j = lane; j < p.k; j += 32) {
            const bool has = (int)j < bq.c.tk_len;
            p.out_ids[row + j] = has ? ix.id_base + tk_id[j] : 0u;
            p.out_scores[row + j] = has ? tk_score[j] : 0.0;
        }
        if (lane == 0) p.out_counts[q] = unsupported ? kCountUnsupported : (uint32_t)bq.c.tk_len;
        __syncwarp();
    }
}

// ---------------- launcher (host) ----------------
size_t bitmap_warp_smem(uint32_t k) { return ((size_t)kBitmapWarpFixedSmem + (size_t)k * 12u + 15u) & ~(size_t)15u; }

cudaError_t launch_bitmap_search(const DevIndex &ix, const SearchParams &p, int sm_count, cudaStream_t stream) {
    const size_t smem = (size_t)kBitmapWarps * p.warp_smem;
    cudaError_t e = cudaFuncSetAttribute(sg_bitmap_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sg_bitmap_search_kernel, kBitmapWarps * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    sg_window_kernel<<<kWindowRows, kWindowThreads, 0, stream>>>(ix, p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int tok_blocks = (int)((p.n_q + kPlanThreads / 32 - 1) / (kPlanThreads / 32));
    sg_tokens_kernel<<<tok_blocks < sm_count * 8 ? tok_blocks : sm_count * 8, kPlanThreads, 0, stream>>>(ix, p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    int blocks = sm_count * per_sm;
    const int need = (int)((p.n_q + kBitmapWarps - 1) / kBitmapWarps);
    if (blocks > need) blocks = need;
    sg_bitmap_search_kernel<<<blocks, kBitmapWarps * 32, smem, stream>>>(ix, p);
    return cudaGetLastError();
}

}  // namespace sg
