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
// Launches per batch.  Suggest top-k on an index with the exact level (sg_fine.cu) runs the count -> resolve pipeline
// (second half of this file):
//   sg_window_kernel         per len(tokens) = 0..128: segment window, T(segment), T(bitmap word)   (tiny; cached per
//                            (metric, similarity))
//   sg_tokens_count_kernel   tokenise every query, count (two words per lane, 8-byte loads), append the bitmap words in
//                            which a bucket reached its threshold to a launch-wide list
//   sg_resolve_kernel        eight lanes per flagged word: exact per-document counts from the bit-per-document level,
//                            survivors scored and linked to their query, k best selected by whoever resolves a query's
//                            last word
//   sg_bitmap_search_kernel  (only_dirty) the queries the pipeline ran out of scratch for; normally looks at one flag
// Autocomplete, the spellchecker's collector, collect mode, k > 1024 and indexes without the exact level:
//   sg_window_kernel, sg_tokens_kernel (tokenise -> len(tokens), term ids), sg_bitmap_search_kernel (count, compare,
//   resolve from the posting lists, score, top-k - all in the warp that owns the query)
#include <cstddef>
#include <map>
#include <mutex>
#include <utility>

#include "sg_common.cuh"
#include "sg_kernels.h"

namespace sg {

namespace {

constexpr int kWindowThreads = 128;
constexpr int kBitmapWarps = 8;            // warps per CTA of sg_bitmap_search_kernel
#ifndef SG_TOKENS_MIN_BLOCKS
#define SG_TOKENS_MIN_BLOCKS 8          // CTAs per SM of sg_tokens_kernel the register allocation aims at (32 registers: its dependent loads want warps)
#endif
#ifndef SG_BITMAP_MIN_BLOCKS
#define SG_BITMAP_MIN_BLOCKS 4          // CTAs per SM the register allocation aims at (64 registers per thread)
#endif
constexpr int kResolveSlots = 1 << kMaxBucketShift;
constexpr uint32_t kRowSlots = kMaxQueryTokens + 8;  // padded to a multiple of 8 with the all-zero row
constexpr uint32_t kTileWords = 32;        // bitmap words per tile: lane l owns word l
constexpr int kSegCache = 256;             // segment starts kept in shared memory per CTA
constexpr int kResolveThreads = 256;       // sg_resolve_kernel
constexpr int kResolveGroup = 8;           // lanes that resolve one flagged bitmap word together (four words per warp)
constexpr int kResolveLists = 4;           // lists a lane takes per round: 32 lists of a query are one round of loads

// Per-warp shared memory of sg_bitmap_search_kernel.  The count loop itself only reads `row`; everything
// else belongs to the cold path (a bucket reached its threshold), which keeps its state here so that the hot loop's
// registers stay free.
struct WarpSmem {
    int32_t tk_len;                    // candidates in the top-k
    int32_t size_a, n_lists;
    uint32_t lm_valid, lm_from, lm_to; // spellchecker completions: rank by the language model (LmContext of the query)
    const uint64_t *lm_vals;
    uint32_t q;                        // number of the query inside the launch (collect mode)
    uint32_t pad[3];
    double *tk_score;                  // the warp's sorted top-k: behind this struct in shared memory, or - k > kSmemTopK - a slice
    uint32_t *tk_id;                   // of SearchParams::tk_global in HBM
    uint32_t flag[32];                 // per lane: buckets of its word that reached the threshold
    uint32_t bias[32];                 // per lane: 2^M - T(word), what the planes started from
    alignas(16) uint32_t row[kRowSlots];   // word offset of the bitmap row of every list
    uint32_t term[kMaxQueryTokens];    // term id of every list
    uint32_t cnt[kResolveSlots];       // bshift > 0: per-document counters of the bucket being resolved;
                                       // bshift = 0: the planes of every lane, cnt[j * 32 + lane]
};
constexpr uint32_t kBitmapWarpFixedSmem = (uint32_t)sizeof(WarpSmem);
static_assert(sizeof(WarpSmem) % 16 == 0, "top-k scores follow and need 8-byte alignment");
static_assert(offsetof(WarpSmem, tk_score) % 8 == 0, "pointer alignment");

// What the cold path needs of the kernel arguments, copied once per CTA (a noinline callee cannot take the address of
// a kernel parameter without a local copy per thread).
struct BlockConsts {
    const uint32_t *postings, *list_off, *perm, *seg_start;
    const uint8_t *seg_thr;   // WindowTables::seg_thr
    uint32_t n_segments, bshift, id_base, k;
    int32_t metric;
    unsigned long long *cand_total;     // collect mode (sg_candidates_batch): SearchParams::cand_*
    unsigned long long cand_cap;
    uint32_t *cand_query, *cand_ids, *cand_overlap, *cand_segment;
    uint32_t seg_cache[kSegCache + 1];  // seg_start[0 .. min(S, kSegCache)]
};

// next query number.  atom.inc, not atom.add: around an add with a uniform address ptxas puts its warp-aggregation
// shuffle, which waits for the result on the spot; here the result is wanted a whole query later.
__device__ __forceinline__ uint32_t take_query(uint32_t *counter) {
    uint32_t q;
    asm volatile("atom.global.inc.u32 %0, [%1], 0xfffffffe;" : "=r"(q) : "l"(counter) : "memory");
    return q;
}

// carry-save adder: (h, l) = a + b + c per bit position; two LOP3
__device__ __forceinline__ void csa(uint32_t &h, uint32_t &l, uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t u = a ^ b;
    h = (a & b) | (u & c);
    l = u ^ c;
}

// one entry of a result row: separate id / score arrays, or 16-byte {id, 0, score} entries (SearchParams::out_packed)
__device__ __forceinline__ void store_entry(uint32_t *out_ids, double *out_scores, uint4 *out_packed, size_t at, uint32_t id, double score) {
    if (out_packed != nullptr) out_packed[at] = make_uint4(id, 0u, (uint32_t)__double2loint(score), (uint32_t)__double2hiint(score));
    else { out_ids[at] = id; out_scores[at] = score; }
}

__device__ __forceinline__ double *warp_tk_score(WarpSmem *ws) { return ws->tk_score; }
__device__ __forceinline__ uint32_t *warp_tk_id(WarpSmem *ws, uint32_t) { return ws->tk_id; }

// Score of a completion (Autocomplete collectors): FirstKCollectorManager.Collect scores a position with -position
// (pkg/suggest/collector.go:104-106); the spellchecker's lmCollector (pkg/spellchecker/collector.go:61-78) with
// ScoreNext(word) = log(count(context, word) / count(context)), or -100 for an unseen continuation - monotone in the count,
// which is what the queue is ordered by here.  Per lane (no warp-wide operation inside).
__device__ __forceinline__ double completion_score(const BlockConsts *bc, const WarpSmem *ws, uint32_t id) {
    const uint32_t word = bc->id_base + id;
    if (!ws->lm_valid) return -(double)word;
    const uint64_t *__restrict__ v = ws->lm_vals;
    uint32_t lo = ws->lm_from, hi = ws->lm_to;
    const uint64_t target = (uint64_t)word << 32;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(v + mid) < target) lo = mid + 1; else hi = mid;
    }
    uint64_t hit = 0;
    if (lo < ws->lm_to) hit = __ldg(v + lo);
    return (uint32_t)(hit >> 32) == word && lo < ws->lm_to ? (double)(uint32_t)hit : 0.0;
}

// Collect mode (sg_candidates_batch): what the mergers hand to Collector.Collect (pkg/merger/collector.go:10-13), appended
// to the launch-wide list; scoring and selection are the caller's (a CollectorManager / metric.Metric of its own).
// Out of line: its registers (a 64-bit atomic and four pointers) stay out of handle_flags and of the kernel around it.
__device__ __noinline__ void collect_candidate(const BlockConsts *bc, uint32_t q, uint32_t id, int count, int size_b) {
    const unsigned long long at = atomicAdd(bc->cand_total, 1ull);
    if (at < bc->cand_cap) {
        bc->cand_query[at] = q;
        bc->cand_ids[at] = bc->id_base + id;
        bc->cand_overlap[at] = (uint32_t)count;
        bc->cand_segment[at] = (uint32_t)size_b;
    }
}

// A document (new id) of segment size_b with an exact overlap count >= T: score it and offer it to the warp's sorted
// top-k (best first, Candidate.Less of pkg/suggest/collector.go:20-26).  All lanes call with identical arguments.
template <bool kCollect>
__device__ void offer_candidate(const BlockConsts *bc, WarpSmem *ws, uint32_t new_id, int count, int size_b, int lane) {
    const uint32_t id = __ldg(bc->perm + new_id);
    if (kCollect) {
        if (lane == 0) collect_candidate(bc, ws->q, id, count, size_b);
        return;
    }
    const double score = bc->metric != kAutocomplete ? metric_score(bc->metric, count, ws->size_a, size_b) : completion_score(bc, ws, id);
    QueryCtx c;
    c.k = bc->k;
    c.tk_len = ws->tk_len;
    c.tk_score = warp_tk_score(ws);
    c.tk_id = warp_tk_id(ws, bc->k);
    topk_insert(c, score, id, lane);
    __syncwarp();
    if (lane == 0) ws->tk_len = c.tk_len;
    __syncwarp();
}

// segment that owns new id `id`: seg_start[B] <= id < seg_start[B + 1]
__device__ __forceinline__ int segment_of_id(const BlockConsts *bc, uint32_t id) {
    int lo = 0, hi = (int)bc->n_segments - 1;  // first segment whose end is above id
    const int cached = min((int)bc->n_segments, kSegCache);
    if (bc->seg_cache[cached] > id) {
        hi = cached - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (bc->seg_cache[mid + 1] <= id) lo = mid + 1; else hi = mid;
        }
        return lo;
    }
    lo = cached;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(bc->seg_start + mid + 1) <= id) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// first position in [a, b) whose posting is >= x; four probes per step, so a list of 64 takes three dependent loads
__device__ __forceinline__ uint32_t lower_bound4(const uint32_t *__restrict__ postings, uint32_t a, uint32_t b, uint32_t x) {
    while (b - a > 4) {
        const uint32_t q = (b - a) / 5 + 1;
        const uint32_t p1 = a + q - 1, p2 = p1 + q, p3 = p2 + q, p4 = min(p3 + q, b - 1);
        const uint32_t v1 = __ldg(postings + p1), v2 = __ldg(postings + p2), v3 = __ldg(postings + min(p3, b - 1)), v4 = __ldg(postings + p4);
        if (v1 >= x) b = p1;
        else if (v2 >= x) { a = p1 + 1; b = p2; }
        else if (p3 >= b || v3 >= x) { a = p2 + 1; b = min(p3, b); }
        else if (v4 >= x) { a = p3 + 1; b = p4; }
        else a = p4 + 1;
    }
    while (a < b && __ldg(postings + a) < x) a++;
    return a;
}

// Buckets that reached the threshold of their word (ws->flag, per lane) in the tile starting at word w0; lane l's flags
// are about word w0 + l.  For every such bucket: find its segment and that segment's own threshold; bshift = 0:
// the planes give the overlap (ws->cnt, ws->bias); bshift > 0: count every document of the bucket exactly - lane l takes
// lists l, l + 32, ..., searches the (term, segment) posting list for the bucket's id range and adds one to the counter
// of every document found - then offer the survivors to the top-k.
template <bool kCollect>
__device__ __noinline__ void handle_flags(const BlockConsts *bc, WarpSmem *ws, uint32_t w0, int M, int lane) {
    const uint32_t bshift = bc->bshift;
    const uint8_t *seg_thr = bc->seg_thr + (size_t)ws->size_a * bc->n_segments;
    const int n_lists = ws->n_lists;
    unsigned lanes = __ballot_sync(kFull, ws->flag[lane] != 0u);
    while (lanes) {
        const int src = __ffs(lanes) - 1;
        lanes &= lanes - 1;
        uint32_t f = ws->flag[src];
        while (f) {
            const int bit = __ffs(f) - 1;
            f &= f - 1;
            const uint32_t bucket = (w0 + (uint32_t)src) * 32u + (uint32_t)bit;
            const uint32_t id_lo = bucket << bshift;
            const int B = segment_of_id(bc, id_lo);
            const int T = (int)__ldg(seg_thr + B);
            if (T == 0) continue;  // the word's threshold came from a neighbouring segment
            if (bshift == 0) {
                int count = (1 << M) - (int)ws->bias[src];  // the planes hold bias + overlap - 2^M
                for (int j = 0; j < M; j++) count += (int)((ws->cnt[j * 32 + src] >> bit) & 1u) << j;
                if (count >= T) offer_candidate<kCollect>(bc, ws, bucket, count, B, lane);
                continue;
            }
            const uint32_t width = 1u << bshift, id_hi = id_lo + width;
            for (uint32_t i = lane; i < width; i += 32) ws->cnt[i] = 0u;
            __syncwarp();
            const uint32_t *__restrict__ postings = bc->postings;
            const size_t stride = (size_t)bc->n_segments + 1;
            for (int j = lane; j < n_lists; j += 32) {
                const uint32_t *o = bc->list_off + (size_t)ws->term[j] * stride + B;
                const uint32_t e = __ldg(o + 1);
                for (uint32_t pos = lower_bound4(postings, __ldg(o), e, id_lo); pos < e; pos++) {
                    const uint32_t x = __ldg(postings + pos);
                    if (x >= id_hi) break;
                    atomicAdd(ws->cnt + (x - id_lo), 1u);
                }
            }
            __syncwarp();
            for (uint32_t base = 0; base < width; base += 32) {
                const uint32_t v = base + lane < width ? ws->cnt[base + lane] : 0u;
                unsigned m = __ballot_sync(kFull, v >= (uint32_t)T);
                while (m) {
                    const int i = __ffs(m) - 1;
                    m &= m - 1;
                    offer_candidate<kCollect>(bc, ws, id_lo + base + (uint32_t)i, (int)__shfl_sync(kFull, v, i), B, lane);
                }
            }
            __syncwarp();
        }
    }
}

// Autocomplete with one or two n-grams (a two- or three-letter prefix): every document of the shorter posting run is a
// candidate, thousands of them, and nearly every bucket of the bitmap would have to be resolved.  Walk the run instead:
// lane l takes posting base + l, checks the other run by binary search, scores its document (language-model lookup or
// -id) - all in parallel - and only what beats the current k-th goes through the warp-serial insert.
// Segments len(tokens)..S-1 (pkg/suggest/autocomplete.go:47) are one contiguous run of each term's postings.
__device__ __noinline__ void complete_from_lists(const BlockConsts *bc, WarpSmem *ws, int lane) {
    const uint32_t S = bc->n_segments;
    const uint32_t b_lo = (uint32_t)ws->size_a;
    if (b_lo >= S) return;
    const size_t stride = (size_t)S + 1;
    const uint32_t *__restrict__ postings = bc->postings;
    const uint32_t *o0 = bc->list_off + (size_t)ws->term[0] * stride;
    uint32_t a_d = __ldg(o0 + b_lo), e_d = __ldg(o0 + S), a_o = 0, e_o = 0;
    const bool two = ws->n_lists == 2;
    if (two) {
        const uint32_t *o1 = bc->list_off + (size_t)ws->term[1] * stride;
        a_o = __ldg(o1 + b_lo);
        e_o = __ldg(o1 + S);
        if (e_o - a_o < e_d - a_d) {  // drive with the shorter run
            const uint32_t ta = a_d, te = e_d;
            a_d = a_o; e_d = e_o; a_o = ta; e_o = te;
        }
    }
    QueryCtx c;
    c.k = bc->k;
    c.tk_len = ws->tk_len;
    c.tk_score = warp_tk_score(ws);
    c.tk_id = warp_tk_id(ws, bc->k);
    for (uint32_t base = a_d; base < e_d; base += 32) {
        const uint32_t i = base + (uint32_t)lane;
        bool valid = i < e_d;
        uint32_t id = 0;
        double score = 0.0;
        if (valid) {
            const uint32_t x = __ldg(postings + i);
            if (two) {
                const uint32_t pos = lower_bound(postings, a_o, e_o, x);
                valid = pos < e_o && __ldg(postings + pos) == x;
            }
            if (valid) {
                id = __ldg(bc->perm + x);
                score = completion_score(bc, ws, id);
                if (c.tk_len == (int)c.k) {  // the k-th only improves: whatever fails against it now fails later too
                    const double ws_ = c.tk_score[c.k - 1];
                    const uint32_t wi = c.tk_id[c.k - 1];
                    valid = score > ws_ || (score == ws_ && id < wi);
                }
            }
        }
        unsigned m = __ballot_sync(kFull, valid);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            topk_insert(c, __shfl_sync(kFull, score, src), __shfl_sync(kFull, id, src), lane);
        }
    }
    __syncwarp();
    if (lane == 0) ws->tk_len = c.tk_len;
    __syncwarp();
}

// The count loop of one query.  The window is walked in tiles of 32 bitmap words (lane l owns word w0 + l of every
// row); inside a tile the lists are taken in blocks of 8.  Per block, seven carry-save adders turn the eight list words
// into one carry of weight 8, which ripples into the planes above.  The planes start at bias = 2^M - T(word), so
// "count >= T" is the carry out of the top plane (kept sticky in ov) and no comparison is needed.  Loads run one block
// ahead of the adders (xa / xb), across tile boundaries, so every warp keeps 8-16 independent 128-byte row reads in
// flight.  The last block of a tile is padded to 8 lists with the all-zero row (1 KB, L1 resident: branch-free loads).
// Returns the first word of the first tile in which a bucket reached its threshold, with the lane's flags, bias and
// planes in ts, or kInf when the window is done.  The caller runs the cold path and resumes behind that tile: called
// from outside this loop its registers do not add to the loop's.
template <int M>
struct TileState {
    uint32_t c[M];   // planes of this lane's word
    uint32_t ov, bias;
};

template <int M>
__device__ __forceinline__ uint32_t count_until_flag(const uint32_t *__restrict__ bitmaps, const uint32_t *s_row,
                                                     const uint8_t *__restrict__ word_thr, uint32_t w_begin, uint32_t win_hi,
                                                     int n_lists, int lane, TileState<M> &ts) {
    const uint32_t n_blocks = ((uint32_t)n_lists + 7u) >> 3;
    const uint32_t n_units = ((win_hi - w_begin + kTileWords - 1) / kTileWords) * n_blocks;
    // loader state: this lane's word of the tile being loaded, next block to load
    const uint32_t *ld_ptr = bitmaps + w_begin + (uint32_t)lane;
    asm volatile("" : "+l"(ld_ptr));  // opaque: row offsets are added to this pointer as 32-bit indices (one IMAD.WIDE per load)
    uint32_t ld_block = 0;
    // (macros, not lambdas over array references: the word registers must stay registers)
#define SG_LOAD_BLOCK(R)                                                                                      \
    do {                                                                                                      \
        const uint4 r0_ = *(const uint4 *)(s_row + ld_block * 8u), r1_ = *(const uint4 *)(s_row + ld_block * 8u + 4u); \
        R##0 = __ldg(ld_ptr + r0_.x);                                                                         \
        R##1 = __ldg(ld_ptr + r0_.y);                                                                         \
        R##2 = __ldg(ld_ptr + r0_.z);                                                                         \
        R##3 = __ldg(ld_ptr + r0_.w);                                                                         \
        R##4 = __ldg(ld_ptr + r1_.x);                                                                         \
        R##5 = __ldg(ld_ptr + r1_.y);                                                                         \
        R##6 = __ldg(ld_ptr + r1_.z);                                                                         \
        R##7 = __ldg(ld_ptr + r1_.w);                                                                         \
        if (++ld_block == n_blocks) {                                                                         \
            ld_block = 0;                                                                                     \
            ld_ptr += kTileWords;                                                                             \
            asm volatile("" : "+l"(ld_ptr));                                                                  \
        }                                                                                                     \
    } while (0)
    // adder state
    uint32_t w0 = w_begin, cons_block = 0;
    const uint8_t *thr_ptr = word_thr + w_begin + (uint32_t)lane;  // threshold of this lane's word of the tile being added
    uint32_t tw_next;                                              // ... of the next tile, loaded a tile ahead
    auto begin_tile = [&](uint32_t T_w) {
        ts.bias = (T_w >> M) ? 0u : (1u << M) - T_w;  // a threshold no count can reach: the planes never overflow
#pragma unroll
        for (int j = 0; j < M; j++) ts.c[j] = 0u - ((ts.bias >> j) & 1u);
        ts.ov = 0u;
    };
    auto consume = [&](uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4, uint32_t x5, uint32_t x6, uint32_t x7) -> bool {
        uint32_t h1, l1, h2, l2, h3, l3, h4, g1, m1, g2, e;
        csa(h1, l1, x0, x1, x2);
        csa(h2, l2, x3, x4, x5);
        csa(h3, l3, x6, x7, ts.c[0]);
        csa(h4, ts.c[0], l1, l2, l3);
        csa(g1, m1, h1, h2, h3);
        csa(g2, ts.c[1], m1, h4, ts.c[1]);
        csa(e, ts.c[2], g1, g2, ts.c[2]);
#pragma unroll
        for (int j = 3; j < M; j++) {
            const uint32_t t = ts.c[j] & e;
            ts.c[j] ^= e;
            e = t;
        }
        ts.ov |= e;
        if (++cons_block == n_blocks) {
            if (__any_sync(kFull, ts.ov != 0u)) return true;  // rare: the caller hands the tile to the cold path
            cons_block = 0;
            w0 += kTileWords;
            thr_ptr += kTileWords;
            begin_tile(tw_next);
            tw_next = w0 + kTileWords < win_hi ? __ldg(thr_ptr + kTileWords) : 255u;
        }
        return false;
    };
    uint32_t xa0, xa1, xa2, xa3, xa4, xa5, xa6, xa7, xb0, xb1, xb2, xb3, xb4, xb5, xb6, xb7;
    xb0 = xb1 = xb2 = xb3 = xb4 = xb5 = xb6 = xb7 = 0u;
    begin_tile(__ldg(thr_ptr));
    tw_next = w0 + kTileWords < win_hi ? __ldg(thr_ptr + kTileWords) : 255u;
    SG_LOAD_BLOCK(xa);
#pragma unroll 1
    for (uint32_t u = 0;;) {
        if (u + 1 < n_units) SG_LOAD_BLOCK(xb);
        if (consume(xa0, xa1, xa2, xa3, xa4, xa5, xa6, xa7)) return w0;
        if (++u == n_units) break;
        if (u + 1 < n_units) SG_LOAD_BLOCK(xa);
        if (consume(xb0, xb1, xb2, xb3, xb4, xb5, xb6, xb7)) return w0;
        if (++u == n_units) break;
    }
#undef SG_LOAD_BLOCK
    return kInf;
}

// One query: count, and for every tile with a hit run the cold path.
template <int M, bool kCollect>
__device__ __forceinline__ void search_query(const uint32_t *__restrict__ bitmaps, const BlockConsts *bc, WarpSmem *ws,
                                             const uint8_t *__restrict__ word_thr, uint32_t win_lo, uint32_t win_hi, int n_lists,
                                             int lane) {
    const bool keep_planes = bc->bshift == 0;
    for (uint32_t w = win_lo & ~(kTileWords - 1); w < win_hi;) {
        TileState<M> ts;
        const uint32_t wf = count_until_flag<M>(bitmaps, ws->row, word_thr, w, win_hi, n_lists, lane, ts);
        if (wf == kInf) break;
        // hand the tile over through shared memory
        ws->flag[lane] = ts.ov;
        ws->bias[lane] = ts.bias;
        if (keep_planes) {
#pragma unroll
            for (int j = 0; j < M; j++) ws->cnt[j * 32 + lane] = ts.c[j];
        }
        __syncwarp();
        handle_flags<kCollect>(bc, ws, wf, M, lane);
        __syncwarp();
        w = wf + kTileWords;
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
    if (p.custom_thr != nullptr) { b_min = 0; b_max = a > 0 ? S - 1 : -1; }  // the table is zero outside the caller's window
    for (int B = threadIdx.x; B < S; B += kWindowThreads) {
        int T = 0;
        if (B >= b_min && B <= b_max) {
            T = p.custom_thr != nullptr ? (int)p.custom_thr[(size_t)a * S + B] : metric_threshold(metric, p.alpha, a, B);
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
// admissible postings / lists of SURVEY.md section 8(d) (the algorithmic bytes of the roofline) and the bitmap words the
// engine reads for the count: 16 bytes per query {postings, lists, bitmap words, 0}.
// ---------------------------------------------------------------------------------------------------------------
// sg_count_kernel's tiles (defined with it below; the stats pass of sg_tokens_kernel repeats its arithmetic)
#ifndef SG_TILE_ALIGN
#define SG_TILE_ALIGN 64
#endif
constexpr uint32_t kTileAlign2 = SG_TILE_ALIGN;   // words the first tile of a window is aligned to
__device__ __forceinline__ uint32_t first_tile_word(uint32_t win_lo, uint32_t win_hi, uint32_t row_words) {
    // SG_TILE_ALIGN=32: tiles start at a 128-byte boundary at or below the window instead of a 256-byte one (a quarter of a
    // tile less per query on average; measured 155 vs 156 us on config #2 - the kernel is bound by the chain of dependent
    // loads in front of a query's count, not by the count's units - so 64 stays).  A row is whole 64-word tiles
    // (choose_layout): a window that ends in the row's last words must not run a tile over the end of the row.
    const uint32_t w = win_lo & ~(kTileAlign2 - 1);
    return w + (win_hi - w + 63u) / 64u * 64u > row_words ? (win_lo & ~63u) : w;
}

template <bool kArrive>
__device__ __forceinline__ void tokens_body(const DevIndex &ix, const SearchParams &p) {
    __shared__ __align__(16) uint32_t s_scratch[kPlanThreads / 32][kMaxRunes + 2 * kMaxQueryTokens];
    __shared__ uint8_t s_ascii[128];
    __shared__ uint32_t s_seen;
    if (kArrive && threadIdx.x == 0) s_seen = p.arrive_base;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < 128) s_ascii[threadIdx.x] = ix.ascii_code[threadIdx.x];
    __syncthreads();
    uint32_t *s_runes = s_scratch[warp];
    uint32_t *s_lterm = s_runes + kMaxRunes;
    uint32_t *s_hash = s_lterm + kMaxQueryTokens;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const size_t stride = (size_t)ix.n_segments + 1;
    if (blockIdx.x == 0 && threadIdx.x < kWorkWords) p.work_counter[threadIdx.x] = 0u;  // counters of the kernels behind this one
    for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < p.n_q; q += n_warps) {
        int size_a = 0, n_lists = 0;
        bool unsupported = false;
        if (!kArrive || wait_for_query(p, q, lane, &s_seen))
            unsupported = tokenize_query<false, kArrive>(ix, p, q, s_runes, s_lterm, s_hash, lane, &size_a, &n_lists, s_ascii);
        if (p.mode == 1 && n_lists < size_a) n_lists = 0;  // a query token that is in no list: nothing can hold them all
        if (unsupported) { size_a = 0; n_lists = 0; }
        uint8_t *plan_base = p.plans + (size_t)q * kTokStride;
        uint32_t *plan_terms = (uint32_t *)(plan_base + kTokTermsOffset);
        for (int j = lane; j < n_lists; j += 32) plan_terms[j] = s_lterm[j];
        if (lane == 0) {
            const WordRange win = p.wt.win[size_a];
            ((uint4 *)plan_base)[0] = make_uint4(unsupported ? 1u : 0u, (uint32_t)size_a, (uint32_t)n_lists, 0u);
            ((uint4 *)plan_base)[1] = make_uint4(win.x, win.y, 0u, 0u);
        }
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
            if (lane == 0) {
                // what the bitmap engine itself reads for the count: every (padded) list's words of the window, whole tiles
                const WordRange win = p.wt.win[size_a];
                const uint32_t tw = p.lean_flags != nullptr ? 64u : kTileWords;  // (sg_count_kernel reads 64-word tiles: kTileWords2)
                const uint32_t w_first = p.lean_flags != nullptr ? first_tile_word(win.x, win.y, ix.row_words) : (win.x & ~(tw - 1));
                const uint32_t tiles = n_lists > 0 && win.y > win.x ? (win.y - w_first + tw - 1) / tw : 0u;
                ((uint4 *)p.stats)[q] = make_uint4(st_postings, st_lists, tiles * tw * (uint32_t)((n_lists + 7) & ~7), 0u);
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kPlanThreads, SG_TOKENS_MIN_BLOCKS) sg_tokens_kernel(const DevIndex ix, const SearchParams p) {
    tokens_body<false>(ix, p);
}

// the same with queries that arrive in chunks while the kernel runs (SearchParams::arrived, sg_search_batch)
__global__ void __launch_bounds__(kPlanThreads, SG_TOKENS_MIN_BLOCKS) sg_tokens_arrive_kernel(const DevIndex ix, const SearchParams p) {
    tokens_body<true>(ix, p);
}

// ---------------------------------------------------------------------------------------------------------------
// sg_bitmap_search_kernel: one warp per query, query numbers from a global counter.
// ---------------------------------------------------------------------------------------------------------------
// kCollect = true is sg_bitmap_collect_kernel (sg_candidates_batch): same count and resolve, every survivor appended to
// the launch-wide candidate list instead of scored into a top-k.  A template so that the top-k kernel's code and register
// allocation are exactly what they are without it.
template <bool kCollect>
__device__ __forceinline__ void bitmap_search_body(const DevIndex &ix, const SearchParams &p) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ BlockConsts s_bc;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        s_bc.postings = ix.postings;
        s_bc.list_off = ix.list_off;
        s_bc.perm = ix.perm;
        s_bc.seg_start = ix.seg_start;
        s_bc.seg_thr = p.wt.seg_thr;
        s_bc.n_segments = ix.n_segments;
        s_bc.bshift = ix.bshift;
        s_bc.id_base = ix.id_base;
        s_bc.k = p.k;
        s_bc.metric = p.mode == 1 ? (int)kAutocomplete : p.metric;
        if (kCollect) {
            s_bc.cand_total = p.cand_total;
            s_bc.cand_cap = p.cand_cap;
            s_bc.cand_query = p.cand_query;
            s_bc.cand_ids = p.cand_ids;
            s_bc.cand_overlap = p.cand_overlap;
            s_bc.cand_segment = p.cand_segment;
        }
    }
    for (uint32_t i = threadIdx.x; i <= min(ix.n_segments, (uint32_t)kSegCache); i += blockDim.x) s_bc.seg_cache[i] = ix.seg_start[i];
    __syncthreads();
    WarpSmem *ws = (WarpSmem *)(smem + (size_t)warp * p.warp_smem);
    if (lane == 0) {
        double *sc = (double *)(ws + 1);
        if (p.tk_global != nullptr) sc = (double *)(p.tk_global + (size_t)(blockIdx.x * kBitmapWarps + warp) * (((size_t)p.k * 12u + 15u) & ~(size_t)15u));
        ws->tk_score = sc;
        ws->tk_id = (uint32_t *)(sc + p.k);
    }
    __syncwarp();
    const double *tk_score = warp_tk_score(ws);
    const uint32_t *tk_id = warp_tk_id(ws, p.k);
    const uint32_t zero_row = ix.n_terms * ix.row_words;
    // behind the count -> resolve pipeline this kernel answers only the queries that ran out of scratch there (kPlanDirty)
    uint32_t *const counter = p.work_counter + (p.only_dirty ? kWorkFallbackQuery : kWorkQuery);

    uint32_t q = 0;
    if (lane == 0) q = take_query(counter);
    q = __shfl_sync(kFull, q, 0);
    while (q < p.n_q) {
        // the plan of this query: header and the first 32 term ids are requested before anything waits on them
        const uint8_t *plan_base = p.plans + (size_t)q * kTokStride;
        const uint4 h0 = __ldg((const uint4 *)plan_base), h1 = __ldg((const uint4 *)plan_base + 1);  // TokenPlan
        const uint32_t t0 = __ldg((const uint32_t *)(plan_base + kTokTermsOffset) + lane);
        uint32_t q_next = 0;
        if (lane == 0) q_next = take_query(counter);  // the next query number travels together with the plan loads
        if (p.only_dirty && !(h0.x & kPlanDirty)) {
            q = __shfl_sync(kFull, q_next, 0);
            continue;
        }
        const bool unsupported = (h0.x & 1u) != 0u;
        const int size_a = (int)h0.y, n_lists = (int)h0.z;
        const WordRange win{h1.x, h1.y};
        int tk_len = 0;

        if (n_lists > 0 && win.y > win.x) {
            if (lane == 0) {
                ws->tk_len = 0;
                ws->size_a = size_a;
                ws->n_lists = n_lists;
                if (kCollect) ws->q = q;
                ws->lm_valid = 0u;
                if (p.lm_ctx != nullptr && p.lm_ctx[q].valid) {
                    const LmContext lc = p.lm_ctx[q];
                    ws->lm_valid = 1u;
                    ws->lm_from = lc.from;
                    ws->lm_to = lc.to;
                    ws->lm_vals = lc.vals;
                }
            }
            const int n_pad = (n_lists + 7) & ~7;
            for (int j = lane; j < n_pad; j += 32) {
                uint32_t row = zero_row;
                if (j < n_lists) {
                    const uint32_t t = j < 32 ? t0 : __ldg((const uint32_t *)(plan_base + kTokTermsOffset) + j);
                    ws->term[j] = t;
                    row = t * ix.row_words;
                }
                ws->row[j] = row;
            }
            __syncwarp();
            const uint8_t *word_thr = p.wt.word_thr + (size_t)size_a * ix.row_words;
            if (p.mode == 1 && n_lists <= 2) complete_from_lists(&s_bc, ws, lane);
            else if (n_lists < 32) search_query<5, kCollect>(ix.bitmaps, &s_bc, ws, word_thr, win.x, win.y, n_lists, lane);
            else search_query<8, kCollect>(ix.bitmaps, &s_bc, ws, word_thr, win.x, win.y, n_lists, lane);
            __syncwarp();
            tk_len = ws->tk_len;
        }

        // results: GetCandidates order, fixed stride k.  sparse_rows: the rows are the caller's page-locked host buffers
        // (sg_search_batch) and every store crosses PCIe - only the valid entries are written
        const size_t row = (size_t)q * p.k;
        const uint32_t n_out = p.sparse_rows ? (uint32_t)tk_len : p.k;
        for (uint32_t j = lane; j < n_out; j += 32) {
            const bool has = (int)j < tk_len;
            store_entry(p.out_ids, p.out_scores, p.out_packed, row + j, has ? ix.id_base + tk_id[j] : 0u, has ? tk_score[j] : 0.0);
        }
        if (lane == 0) {
            // collect mode reports len(tokens): the caller's Distance(inter, sizeA, sizeB) needs it
            // (re-read from the plan: keeping size_a live across the search costs the count loop a register)
            uint32_t n_report = (uint32_t)tk_len;
            if (kCollect) n_report = __ldg((const uint32_t *)(p.plans + (size_t)q * kTokStride) + 1);
            p.out_counts[q] = unsupported ? kCountUnsupported : n_report;
            if (unsupported && p.too_long_flag != nullptr) *p.too_long_flag = 1u;
        }
        __syncwarp();  // every lane is done with this query's shared state before lane 0 resets it for the next
        q = __shfl_sync(kFull, q_next, 0);
    }
}

__global__ void __launch_bounds__(kBitmapWarps * 32, SG_BITMAP_MIN_BLOCKS) sg_bitmap_search_kernel(const DevIndex ix, const SearchParams p) {
    if (p.only_dirty && ((const volatile uint32_t *)p.work_counter)[kWorkDirtyAny] == 0u) return;  // the usual case: nothing to redo
    bitmap_search_body<false>(ix, p);
}

__global__ void __launch_bounds__(kBitmapWarps * 32, SG_BITMAP_MIN_BLOCKS) sg_bitmap_collect_kernel(const DevIndex ix, const SearchParams p) {
    bitmap_search_body<true>(ix, p);
}


// ---------------------------------------------------------------------------------------------------------------
// The count -> resolve pipeline (Suggest top-k on an index with the exact level, sg_fine.cu): the same bit-sliced count
// as above, but a bucket that reaches its threshold is not resolved by the warp that found it.  sg_count_kernel only
// counts and appends {query, bitmap word, flagged buckets} to a launch-wide list; sg_resolve_kernel takes one flagged
// word per warp - so the resolves of a query with thousands of flagged buckets (frequent n-grams, low thresholds) spread
// over the whole GPU instead of serialising one warp - counts the bucket's documents exactly from the pairs' bits
// (DevIndex::rank4 / fine: two loads per list, no posting-list search), links the survivors to their query, and the
// warp that resolves the LAST flagged word of a query scores its survivors, selects the k best and writes the row.
// A query without a flagged word is finished by sg_count_kernel itself.  If a launch runs out of scratch the queries
// concerned are marked kPlanDirty and answered by sg_bitmap_search_kernel (only_dirty) behind the pipeline.
// ---------------------------------------------------------------------------------------------------------------
#ifndef SG_COUNT_MIN_BLOCKS
#define SG_COUNT_MIN_BLOCKS 4           // CTAs per SM the register allocation of sg_count_kernel aims at (64 registers: two sets of
                                        // eight 8-byte row words in flight plus 2 x M planes)
#endif

// The same count with every lane owning TWO adjacent words of a 64-word tile: row reads are 8-byte loads (LDG.64, 256
// bytes per warp-load).  A B200 SM issues 4-byte warp-loads at a rate that caps L2 -> SM traffic near 9.5 TB/s chip-wide;
// with 8-byte loads the same cache delivers ~17 TB/s (tools/l2bench.cu, profiles/l2_peak.json), and the count of
// config #2 is bound by exactly that.  Used by sg_count_kernel.
constexpr uint32_t kTileWords2 = 64;  // (sg_tokens_kernel's stats pass repeats the 64)
template <int M>
struct TileState2 {
    uint32_t c[M][2];   // planes of this lane's two words
    uint32_t ov[2], bias[2];
};

// two adjacent row words
__device__ __forceinline__ uint2 ld_row2(const uint32_t *p) { return __ldg((const uint2 *)p); }

template <int M>
__device__ __forceinline__ uint32_t count_until_flag2(const uint32_t *__restrict__ bitmaps, const uint32_t *s_row,
                                                      const uint8_t *__restrict__ word_thr, uint32_t w_begin, uint32_t win_hi,
                                                      int n_lists, int lane, TileState2<M> &ts) {
    const uint32_t n_blocks = ((uint32_t)n_lists + 7u) >> 3;
    const uint32_t n_units = ((win_hi - w_begin + kTileWords2 - 1) / kTileWords2) * n_blocks;
    const uint32_t *ld_ptr = bitmaps + w_begin + 2u * (uint32_t)lane;
    asm volatile("" : "+l"(ld_ptr));  // opaque: row offsets are added to this pointer as 32-bit indices (one IMAD.WIDE per load)
    uint32_t ld_block = 0;
#define SG_LOAD_BLOCK2(R)                                                                                     \
    do {                                                                                                      \
        const uint4 r0_ = *(const uint4 *)(s_row + ld_block * 8u), r1_ = *(const uint4 *)(s_row + ld_block * 8u + 4u); \
        R##0 = ld_row2(ld_ptr + r0_.x);                                                                       \
        R##1 = ld_row2(ld_ptr + r0_.y);                                                                       \
        R##2 = ld_row2(ld_ptr + r0_.z);                                                                       \
        R##3 = ld_row2(ld_ptr + r0_.w);                                                                       \
        R##4 = ld_row2(ld_ptr + r1_.x);                                                                       \
        R##5 = ld_row2(ld_ptr + r1_.y);                                                                       \
        R##6 = ld_row2(ld_ptr + r1_.z);                                                                       \
        R##7 = ld_row2(ld_ptr + r1_.w);                                                                       \
        if (++ld_block == n_blocks) {                                                                         \
            ld_block = 0;                                                                                     \
            ld_ptr += kTileWords2;                                                                            \
            asm volatile("" : "+l"(ld_ptr));                                                                  \
        }                                                                                                     \
    } while (0)
    uint32_t w0 = w_begin, cons_block = 0;
    const uint8_t *thr_ptr = word_thr + w_begin + 2u * (uint32_t)lane;  // thresholds of this lane's two words (one 16-bit load)
    uint32_t tw_next;
    auto begin_tile = [&](uint32_t T_pair) {
#pragma unroll
        for (int x = 0; x < 2; x++) {
            const uint32_t T_w = (T_pair >> (8 * x)) & 0xFFu;
            ts.bias[x] = (T_w >> M) ? 0u : (1u << M) - T_w;  // a threshold no count can reach: the planes never overflow
#pragma unroll
            for (int j = 0; j < M; j++) ts.c[j][x] = 0u - ((ts.bias[x] >> j) & 1u);
            ts.ov[x] = 0u;
        }
    };
    auto add8 = [&](int x, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4, uint32_t x5, uint32_t x6, uint32_t x7) {
        uint32_t h1, l1, h2, l2, h3, l3, h4, g1, m1, g2, e;
        csa(h1, l1, x0, x1, x2);
        csa(h2, l2, x3, x4, x5);
        csa(h3, l3, x6, x7, ts.c[0][x]);
        csa(h4, ts.c[0][x], l1, l2, l3);
        csa(g1, m1, h1, h2, h3);
        csa(g2, ts.c[1][x], m1, h4, ts.c[1][x]);
        csa(e, ts.c[2][x], g1, g2, ts.c[2][x]);
#pragma unroll
        for (int j = 3; j < M; j++) {
            const uint32_t t = ts.c[j][x] & e;
            ts.c[j][x] ^= e;
            e = t;
        }
        ts.ov[x] |= e;
    };
    auto consume = [&](uint2 x0, uint2 x1, uint2 x2, uint2 x3, uint2 x4, uint2 x5, uint2 x6, uint2 x7) -> bool {
        add8(0, x0.x, x1.x, x2.x, x3.x, x4.x, x5.x, x6.x, x7.x);
        add8(1, x0.y, x1.y, x2.y, x3.y, x4.y, x5.y, x6.y, x7.y);
        if (++cons_block == n_blocks) {
            if (__any_sync(kFull, (ts.ov[0] | ts.ov[1]) != 0u)) return true;  // rare: the caller records the flagged words
            cons_block = 0;
            w0 += kTileWords2;
            thr_ptr += kTileWords2;
            begin_tile(tw_next);
            tw_next = w0 + kTileWords2 < win_hi ? (uint32_t)__ldg((const uint16_t *)(thr_ptr + kTileWords2)) : 0xFFFFu;
        }
        return false;
    };
    uint2 xa0, xa1, xa2, xa3, xa4, xa5, xa6, xa7, xb0, xb1, xb2, xb3, xb4, xb5, xb6, xb7;
    xb0 = xb1 = xb2 = xb3 = xb4 = xb5 = xb6 = xb7 = make_uint2(0u, 0u);
    begin_tile((uint32_t)__ldg((const uint16_t *)thr_ptr));
    tw_next = w0 + kTileWords2 < win_hi ? (uint32_t)__ldg((const uint16_t *)(thr_ptr + kTileWords2)) : 0xFFFFu;
    SG_LOAD_BLOCK2(xa);
#pragma unroll 1
    for (uint32_t u = 0;;) {
        if (u + 1 < n_units) SG_LOAD_BLOCK2(xb);
        if (consume(xa0, xa1, xa2, xa3, xa4, xa5, xa6, xa7)) return w0;
        if (++u == n_units) break;
        if (u + 1 < n_units) SG_LOAD_BLOCK2(xa);
        if (consume(xb0, xb1, xb2, xb3, xb4, xb5, xb6, xb7)) return w0;
        if (++u == n_units) break;
    }
#undef SG_LOAD_BLOCK2
    return kInf;
}


template <int M>
__device__ __forceinline__ uint32_t count_and_flag(const DevIndex &ix, const SearchParams &p, const uint32_t *s_row,
                                                   const uint8_t *__restrict__ word_thr, uint32_t win_lo, uint32_t win_hi, int n_lists,
                                                   uint32_t q, int lane, bool &dirty) {
    const uint32_t cap = p.flag_cap;
    uint32_t n_entries = 0;
    for (uint32_t w = first_tile_word(win_lo, win_hi, ix.row_words); w < win_hi;) {
        TileState2<M> ts;
        const uint32_t wf = count_until_flag2<M>(ix.bitmaps, s_row, word_thr, w, win_hi, n_lists, lane, ts);
        if (wf == kInf) break;
        const unsigned m0 = __ballot_sync(kFull, ts.ov[0] != 0u), m1 = __ballot_sync(kFull, ts.ov[1] != 0u);
        const uint32_t n0 = (uint32_t)__popc(m0), n = n0 + (uint32_t)__popc(m1);
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(p.work_counter + kWorkFlagCursor, n);
        base = __shfl_sync(kFull, base, 0);
        // Every slot below cap that was reserved is written, also when the reservation runs over the end of the list:
        // sg_resolve_kernel reads min(cursor, cap) entries (and skips those of a dirty query).
        const unsigned below = (1u << lane) - 1u;
        const uint32_t i0 = base + (uint32_t)__popc(m0 & below), i1 = base + n0 + (uint32_t)__popc(m1 & below);
        if (ts.ov[0] != 0u && i0 < cap) p.lean_flags[i0] = make_uint4(q, wf + 2u * (uint32_t)lane, ts.ov[0], 0u);
        if (ts.ov[1] != 0u && i1 < cap) p.lean_flags[i1] = make_uint4(q, wf + 2u * (uint32_t)lane + 1u, ts.ov[1], 0u);
        if (base + n <= cap) n_entries += n;
        else dirty = true;
        w = wf + kTileWords2;
    }
    return n_entries;
}

// kFused: the kernel tokenizes the query itself (tokenize_query, the same code sg_tokens_kernel runs) and writes the plan for
// sg_resolve_kernel, instead of reading a plan sg_tokens_kernel wrote: one launch and one pass over the plans less, and the
// tokenizer's chains of dependent loads (offsets -> bytes -> hash probe) hide under the row reads of the other warps.
template <bool kFused, bool kArrive>
__device__ __forceinline__ void count_body(const DevIndex &ix, const SearchParams &p) {
    __shared__ __align__(16) uint32_t s_rows[kBitmapWarps][kRowSlots];
    __shared__ __align__(16) uint32_t s_tok[kFused ? kBitmapWarps : 1][kFused ? kMaxRunes + 2 * kMaxQueryTokens : 1];
    __shared__ uint8_t s_ascii[128];
    __shared__ uint32_t s_seen;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (kFused) {
        if (kArrive && threadIdx.x == 0) s_seen = p.arrive_base;
        if (threadIdx.x < 128) s_ascii[threadIdx.x] = ix.ascii_code[threadIdx.x];
        __syncthreads();
    }
    uint32_t *row = s_rows[warp];
    uint32_t *s_runes = s_tok[kFused ? warp : 0];
    uint32_t *s_lterm = s_runes + kMaxRunes;
    uint32_t *s_hash = s_lterm + kMaxQueryTokens;
    const uint32_t zero_row = ix.n_terms * ix.row_words;
    uint32_t q = 0;
    if (lane == 0) q = take_query(p.work_counter + kWorkQuery);
    q = __shfl_sync(kFull, q, 0);
    while (q < p.n_q) {
        uint8_t *plan_base = p.plans + (size_t)q * kTokStride;
        uint32_t q_next = 0;
        bool unsupported;
        int size_a, n_lists;
        WordRange win;
        uint32_t t0 = 0;
        if (kFused) {
            if (lane == 0) q_next = take_query(p.work_counter + kWorkQuery);
            unsupported = false;
            size_a = n_lists = 0;
            if (!kArrive || wait_for_query(p, q, lane, &s_seen))
                unsupported = tokenize_query<false, kArrive>(ix, p, q, s_runes, s_lterm, s_hash, lane, &size_a, &n_lists, s_ascii);
            if (unsupported) { size_a = 0; n_lists = 0; }
            win = p.wt.win[size_a];
            for (int j = lane; j < n_lists; j += 32) ((uint32_t *)(plan_base + kTokTermsOffset))[j] = s_lterm[j];  // for sg_resolve_kernel
        } else {
            const uint4 h0 = __ldg((const uint4 *)plan_base), h1 = __ldg((const uint4 *)plan_base + 1);  // TokenPlan
            t0 = __ldg((const uint32_t *)(plan_base + kTokTermsOffset) + lane);
            if (lane == 0) q_next = take_query(p.work_counter + kWorkQuery);
            unsupported = (h0.x & 1u) != 0u;
            size_a = (int)h0.y;
            n_lists = (int)h0.z;
            win = WordRange{h1.x, h1.y};
        }
        uint32_t n_entries = 0;
        bool dirty = false;
        if (n_lists > 0 && win.y > win.x) {
            const int n_pad = (n_lists + 7) & ~7;
            for (int j = lane; j < n_pad; j += 32) {
                uint32_t r = zero_row;
                if (j < n_lists) {
                    if (kFused) r = s_lterm[j] * ix.row_words;
                    else r = (j < 32 ? t0 : __ldg((const uint32_t *)(plan_base + kTokTermsOffset) + j)) * ix.row_words;
                }
                row[j] = r;
            }
            __syncwarp();
            const uint8_t *word_thr = p.wt.word_thr + (size_t)size_a * ix.row_words;
            if (n_lists < 32) n_entries = count_and_flag<5>(ix, p, row, word_thr, win.x, win.y, n_lists, q, lane, dirty);
            else n_entries = count_and_flag<8>(ix, p, row, word_thr, win.x, win.y, n_lists, q, lane, dirty);
            __syncwarp();
        }
        if (lane == 0) {
            p.lean_head[q] = kNilNode;
            p.lean_pending[q] = n_entries;
            if (kFused) {
                ((uint4 *)plan_base)[0] = make_uint4((unsupported ? 1u : 0u) | (dirty ? kPlanDirty : 0u), (uint32_t)size_a, (uint32_t)n_lists, 0u);
                ((uint4 *)plan_base)[1] = make_uint4(win.x, win.y, n_entries, 0u);
            } else {
                ((uint32_t *)plan_base)[6] = n_entries;  // TokenPlan::n_flagged
                if (dirty) atomicOr((uint32_t *)plan_base, kPlanDirty);
            }
            if (dirty) p.work_counter[kWorkDirtyAny] = 1u;
        }
        if (n_entries == 0u && !dirty) {  // nothing reached its threshold: the answer is the empty row
            if (!p.sparse_rows) {
                const size_t r0 = (size_t)q * p.k;
                for (uint32_t j = lane; j < p.k; j += 32) store_entry(p.out_ids, p.out_scores, p.out_packed, r0 + j, 0u, 0.0);
            }
            if (lane == 0) {
                p.out_counts[q] = unsupported ? kCountUnsupported : 0u;
                if (unsupported && p.too_long_flag != nullptr) *p.too_long_flag = 1u;
            }
        }
        __syncwarp();
        q = __shfl_sync(kFull, q_next, 0);
    }
}

__global__ void __launch_bounds__(kBitmapWarps * 32, SG_COUNT_MIN_BLOCKS) sg_count_kernel(const DevIndex ix, const SearchParams p) {
    count_body<false, false>(ix, p);
}

__global__ void __launch_bounds__(kBitmapWarps * 32, SG_COUNT_MIN_BLOCKS) sg_tokens_count_kernel(const DevIndex ix, const SearchParams p) {
    count_body<true, false>(ix, p);
}

// the same with queries that arrive in chunks while the kernel runs (SearchParams::arrived, sg_search_batch)
__global__ void __launch_bounds__(kBitmapWarps * 32, SG_COUNT_MIN_BLOCKS) sg_tokens_count_arrive_kernel(const DevIndex ix, const SearchParams p) {
    count_body<true, true>(ix, p);
}

// ---- sg_resolve_kernel: eight lanes per flagged bitmap word, four words per warp ----
// A flagged bucket is counted per document from the bits of its (term, bucket) pairs (DevIndex::fine, found through rank4 +
// popcounts: one 16-byte and one 4-byte load, then the pair's bits): lane g of the group takes lists g, g + 8, ..., two at a
// time so that their loads are in flight together, and adds one to a byte counter in shared memory for every document a
// list has; then every lane compares 16 counters with the threshold (SIMD-in-a-word).  128 documents per pass.
// The chain of dependent loads of a word is short (flag entry -> plan -> {row word group, rank} -> pair bits) and
// thousands of words are in flight per SM; a warp per word (round 2's first version) left the GPU waiting on ~10
// dependent round trips per word at 5 words per warp, a thread per word spent ~3,500 instructions on bit-sliced adds.
// Survivors: a query whose only flagged word (TokenPlan::n_flagged == 1, the usual case) holds one survivor is written
// straight to its row.  Otherwise survivors are scored where they are found and linked to the query ({id, next, score}),
// every flagged word of the query "arrives" (atomicSub on lean_pending), and the thread that arrives last selects the k
// best ((score desc, id asc), Candidate.Less, pkg/suggest/collector.go:20-26).  The list lives in L2 (.cg accesses): other
// SMs wrote it.
__device__ __forceinline__ bool plan_is_dirty(const uint8_t *plan_base) { return (*(const volatile uint32_t *)plan_base & kPlanDirty) != 0u; }

__device__ __forceinline__ void mark_dirty(const SearchParams &p, uint32_t q) {
    atomicOr((uint32_t *)(p.plans + (size_t)q * kTokStride), kPlanDirty);
    p.work_counter[kWorkDirtyAny] = 1u;
}

// A survivor is scored by the lane that found it (original id from perm, 1 - Distance in float64) and linked in front of
// its query's list: node = {id, next | taken << 31, score}.
__device__ __forceinline__ void link_survivor(const DevIndex &ix, const SearchParams &p, uint32_t q, int size_a, uint32_t slot, int count,
                                              int size_b) {
    const uint32_t idx = atomicAdd(p.work_counter + kWorkNodeCursor, 1u);
    if (idx < p.node_cap) {
        const uint32_t id = __ldg(ix.perm + slot);
        const double score = metric_score(p.metric, count, size_a, size_b);
        const uint32_t prev = atomicExch(p.lean_head + q, idx);
        __stcg(p.lean_nodes + idx, make_uint4(id, prev & 0x7FFFFFFFu, (uint32_t)__double2loint(score), (uint32_t)__double2hiint(score)));  // bit 31: taken
    } else {
        mark_dirty(p, q);  // out of nodes: the fallback kernel answers this query
    }
}

__device__ __forceinline__ void write_row_end(const DevIndex &ix, const SearchParams &p, uint32_t q, uint32_t n_out) {
    if (!p.sparse_rows) {
        const size_t row = (size_t)q * p.k;
        for (uint32_t j = n_out; j < p.k; j++) store_entry(p.out_ids, p.out_scores, p.out_packed, row + j, 0u, 0.0);
    }
    p.out_counts[q] = n_out;  // (a query with too many n-grams has no lists and never gets here)
}

// Every flagged word of query q is resolved and this thread arrived last: the k best survivors, (score desc, id asc).
// k <= kSelectLocal: one walk of the list, insertion into a sorted array in local memory; larger k: the best node that is
// not taken yet, k times (nodes carry a taken bit).  One dependent load per node either way: scores travel with the nodes.
constexpr uint32_t kSelectLocal = 32;
struct SelectArgs {  // by value: a reference to the kernel parameters would make every thread copy them to local memory
    uint32_t k, id_base;
    int32_t sparse_rows;
    const uint32_t *lean_head;
    uint4 *lean_nodes;
    uint32_t *out_ids, *out_counts;
    double *out_scores;
    uint4 *out_packed;
};
__device__ __noinline__ void select_survivors(const SelectArgs p, uint32_t q) {
    const size_t row = (size_t)q * p.k;
    uint32_t n_out = 0;
    if (p.k <= kSelectLocal) {
        double bs[kSelectLocal];
        uint32_t bi[kSelectLocal];
        for (uint32_t n = __ldcg(p.lean_head + q); n != kNilNode;) {
            const uint4 node = __ldcg(p.lean_nodes + n);
            const double sc = __hiloint2double((int)node.w, (int)node.z);
            n = node.y & 0x7FFFFFFFu;
            if (n == 0x7FFFFFFFu) n = kNilNode;
            uint32_t at = n_out;  // position of the new entry: behind everything that is better
            while (at > 0 && (bs[at - 1] < sc || (bs[at - 1] == sc && bi[at - 1] > node.x))) at--;
            if (at >= p.k) continue;
            const uint32_t last = n_out < p.k ? n_out : p.k - 1;
            for (uint32_t j = last; j > at; j--) { bs[j] = bs[j - 1]; bi[j] = bi[j - 1]; }
            bs[at] = sc;
            bi[at] = node.x;
            if (n_out < p.k) n_out++;
        }
        for (uint32_t j = 0; j < n_out; j++) store_entry(p.out_ids, p.out_scores, p.out_packed, row + j, p.id_base + bi[j], bs[j]);
    } else {
        for (; n_out < p.k; n_out++) {
            uint32_t best = kNilNode, best_id = 0;
            double best_score = 0.0;
            for (uint32_t n = __ldcg(p.lean_head + q); n != kNilNode;) {
                const uint4 node = __ldcg(p.lean_nodes + n);
                if (!(node.y >> 31)) {
                    const double sc = __hiloint2double((int)node.w, (int)node.z);
                    if (best == kNilNode || sc > best_score || (sc == best_score && node.x < best_id)) { best = n; best_score = sc; best_id = node.x; }
                }
                n = node.y & 0x7FFFFFFFu;
                if (n == 0x7FFFFFFFu) n = kNilNode;
            }
            if (best == kNilNode) break;
            atomicOr(&p.lean_nodes[best].y, 0x80000000u);
            store_entry(p.out_ids, p.out_scores, p.out_packed, row + n_out, p.id_base + best_id, best_score);
        }
    }
    if (!p.sparse_rows)
        for (uint32_t j = n_out; j < p.k; j++) store_entry(p.out_ids, p.out_scores, p.out_packed, row + j, 0u, 0.0);
    p.out_counts[q] = n_out;
}

#ifndef SG_RESOLVE_MIN_BLOCKS
#define SG_RESOLVE_MIN_BLOCKS 4           // 64 registers; 5 (48 registers) measures the same, 6 and 8 spill and are slower, 1 (more registers, fewer warps) is 35 % slower
#endif
__global__ void __launch_bounds__(kResolveThreads, SG_RESOLVE_MIN_BLOCKS) sg_resolve_kernel(const DevIndex ix, const SearchParams p) {
    __shared__ uint32_t s_seg[kSegCache + 1];
    __shared__ __align__(16) uint32_t s_cnt[kResolveThreads / kResolveGroup][32];  // per group: 128 byte counters, one per document
    for (uint32_t i = threadIdx.x; i <= min(ix.n_segments, (uint32_t)kSegCache); i += blockDim.x) s_seg[i] = ix.seg_start[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int gl = lane & (kResolveGroup - 1);                       // lane inside the group
    const unsigned gmask = ((1u << kResolveGroup) - 1u) << (lane & ~(kResolveGroup - 1));
    uint32_t *cnt = s_cnt[threadIdx.x / kResolveGroup];
    const uint32_t bshift = ix.bshift, W = 1u << bshift;
    const uint32_t n_entries = min(((const volatile uint32_t *)p.work_counter)[kWorkFlagCursor], p.flag_cap);
    const uint32_t n_groups = gridDim.x * (kResolveThreads / kResolveGroup);
    const int S = (int)ix.n_segments, cached = min(S, kSegCache);
    for (uint32_t e = blockIdx.x * (kResolveThreads / kResolveGroup) + threadIdx.x / kResolveGroup; e < n_entries; e += n_groups) {
        const uint4 f = __ldg(p.lean_flags + e);
        const uint32_t q = f.x, w = f.y;
        uint32_t mask = f.z;
        const uint8_t *plan_base = p.plans + (size_t)q * kTokStride;
        const uint4 h0 = __ldg((const uint4 *)plan_base);  // len(tokens), list count: never change (the flags word may)
        const uint32_t n_flagged = __ldcg((const uint32_t *)plan_base + 6);  // written by sg_count_kernel
        const uint32_t *terms = (const uint32_t *)(plan_base + kTokTermsOffset);
        // the first round's term ids travel with the header, not behind it: the plan's term area is 128 entries whatever the
        // list count (entries past it are stale and ignored below)
        uint32_t term0[kResolveLists];
#pragma unroll
        for (int u = 0; u < kResolveLists; u++) term0[u] = __ldg(terms + gl + u * kResolveGroup);
        const int size_a = (int)h0.y, n_lists = (int)h0.z;
        // answered by the fallback kernel; nobody waits for this word.  ONE lane looks: the flag can change between the looks
        // of two lanes, and a group that disagreed about it would split around the shuffles and barriers below.
        uint32_t is_dirty = gl == 0 ? (uint32_t)plan_is_dirty(plan_base) : 0u;
        is_dirty = __shfl_sync(gmask, is_dirty, lane & ~(kResolveGroup - 1));
        if (is_dirty) continue;
        const uint8_t *seg_thr = p.wt.seg_thr + (size_t)size_a * ix.n_segments;
        const bool sole = n_flagged == 1u;
        if (p.debug & 4u) continue;
        // this lane's first survivor stays in registers: usually it is the only one of the query
        uint32_t n_surv = 0, s_slot = 0;
        int s_count = 0, s_b = 0;
        auto survivor = [&](uint32_t slot, int count, int B) {
            if (p.debug & 2u) return;
            if (n_surv == 0) { s_slot = slot; s_count = count; s_b = B; }
            else link_survivor(ix, p, q, size_a, slot, count, B);
            n_surv++;
        };
        while (mask) {
            const uint32_t bit = (uint32_t)__ffs(mask) - 1u;
            mask &= mask - 1u;
            const uint32_t bucket = w * 32u + bit;
            const uint32_t id_lo = bucket << bshift;
            int B;
            {   // segment that owns the bucket: seg_start[B] <= id_lo < seg_start[B + 1]
                int lo = 0, hi = S - 1;
                if (s_seg[cached] > id_lo) {
                    hi = cached - 1;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_seg[mid + 1] <= id_lo) lo = mid + 1; else hi = mid; }
                } else {
                    lo = cached;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(ix.seg_start + mid + 1) <= id_lo) lo = mid + 1; else hi = mid; }
                }
                B = lo;
            }
            const int T = (int)__ldg(seg_thr + B);
            if (T == 0) continue;  // the word's threshold came from a neighbouring segment
            if (bshift == 0u) {   // one bit per document: the overlap is the number of lists with the bit
                int count = 0;
                for (int j = gl; j < n_lists; j += kResolveGroup)
                    count += (int)((__ldg(ix.bitmaps + (size_t)__ldg(terms + j) * ix.row_words + w) >> bit) & 1u);
#pragma unroll
                for (int o = kResolveGroup / 2; o; o >>= 1) count += __shfl_xor_sync(gmask, count, o);
                if (count >= T && gl == 0) survivor(bucket, count, B);
                continue;
            }
            const uint32_t chunk_bits = W < 128u ? W : 128u;  // documents counted per pass
            for (uint32_t c0 = 0; c0 < W; c0 += 128u) {
                *(uint4 *)(cnt + 4 * gl) = make_uint4(0u, 0u, 0u, 0u);
                __syncwarp(gmask);
                // lists gl, gl + 8, gl + 16, gl + 24 of a round together: every load of a level is issued before anything waits
                // (absent lists and lists without the bucket read a harmless address instead of branching around the load)
                for (int j0 = gl; j0 < n_lists; j0 += kResolveLists * kResolveGroup) {
                    uint32_t term[kResolveLists];
#pragma unroll
                    for (int u = 0; u < kResolveLists; u++) {
                        const int j = j0 + u * kResolveGroup;
                        term[u] = j0 == gl ? term0[u] : __ldg(terms + (j < n_lists ? j : j0));
                        if (j >= n_lists) term[u] = ix.n_terms;  // absent list: the all-zero row (1 KB, stays in L1), no bucket, no pair
                    }
                    uint4 grp[kResolveLists];
                    uint32_t rk[kResolveLists];
#pragma unroll
                    for (int u = 0; u < kResolveLists; u++) {
                        const size_t at = (size_t)term[u] * ix.row_words + w;
                        grp[u] = __ldg((const uint4 *)(ix.bitmaps + (at & ~(size_t)3)));
                        rk[u] = __ldg(ix.rank4 + (at >> 2));
                    }
                    uint4 m[kResolveLists];
                    bool has[kResolveLists];
#pragma unroll
                    for (int u = 0; u < kResolveLists; u++) {
                        const uint32_t in = (uint32_t)(((size_t)term[u] * ix.row_words + w) & 3);
                        const uint32_t rw = in == 0 ? grp[u].x : in == 1 ? grp[u].y : in == 2 ? grp[u].z : grp[u].w;
                        has[u] = j0 + u * kResolveGroup < n_lists && ((rw >> bit) & 1u);
                        uint32_t pair = rk[u] + (uint32_t)__popc(rw & ((1u << bit) - 1u));
                        if (in > 0) pair += (uint32_t)__popc(grp[u].x);
                        if (in > 1) pair += (uint32_t)__popc(grp[u].y);
                        if (in > 2) pair += (uint32_t)__popc(grp[u].z);
                        const uint64_t bitpos = has[u] && !(p.debug & 1u) ? ((uint64_t)pair << bshift) + c0 : 0ull;
                        m[u] = make_uint4(0u, 0u, 0u, 0u);
                        if (chunk_bits == 128u) m[u] = __ldg((const uint4 *)(ix.fine + (bitpos >> 5)));
                        else if (chunk_bits == 64u) { const uint2 v = __ldg((const uint2 *)(ix.fine + (bitpos >> 5))); m[u].x = v.x; m[u].y = v.y; }
                        else if (chunk_bits == 32u) m[u].x = __ldg(ix.fine + (bitpos >> 5));
                        else m[u].x = (__ldg(ix.fine + (bitpos >> 5)) >> (uint32_t)(bitpos & 31u)) & ((1u << chunk_bits) - 1u);
                    }
#pragma unroll
                    for (int u = 0; u < kResolveLists; u++) {
                        if (!has[u]) continue;
                        const uint32_t mw[4] = {m[u].x, m[u].y, m[u].z, m[u].w};
#pragma unroll
                        for (int x = 0; x < 4; x++) {
                            uint32_t o = mw[x];
                            while (o) {  // one byte counter per document: a list adds one to every document it has
                                const uint32_t b = 32u * (uint32_t)x + (uint32_t)__ffs(o) - 1u;
                                o &= o - 1u;
                                atomicAdd(cnt + (b >> 2), 1u << (8u * (b & 3u)));
                            }
                        }
                    }
                }
                __syncwarp(gmask);
                // documents whose counter reached the threshold: lane gl scans counters 16 gl .. 16 gl + 15
                const uint4 mine = *(const uint4 *)(cnt + 4 * gl);
                const uint32_t cw[4] = {mine.x, mine.y, mine.z, mine.w};
                const uint32_t t4 = (uint32_t)T * 0x01010101u;
#pragma unroll
                for (int x = 0; x < 4; x++) {
                    uint32_t ge = __vcmpgeu4(cw[x], t4);  // 0xFF in every byte that is >= T
                    while (ge) {
                        const uint32_t byte = ((uint32_t)__ffs(ge) - 1u) >> 3;
                        ge &= ~(0xFFu << (8u * byte));
                        survivor(id_lo + c0 + 16u * (uint32_t)gl + 4u * (uint32_t)x + byte, (int)((cw[x] >> (8u * byte)) & 0xFFu), B);
                    }
                }
                __syncwarp(gmask);
            }
        }
        // what the group found
        uint32_t total = n_surv;
#pragma unroll
        for (int o = kResolveGroup / 2; o; o >>= 1) total += __shfl_xor_sync(gmask, total, o);
        if (sole && total <= 1u) {  // the whole answer of the query is known here: no list, no atomics
            if (n_surv == 1u) {
                const size_t row = (size_t)q * p.k;
                store_entry(p.out_ids, p.out_scores, p.out_packed, row, ix.id_base + __ldg(ix.perm + s_slot), metric_score(p.metric, s_count, size_a, s_b));
            }
            if (gl == 0) write_row_end(ix, p, q, total);
            continue;
        }
        if (n_surv >= 1u) link_survivor(ix, p, q, size_a, s_slot, s_count, s_b);
        __syncwarp(gmask);
        if (gl == 0) {
            __threadfence();  // the nodes of every lane of the group (ordered before this by the barrier) are visible before the arrival is
            bool last = sole;
            if (!sole) {
                last = atomicSub(p.lean_pending + q, 1u) == 1u;
                if (last) __threadfence();
            }
            if (last && !plan_is_dirty(plan_base))
                select_survivors(SelectArgs{p.k, ix.id_base, p.sparse_rows, p.lean_head, p.lean_nodes, p.out_ids, p.out_counts, p.out_scores, p.out_packed}, q);
        }
        __syncwarp(gmask);
    }
}

// ---------------- launcher (host) ----------------
// k > kSmemTopK: the top-k lives in HBM (SearchParams::tk_global), the warp keeps only its fixed state in shared memory
size_t bitmap_warp_smem(uint32_t k) { return ((size_t)kBitmapWarpFixedSmem + (k <= kSmemTopK ? (size_t)k * 12u : 0u) + 15u) & ~(size_t)15u; }

// The dynamic shared-memory opt-in is a per-device attribute of a kernel, shared by every index and host thread of
// the process: only ever raise it.  CTAs per SM are cached per (kernel, device, k).
static cudaError_t kernel_occupancy(const void *func, int which, int device, uint32_t k, size_t smem, int *blocks_per_sm) {
    static std::mutex mu;
    static size_t opted_in[3][64] = {{0}};
    static std::map<std::pair<int, std::pair<int, uint32_t>>, int> cache;
    std::lock_guard<std::mutex> lock(mu);
    if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
    if (smem > opted_in[which][device]) {
        cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        opted_in[which][device] = smem;
    }
    auto it = cache.find({which, {device, k}});
    if (it == cache.end()) {
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, func, kBitmapWarps * 32, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorInvalidConfiguration;
        it = cache.emplace(std::make_pair(which, std::make_pair(device, k)), per_sm).first;
    }
    *blocks_per_sm = it->second;
    return cudaSuccess;
}

cudaError_t bitmap_search_occupancy(int device, uint32_t k, int *blocks_per_sm) {
    return kernel_occupancy((const void *)sg_bitmap_search_kernel, 0, device, k, (size_t)kBitmapWarps * bitmap_warp_smem(k), blocks_per_sm);
}

cudaError_t lean_occupancy(int device, uint32_t k, int *count_per_sm, int *resolve_per_sm) {
    (void)k;
    cudaError_t e = kernel_occupancy((const void *)sg_tokens_count_kernel, 1, device, 0, 0, count_per_sm);  // (the unfused kernel needs no more)
    if (e != cudaSuccess) return e;
    *resolve_per_sm = 2048 / kResolveThreads;  // CTAs of kResolveThreads per SM the grid is sized for (the kernel strides over the flagged words)
    return cudaSuccess;
}

// CUDA loads a kernel's code at its first launch (lazy module loading), and that load can wait for the device to go idle.
// sg_search_batch launches kernels that wait on the device for copies the host has yet to enqueue: a first launch behind
// such a kernel would never return.  Every kernel of this file is therefore loaded when an index is created.
cudaError_t preload_bitmap_kernels() {
    const void *kernels[] = {(const void *)sg_window_kernel, (const void *)sg_tokens_kernel, (const void *)sg_tokens_arrive_kernel,
                             (const void *)sg_count_kernel, (const void *)sg_tokens_count_kernel, (const void *)sg_tokens_count_arrive_kernel,
                             (const void *)sg_resolve_kernel, (const void *)sg_bitmap_search_kernel, (const void *)sg_bitmap_collect_kernel};
    for (const void *k : kernels) {
        cudaFuncAttributes attr;
        cudaError_t e = cudaFuncGetAttributes(&attr, k);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_window(const DevIndex &ix, const SearchParams &p, cudaStream_t stream) {
    sg_window_kernel<<<kWindowRows, kWindowThreads, 0, stream>>>(ix, p);
    return cudaGetLastError();
}

cudaError_t launch_bitmap_search(const DevIndex &ix, const SearchParams &p, int sm_count, int per_sm, bool run_window,
                                 cudaStream_t stream, cudaEvent_t *stage_events) {
    const size_t smem = (size_t)kBitmapWarps * p.warp_smem;
    cudaError_t e;
    if (stage_events) cudaEventRecord(stage_events[0], stream);
    if (run_window) {
        e = launch_window(ix, p, stream);
        if (e != cudaSuccess) return e;
    }
    if (stage_events) cudaEventRecord(stage_events[1], stream);
    const int tok_blocks = (int)((p.n_q + kPlanThreads / 32 - 1) / (kPlanThreads / 32));
    if (p.arrived != nullptr) sg_tokens_arrive_kernel<<<tok_blocks < sm_count * 8 ? tok_blocks : sm_count * 8, kPlanThreads, 0, stream>>>(ix, p);
    else sg_tokens_kernel<<<tok_blocks < sm_count * 8 ? tok_blocks : sm_count * 8, kPlanThreads, 0, stream>>>(ix, p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (stage_events) cudaEventRecord(stage_events[2], stream);
    int blocks = sm_count * per_sm;
    const int need = (int)((p.n_q + kBitmapWarps - 1) / kBitmapWarps);
    if (blocks > need) blocks = need;
    if (p.cand_total != nullptr) sg_bitmap_collect_kernel<<<blocks, kBitmapWarps * 32, smem, stream>>>(ix, p);  // k = 1: under 48 KB, no opt-in
    else sg_bitmap_search_kernel<<<blocks, kBitmapWarps * 32, smem, stream>>>(ix, p);
    if (stage_events) cudaEventRecord(stage_events[3], stream);
    return cudaGetLastError();
}

// sg_window_kernel (optional) + sg_tokens_kernel + sg_count_kernel + sg_resolve_kernel + sg_bitmap_search_kernel (only_dirty);
// stage_events: 6 events around the five kernels
cudaError_t launch_lean_search(const DevIndex &ix, const SearchParams &p, int sm_count, int count_per_sm, int resolve_per_sm,
                               int search_per_sm, bool run_window, bool fused, cudaStream_t stream, cudaEvent_t *stage_events) {
    const size_t smem = (size_t)kBitmapWarps * p.warp_smem;
    cudaError_t e;
    if (stage_events) cudaEventRecord(stage_events[0], stream);
    if (run_window) {
        e = launch_window(ix, p, stream);
        if (e != cudaSuccess) return e;
    }
    if (stage_events) cudaEventRecord(stage_events[1], stream);
    const int need = (int)((p.n_q + kBitmapWarps - 1) / kBitmapWarps);
    int blocks = sm_count * count_per_sm;
    if (blocks > need) blocks = need;
    if (fused) {
        // (the counters of the launch are zeroed by the caller: no kernel runs before this one)
        if (stage_events) cudaEventRecord(stage_events[2], stream);
        if (p.arrived != nullptr) sg_tokens_count_arrive_kernel<<<blocks, kBitmapWarps * 32, 0, stream>>>(ix, p);
        else sg_tokens_count_kernel<<<blocks, kBitmapWarps * 32, 0, stream>>>(ix, p);
    } else {
        const int tok_blocks = (int)((p.n_q + kPlanThreads / 32 - 1) / (kPlanThreads / 32));
        if (p.arrived != nullptr) sg_tokens_arrive_kernel<<<tok_blocks < sm_count * 8 ? tok_blocks : sm_count * 8, kPlanThreads, 0, stream>>>(ix, p);
        else sg_tokens_kernel<<<tok_blocks < sm_count * 8 ? tok_blocks : sm_count * 8, kPlanThreads, 0, stream>>>(ix, p);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (stage_events) cudaEventRecord(stage_events[2], stream);
        sg_count_kernel<<<blocks, kBitmapWarps * 32, 0, stream>>>(ix, p);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (stage_events) cudaEventRecord(stage_events[3], stream);
    blocks = sm_count * resolve_per_sm;
    sg_resolve_kernel<<<blocks, kResolveThreads, 0, stream>>>(ix, p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (stage_events) cudaEventRecord(stage_events[4], stream);
    // normally no query is dirty and every CTA of this kernel only looks at the flag
    SearchParams pf = p;
    pf.only_dirty = 1;
    blocks = sm_count * search_per_sm;
    if (blocks > need) blocks = need;
    sg_bitmap_search_kernel<<<blocks, kBitmapWarps * 32, smem, stream>>>(ix, pf);
    if (stage_events) cudaEventRecord(stage_events[5], stream);
    return cudaGetLastError();
}

}  // namespace sg
