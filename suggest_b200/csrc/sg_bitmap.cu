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

// Per-warp shared memory of sg_bitmap_search_kernel.  The count loop itself only reads `row`; everything
// else belongs to the cold path (a bucket reached its threshold), which keeps its state here so that the hot loop's
// registers stay free.
struct WarpSmem {
    int32_t tk_len;                    // candidates in the top-k
    int32_t size_a, n_lists;
    uint32_t lm_valid, lm_from, lm_to; // spellchecker completions: rank by the language model (LmContext of the query)
    const uint64_t *lm_vals;
    uint32_t q;                        // number of the query inside the launch (collect mode)
    uint32_t pad[7];
    uint32_t flag[32];                 // per lane: buckets of its word that reached the threshold
    uint32_t bias[32];                 // per lane: 2^M - T(word), what the planes started from
    alignas(16) uint32_t row[kRowSlots];   // word offset of the bitmap row of every list
    uint32_t term[kMaxQueryTokens];    // term id of every list
    uint32_t cnt[kResolveSlots];       // bshift > 0: per-document counters of the bucket being resolved;
                                       // bshift = 0: the planes of every lane, cnt[j * 32 + lane]
};
constexpr uint32_t kBitmapWarpFixedSmem = (uint32_t)sizeof(WarpSmem);
static_assert(sizeof(WarpSmem) % 16 == 0, "top-k scores follow and need 8-byte alignment");

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

__device__ __forceinline__ double *warp_tk_score(WarpSmem *ws) { return (double *)(ws + 1); }
__device__ __forceinline__ uint32_t *warp_tk_id(WarpSmem *ws, uint32_t k) { return (uint32_t *)(warp_tk_score(ws) + k); }

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
__global__ void __launch_bounds__(kPlanThreads, SG_TOKENS_MIN_BLOCKS) sg_tokens_kernel(const DevIndex ix, const SearchParams p) {
    __shared__ __align__(16) uint32_t s_scratch[kPlanThreads / 32][kMaxRunes + 2 * kMaxQueryTokens];
    __shared__ uint8_t s_ascii[128];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < 128) s_ascii[threadIdx.x] = ix.ascii_code[threadIdx.x];
    __syncthreads();
    uint32_t *s_runes = s_scratch[warp];
    uint32_t *s_lterm = s_runes + kMaxRunes;
    uint32_t *s_hash = s_lterm + kMaxQueryTokens;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const size_t stride = (size_t)ix.n_segments + 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) *p.work_counter = 0u;  // query counter of the search kernel behind this one
    for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < p.n_q; q += n_warps) {
        int size_a = 0, n_lists = 0;
        const bool unsupported = tokenize_query(ix, p, q, s_runes, s_lterm, s_hash, lane, &size_a, &n_lists, s_ascii);
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
                const uint32_t tiles = n_lists > 0 && win.y > win.x ? (win.y - (win.x & ~(kTileWords - 1)) + kTileWords - 1) / kTileWords : 0u;
                ((uint4 *)p.stats)[q] = make_uint4(st_postings, st_lists, tiles * kTileWords * (uint32_t)((n_lists + 7) & ~7), 0u);
            }
        }
        __syncwarp();
    }
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
    const double *tk_score = warp_tk_score(ws);
    const uint32_t *tk_id = warp_tk_id(ws, p.k);
    const uint32_t zero_row = ix.n_terms * ix.row_words;

    uint32_t q = 0;
    if (lane == 0) q = take_query(p.work_counter);
    q = __shfl_sync(kFull, q, 0);
    while (q < p.n_q) {
        // the plan of this query: header and the first 32 term ids are requested before anything waits on them
        const uint8_t *plan_base = p.plans + (size_t)q * kTokStride;
        const uint4 h0 = __ldg((const uint4 *)plan_base), h1 = __ldg((const uint4 *)plan_base + 1);  // TokenPlan
        const uint32_t t0 = __ldg((const uint32_t *)(plan_base + kTokTermsOffset) + lane);
        uint32_t q_next = 0;
        if (lane == 0) q_next = take_query(p.work_counter);  // the next query number travels together with the plan loads
        const bool unsupported = h0.x != 0u;
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
            p.out_ids[row + j] = has ? ix.id_base + tk_id[j] : 0u;
            p.out_scores[row + j] = has ? tk_score[j] : 0.0;
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
    bitmap_search_body<false>(ix, p);
}

__global__ void __launch_bounds__(kBitmapWarps * 32, SG_BITMAP_MIN_BLOCKS) sg_bitmap_collect_kernel(const DevIndex ix, const SearchParams p) {
    bitmap_search_body<true>(ix, p);
}

// ---------------- launcher (host) ----------------
size_t bitmap_warp_smem(uint32_t k) { return ((size_t)kBitmapWarpFixedSmem + (size_t)k * 12u + 15u) & ~(size_t)15u; }

// The dynamic shared-memory opt-in is a per-device attribute of the kernel, shared by every index and host thread of
// the process: only ever raise it.  CTAs per SM are cached per (device, k).
cudaError_t bitmap_search_occupancy(int device, uint32_t k, int *blocks_per_sm) {
    static std::mutex mu;
    static size_t opted_in[64] = {0};
    static std::map<std::pair<int, uint32_t>, int> cache;
    std::lock_guard<std::mutex> lock(mu);
    const size_t smem = (size_t)kBitmapWarps * bitmap_warp_smem(k);
    if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
    if (smem > opted_in[device]) {
        cudaError_t e = cudaFuncSetAttribute(sg_bitmap_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        opted_in[device] = smem;
    }
    auto it = cache.find({device, k});
    if (it == cache.end()) {
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sg_bitmap_search_kernel, kBitmapWarps * 32, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorInvalidConfiguration;
        it = cache.emplace(std::make_pair(device, k), per_sm).first;
    }
    *blocks_per_sm = it->second;
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
    sg_tokens_kernel<<<tok_blocks < sm_count * 8 ? tok_blocks : sm_count * 8, kPlanThreads, 0, stream>>>(ix, p);
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

}  // namespace sg
