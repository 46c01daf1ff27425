// sg_device.h — structures shared by the host side and the sm_100a kernels of libsuggest_b200.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define SG_HD __host__ __device__
#else
#define SG_HD
struct alignas(16) uint4 {  // layout of CUDA's uint4 for the host-only translation units
    unsigned int x, y, z, w;
};
#endif

namespace sg {

constexpr int kMaxNgram = 8;          // pkg/analysis/ngram_tokenizer.go:3
constexpr int kMaxQueryTokens = 128;  // SG_MAX_QUERY_TOKENS
constexpr int kMaxWrapRunes = 8;
constexpr int kMaxRunes = kMaxQueryTokens + kMaxNgram;  // wrapped query runes staged per warp
constexpr uint32_t kNoTerm = 0xFFFFFFFFu;
// per-warp shared memory of sg_search_kernel: [counter table | TMA ring | mbarriers | run slices | thresholds | top-k]
#ifndef SG_SLICE_BYTES
#define SG_SLICE_BYTES 2048
#endif
#ifndef SG_RING_SLOTS
#define SG_RING_SLOTS 2
#endif
constexpr uint32_t kSliceBytes = SG_SLICE_BYTES;  // one TMA bulk copy: 512 postings
constexpr uint32_t kRingSlots = SG_RING_SLOTS;    // slices in flight per warp (at most 8)
constexpr uint32_t kWarpFixedSmem = kRingSlots * kSliceBytes + 64 + 128 + 64 + 3 * kMaxQueryTokens * 4 + 256;
constexpr int kMaxSearchThreads = 512;       // launch bound of sg_search_kernel (16 warps)
constexpr uint32_t kCountUnsupported = 0xFFFFFFFFu;     // SG_COUNT_UNSUPPORTED
constexpr uint32_t kMaxBucketShift = 8;      // at most 256 documents per bitmap bucket
constexpr uint32_t kSmemTopK = 1024;         // largest k whose per-warp top-k is kept in shared memory (bitmap engine)

// non-ASCII alphabet interval: rune r in [lo, hi] has symbol code base + (r - lo)
struct RuneRange {
    uint32_t lo, hi, base;
};

// Read-only view of an index in HBM; passed to kernels by value.
struct DevIndex {
    // ---- tokenizer (pkg/suggest/tokenizer.go:9-20) ----
    int32_t n;               // nGramSize
    int32_t bits;            // bits per symbol code inside a packed term key
    uint32_t pad_code;       // code written for a rune outside the alphabet (pkg/analysis/normalizer.go:29-33)
    int32_t n_wrap_start, n_wrap_end;
    int32_t wrap_ascii;      // 1: every wrap rune is ASCII (the tokenizer's one-rune-per-lane fast path applies)
    uint32_t wrap_start[kMaxWrapRunes], wrap_end[kMaxWrapRunes];  // Wrap[0], Wrap[1] as lower-cased runes
    uint8_t ascii_code[128]; // 0 = not in the alphabet
    const RuneRange *ranges; // sorted, disjoint
    int32_t n_ranges;
    // ---- term dictionary: packed key -> term id, open addressing, key 0 = empty; one 16-byte entry per slot
    // {key lo, key hi, term id, 0} so that a probe is a single load ----
    const uint4 *term_table;
    uint32_t term_mask;
    uint32_t n_terms;
    // ---- inverted index, CSR in HBM ----
    // Documents are renumbered by (cardinality segment, original id); segment B owns new ids
    // [seg_start[B], seg_start[B+1]).  All postings of one term are contiguous, ordered by
    // (segment, new id): list (term t, segment B) = postings[list_off[t*(S+1)+B] .. list_off[t*(S+1)+B+1]).
    uint32_t n_segments;     // S = InvertedIndexIndices.Size()
    uint32_t n_docs;
    uint32_t id_base;        // added to every returned id (record-id-range shards)
    const uint32_t *seg_start;  // S + 1
    const uint32_t *list_off;   // n_terms * (S + 1)
    const uint32_t *postings;   // 16-byte aligned, padded with 4 trailing entries
    const uint32_t *perm;       // new id -> original document id (without id_base); n_ids entries, holes are never read
    // ---- bucket bitmaps (sg_bitmap.cu) ----
    // Segment starts are aligned to 2^bshift new ids, so a bucket (2^bshift consecutive new ids) lies inside one
    // segment.  Row t holds one bit per bucket: set iff term t has a posting in that bucket.  Rows are row_words
    // 32-bit words (a multiple of 64: whole tiles of the search kernel) apart; row n_terms is all zero (padding lists of a query).
    uint32_t n_ids;          // seg_start[S]: new ids including the alignment holes
    uint32_t bshift;         // log2(new ids per bucket), 0..kMaxBucketShift
    uint32_t row_words;      // 0: the index has no bitmaps (over the memory budget) and is searched by sg_search_kernel
    const uint32_t *bitmaps; // (n_terms + 1) * row_words
    // ---- exact level under the bucket bitmaps (sg_fine.cu), read by sg_resolve_kernel ----
    // Every set bit of the bitmaps is a (term, bucket) pair; pairs are numbered in (term, bucket) order.  rank4[g] = number
    // of pairs before the g-th group of four bitmap words (groups never straddle rows: row_words is a multiple of 64), and
    // fine holds, for pair p, one bit per document of the bucket (2^bshift bits at bit offset p << bshift): set iff the
    // term has that document.  A flagged bucket is therefore resolved with one 16-byte load (the group) and one load of
    // the pair's bits per list, without touching the posting lists.  nullptr: not built (bshift = 0 needs neither).
    const uint32_t *rank4;
    const uint32_t *fine;
};

// One per query, written by sg_plan_kernel and read by sg_search_kernel: kPlanStride bytes =
// [QueryPlan (32) | thresholds of the window segments b_min + i (256) | posting runs {first, end} per list (128 x 8)]
struct QueryPlan {
    uint32_t flags;      // 1: more than 128 n-grams (SG_ERR_QUERY_TOO_LONG)
    int32_t size_a;      // len(tokens), suggester.go:53
    int32_t b_min;       // MinY
    int32_t b_lo, b_hi;  // first / last admissible non-empty segment; b_hi < b_lo: nothing to search
    int32_t n_lists;     // posting runs to read; 0: nothing to search
    int32_t shift;       // log2(documents per counter bucket)
    int32_t reserved;
};
constexpr uint32_t kPlanThrOffset = 32, kPlanRunsOffset = 32 + 256;
constexpr uint32_t kPlanStride = kPlanRunsOffset + kMaxQueryTokens * 8;

// Bitmap engine: one per query, written by sg_tokens_kernel and read by sg_bitmap_search_kernel:
// kTokStride bytes = [TokenPlan (32) | term id of every list to open (128 x 4)]
struct TokenPlan {
    uint32_t flags;      // bit 0: more than 128 n-grams (SG_ERR_QUERY_TOO_LONG); bit 1 (kPlanDirty): the count / resolve
                         // pipeline ran out of scratch for this query, sg_bitmap_search_kernel answers it instead
    int32_t size_a;      // len(tokens), suggester.go:53
    int32_t n_lists;     // tokens that are terms of the index, with multiplicity
    int32_t reserved;
    uint32_t win_lo, win_hi;  // WindowTables::win[size_a], copied so that the search kernel has it with the header
    uint32_t n_flagged;       // count -> resolve pipeline: bitmap words of the query with a bucket at its threshold (sg_count_kernel)
    uint32_t reserved2;
};
constexpr uint32_t kTokTermsOffset = 32;
constexpr uint32_t kTokStride = kTokTermsOffset + kMaxQueryTokens * 4;
constexpr uint32_t kPlanDirty = 2u;

// ---- count -> resolve pipeline (sg_count_kernel, sg_resolve_kernel): scratch of one launch, sized per query ----
// flagged bitmap words {query, word, buckets that reached their threshold, 0}: sg_count_kernel -> sg_resolve_kernel
// survivors {original id, next node of the query | taken << 31, score (two words)}, a linked list per query
#ifndef SG_FLAGS_PER_QUERY
#define SG_FLAGS_PER_QUERY 64           // (Zipf-lettered Cosine 0.5 flags 19 words per query)
#endif
#ifndef SG_NODES_PER_QUERY
#define SG_NODES_PER_QUERY 8
#endif
constexpr uint32_t kFlagsPerQuery = SG_FLAGS_PER_QUERY, kNodesPerQuery = SG_NODES_PER_QUERY;
constexpr uint32_t kLeanScratchPerQuery = kFlagsPerQuery * 16 + kNodesPerQuery * 16 + 8;  // + pending[q], head[q]
constexpr uint32_t kNilNode = 0xFFFFFFFFu;
constexpr uint32_t kArriveOffPad = 32;       // offsets of chunk c start at c * (chunk_queries + kArriveOffPad): no 128-byte line is shared by two chunks
// counters of one launch, zeroed by sg_tokens_kernel: SearchParams::work_counter points at kWorkWords of them
enum { kWorkQuery = 0, kWorkFallbackQuery = 1, kWorkFlagCursor = 2, kWorkNodeCursor = 3, kWorkDirtyAny = 4, kWorkWords = 8 };

// Per call (metric, similarity, mode are per call): everything that depends on the query only through len(tokens).
// Row a = len(tokens) in 0..kMaxQueryTokens.
struct alignas(8) WordRange {
    uint32_t x, y;       // {first, one past last} bitmap word
};
struct WindowTables {
    uint8_t *seg_thr;    // [129][S]          Threshold(alpha, a, B) of an admissible, non-empty segment of the window, else 0
    uint8_t *word_thr;   // [129][row_words]  smallest such threshold over the segments owning buckets of the bitmap word; 255: none
    WordRange *win;      // [129]             bitmap words with a threshold
};
constexpr uint32_t kWindowRows = kMaxQueryTokens + 1;

// ---- n-gram language model (pkg/lm), sg_lm.cu ----
constexpr int kMaxLmOrder = 8;
constexpr uint32_t kLmInvalidContext = 0xFFFFFFFDu;   // InvalidContextOffset, pkg/lm/ngram_vector.go:31-35
constexpr double kLmUnknownWordScore = -100.0;        // pkg/lm/ngram_model.go:24
// Level i holds the (i+1)-grams: values[] = word << 32 | count ordered by (context, word), containers[] = context << 32 |
// first value of that context; the context of an entry is the position of its prefix in level i-1 (pkg/lm/packed_array.go).
struct DevLm {
    uint32_t order;
    uint32_t n_containers[kMaxLmOrder], n_values[kMaxLmOrder], totals[kMaxLmOrder];
    const uint64_t *containers[kMaxLmOrder], *values[kMaxLmOrder];
};
// What nGramModel.Next(context) leaves for ScoreNext (pkg/lm/ngram_model.go:67-99, scorer_next.go:9-23): the values of the
// context's continuations and the count of the context itself.  valid = 0: no scorer (unknown context or none given).
struct LmContext {
    const uint64_t *vals;   // values of the continuation level
    uint32_t from, to;      // [from, to) = continuations of the context
    uint32_t ctx_count;     // count of the full context (the denominator of ScoreNext)
    uint32_t valid;
};

struct SearchParams {
    const char *q_bytes;
    const uint32_t *q_off;
    uint32_t n_q;
    int32_t metric;
    double alpha;
    uint32_t k;
    uint32_t *out_ids;
    double *out_scores;
    uint32_t *out_counts;
    uint4 *out_packed;        // non-null: rows of 16-byte {id, 0, score} entries (sg_candidate = suggest.Candidate's layout) instead of
                              // out_ids / out_scores: one store - one PCIe write when the rows are page-locked host memory - per entry
    uint32_t *stats;          // optional, 16-byte aligned: {admissible postings, admissible lists, 32-bit words the engine reads for the count, 0} per query
    uint32_t *work_counter;   // kWorkWords counters of the launch (bitmap engine: zeroed by sg_tokens_kernel; scan-count engine: word 0, zeroed before launch)
    uint8_t *plans;           // n_q * kPlanStride bytes of scratch
    uint32_t tbl_bytes;       // per-warp count table size (power of two)
    uint32_t warp_smem;       // bytes of shared memory owned by one warp
    int32_t force_shift;      // < 0: cost model picks the bucket width; otherwise log2(bucket width)
    int32_t mode;             // 0: Suggest; 1: Autocomplete (no tail wrap, every token required, lowest ids win)
    WindowTables wt;          // bitmap engine only
    const LmContext *lm_ctx;  // mode 1 only, optional: rank the completions by the language model (spellchecker collector)
    uint32_t *too_long_flag;  // optional: set to 1 if any query of the launch is reported as SG_COUNT_UNSUPPORTED
    int32_t sparse_rows;      // 1: write only the out_counts[q] valid entries of a row (rows in page-locked host memory)
    // ---- count -> resolve pipeline (bitmap engine, Suggest top-k) ----
    uint4 *lean_flags;        // n_q * kFlagsPerQuery entries; nullptr: the launch runs sg_bitmap_search_kernel only
    uint4 *lean_nodes;        // n_q * kNodesPerQuery entries
    uint32_t flag_cap, node_cap;  // entries of the two lists this launch may use (<= what is allocated; the tests shrink them)
    uint32_t *lean_pending;   // [n_q] flagged words of the query not yet resolved
    uint32_t *lean_head;      // [n_q] first survivor node of the query, kNilNode: none
    int32_t only_dirty;       // sg_bitmap_search_kernel: answer only the queries marked kPlanDirty (and exit at once if none is)
    // ---- queries that arrive while the kernel runs (sg_search_batch, page-locked rows) ----
    const uint32_t *arrived;  // nullptr: everything is there.  Else: chunks copied so far = *arrived - arrive_base
    uint32_t arrive_base;
    uint32_t chunk_queries;   // queries per chunk; != 0 also means: q_off holds n + 1 offsets per chunk, kArriveOffPad entries apart
    uint32_t debug;           // SG_RESOLVE_DEBUG (experiments only; results are wrong with any bit set): 1 skip the pair bits, 2 skip
                              // survivors, 4 stop behind the plan loads
    uint8_t *tk_global;       // k > kSmemTopK: 12 * k bytes per warp of the launch for its sorted top-k (scores, then ids); else nullptr
    // ---- sg_candidates_batch (bitmap engine): every candidate of the T-occurrence count instead of a top-k ----
    const uint8_t *custom_thr;        // optional [129][S]: Threshold(alpha, a, B) tabulated by the caller for a metric.Metric that
                                      // is not built in (0 outside [MinY, MaxY]); replaces metric / alpha in sg_window_kernel
    unsigned long long *cand_total;   // non-null: collect mode; candidates found so far (may exceed cand_cap)
    unsigned long long cand_cap;      // entries the four arrays below hold
    uint32_t *cand_query, *cand_ids, *cand_overlap, *cand_segment;  // MergeCandidate (pkg/merger/list_merger.go:33-48) + its query and segment
};

// queries of more than kMaxQueryTokens n-grams, tokenized on the host (sg_long.cu)
struct LongParams {
    const uint64_t *keys;       // packed term keys of every token, query after query (duplicates after normalisation kept)
    const uint32_t *key_off;    // n_long + 1
    uint32_t n_long;
    uint32_t *terms;            // scratch, one per key: term id or kNoTerm
    uint32_t *counters;         // scratch, n_ids per CTA of the launch
    int32_t metric, mode;       // mode 1: Autocomplete
    double alpha;
    uint32_t k;                 // <= kSmemTopK
    uint32_t *out_ids;          // [n_long][k]
    double *out_scores;
    uint32_t *out_counts;       // [n_long]
};

SG_HD static inline uint64_t mix64(uint64_t x) {  // splitmix64 finaliser; host and device hash term keys with it
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

}  // namespace sg
