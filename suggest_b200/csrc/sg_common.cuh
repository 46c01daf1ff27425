// sg_common.cuh — device helpers shared by the kernels of libsuggest_b200: the pkg/metric arithmetic, the query
// tokenizer (pkg/suggest/tokenizer.go:9-20 chain), the per-warp sorted top-k and the candidate emit.
#pragma once
#include <cuda_runtime.h>

#include <cfloat>
#include <climits>

#include "sg_device.h"

namespace sg {

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint32_t kInf = 0xFFFFFFFFu;
constexpr int kThrCache = 256;  // window segments whose threshold is cached in shared memory
constexpr uint32_t kSlicePostings = kSliceBytes / 4;
constexpr int kPlanThreads = 256;  // sg_plan_kernel: 8 queries per CTA

enum { kJaccard = 0, kCosine = 1, kDice = 2, kOverlap = 3, kExact = 4, kAutocomplete = 5 };

// ---------------- pkg/metric, float64, one rounding per reference operation ----------------
__device__ __forceinline__ int f2i(double v) {  // int(float64) with saturation; callers clamp anyway
    if (!(v < 2147483647.0)) return 2147483647;
    if (v < -2147483647.0) return -2147483647;
    return (int)v;
}

__device__ int metric_min_y(int m, double a, int size) {
    switch (m) {
    case kJaccard: return f2i(ceil(__dmul_rn(a, (double)size)));                                   // jaccard.go:13
    case kCosine: return f2i(ceil(__dmul_rn(__dmul_rn(a, a), (double)size)));                      // cosine.go:13
    case kDice: return f2i(ceil(__dmul_rn(__ddiv_rn(a, __dsub_rn(2.0, a)), (double)size)));        // dice.go:13
    case kOverlap: return 1;                                                                       // overlap.go:13
    default: return size;                                                                          // exact.go:11
    }
}

__device__ int metric_max_y(int m, double a, int size) {
    switch (m) {
    case kJaccard: return f2i(floor(__ddiv_rn((double)size, a)));
    case kCosine: return f2i(floor(__ddiv_rn((double)size, __dmul_rn(a, a))));
    case kDice: return f2i(floor(__dmul_rn(__ddiv_rn(__dsub_rn(2.0, a), a), (double)size)));
    case kOverlap: return 32767;  // math.MaxInt16
    case kAutocomplete: return 2147483647;  // every segment from len(terms) up, pkg/suggest/autocomplete.go:47
    default: return size;
    }
}

__device__ __noinline__ int metric_threshold(int m, double a, int sa, int sb) {
    switch (m) {
    case kJaccard: return f2i(ceil(__ddiv_rn(__dmul_rn(a, (double)(sa + sb)), __dadd_rn(1.0, a))));
    case kCosine: return f2i(ceil(__dmul_rn(a, __dsqrt_rn((double)((long long)sa * sb)))));  // Go's int is 64-bit: no overflow for long queries
    case kDice: return f2i(ceil(__dmul_rn(__dmul_rn(0.5, a), (double)(sa + sb))));
    case kOverlap: return f2i(ceil(__dmul_rn(a, (double)min(sa, sb))));
    default: return sa;
    }
}

// scorer.go:29-31: 1 - Distance(overlap, sizeA, sizeB)
__device__ __noinline__ double metric_score(int m, int c, int sa, int sb) {
    double d;
    switch (m) {
    case kJaccard: d = __dsub_rn(1.0, __ddiv_rn((double)c, (double)(sa + sb - c))); break;
    case kCosine: d = __dsub_rn(1.0, __ddiv_rn((double)c, __dsqrt_rn((double)((long long)sa * sb)))); break;
    case kDice: d = __dsub_rn(1.0, __ddiv_rn((double)(2 * c), (double)(sa + sb))); break;
    case kOverlap: d = __dsub_rn(1.0, __ddiv_rn((double)c, (double)min(sa, sb))); break;
    default: d = 0.0; break;
    }
    return __dsub_rn(1.0, d);
}

__device__ __forceinline__ bool threshold_admits(int T, int sa, int sb) {  // suggester.go:76
    return T != 0 && T <= sb && T <= sa;
}

// ---------------- small helpers ----------------
__device__ __forceinline__ uint32_t symbol_code(const DevIndex &ix, uint32_t r) {  // normalizer.go:29-33
    if (r < 128) {
        uint32_t c = ix.ascii_code[r];
        return c ? c : ix.pad_code;
    }
    int lo = 0, hi = ix.n_ranges;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (ix.ranges[mid].hi < r) lo = mid + 1; else hi = mid;
    }
    if (lo < ix.n_ranges && r >= ix.ranges[lo].lo) return ix.ranges[lo].base + (r - ix.ranges[lo].lo);
    return ix.pad_code;
}

__device__ __forceinline__ uint32_t term_lookup(const DevIndex &ix, uint64_t key) {
    uint32_t h = (uint32_t)mix64(key) & ix.term_mask;
    for (;;) {
        const uint4 e = __ldg(ix.term_table + h);
        const uint64_t k = (uint64_t)e.y << 32 | e.x;
        if (k == key) return e.z;
        if (k == 0) return kNoTerm;
        h = (h + 1) & ix.term_mask;
    }
}

// first position in [a, b) whose posting is >= x
__device__ __forceinline__ uint32_t lower_bound(const uint32_t *__restrict__ postings, uint32_t a, uint32_t b, uint32_t x) {
    while (a < b) {
        uint32_t mid = a + ((b - a) >> 1);
        if (__ldg(postings + mid) < x) a = mid + 1; else b = mid;
    }
    return a;
}

// Go `for range` decoding step (utf8.DecodeRune acceptance), invalid byte -> U+FFFD, width 1
// Query bytes and offsets are read past L1 (.cg): sg_search_batch lets them arrive by DMA while the kernel is already
// running (SearchParams::arrived), and a line cached before its neighbour chunk landed would be stale.
template <bool kArrive>
__device__ __forceinline__ uint32_t ldq(const uint8_t *p) { return kArrive ? (uint32_t)__ldcg(p) : (uint32_t)*p; }

__device__ __noinline__ int utf8_step(const uint8_t *s, uint32_t len, uint32_t *rune) {
    uint32_t b0 = s[0];  // (plain loads: non-ASCII chunks of an arriving batch live in a 128-byte aligned area of their own)
    if (b0 < 0x80) { *rune = b0; return 1; }
    int need;
    uint32_t r, lo = 0x80, hi = 0xBF;
    if (b0 >= 0xC2 && b0 <= 0xDF) { need = 1; r = b0 & 0x1F; }
    else if (b0 >= 0xE0 && b0 <= 0xEF) { need = 2; r = b0 & 0x0F; if (b0 == 0xE0) lo = 0xA0; if (b0 == 0xED) hi = 0x9F; }
    else if (b0 >= 0xF0 && b0 <= 0xF4) { need = 3; r = b0 & 0x07; if (b0 == 0xF0) lo = 0x90; if (b0 == 0xF4) hi = 0x8F; }
    else { *rune = 0xFFFD; return 1; }
    if (len < (uint32_t)need + 1) { *rune = 0xFFFD; return 1; }
    for (int i = 1; i <= need; i++) {
        uint32_t b = s[i];
        uint32_t l = i == 1 ? lo : 0x80, h = i == 1 ? hi : 0xBF;
        if (b < l || b > h) { *rune = 0xFFFD; return 1; }
        r = (r << 6) | (b & 0x3F);
    }
    *rune = r;
    return need + 1;
}

// sg_search_batch copies the queries to the device in chunks while the kernel that tokenizes them is already running; after
// every chunk a counter in device memory is bumped (a 4-byte copy on the same copy stream, so it lands behind the chunk).
// A warp waits here until the chunk of its query has arrived.  Copies do not depend on kernels: no deadlock.
// Returns false if the chunk did not arrive within ~2 s (a failed copy the host could not report in time): the query is
// then answered as empty and the host turns the call into an error (too_long_flag[1]) - a kernel must never spin for good.
// s_seen: per-CTA copy of the counter in shared memory, so that a warp whose chunk is known to be there does not touch
// the counter's L2 line at all (thousands of warps polling one line every few hundred nanoseconds is a hot spot of its own).
__device__ __noinline__ int wait_for_chunk(const uint32_t *arrived, uint32_t need, uint32_t *gave_up_flag, volatile uint32_t *s_seen) {
    if ((int32_t)(*s_seen - need) >= 0) return 1;
    const long long t0 = clock64();
    for (;;) {
        const uint32_t now = *(const volatile uint32_t *)arrived;
        if ((int32_t)(now - *s_seen) > 0) *s_seen = now;
        if ((int32_t)(now - need) >= 0) return 1;
        if (clock64() - t0 > 4000000000ll) {
            if (gave_up_flag != nullptr) gave_up_flag[1] = 1u;
            return 0;
        }
        __nanosleep(1000);
    }
}

__device__ __forceinline__ bool wait_for_query(const SearchParams &p, uint32_t q, int lane, volatile uint32_t *s_seen) {
    int ok = 1;
    if (lane == 0) ok = wait_for_chunk(p.arrived, p.arrive_base + q / p.chunk_queries + 1u, p.too_long_flag, s_seen);
    ok = __shfl_sync(kFull, ok, 0);
    __threadfence();
    return ok != 0;
}

// The tokenizer chain of pkg/suggest/tokenizer.go:9-20 for query q of the batch, by one converged warp: wrap ->
// lower -> trim -> rune n-gram windows with first-occurrence dedupe -> alphabet normalisation -> term lookup.
// s_runes[kMaxRunes], s_lterm[kMaxQueryTokens] (out: term id of every token that is a term of the index, duplicates
// after normalisation stay separate lists as in the reference), s_hash[kMaxQueryTokens] are per-warp shared scratch.
// *size_a = len(tokens) (suggester.go:53).  Returns true if the query has more than kMaxQueryTokens n-grams.
// s_ascii (optional): a shared-memory copy of ix.ascii_code, which enables the fast path below.
// kKeys = true (documents, sg_gpubuild.cu): no term lookup; the packed key of every token goes to s_keys[kMaxQueryTokens]
// in token order and *n_lists_out = len(tokens).
template <bool kKeys = false, bool kArrive = false>
__device__ __forceinline__ bool tokenize_query(const DevIndex &ix, const SearchParams &p, uint32_t q, uint32_t *s_runes,
                                               uint32_t *s_lterm, uint32_t *s_hash, int lane, int *size_a_out, int *n_lists_out,
                                               const uint8_t *s_ascii = nullptr, uint64_t *s_keys = nullptr) {
    bool unsupported = false;
    int size_a = 0;
    // chunked arrival: every chunk of chunk_queries queries has its own n + 1 offsets, kArriveOffPad entries apart
    uint32_t qb, qe;
    if (kArrive) {
        const uint32_t qi = q + (q / p.chunk_queries) * kArriveOffPad;
        qb = __ldcg(p.q_off + qi);
        qe = __ldcg(p.q_off + qi + 1);
    } else {
        qb = __ldg(p.q_off + q);
        qe = __ldg(p.q_off + q + 1);
    }
    const uint32_t qlen = qe - qb;
    const uint8_t *qp = (const uint8_t *)p.q_bytes + qb;
    const int nws = ix.n_wrap_start, nwe = p.mode == 1 ? 0 : ix.n_wrap_end;  // NewAutocompleteTokenizer: no tail wrap
    if (s_ascii != nullptr && ix.wrap_ascii && nws + qlen + nwe <= 32u) {
        // Fast path, the common case: an all-ASCII text of at most 32 runes, one rune per lane.  Trim is a ballot, a raw
        // n-gram window is at most 8 bytes, so first-occurrence dedupe (appendUnique, ngram_tokenizer.go:46-54) is one
        // exact MATCH on the packed window; bytes = runes for the early-out of ngram_tokenizer.go:18.
        const int nr = nws + (int)qlen + nwe;
        uint32_t r = ' ';
        if (lane < nws) r = ix.wrap_start[lane];
        else if (lane < nws + (int)qlen) r = ldq<kArrive>(qp + lane - nws);
        else if (lane < nr) r = ix.wrap_end[lane - nws - (int)qlen];
        if (!__any_sync(kFull, r >= 0x80u)) {
            if (r >= 'A' && r <= 'Z') r += 32;
            s_runes[lane] = r;
            __syncwarp();
            const unsigned nonspace = __ballot_sync(kFull, lane < nr && r != ' ');
            int n_lists = 0;
            if (nonspace != 0u) {
                const int f = __ffs(nonspace) - 1, R = 32 - __clz(nonspace) - f;
                const int n_win = R >= ix.n ? R - ix.n + 1 : 0;
                const bool active = lane < n_win;
                unsigned long long raw = 1ull << 63 | (unsigned)lane;  // ASCII windows never set bit 63
                uint64_t key = 0;
                if (active) {
                    raw = 0;
                    for (int cpos = 0; cpos < ix.n; cpos++) {
                        const uint32_t ch = s_runes[f + lane + cpos];
                        raw |= (unsigned long long)ch << (8 * cpos);
                        const uint32_t code = s_ascii[ch];
                        key |= (uint64_t)(code ? code : ix.pad_code) << (ix.bits * cpos);
                    }
                }
                const unsigned same = __match_any_sync(kFull, raw);  // every lane takes part: no short-circuit around it
                const bool keep = active && (same & ((1u << lane) - 1u)) == 0u;
                const unsigned km = __ballot_sync(kFull, keep);
                size_a = __popc(km);
                if (kKeys) {
                    if (keep) s_keys[__popc(km & ((1u << lane) - 1u))] = key;
                    n_lists = size_a;
                } else {
                    const uint32_t term = keep ? term_lookup(ix, key) : kNoTerm;
                    const unsigned tm = __ballot_sync(kFull, term != kNoTerm);
                    if (term != kNoTerm) s_lterm[__popc(tm & ((1u << lane) - 1u))] = term;
                    n_lists = __popc(tm);
                }
            }
            __syncwarp();
            *size_a_out = size_a;
            *n_lists_out = n_lists;
            return false;
        }
    }
    bool nonascii = false;
    for (uint32_t i = lane; i < qlen; i += 32) nonascii |= ldq<kArrive>(qp + i) >= 0x80;
    nonascii = __any_sync(kFull, nonascii);
    int nq_runes = 0;
    if (!nonascii) {
        if (nws + qlen + nwe > (uint32_t)kMaxRunes) unsupported = true;
        else {
            for (uint32_t i = lane; i < qlen; i += 32) {
                uint32_t ch = ldq<kArrive>(qp + i);
                s_runes[nws + i] = (ch >= 'A' && ch <= 'Z') ? ch + 32 : ch;
            }
            nq_runes = (int)qlen;
        }
    } else {
        // queries holding non-ASCII bytes arrive lower-cased (sg_search_batch does it on the host)
        int cnt = 0;
        if (lane == 0) {
            uint32_t i = 0;
            while (i < qlen) {
                uint32_t r;
                i += (uint32_t)utf8_step(qp + i, qlen - i, &r);
                if (nws + cnt + nwe >= kMaxRunes) { cnt = -1; break; }
                s_runes[nws + cnt++] = (r >= 'A' && r <= 'Z') ? r + 32 : r;
            }
        }
        cnt = __shfl_sync(kFull, cnt, 0);
        if (cnt < 0) unsupported = true; else nq_runes = cnt;
    }
    int n_win = 0, wlen = 0, first = 0;
    if (!unsupported) {
        if (lane < nws) s_runes[lane] = ix.wrap_start[lane];
        if (lane < nwe) s_runes[nws + nq_runes + lane] = ix.wrap_end[lane];
        __syncwarp();
        const int nr = nws + nq_runes + nwe;
        int f = INT_MAX, l = -1, bytes = 0;
        for (int i = lane; i < nr; i += 32) {  // strings.Trim(text, " ")
            if (s_runes[i] != ' ') { f = min(f, i); l = max(l, i); }
        }
        f = __reduce_min_sync(kFull, f);
        l = __reduce_max_sync(kFull, l);
        if (l >= 0) {
            for (int i = f + lane; i <= l; i += 32) {
                uint32_t r = s_runes[i];
                bytes += r < 0x80 ? 1 : r < 0x800 ? 2 : r < 0x10000 ? 3 : 4;
            }
        }
        bytes = __reduce_add_sync(kFull, bytes);
        const int R = l >= 0 ? l - f + 1 : 0;
        first = f;
        if (R > 0 && bytes >= ix.n) {  // ngram_tokenizer.go:18: the early-out compares bytes
            if (R < ix.n) { n_win = 1; wlen = R; } else { n_win = R - ix.n + 1; wlen = ix.n; }
        }
        if (n_win > kMaxQueryTokens) { unsupported = true; n_win = 0; }
    }
    int n_lists = 0;
    for (int i = lane; i < n_win; i += 32) {
        uint32_t h = 2166136261u;
        for (int cpos = 0; cpos < wlen; cpos++) h = (h ^ s_runes[first + i + cpos]) * 16777619u;
        s_hash[i] = h;
    }
    __syncwarp();
    for (int base = 0; base < n_win; base += 32) {
        const int i = base + lane;
        bool keep = i < n_win;
        // appendUnique, ngram_tokenizer.go:46-54: raw windows, first occurrence wins (hash first, runes on a hit)
        const uint32_t my_hash = keep ? s_hash[i] : 0u;
        // windows of earlier groups of 32, then the windows of this group that share the hash (one MATCH instead of a loop)
        for (int j = 0; j < base; j++) {
            if (keep && s_hash[j] == my_hash) {
                bool eq = true;
                for (int cpos = 0; cpos < wlen; cpos++) eq &= s_runes[first + i + cpos] == s_runes[first + j + cpos];
                keep = !eq;
            }
        }
        const unsigned long long tag = keep ? (unsigned long long)my_hash : (1ull << 32 | (unsigned)lane);
        unsigned earlier = __match_any_sync(kFull, tag) & ((1u << lane) - 1u);
        while (keep && earlier) {
            const int j = base + __ffs(earlier) - 1;
            earlier &= earlier - 1;
            bool eq = true;
            for (int cpos = 0; cpos < wlen; cpos++) eq &= s_runes[first + i + cpos] == s_runes[first + j + cpos];
            keep = !eq;
        }
        uint32_t term = kNoTerm;
        uint64_t key = 0;
        if (keep) {
            for (int cpos = 0; cpos < wlen; cpos++)
                key |= (uint64_t)symbol_code(ix, s_runes[first + i + cpos]) << (ix.bits * cpos);
            if (!kKeys) term = term_lookup(ix, key);
        }
        const unsigned km = __ballot_sync(kFull, keep);
        size_a += __popc(km);
        if (kKeys) {
            if (keep) s_keys[n_lists + __popc(km & ((1u << lane) - 1u))] = key;
            n_lists += __popc(km);
        } else {
            const unsigned tm = __ballot_sync(kFull, term != kNoTerm);
            if (term != kNoTerm) s_lterm[n_lists + __popc(tm & ((1u << lane) - 1u))] = term;
            n_lists += __popc(tm);
        }
    }
    __syncwarp();

    *size_a_out = size_a;
    *n_lists_out = n_lists;
    return unsupported;
}

__device__ __forceinline__ float ln_factorial(float n) {  // Stirling, good enough for the cost model
    if (n < 2.0f) return 0.0f;
    return n * __logf(n) - n + 0.5f * __logf(6.2831853f * n) + 1.0f / (12.0f * n);
}

// Per-query state that lives in registers, identical in every lane of the warp.
struct QueryCtx {
    int metric;
    double alpha;
    int size_a;
    int b_lo, b_hi;   // admissible, non-empty segment range
    uint32_t k;
    int tk_len;
    double *tk_score;
    uint32_t *tk_id;
};

// sorted insert into the warp's top-k (best first); all lanes call with identical arguments
__device__ __noinline__ void topk_insert(QueryCtx &c, double score, uint32_t id, int lane) {
    const int k = (int)c.k;
    if (c.tk_len == k) {
        double ws = c.tk_score[k - 1];
        uint32_t wi = c.tk_id[k - 1];
        if (!(score > ws || (score == ws && id < wi))) return;  // Candidate.Less, collector.go:20-26
    }
    int better = 0;
    for (int j = lane; j < c.tk_len; j += 32) {
        double s = c.tk_score[j];
        better += (s > score || (s == score && c.tk_id[j] < id)) ? 1 : 0;
    }
    const int pos = __reduce_add_sync(kFull, better);
    const int new_len = min(c.tk_len + 1, k);
    for (int hi = new_len - 1; hi > pos; hi -= 32) {
        int j = hi - lane;
        bool act = j > pos;
        double sv = 0.0;
        uint32_t iv = 0;
        if (act) { sv = c.tk_score[j - 1]; iv = c.tk_id[j - 1]; }
        __syncwarp();
        if (act) { c.tk_score[j] = sv; c.tk_id[j] = iv; }
        __syncwarp();
    }
    __syncwarp();  // the reads of the ranking loop above are ordered before this write (racecheck cannot see it through the reduce)
    if (lane == 0) { c.tk_score[pos] = score; c.tk_id[pos] = id; }
    __syncwarp();
    c.tk_len = new_len;
}

// A document (new id) of segment size_b with an exact overlap count: apply the segment's threshold T,
// score, offer to the top-k.  All lanes call with identical arguments.
__device__ __forceinline__ void emit_candidate(const DevIndex &ix, QueryCtx &c, uint32_t new_id, int count, int size_b, int T,
                                               int lane) {
    if (count < T) return;
    const uint32_t id = __ldg(ix.perm + new_id);
    // FirstKCollectorManager.Collect scores a position with -position (pkg/suggest/collector.go:104-106)
    const double score = c.metric == kAutocomplete ? -(double)(ix.id_base + id) : metric_score(c.metric, count, c.size_a, size_b);
    topk_insert(c, score, id, lane);
}

// A bucket whose counter reached its segment's threshold: find the documents of [blo, bhi) exactly.  Every lane
// binary-searches the runs it owns (lists lane, lane + 32, ...) for the range, then the warp repeatedly takes the smallest
// id and counts the lists holding it.  Rare (about 1.1 buckets per query on config #2, the true match included), so out of line.
__device__ __noinline__ void resolve_bucket(const DevIndex &ix, QueryCtx &c, const uint32_t *__restrict__ postings,
                                            const uint32_t *s_cur, const uint32_t *s_end, int n_lists, uint32_t blo,
                                            unsigned long long bhi64, int size_b, int T, int lane) {
    uint32_t pp[4], pe[4], vv[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const int j = lane + 32 * g;
        vv[g] = kInf;
        pp[g] = pe[g] = 0;
        if (j < n_lists) {
            pe[g] = s_end[j];
            pp[g] = lower_bound(postings, s_cur[j], pe[g], blo);
            if (pp[g] < pe[g]) {
                const uint32_t x2 = __ldg(postings + pp[g]);
                if ((unsigned long long)x2 < bhi64) vv[g] = x2;
            }
        }
    }
    for (;;) {
        const uint32_t mine = min(min(vv[0], vv[1]), min(vv[2], vv[3]));
        const uint32_t m = __reduce_min_sync(kFull, mine);
        if (m == kInf) break;
        int cnt = 0;
#pragma unroll
        for (int g = 0; g < 4; g++) cnt += vv[g] == m;
        cnt = __reduce_add_sync(kFull, cnt);
#pragma unroll
        for (int g = 0; g < 4; g++) {
            if (vv[g] == m) {
                vv[g] = kInf;
                if (++pp[g] < pe[g]) {
                    const uint32_t x2 = __ldg(postings + pp[g]);
                    if ((unsigned long long)x2 < bhi64) vv[g] = x2;
                }
            }
        }
        emit_candidate(ix, c, m, cnt, size_b, T, lane);
    }
}

// P(Poisson(lam) >= t), upper-ish estimate for the bucket-width cost model
__device__ __forceinline__ float poisson_tail(float lam, int t) {
    if (lam <= 0.0f) return 0.0f;
    const float ft = (float)t;
    if (lam >= 0.7f * ft) return 1.0f;
    return fminf(1.0f, __expf(-lam + ft * __logf(lam) - ln_factorial(ft)) / (1.0f - lam / (ft + 1.0f)));
}

}  // namespace

}  // namespace sg
