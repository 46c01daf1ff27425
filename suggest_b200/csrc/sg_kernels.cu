// sg_kernels.cu — sm_100a kernels of libsuggest_b200.
//
// sg_search_kernel: one warp owns one query from tokenisation to its k results (persistent warps
// pull query numbers from a global counter).  What the reference does per query with goroutines,
// lazily decoded posting-list iterators, CPMerge and a heap
//   (pkg/suggest/suggester.go:46-131, pkg/index/searcher.go:28-78, pkg/merger/cp_merge.go:19-120,
//    pkg/suggest/collector.go:117-191, pkg/suggest/topk.go:66-175)
// becomes:
//   1. tokenise in shared memory (pkg/suggest/tokenizer.go:9-20 chain), look the n-grams up in the
//      term hash table -> one contiguous posting run per query token (documents are renumbered by
//      cardinality segment, so the segment window [MinY, MaxY] of a term is one slice of HBM);
//   2. T-occurrence count (ScanCount semantics, pkg/merger/scan_count.go:14-88) in a warp-private
//      table of byte counters in shared memory, updated with 32-bit shared atomics (ATOMS.ADD on the
//      byte's field; measured on B200 at the cost of a plain scattered store, 2.7x cheaper than a
//      byte load + store).  The table covers the window's id range at a bucket width of 2^s
//      documents chosen per query by a cost model: s = 0 counts documents exactly (dense
//      dictionaries); s > 0 counts "lists that hit the bucket" - every (list, bucket) pair adds one,
//      so a counter is bounded by the number of lists (<= 128) and bounds every overlap inside the
//      bucket - then the table is scanned segment by segment against that segment's threshold and
//      the few surviving buckets are resolved exactly by a warp-wide merge of the run slices that
//      fall into them;
//   3. threshold / score in float64 with the reference's operation order (pkg/metric/*.go,
//      pkg/suggest/scorer.go:29-31) and a sorted top-k in shared memory ordered (score desc, id asc)
//      (pkg/suggest/collector.go:20-26).
//
// sg_merge_topk_kernel: k best of the per-shard top-k lists for record-id-range shards.
#include <mutex>

#include "sg_common.cuh"
#include "sg_kernels.h"

namespace sg {

namespace {


// ---------------- shared-memory staging: TMA bulk copies + mbarriers, red.shared counters ----------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}

// Converged warp: one elected lane arms the mbarrier with the byte count and issues the global -> shared bulk copy
// (UBLKCP); the copy's completion is counted in bytes on the same mbarrier.
__device__ __forceinline__ void tma_issue(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n\t"
        "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t"
        "}" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(mbar)
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(mbar), "r"(parity)
                     : "memory");
    } while (!done);
}

// counter += inc: the byte counters are fields of 32-bit words; ATOMS.ADD without a return value, never branched
// around (ptxas turns a predicated red into a branch, and a divergent branch costs more than an add of zero)
// No "memory" clobber on purpose: the table is only read back after a __syncwarp(), and without the clobber the
// compiler may issue the next group's shared loads ahead of these.
__device__ __forceinline__ void red_add(uint32_t saddr, uint32_t inc) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(saddr), "r"(inc));
}

// Count one group of 128 postings (4 per lane, ascending over the warp): every (list, bucket) pair adds exactly one to
// the bucket's byte counter, so a counter never exceeds the number of lists (<= 128).  `prev` is the bucket of the
// posting just before this lane's four (read from the slice in shared memory, or carried over from the previous slice).
// A posting that is not the first of its bucket adds zero to its own word.
__device__ __forceinline__ void count_group_full(uint32_t tbl_saddr, const uint4 x, uint32_t prev, uint32_t lo_id, int shift) {
    const uint32_t b0 = (x.x - lo_id) >> shift, b1 = (x.y - lo_id) >> shift, b2 = (x.z - lo_id) >> shift, b3 = (x.w - lo_id) >> shift;
    // head << ((bucket & 3) * 8): the funnel shift takes its amount modulo 32
    const uint32_t i0 = __funnelshift_l(0u, (uint32_t)(b0 != prev), b0 << 3), i1 = __funnelshift_l(0u, (uint32_t)(b1 != b0), b1 << 3);
    const uint32_t i2 = __funnelshift_l(0u, (uint32_t)(b2 != b1), b2 << 3), i3 = __funnelshift_l(0u, (uint32_t)(b3 != b2), b3 << 3);
    red_add(tbl_saddr + (b0 & ~3u), i0);
    red_add(tbl_saddr + (b1 & ~3u), i1);
    red_add(tbl_saddr + (b2 & ~3u), i2);
    red_add(tbl_saddr + (b3 & ~3u), i3);
}

// First / last group of a list (about two in seven), some slots lie outside the list's slice [a, b): those add zero to
// this lane's word of a scratch line.  pos = index of this lane's first posting.  Out of line so that the hot loop stays
// small in the instruction cache.
__device__ __noinline__ void count_group_partial(uint32_t tbl_saddr, uint32_t scratch_saddr, const uint4 x, uint32_t prev, uint32_t pos,
                                                 uint32_t a, uint32_t b, uint32_t lo_id, int shift) {
    uint32_t b0 = (x.x - lo_id) >> shift, b1 = (x.y - lo_id) >> shift, b2 = (x.z - lo_id) >> shift, b3 = (x.w - lo_id) >> shift;
    const uint32_t d = pos - a, len = b - a;  // unsigned: positions before a wrap around and fail the test too
    if (d - 1 >= len) prev = kInf;
    if (d >= len) b0 = kInf;
    if (d + 1 >= len) b1 = kInf;
    if (d + 2 >= len) b2 = kInf;
    if (d + 3 >= len) b3 = kInf;
    uint32_t i0 = __funnelshift_l(0u, (uint32_t)(b0 != prev), b0 << 3), i1 = __funnelshift_l(0u, (uint32_t)(b1 != b0), b1 << 3);
    uint32_t i2 = __funnelshift_l(0u, (uint32_t)(b2 != b1), b2 << 3), i3 = __funnelshift_l(0u, (uint32_t)(b3 != b2), b3 << 3);
    uint32_t a0 = tbl_saddr + (b0 & ~3u), a1 = tbl_saddr + (b1 & ~3u), a2 = tbl_saddr + (b2 & ~3u), a3 = tbl_saddr + (b3 & ~3u);
    if (b0 == kInf) { a0 = scratch_saddr; i0 = 0u; }
    if (b1 == kInf) { a1 = scratch_saddr; i1 = 0u; }
    if (b2 == kInf) { a2 = scratch_saddr; i2 = 0u; }
    if (b3 == kInf) { a3 = scratch_saddr; i3 = 0u; }
    red_add(a0, i0);
    red_add(a1, i1);
    red_add(a2, i2);
    red_add(a3, i3);
}

// Walks the posting-run slices of one chunk in order, kSlicePostings at a time (the TMA producer's cursor).
struct SliceWalker {
    int j;               // list
    uint32_t a, b, base; // slice of the list inside the chunk, first posting (multiple of 4) of the current piece
    bool valid;
    __device__ __forceinline__ void start(const uint32_t *s_cur, const uint32_t *s_end, int n_lists) {
        j = -1;
        a = b = base = 0;
        valid = true;
        next(s_cur, s_end, n_lists);
    }
    __device__ __forceinline__ void next(const uint32_t *s_cur, const uint32_t *s_end, int n_lists) {
        if (j >= 0 && base + kSlicePostings < b) { base += kSlicePostings; return; }
        for (++j; j < n_lists; ++j) {
            a = s_cur[j];
            b = s_end[j];
            if (a < b) { base = a & ~3u; return; }
        }
        valid = false;
    }
    __device__ __forceinline__ uint32_t bytes() const {  // whole 16-byte units; the postings array is padded
        const uint32_t end = min(base + kSlicePostings, (b + 3u) & ~3u);
        return (end - base) * 4u;
    }
};

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// sg_plan_kernel: steps 1-4 for every query (tokenise, segment window + thresholds, posting runs, bucket width),
// written as one QueryPlan per query.  Light on shared memory, so it runs at full occupancy and its dependent global
// loads (query bytes -> term hash -> list offsets) hide behind other warps; keeping this code out of sg_search_kernel
// also keeps that kernel's hot loop resident in the instruction cache.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPlanThreads) sg_plan_kernel(const DevIndex ix, const SearchParams p) {
    __shared__ __align__(16) uint32_t s_scratch[kPlanThreads / 32][kMaxRunes + 2 * kMaxQueryTokens + kThrCache / 4];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint32_t *s_runes = s_scratch[warp];                        // [kMaxRunes]
    uint32_t *s_lterm = s_runes + kMaxRunes;                    // [128] term id of every list to open
    uint32_t *s_hash = s_lterm + kMaxQueryTokens;               // [128] hash of every raw n-gram window
    uint8_t *s_thr = (uint8_t *)(s_hash + kMaxQueryTokens);     // [256] threshold of window segment b_min + i, 0 = skip
    const uint32_t S = ix.n_segments;
    const uint32_t stride = S + 1;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;

    for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < p.n_q; q += n_warps) {
        QueryCtx c;
        c.metric = p.mode == 1 ? (int)kAutocomplete : p.metric;
        c.alpha = p.alpha;
        c.size_a = 0;
        c.b_lo = 0;
        c.b_hi = -1;
        bool unsupported = false;
        uint32_t st_postings = 0, st_lists = 0;
        for (int i = lane; i < kThrCache / 4; i += 32) ((uint32_t *)s_thr)[i] = 0u;
        __syncwarp();

        // ---------------- 1. tokenise: wrap -> lower -> trim -> n-gram windows (dedupe) -> normalise ----------------
        int n_lists = 0;
        unsupported = tokenize_query(ix, p, q, s_runes, s_lterm, s_hash, lane, &c.size_a, &n_lists);

        // ---------------- 2. segment window and thresholds (suggester.go:53-59, :73-78) ----------------
        int b_min = 0;
        if (c.size_a > 0) {
            b_min = max(metric_min_y(c.metric, c.alpha, c.size_a), 0);
            int b_max = metric_max_y(c.metric, c.alpha, c.size_a);
            if (b_max >= (int)S) b_max = (int)S - 1;
            int lo = INT_MAX, hi = -1;
            for (int B = b_min + lane; B <= b_max; B += 32) {
                const int T = metric_threshold(c.metric, c.alpha, c.size_a, B);
                const bool use = threshold_admits(T, c.size_a, B) && __ldg(ix.seg_start + B + 1) > __ldg(ix.seg_start + B);
                if (use) { lo = min(lo, B); hi = max(hi, B); }
                if (B - b_min < kThrCache) s_thr[B - b_min] = use ? (uint8_t)T : (uint8_t)0;
            }
            c.b_lo = __reduce_min_sync(kFull, lo);
            c.b_hi = __reduce_max_sync(kFull, hi);
            if (p.stats != nullptr) {  // admissible postings and lists
                for (int B = b_min; B <= b_max; B++) {
                    const int T = metric_threshold(c.metric, c.alpha, c.size_a, B);
                    if (!threshold_admits(T, c.size_a, B)) continue;
                    for (int j = lane; j < n_lists; j += 32) {
                        const uint32_t *o = ix.list_off + (size_t)s_lterm[j] * stride + B;
                        const uint32_t len = __ldg(o + 1) - __ldg(o);
                        st_postings += len;
                        st_lists += len != 0;
                    }
                }
                st_postings = __reduce_add_sync(kFull, st_postings);
                st_lists = __reduce_add_sync(kFull, st_lists);
            }
            __syncwarp();
        }
        // threshold of an admissible, non-empty window segment; 0 = nothing to find there
        auto thr_of = [&](int B) -> int {
            if (B - b_min < kThrCache) return (int)s_thr[B - b_min];
            const int T = metric_threshold(c.metric, c.alpha, c.size_a, B);
            return (threshold_admits(T, c.size_a, B) && __ldg(ix.seg_start + B + 1) > __ldg(ix.seg_start + B)) ? T : 0;
        };

        if (p.mode == 1 && n_lists < c.size_a) n_lists = 0;  // a query token that is in no list: nothing can hold them all
        uint8_t *plan_base = p.plans + (size_t)q * kPlanStride;
        QueryPlan *plan = (QueryPlan *)plan_base;
        uint2 *plan_runs = (uint2 *)(plan_base + kPlanRunsOffset);
        int shift = 0;
        if (c.b_hi >= 0 && n_lists > 0) {
            // ---------------- 3. one posting run per list ----------------
            float total = 0.0f;
            for (int j = lane; j < n_lists; j += 32) {
                const uint32_t *o = ix.list_off + (size_t)s_lterm[j] * stride;
                const uint32_t r0 = __ldg(o + c.b_lo), r1 = __ldg(o + c.b_hi + 1);
                plan_runs[j] = make_uint2(r0, r1);
                total += (float)(r1 - r0);
            }
            for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(kFull, total, o);
            __syncwarp();

            // ---------------- 4. bucket width ----------------
            // shift = 0 counts documents exactly but may need many passes over the table; a wider bucket counts
            // "lists that hit the bucket" (an upper bound of every overlap inside it) in one pass and pays an exact
            // merge for the buckets that reach their segment's threshold.  Minimise the estimated cost.
            const uint32_t c_base = __ldg(ix.seg_start + c.b_lo);
            const uint32_t D = __ldg(ix.seg_start + c.b_hi + 1) - c_base;
            const uint32_t NB = p.tbl_bytes;
            if (p.force_shift >= 0) shift = min(p.force_shift, 30);
            else if (D > NB) {
                int s1 = 1;
                while ((((uint64_t)(D - 1) >> s1) + 1) > NB) s1++;
                const float fl = (float)n_lists, fd = (float)D;
                const float chunk_cost = (float)NB * (1.0f / 32.0f) + 400.0f + 30.0f * fl;
                float best = (float)((D + NB - 1) / NB) * chunk_cost;  // shift 0
                for (int s = max(1, s1 - 3); s <= s1; s++) {
                    const float w = (float)(1u << s);
                    const float lam = fl * (1.0f - __expf(-(total / fl) * w / fd));
                    float est = 0.0f;
                    for (int B = c.b_lo + lane; B <= c.b_hi; B += 32) {
                        const int T = thr_of(B);
                        if (T == 0) continue;
                        const float docs = (float)(__ldg(ix.seg_start + B + 1) - __ldg(ix.seg_start + B));
                        est += (docs / w + 1.0f) * poisson_tail(lam, T);
                    }
                    for (int o = 16; o; o >>= 1) est += __shfl_xor_sync(kFull, est, o);
                    const uint32_t nb_total = ((D - 1) >> s) + 1;
                    const float cost = (float)((nb_total + NB - 1) / NB) * chunk_cost + est * 2000.0f;
                    if (cost < best) { best = cost; shift = s; }
                }
            }
            while (shift < 30 && (((uint64_t)(D - 1) >> shift) + NB) / NB > 0xFFFFu) shift++;  // keep the chunk loop bounded
        } else {
            n_lists = 0;
        }
        for (int i = lane; i < kThrCache / 16; i += 32) ((uint4 *)(plan_base + kPlanThrOffset))[i] = ((const uint4 *)s_thr)[i];
        if (lane == 0) {
            plan->flags = unsupported ? 1u : 0u;
            plan->size_a = c.size_a;
            plan->b_min = b_min;
            plan->b_lo = c.b_lo;
            plan->b_hi = c.b_hi;
            plan->n_lists = n_lists;
            plan->shift = shift;
            if (p.stats != nullptr) ((uint4 *)p.stats)[q] = make_uint4(st_postings, st_lists, st_postings, 0u);  // this engine reads the postings themselves
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// sg_search_kernel: steps 5-8 (count, scan, resolve, score / top-k) from the QueryPlans.  Persistent: one CTA per SM,
// one warp per query, query numbers from a global counter.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMaxSearchThreads, 1) sg_search_kernel(const DevIndex ix, const SearchParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t *wsm = smem + (size_t)warp * p.warp_smem;
    uint8_t *tbl = wsm;                                         // [tbl_bytes] byte counters, updated as fields of 32-bit words
    uint8_t *ring = wsm + p.tbl_bytes;                          // [kRingSlots][kSliceBytes] posting slices landed by TMA
    uint64_t *mbar = (uint64_t *)(ring + kRingSlots * kSliceBytes);  // [kRingSlots] one mbarrier per slot (64 bytes) + scratch line (128)
    uint4 *s_meta = (uint4 *)(ring + kRingSlots * kSliceBytes + 192);  // [kRingSlots] {a, b, base} of the slice in each slot (64 bytes)
    uint32_t *s_cur = (uint32_t *)(ring + kRingSlots * kSliceBytes + 256);  // [128] start of the not yet counted part of a run
    uint32_t *s_end = s_cur + kMaxQueryTokens;                  // [128] end of the run slice inside the current chunk
    uint32_t *s_rend = s_end + kMaxQueryTokens;                 // [128] end of the run (segment window)
    uint8_t *s_thr = (uint8_t *)(s_rend + kMaxQueryTokens);     // [256] threshold of window segment b_min + i, 0 = skip
    double *tk_score = (double *)(s_thr + kThrCache);           // [k]
    uint32_t *tk_id = (uint32_t *)(tk_score + p.k);             // [k]

    const uint32_t *__restrict__ postings = ix.postings;
    const uint32_t tbl_saddr = smem_u32(tbl), ring_saddr = smem_u32(ring), mbar_saddr = smem_u32(mbar);
    const uint32_t scratch_saddr = mbar_saddr + 64 + lane * 4;  // this lane's word of the scratch line behind the mbarriers
    if (lane == 0) {
        for (uint32_t sl = 0; sl < kRingSlots; sl++) mbar_init(mbar_saddr + 8 * sl, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t n_filled = 0, n_used = 0;  // slices issued to / consumed from the ring since the kernel started

    for (;;) {
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(p.work_counter, 1u);
        q = __shfl_sync(kFull, q, 0);
        if (q >= p.n_q) break;

        const uint8_t *plan_base = p.plans + (size_t)q * kPlanStride;
        const uint4 h0 = __ldg((const uint4 *)plan_base), h1 = __ldg((const uint4 *)plan_base + 1);  // QueryPlan
        QueryCtx c;
        c.metric = p.mode == 1 ? (int)kAutocomplete : p.metric;
        c.alpha = p.alpha;
        c.k = p.k;
        c.tk_len = 0;
        c.tk_score = tk_score;
        c.tk_id = tk_id;
        c.size_a = (int)h0.y;
        c.b_lo = (int)h0.w;
        c.b_hi = (int)h1.x;
        const bool unsupported = h0.x != 0u;
        const int b_min = (int)h0.z;
        const int n_lists = (int)h1.y;
        const int shift = (int)h1.z;
        // threshold of an admissible, non-empty window segment; 0 = nothing to find there
        auto thr_of = [&](int B) -> int {
            if (B - b_min < kThrCache) return (int)s_thr[B - b_min];
            const int T = metric_threshold(c.metric, c.alpha, c.size_a, B);
            return (threshold_admits(T, c.size_a, B) && __ldg(ix.seg_start + B + 1) > __ldg(ix.seg_start + B)) ? T : 0;
        };

        if (n_lists > 0) {
            for (int i = lane; i < kThrCache / 16; i += 32) ((uint4 *)s_thr)[i] = __ldg((const uint4 *)(plan_base + kPlanThrOffset) + i);
            for (int j = lane; j < n_lists; j += 32) {
                const uint2 r = __ldg((const uint2 *)(plan_base + kPlanRunsOffset) + j);
                s_cur[j] = r.x;
                s_rend[j] = r.y;
            }
            const uint32_t c_base = __ldg(ix.seg_start + c.b_lo);
            const uint32_t D = __ldg(ix.seg_start + c.b_hi + 1) - c_base;
            const uint32_t NB = p.tbl_bytes;
            __syncwarp();
            const uint64_t chunk_docs = (uint64_t)NB << shift;

            for (uint64_t cs = 0; cs < D; cs += chunk_docs) {
                const uint32_t ce = (uint32_t)min((unsigned long long)D, (unsigned long long)(cs + chunk_docs));
                const uint32_t lo_id = c_base + (uint32_t)cs;  // first id of the chunk
                const uint32_t hi_id = c_base + ce;            // one past its last id
                const uint32_t n_buckets = ((ce - (uint32_t)cs - 1) >> shift) + 1;
                // slice of every run inside the chunk
                for (int j = lane; j < n_lists; j += 32)
                    s_end[j] = ce == D ? s_rend[j] : lower_bound(postings, s_cur[j], s_rend[j], hi_id);
                const uint32_t n_vec = (n_buckets + 15) >> 4;
                for (uint32_t w = lane; w < n_vec; w += 32) ((uint4 *)tbl)[w] = make_uint4(0, 0, 0, 0);
                __syncwarp();

                // ---- count: the runs stream through a ring of shared-memory slots filled by TMA bulk copies ----
                {
                    SliceWalker prod;
                    prod.start(s_cur, s_end, n_lists);
                    auto fill = [&]() {  // next slice -> slot n_filled % kRingSlots
                        if (!prod.valid) return;
                        const uint32_t slot = n_filled % kRingSlots;
                        s_meta[slot] = make_uint4(prod.a, prod.b, prod.base, 0u);  // every lane stores the same words
                        tma_issue(ring_saddr + slot * kSliceBytes, postings + prod.base, prod.bytes(), mbar_saddr + 8 * slot);
                        n_filled++;
                        prod.next(s_cur, s_end, n_lists);
                    };
                    for (uint32_t sl = 0; sl < kRingSlots; sl++) fill();
                    uint32_t carry = kInf;
                    while (n_used != n_filled) {
                        const uint32_t slot = n_used % kRingSlots;
                        mbar_wait(mbar_saddr + 8 * slot, (n_used / kRingSlots) & 1u);
                        const uint4 meta = s_meta[slot];
                        const uint32_t a = meta.x, b = meta.y;
                        if (meta.z == (a & ~3u)) carry = kInf;  // first slice of a list
                        const uint32_t end = min(meta.z + kSlicePostings, b);
                        // Groups of 128 postings, lane l takes words 4l..4l+3 of a group.  The posting before a lane's four
                        // comes from the word just below them (bank-conflict 4, but no shuffle and no convergence branch);
                        // lane 0 of the slice's first group takes the carry from the previous slice instead.
                        const uint32_t *w = (const uint32_t *)(ring + slot * kSliceBytes) + 4 * lane;
                        const bool use_carry = lane == 0;
                        uint32_t g = meta.z;
                        {   // first group: the only one that can start before a, and the one that needs the carry
                            const uint4 x = *(const uint4 *)w;
                            uint32_t prev = (w[-1] - lo_id) >> shift;  // lane 0 reads the word below the slot: unused
                            if (use_carry) prev = carry;
                            if (g >= a && g + 128 <= b) count_group_full(tbl_saddr, x, prev, lo_id, shift);
                            else count_group_partial(tbl_saddr, scratch_saddr, x, prev, g + lane * 4, a, b, lo_id, shift);
                            g += 128;
                            w += 128;
                        }
                        for (; g + 128 <= end; g += 128, w += 128)  // groups wholly inside [a, b)
                            count_group_full(tbl_saddr, *(const uint4 *)w, (w[-1] - lo_id) >> shift, lo_id, shift);
                        if (g < end)  // the list ends inside this group
                            count_group_partial(tbl_saddr, scratch_saddr, *(const uint4 *)w, (w[-1] - lo_id) >> shift, g + lane * 4, a, b, lo_id,
                                                shift);
                        carry = (((const uint32_t *)(ring + slot * kSliceBytes))[end - 1 - meta.z] - lo_id) >> shift;
                        __syncwarp();  // every lane has read the slot before it is refilled
                        n_used++;
                        fill();
                    }
                }
                __syncwarp();

                // ---- scan, segment by segment: buckets whose counter reaches the segment's threshold ----
                for (int B = c.b_lo; B <= c.b_hi; B++) {
                    const int T = thr_of(B);
                    if (T == 0) continue;
                    const uint32_t r0 = max(__ldg(ix.seg_start + B), lo_id), r1 = min(__ldg(ix.seg_start + B + 1), hi_id);
                    if (r0 >= r1) continue;
                    const uint32_t bk0 = (r0 - lo_id) >> shift, bk1 = (r1 - 1 - lo_id) >> shift;  // inclusive
                    const uint32_t th4 = (uint32_t)T * 0x01010101u;
                    for (uint32_t w0 = bk0 >> 4; w0 <= (bk1 >> 4); w0 += 32) {
                        const uint32_t w = w0 + lane;
                        uint4 x = make_uint4(0, 0, 0, 0);
                        if (w <= (bk1 >> 4)) x = ((const uint4 *)tbl)[w];
                        const unsigned g0 = __vcmpgeu4(x.x, th4), g1 = __vcmpgeu4(x.y, th4), g2 = __vcmpgeu4(x.z, th4),
                                       g3 = __vcmpgeu4(x.w, th4);
                        unsigned m16 = 0;
                        if ((g0 | g1 | g2 | g3) != 0) {
                            m16 = (((g0 & 0x01010101u) * 0x01020408u) >> 24) & 0xFu;
                            m16 |= ((((g1 & 0x01010101u) * 0x01020408u) >> 24) & 0xFu) << 4;
                            m16 |= ((((g2 & 0x01010101u) * 0x01020408u) >> 24) & 0xFu) << 8;
                            m16 |= ((((g3 & 0x01010101u) * 0x01020408u) >> 24) & 0xFu) << 12;
                            const uint32_t vb = w * 16u;  // first bucket of this vector; keep buckets inside [bk0, bk1]
                            if (vb < bk0) m16 &= ~((1u << (bk0 - vb)) - 1u);
                            if (vb + 15u > bk1) m16 &= (2u << (bk1 - vb)) - 1u;
                        }
                        unsigned bal;
                        while ((bal = __ballot_sync(kFull, m16 != 0)) != 0) {
                            const int src = __ffs(bal) - 1;
                            unsigned mm = __shfl_sync(kFull, m16, src);
                            if (lane == src) m16 = 0;
                            while (mm) {
                                const int bit = __ffs(mm) - 1;
                                mm &= mm - 1;
                                const uint32_t bucket = (w0 + (uint32_t)src) * 16u + (uint32_t)bit;
                                if (shift == 0) {
                                    emit_candidate(ix, c, lo_id + bucket, (int)tbl[bucket], B, T, lane);
                                    continue;
                                }
                                // resolve the bucket exactly: warp-wide merge of the run slices inside its id range
                                const uint32_t blo = max(lo_id + (bucket << shift), r0);
                                const unsigned long long bhi64 =
                                    min((unsigned long long)lo_id + ((unsigned long long)(bucket + 1) << shift), (unsigned long long)r1);
                                resolve_bucket(ix, c, postings, s_cur, s_end, n_lists, blo, bhi64, B, T, lane);
                            }
                        }
                    }
                }
                __syncwarp();
                for (int j = lane; j < n_lists; j += 32) s_cur[j] = s_end[j];
                __syncwarp();
            }
        }

        // ---------------- 5. results: GetCandidates order, fixed stride k ----------------
        const size_t row = (size_t)q * p.k;
        for (uint32_t j = lane; j < p.k; j += 32) {
            const bool has = (int)j < c.tk_len;
            p.out_ids[row + j] = has ? ix.id_base + tk_id[j] : 0u;
            p.out_scores[row + j] = has ? tk_score[j] : 0.0;
        }
        if (lane == 0) {
            p.out_counts[q] = unsupported ? kCountUnsupported : (uint32_t)c.tk_len;
        }
        __syncwarp();
    }
}

// k best of n_parts sorted lists per query, (score desc, id asc).  One warp per query, lane = part.
// Part p's rows start stride_* elements behind part p - 1's (separate [part][query][k] arrays, or one packed block per part).
__global__ void sg_merge_topk_kernel(uint32_t n_parts, uint32_t n_q, uint32_t k, const uint32_t *__restrict__ part_ids,
                                     const double *__restrict__ part_scores, const uint32_t *__restrict__ part_counts,
                                     size_t stride_ids, size_t stride_scores, size_t stride_counts,
                                     uint32_t *out_ids, double *out_scores, uint32_t *out_counts, int sparse) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < n_q; q += warps) {
        uint32_t cur = 0, cnt = 0;
        size_t row = 0;
        bool unsupported = false;
        const uint32_t *my_ids = part_ids + (size_t)lane * stride_ids;
        const double *my_scores = part_scores + (size_t)lane * stride_scores;
        if ((uint32_t)lane < n_parts) {
            cnt = part_counts[(size_t)lane * stride_counts + q];
            row = (size_t)q * k;
            if (cnt == kCountUnsupported) { unsupported = true; cnt = 0; }
            if (cnt > k) cnt = k;
        }
        unsupported = __any_sync(kFull, unsupported);
        uint32_t n_out = 0;
        for (; n_out < k; n_out++) {
            bool has = cur < cnt;
            double s = has ? my_scores[row + cur] : 0.0;
            uint32_t id = has ? my_ids[row + cur] : kInf;
            int who = lane;
            for (int o = 16; o; o >>= 1) {
                const bool oh = __shfl_xor_sync(kFull, has, o);
                const double os = __shfl_xor_sync(kFull, s, o);
                const uint32_t oi = __shfl_xor_sync(kFull, id, o);
                const int ow = __shfl_xor_sync(kFull, who, o);
                const bool take = oh && (!has || os > s || (os == s && (oi < id || (oi == id && ow < who))));
                if (take) { has = oh; s = os; id = oi; who = ow; }
            }
            if (!has) break;
            if (lane == 0) { out_ids[(size_t)q * k + n_out] = id; out_scores[(size_t)q * k + n_out] = s; }
            if (lane == who) cur++;
        }
        // sparse: the rows are page-locked host memory (every store crosses PCIe): only the valid entries are written
        if (!sparse) for (uint32_t j = n_out + lane; j < k; j += 32) { out_ids[(size_t)q * k + j] = 0; out_scores[(size_t)q * k + j] = 0.0; }
        if (lane == 0) out_counts[q] = unsupported ? kCountUnsupported : n_out;
    }
}

// The same selection with one pointer per part instead of a stride: part p's packed block [scores | ids | counts] lives in
// the HBM of the GPU that searched shard p and is read from there over NVLink peer access, only the entries that are
// needed - the gather of the shard exchange is these loads, there is no copy of the blocks (sg_sharded_search_batch).
// A warp takes 32 consecutive queries at a time: the counts of a part for those queries are one coalesced 128-byte
// (remote) load, staged in shared memory; the 32 result counts leave as one 128-byte store (one PCIe write when the rows
// are page-locked host memory, instead of 32).  Within a query lane = part, as in sg_merge_topk_kernel.
__global__ void __launch_bounds__(256) sg_merge_topk_peer_kernel(uint32_t n_parts, uint32_t n_q, uint32_t k, const void *const *__restrict__ parts,
                                                                 uint32_t *out_ids, double *out_scores, uint32_t *out_counts, int sparse) {
    __shared__ uint32_t s_cnt[8][32][33];  // [warp][part][query of the chunk] (+1: conflict-free column reads)
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const double *my_scores = nullptr;
    const uint32_t *my_ids = nullptr;
    if ((uint32_t)lane < n_parts) {
        my_scores = (const double *)parts[lane];
        my_ids = (const uint32_t *)(my_scores + (size_t)n_q * k);
    }
    const uint32_t n_chunks = (n_q + 31) >> 5;
    for (uint32_t chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < n_chunks; chunk += warps) {
        const uint32_t q0 = chunk << 5;
        for (uint32_t p = 0; p < n_parts; p++) {
            const uint32_t *cnts = (const uint32_t *)((const double *)parts[p] + (size_t)n_q * k) + (size_t)n_q * k;
            s_cnt[warp][p][lane] = q0 + lane < n_q ? cnts[q0 + lane] : 0u;
        }
        __syncwarp();
        uint32_t my_out = 0;
        const uint32_t q_end = min(32u, n_q - q0);
        for (uint32_t i = 0; i < q_end; i++) {
            const size_t row = (size_t)(q0 + i) * k;
            uint32_t cur = 0, cnt = (uint32_t)lane < n_parts ? s_cnt[warp][lane][i] : 0u;
            bool unsupported = cnt == kCountUnsupported;
            if (unsupported) cnt = 0;
            if (cnt > k) cnt = k;
            unsupported = __any_sync(kFull, unsupported);
            uint32_t n_out = 0;
            if (__any_sync(kFull, cnt != 0u)) {
                // head of this lane's list, fetched once per advance (a remote load each)
                bool has = cur < cnt;
                double s = has ? my_scores[row] : 0.0;
                uint32_t id = has ? my_ids[row] : kInf;
                for (; n_out < k; n_out++) {
                    bool bh = has;
                    double bs = s;
                    uint32_t bi = id;
                    int who = lane;
                    for (int o = 16; o; o >>= 1) {
                        const bool oh = __shfl_xor_sync(kFull, bh, o);
                        const double os = __shfl_xor_sync(kFull, bs, o);
                        const uint32_t oi = __shfl_xor_sync(kFull, bi, o);
                        const int ow = __shfl_xor_sync(kFull, who, o);
                        const bool take = oh && (!bh || os > bs || (os == bs && (oi < bi || (oi == bi && ow < who))));
                        if (take) { bh = oh; bs = os; bi = oi; who = ow; }
                    }
                    if (!bh) break;
                    if (lane == 0) { out_ids[row + n_out] = bi; out_scores[row + n_out] = bs; }
                    if (lane == who) {
                        cur++;
                        has = cur < cnt;
                        s = has ? my_scores[row + cur] : 0.0;
                        id = has ? my_ids[row + cur] : kInf;
                    }
                }
            }
            if (!sparse) for (uint32_t j = n_out + lane; j < k; j += 32) { out_ids[row + j] = 0; out_scores[row + j] = 0.0; }
            if ((uint32_t)lane == i) my_out = unsupported ? kCountUnsupported : n_out;
        }
        if (q0 + lane < n_q) out_counts[q0 + lane] = my_out;
        __syncwarp();
    }
}

// ---------------- launchers (host) ----------------
cudaError_t launch_search(const DevIndex &ix, const SearchParams &p, int blocks, int warps_per_block, size_t smem_bytes,
                          cudaStream_t stream, cudaEvent_t *stage_events) {
    // per-device attribute shared by every index and host thread of the process: only ever raise it (a thread with a
    // smaller k must not lower it between another thread's set and its launch)
    cudaError_t e = cudaSuccess;
    {
        static std::mutex mu;
        static size_t opted_in[64] = {0};
        int device = 0;
        e = cudaGetDevice(&device);
        if (e != cudaSuccess) return e;
        if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
        std::lock_guard<std::mutex> lock(mu);
        if (smem_bytes > opted_in[device]) {
            e = cudaFuncSetAttribute(sg_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
            if (e != cudaSuccess) return e;
            opted_in[device] = smem_bytes;
        }
    }
    const int plan_blocks = (int)((p.n_q + kPlanThreads / 32 - 1) / (kPlanThreads / 32));
    if (stage_events) cudaEventRecord(stage_events[0], stream);
    sg_plan_kernel<<<plan_blocks < 148 * 8 ? plan_blocks : 148 * 8, kPlanThreads, 0, stream>>>(ix, p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (stage_events) cudaEventRecord(stage_events[1], stream);
    sg_search_kernel<<<blocks, warps_per_block * 32, smem_bytes, stream>>>(ix, p);
    if (stage_events) cudaEventRecord(stage_events[2], stream);
    return cudaGetLastError();
}

cudaError_t launch_merge_topk(uint32_t n_parts, uint32_t n_q, uint32_t k, const uint32_t *part_ids, const double *part_scores,
                              const uint32_t *part_counts, size_t stride_ids, size_t stride_scores, size_t stride_counts,
                              uint32_t *out_ids, double *out_scores, uint32_t *out_counts, int blocks, cudaStream_t stream, int sparse) {
    sg_merge_topk_kernel<<<blocks, 256, 0, stream>>>(n_parts, n_q, k, part_ids, part_scores, part_counts, stride_ids, stride_scores,
                                                     stride_counts, out_ids, out_scores, out_counts, sparse);
    return cudaGetLastError();
}

cudaError_t launch_merge_topk_peer(uint32_t n_parts, uint32_t n_q, uint32_t k, const void *const *parts, uint32_t *out_ids,
                                   double *out_scores, uint32_t *out_counts, int blocks, cudaStream_t stream, int sparse) {
    sg_merge_topk_peer_kernel<<<blocks, 256, 0, stream>>>(n_parts, n_q, k, parts, out_ids, out_scores, out_counts, sparse);
    return cudaGetLastError();
}

}  // namespace sg
