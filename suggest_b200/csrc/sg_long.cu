// sg_long.cu — queries of more than 128 n-grams (kMaxQueryTokens), which the batched kernels refuse.
//
// The reference answers them like any other query (its mergers count overlaps up to 0xFFFF, pkg/merger/list_merger.go:9,
// 51-57); a drop-in must not fail where the reference answers.  They are rare and only matter when the dictionary holds
// documents that long, so this path is simple rather than fast: the host tokenizes (sg_text.cpp, the same chain as the
// index build), and one warp per query runs ScanCount as the reference defines it (pkg/merger/scan_count.go:14-88) - one
// 32-bit counter per document slot of the window in HBM, one atomic add per posting of every (token, window) run, then a
// scan of the counters against Threshold(alpha, sizeA, sizeB) of the slot's segment, score, sorted top-k.
#include <cuda_runtime.h>

#include "sg_common.cuh"
#include "sg_kernels.h"

namespace sg {

__global__ void __launch_bounds__(32) sg_long_query_kernel(const DevIndex ix, const LongParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x;
    QueryCtx c{};
    c.k = p.k;
    c.tk_score = (double *)smem;
    c.tk_id = (uint32_t *)(c.tk_score + p.k);
    const int S = (int)ix.n_segments;
    const size_t stride = (size_t)S + 1;
    uint32_t *counters = p.counters + (size_t)blockIdx.x * ix.n_ids;
    const int metric = p.mode == 1 ? (int)kAutocomplete : p.metric;
    for (uint32_t qi = blockIdx.x; qi < p.n_long; qi += gridDim.x) {
        const uint32_t k0 = p.key_off[qi], k1 = p.key_off[qi + 1];
        const int size_a = (int)(k1 - k0);  // len(tokens), duplicates after normalisation included (suggester.go:53)
        int present = 0;
        for (uint32_t j = k0 + lane; j < k1; j += 32) {
            const uint32_t t = term_lookup(ix, p.keys[j]);
            p.terms[j] = t;
            present += t != kNoTerm;
        }
        present = __reduce_add_sync(kFull, present);
        __syncwarp();
        c.tk_len = 0;
        int b_min = 0, b_max = -1;
        if (size_a > 0 && !(p.mode == 1 && present < size_a)) {  // Autocomplete: every token must be a term (autocomplete.go:40-77)
            b_min = max(metric_min_y(metric, p.alpha, size_a), 0);
            b_max = min(metric_max_y(metric, p.alpha, size_a), S - 1);
        }
        if (b_max >= b_min && present > 0) {
            const uint32_t lo = ix.seg_start[b_min], hi = ix.seg_start[b_max + 1];
            for (uint32_t s = lo + lane; s < hi; s += 32) counters[s] = 0u;
            __syncwarp();
            for (uint32_t j = k0; j < k1; j++) {
                const uint32_t t = p.terms[j];
                if (t == kNoTerm) continue;
                const uint32_t *o = ix.list_off + (size_t)t * stride;
                const uint32_t a = __ldg(o + b_min), e = __ldg(o + b_max + 1);
                for (uint32_t pos = a + lane; pos < e; pos += 32) atomicAdd(counters + __ldg(ix.postings + pos), 1u);
            }
            __threadfence_block();
            __syncwarp();
            for (int B = b_min; B <= b_max; B++) {
                const int T = metric_threshold(metric, p.alpha, size_a, B);
                if (!threshold_admits(T, size_a, B)) continue;  // suggester.go:76
                const uint32_t s0 = ix.seg_start[B], s1 = ix.seg_start[B + 1];
                for (uint32_t base = s0; base < s1; base += 32) {
                    const uint32_t slot = base + lane;
                    const uint32_t cnt = slot < s1 ? counters[slot] : 0u;
                    unsigned m = __ballot_sync(kFull, cnt >= (uint32_t)T);
                    while (m) {
                        const int i = __ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t id = __ldg(ix.perm + base + i);
                        const int overlap = (int)min(__shfl_sync(kFull, cnt, i), 0xFFFFu);  // MaxOverlap, list_merger.go:9
                        const double score = p.mode == 1 ? -(double)(ix.id_base + id) : metric_score(metric, overlap, size_a, B);
                        topk_insert(c, score, id, lane);
                    }
                }
            }
        }
        __syncwarp();
        const size_t row = (size_t)qi * p.k;
        for (uint32_t j = lane; j < p.k; j += 32) {
            const bool has = (int)j < c.tk_len;
            p.out_ids[row + j] = has ? ix.id_base + c.tk_id[j] : 0u;
            p.out_scores[row + j] = has ? c.tk_score[j] : 0.0;
        }
        if (lane == 0) p.out_counts[qi] = (uint32_t)c.tk_len;
        __syncwarp();
    }
}

cudaError_t launch_long_queries(const DevIndex &ix, const LongParams &p, int blocks, cudaStream_t stream) {
    const size_t smem = (size_t)p.k * 12;
    sg_long_query_kernel<<<blocks, 32, smem, stream>>>(ix, p);
    return cudaGetLastError();
}

}  // namespace sg
