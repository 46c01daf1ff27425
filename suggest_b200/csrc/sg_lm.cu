// sg_lm.cu — the n-gram language model of pkg/lm on the device and the batched spellchecker of pkg/spellchecker
// (BASELINE.json config #5).
//
//   pkg/lm/packed_array.go:163-210   find / findContainerPos      -> lm_find (two binary searches over sorted uint64)
//   pkg/lm/ngram_model.go:44-64      nGramModel.Score             -> sg_lm_score_kernel (stupid back-off, alpha = 0.4)
//   pkg/lm/ngram_model.go:67-99      nGramModel.Next              -> sg_lm_context_kernel
//   pkg/lm/scorer_next.go:15-23      scorerNext.ScoreNext         -> lm_next_count + calcScore (sg_lm_score_next_kernel)
//   pkg/spellchecker/spellchecker.go:40-92  SpellChecker.Predict  -> sg_predict_batch: completions of the last word ranked
//       by the model (sg_bitmap_search_kernel with an LmContext per query), topped up with fuzzy Cosine candidates
//       (the Suggest path), merged, stably re-sorted by the model's score and cut (sg_predict_merge_kernel).
// ScoreNext of a query's candidates is log(count(context, word) / count(context)) with one denominator per query, so every
// ranking here compares the integer counts: exact, whatever the last bit of a device log() is.
#include <cuda_runtime.h>

#include <cmath>

#include "sg_device.h"
#include "sg_kernels.h"

namespace sg {

namespace {

// packedArray.findContainerPos: index of the container of `context`, or -1
__device__ __forceinline__ int lm_container_pos(const uint64_t *__restrict__ c, uint32_t n, uint32_t context) {
    if (n == 0 || (uint32_t)(c[0] >> 32) > context || (uint32_t)(c[n - 1] >> 32) < context) return -1;
    const uint64_t target = (uint64_t)context << 32;
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (c[mid] < target) lo = mid + 1; else hi = mid;
    }
    if (lo >= n || (uint32_t)(c[lo] >> 32) != context) return -1;
    return (int)lo;
}

// count of `word` among values[from, to) (ordered by word), 0 if absent; *pos = its position
__device__ __forceinline__ uint32_t lm_range_count(const uint64_t *__restrict__ v, uint32_t from, uint32_t to, uint32_t word, uint32_t *pos) {
    *pos = kLmInvalidContext;
    if (from >= to) return 0;
    if ((uint32_t)(v[from] >> 32) > word || (uint32_t)(v[to - 1] >> 32) < word) return 0;
    const uint64_t target = (uint64_t)word << 32;
    uint32_t lo = from, hi = to;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (v[mid] < target) lo = mid + 1; else hi = mid;
    }
    if (lo >= to || (uint32_t)(v[lo] >> 32) != word) return 0;
    *pos = lo;
    return (uint32_t)v[lo];
}

// packedArray.GetCount(word, context) at one level
__device__ __forceinline__ uint32_t lm_get_count(const DevLm &lm, int level, uint32_t word, uint32_t context, uint32_t *pos) {
    *pos = kLmInvalidContext;
    const int i = lm_container_pos(lm.containers[level], lm.n_containers[level], context);
    if (i < 0) return 0;
    const uint32_t from = (uint32_t)lm.containers[level][i];
    const uint32_t to = (uint32_t)i == lm.n_containers[level] - 1 ? lm.n_values[level] : (uint32_t)lm.containers[level][i + 1];
    return lm_range_count(lm.values[level], from, to, word, pos);
}

// calcScore, pkg/lm/ngram_model.go:163-175; counts[0..n)
__device__ __forceinline__ double lm_calc_score(const uint32_t *counts, int n) {
    double factor = 1.0;
    for (int i = n - 1; i >= 1; i--) {
        if (counts[i] > 0) return log(__ddiv_rn(__dmul_rn(factor, (double)counts[i]), (double)counts[i - 1]));
        factor = __dmul_rn(factor, 0.4);
    }
    return kLmUnknownWordScore;
}

// nGramModel.Next for one context -> LmContext (valid = 0 where the reference returns a nil scorer)
__device__ __forceinline__ LmContext lm_next(const DevLm &lm, const uint32_t *ctx, uint32_t len) {
    LmContext out{nullptr, 0u, 0u, 0u, 0u};
    if (len == 0 || len >= lm.order) return out;
    uint32_t parent = kLmInvalidContext, count = 0;
    for (uint32_t level = 0; level < len; level++) {
        uint32_t pos;
        count = lm_get_count(lm, (int)level, ctx[level], parent, &pos);
        if (count == 0) return out;
        parent = pos;
    }
    const int i = lm_container_pos(lm.containers[len], lm.n_containers[len], parent);  // SubVector(parent)
    if (i < 0) return out;
    out.vals = lm.values[len];
    out.from = (uint32_t)lm.containers[len][i];
    out.to = (uint32_t)i == lm.n_containers[len] - 1 ? lm.n_values[len] : (uint32_t)lm.containers[len][i + 1];
    out.ctx_count = count;
    out.valid = 1u;
    return out;
}

}  // namespace

__global__ void sg_lm_context_kernel(const DevLm lm, const uint32_t *__restrict__ ctx_ids, const uint32_t *__restrict__ ctx_off,
                                     uint32_t n_q, LmContext *out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_q) return;
    out[q] = lm_next(lm, ctx_ids + ctx_off[q], ctx_off[q + 1] - ctx_off[q]);
}

// nGramModel.Score of n-gram q = ids[off[q] .. off[q+1])
__global__ void sg_lm_score_kernel(const DevLm lm, const uint32_t *__restrict__ ids, const uint32_t *__restrict__ off, uint32_t n,
                                   double *out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const uint32_t *g = ids + off[q];
    uint32_t len = off[q + 1] - off[q];
    if (len > lm.order) len = lm.order;
    uint32_t counts[kMaxLmOrder + 1];
    uint32_t parent = kLmInvalidContext;
    for (uint32_t i = 0; i < len; i++) {
        if (i == 0) counts[0] = lm.totals[0];
        uint32_t pos;
        counts[i + 1] = lm_get_count(lm, (int)i, g[i], parent, &pos);
        parent = pos;  // an unseen prefix has no continuation: InvalidContextOffset matches no container
    }
    out[q] = lm_calc_score(counts, (int)len + 1);
}

// ScoreNext of candidate c under the context of its query; has_scorer[q] = 0 where nGramModel.Next returns nil
__global__ void sg_lm_score_next_kernel(const LmContext *__restrict__ ctx, const uint32_t *__restrict__ cand_ids,
                                        const uint32_t *__restrict__ cand_off, uint32_t n_q, double *out) {
    const uint32_t q = blockIdx.y;
    if (q >= n_q) return;
    const LmContext c = ctx[q];
    for (uint32_t i = cand_off[q] + blockIdx.x * blockDim.x + threadIdx.x; i < cand_off[q + 1]; i += gridDim.x * blockDim.x) {
        double s = kLmUnknownWordScore;
        if (c.valid) {
            uint32_t pos;
            const uint32_t count = lm_range_count(c.vals, c.from, c.to, cand_ids[i], &pos);
            if (count) s = log(__ddiv_rn((double)count, (double)c.ctx_count));  // calcScore with the last count present
        }
        out[i] = s;
    }
}

// SpellChecker.Predict, the part behind the two searches (spellchecker.go:56-91): one thread per query.
//   candidates = completions (already the k best by the model, (count desc, id asc)); if fewer than k, append the fuzzy
//   candidates not among them (merge, :134-151); with a scorer, sort.SliceStable by ScoreNext desc (:126-131);
//   if k < len keep k + 1 (sic, :87-89).  Output rows have stride k + 1.
__global__ void sg_predict_merge_kernel(const LmContext *__restrict__ ctx, uint32_t n_q, uint32_t k, const uint32_t *__restrict__ ac_ids,
                                        const uint32_t *__restrict__ ac_cnt, const uint32_t *__restrict__ fz_ids,
                                        const uint32_t *__restrict__ fz_cnt, uint32_t *out_ids, uint32_t *out_cnt, uint32_t *scratch) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_q) return;
    uint32_t *ids = scratch + (size_t)q * 4 * k;   // [2k] ids, [2k] counts
    uint32_t *cnt = ids + 2 * k;
    uint32_t n = ac_cnt[q] == kCountUnsupported ? 0u : ac_cnt[q];
    for (uint32_t i = 0; i < n; i++) ids[i] = ac_ids[(size_t)q * k + i];
    if (n < k) {
        const uint32_t nf = fz_cnt[q] == kCountUnsupported ? 0u : fz_cnt[q];
        const uint32_t n_ac = n;
        for (uint32_t j = 0; j < nf; j++) {
            const uint32_t y = fz_ids[(size_t)q * k + j];
            bool unique = true;
            for (uint32_t i = 0; i < n_ac && unique; i++) unique = ids[i] != y;
            if (unique) ids[n++] = y;
        }
    }
    const LmContext c = ctx[q];
    if (c.valid) {
        for (uint32_t i = 0; i < n; i++) {
            uint32_t pos;
            cnt[i] = lm_range_count(c.vals, c.from, c.to, ids[i], &pos);
        }
        for (uint32_t i = 1; i < n; i++) {  // stable insertion sort, count desc
            const uint32_t id = ids[i], ci = cnt[i];
            uint32_t j = i;
            while (j > 0 && cnt[j - 1] < ci) { ids[j] = ids[j - 1]; cnt[j] = cnt[j - 1]; j--; }
            ids[j] = id;
            cnt[j] = ci;
        }
    }
    if (k < n) n = k + 1;
    for (uint32_t i = 0; i < k + 1; i++) out_ids[(size_t)q * (k + 1) + i] = i < n ? ids[i] : 0u;
    out_cnt[q] = n;
}

cudaError_t launch_lm_context(const DevLm &lm, const uint32_t *ctx_ids, const uint32_t *ctx_off, uint32_t n_q, LmContext *out,
                              cudaStream_t stream) {
    if (n_q) sg_lm_context_kernel<<<(n_q + 127) / 128, 128, 0, stream>>>(lm, ctx_ids, ctx_off, n_q, out);
    return cudaGetLastError();
}

cudaError_t launch_lm_score(const DevLm &lm, const uint32_t *ids, const uint32_t *off, uint32_t n, double *out, cudaStream_t stream) {
    if (n) sg_lm_score_kernel<<<(n + 127) / 128, 128, 0, stream>>>(lm, ids, off, n, out);
    return cudaGetLastError();
}

cudaError_t launch_lm_score_next(const LmContext *ctx, const uint32_t *cand_ids, const uint32_t *cand_off, uint32_t n_q, double *out,
                                 cudaStream_t stream) {
    for (uint32_t q0 = 0; q0 < n_q; q0 += 65535) {
        const uint32_t nq = n_q - q0 < 65535 ? n_q - q0 : 65535;
        sg_lm_score_next_kernel<<<dim3(4, nq), 128, 0, stream>>>(ctx + q0, cand_ids, cand_off + q0, nq, out);
    }
    return cudaGetLastError();
}

cudaError_t launch_predict_merge(const LmContext *ctx, uint32_t n_q, uint32_t k, const uint32_t *ac_ids, const uint32_t *ac_cnt,
                                 const uint32_t *fz_ids, const uint32_t *fz_cnt, uint32_t *out_ids, uint32_t *out_cnt, uint32_t *scratch,
                                 cudaStream_t stream) {
    if (n_q) sg_predict_merge_kernel<<<(n_q + 127) / 128, 128, 0, stream>>>(ctx, n_q, k, ac_ids, ac_cnt, fz_ids, fz_cnt, out_ids, out_cnt, scratch);
    return cudaGetLastError();
}

}  // namespace sg
