// sg_kernels.h — host-callable launchers of the sm_100a kernels in sg_kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include "sg_device.h"

namespace sg {

// stage_events (optional): n_kernels + 1 events recorded before, between and after the kernels of the launch
cudaError_t launch_search(const DevIndex &ix, const SearchParams &p, int blocks, int warps_per_block, size_t smem_bytes,
                          cudaStream_t stream, cudaEvent_t *stage_events = nullptr);
// bitmap engine (sg_bitmap.cu): sg_window_kernel + sg_tokens_kernel + sg_bitmap_search_kernel; p.warp_smem = bitmap_warp_smem(k)
size_t bitmap_warp_smem(uint32_t k);
// opt in to the shared memory the kernel needs at this k (raised, never lowered) and how many CTAs fit an SM (cached)
cudaError_t bitmap_search_occupancy(int device, uint32_t k, int *blocks_per_sm);
cudaError_t preload_bitmap_kernels();  // forces the lazy load of every kernel of sg_bitmap.cu (see there)
cudaError_t launch_window(const DevIndex &ix, const SearchParams &p, cudaStream_t stream);  // sg_window_kernel alone (fills p.wt)
cudaError_t launch_bitmap_search(const DevIndex &ix, const SearchParams &p, int sm_count, int blocks_per_sm, bool run_window,
                                 cudaStream_t stream, cudaEvent_t *stage_events = nullptr);
// count -> resolve pipeline (sg_count_kernel, sg_resolve_kernel, then sg_bitmap_search_kernel for the queries that ran out of
// scratch); p.lean_* set; stage_events (optional): 6 events around the five kernels
cudaError_t lean_occupancy(int device, uint32_t k, int *count_per_sm, int *resolve_per_sm);
// fused: sg_tokens_count_kernel tokenizes and counts in one launch (the caller has zeroed p.work_counter's kWorkWords);
// else sg_tokens_kernel + sg_count_kernel (needed for p.stats)
cudaError_t launch_lean_search(const DevIndex &ix, const SearchParams &p, int sm_count, int count_per_sm, int resolve_per_sm,
                               int search_per_sm, bool run_window, bool fused, cudaStream_t stream, cudaEvent_t *stage_events = nullptr);
// sg_long.cu: one warp per query of more than 128 n-grams (host-tokenized), ScanCount over HBM counters; k <= kSmemTopK
cudaError_t launch_long_queries(const DevIndex &ix, const LongParams &p, int blocks, cudaStream_t stream);
cudaError_t launch_merge_topk(uint32_t n_parts, uint32_t n_q, uint32_t k, const uint32_t *part_ids, const double *part_scores,
                              const uint32_t *part_counts, size_t stride_ids, size_t stride_scores, size_t stride_counts,
                              uint32_t *out_ids, double *out_scores, uint32_t *out_counts, int blocks, cudaStream_t stream,
                              int sparse = 0);  // sparse: write only the valid entries of a row (rows in page-locked host memory)
// parts: device array of n_parts pointers to packed blocks (sg_packed_rows_bytes), each in the HBM of its shard's GPU (peer access)
cudaError_t launch_merge_topk_peer(uint32_t n_parts, uint32_t n_q, uint32_t k, const void *const *parts, uint32_t *out_ids,
                                   double *out_scores, uint32_t *out_counts, int blocks, cudaStream_t stream, int sparse = 0);

// language model (sg_lm.cu)
cudaError_t launch_lm_context(const DevLm &lm, const uint32_t *ctx_ids, const uint32_t *ctx_off, uint32_t n_q, LmContext *out,
                              cudaStream_t stream);
cudaError_t launch_lm_score(const DevLm &lm, const uint32_t *ids, const uint32_t *off, uint32_t n, double *out, cudaStream_t stream);
cudaError_t launch_lm_score_next(const LmContext *ctx, const uint32_t *cand_ids, const uint32_t *cand_off, uint32_t n_q, double *out,
                                 cudaStream_t stream);
cudaError_t launch_predict_merge(const LmContext *ctx, uint32_t n_q, uint32_t k, const uint32_t *ac_ids, const uint32_t *ac_cnt,
                                 const uint32_t *fz_ids, const uint32_t *fz_cnt, uint32_t *out_ids, uint32_t *out_cnt, uint32_t *scratch,
                                 cudaStream_t stream);

}  // namespace sg
