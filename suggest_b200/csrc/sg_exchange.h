// sg_exchange.h — parameters of sg_exchange_merge_kernel (sg_exchange.cu): the fused shard exchange + merge.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

int sg_internal_fail(int code, const std::string &msg);  // sg_api.cu: sets sg_last_error() of the calling thread

namespace sg {

constexpr uint32_t kMaxExchangeRanks = 16;  // GPUs of one box; the region starts with 2 x 16 64-bit flags

struct ExchangeParams {
    uint32_t rank, world, n_q, k;
    unsigned long long step;                        // this call's number (1, 2, ...): what the flags count up to
    unsigned long long *regions[kMaxExchangeRanks];  // every rank's region: [start flags x 16 | end flags x 16 | ...]
    const void *shard_rows[kMaxExchangeRanks];       // packed rows of shard s (sg_packed_rows_bytes layout), in rank s's HBM
    void *merged_rows[kMaxExchangeRanks];            // merged rows of the whole batch, one copy per rank
    unsigned int *done_counter;                      // local: blocks of this launch that have finished
    unsigned int *status;                            // local: 0 ok, 1 start barrier timed out, 2 end barrier timed out
};

cudaError_t launch_exchange_merge(const ExchangeParams &p, int sm_count, cudaStream_t stream);

}  // namespace sg
