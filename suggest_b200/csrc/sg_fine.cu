// sg_fine.cu — the exact level under the bucket bitmaps (DevIndex::rank4 / fine), built on the device from an index that
// is already in HBM (either build path: sg_gpubuild.cu or the host build's upload).
//
// The reference resolves nothing: its mergers walk the decoded posting lists themselves (pkg/merger/cp_merge.go:19-120).
// Here a bucket that reaches its threshold in the bit-sliced count has to be counted per document; round 1 did that with
// a 4-ary search of every (term, segment) posting list for the bucket's slot range - three rounds of dependent loads per
// list.  With this level the bits of a (term, bucket) pair are addressed directly: the pair's number is the rank of its
// bit in the term's row (a prefix count per group of four words + popcounts inside the group), and fine[] holds the
// pair's 2^bshift document bits at that number.
#include <cub/cub.cuh>

#include <string>
#include <vector>

#include "sg_device.h"
#include "sg_host.h"

namespace sg {

namespace {

__global__ void sg_group_popc_kernel(const uint4 *__restrict__ bitmaps, uint64_t n_groups, uint32_t *out) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint4 v = bitmaps[g];
    out[g] = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
}

// one thread per posting: its term from the term starts (list_off[t * (S + 1)], ascending), its pair from the rank
__global__ void sg_fine_fill_kernel(const DevIndex ix, uint64_t n_postings, uint32_t *fine) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_postings) return;
    const size_t stride = (size_t)ix.n_segments + 1;
    uint32_t lo = 0, hi = ix.n_terms;  // last term whose first posting is <= i
    while (hi - lo > 1) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if ((uint64_t)ix.list_off[(size_t)mid * stride] <= i) lo = mid; else hi = mid;
    }
    const uint32_t t = lo;
    const uint32_t slot = ix.postings[i];
    const uint32_t bucket = slot >> ix.bshift, w = bucket >> 5, bit = bucket & 31u;
    const size_t at = (size_t)t * ix.row_words + w;
    const uint4 grp = *(const uint4 *)(ix.bitmaps + (at & ~(size_t)3));
    const uint32_t ws[4] = {grp.x, grp.y, grp.z, grp.w};
    uint32_t pair = ix.rank4[at >> 2];
    const uint32_t in = (uint32_t)(at & 3);
    for (uint32_t j = 0; j < in; j++) pair += __popc(ws[j]);
    pair += __popc(ws[in] & ((1u << bit) - 1u));
    const uint64_t bitpos = ((uint64_t)pair << ix.bshift) + (slot & ((1u << ix.bshift) - 1u));
    atomicOr(fine + (bitpos >> 5), 1u << (bitpos & 31u));
}

}  // namespace

// Adds rank4 / fine to a DevIndex whose bitmaps, postings and list offsets are in HBM.  Returns "" (built, or not needed:
// bshift = 0, no bitmaps), "skip: ..." (over the budget: the index is searched without the level), or a CUDA error text.
std::string build_fine_level(DevIndex *ix, uint64_t n_postings, uint64_t budget_bytes, std::vector<void *> *allocs,
                             uint64_t *device_bytes) {
    ix->rank4 = nullptr;
    ix->fine = nullptr;
    if (ix->row_words == 0 || ix->bshift == 0 || ix->bitmaps == nullptr || n_postings == 0) return "";
    const uint64_t n_groups = ((uint64_t)ix->n_terms + 1) * ix->row_words / 4;
    // pairs <= postings; fine bytes are known only after the scan, bound them first
    const uint64_t fine_bound = ((n_postings << ix->bshift) + 7) / 8 + 64;
    if (n_groups * 4 + fine_bound > budget_bytes) return "skip: the exact level would not fit its memory budget";
    if (n_groups > 0x7FFFFFF0ull || n_postings > 0x7FFFFFF0ull) return "skip: too many bitmap groups for 32-bit ranks";
#define FN_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess) { cudaGetLastError(); return std::string("cuda: ") + cudaGetErrorString(e__); } \
    } while (0)
    uint32_t *d_rank = nullptr, *d_fine = nullptr;
    FN_CUDA(cudaMalloc((void **)&d_rank, (n_groups + 1) * 4));
    allocs->push_back(d_rank);
    const int threads = 256;
    sg_group_popc_kernel<<<(unsigned)((n_groups + threads - 1) / threads), threads>>>((const uint4 *)ix->bitmaps, n_groups, d_rank);
    FN_CUDA(cudaGetLastError());
    uint32_t last_count = 0;
    FN_CUDA(cudaMemcpy(&last_count, d_rank + n_groups - 1, 4, cudaMemcpyDeviceToHost));
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    FN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_rank, d_rank, (int)n_groups));
    FN_CUDA(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 4));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_rank, d_rank, (int)n_groups);
    uint32_t last_rank = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&last_rank, d_rank + n_groups - 1, 4, cudaMemcpyDeviceToHost);
    cudaFree(d_tmp);
    FN_CUDA(e);
    const uint64_t n_pairs = (uint64_t)last_rank + last_count;
    const uint64_t fine_words = ((n_pairs << ix->bshift) + 31) / 32 + 8;  // + slack: a pair's words are read as whole words
    FN_CUDA(cudaMalloc((void **)&d_fine, fine_words * 4));
    allocs->push_back(d_fine);
    FN_CUDA(cudaMemset(d_fine, 0, fine_words * 4));
    ix->rank4 = d_rank;
    sg_fine_fill_kernel<<<(unsigned)((n_postings + threads - 1) / threads), threads>>>(*ix, n_postings, d_fine);
    FN_CUDA(cudaGetLastError());
    FN_CUDA(cudaDeviceSynchronize());
    ix->fine = d_fine;
    if (device_bytes) *device_bytes += (n_groups + 1) * 4 + fine_words * 4;
#undef FN_CUDA
    return "";
}

}  // namespace sg
