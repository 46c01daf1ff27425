// sg_gpubuild.cu — index build on the device.
//
// What suggest.Index + index.Writer.AddDocument/Commit + index.Reader.Read do on the CPU
//   (pkg/suggest/indexer.go:14-45, pkg/index/indexer_writer.go:66-145, pkg/index/index_reader.go:29-120)
// for a dictionary given as text: tokenise every document (the same device tokenizer as the query path, emitting packed
// n-gram keys), file it under its cardinality segment, and lay the posting lists out as sg_device.h describes —
// segment-aligned slots, CSR posting lists per (term, segment), one bucket bitmap per term, the term hash table.
// Sorting does the grouping (CUB radix sorts) instead of the reference's map-of-maps:
//   1. sg_doc_tokens_kernel (count), exclusive scan, sg_doc_tokens_kernel (emit): (key, document) pairs, distinct per document
//   2. sort the keys, run-length encode: the terms (id = rank of the key) and their posting counts
//   3. cardinality histogram -> bucket width (host, from the term frequencies) -> aligned segment starts
//   4. stable sort of the documents by cardinality -> slots (perm and its inverse)
//   5. (term id << 32 | slot) per pair, sorted: the posting lists, in order
//   6. list offsets by binary search, bitmaps by atomicOr, hash table by atomicCAS
// The result is the same index the host build (sg_index.cpp) produces, up to the numbering of the terms, which is
// internal.  Documents with more than 128 n-grams, or more than 4 GB of text, make the caller fall back to the host build.
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "sg_common.cuh"
#include "sg_host.h"
#include "sg_kernels.h"

namespace sg {

namespace {

constexpr int kDocThreads = 256;

// Tokenise documents [0, n_docs); one warp per document.
//   emit = false: card[d] = len(tokens) (duplicates after normalisation included, indexer_writer.go:67),
//                 nkeys[d] = distinct keys (a posting list holds the document once), *too_long |= more than 128 n-grams
//   emit = true : the distinct keys of d go to pair_key[key_off[d] ..], pair_doc[..] = d
__global__ void __launch_bounds__(kDocThreads) sg_doc_tokens_kernel(const DevIndex ix, const SearchParams p, bool emit, uint32_t *card,
                                                                    uint32_t *nkeys, const uint32_t *__restrict__ key_off,
                                                                    uint64_t *pair_key, uint32_t *pair_doc, uint32_t *too_long) {
    __shared__ __align__(16) uint32_t s_scratch[kDocThreads / 32][kMaxRunes + 2 * kMaxQueryTokens];
    __shared__ __align__(16) uint64_t s_keys_all[kDocThreads / 32][kMaxQueryTokens];
    __shared__ uint8_t s_ascii[128];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < 128) s_ascii[threadIdx.x] = ix.ascii_code[threadIdx.x];
    __syncthreads();
    uint32_t *s_runes = s_scratch[warp];
    uint32_t *s_lterm = s_runes + kMaxRunes;
    uint32_t *s_hash = s_lterm + kMaxQueryTokens;
    uint64_t *s_keys = s_keys_all[warp];
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; d < p.n_q; d += n_warps) {
        int size_a = 0, n_tok = 0;
        const bool unsupported = tokenize_query<true>(ix, p, d, s_runes, s_lterm, s_hash, lane, &size_a, &n_tok, s_ascii, s_keys);
        __syncwarp();
        if (unsupported) {
            if (lane == 0) atomicOr(too_long, 1u);
            size_a = n_tok = 0;
        }
        // distinct keys, first occurrence kept
        uint32_t n_distinct = 0;
        for (int base = 0; base < n_tok; base += 32) {
            const int i = base + lane;
            bool keep = i < n_tok;
            const uint64_t key = keep ? s_keys[i] : 0ull;
            for (int j = 0; j < i && keep; j++) keep = s_keys[j] != key;
            const unsigned km = __ballot_sync(kFull, keep);
            if (emit && keep) {
                const uint32_t at = __ldg(key_off + d) + n_distinct + (uint32_t)__popc(km & ((1u << lane) - 1u));
                pair_key[at] = key;
                pair_doc[at] = d;
            }
            n_distinct += (uint32_t)__popc(km);
        }
        if (!emit && lane == 0) { card[d] = (uint32_t)size_a; nkeys[d] = n_distinct; }
        __syncwarp();
    }
}

// 64-bit sum of per-document key counts: the 32-bit exclusive scan below wraps silently past 2^32 pairs
__global__ void sg_sum64_kernel(const uint32_t *__restrict__ a, uint32_t n, unsigned long long *out) {
    unsigned long long s = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += a[i];
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

__global__ void sg_iota_kernel(uint32_t *a, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}

__global__ void sg_hist_kernel(const uint32_t *__restrict__ card, uint32_t n, uint32_t *hist) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(hist + card[i], 1u);
}

// term hash table: same probing as term_lookup (sg_common.cuh) and HostIndex::build_hash
__global__ void sg_hash_insert_kernel(const uint64_t *__restrict__ keys, uint32_t n_terms, uint4 *table, uint32_t mask) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_terms) return;
    const uint64_t key = keys[t];
    uint32_t h = (uint32_t)mix64(key) & mask;
    for (;;) {  // the key occupies the first 8 bytes of the 16-byte entry
        const unsigned long long prev = atomicCAS((unsigned long long *)(table + h), 0ull, (unsigned long long)key);
        if (prev == 0ull) { table[h].z = t; return; }
        h = (h + 1) & mask;
    }
}

// slot of every document: sorted position j (documents ordered by cardinality, then id) -> seg_start[c] + rank inside the segment
__global__ void sg_slots_kernel(const uint32_t *__restrict__ sorted_card, const uint32_t *__restrict__ sorted_doc, uint32_t n_docs,
                                const uint32_t *__restrict__ seg_start, const uint32_t *__restrict__ dense_start, uint32_t *slot_of_doc,
                                uint32_t *perm) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_docs) return;
    const uint32_t c = sorted_card[j], d = sorted_doc[j];
    const uint32_t slot = seg_start[c] + (j - dense_start[c]);
    slot_of_doc[d] = slot;
    perm[slot] = d;
}

__global__ void sg_posting_keys_kernel(const DevIndex ix, const uint64_t *__restrict__ pair_key, const uint32_t *__restrict__ pair_doc,
                                       const uint32_t *__restrict__ slot_of_doc, uint64_t n_pairs, uint64_t *pk) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const uint32_t t = term_lookup(ix, pair_key[i]);
    pk[i] = (uint64_t)t << 32 | slot_of_doc[pair_doc[i]];
}

__global__ void sg_postings_bitmaps_kernel(const uint64_t *__restrict__ pk, uint64_t n_pairs, uint32_t *postings, uint32_t *bitmaps,
                                           uint32_t row_words, uint32_t bshift) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const uint32_t t = (uint32_t)(pk[i] >> 32), slot = (uint32_t)pk[i];
    postings[i] = slot;
    if (row_words) {
        const uint32_t bucket = slot >> bshift;
        atomicOr(bitmaps + (size_t)t * row_words + (bucket >> 5), 1u << (bucket & 31));
    }
}

// list_off[t][B] = first posting of term t whose slot is >= seg_start[B]; counts the non-empty lists
__global__ void sg_list_offsets_kernel(const uint64_t *__restrict__ pk, uint64_t n_pairs, const uint32_t *__restrict__ seg_start,
                                       uint32_t n_terms, uint32_t S, uint32_t *list_off, unsigned long long *n_lists) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)n_terms * (S + 1)) return;
    const uint32_t t = (uint32_t)(i / (S + 1)), B = (uint32_t)(i % (S + 1));
    auto bound = [&](uint32_t b) {
        const uint64_t target = (uint64_t)t << 32 | seg_start[b];
        uint64_t lo = 0, hi = n_pairs;
        while (lo < hi) {
            const uint64_t mid = lo + ((hi - lo) >> 1);
            if (pk[mid] < target) lo = mid + 1; else hi = mid;
        }
        return (uint32_t)lo;
    };
    const uint32_t a = bound(B);
    list_off[i] = a;
    if (B < S && bound(B + 1) > a) atomicAdd(n_lists, 1ull);
}

struct Temp {  // device temporaries of the build
    std::vector<void *> ptrs;
    template <typename T>
    cudaError_t get(T **p, size_t n) {
        cudaError_t e = cudaMalloc((void **)p, (n ? n : 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
    ~Temp() { for (void *p : ptrs) cudaFree(p); }
};

#define GB_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t e__ = (expr);                                                       \
        if (e__ != cudaSuccess) return std::string("cuda: ") + #expr + ": " + cudaGetErrorString(e__); \
    } while (0)

}  // namespace

// Returns "" on success, "fallback: ..." when the dictionary is not eligible (the caller then builds on the host), any
// other text on a CUDA failure.  On success `out` owns the device arrays (pushed to *allocs as well).
std::string gpu_build(const DevIndex &text, const char *doc_bytes, const uint64_t *doc_off, uint32_t n_docs, int want_bshift,
                      uint64_t bitmap_budget, GpuBuilt *out, std::vector<void *> *allocs) {
    const uint64_t total = doc_off[n_docs];
    if (total > 0xFFFFFF00ull) return "fallback: more than 4 GB of dictionary text";
    if (n_docs == 0) return "fallback: empty dictionary";
    // documents with non-ASCII bytes arrive lower-cased (strings.ToLower on the host, as for queries)
    std::string low;
    std::vector<uint32_t> off32((size_t)n_docs + 1);
    unsigned char high = 0;
    for (uint64_t i = 0; i < total; i++) high |= (unsigned char)doc_bytes[i];
    const char *src = doc_bytes;
    if (high & 0x80) {
        low.reserve(total + total / 8 + 16);
        for (uint32_t d = 0; d < n_docs; d++) {
            off32[d] = (uint32_t)low.size();
            to_lower((const uint8_t *)doc_bytes + doc_off[d], (size_t)(doc_off[d + 1] - doc_off[d]), &low);
            if (low.size() > 0xFFFFFF00ull) return "fallback: more than 4 GB of dictionary text";
        }
        off32[n_docs] = (uint32_t)low.size();
        src = low.data();
    } else {
        for (uint32_t d = 0; d <= n_docs; d++) off32[d] = (uint32_t)doc_off[d];
    }
    const size_t n_bytes = off32[n_docs];

    Temp tmp;
    char *d_bytes;
    uint32_t *d_off, *d_card, *d_nkeys, *d_key_off, *d_flag;
    GB_CUDA(tmp.get(&d_bytes, n_bytes + 16));
    GB_CUDA(tmp.get(&d_off, (size_t)n_docs + 1));
    GB_CUDA(tmp.get(&d_card, n_docs));
    GB_CUDA(tmp.get(&d_nkeys, (size_t)n_docs + 1));
    GB_CUDA(tmp.get(&d_key_off, (size_t)n_docs + 1));
    GB_CUDA(tmp.get(&d_flag, 4));
    if (n_bytes) GB_CUDA(cudaMemcpy(d_bytes, src, n_bytes, cudaMemcpyHostToDevice));
    GB_CUDA(cudaMemcpy(d_off, off32.data(), ((size_t)n_docs + 1) * 4, cudaMemcpyHostToDevice));
    GB_CUDA(cudaMemset(d_flag, 0, 16));
    GB_CUDA(cudaMemset(d_nkeys + n_docs, 0, 4));

    SearchParams p{};
    p.q_bytes = d_bytes;
    p.q_off = d_off;
    p.n_q = n_docs;
    p.mode = 0;
    const int blocks = (int)std::min<uint64_t>(((uint64_t)n_docs + kDocThreads / 32 - 1) / (kDocThreads / 32), 148 * 16);
    auto grid1 = [](uint64_t n) { return (unsigned)((n + 255) / 256); };

    // 1. count, scan, emit
    sg_doc_tokens_kernel<<<blocks, kDocThreads>>>(text, p, false, d_card, d_nkeys, nullptr, nullptr, nullptr, d_flag);
    GB_CUDA(cudaGetLastError());
    uint32_t flag = 0;
    GB_CUDA(cudaMemcpy(&flag, d_flag, 4, cudaMemcpyDeviceToHost));
    if (flag) return "fallback: a document has more than 128 n-grams";
    void *d_cub = nullptr;
    size_t cub_bytes = 0;
    auto cub_reserve = [&](size_t need) -> cudaError_t {
        if (need <= cub_bytes) return cudaSuccess;
        void *pnew = nullptr;
        cudaError_t e = cudaMalloc(&pnew, need);
        if (e != cudaSuccess) return e;
        tmp.ptrs.push_back(pnew);  // the old block stays until the end: simpler than tracking it
        d_cub = pnew;
        cub_bytes = need;
        return cudaSuccess;
    };
    {   // the pair count in 64 bits first: offsets, CUB's int counts and every buffer below assume it fits 2^31
        unsigned long long *d_total = nullptr, total = 0;
        GB_CUDA(tmp.get(&d_total, 1));
        GB_CUDA(cudaMemset(d_total, 0, sizeof(unsigned long long)));
        sg_sum64_kernel<<<296, 256>>>(d_nkeys, n_docs, d_total);
        GB_CUDA(cudaGetLastError());
        GB_CUDA(cudaMemcpy(&total, d_total, sizeof(total), cudaMemcpyDeviceToHost));
        if (total > 0x7FFFFFF0ull) return "fallback: more than 2^31 postings";
    }
    size_t need = 0;
    GB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, d_nkeys, d_key_off, (int)n_docs + 1));
    GB_CUDA(cub_reserve(need));
    GB_CUDA(cub::DeviceScan::ExclusiveSum(d_cub, need, d_nkeys, d_key_off, (int)n_docs + 1));
    uint32_t n_pairs32 = 0;
    GB_CUDA(cudaMemcpy(&n_pairs32, d_key_off + n_docs, 4, cudaMemcpyDeviceToHost));
    const uint64_t n_pairs = n_pairs32;
    if (n_pairs == 0) return "fallback: no document has an n-gram";
    if (n_pairs > 0x7FFFFFF0ull) return "fallback: more than 2^31 postings";
    uint64_t *d_pair_key, *d_keys_sorted, *d_terms;
    uint32_t *d_pair_doc, *d_freq, *d_n_terms;
    GB_CUDA(tmp.get(&d_pair_key, n_pairs));
    GB_CUDA(tmp.get(&d_pair_doc, n_pairs));
    sg_doc_tokens_kernel<<<blocks, kDocThreads>>>(text, p, true, nullptr, nullptr, d_key_off, d_pair_key, d_pair_doc, d_flag);
    GB_CUDA(cudaGetLastError());

    // 2. terms = distinct keys in key order, with their posting counts
    GB_CUDA(tmp.get(&d_keys_sorted, n_pairs));
    GB_CUDA(tmp.get(&d_terms, n_pairs));
    GB_CUDA(tmp.get(&d_freq, n_pairs));
    GB_CUDA(tmp.get(&d_n_terms, 1));
    GB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, need, d_pair_key, d_keys_sorted, (int)n_pairs));
    GB_CUDA(cub_reserve(need));
    GB_CUDA(cub::DeviceRadixSort::SortKeys(d_cub, need, d_pair_key, d_keys_sorted, (int)n_pairs));
    GB_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, need, d_keys_sorted, d_terms, d_freq, d_n_terms, (int)n_pairs));
    GB_CUDA(cub_reserve(need));
    GB_CUDA(cub::DeviceRunLengthEncode::Encode(d_cub, need, d_keys_sorted, d_terms, d_freq, d_n_terms, (int)n_pairs));
    uint32_t n_terms = 0;
    GB_CUDA(cudaMemcpy(&n_terms, d_n_terms, 4, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> freq(n_terms);
    GB_CUDA(cudaMemcpy(freq.data(), d_freq, (size_t)n_terms * 4, cudaMemcpyDeviceToHost));

    // 3. segments: cardinality histogram, bucket width, aligned starts (the same rules as sg_index.cpp)
    uint32_t *d_max;
    GB_CUDA(tmp.get(&d_max, 1));
    GB_CUDA(cub::DeviceReduce::Max(nullptr, need, d_card, d_max, (int)n_docs));
    GB_CUDA(cub_reserve(need));
    GB_CUDA(cub::DeviceReduce::Max(d_cub, need, d_card, d_max, (int)n_docs));
    uint32_t max_card = 0;
    GB_CUDA(cudaMemcpy(&max_card, d_max, 4, cudaMemcpyDeviceToHost));
    const uint32_t S = max_card + 1;
    uint32_t *d_hist;
    GB_CUDA(tmp.get(&d_hist, S));
    GB_CUDA(cudaMemset(d_hist, 0, (size_t)S * 4));
    sg_hist_kernel<<<grid1(n_docs), 256>>>(d_card, n_docs, d_hist);
    GB_CUDA(cudaGetLastError());
    std::vector<uint32_t> seg_count(S);
    GB_CUDA(cudaMemcpy(seg_count.data(), d_hist, (size_t)S * 4, cudaMemcpyDeviceToHost));
    if ((uint64_t)n_terms * (S + 1) > 0xFFFFFFF0ull) return "fallback: term x segment offset table exceeds 32 bits";
    std::vector<uint32_t> seg_start;
    uint32_t bs = 0, row_words = 0, n_ids = 0;
    std::string err = choose_layout(seg_count, freq, n_docs, n_pairs, want_bshift, bitmap_budget, &bs, &row_words, &seg_start, &n_ids);
    if (!err.empty()) return "fallback: " + err;
    std::vector<uint32_t> dense_start(S + 1, 0);
    for (uint32_t b = 0; b < S; b++) dense_start[b + 1] = dense_start[b] + seg_count[b];

    // persistent arrays
    auto keep = [&](void **p, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(p, bytes ? bytes : 4);
        if (e == cudaSuccess) allocs->push_back(*p);
        return e;
    };
    uint32_t *d_seg_start, *d_dense_start, *d_perm, *d_slot_of_doc;
    GB_CUDA(keep((void **)&d_seg_start, ((size_t)S + 1) * 4));
    GB_CUDA(tmp.get(&d_dense_start, (size_t)S + 1));
    GB_CUDA(keep((void **)&d_perm, (size_t)n_ids * 4));
    GB_CUDA(tmp.get(&d_slot_of_doc, n_docs));
    GB_CUDA(cudaMemcpy(d_seg_start, seg_start.data(), ((size_t)S + 1) * 4, cudaMemcpyHostToDevice));
    GB_CUDA(cudaMemcpy(d_dense_start, dense_start.data(), ((size_t)S + 1) * 4, cudaMemcpyHostToDevice));
    GB_CUDA(cudaMemset(d_perm, 0xFF, (size_t)n_ids * 4));

    // 4. slots: stable sort of the documents by cardinality keeps the id order inside a segment
    uint32_t *d_iota, *d_sorted_card, *d_sorted_doc;
    GB_CUDA(tmp.get(&d_iota, n_docs));
    GB_CUDA(tmp.get(&d_sorted_card, n_docs));
    GB_CUDA(tmp.get(&d_sorted_doc, n_docs));
    sg_iota_kernel<<<grid1(n_docs), 256>>>(d_iota, n_docs);
    int card_bits = 1;
    while ((1u << card_bits) <= max_card && card_bits < 32) card_bits++;
    GB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, d_card, d_sorted_card, d_iota, d_sorted_doc, (int)n_docs, 0, card_bits));
    GB_CUDA(cub_reserve(need));
    GB_CUDA(cub::DeviceRadixSort::SortPairs(d_cub, need, d_card, d_sorted_card, d_iota, d_sorted_doc, (int)n_docs, 0, card_bits));
    sg_slots_kernel<<<grid1(n_docs), 256>>>(d_sorted_card, d_sorted_doc, n_docs, d_seg_start, d_dense_start, d_slot_of_doc, d_perm);
    GB_CUDA(cudaGetLastError());

    // 6a. term hash table (needed by step 5)
    size_t cap = 16;
    while (cap < (size_t)n_terms * 2 + 2) cap <<= 1;
    uint4 *d_table;
    GB_CUDA(keep((void **)&d_table, cap * sizeof(uint4)));
    GB_CUDA(cudaMemset(d_table, 0, cap * sizeof(uint4)));
    sg_hash_insert_kernel<<<grid1(n_terms), 256>>>(d_terms, n_terms, d_table, (uint32_t)cap - 1);
    GB_CUDA(cudaGetLastError());
    DevIndex look = text;
    look.term_table = d_table;
    look.term_mask = (uint32_t)cap - 1;

    // 5. posting lists: (term << 32 | slot) sorted
    uint64_t *d_pk, *d_pk_sorted;
    GB_CUDA(tmp.get(&d_pk, n_pairs));
    GB_CUDA(tmp.get(&d_pk_sorted, n_pairs));
    sg_posting_keys_kernel<<<grid1(n_pairs), 256>>>(look, d_pair_key, d_pair_doc, d_slot_of_doc, n_pairs, d_pk);
    GB_CUDA(cudaGetLastError());
    int term_bits = 1;
    while ((1ull << term_bits) < (uint64_t)n_terms && term_bits < 32) term_bits++;
    GB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, need, d_pk, d_pk_sorted, (int)n_pairs, 0, 32 + term_bits));
    GB_CUDA(cub_reserve(need));
    GB_CUDA(cub::DeviceRadixSort::SortKeys(d_cub, need, d_pk, d_pk_sorted, (int)n_pairs, 0, 32 + term_bits));

    // 6b. postings, bitmaps, list offsets
    uint32_t *d_postings, *d_bitmaps = nullptr, *d_list_off;
    unsigned long long *d_n_lists;
    const size_t n_post_alloc = ((size_t)n_pairs + 3) / 4 * 4 + 4;
    GB_CUDA(keep((void **)&d_postings, n_post_alloc * 4));
    GB_CUDA(cudaMemset(d_postings, 0xFF, n_post_alloc * 4));
    if (row_words) {
        GB_CUDA(keep((void **)&d_bitmaps, ((size_t)n_terms + 1) * row_words * 4));
        GB_CUDA(cudaMemset(d_bitmaps, 0, ((size_t)n_terms + 1) * row_words * 4));
    }
    GB_CUDA(keep((void **)&d_list_off, (size_t)n_terms * (S + 1) * 4));
    GB_CUDA(tmp.get(&d_n_lists, 1));
    GB_CUDA(cudaMemset(d_n_lists, 0, 8));
    sg_postings_bitmaps_kernel<<<grid1(n_pairs), 256>>>(d_pk_sorted, n_pairs, d_postings, d_bitmaps, row_words, bs);
    GB_CUDA(cudaGetLastError());
    sg_list_offsets_kernel<<<grid1((uint64_t)n_terms * (S + 1)), 256>>>(d_pk_sorted, n_pairs, d_seg_start, n_terms, S, d_list_off, d_n_lists);
    GB_CUDA(cudaGetLastError());
    unsigned long long n_lists = 0;
    GB_CUDA(cudaMemcpy(&n_lists, d_n_lists, 8, cudaMemcpyDeviceToHost));
    GB_CUDA(cudaDeviceSynchronize());

    out->term_table = d_table;
    out->term_mask = (uint32_t)cap - 1;
    out->n_terms = n_terms;
    out->n_segments = S;
    out->n_ids = n_ids;
    out->bshift = bs;
    out->row_words = row_words;
    out->seg_start = d_seg_start;
    out->list_off = d_list_off;
    out->postings = d_postings;
    out->perm = d_perm;
    out->bitmaps = d_bitmaps;
    out->n_postings = n_pairs;
    out->n_lists = n_lists;
    out->device_bytes = cap * 16 + ((size_t)S + 1) * 4 + (size_t)n_ids * 4 + n_post_alloc * 4 + (size_t)n_terms * (S + 1) * 4 +
                        (row_words ? ((size_t)n_terms + 1) * row_words * 4 : 0);
    out->kernel_launches = 11;
    return "";
}

}  // namespace sg
