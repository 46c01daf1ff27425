// microbench.cu — cost of the shared-memory counter update variants considered for sg_search_kernel
// (scattered byte read-modify-write vs 32-bit shared atomics), at the kernel's real geometry.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(uint32_t tbl_bytes, int iters, uint32_t *sink, const uint32_t *ids, uint32_t n_ids) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *tbl = smem + (size_t)warp * tbl_bytes;
    uint32_t *tbl32 = (uint32_t *)tbl;
    for (uint32_t i = lane; i < tbl_bytes / 4; i += 32) tbl32[i] = 0;
    __syncwarp();
    const uint32_t mask = tbl_bytes - 1;
    uint32_t pos = ((blockIdx.x * 32 + warp) * 9973u) % (n_ids - 4096);
    uint32_t acc = 0;
    for (int it = 0; it < iters; it++) {
        // sorted ids as in a posting run: lane handles 4 consecutive
        const uint4 v = *(const uint4 *)(ids + ((pos + it * 128 + lane * 4) % (n_ids - 4096) & ~3u));
        const uint32_t b0 = v.x & mask, b1 = v.y & mask, b2 = v.z & mask, b3 = v.w & mask;
        if (MODE == 0) {  // byte RMW, loads first then stores (current kernel)
            uint32_t c0 = tbl[b0], c1 = tbl[b1], c2 = tbl[b2], c3 = tbl[b3];
            tbl[b0] = (uint8_t)(c0 + (c0 != 255u));
            tbl[b1] = (uint8_t)(c1 + (c1 != 255u));
            tbl[b2] = (uint8_t)(c2 + (c2 != 255u));
            tbl[b3] = (uint8_t)(c3 + (c3 != 255u));
            __syncwarp();
        } else if (MODE == 1) {  // 32-bit shared atomic add, byte field
            atomicAdd(tbl32 + (b0 >> 2), 1u << ((b0 & 3) * 8));
            atomicAdd(tbl32 + (b1 >> 2), 1u << ((b1 & 3) * 8));
            atomicAdd(tbl32 + (b2 >> 2), 1u << ((b2 & 3) * 8));
            atomicAdd(tbl32 + (b3 >> 2), 1u << ((b3 & 3) * 8));
        } else if (MODE == 2) {  // byte stores only
            tbl[b0] = 1; tbl[b1] = 1; tbl[b2] = 1; tbl[b3] = 1;
        } else if (MODE == 3) {  // 32-bit atomic with return value
            acc += atomicAdd(tbl32 + (b0 >> 2), 1u << ((b0 & 3) * 8));
            acc += atomicAdd(tbl32 + (b1 >> 2), 1u << ((b1 & 3) * 8));
            acc += atomicAdd(tbl32 + (b2 >> 2), 1u << ((b2 & 3) * 8));
            acc += atomicAdd(tbl32 + (b3 >> 2), 1u << ((b3 & 3) * 8));
        } else if (MODE == 4) {  // 16-bit field atomics (half as many buckets per byte of table)
            const uint32_t m2 = mask >> 1;
            atomicAdd(tbl32 + ((b0 & m2) >> 1), 1u << ((b0 & 1) * 16));
            atomicAdd(tbl32 + ((b1 & m2) >> 1), 1u << ((b1 & 1) * 16));
            atomicAdd(tbl32 + ((b2 & m2) >> 1), 1u << ((b2 & 1) * 16));
            atomicAdd(tbl32 + ((b3 & m2) >> 1), 1u << ((b3 & 1) * 16));
        } else if (MODE == 5) {  // loads only (global traffic floor of this loop)
            acc += b0 + b1 + b2 + b3;
        }
    }
    __syncwarp();
    for (uint32_t i = lane; i < tbl_bytes / 4; i += 32) acc += tbl32[i];
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int MODE>
void run(const char *name, int warps, uint32_t tbl_bytes, const uint32_t *d_ids, uint32_t n_ids, uint32_t *sink) {
    const int iters = 4096, blocks = 148;
    size_t smem = (size_t)warps * tbl_bytes;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, warps * 32, smem>>>(tbl_bytes, iters, sink, d_ids, n_ids);
    cudaEventRecord(a);
    k<MODE><<<blocks, warps * 32, smem>>>(tbl_bytes, iters, sink, d_ids, n_ids);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    double postings = (double)blocks * warps * iters * 128.0;
    printf("%-34s warps/SM=%2d tbl=%6u  %8.3f ms  %7.1f Gpostings/s  %6.2f ns per warp-iter(128)  %s\n", name, warps, tbl_bytes, ms,
           postings / ms / 1e6, ms * 1e6 / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    const uint32_t n_ids = 20u << 20;  // 80 MB like the index
    std::vector<uint32_t> h(n_ids);
    uint64_t s = 88172645463325252ull;
    uint32_t cur = 0;
    for (uint32_t i = 0; i < n_ids; i++) {  // ascending ids with random gaps ~ 950/20 (one query's merged density is irrelevant here)
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        cur += 1 + (uint32_t)(s % 1900);
        h[i] = cur;
    }
    uint32_t *d_ids, *sink;
    cudaMalloc(&d_ids, (size_t)n_ids * 4);
    cudaMalloc(&sink, 64);
    cudaMemcpy(d_ids, h.data(), (size_t)n_ids * 4, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; rep++) {
        const int warps[3] = {12, 24, 32};
        const uint32_t tbl[3] = {16384, 8192, 4096};
        for (int c = 0; c < 3; c++) {
            run<0>("byte LDS+STS (current)", warps[c], tbl[c], d_ids, n_ids, sink);
            run<1>("atomicAdd u32, 8-bit field", warps[c], tbl[c], d_ids, n_ids, sink);
            run<4>("atomicAdd u32, 16-bit field", warps[c], tbl[c], d_ids, n_ids, sink);
            run<3>("atomicAdd u32 with return", warps[c], tbl[c], d_ids, n_ids, sink);
            run<2>("byte STS only", warps[c], tbl[c], d_ids, n_ids, sink);
            run<5>("global loads only", warps[c], tbl[c], d_ids, n_ids, sink);
        }
    }
    return 0;
}
