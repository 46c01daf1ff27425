// l2bench.cu — L2 -> SM read bandwidth of a B200 for an L2-resident working set, the denominator of bench.py's
// roofline_engine (the bucket bitmaps of config #2 are 19.4 MB and stay in the 126 MB L2; the buffer here is 16 MB so that the
// addressing is masks, not divisions: with LDG.32 the loop would otherwise be bound by its own arithmetic).
// Every warp streams whole rows (1 KB, like a term's bitmap row) picked pseudo-randomly from a `mb`-MB buffer with
// coalesced loads of 4 / 8 / 16 bytes per lane, `unroll` independent loads in flight per lane.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2bench tools/l2bench.cu
// usage: tools/l2bench [out.json]
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

template <typename V, int UNROLL>
__global__ void __launch_bounds__(256) read_rows(const V *__restrict__ buf, uint32_t n_rows, uint32_t row_vecs, int iters, uint32_t *sink) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t acc = 0, r = warp * 2654435761u;
    for (int it = 0; it < iters; it++) {
        V v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            r = r * 1664525u + 1013904223u;
            const uint32_t row = (r >> 8) & (n_rows - 1);  // n_rows is a power of two
            // one warp-load covers 32 consecutive vectors of the row; rows are row_vecs vectors long
            v[u] = __ldg(buf + (size_t)row * row_vecs + (((it * UNROLL + u) * 32 + lane) & (row_vecs - 1)));
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const uint32_t *w = (const uint32_t *)&v[u];
            for (int i = 0; i < (int)(sizeof(V) / 4); i++) acc ^= w[i];
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <typename V, int UNROLL>
double run(const char *name, const void *buf, size_t bytes, int ctas_per_sm, uint32_t *sink, FILE *json, bool first) {
    const int iters = 2048 / UNROLL * 8, blocks = 148 * ctas_per_sm;
    const uint32_t row_vecs = 1024 / sizeof(V), n_rows = (uint32_t)(bytes / 1024);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    read_rows<V, UNROLL><<<blocks, 256>>>((const V *)buf, n_rows, row_vecs, iters, sink);  // warms L2
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(a);
        read_rows<V, UNROLL><<<blocks, 256>>>((const V *)buf, n_rows, row_vecs, iters, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double total = (double)blocks * 8 * iters * UNROLL * 32 * sizeof(V);
    const double gbs = total / best / 1e6;
    printf("%-10s unroll %2d  %2d CTAs/SM (%2d warps)  %8.3f ms  %8.1f GB/s  %s\n", name, UNROLL, ctas_per_sm, ctas_per_sm * 8, best, gbs,
           cudaGetLastError() == cudaSuccess ? "" : "ERROR");
    if (json) fprintf(json, "%s{\"load\": \"%s\", \"unroll\": %d, \"warps_per_sm\": %d, \"gbs\": %.1f}", first ? "" : ", ", name, UNROLL, ctas_per_sm * 8, gbs);
    return gbs;
}

int main(int argc, char **argv) {
    const size_t mb = 16, bytes = mb << 20;
    void *buf;
    uint32_t *sink;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&sink, 64);
    cudaMemset(buf, 1, bytes);
    FILE *json = argc > 1 ? fopen(argv[1], "w") : nullptr;
    if (json) fprintf(json, "{\"working_set_mb\": %zu, \"runs\": [", mb);
    double best = 0;
    bool first = true;
    for (int ctas : {2, 4, 6, 8}) {
        double g;
        g = run<uint32_t, 8>("LDG.32", buf, bytes, ctas, sink, json, first); first = false; if (g > best) best = g;
        g = run<uint32_t, 16>("LDG.32", buf, bytes, ctas, sink, json, first); if (g > best) best = g;
        g = run<uint2, 8>("LDG.64", buf, bytes, ctas, sink, json, first); if (g > best) best = g;
        g = run<uint4, 4>("LDG.128", buf, bytes, ctas, sink, json, first); if (g > best) best = g;
        g = run<uint4, 8>("LDG.128", buf, bytes, ctas, sink, json, first); if (g > best) best = g;
    }
    printf("best L2 -> SM read bandwidth over a %zu MB working set: %.1f GB/s\n", mb, best);
    if (json) {
        fprintf(json, "], \"l2_read_gbs\": %.1f, \"how\": \"tools/l2bench.cu: warps stream random 1 KB rows of a %zu MB buffer (L2-resident), best of LDG.32/64/128 x 16-64 warps per SM, CUDA events\"}\n", best, mb);
        fclose(json);
    }
    return 0;
}
