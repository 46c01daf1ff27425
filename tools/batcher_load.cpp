// batcher_load.cpp — load generator for sg_batcher_* (include/suggest_b200.h): T host threads, each issuing ONE query per
// call (sg_suggest_one), the way the reference's HTTP handlers call Service.Suggest (internal/suggest/api/suggest_handler.go:56).
// Reports aggregate queries/s and the per-call latency distribution as one JSON line.
// build: g++ -O2 -std=c++17 tools/batcher_load.cpp -Iinclude -Lsuggest_b200 -lsuggest_b200 -Wl,-rpath,$PWD/suggest_b200 -lpthread -o tools/batcher_load
// usage: tools/batcher_load workload.bin threads calls_per_thread max_batch max_wait_us
//   workload.bin (written by bench.py): u32 n_docs, u64 doc_off[n_docs+1], doc bytes, u32 n_q, u32 q_off[n_q+1], query bytes
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "suggest_b200.h"

static bool read_all(FILE *f, void *p, size_t n) { return fread(p, 1, n, f) == n; }

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: %s workload.bin threads calls_per_thread max_batch max_wait_us\n", argv[0]); return 2; }
    const int threads = atoi(argv[2]), calls = atoi(argv[3]);
    const uint32_t max_batch = (uint32_t)atoi(argv[4]), max_wait = (uint32_t)atoi(argv[5]);
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror("workload"); return 2; }
    uint32_t n_docs = 0, n_q = 0;
    read_all(f, &n_docs, 4);
    std::vector<uint64_t> doc_off(n_docs + 1);
    read_all(f, doc_off.data(), doc_off.size() * 8);
    std::vector<char> docs(doc_off[n_docs]);
    read_all(f, docs.data(), docs.size());
    read_all(f, &n_q, 4);
    std::vector<uint32_t> q_off(n_q + 1);
    read_all(f, q_off.data(), q_off.size() * 4);
    std::vector<char> qs(q_off[n_q]);
    if (!read_all(f, qs.data(), qs.size())) { fprintf(stderr, "short workload file\n"); return 2; }
    fclose(f);
    const char *alphabet[] = {"english", "russian", "numbers", "$"};
    sg_config cfg{3, "$", "$", "$", alphabet, 4, 0};
    sg_index *ix = nullptr;
    if (sg_index_build(&cfg, docs.data(), doc_off.data(), n_docs, 0, &ix) != SG_OK) { fprintf(stderr, "build: %s\n", sg_last_error()); return 1; }
    sg_batcher *b = nullptr;
    const uint32_t k = 10;
    if (sg_batcher_create(ix, max_batch, max_wait, k, &b) != SG_OK) { fprintf(stderr, "batcher: %s\n", sg_last_error()); return 1; }
    std::vector<std::vector<float>> lat(threads);
    std::vector<uint64_t> found(threads, 0);
    auto body = [&](int t, int n_calls, bool record) {
        uint32_t ids[k];
        double scores[k];
        uint32_t count = 0;
        if (record) lat[t].reserve(n_calls);
        for (int i = 0; i < n_calls; i++) {
            const uint32_t q = (uint32_t)(((uint64_t)t * 7919u + (uint64_t)i * 104729u) % n_q);
            const auto t0 = std::chrono::steady_clock::now();
            const int rc = sg_suggest_one(b, qs.data() + q_off[q], q_off[q + 1] - q_off[q], SG_JACCARD, 0.5, k, ids, scores, &count);
            const auto t1 = std::chrono::steady_clock::now();
            if (rc != SG_OK) { fprintf(stderr, "sg_suggest_one: %s\n", sg_last_error()); exit(1); }
            if (record) { lat[t].push_back(std::chrono::duration<float, std::micro>(t1 - t0).count()); found[t] += count; }
        }
    };
    {   // warm-up
        std::vector<std::thread> w;
        for (int t = 0; t < threads; t++) w.emplace_back(body, t, std::max(calls / 10, 10), false);
        for (auto &th : w) th.join();
    }
    sg_batcher_stats s0{}, s1{};
    sg_batcher_get_stats(b, &s0);
    const auto w0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(body, t, calls, true);
    for (auto &th : pool) th.join();
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
    sg_batcher_get_stats(b, &s1);
    std::vector<float> all;
    uint64_t total_found = 0;
    for (int t = 0; t < threads; t++) { all.insert(all.end(), lat[t].begin(), lat[t].end()); total_found += found[t]; }
    std::sort(all.begin(), all.end());
    auto pct = [&](double p) { return all.empty() ? 0.0f : all[std::min(all.size() - 1, (size_t)(p * all.size()))]; };
    const double n_calls = (double)threads * calls;
    printf("{\"threads\": %d, \"calls\": %.0f, \"qps\": %.1f, \"p50_us\": %.1f, \"p90_us\": %.1f, \"p99_us\": %.1f, \"max_us\": %.1f, "
           "\"batches\": %llu, \"mean_batch\": %.1f, \"largest_batch\": %u, \"max_batch\": %u, \"max_wait_us\": %u, \"candidates_per_query\": %.3f}\n",
           threads, n_calls, n_calls / wall, pct(0.5), pct(0.9), pct(0.99), all.empty() ? 0.0f : all.back(),
           (unsigned long long)(s1.batches - s0.batches), (double)(s1.queries - s0.queries) / std::max<double>(1.0, (double)(s1.batches - s0.batches)),
           s1.largest_batch, max_batch, max_wait, (double)total_found / n_calls);
    sg_batcher_free(b);
    sg_index_free(ix);
    return 0;
}
