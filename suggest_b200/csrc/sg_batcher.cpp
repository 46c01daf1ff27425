// sg_batcher.cpp — sg_batcher_*: concurrent single-query Suggest calls coalesced into sg_search_batch calls.
//
// The reference's callers issue one query per goroutine (internal/suggest/api/suggest_handler.go:42-76 per HTTP request,
// cmd/suggest/cmd/eval.go:60, pkg/spellchecker/spellchecker.go:67); a GPU call per query would spend ~60 us of launch and
// copy latency on ~4 ns of work.  A batcher owns one worker thread: callers (any number of host threads - cgo calls block
// an OS thread each) append their query to the open batch and sleep; a worker closes a batch when it is full, when its
// oldest query has waited max_wait_us, or - the usual case under load - as soon as its previous batch has returned, so
// the batch size follows the arrival rate by itself.  Two workers (SG_BATCHER_WORKERS) keep two batches in flight: one
// batch's copies and host-side work run under the other's kernels.  The batch runs through sg_search_batch with page-locked buffers
// owned by the batcher (the kernel stores the result rows straight into them; a caller's Go-heap slices never force the
// staged path), the rows are handed out, the callers wake.  Queries of one batch share (metric, similarity); k is the
// largest asked for (a top-k list is a prefix of every longer one).
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/suggest_b200.h"
#include "sg_exchange.h"  // sg_internal_fail

namespace {

struct Request {
    const char *query;
    uint32_t len;
    int metric;
    double alpha;
    uint32_t k;
    uint32_t *out_ids;
    double *out_scores;
    uint32_t *out_count;
    int rc = SG_OK;
    std::string err;
    bool done = false;
    std::chrono::steady_clock::time_point arrived;
};

}  // namespace

struct Worker {  // one batch in flight: a thread and the page-locked staging it owns
    std::thread thread;
    char *q_bytes = nullptr;
    size_t q_cap = 0;
    uint32_t *q_off = nullptr, *ids = nullptr, *counts = nullptr;
    double *scores = nullptr;
};

struct sg_batcher {
    sg_index *ix = nullptr;
    uint32_t max_batch = 0, max_wait_us = 0, max_k = 0;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<Request *> queue;
    bool stop = false;
    std::vector<Worker> workers;
    // statistics
    std::atomic<uint64_t> n_batches{0}, n_queries{0}, max_seen{0};
};

namespace {

void run_batch(sg_batcher *bt, Worker *b, std::vector<Request *> &batch) {
    const uint32_t n = (uint32_t)batch.size();
    uint32_t k = 1;
    size_t total = 0;
    for (Request *r : batch) { if (r->k > k) k = r->k; total += r->len; }
    int rc = SG_OK;
    std::string err;
    if (total > b->q_cap) {  // grow the query staging (rare: queries are short)
        void *p = nullptr;
        size_t cap = total * 2 + 4096;
        if (sg_pinned_alloc(cap, &p) == SG_OK) {
            sg_pinned_free(b->q_bytes);
            b->q_bytes = (char *)p;
            b->q_cap = cap;
        } else {
            rc = SG_ERR_NOMEM;
            err = sg_last_error();
        }
    }
    if (rc == SG_OK) {
        size_t at = 0;
        for (uint32_t i = 0; i < n; i++) {
            b->q_off[i] = (uint32_t)at;
            if (batch[i]->len) std::memcpy(b->q_bytes + at, batch[i]->query, batch[i]->len);
            at += batch[i]->len;
        }
        b->q_off[n] = (uint32_t)at;
        rc = sg_search_batch(bt->ix, b->q_bytes, b->q_off, n, batch[0]->metric, batch[0]->alpha, k, b->ids, b->scores, b->counts);
        if (rc != SG_OK) err = sg_last_error();
    }
    if (rc == SG_ERR_QUERY_TOO_LONG) {
        // one query of the batch was refused: the others have their rows (counts), only the refused ones fail
        rc = SG_OK;
    }
    for (uint32_t i = 0; i < n; i++) {
        Request *r = batch[i];
        r->rc = rc;
        if (rc != SG_OK) { r->err = err; continue; }
        uint32_t c = b->counts[i];
        if (c == SG_COUNT_UNSUPPORTED) { r->rc = SG_ERR_QUERY_TOO_LONG; r->err = "the query has too many n-grams"; continue; }
        if (c > r->k) c = r->k;
        std::memcpy(r->out_ids, b->ids + (size_t)i * k, (size_t)c * sizeof(uint32_t));
        std::memcpy(r->out_scores, b->scores + (size_t)i * k, (size_t)c * sizeof(double));
        *r->out_count = c;
    }
    bt->n_batches.fetch_add(1, std::memory_order_relaxed);
    bt->n_queries.fetch_add(n, std::memory_order_relaxed);
    uint64_t seen = bt->max_seen.load(std::memory_order_relaxed);
    while (n > seen && !bt->max_seen.compare_exchange_weak(seen, n, std::memory_order_relaxed)) {}
}

void worker_loop(sg_batcher *b, Worker *w) {
    std::vector<Request *> batch;
    for (;;) {
        batch.clear();
        {
            std::unique_lock<std::mutex> lk(b->mu);
            b->cv_work.wait(lk, [b] { return b->stop || !b->queue.empty(); });
            if (b->queue.empty()) return;  // stop, nothing left to serve
            // an idle batcher gives the first query's companions max_wait_us to arrive; under load the queue has filled while
            // the previous batch ran and this does not wait at all
            const auto deadline = b->queue.front()->arrived + std::chrono::microseconds(b->max_wait_us);
            while (!b->stop && b->queue.size() < b->max_batch && std::chrono::steady_clock::now() < deadline)
                b->cv_work.wait_until(lk, deadline);
            if (b->queue.empty()) continue;  // another worker took the batch this one was waiting to fill
            const int metric = b->queue.front()->metric;
            const double alpha = b->queue.front()->alpha;
            for (auto it = b->queue.begin(); it != b->queue.end() && batch.size() < b->max_batch;) {
                if ((*it)->metric == metric && (*it)->alpha == alpha) {
                    batch.push_back(*it);
                    it = b->queue.erase(it);
                } else {
                    ++it;
                }
            }
        }
        run_batch(b, w, batch);
        {
            std::lock_guard<std::mutex> lk(b->mu);
            for (Request *r : batch) r->done = true;
        }
        b->cv_done.notify_all();
    }
}

void release(sg_batcher *b) {
    for (Worker &w : b->workers) {
        sg_pinned_free(w.q_bytes);
        sg_pinned_free(w.q_off);
        sg_pinned_free(w.ids);
        sg_pinned_free(w.counts);
        sg_pinned_free(w.scores);
    }
    delete b;
}

}  // namespace

extern "C" {

int sg_batcher_create(sg_index *ix, uint32_t max_batch, uint32_t max_wait_us, uint32_t max_k, sg_batcher **out) {
    if (!out) return sg_internal_fail(SG_ERR_INVALID, "null out");
    *out = nullptr;
    if (!ix) return sg_internal_fail(SG_ERR_INVALID, "null index");
    if (max_batch < 1 || max_batch > (1u << 20) || max_k < 1 || max_k > SG_MAX_TOPK_SHARED) return sg_internal_fail(SG_ERR_INVALID, "max_batch / max_k out of range");
    sg_batcher *b = new (std::nothrow) sg_batcher();
    if (!b) return sg_internal_fail(SG_ERR_NOMEM, "out of host memory");
    b->ix = ix;
    b->max_batch = max_batch;
    b->max_wait_us = max_wait_us;
    b->max_k = max_k;
    int n_workers = 2;  // batches in flight
    if (const char *v = std::getenv("SG_BATCHER_WORKERS")) n_workers = std::atoi(v);
    if (n_workers < 1) n_workers = 1;
    if (n_workers > 8) n_workers = 8;
    b->workers.resize((size_t)n_workers);  // (never resized again: the threads hold pointers into it)
    int rc = SG_OK;
    for (Worker &w : b->workers) {
        w.q_cap = (size_t)max_batch * 64 + 4096;
        void *p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        if (rc == SG_OK) rc = sg_pinned_alloc(w.q_cap, &p[0]);
        if (rc == SG_OK) rc = sg_pinned_alloc(((size_t)max_batch + 1) * sizeof(uint32_t), &p[1]);
        if (rc == SG_OK) rc = sg_pinned_alloc((size_t)max_batch * max_k * sizeof(uint32_t), &p[2]);
        if (rc == SG_OK) rc = sg_pinned_alloc((size_t)max_batch * sizeof(uint32_t), &p[3]);
        if (rc == SG_OK) rc = sg_pinned_alloc((size_t)max_batch * max_k * sizeof(double), &p[4]);
        w.q_bytes = (char *)p[0];
        w.q_off = (uint32_t *)p[1];
        w.ids = (uint32_t *)p[2];
        w.counts = (uint32_t *)p[3];
        w.scores = (double *)p[4];
    }
    if (rc != SG_OK) { release(b); return rc; }
    for (Worker &w : b->workers) w.thread = std::thread(worker_loop, b, &w);
    *out = b;
    return SG_OK;
}

int sg_suggest_one(sg_batcher *b, const char *query, uint32_t len, int metric, double alpha, uint32_t k, uint32_t *out_ids,
                   double *out_scores, uint32_t *out_count) {
    if (!b || !out_ids || !out_scores || !out_count || (len && !query)) return sg_internal_fail(SG_ERR_INVALID, "null argument");
    if (k < 1 || k > b->max_k) return sg_internal_fail(SG_ERR_INVALID, "topK is invalid (above the batcher's max_k?)");      // search.go:18-21
    if (!(alpha > 0.0) || alpha > 1.0) return sg_internal_fail(SG_ERR_INVALID, "similarity shoud be in (0.0, 1.0]");          // search.go:23-25
    if (metric < SG_JACCARD || metric > SG_EXACT) return sg_internal_fail(SG_ERR_INVALID, "unknown metric");
    Request r;
    r.query = query;
    r.len = len;
    r.metric = metric;
    r.alpha = alpha;
    r.k = k;
    r.out_ids = out_ids;
    r.out_scores = out_scores;
    r.out_count = out_count;
    r.arrived = std::chrono::steady_clock::now();
    *out_count = 0;
    {
        std::unique_lock<std::mutex> lk(b->mu);
        if (b->stop) return sg_internal_fail(SG_ERR_INVALID, "the batcher is being freed");
        b->queue.push_back(&r);
        b->cv_work.notify_one();
        b->cv_done.wait(lk, [&r] { return r.done; });
    }
    if (r.rc != SG_OK) return sg_internal_fail(r.rc, r.err);
    return SG_OK;
}

int sg_batcher_get_stats(const sg_batcher *b, sg_batcher_stats *stats) {
    if (!b || !stats) return sg_internal_fail(SG_ERR_INVALID, "null argument");
    stats->batches = b->n_batches.load(std::memory_order_relaxed);
    stats->queries = b->n_queries.load(std::memory_order_relaxed);
    stats->largest_batch = (uint32_t)b->max_seen.load(std::memory_order_relaxed);
    stats->max_batch = b->max_batch;
    stats->max_wait_us = b->max_wait_us;
    stats->reserved = 0;
    return SG_OK;
}

void sg_batcher_free(sg_batcher *b) {
    if (!b) return;
    {
        std::lock_guard<std::mutex> lk(b->mu);
        b->stop = true;
    }
    b->cv_work.notify_all();
    for (Worker &w : b->workers)
        if (w.thread.joinable()) w.thread.join();  // serve what is still queued, then return
    release(b);
}

}  // extern "C"
