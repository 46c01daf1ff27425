// sg_submit.cpp — sg_search_batch_candidates_submit / sg_ticket_wait: batches in flight from ONE host thread.
//
// A synchronous caller of sg_search_batch_candidates leaves the GPU idle for about a third of every call (first copy in,
// last copy out, wake-up, ~100 us of enqueueing on the host).  The reference gets its overlap from goroutines
// (internal/suggest/api/suggest_handler.go:42-76: one per request); a host that drives the library from one thread - a
// batch pipeline, a Python process - gets it here: submit returns at once with a ticket, a few native worker threads per
// index (SG_SUBMIT_WORKERS, default 3) run the very same synchronous call, wait returns the call's status.  The callers'
// buffers belong to the library from submit to wait.
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/suggest_b200.h"
#include "sg_exchange.h"  // sg_internal_fail

struct sg_ticket {
    // the call
    sg_index *ix;
    const char *q_bytes;
    const uint32_t *q_off;
    uint32_t n_q;
    int metric;
    double alpha;
    uint32_t k;
    sg_candidate *out_rows;
    uint32_t *out_counts;
    // its outcome
    std::mutex mu;
    std::condition_variable cv;
    bool done = false;
    int rc = SG_OK;
    std::string err;
};

namespace {

struct Pool {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<sg_ticket *> queue;
    bool stop = false;
    std::vector<std::thread> workers;
};

std::mutex g_pools_mu;
std::unordered_map<sg_index *, Pool *> g_pools;  // one pool per index that was ever submitted to

void worker_loop(Pool *p) {
    for (;;) {
        sg_ticket *t = nullptr;
        {
            std::unique_lock<std::mutex> lk(p->mu);
            p->cv.wait(lk, [p] { return p->stop || !p->queue.empty(); });
            if (p->queue.empty()) return;  // stop, nothing left to serve
            t = p->queue.front();
            p->queue.pop_front();
        }
        const int rc = sg_search_batch_candidates(t->ix, t->q_bytes, t->q_off, t->n_q, t->metric, t->alpha, t->k, t->out_rows, t->out_counts);
        std::string err = rc != SG_OK ? sg_last_error() : "";
        {
            std::lock_guard<std::mutex> lk(t->mu);
            t->rc = rc;
            t->err.swap(err);
            t->done = true;
            t->cv.notify_all();  // under the lock: the waiter deletes the ticket as soon as it holds the mutex again
        }
    }
}

Pool *pool_of(sg_index *ix) {
    std::lock_guard<std::mutex> lk(g_pools_mu);
    auto it = g_pools.find(ix);
    if (it != g_pools.end()) return it->second;
    Pool *p = new (std::nothrow) Pool();
    if (!p) return nullptr;
    int n = 3;
    if (const char *v = std::getenv("SG_SUBMIT_WORKERS")) n = std::atoi(v);
    if (n < 1) n = 1;
    if (n > 16) n = 16;
    for (int i = 0; i < n; i++) p->workers.emplace_back(worker_loop, p);
    g_pools[ix] = p;
    return p;
}

}  // namespace

// sg_index_free calls this first: every submitted call is served, then the workers stop
extern "C" void sg_internal_drop_submit_pool(sg_index *ix) {
    Pool *p = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_pools_mu);
        auto it = g_pools.find(ix);
        if (it == g_pools.end()) return;
        p = it->second;
        g_pools.erase(it);
    }
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop = true;
    }
    p->cv.notify_all();
    for (std::thread &w : p->workers)
        if (w.joinable()) w.join();
    delete p;
}

extern "C" {

int sg_search_batch_candidates_submit(sg_index *ix, const char *q_bytes, const uint32_t *q_off, uint32_t n_q, int metric, double alpha,
                                      uint32_t k, sg_candidate *out_rows, uint32_t *out_counts, sg_ticket **ticket) {
    if (!ticket) return sg_internal_fail(SG_ERR_INVALID, "null ticket");
    *ticket = nullptr;
    if (!ix) return sg_internal_fail(SG_ERR_INVALID, "null index");
    Pool *p = pool_of(ix);
    sg_ticket *t = new (std::nothrow) sg_ticket();
    if (!p || !t) { delete t; return sg_internal_fail(SG_ERR_NOMEM, "out of host memory"); }
    t->ix = ix;
    t->q_bytes = q_bytes;
    t->q_off = q_off;
    t->n_q = n_q;
    t->metric = metric;
    t->alpha = alpha;
    t->k = k;
    t->out_rows = out_rows;
    t->out_counts = out_counts;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        if (p->stop) { delete t; return sg_internal_fail(SG_ERR_INVALID, "the index is being freed"); }
        p->queue.push_back(t);
    }
    p->cv.notify_one();
    *ticket = t;
    return SG_OK;
}

int sg_ticket_wait(sg_ticket *t) {
    if (!t) return sg_internal_fail(SG_ERR_INVALID, "null ticket");
    {
        std::unique_lock<std::mutex> lk(t->mu);
        t->cv.wait(lk, [t] { return t->done; });
    }
    const int rc = t->rc;
    const std::string err = t->err;
    delete t;
    return rc == SG_OK ? SG_OK : sg_internal_fail(rc, err);
}

}  // extern "C"
