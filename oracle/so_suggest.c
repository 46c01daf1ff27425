/*
 * so_suggest.c — metrics, scorer, top-k queue and nGramSuggester.Suggest, restated.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows:
 *   pkg/metric/{jaccard,cosine,dice,overlap,exact}.go     float64, operation for operation
 *   pkg/suggest/scorer.go:29-31                           score = 1 - Distance(overlap, sizeA, sizeB)
 *   pkg/suggest/collector.go:20-26, 117-191               Candidate.Less, fuzzyCollector, FuzzyCollectorManager
 *   pkg/suggest/topk.go:66-175                            bounded heap over container/heap
 *   pkg/suggest/suggester.go:46-131                       segment window, thresholds, dynamic alpha
 *   pkg/index/searcher.go:28-78                           filterTermsByExistence, iterators, merge
 *
 * Build with -ffp-contract=off: the Go amd64 compiler never fuses multiply-add.
 */
#include "so_internal.h"
#include <math.h>
#include <pthread.h>

/* ---------------- pkg/metric ---------------- */
int so_metric_min_y(int m, double alpha, int size) {
    switch (m) {
    case SO_JACCARD: return (int)ceil(alpha * (double)size);
    case SO_COSINE: return (int)ceil(alpha * alpha * (double)size);
    case SO_DICE: return (int)ceil(alpha / (2 - alpha) * (double)size);
    case SO_OVERLAP: return 1;
    case SO_EXACT: return size;
    }
    return 0;
}

/* int(float64) of an out-of-range value is implementation-defined in Go; every caller clamps the
 * result to indices.Size()-1 straight away (suggester.go:57-59), so saturate instead. */
static int so_f2i(double v) {
    if (!(v < 2147483647.0)) return 2147483647;
    if (v < -2147483647.0) return -2147483647;
    return (int)v;
}

int so_metric_max_y(int m, double alpha, int size) {
    switch (m) {
    case SO_JACCARD: return so_f2i(floor((double)size / alpha));
    case SO_COSINE: return so_f2i(floor((double)size / (alpha * alpha)));
    case SO_DICE: return so_f2i(floor((2 - alpha) / alpha * (double)size));
    case SO_OVERLAP: return 32767; /* math.MaxInt16 */
    case SO_EXACT: return size;
    }
    return 0;
}

int so_metric_threshold(int m, double alpha, int a, int b) {
    switch (m) {
    case SO_JACCARD: return so_f2i(ceil(alpha * (double)(a + b) / (1 + alpha)));
    case SO_COSINE: return so_f2i(ceil(alpha * sqrt((double)(a * b))));
    case SO_DICE: return so_f2i(ceil(0.5 * alpha * (double)(a + b)));
    case SO_OVERLAP: return so_f2i(ceil(alpha * fmin((double)a, (double)b)));
    case SO_EXACT: return a;
    }
    return 0;
}

double so_metric_distance(int m, int inter, int a, int b) {
    switch (m) {
    case SO_JACCARD: return 1 - (double)inter / (double)(a + b - inter);
    case SO_COSINE: return 1 - (double)inter / sqrt((double)(a * b));
    case SO_DICE: return 1 - (double)(2 * inter) / (double)(a + b);
    case SO_OVERLAP: return 1 - (double)inter / fmin((double)a, (double)b);
    case SO_EXACT: return 0;
    }
    return 1;
}

double so_score(int m, int inter, int a, int b) { return 1 - so_metric_distance(m, inter, a, b); }

/* ---------------- pkg/suggest/topk.go ---------------- */
typedef struct { uint32_t key; double score; } so_candidate;
typedef struct { so_candidate *h; int len, top_k; } so_queue;

static int cand_less(so_candidate c, so_candidate o) { /* collector.go:20-26 */
    if (c.score == o.score) return c.key > o.key;
    return c.score < o.score;
}
static void q_swap(so_queue *q, int i, int j) { so_candidate t = q->h[i]; q->h[i] = q->h[j]; q->h[j] = t; }
static void q_up(so_queue *q, int j) {
    for (;;) {
        int i = (j - 1) / 2;
        if (i == j || !cand_less(q->h[j], q->h[i])) break;
        q_swap(q, i, j);
        j = i;
    }
}
static int q_down(so_queue *q, int i0, int n) {
    int i = i0;
    for (;;) {
        int j1 = 2 * i + 1;
        if (j1 >= n || j1 < 0) break;
        int j = j1, j2 = j1 + 1;
        if (j2 < n && cand_less(q->h[j2], q->h[j1])) j = j2;
        if (!cand_less(q->h[j], q->h[i])) break;
        q_swap(q, i, j);
        i = j;
    }
    return i > i0;
}
static void q_init(so_queue *q, int k) { q->h = (so_candidate *)malloc(sizeof(so_candidate) * (size_t)(k > 0 ? k : 1)); q->len = 0; q->top_k = k; }
static void q_free(so_queue *q) { free(q->h); }
static int q_full(const so_queue *q) { return q->len == q->top_k; }
static int q_can_take(const so_queue *q, double score) { return !q_full(q) || q->h[0].score <= score; }
static void q_add(so_queue *q, uint32_t pos, double score) { /* topk.go:82-101 */
    if (!q_can_take(q, score)) return;
    so_candidate c = {pos, score};
    if (q->len < q->top_k) {
        q->h[q->len++] = c; /* heap.Push */
        q_up(q, q->len - 1);
        return;
    }
    if (cand_less(q->h[0], c)) { /* updateTop: h[0] = c; heap.Fix(&h, 0) */
        q->h[0] = c;
        if (!q_down(q, 0, q->len)) q_up(q, 0);
    }
}
static double q_lowest(const so_queue *q) { return q->len > 0 ? q->h[0].score : -INFINITY; }
static void q_merge(so_queue *q, const so_queue *other) { /* topk.go:150-165: iterates the heap array in place */
    for (int i = 0; i < other->len; i++) q_add(q, other->h[i].key, other->h[i].score);
}
/* GetCandidates, topk.go:127-147: pop everything, worst first, into a descending array */
static int q_drain(so_queue *q, uint32_t *ids, double *scores) {
    int n = q->len;
    while (q->len > 0) {
        int last = q->len - 1;
        q_swap(q, 0, last);
        q_down(q, 0, last);
        q->len--;
        ids[q->len] = q->h[last].key;
        scores[q->len] = q->h[last].score;
    }
    return n;
}

int so_topk(const uint32_t *ids, const double *scores, uint32_t n, uint32_t k, uint32_t *out_ids,
            double *out_scores, double *lowest_score) {
    so_queue q;
    q_init(&q, (int)k);
    for (uint32_t i = 0; i < n; i++) q_add(&q, ids[i], scores[i]);
    *lowest_score = q_lowest(&q);
    int cnt = q_drain(&q, out_ids, out_scores);
    q_free(&q);
    return cnt;
}

/* ---------------- per-query scratch ---------------- */
typedef struct {
    so_tokens toks;
    so_bytes scratch;
    uint16_t *counts;   /* n_docs, canonical mode */
    uint32_t *touched;
    size_t n_touched, cap_touched;
} so_work;

static void work_init(so_work *w, const so_index *ix) {
    memset(w, 0, sizeof(*w));
    w->counts = (uint16_t *)calloc(ix->n_docs ? ix->n_docs : 1, sizeof(uint16_t));
}
static void work_free(so_work *w) {
    so_tokens_free(&w->toks);
    free(w->scratch.p); free(w->counts); free(w->touched);
}

/* ---------------- canonical search (SURVEY.md §8c rules 5-9) ---------------- */
static int so_suggest_canonical(const so_index *ix, so_work *w, const uint8_t *q, size_t qlen, int metric,
                                double alpha, uint32_t k, uint32_t *out_ids, double *out_scores) {
    so_tokenize_into(ix, q, qlen, &w->toks, &w->scratch);
    int size_a = (int)so_tokens_count(&w->toks);
    if (size_a == 0) return 0;
    int b_min = so_metric_min_y(metric, alpha, size_a), b_max = so_metric_max_y(metric, alpha, size_a);
    int len_indices = (int)ix->n_segs;
    if (b_max >= len_indices) b_max = len_indices - 1;
    so_queue global;
    q_init(&global, (int)k);
    for (int size_b = b_min; size_b <= b_max; size_b++) {
        if (size_b < 0) continue;
        int threshold = so_metric_threshold(metric, alpha, size_a, size_b);
        if (threshold == 0 || threshold > size_b || threshold > size_a) continue;
        if (ix->segs[size_b].used == 0) continue; /* indices.Get(sizeB) == nil */
        int n_present = 0;
        for (int t = 0; t < size_a; t++)
            if (so_index_find(ix, (uint32_t)size_b, w->toks.bytes.p + w->toks.off.p[t],
                              w->toks.off.p[t + 1] - w->toks.off.p[t])) n_present++;
        if (n_present < threshold) continue; /* searcher.go:32-34 */
        w->n_touched = 0;
        for (int t = 0; t < size_a; t++) {
            so_list *l = so_index_find(ix, (uint32_t)size_b, w->toks.bytes.p + w->toks.off.p[t],
                                       w->toks.off.p[t + 1] - w->toks.off.p[t]);
            if (!l) continue;
            uint32_t prev = 0xFFFFFFFFu;
            for (size_t i = 0; i < l->ids.n; i++) {
                uint32_t id = l->ids.p[i];
                if (id == prev) continue; /* an id repeated inside one list counts once */
                prev = id;
                if (w->counts[id] == 0) {
                    if (w->n_touched == w->cap_touched) {
                        w->cap_touched = w->cap_touched ? w->cap_touched * 2 : 1024;
                        w->touched = (uint32_t *)realloc(w->touched, w->cap_touched * sizeof(uint32_t));
                    }
                    w->touched[w->n_touched++] = id;
                }
                w->counts[id]++;
            }
        }
        for (size_t i = 0; i < w->n_touched; i++) {
            uint32_t id = w->touched[i];
            int c = w->counts[id];
            w->counts[id] = 0;
            if (c >= threshold) q_add(&global, id, so_score(metric, c, size_a, size_b));
        }
    }
    int cnt = q_drain(&global, out_ids, out_scores);
    q_free(&global);
    return cnt;
}

/* ---------------- faithful search ---------------- */
typedef struct { so_queue *q; int metric, size_a, size_b; } so_fuzzy_collector;
static int fuzzy_collect(void *ctx, uint64_t cand) { /* collector.go:124-128 */
    so_fuzzy_collector *c = (so_fuzzy_collector *)ctx;
    q_add(c->q, so_cand_pos(cand), so_score(c->metric, so_cand_overlap(cand), c->size_a, c->size_b));
    return 0;
}

static int so_search_segment(const so_index *ix, so_work *w, int size_a, int size_b, int threshold, int algo,
                             so_collect_fn collect, void *ctx) {
    /* filterTermsByExistence, searcher.go:67-78 */
    int n = size_a, n_filtered = 0;
    so_list **filtered = (so_list **)malloc(sizeof(so_list *) * (size_t)n);
    for (int i = 0; i < n && (n_filtered + n - i) >= threshold; i++) {
        so_list *l = so_index_find(ix, (uint32_t)size_b, w->toks.bytes.p + w->toks.off.p[i],
                                   w->toks.off.p[i + 1] - w->toks.off.p[i]);
        if (l) filtered[n_filtered++] = l;
    }
    int rc = 0;
    if (n_filtered >= threshold) {
        so_iter *its = (so_iter *)malloc(sizeof(so_iter) * (size_t)n_filtered);
        so_iter **rid = (so_iter **)malloc(sizeof(so_iter *) * (size_t)n_filtered);
        for (int i = 0; i < n_filtered && rc == 0; i++) { /* resolvePostingList, codec.go:76-88 */
            so_list *l = filtered[i];
            if (l->codec == 0) rc = so_iter_init_vb(&its[i], l->enc, l->enc_len, (int)l->ids.n);
            else if (l->codec == 1) rc = so_iter_init_skipping(&its[i], l->enc, l->enc_len, (int)l->ids.n, 64);
            else so_iter_init_slice(&its[i], l->ids.p, (int)l->ids.n);
            rid[i] = &its[i];
        }
        if (rc == 0) rc = so_merger_merge(algo, rid, n_filtered, threshold, collect, ctx);
        free(its); free(rid);
    }
    free(filtered);
    return rc;
}

static int so_suggest_faithful(const so_index *ix, so_work *w, const uint8_t *q, size_t qlen, int metric,
                               double alpha, uint32_t k, int algo, uint32_t *out_ids, double *out_scores) {
    if (!ix->committed) return -1;
    so_tokenize_into(ix, q, qlen, &w->toks, &w->scratch);
    int size_a = (int)so_tokens_count(&w->toks);
    if (size_a == 0) return 0;
    int b_min = so_metric_min_y(metric, alpha, size_a), b_max = so_metric_max_y(metric, alpha, size_a);
    int len_indices = (int)ix->n_segs;
    if (b_max >= len_indices) b_max = len_indices - 1;
    if (b_max - b_min + 1 <= 0) return 0; /* the reference panics / deadlocks here; return empty */
    so_queue global, local;
    q_init(&global, (int)k);
    q_init(&local, (int)k);
    double similarity = alpha; /* similarityHolder */
    int rc = 0;
    /* the feed order of suggester.go:113-121, consumed here by one worker in order */
    for (int i = size_a, j = size_a + 1; (i >= b_min || j <= b_max) && rc == 0; i--, j++) {
        for (int side = 0; side < 2 && rc == 0; side++) {
            int size_b;
            if (side == 0) { if (i < b_min) continue; size_b = i; }
            else { if (j > b_max) continue; size_b = j; }
            int threshold = so_metric_threshold(metric, similarity, size_a, size_b);
            if (threshold == 0 || threshold > size_b || threshold > size_a) continue;
            if (size_b < 0 || size_b >= len_indices || ix->segs[size_b].used == 0) continue;
            local.len = 0; /* collectorManager.Create() */
            so_fuzzy_collector col = {&local, metric, size_a, size_b};
            rc = so_search_segment(ix, w, size_a, size_b, threshold, algo, fuzzy_collect, &col);
            if (rc != 0) break;
            q_merge(&global, &local);
            double lowest = q_full(&global) ? q_lowest(&global) : -INFINITY; /* GetLowestScore, collector.go:185-191 */
            if (lowest > similarity) similarity = lowest;
        }
    }
    int cnt = rc == 0 ? q_drain(&global, out_ids, out_scores) : -1;
    q_free(&global); q_free(&local);
    return cnt;
}

static int so_suggest_with(const so_index *ix, so_work *w, const char *query, uint32_t qlen, int metric,
                           double alpha, uint32_t k, int mode, int algo, uint32_t *out_ids, double *out_scores) {
    if (k == 0) return -1; /* NewSearchConfig: topK >= 1, search.go:18-21 */
    if (!(alpha > 0) || alpha > 1) return -1;
    if (mode == SO_MODE_CANONICAL)
        return so_suggest_canonical(ix, w, (const uint8_t *)query, qlen, metric, alpha, k, out_ids, out_scores);
    return so_suggest_faithful(ix, w, (const uint8_t *)query, qlen, metric, alpha, k, algo, out_ids, out_scores);
}

int so_suggest(const so_index *ix, const char *query, uint32_t qlen, int metric, double alpha, uint32_t k,
               int mode, int merger_algo, uint32_t *out_ids, double *out_scores) {
    so_work w;
    work_init(&w, ix);
    int rc = so_suggest_with(ix, &w, query, qlen, metric, alpha, k, mode, merger_algo, out_ids, out_scores);
    work_free(&w);
    return rc;
}

/* ---------------- batch: one query per thread at a time ---------------- */
typedef struct {
    const so_index *ix;
    const char *q_bytes;
    const uint64_t *q_off;
    uint32_t n_q;
    int metric, mode, algo;
    double alpha;
    uint32_t k;
    uint32_t *out_ids;
    double *out_scores;
    uint32_t *out_counts;
    uint32_t *next;
    int *failed;
} so_batch;

static void *so_batch_worker(void *arg) {
    so_batch *b = (so_batch *)arg;
    so_work w;
    work_init(&w, b->ix);
    for (;;) {
        uint32_t i = __atomic_fetch_add(b->next, 16, __ATOMIC_RELAXED);
        if (i >= b->n_q) break;
        uint32_t end = i + 16 < b->n_q ? i + 16 : b->n_q;
        for (; i < end; i++) {
            int rc = so_suggest_with(b->ix, &w, b->q_bytes + b->q_off[i], (uint32_t)(b->q_off[i + 1] - b->q_off[i]),
                                     b->metric, b->alpha, b->k, b->mode, b->algo,
                                     b->out_ids + (size_t)i * b->k, b->out_scores + (size_t)i * b->k);
            if (rc < 0) { __atomic_store_n(b->failed, 1, __ATOMIC_RELAXED); rc = 0; }
            b->out_counts[i] = (uint32_t)rc;
        }
    }
    work_free(&w);
    return NULL;
}

int so_suggest_batch(const so_index *ix, const char *q_bytes, const uint64_t *q_off, uint32_t n_q, int metric,
                     double alpha, uint32_t k, int mode, int merger_algo, int n_threads, uint32_t *out_ids,
                     double *out_scores, uint32_t *out_counts) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    uint32_t next = 0;
    int failed = 0;
    so_batch b = {ix, q_bytes, q_off, n_q, metric, mode, merger_algo, alpha, k, out_ids, out_scores, out_counts,
                  &next, &failed};
    pthread_t th[256];
    for (int t = 1; t < n_threads; t++) pthread_create(&th[t], NULL, so_batch_worker, &b);
    so_batch_worker(&b);
    for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
    return failed ? -1 : 0;
}

/* ---------------- autocomplete (pkg/suggest/autocomplete.go:40-77) ---------------- */
int so_autocomplete(const so_index *ix, const char *query, uint32_t qlen, uint32_t limit, uint32_t *out_ids,
                    double *out_scores) {
    if (limit == 0) return -1;
    so_work w;
    work_init(&w, ix);
    so_tokenize_mode(ix, (const uint8_t *)query, qlen, &w.toks, &w.scratch, 0);
    int terms_len = (int)so_tokens_count(&w.toks);
    so_queue q;
    q_init(&q, (int)limit);
    /* termsLen == 0: mergerOptimizer.Merge returns at once for an empty rid (list_merger.go:76-78) */
    for (int size = terms_len; terms_len > 0 && size < (int)ix->n_segs; size++) {
        if (ix->segs[size].used == 0) continue;
        /* searcher.Search with threshold = termsLen: every token must be present, then k-way intersection */
        int present = 1;
        for (int t = 0; t < terms_len && present; t++)
            present = so_index_find(ix, (uint32_t)size, w.toks.bytes.p + w.toks.off.p[t], w.toks.off.p[t + 1] - w.toks.off.p[t]) != NULL;
        if (!present) continue;
        w.n_touched = 0;
        for (int t = 0; t < terms_len; t++) {
            so_list *l = so_index_find(ix, (uint32_t)size, w.toks.bytes.p + w.toks.off.p[t], w.toks.off.p[t + 1] - w.toks.off.p[t]);
            uint32_t prev = 0xFFFFFFFFu;
            for (size_t i = 0; i < l->ids.n; i++) {
                uint32_t id = l->ids.p[i];
                if (id == prev) continue;
                prev = id;
                if (w.counts[id] == 0) {
                    if (w.n_touched == w.cap_touched) {
                        w.cap_touched = w.cap_touched ? w.cap_touched * 2 : 1024;
                        w.touched = (uint32_t *)realloc(w.touched, w.cap_touched * sizeof(uint32_t));
                    }
                    w.touched[w.n_touched++] = id;
                }
                w.counts[id]++;
            }
        }
        for (size_t i = 0; i < w.n_touched; i++) {
            uint32_t id = w.touched[i];
            int c = w.counts[id];
            w.counts[id] = 0;
            if (c >= terms_len) q_add(&q, id, -(double)id); /* FirstKCollectorManager.Collect, collector.go:104-106 */
        }
    }
    int cnt = q_drain(&q, out_ids, out_scores);
    q_free(&q);
    work_free(&w);
    return cnt;
}

/* ---------------- admissible postings / lists of one query ---------------- */
int so_query_stats(const so_index *ix, const char *query, uint32_t qlen, int metric, double alpha,
                   uint64_t *postings, uint64_t *lists, uint32_t *segments, uint32_t *size_a_out) {
    so_tokens toks = {0};
    so_bytes scratch = {0};
    so_tokenize_into(ix, (const uint8_t *)query, qlen, &toks, &scratch);
    int size_a = (int)so_tokens_count(&toks);
    *postings = 0; *lists = 0; *segments = 0; *size_a_out = (uint32_t)size_a;
    if (size_a > 0) {
        int b_min = so_metric_min_y(metric, alpha, size_a), b_max = so_metric_max_y(metric, alpha, size_a);
        if (b_max >= (int)ix->n_segs) b_max = (int)ix->n_segs - 1;
        for (int size_b = b_min < 0 ? 0 : b_min; size_b <= b_max; size_b++) {
            int threshold = so_metric_threshold(metric, alpha, size_a, size_b);
            if (threshold == 0 || threshold > size_b || threshold > size_a) continue;
            if (ix->segs[size_b].used == 0) continue;
            (*segments)++;
            for (int t = 0; t < size_a; t++) {
                so_list *l = so_index_find(ix, (uint32_t)size_b, toks.bytes.p + toks.off.p[t],
                                           toks.off.p[t + 1] - toks.off.p[t]);
                if (l) { (*lists)++; *postings += l->ids.n; }
            }
        }
    }
    so_tokens_free(&toks);
    free(scratch.p);
    return 0;
}
