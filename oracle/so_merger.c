/*
 * so_merger.c — the four T-occurrence solvers and the intersector, restated.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows:
 *   pkg/merger/list_merger.go:73-85       mergerOptimizer.Merge (n<T -> nothing, n==T -> intersect)
 *   pkg/merger/list_intersector.go:23-81  Intersect
 *   pkg/merger/scan_count.go:14-88        ScanCount
 *   pkg/merger/cp_merge.go:19-120         CPMerge (production choice, pkg/suggest/ngram_index_builder.go:69)
 *   pkg/merger/merge_skip.go:52-151       MergeSkip
 *   pkg/merger/divide_skip.go:25-74       DivideSkip
 */
#include "so_internal.h"
#include <math.h>

/* sort.Sort(rid) by Len(): Go's sort is not stable; ties may land in any order there.  A stable
 * insertion sort is one of the permitted outcomes. */
static void so_sort_by_len(so_iter **rid, int n, int descending) {
    for (int i = 1; i < n; i++) {
        so_iter *x = rid[i];
        int lx = x->vt->len(x), j = i - 1;
        while (j >= 0 && (descending ? rid[j]->vt->len(rid[j]) < lx : rid[j]->vt->len(rid[j]) > lx)) {
            rid[j + 1] = rid[j];
            j--;
        }
        rid[j + 1] = x;
    }
}

static int so_increment(uint64_t *c) {
    if ((uint32_t)(*c & 0xFFFFFFFFu) == SO_MAX_OVERLAP) return -1; /* panic("overlap overflow") */
    (*c)++;
    return 0;
}

/* one step of the 2-way merge shared by scan_count.go:19-68 and cp_merge.go:32-81 */
static int so_merge_list_into(so_iter *list, so_u64s *candidates, so_u64s *tmp) {
    int is_valid = 1;
    uint32_t current;
    int rc = list->vt->get(list, &current);
    if (rc == SO_IT_NOT_DEREF) is_valid = 0; else if (rc != SO_IT_OK) return -1;
    tmp->n = 0;
    size_t j = 0, end = candidates->n;
    while (j < end || is_valid) {
        if (j >= end || (is_valid && so_cand_pos(candidates->p[j]) > current)) {
            so_u64s_push(tmp, so_cand(current, 1));
            if (list->vt->has_next(list)) {
                if (list->vt->next(list, &current) != SO_IT_OK) return -1;
            } else is_valid = 0;
        } else if (!is_valid || (j < end && so_cand_pos(candidates->p[j]) < current)) {
            so_u64s_push(tmp, candidates->p[j]);
            j++;
        } else {
            if (so_increment(&candidates->p[j]) < 0) return -1;
            so_u64s_push(tmp, candidates->p[j]);
            j++;
            if (list->vt->has_next(list)) {
                if (list->vt->next(list, &current) != SO_IT_OK) return -1;
            } else is_valid = 0;
        }
    }
    so_u64s sw = *candidates; *candidates = *tmp; *tmp = sw;
    return 0;
}

static int so_emit_ge(const so_u64s *candidates, int threshold, so_collect_fn collect, void *ctx) {
    for (size_t i = 0; i < candidates->n; i++) {
        if (so_cand_overlap(candidates->p[i]) >= threshold) {
            int rc = collect(ctx, candidates->p[i]);
            if (rc == 1) return 0; /* ErrCollectionTerminated is swallowed */
            if (rc < 0) return rc;
        }
    }
    return 0;
}

static int so_scan_count(so_iter **rid, int n, int threshold, so_collect_fn collect, void *ctx) {
    so_u64s candidates = {0}, tmp = {0};
    int rc = 0;
    for (int i = 0; i < n && rc == 0; i++) rc = so_merge_list_into(rid[i], &candidates, &tmp);
    if (rc == 0) rc = so_emit_ge(&candidates, threshold, collect, ctx);
    free(candidates.p); free(tmp.p);
    return rc;
}

static int so_cp_merge(so_iter **rid, int n, int threshold, so_collect_fn collect, void *ctx) {
    int min_queries = n - threshold + 1;
    so_sort_by_len(rid, n, 0);
    so_u64s candidates = {0}, tmp = {0};
    int rc = 0;
    for (int i = 0; i < min_queries && rc == 0; i++) rc = so_merge_list_into(rid[i], &candidates, &tmp);
    for (int i = min_queries; i < n && candidates.n > 0 && rc == 0; i++) { /* cp_merge.go:83-103 */
        tmp.n = 0;
        for (size_t c = 0; c < candidates.n; c++) {
            uint64_t cand = candidates.p[c];
            uint32_t current;
            int lr = rid[i]->vt->lower_bound(rid[i], so_cand_pos(cand), &current);
            if (lr == SO_IT_OK && current == so_cand_pos(cand)) {
                if (so_increment(&cand) < 0) { rc = -1; break; }
            }
            if (lr != SO_IT_OK && lr != SO_IT_NOT_DEREF) { rc = -1; break; }
            if (so_cand_overlap(cand) + (n - i - 1) >= threshold) so_u64s_push(&tmp, cand);
        }
        so_u64s sw = candidates; candidates = tmp; tmp = sw;
    }
    if (rc == 0) rc = so_emit_ge(&candidates, threshold, collect, ctx);
    free(candidates.p); free(tmp.p);
    return rc;
}

/* ---- MergeSkip: container/heap over records (ridID, position) keyed by position ---- */
typedef struct { uint32_t rid_id, position; } so_record;
typedef struct { so_record *slice; int size; } so_rheap;

static void rh_swap(so_rheap *h, int i, int j) { so_record t = h->slice[i]; h->slice[i] = h->slice[j]; h->slice[j] = t; }
static int rh_less(so_rheap *h, int i, int j) { return h->slice[i].position < h->slice[j].position; }
static void rh_up(so_rheap *h, int j) { /* container/heap.up */
    for (;;) {
        int i = (j - 1) / 2;
        if (i == j || !rh_less(h, j, i)) break;
        rh_swap(h, i, j);
        j = i;
    }
}
static int rh_down(so_rheap *h, int i0, int n) { /* container/heap.down */
    int i = i0;
    for (;;) {
        int j1 = 2 * i + 1;
        if (j1 >= n || j1 < 0) break;
        int j = j1, j2 = j1 + 1;
        if (j2 < n && rh_less(h, j2, j1)) j = j2;
        if (!rh_less(h, j, i)) break;
        rh_swap(h, i, j);
        i = j;
    }
    return i > i0;
}
static void rh_init(so_rheap *h) { int n = h->size; for (int i = n / 2 - 1; i >= 0; i--) rh_down(h, i, n); }
static void rh_pop(so_rheap *h) { /* heap.Pop: swap(0,n-1); down(0,n-1); h.Pop() => size-- */
    int n = h->size - 1;
    rh_swap(h, 0, n);
    rh_down(h, 0, n);
    h->size--;
}
static void rh_push(so_rheap *h) { /* heap.Push: h.Push(x) => size++; up(Len()-1) */
    h->size++;
    rh_up(h, h->size - 1);
}

static int so_merge_skip(so_iter **rid, int n, int threshold, so_collect_fn collect, void *ctx) {
    so_rheap h;
    h.slice = (so_record *)calloc((size_t)(n > 0 ? n : 1), sizeof(so_record));
    h.size = n;
    int ret = 0;
    for (int i = 0; i < n; i++) {
        uint32_t r = 0;
        int rc = rid[i]->vt->get(rid[i], &r);
        if (rc != SO_IT_OK && rc != SO_IT_NOT_DEREF) { free(h.slice); return -1; }
        h.slice[i].rid_id = (uint32_t)i;
        h.slice[i].position = r;
    }
    rh_init(&h);
    while (h.size > 0) {
        int popped = 0;
        so_record t = h.slice[0];
        while (h.size > 0 && t.position >= h.slice[0].position) { rh_pop(&h); popped++; }
        if (popped >= threshold) {
            int rc = collect(ctx, so_cand(t.position, (uint32_t)popped));
            if (rc == 1) break;
            if (rc < 0) { ret = rc; break; }
            int start = h.size;
            for (int i = 0; i < popped; i++) {
                so_record item = h.slice[start + i];
                so_iter *cur = rid[item.rid_id];
                if (cur->vt->has_next(cur)) {
                    uint32_t r;
                    if (cur->vt->next(cur, &r) != SO_IT_OK) { ret = -1; goto done; }
                    h.slice[h.size].rid_id = item.rid_id;
                    h.slice[h.size].position = r;
                    rh_push(&h);
                }
            }
        } else {
            for (int j = threshold - 1 - popped; j > 0 && h.size > 0; j--) { rh_pop(&h); popped++; }
            if (h.size == 0) break;
            uint32_t top_pos = h.slice[0].position;
            int start = h.size;
            for (int i = 0; i < popped; i++) {
                so_record item = h.slice[start + i];
                so_iter *cur = rid[item.rid_id];
                if (cur->vt->len(cur) == 0) continue;
                uint32_t r;
                int rc = cur->vt->lower_bound(cur, top_pos, &r);
                if (rc != SO_IT_OK && rc != SO_IT_NOT_DEREF) { ret = -1; goto done; }
                if (rc == SO_IT_OK) {
                    h.slice[h.size].rid_id = item.rid_id;
                    h.slice[h.size].position = r;
                    rh_push(&h);
                }
            }
        }
    }
done:
    free(h.slice);
    return ret;
}

static int so_collect_simple(void *ctx, uint64_t c) { so_u64s_push((so_u64s *)ctx, c); return 0; }

static int so_divide_skip(so_iter **rid, int n, int threshold, double mu, so_collect_fn collect, void *ctx) {
    so_sort_by_len(rid, n, 1); /* sort.Reverse */
    double M = (double)rid[0]->vt->len(rid[0]);
    int l = (int)((double)threshold / (mu * log(M) + 1));
    if (l > n) l = n;
    so_iter **l_long = rid, **l_short = rid + l;
    int n_short = n - l;
    if (n_short == 0) return so_merger_merge(SO_MERGE_SKIP, rid, n, threshold, collect, ctx);
    so_u64s res = {0};
    int rc = so_merger_merge(SO_MERGE_SKIP, l_short, n_short, threshold - l, so_collect_simple, &res);
    for (size_t i = 0; i < res.n && rc == 0; i++) {
        uint64_t c = res.p[i];
        uint32_t position = so_cand_pos(c);
        for (int k = 0; k < l; k++) {
            uint32_t r;
            int lr = l_long[k]->vt->lower_bound(l_long[k], position, &r);
            if (lr != SO_IT_OK && lr != SO_IT_NOT_DEREF) { rc = -1; break; }
            if (lr == SO_IT_OK && r == position && so_increment(&c) < 0) { rc = -1; break; }
        }
        if (rc == 0 && so_cand_overlap(c) >= threshold) {
            int cr = collect(ctx, c);
            if (cr == 1) break;
            if (cr < 0) rc = cr;
        }
    }
    free(res.p);
    return rc;
}

int so_merger_intersect(so_iter **rid, int n, so_collect_fn collect, void *ctx) {
    if (n == 0) return 0;
    so_sort_by_len(rid, n, 0);
    so_iter *first = rid[0];
    uint32_t item;
    int rc = first->vt->get(first, &item);
    if (rc != SO_IT_OK) return rc == SO_IT_NOT_DEREF ? -2 : -1; /* the Go code returns this error */
    for (;;) {
        int good = 1;
        for (int k = 1; k < n; k++) {
            uint32_t lower;
            int lr = rid[k]->vt->lower_bound(rid[k], item, &lower);
            if (lr == SO_IT_NOT_DEREF || (lr == SO_IT_OK && lower != item)) { good = 0; break; }
            if (lr != SO_IT_OK) return -1;
        }
        if (good) {
            int cr = collect(ctx, so_cand(item, (uint32_t)n));
            if (cr == 1) return 0;
            if (cr < 0) return cr;
        }
        if (!first->vt->has_next(first)) break;
        if (first->vt->next(first, &item) != SO_IT_OK) return -1;
    }
    return 0;
}

/* mergerOptimizer.Merge, list_merger.go:73-85 */
int so_merger_merge(int algo, so_iter **rid, int n, int threshold, so_collect_fn collect, void *ctx) {
    if (n < threshold || n == 0 || threshold < 0) return 0;
    if (n == threshold) return so_merger_intersect(rid, n, collect, ctx);
    switch (algo) {
    case SO_SCAN_COUNT: return so_scan_count(rid, n, threshold, collect, ctx);
    case SO_CP_MERGE: return so_cp_merge(rid, n, threshold, collect, ctx);
    case SO_MERGE_SKIP: return so_merge_skip(rid, n, threshold, collect, ctx);
    case SO_DIVIDE_SKIP: return so_divide_skip(rid, n, threshold, 0.01, collect, ctx);
    }
    return -1;
}

/* ---- flat-array entry points for the known-answer tests ---- */
typedef struct { uint64_t *out; uint64_t cap, n; int overflow; } so_out_ctx;
static int so_collect_out(void *ctx, uint64_t c) {
    so_out_ctx *o = (so_out_ctx *)ctx;
    if (o->n >= o->cap) { o->overflow = 1; return 1; }
    o->out[o->n++] = c;
    return 0;
}

int64_t so_merge(int algo, const uint32_t *ids, const uint32_t *off, uint32_t n_lists, int threshold,
                 uint64_t *out, uint64_t cap) {
    so_iter *its = (so_iter *)calloc(n_lists ? n_lists : 1, sizeof(so_iter));
    so_iter **rid = (so_iter **)calloc(n_lists ? n_lists : 1, sizeof(so_iter *));
    for (uint32_t i = 0; i < n_lists; i++) {
        so_iter_init_slice(&its[i], ids + off[i], (int)(off[i + 1] - off[i]));
        rid[i] = &its[i];
    }
    so_out_ctx o = {out, cap, 0, 0};
    int rc = so_merger_merge(algo, rid, (int)n_lists, threshold, so_collect_out, &o);
    free(its); free(rid);
    if (rc < 0 || o.overflow) return -1;
    return (int64_t)o.n;
}

int64_t so_intersect(const uint32_t *ids, const uint32_t *off, uint32_t n_lists, uint64_t *out, uint64_t cap) {
    so_iter *its = (so_iter *)calloc(n_lists ? n_lists : 1, sizeof(so_iter));
    so_iter **rid = (so_iter **)calloc(n_lists ? n_lists : 1, sizeof(so_iter *));
    for (uint32_t i = 0; i < n_lists; i++) {
        so_iter_init_slice(&its[i], ids + off[i], (int)(off[i + 1] - off[i]));
        rid[i] = &its[i];
    }
    so_out_ctx o = {out, cap, 0, 0};
    int rc = so_merger_intersect(rid, (int)n_lists, so_collect_out, &o);
    free(its); free(rid);
    if (rc < 0 || o.overflow) return -1;
    return (int64_t)o.n;
}
