/*
 * so_index.c — inverted index by n-gram cardinality segment, restated.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows:
 *   pkg/suggest/indexer.go:14-45          Index(): for every (id, value): AddDocument(id, Tokenize(value))
 *   pkg/index/indexer_writer.go:66-86     AddDocument: indices[len(tokens)][token] = append(.., id)
 *   pkg/index/indexer_writer.go:89-145    Commit: encode each list with the length-class codec
 *   pkg/index/codec.go:39-51              <=65 VB, 66..256 skipping(64), >256 roaring
 *   pkg/index/index_reader.go:84-120      segment without terms stays nil
 */
#include "so_internal.h"

static uint64_t so_hash(const uint8_t *p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

so_index *so_index_new(int ngram_size, const char *wrap_start, const char *wrap_end, const char *pad,
                       const char *const *alphabet, int n_alphabet) {
    if (ngram_size < 1 || ngram_size > SO_MAX_N) return NULL;
    so_index *ix = (so_index *)calloc(1, sizeof(so_index));
    ix->n = ngram_size;
    so_bytes_push(&ix->wrap_start, wrap_start, strlen(wrap_start));
    so_bytes_push(&ix->wrap_end, wrap_end, strlen(wrap_end));
    so_bytes_push(&ix->pad, pad, strlen(pad));
    so_alphabet_init(&ix->alphabet, alphabet, n_alphabet);
    return ix;
}

void so_index_free(so_index *ix) {
    if (!ix) return;
    for (size_t i = 0; i < ix->n_lists; i++) { free(ix->lists[i].ids.p); free(ix->lists[i].enc); }
    free(ix->lists);
    for (size_t s = 0; s < ix->n_segs; s++) free(ix->segs[s].slots);
    free(ix->segs);
    free(ix->term_bytes.p);
    free(ix->wrap_start.p); free(ix->wrap_end.p); free(ix->pad.p);
    so_alphabet_free(&ix->alphabet);
    free(ix);
}

static void so_segmap_grow(so_index *ix, so_segmap *m) {
    size_t ncap = m->cap ? m->cap * 2 : 16;
    int64_t *ns = (int64_t *)malloc(ncap * sizeof(int64_t));
    for (size_t i = 0; i < ncap; i++) ns[i] = -1;
    for (size_t i = 0; i < m->cap; i++) {
        int64_t li = m->slots[i];
        if (li < 0) continue;
        so_list *l = &ix->lists[li];
        size_t h = so_hash(ix->term_bytes.p + l->term_off, l->term_len) & (ncap - 1);
        while (ns[h] >= 0) h = (h + 1) & (ncap - 1);
        ns[h] = li;
    }
    free(m->slots);
    m->slots = ns;
    m->cap = ncap;
}

so_list *so_index_find(const so_index *ix, uint32_t segment, const uint8_t *term, uint32_t term_len) {
    if (segment >= ix->n_segs) return NULL;
    const so_segmap *m = &ix->segs[segment];
    if (!m->cap) return NULL;
    size_t h = so_hash(term, term_len) & (m->cap - 1);
    while (m->slots[h] >= 0) {
        so_list *l = &ix->lists[m->slots[h]];
        if (l->term_len == term_len && memcmp(ix->term_bytes.p + l->term_off, term, term_len) == 0) return l;
        h = (h + 1) & (m->cap - 1);
    }
    return NULL;
}

static so_list *so_index_find_or_add(so_index *ix, uint32_t segment, const uint8_t *term, uint32_t term_len) {
    so_segmap *m = &ix->segs[segment];
    if ((m->used + 1) * 2 > m->cap) so_segmap_grow(ix, m);
    size_t h = so_hash(term, term_len) & (m->cap - 1);
    while (m->slots[h] >= 0) {
        so_list *l = &ix->lists[m->slots[h]];
        if (l->term_len == term_len && memcmp(ix->term_bytes.p + l->term_off, term, term_len) == 0) return l;
        h = (h + 1) & (m->cap - 1);
    }
    if (ix->n_lists == ix->cap_lists) {
        ix->cap_lists = ix->cap_lists ? ix->cap_lists * 2 : 1024;
        ix->lists = (so_list *)realloc(ix->lists, ix->cap_lists * sizeof(so_list));
    }
    so_list *l = &ix->lists[ix->n_lists];
    memset(l, 0, sizeof(*l));
    l->segment = segment;
    l->term_off = (uint32_t)ix->term_bytes.n;
    l->term_len = term_len;
    so_bytes_push(&ix->term_bytes, term, term_len);
    m->slots[h] = (int64_t)ix->n_lists;
    m->used++;
    ix->n_lists++;
    return l;
}

int so_index_add_docs(so_index *ix, const char *bytes, const uint64_t *off, uint32_t n_docs) {
    so_tokens toks = {0};
    so_bytes scratch = {0};
    for (uint32_t d = 0; d < n_docs; d++) {
        uint32_t id = ix->n_docs + d;
        so_tokenize_into(ix, (const uint8_t *)bytes + off[d], (size_t)(off[d + 1] - off[d]), &toks, &scratch);
        size_t card = so_tokens_count(&toks);
        if (ix->n_segs <= card) { /* indexer_writer.go:69-73 */
            ix->segs = (so_segmap *)realloc(ix->segs, (card + 1) * sizeof(so_segmap));
            memset(ix->segs + ix->n_segs, 0, (card + 1 - ix->n_segs) * sizeof(so_segmap));
            ix->n_segs = card + 1;
        }
        for (size_t t = 0; t < card; t++) { /* duplicates after normalisation append the id again */
            so_list *l = so_index_find_or_add(ix, (uint32_t)card, toks.bytes.p + toks.off.p[t],
                                              toks.off.p[t + 1] - toks.off.p[t]);
            so_u32s_push(&l->ids, id);
        }
    }
    ix->n_docs += n_docs;
    so_tokens_free(&toks);
    free(scratch.p);
    return 0;
}

int so_index_commit(so_index *ix) {
    for (size_t i = 0; i < ix->n_lists; i++) {
        so_list *l = &ix->lists[i];
        free(l->enc);
        l->enc = NULL;
        uint32_t n = (uint32_t)l->ids.n;
        if (n <= 64 + 1) l->codec = 0;
        else if (n <= 256) l->codec = 1;
        else { l->codec = 2; continue; } /* roaring class: served from the sorted slice (not restated) */
        uint64_t cap = (uint64_t)n * 5 + 2 * (n / 64 + 2);
        l->enc = (uint8_t *)malloc(cap);
        l->enc_len = so_encode(l->codec == 0 ? SO_CODEC_VB : SO_CODEC_SKIPPING, 64, l->ids.p, n, l->enc, cap);
        if (l->enc_len < 0) return -1;
    }
    ix->committed = 1;
    return 0;
}

uint32_t so_index_segments(const so_index *ix) { return (uint32_t)ix->n_segs; }
uint64_t so_index_lists(const so_index *ix) { return ix->n_lists; }
uint64_t so_index_postings(const so_index *ix) {
    uint64_t s = 0;
    for (size_t i = 0; i < ix->n_lists; i++) s += ix->lists[i].ids.n;
    return s;
}

int64_t so_index_get_list(const so_index *ix, uint32_t segment, const char *term, uint32_t term_len,
                          uint32_t *out, uint64_t cap) {
    so_list *l = so_index_find(ix, segment, (const uint8_t *)term, term_len);
    if (!l) return -1;
    if (out) {
        if (l->ids.n > cap) return -2;
        memcpy(out, l->ids.p, l->ids.n * sizeof(uint32_t));
    }
    return (int64_t)l->ids.n;
}

int so_index_list_at(const so_index *ix, uint64_t i, uint32_t *segment, const char **term, uint32_t *term_len,
                     const uint32_t **ids, uint32_t *n_ids) {
    if (i >= ix->n_lists) return -1;
    const so_list *l = &ix->lists[i];
    *segment = l->segment;
    *term = (const char *)ix->term_bytes.p + l->term_off;
    *term_len = l->term_len;
    *ids = l->ids.p;
    *n_ids = (uint32_t)l->ids.n;
    return 0;
}
