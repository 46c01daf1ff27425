/* so_internal.h — shared declarations of the CPU oracle.  TEST INFRASTRUCTURE ONLY (see suggest_oracle.h). */
#ifndef SO_INTERNAL_H
#define SO_INTERNAL_H

#include "suggest_oracle.h"
#include <stdlib.h>
#include <string.h>

#define SO_RUNE_ERROR 0xFFFDu
#define SO_MAX_N 8 /* pkg/analysis/ngram_tokenizer.go:3 */

/* ---------------- growable buffers ---------------- */
typedef struct { uint8_t *p; size_t n, cap; } so_bytes;
typedef struct { uint32_t *p; size_t n, cap; } so_u32s;
typedef struct { uint64_t *p; size_t n, cap; } so_u64s;

static inline void so_bytes_reserve(so_bytes *b, size_t extra) {
    if (b->n + extra > b->cap) {
        size_t c = b->cap ? b->cap * 2 : 64;
        while (c < b->n + extra) c *= 2;
        b->p = (uint8_t *)realloc(b->p, c);
        b->cap = c;
    }
}
static inline void so_bytes_push(so_bytes *b, const void *src, size_t n) {
    so_bytes_reserve(b, n);
    memcpy(b->p + b->n, src, n);
    b->n += n;
}
static inline void so_bytes_push1(so_bytes *b, uint8_t v) { so_bytes_push(b, &v, 1); }
static inline void so_u32s_push(so_u32s *a, uint32_t v) {
    if (a->n == a->cap) {
        a->cap = a->cap ? a->cap * 2 : 8;
        a->p = (uint32_t *)realloc(a->p, a->cap * sizeof(uint32_t));
    }
    a->p[a->n++] = v;
}
static inline void so_u64s_push(so_u64s *a, uint64_t v) {
    if (a->n == a->cap) {
        a->cap = a->cap ? a->cap * 2 : 8;
        a->p = (uint64_t *)realloc(a->p, a->cap * sizeof(uint64_t));
    }
    a->p[a->n++] = v;
}

/* ---------------- text ---------------- */
/* Go `for _, r := range s` step: returns width, stores rune (invalid byte -> U+FFFD, width 1) */
int so_utf8_decode(const uint8_t *s, size_t len, uint32_t *rune);
int so_utf8_encode(uint32_t rune, uint8_t out[4]);
uint32_t so_rune_lower(uint32_t r);
void so_lower_into(const uint8_t *s, size_t len, so_bytes *out);

/* pkg/alphabet */
typedef struct {
    int kind; /* 0 sequential [lo,hi]; 1 russian (sequential + U+0451 -> U+0435); 2 simple set */
    uint32_t lo, hi;
    uint32_t *set;
    size_t n_set;
} so_alpha_part;
typedef struct {
    so_alpha_part *parts;
    int n_parts;
} so_alphabet;
void so_alphabet_init(so_alphabet *a, const char *const *desc, int n);
void so_alphabet_free(so_alphabet *a);
int so_alphabet_contains(const so_alphabet *a, uint32_t r);

/* token list: bytes back to back + offsets */
typedef struct {
    so_bytes bytes;
    so_u32s off; /* ntok+1 entries once non-empty */
} so_tokens;
static inline size_t so_tokens_count(const so_tokens *t) { return t->off.n ? t->off.n - 1 : 0; }
void so_tokens_reset(so_tokens *t);
void so_tokens_free(so_tokens *t);
void so_ngram_tokens(const uint8_t *text, size_t len, int n, so_tokens *out);

/* ---------------- codec input cursor (pkg/store/byte_input.go) ---------------- */
typedef struct { const uint8_t *buf; int64_t len, i; } so_input;
int so_read_vu32(so_input *in, uint32_t *v); /* 0 ok, <0 error */
int so_read_u16(so_input *in, uint16_t *v);

/* ---------------- merger.ListIterator (pkg/merger/list_iterator.go:14-26) ---------------- */
#define SO_IT_OK 0
#define SO_IT_NOT_DEREF 1 /* ErrIteratorIsNotDereferencable */
#define SO_IT_ERR (-1)

typedef struct so_iter so_iter;
typedef struct {
    int (*get)(so_iter *, uint32_t *);
    int (*has_next)(so_iter *);
    int (*next)(so_iter *, uint32_t *);
    int (*lower_bound)(so_iter *, uint32_t, uint32_t *);
    int (*len)(so_iter *);
} so_iter_vt;

struct so_iter {
    const so_iter_vt *vt;
    /* slice iterator */
    const uint32_t *slice;
    int index, size;
    /* encoded iterators (pkg/index/posting_list.go, skipping_posting_list.go) */
    so_input in;
    uint32_t current, current_skip_value;
    int next_skip_position, gap, is_last_block;
};

void so_iter_init_slice(so_iter *it, const uint32_t *ids, int n);
int so_iter_init_vb(so_iter *it, const uint8_t *buf, int64_t len, int list_size);
int so_iter_init_skipping(so_iter *it, const uint8_t *buf, int64_t len, int list_size, int gap);

/* merger.Collector: return 0 to continue, 1 for ErrCollectionTerminated, <0 error */
typedef int (*so_collect_fn)(void *ctx, uint64_t candidate);

int so_merger_merge(int algo, so_iter **rid, int n, int threshold, so_collect_fn collect, void *ctx);
int so_merger_intersect(so_iter **rid, int n, so_collect_fn collect, void *ctx);

#define SO_MAX_OVERLAP 0xFFFFu
static inline uint64_t so_cand(uint32_t pos, uint32_t overlap) { return ((uint64_t)pos << 32) | overlap; }
static inline uint32_t so_cand_pos(uint64_t c) { return (uint32_t)(c >> 32); }
static inline int so_cand_overlap(uint64_t c) { return (int)(uint32_t)(c & 0xFFFFFFFFu); }

/* ---------------- index ---------------- */
typedef struct {
    uint32_t segment;
    uint32_t term_off, term_len; /* into so_index.term_bytes */
    so_u32s ids;                 /* as appended by Writer.AddDocument (duplicates kept) */
    uint8_t *enc;                /* encoded bytes after commit */
    int64_t enc_len;
    int codec; /* 0 VB, 1 skipping(64), 2 "bitmap" class (>256), kept as a sorted slice */
} so_list;

typedef struct {
    int64_t *slots; /* open addressing: index into lists[], -1 empty */
    size_t cap, used;
} so_segmap;

struct so_index {
    int n;
    so_bytes wrap_start, wrap_end, pad;
    so_alphabet alphabet;
    so_list *lists;
    size_t n_lists, cap_lists;
    so_bytes term_bytes;
    so_segmap *segs; /* one map per segment (cardinality) */
    size_t n_segs;   /* == len(iw.indices), pkg/index/indexer_writer.go:69-73 */
    uint32_t n_docs;
    int committed;
};

so_list *so_index_find(const so_index *ix, uint32_t segment, const uint8_t *term, uint32_t term_len);
void so_tokenize_into(const so_index *ix, const uint8_t *text, size_t len, so_tokens *out, so_bytes *scratch);
void so_tokenize_mode(const so_index *ix, const uint8_t *text, size_t len, so_tokens *out, so_bytes *scratch, int tail_wrap);

#endif
