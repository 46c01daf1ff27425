/*
 * so_text.c — tokenizer chain of the reference, restated for the CPU oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows:
 *   pkg/suggest/tokenizer.go:9-20            composition wrap(filter(ngram, normalise))
 *   pkg/analysis/wrap_tokenizer.go:18-20     start + text + end
 *   pkg/analysis/filter_tokenizer.go:20-27   strings.ToLower, strings.Trim(text, " ")
 *   pkg/analysis/ngram_tokenizer.go:17-54    rune windows, byte-length early-out, appendUnique
 *   pkg/analysis/normalizer.go:21-37         rune not in alphabet -> pad
 *   pkg/alphabet (all files)                      english / russian (ё) / numbers / literal sets, composite = OR
 */
#include "so_internal.h"

/* ---- UTF-8, with Go's utf8.DecodeRune acceptance rules ---- */
int so_utf8_decode(const uint8_t *s, size_t len, uint32_t *rune) {
    uint8_t b0 = s[0];
    if (b0 < 0x80) { *rune = b0; return 1; }
    int need;
    uint8_t lo = 0x80, hi = 0xBF;
    uint32_t r;
    if (b0 >= 0xC2 && b0 <= 0xDF) { need = 1; r = b0 & 0x1F; }
    else if (b0 >= 0xE0 && b0 <= 0xEF) {
        need = 2; r = b0 & 0x0F;
        if (b0 == 0xE0) lo = 0xA0;
        if (b0 == 0xED) hi = 0x9F;
    } else if (b0 >= 0xF0 && b0 <= 0xF4) {
        need = 3; r = b0 & 0x07;
        if (b0 == 0xF0) lo = 0x90;
        if (b0 == 0xF4) hi = 0x8F;
    } else { *rune = SO_RUNE_ERROR; return 1; }
    if (len < (size_t)need + 1) { *rune = SO_RUNE_ERROR; return 1; }
    for (int i = 1; i <= need; i++) {
        uint8_t b = s[i];
        uint8_t l = (i == 1) ? lo : 0x80, h = (i == 1) ? hi : 0xBF;
        if (b < l || b > h) { *rune = SO_RUNE_ERROR; return 1; }
        r = (r << 6) | (b & 0x3F);
    }
    *rune = r;
    return need + 1;
}

int so_utf8_encode(uint32_t r, uint8_t out[4]) {
    if (r > 0x10FFFF || (r >= 0xD800 && r <= 0xDFFF)) r = SO_RUNE_ERROR; /* utf8.EncodeRune */
    if (r < 0x80) { out[0] = (uint8_t)r; return 1; }
    if (r < 0x800) { out[0] = 0xC0 | (r >> 6); out[1] = 0x80 | (r & 0x3F); return 2; }
    if (r < 0x10000) {
        out[0] = 0xE0 | (r >> 12); out[1] = 0x80 | ((r >> 6) & 0x3F); out[2] = 0x80 | (r & 0x3F);
        return 3;
    }
    out[0] = 0xF0 | (r >> 18); out[1] = 0x80 | ((r >> 12) & 0x3F); out[2] = 0x80 | ((r >> 6) & 0x3F);
    out[3] = 0x80 | (r & 0x3F);
    return 4;
}

/* ---- unicode.ToLower (simple case mapping) ---- */
typedef struct { uint32_t lo, hi; int32_t delta; uint32_t step; } so_case_range;
static const so_case_range so_lower_ranges[] = {
#include "unicode_lower.inc"
};

uint32_t so_rune_lower(uint32_t r) {
    if (r < 0x80) return (r >= 'A' && r <= 'Z') ? r + 32 : r;
    size_t lo = 0, hi = sizeof(so_lower_ranges) / sizeof(so_lower_ranges[0]);
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (so_lower_ranges[mid].hi < r) lo = mid + 1; else hi = mid;
    }
    if (lo < sizeof(so_lower_ranges) / sizeof(so_lower_ranges[0])) {
        const so_case_range *c = &so_lower_ranges[lo];
        if (r >= c->lo && r <= c->hi && (r - c->lo) % c->step == 0) return (uint32_t)((int32_t)r + c->delta);
    }
    return r;
}

/* strings.ToLower: pure-ASCII strings are lowered bytewise; anything else goes through strings.Map,
 * whose net effect is decode (invalid byte -> U+FFFD) / map / re-encode for every rune. */
void so_lower_into(const uint8_t *s, size_t len, so_bytes *out) {
    int ascii = 1;
    for (size_t i = 0; i < len; i++) if (s[i] >= 0x80) { ascii = 0; break; }
    if (ascii) {
        so_bytes_reserve(out, len);
        for (size_t i = 0; i < len; i++) {
            uint8_t c = s[i];
            out->p[out->n++] = (c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c;
        }
        return;
    }
    size_t i = 0;
    while (i < len) {
        uint32_t r;
        int w = so_utf8_decode(s + i, len - i, &r);
        uint8_t enc[4];
        int e = so_utf8_encode(so_rune_lower(r), enc);
        so_bytes_push(out, enc, (size_t)e);
        i += (size_t)w;
    }
}

int so_to_lower(const char *text, uint32_t len, char *out, uint32_t cap) {
    so_bytes b = {0};
    so_lower_into((const uint8_t *)text, len, &b);
    int n = (int)b.n;
    if (b.n > cap) n = -1; else memcpy(out, b.p, b.n);
    free(b.p);
    return n;
}

/* ---- alphabet ---- */
void so_alphabet_init(so_alphabet *a, const char *const *desc, int n) {
    a->parts = (so_alpha_part *)calloc((size_t)(n > 0 ? n : 1), sizeof(so_alpha_part));
    a->n_parts = n;
    for (int i = 0; i < n; i++) {
        so_alpha_part *p = &a->parts[i];
        if (strcmp(desc[i], "english") == 0) { p->kind = 0; p->lo = 'a'; p->hi = 'z'; }         /* english_alphabet.go */
        else if (strcmp(desc[i], "numbers") == 0) { p->kind = 0; p->lo = '0'; p->hi = '9'; }    /* number_alphabet.go */
        else if (strcmp(desc[i], "russian") == 0) { p->kind = 1; p->lo = 0x430; p->hi = 0x44F; } /* russian_alphabet.go */
        else { /* simple_alphabet.go: []rune(symbols) */
            p->kind = 2;
            size_t len = strlen(desc[i]);
            p->set = (uint32_t *)malloc((len + 1) * sizeof(uint32_t));
            size_t k = 0;
            while (k < len) {
                uint32_t r;
                int w = so_utf8_decode((const uint8_t *)desc[i] + k, len - k, &r);
                p->set[p->n_set++] = r;
                k += (size_t)w;
            }
        }
    }
}

void so_alphabet_free(so_alphabet *a) {
    for (int i = 0; i < a->n_parts; i++) free(a->parts[i].set);
    free(a->parts);
}

int so_alphabet_contains(const so_alphabet *a, uint32_t r) {
    for (int i = 0; i < a->n_parts; i++) {
        const so_alpha_part *p = &a->parts[i];
        uint32_t c = r;
        if (p->kind == 1 && c == 0x451) c = 0x435; /* ё tested as е, russian_alphabet.go:16-22 */
        if (p->kind == 2) {
            for (size_t k = 0; k < p->n_set; k++) if (p->set[k] == c) return 1;
        } else if (c >= p->lo && c <= p->hi) return 1;
    }
    return 0;
}

/* ---- token lists ---- */
void so_tokens_reset(so_tokens *t) { t->bytes.n = 0; t->off.n = 0; }
void so_tokens_free(so_tokens *t) { free(t->bytes.p); free(t->off.p); memset(t, 0, sizeof(*t)); }

/* appendUnique, ngram_tokenizer.go:46-54 */
static void so_append_unique(so_tokens *t, const uint8_t *x, size_t n) {
    size_t cnt = so_tokens_count(t);
    for (size_t i = 0; i < cnt; i++) {
        size_t l = t->off.p[i + 1] - t->off.p[i];
        if (l == n && memcmp(t->bytes.p + t->off.p[i], x, n) == 0) return;
    }
    if (t->off.n == 0) so_u32s_push(&t->off, 0);
    so_bytes_push(&t->bytes, x, n);
    so_u32s_push(&t->off, (uint32_t)t->bytes.n);
}

/* nGramTokenizer.Tokenize, ngram_tokenizer.go:17-43, kept statement for statement */
void so_ngram_tokens(const uint8_t *text, size_t len, int n, so_tokens *out) {
    so_tokens_reset(out);
    if ((int64_t)len < (int64_t)n) return;
    size_t prev_indexes[SO_MAX_N] = {0};
    int i = 0;
    size_t index = 0;
    while (index < len) {
        uint32_t r;
        int w = so_utf8_decode(text + index, len - index, &r);
        i++;
        if (i > n) {
            size_t top = prev_indexes[(i - n) % n];
            so_append_unique(out, text + top, index - top);
        }
        prev_indexes[i % n] = index;
        index += (size_t)w;
    }
    size_t top = prev_indexes[(i + 1) % n];
    so_append_unique(out, text + top, len - top);
}

int so_ngram_tokenize(const char *text, uint32_t len, int n, char *out, uint32_t cap, uint32_t *tok_off,
                      uint32_t max_tok) {
    so_tokens t = {0};
    so_ngram_tokens((const uint8_t *)text, len, n, &t);
    size_t cnt = so_tokens_count(&t);
    int ret = (int)cnt;
    if (t.bytes.n > cap || cnt > max_tok) ret = -1;
    else {
        memcpy(out, t.bytes.p, t.bytes.n);
        for (size_t i = 0; i <= cnt; i++) tok_off[i] = cnt ? t.off.p[i] : 0;
    }
    so_tokens_free(&t);
    return ret;
}

/* wrap -> lower -> trim -> ngram -> normalise */
void so_tokenize_into(const so_index *ix, const uint8_t *text, size_t len, so_tokens *out, so_bytes *scratch) {
    so_tokenize_mode(ix, text, len, out, scratch, 1);
}

/* tail_wrap = 0: NewAutocompleteTokenizer, pkg/suggest/tokenizer.go:23-34 (no wrap symbol at the tail of the query) */
void so_tokenize_mode(const so_index *ix, const uint8_t *text, size_t len, so_tokens *out, so_bytes *scratch, int tail_wrap) {
    so_bytes wrapped = {0};
    so_bytes_push(&wrapped, ix->wrap_start.p, ix->wrap_start.n);
    so_bytes_push(&wrapped, text, len);
    if (tail_wrap) so_bytes_push(&wrapped, ix->wrap_end.p, ix->wrap_end.n);
    scratch->n = 0;
    so_lower_into(wrapped.p, wrapped.n, scratch);
    free(wrapped.p);
    size_t b = 0, e = scratch->n;
    while (b < e && scratch->p[b] == ' ') b++;     /* strings.Trim(text, " ") */
    while (e > b && scratch->p[e - 1] == ' ') e--;
    so_tokens raw = {0};
    so_ngram_tokens(scratch->p + b, e - b, ix->n, &raw);
    /* normalizeFilter.Filter: tokens are rewritten in place, duplicates after rewriting are kept */
    so_tokens_reset(out);
    size_t cnt = so_tokens_count(&raw);
    if (cnt) so_u32s_push(&out->off, 0);
    for (size_t t = 0; t < cnt; t++) {
        const uint8_t *tok = raw.bytes.p + raw.off.p[t];
        size_t tl = raw.off.p[t + 1] - raw.off.p[t], k = 0;
        while (k < tl) {
            uint32_t r;
            int w = so_utf8_decode(tok + k, tl - k, &r);
            if (so_alphabet_contains(&ix->alphabet, r)) {
                uint8_t enc[4];
                int el = so_utf8_encode(r, enc); /* res += string(r) */
                so_bytes_push(&out->bytes, enc, (size_t)el);
            } else {
                so_bytes_push(&out->bytes, ix->pad.p, ix->pad.n);
            }
            k += (size_t)w;
        }
        so_u32s_push(&out->off, (uint32_t)out->bytes.n);
    }
    so_tokens_free(&raw);
}

int so_tokenize(const so_index *ix, const char *text, uint32_t len, char *out, uint32_t cap, uint32_t *tok_off,
                uint32_t max_tok) {
    so_tokens t = {0};
    so_bytes scratch = {0};
    so_tokenize_into(ix, (const uint8_t *)text, len, &t, &scratch);
    size_t cnt = so_tokens_count(&t);
    int ret = (int)cnt;
    if (t.bytes.n > cap || cnt > max_tok) ret = -1;
    else {
        if (t.bytes.n) memcpy(out, t.bytes.p, t.bytes.n);
        for (size_t i = 0; i <= cnt; i++) tok_off[i] = cnt ? t.off.p[i] : 0;
    }
    so_tokens_free(&t);
    free(scratch.p);
    return ret;
}

int so_alphabet_has(const so_index *ix, uint32_t rune) { return so_alphabet_contains(&ix->alphabet, rune); }

int so_has_duplicate_tokens(const so_index *ix, const char *text, uint32_t len) {
    so_tokens t = {0};
    so_bytes scratch = {0};
    so_tokenize_into(ix, (const uint8_t *)text, len, &t, &scratch);
    size_t cnt = so_tokens_count(&t);
    int dup = 0;
    for (size_t i = 0; i < cnt && !dup; i++)
        for (size_t j = 0; j < i; j++) {
            size_t li = t.off.p[i + 1] - t.off.p[i], lj = t.off.p[j + 1] - t.off.p[j];
            if (li == lj && memcmp(t.bytes.p + t.off.p[i], t.bytes.p + t.off.p[j], li) == 0) { dup = 1; break; }
        }
    so_tokens_free(&t);
    free(scratch.p);
    return dup;
}
