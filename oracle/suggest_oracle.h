/*
 * suggest_oracle.h — CPU restatement ("oracle") of the suggest-go/suggest Suggest hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under suggest_b200/ may include, link or call this.  It is
 * used by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs as the checker and the reported CPU baseline.
 *
 * Parity status: PINNED against the reference's own golden vectors (tests/test_oracle_*.py):
 *   pkg/analysis/ngram_tokenizer_test.go:16-45, pkg/merger/list_merger_test.go:48-140,
 *   pkg/merger/list_intersector_test.go:14-46, pkg/suggest/topk_test.go:10-39,
 *   pkg/suggest/ngram_index_test.go:16-39, pkg/suggest/example_test.go:17-71,
 *   pkg/suggest/service_test.go:35-59, pkg/compression/compression_test.go:28-56,
 *   pkg/index/posting_list_test.go:39-132, and the bytes of pkg/suggest/testdata/db/cars.{hd,dl}.
 * Float scores are not asserted by any reference test (pkg/metric has none); they are pinned only
 * by restating pkg/metric/{jaccard,cosine,dice,overlap,exact}.go and pkg/suggest/scorer.go:30 operation for operation.
 *
 * The reference (Go) cannot be built in this image (no Go toolchain), so there is no oracle/_ref.
 */
#ifndef SUGGEST_ORACLE_H
#define SUGGEST_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* pkg/metric: jaccard.go, cosine.go, dice.go, overlap.go, exact.go */
enum { SO_JACCARD = 0, SO_COSINE = 1, SO_DICE = 2, SO_OVERLAP = 3, SO_EXACT = 4 };

/* pkg/merger: scan_count.go, cp_merge.go, merge_skip.go, divide_skip.go */
enum { SO_SCAN_COUNT = 0, SO_CP_MERGE = 1, SO_MERGE_SKIP = 2, SO_DIVIDE_SKIP = 3 };

/* search modes of so_suggest */
enum {
    SO_MODE_CANONICAL = 0, /* rule set of SURVEY.md §8(c): exact overlap, no dynamic alpha, each doc once */
    SO_MODE_FAITHFUL = 1   /* line-faithful: segment order, lazy codecs, CPMerge, per-segment queues, dynamic alpha */
};

typedef struct so_index so_index;

/* ---- index description (pkg/suggest/config.go:25-35) and build (pkg/suggest/indexer.go:14-45) ---- */
so_index *so_index_new(int ngram_size, const char *wrap_start, const char *wrap_end, const char *pad,
                       const char *const *alphabet, int n_alphabet);
/* docs: concatenated bytes; off[n_docs+1]; document id = position (pkg/dictionary/helpers.go:38-45) */
int so_index_add_docs(so_index *ix, const char *bytes, const uint64_t *off, uint32_t n_docs);
/* encode every list with the length-class codec (pkg/index/codec.go:39-51); required before FAITHFUL search */
int so_index_commit(so_index *ix);
void so_index_free(so_index *ix);
uint32_t so_index_segments(const so_index *ix);     /* InvertedIndexIndices.Size() */
uint64_t so_index_postings(const so_index *ix);     /* total stored postings (with in-list duplicates) */
uint64_t so_index_lists(const so_index *ix);        /* non-empty (segment, term) lists */
/* copy out one list; returns its length or -1 if absent.  out may be NULL to query the length */
int64_t so_index_get_list(const so_index *ix, uint32_t segment, const char *term, uint32_t term_len,
                          uint32_t *out, uint64_t cap);
/* iterate lists: fills term bytes/segment/length for list number i (arbitrary but stable order) */
int so_index_list_at(const so_index *ix, uint64_t i, uint32_t *segment, const char **term, uint32_t *term_len,
                     const uint32_t **ids, uint32_t *n_ids);

/* ---- tokenizer chain (pkg/suggest/tokenizer.go:9-20 and pkg/analysis) ---- */
/* plain n-gram tokenizer, pkg/analysis/ngram_tokenizer.go:17-43.  Tokens are written NUL-free,
 * back to back into out; tok_off[ntok+1].  Returns the token count or -1 when out/tok_off is too small. */
int so_ngram_tokenize(const char *text, uint32_t len, int n, char *out, uint32_t cap, uint32_t *tok_off,
                      uint32_t max_tok);
/* full chain wrap -> lower -> trim -> n-gram -> normalise */
int so_tokenize(const so_index *ix, const char *text, uint32_t len, char *out, uint32_t cap, uint32_t *tok_off,
                uint32_t max_tok);
/* alphabet membership, pkg/alphabet/composite_alphabet.go:35-45 */
int so_alphabet_has(const so_index *ix, uint32_t rune);
/* strings.ToLower restated (Go semantics incl. invalid UTF-8 -> U+FFFD); returns output length */
int so_to_lower(const char *text, uint32_t len, char *out, uint32_t cap);

/* ---- metric (pkg/metric) ---- */
int so_metric_min_y(int metric, double alpha, int size);
int so_metric_max_y(int metric, double alpha, int size);
int so_metric_threshold(int metric, double alpha, int size_a, int size_b);
double so_metric_distance(int metric, int inter, int size_a, int size_b);
double so_score(int metric, int inter, int size_a, int size_b); /* scorer.go:29-31 */

/* ---- mergers over plain sorted slices (pkg/merger) ---- */
/* lists: flat ids + off[n_lists+1].  out receives MergeCandidates (pos<<32|overlap) in emission order.
 * Returns the number emitted, or -1 on overflow of cap. */
int64_t so_merge(int algo, const uint32_t *ids, const uint32_t *off, uint32_t n_lists, int threshold,
                 uint64_t *out, uint64_t cap);
int64_t so_intersect(const uint32_t *ids, const uint32_t *off, uint32_t n_lists, uint64_t *out, uint64_t cap);

/* ---- codecs (pkg/compression) and posting-list iterators (pkg/index) ---- */
enum { SO_CODEC_VB = 0, SO_CODEC_SKIPPING = 1, SO_CODEC_BINARY = 2 };
int64_t so_encode(int codec, int gap, const uint32_t *list, uint32_t n, uint8_t *out, uint64_t cap);
int64_t so_decode(int codec, int gap, const uint8_t *in, uint64_t in_len, uint32_t *out, uint32_t n);
/* posting_list_test.go: encode -> Init -> LowerBound(to) -> drain.  kind: SO_CODEC_VB or SO_CODEC_SKIPPING.
 * *lb = value returned by LowerBound, *err = 1 when it reported "not dereferencable".  Returns tail length. */
int so_posting_lower_bound_tail(int kind, int gap, const uint32_t *list, uint32_t n, uint32_t to, uint32_t *lb,
                                int *err, uint32_t *tail, uint32_t cap);

/* ---- top-k queue (pkg/suggest/topk.go) ---- */
/* adds (ids[i], scores[i]) in order to a queue of size k, returns GetCandidates() */
int so_topk(const uint32_t *ids, const double *scores, uint32_t n, uint32_t k, uint32_t *out_ids,
            double *out_scores, double *lowest_score);

/* ---- the path itself: nGramSuggester.Suggest with a FuzzyCollectorManager(k) ---- */
/* returns the number of candidates (<= k), ordered score desc / id asc; -1 on error */
int so_suggest(const so_index *ix, const char *query, uint32_t qlen, int metric, double alpha, uint32_t k,
               int mode, int merger_algo, uint32_t *out_ids, double *out_scores);
/* batch over n_threads pthreads (one query per thread at a time); out stride k; counts[n_q] */
int so_suggest_batch(const so_index *ix, const char *q_bytes, const uint64_t *q_off, uint32_t n_q, int metric,
                     double alpha, uint32_t k, int mode, int merger_algo, int n_threads, uint32_t *out_ids,
                     double *out_scores, uint32_t *out_counts);
/* nGramAutocomplete.Autocomplete with a FirstKCollectorManager(limit) (pkg/suggest/autocomplete.go:40-77,
 * collector.go:48-115): documents of every segment >= len(tokens) that hold all query tokens (tokenizer without the
 * tail wrap), the `limit` lowest ids, score = -id.  Returns the count, -1 on error. */
int so_autocomplete(const so_index *ix, const char *query, uint32_t qlen, uint32_t limit, uint32_t *out_ids,
                    double *out_scores);
/* SURVEY.md §8(d) algorithmic-bytes ingredients for one query */
int so_query_stats(const so_index *ix, const char *query, uint32_t qlen, int metric, double alpha,
                   uint64_t *postings, uint64_t *lists, uint32_t *segments, uint32_t *size_a);
/* does this text's token list contain duplicates after normalisation? (SURVEY §8c rule 5b whitelist) */
int so_has_duplicate_tokens(const so_index *ix, const char *text, uint32_t len);

#ifdef __cplusplus
}
#endif
#endif
