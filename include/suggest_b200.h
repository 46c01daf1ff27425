/*
 * suggest_b200.h — C ABI of libsuggest_b200.so, the B200 (sm_100a) implementation of the
 * suggest-go/suggest `Service.Suggest -> NGramIndex.Suggest` hot path.
 *
 * The reference has no FFI of its own (pure Go, CGO_ENABLED=0); its seam is the Go interface pair
 *   suggest.Builder{ Build() (NGramIndex, error) }            pkg/suggest/ngram_index_builder.go:14-17
 *   suggest.Suggester{ Suggest(query, similarity, metric, factory) ([]Candidate, error) }
 *                                                              pkg/suggest/suggester.go:17-20
 * A cgo shim implements those two interfaces on top of the entry points below,
 * so `Service.AddIndex(name, dict, builder)` (pkg/suggest/service.go:78-91) accepts the GPU index
 * unchanged.  Every entry point names the reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every input and output buffer and the library
 *     never keeps a pointer past the call (cgo pointer rules);
 *   - return 0 on success, a negative sg_status otherwise; sg_last_error() gives the thread-local
 *     message (Go side: errors.New(C.GoString(C.sg_last_error())));
 *   - an sg_index is immutable after creation and may be searched from any number of host threads
 *     at once (pkg/suggest/service_test.go:19-80 exercises exactly that); sg_index_free waits for
 *     calls in flight;
 *   - results are fixed stride: row q of out_ids/out_scores holds out_counts[q] <= k candidates
 *     ordered (score desc, id asc) as pkg/suggest/collector.go:20-26 + topk.go:127-147 emit them.
 */
#ifndef SUGGEST_B200_H
#define SUGGEST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    SG_OK = 0,
    SG_ERR_INVALID = -1,      /* bad argument (NewSearchConfig rules: k >= 1, 0 < similarity <= 1) */
    SG_ERR_UNSUPPORTED = -2,  /* description outside what the device tokenizer encodes (see sg_index_build) */
    SG_ERR_CUDA = -3,         /* CUDA runtime failure, message holds cudaGetErrorString */
    SG_ERR_NOMEM = -4,
    SG_ERR_QUERY_TOO_LONG = -5, /* device-buffer entry points: some query has more than SG_MAX_QUERY_TOKENS n-grams, its count is
                                   SG_COUNT_UNSUPPORTED.  Host-buffer entry points answer such queries (up to 65535 n-grams, topK <= 1024) */
    SG_ERR_IO = -6,
    SG_ERR_FORMAT = -7        /* on-disk index is not "v5.1" or is corrupt */
} sg_status;

/* pkg/metric: JaccardMetric, CosineMetric, DiceMetric, OverlapMetric, ExactMetric */
typedef enum { SG_JACCARD = 0, SG_COSINE = 1, SG_DICE = 2, SG_OVERLAP = 3, SG_EXACT = 4 } sg_metric;

#define SG_MAX_NGRAM 8               /* pkg/analysis/ngram_tokenizer.go:3 (maxN) */
#define SG_MAX_QUERY_TOKENS 128      /* n-grams per query the batched kernels take; longer queries (the reference saturates at 0xFFFF,
                                        pkg/merger/list_merger.go:9) are answered by sg_search_batch / sg_autocomplete_batch / sg_suggest_one
                                        on a slower path of their own (host tokenization + sg_long_query_kernel) */
#define SG_MAX_TOPK 16384            /* sg_search_batch / _device; above 1024 the per-warp top-k lives in HBM instead of shared memory.
                                        The shard merges, sg_candidates' replay aside, the batcher and sg_predict_batch take k <= SG_MAX_TOPK_SHARED */
#define SG_MAX_TOPK_SHARED 1024
#define SG_COUNT_UNSUPPORTED 0xFFFFFFFFu

/* suggest.IndexDescription, pkg/suggest/config.go:25-35 (tokenizer-relevant fields) */
typedef struct {
    int32_t ngram_size;            /* NGramSize, 1..8 */
    const char *wrap_start;        /* Wrap[0], UTF-8 */
    const char *wrap_end;          /* Wrap[1] */
    const char *pad;               /* Pad: must decode to exactly one rune */
    const char *const *alphabet;   /* Alphabet: "english" | "russian" | "numbers" | literal characters */
    int32_t n_alphabet;
    int32_t device;                /* CUDA device ordinal that will hold the index */
} sg_config;

typedef struct sg_index sg_index;

typedef struct {
    uint32_t n_docs;
    uint32_t n_segments;     /* InvertedIndexIndices.Size() = largest cardinality + 1 */
    uint32_t n_terms;        /* distinct n-grams */
    uint64_t n_lists;        /* non-empty (segment, term) posting lists */
    uint64_t n_postings;
    uint64_t device_bytes;   /* HBM held by the handle */
    uint32_t id_base;
    int32_t device;
} sg_index_info;

/*
 * Build the index on the GPU from the dictionary text.
 * Replaces suggest.NewRAMBuilder(dict, description).Build()  (pkg/suggest/ngram_index_builder.go:27-83),
 * i.e. suggest.Index (pkg/suggest/indexer.go:14-45) + index.Writer.AddDocument/Commit
 * (pkg/index/indexer_writer.go:66-145) + index.Reader.Read (pkg/index/index_reader.go:29-120), with
 * the posting lists kept decoded in HBM as CSR instead of VB/skipping/roaring bytes.
 * Document i gets id id_base + i (pkg/dictionary/helpers.go:38-45: id = line number); id_base > 0
 * is for record-id-range shards.
 * The build itself runs on the device (tokenise, sort, CSR, bitmaps: sg_gpubuild.cu) unless a document has more than 128
 * n-grams or the text exceeds 4 GB, in which case the host build is used; SG_BUILD=host|gpu forces one of them.
 */
int sg_index_build(const sg_config *cfg, const char *doc_bytes, const uint64_t *doc_off, uint32_t n_docs,
                   uint32_t id_base, sg_index **out);

/*
 * Build from posting lists the caller already decoded (for a Go shim that read `.hd/.dl` itself).
 * list i belongs to cardinality segment list_segment[i] and term list_term_off[i]..list_term_off[i+1]
 * of term_bytes, and holds ids[list_off[i]..list_off[i+1]) in ascending order.
 * Replaces index.Reader.Read (pkg/index/index_reader.go:29-120).
 */
int sg_index_from_lists(const sg_config *cfg, uint32_t n_segments, uint64_t n_lists, const uint32_t *list_segment,
                        const char *term_bytes, const uint64_t *list_term_off, const uint32_t *ids,
                        const uint64_t *list_off, sg_index **out);

/*
 * Open an index written by the reference's `suggest indexer` (header `<name>.hd`, gob; lists
 * `<name>.dl`, VB / skipping(64) / roaring by length class).  Replaces suggest.NewFSBuilder(...).Build()
 * (pkg/suggest/ngram_index_builder.go:38-83, pkg/index/index_reader.go:29-120, pkg/compression).
 */
int sg_index_open_disk(const sg_config *cfg, const char *hd_path, const char *dl_path, sg_index **out);

void sg_index_free(sg_index *ix);
int sg_index_get_info(const sg_index *ix, sg_index_info *info);

/*
 * How the handle lays the index out in HBM.  Documents are renumbered ("slots") by cardinality segment
 * (pkg/index/indexer_writer.go:66-86 files a document under len(tokens)); next to the decoded posting lists every
 * term owns one row of bits, one bit per bucket of 2^bucket_shift consecutive slots, which is what the bitmap engine
 * reads.  engine 0 = scan-count kernel over the posting lists (used when the bitmaps exceed their memory budget or
 * SG_ENGINE=scancount), engine 1 = bitmap kernel.
 */
typedef struct {
    uint32_t n_slots;        /* document slots, alignment holes between segments included */
    uint32_t bucket_shift;
    uint32_t row_words;      /* 32-bit words per bitmap row; 0: built without bitmaps */
    uint32_t engine;
    uint32_t built_on_device; /* 1: sg_index_build ran the device build (sg_gpubuild.cu); 0: host build */
    uint32_t pipeline;       /* engine 1 only: 1 = Suggest runs the count -> resolve pipeline (sg_count_kernel, sg_resolve_kernel over the
                                exact level, sg_fine.cu); 0 = sg_bitmap_search_kernel alone (SG_PIPELINE=classic, or the level did not fit) */
    uint64_t bitmap_bytes;
} sg_index_layout;
int sg_index_get_layout(const sg_index *ix, sg_index_layout *layout);

/*
 * Batched NGramIndex.Suggest with a FuzzyCollectorManager(k): for every query
 *   tokenise (pkg/suggest/tokenizer.go:9-20, pkg/analysis) -> segment window and thresholds
 *   (pkg/suggest/suggester.go:46-131, pkg/metric) -> posting fetch + T-occurrence count
 *   (pkg/index/searcher.go:28-78, pkg/merger) -> score (pkg/suggest/scorer.go:29-31) -> top-k
 *   (pkg/suggest/collector.go:117-191, topk.go).
 * HOST buffers; the call copies queries to the device, runs the kernels and copies results back.
 * q_off has n_q + 1 entries.  out_ids / out_scores hold n_q * k entries, out_counts n_q.
 * Row q holds out_counts[q] <= k candidates in GetCandidates order (score descending, id ascending); the entries of a
 * row at and behind out_counts[q] are unspecified (the Go shim slices a row by its count).
 * If all three output buffers are page-locked memory the device can address (sg_pinned_alloc, cudaHostAlloc,
 * cudaHostRegister) the search kernel stores the candidates straight into them while it runs: no staging copy in HBM,
 * no device-to-host copy behind the kernel, and only the valid entries cross PCIe.  Pageable buffers (plain Go or
 * malloc memory) go through staging and cudaMemcpyAsync.  SG_DIRECT_OUT=0 forces the staged path.
 * A single Suggest call is a batch of one.
 */
int sg_search_batch(sg_index *ix, const char *q_bytes, const uint32_t *q_off, uint32_t n_q, int metric, double alpha,
                    uint32_t k, uint32_t *out_ids, double *out_scores, uint32_t *out_counts);

/*
 * sg_search_batch with the rows as suggest.Candidate lays them out in memory (pkg/suggest/collector.go:12-17:
 * {Key index.Position (uint32), Score float64} = 16 bytes with the padding): out_rows[q * k + i] for i < out_counts[q].
 * A Go caller can view page-locked rows as []Candidate without converting them, and with page-locked rows the kernels
 * store one 16-byte entry per candidate instead of an id and a score into separate arrays - half the PCIe writes, which is
 * what bounds the direct result path.  Needs an index with bitmaps (engine 1).
 */
typedef struct {
    uint32_t key;
    uint32_t reserved;   /* written as 0 */
    double score;
} sg_candidate;
int sg_search_batch_candidates(sg_index *ix, const char *q_bytes, const uint32_t *q_off, uint32_t n_q, int metric, double alpha,
                               uint32_t k, sg_candidate *out_rows, uint32_t *out_counts);

/*
 * The same call without waiting for it: several batches in flight from ONE host thread.  The reference overlaps its
 * requests with goroutines (internal/suggest/api/suggest_handler.go:42-76); a host that drives the library from a single
 * thread gets the same overlap here - one call's copies and host-side work run under another call's kernels.
 * submit returns at once with a ticket; a few worker threads of the library (SG_SUBMIT_WORKERS, default 3) run
 * sg_search_batch_candidates itself; sg_ticket_wait blocks until the call has returned, hands back its status (and its
 * message through sg_last_error) and frees the ticket.  Every buffer of the call belongs to the library from submit to
 * wait.  Every ticket must be waited for exactly once; sg_index_free serves what was submitted before it frees the index.
 */
typedef struct sg_ticket sg_ticket;
int sg_search_batch_candidates_submit(sg_index *ix, const char *q_bytes, const uint32_t *q_off, uint32_t n_q, int metric,
                                      double alpha, uint32_t k, sg_candidate *out_rows, uint32_t *out_counts, sg_ticket **ticket);
int sg_ticket_wait(sg_ticket *ticket);

/*
 * Batched NGramIndex.Autocomplete with a FirstKCollectorManager(limit)  (pkg/suggest/autocomplete.go:40-77,
 * collector.go:48-115): the query is tokenised without the tail wrap (pkg/suggest/tokenizer.go:23-34), a candidate must
 * hold every query n-gram (threshold = len(tokens), segments len(tokens)..Size()-1) and the `limit` lowest ids win,
 * returned ascending with score = -id as the reference's queue scores them.  Same buffers as sg_search_batch.
 */
int sg_autocomplete_batch(sg_index *ix, const char *q_bytes, const uint32_t *q_off, uint32_t n_q, uint32_t limit,
                          uint32_t *out_ids, double *out_scores, uint32_t *out_counts);

/*
 * Every candidate of the T-occurrence count, for callers whose CollectorManager is not a FuzzyCollectorManager /
 * FirstKCollectorManager or whose metric.Metric is not one of the five built-ins: what searcher.Search + the mergers hand
 * to Collector.Collect (pkg/index/searcher.go:28-78, pkg/merger/collector.go:10-13, MergeCandidate = position + overlap,
 * pkg/merger/list_merger.go:33-48) over every admissible segment of nGramSuggester.Suggest's window
 * (pkg/suggest/suggester.go:53-78).  The Go shim replays them through factory().Create() / SetScorer / Collect /
 * manager.Collect on the host, segment by segment in the reference's feed order (suggester.go:110-118).
 *   thresholds == NULL: window and thresholds of the built-in `metric` at `alpha`, computed on the device as in
 *                       sg_search_batch;
 *   thresholds != NULL: [SG_MAX_QUERY_TOKENS + 1][n_segments] bytes tabulated by the caller from its own metric.Metric:
 *                       thresholds[a * n_segments + B] = Threshold(alpha, a, B) for MinY(alpha, a) <= B <= MaxY(alpha, a),
 *                       0 elsewhere (values above 255 as 255: never admissible, a <= 128).  `metric` / `alpha` are ignored.
 *                       The library applies suggester.go:76 (T == 0, T > sizeB, T > sizeA: skip) and skips empty segments.
 * Output: one entry per candidate in no particular order: out_query (number of the query in the batch), out_ids
 * (document id), out_overlap (exact overlap count, rule 5 of SURVEY.md 8c), out_segment (sizeB).  *out_total is the
 * number found; only the first min(cap, *out_total) are written - call again with a larger cap if it was exceeded.
 * out_size_a[q] = len(tokens) of query q (the sizeA of Distance / Threshold).  HOST buffers.  Bitmap-engine indexes only
 * (SG_ERR_UNSUPPORTED otherwise).
 */
int sg_candidates_batch(sg_index *ix, const char *q_bytes, const uint32_t *q_off, uint32_t n_q, int metric, double alpha,
                        const uint8_t *thresholds, uint64_t cap, uint32_t *out_query, uint32_t *out_ids, uint32_t *out_overlap,
                        uint32_t *out_segment, uint64_t *out_total, uint32_t *out_size_a);

/*
 * Same, every buffer already resident on the index's device; enqueued on `stream` (a cudaStream_t,
 * NULL = default stream) without synchronising.  Queries must already be lower-cased if they hold
 * non-ASCII bytes (sg_search_batch does that on the host, strings.ToLower semantics).
 * d_stats may be NULL; otherwise (16-byte aligned, 4 uint32 per query) it receives {admissible postings, admissible
 * lists} of SURVEY.md section 8(d), the 32-bit words the engine itself reads for the count, and 0.
 */
int sg_search_batch_device(sg_index *ix, const char *d_q_bytes, const uint32_t *d_q_off, uint32_t n_q, int metric,
                           double alpha, uint32_t k, uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts,
                           uint32_t *d_stats, void *stream);

/*
 * Measurement aid: one sg_search_batch_device launch with CUDA events between its kernels on `stream`; synchronises.
 * ms_out receives one duration per kernel (at most 8), names_out their comma-separated names.  Returns the number of
 * kernels.  bench.py uses it for the per-kernel roofline; it is not part of the reference-facing path.
 */
int sg_search_stage_times(sg_index *ix, const char *d_q_bytes, const uint32_t *d_q_off, uint32_t n_q, int metric, double alpha,
                          uint32_t k, uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts, void *stream,
                          float *ms_out, char *names_out, uint32_t names_cap);

/*
 * sg_autocomplete_batch with every buffer resident on the index's device (as sg_search_batch_device is to
 * sg_search_batch; d_stats as there, for the admissible lists of NGramIndex.Autocomplete: every segment from len(tokens)
 * up), and the measurement aid for it.
 */
int sg_autocomplete_batch_device(sg_index *ix, const char *d_q_bytes, const uint32_t *d_q_off, uint32_t n_q, uint32_t limit,
                                 uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts, uint32_t *d_stats, void *stream);
int sg_autocomplete_stage_times(sg_index *ix, const char *d_q_bytes, const uint32_t *d_q_off, uint32_t n_q, uint32_t limit,
                                uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts, void *stream, float *ms_out,
                                char *names_out, uint32_t names_cap);

/*
 * Cross-shard reduce for record-id-range shards: for every query pick the k best of n_parts
 * per-shard results (layout [part][query][k] as an all-gather of sg_search_batch_device outputs
 * delivers them) under the same (score desc, id asc) order.  Device buffers.
 * The reference has no counterpart (single process); the order is FuzzyCollectorManager.Collect's
 * queue merge, pkg/suggest/collector.go:165-178.
 */
int sg_merge_topk_device(int device, uint32_t n_parts, uint32_t n_q, uint32_t k, const uint32_t *d_part_ids,
                         const double *d_part_scores, const uint32_t *d_part_counts, uint32_t *d_out_ids,
                         double *d_out_scores, uint32_t *d_out_counts, void *stream);

/*
 * The same pair with one packed block per shard, so that the exchange is a single all-gather:
 * block = [scores: n_q * k doubles | ids: n_q * k uint32 | counts: n_q uint32], padded to sg_packed_rows_bytes(n_q, k).
 * sg_search_batch_packed_device writes this shard's block; after an all-gather of the blocks ([part][block]),
 * sg_merge_topk_packed_device picks the k best per query.
 */
uint64_t sg_packed_rows_bytes(uint32_t n_q, uint32_t k);
int sg_search_batch_packed_device(sg_index *ix, const char *d_q_bytes, const uint32_t *d_q_off, uint32_t n_q, int metric,
                                  double alpha, uint32_t k, void *d_packed, void *stream);
int sg_merge_topk_packed_device(int device, uint32_t n_parts, uint32_t n_q, uint32_t k, const void *d_parts, uint32_t *d_out_ids,
                                double *d_out_scores, uint32_t *d_out_counts, void *stream);

/*
 * Record-id-range shards of one dictionary over the GPUs of a box, driven by ONE host process (the Go service cannot be
 * one process per GPU).  No reference counterpart (single process, single index); SURVEY.md 8(e), BASELINE.json config #4.
 * Shard s indexes documents [n_docs * s / n_shards, n_docs * (s + 1) / n_shards) on CUDA device devices[s] with id_base =
 * its first document, so returned ids are the dictionary's and the (score desc, id asc) order of pkg/suggest/collector.go:20-26
 * holds across shards.  A query's result is the k best of the per-shard top-k lists (FuzzyCollectorManager.Collect merging
 * per-segment queues, collector.go:165-178, is the same reduction).
 * sg_sharded_search_batch: HOST buffers as sg_search_batch.  The queries go to the first GPU once and from there to the
 * others over NVLink; every shard searches on its own stream; the merge kernel on the first GPU reads the per-shard rows
 * straight from the other GPUs' HBM (peer access; `peer_reads` of sg_sharded_get_info), or from peer copies of the blocks
 * when peer access is not available (SG_SHARD_GATHER_COPY=1 forces that).  One call at a time per handle.
 * devices may name the same GPU more than once (shards then share it).
 */
typedef struct sg_sharded sg_sharded;
int sg_sharded_build(const sg_config *cfg, const char *doc_bytes, const uint64_t *doc_off, uint32_t n_docs, const int32_t *devices,
                     uint32_t n_shards, sg_sharded **out);
int sg_sharded_search_batch(sg_sharded *sx, const char *q_bytes, const uint32_t *q_off, uint32_t n_q, int metric, double alpha,
                            uint32_t k, uint32_t *out_ids, double *out_scores, uint32_t *out_counts);
int sg_sharded_get_info(const sg_sharded *sx, uint32_t *n_shards, uint32_t *n_docs, int32_t *peer_reads);
sg_index *sg_sharded_shard(const sg_sharded *sx, uint32_t shard);   /* borrowed: sg_index_get_info / _layout of one shard */
void sg_sharded_free(sg_sharded *sx);

/*
 * The same shards with ONE PROCESS PER GPU (torch.distributed / MPI ranks): the exchange of the per-shard top-k and the merge
 * as one kernel over NVLink peer memory (sg_exchange.cu).  No reference counterpart; replaces "NCCL all-gather + merge" of
 * BASELINE.json config #4 (kept as SG_SHARD_EXCHANGE=nccl in suggest_b200/sharding.py).
 * Every rank creates an exchange (a region of its HBM: flags, its shard's rows, the merged rows), publishes the region's
 * CUDA IPC handle (SG_EXCHANGE_HANDLE_BYTES) to the others by whatever transport the host has, and connects with the
 * handles of all ranks in rank order.  sg_exchange_search is collective: every rank calls it with the same batch in device
 * memory, in the same order.  Rank r searches its shard, then merges queries [n_q * r / world, n_q * (r + 1) / world) from
 * all shards' rows (peer loads, valid entries only) and stores the winners into the merged rows of every rank (peer stores);
 * two flag barriers inside the kernel order the steps.  On return (stream order) d_out_* (optional) or sg_exchange_result
 * hold the k best of the whole dictionary for every query, on every rank; entries behind a row's count are unspecified.
 * max_queries / max_k must be the same on every rank.  sg_exchange_status: SG_OK, or an error if a barrier timed out
 * (a rank did not arrive within ~4 s).
 */
#define SG_EXCHANGE_HANDLE_BYTES 64
typedef struct sg_exchange sg_exchange;
int sg_exchange_create(int device, uint32_t rank, uint32_t world, uint32_t max_queries, uint32_t max_k, sg_exchange **out);
int sg_exchange_handle(sg_exchange *ex, void *handle_out);
int sg_exchange_connect(sg_exchange *ex, const void *handles);
int sg_exchange_search(sg_exchange *ex, sg_index *ix, const char *d_q_bytes, const uint32_t *d_q_off, uint32_t n_q, int metric,
                       double alpha, uint32_t k, uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts, void *stream);
int sg_exchange_result(sg_exchange *ex, uint32_t n_q, uint32_t k, const uint32_t **d_ids, const double **d_scores,
                       const uint32_t **d_counts);
int sg_exchange_status(sg_exchange *ex, void *stream);
void sg_exchange_free(sg_exchange *ex);

/*
 * ---- single-query callers: a micro-batcher in front of sg_search_batch ----
 * Replaces nothing of the reference and serves its calling pattern: Service.Suggest is called with ONE query per goroutine
 * (internal/suggest/api/suggest_handler.go:42-76, cmd/suggest/cmd/eval.go:60, pkg/spellchecker/spellchecker.go:67).
 * sg_suggest_one blocks the calling thread until its row is there; any number of host threads may call it at once.
 * Their queries are coalesced by worker threads (two: two batches in flight; SG_BATCHER_WORKERS) into sg_search_batch
 * calls over page-locked buffers the batcher owns: a batch closes when it holds max_batch queries, when its oldest query
 * has waited max_wait_us, or - under load - as soon as a worker's previous batch has returned.  Queries of one batch share (metric, similarity); k is per query (<= max_k).
 * out_ids / out_scores: room for k entries; *out_count receives the number of candidates ((score desc, id asc) order).
 * Errors are per query (SG_ERR_QUERY_TOO_LONG for one query does not fail its batch mates).
 * sg_batcher_free serves what is queued, then stops the workers; the index must outlive the batcher.
 */
typedef struct sg_batcher sg_batcher;
typedef struct {
    uint64_t batches;        /* sg_search_batch calls so far */
    uint64_t queries;        /* queries served so far */
    uint32_t largest_batch;
    uint32_t max_batch, max_wait_us;
    uint32_t reserved;
} sg_batcher_stats;
int sg_batcher_create(sg_index *ix, uint32_t max_batch, uint32_t max_wait_us, uint32_t max_k, sg_batcher **out);
int sg_suggest_one(sg_batcher *b, const char *query, uint32_t len, int metric, double alpha, uint32_t k, uint32_t *out_ids,
                   double *out_scores, uint32_t *out_count);
int sg_batcher_get_stats(const sg_batcher *b, sg_batcher_stats *stats);
void sg_batcher_free(sg_batcher *b);

/*
 * ---- language model and spellchecker (BASELINE.json config #5) ----
 * sg_lm: the stupid-back-off n-gram model of pkg/lm in HBM.  Level i (0-based) holds the (i+1)-grams as the reference's
 * packed arrays (pkg/lm/packed_array.go): values[] = word << 32 | count ordered by (context, word), containers[] =
 * context << 32 | index of the context's first value.  Word ids are the ids of the vocabulary dictionary, which is also
 * the dictionary of the suggest index.
 *   sg_lm_create            NewNGramModel(indices)                pkg/lm/ngram_model.go:36-41
 *   sg_lm_open              nGramModel.Load of the binary model   pkg/lm/ngram_model.go:126-160 (the trailing MPH table is not read)
 *   sg_lm_score_batch       nGramModel.Score per n-gram           pkg/lm/ngram_model.go:44-64, calcScore :163-175
 *   sg_lm_score_next_batch  nGramModel.Next(context) then ScorerNext.ScoreNext(candidate)   ngram_model.go:67-99, scorer_next.go:15-23;
 *                           out_has_scorer[q] = 0 where Next returns a nil scorer (every candidate then scores -100)
 *   sg_predict_batch        SpellChecker.Predict                  pkg/spellchecker/spellchecker.go:40-92, for queries already split
 *                           into the last word (w_bytes / w_off) and the word ids of the context as languageModel.Next passes
 *                           them to the model (language_model.go:103-115).  Rows of out_ids have stride k + 1 (the reference
 *                           keeps k + 1 candidates when it has more than k, spellchecker.go:87-89).
 * Host buffers throughout.
 */
typedef struct sg_lm sg_lm;
int sg_lm_create(uint32_t order, const uint64_t *const *containers, const uint64_t *n_containers, const uint64_t *const *values,
                 const uint64_t *n_values, const uint32_t *totals, int device, sg_lm **out);
int sg_lm_open(const char *path, int device, sg_lm **out);
void sg_lm_free(sg_lm *lm);
int sg_lm_score_batch(sg_lm *lm, const uint32_t *ids, const uint32_t *off, uint32_t n, double *out_scores);
int sg_lm_score_next_batch(sg_lm *lm, const uint32_t *ctx_ids, const uint32_t *ctx_off, uint32_t n_q, const uint32_t *cand_ids,
                           const uint32_t *cand_off, double *out_scores, uint8_t *out_has_scorer);
int sg_predict_batch(sg_index *ix, sg_lm *lm, const char *w_bytes, const uint32_t *w_off, const uint32_t *ctx_ids,
                     const uint32_t *ctx_off, uint32_t n_q, double similarity, uint32_t k, uint32_t *out_ids, uint32_t *out_counts);

/*
 * Page-locked, device-addressable host memory for query and result buffers (a micro-batcher allocates its buffers once
 * and reuses them for every sg_search_batch call).  sg_is_pinned: 1 if [p, p + bytes) is such memory - what
 * sg_search_batch checks to pick the direct result path - else 0.  No counterpart in the reference (Go heap memory).
 */
int sg_pinned_alloc(uint64_t bytes, void **out);
void sg_pinned_free(void *p);
int sg_is_pinned(const void *p, uint64_t bytes);

/* number of kernels this library has launched in the calling process (bench.py "gpu_launches") */
uint64_t sg_kernel_launches(void);

const char *sg_last_error(void);
const char *sg_version(void);

/*
 * Host-only introspection (no GPU is touched): the tokenizer chain and the CSR build exactly as the
 * library runs them before the upload, exposed so that the CPU test-suite can compare them with the
 * reference's rules (pkg/suggest/tokenizer.go:9-20, pkg/index/indexer_writer.go:66-86).  Searching
 * has no host implementation.
 */
typedef struct sg_host_index sg_host_index;
int sg_host_tokenize(const sg_config *cfg, const char *text, uint32_t len, char *out_bytes, uint32_t cap,
                     uint32_t *tok_off, uint32_t max_tok);   /* returns the token count; tok_off[count + 1] */
int sg_host_to_lower(const char *text, uint32_t len, char *out, uint32_t cap); /* strings.ToLower; returns length */
int sg_host_index_build(const sg_config *cfg, const char *doc_bytes, const uint64_t *doc_off, uint32_t n_docs,
                        sg_host_index **out);
int sg_host_index_open_disk(const sg_config *cfg, const char *hd_path, const char *dl_path, sg_host_index **out);
void sg_host_index_free(sg_host_index *hi);
int sg_host_index_get_info(const sg_host_index *hi, sg_index_info *info);
/* original ids of list (segment, term), ascending; returns the length, -1 if absent, -2 if cap is too small */
int64_t sg_host_index_get_list(const sg_host_index *hi, uint32_t segment, const char *term, uint32_t term_len,
                               uint32_t *out, uint64_t cap);
int sg_host_index_get_layout(const sg_host_index *hi, sg_index_layout *layout);
/* first slot of every segment, n_segments + 1 entries; returns the count or -2 if cap is too small */
int64_t sg_host_index_get_segments(const sg_host_index *hi, uint32_t *out, uint64_t cap);
/* slots (not original ids) of list (segment, term); same return convention as sg_host_index_get_list */
int64_t sg_host_index_get_list_slots(const sg_host_index *hi, uint32_t segment, const char *term, uint32_t term_len,
                                     uint32_t *out, uint64_t cap);
/* the term's bitmap row, row_words entries; -1 if the term is absent or the index has no bitmaps, -2 if cap is too small */
int64_t sg_host_index_get_bitmap(const sg_host_index *hi, const char *term, uint32_t term_len, uint32_t *out, uint64_t cap);
const char *sg_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
