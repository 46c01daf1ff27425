// sg_host.h — host-side pieces of libsuggest_b200: text handling, index description, CSR build.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "sg_device.h"

namespace sg {

// ---------------- text (pkg/analysis, pkg/alphabet, Go strings/unicode semantics) ----------------
constexpr uint32_t kRuneError = 0xFFFD;
int utf8_decode(const uint8_t *s, size_t len, uint32_t *rune);  // one `for range` step: width, invalid byte -> U+FFFD
int utf8_encode(uint32_t rune, uint8_t out[4]);
uint32_t rune_lower(uint32_t r);                                 // unicode.ToLower
void to_lower(const uint8_t *s, size_t len, std::string *out);   // strings.ToLower (appends)

// suggest.IndexDescription, pkg/suggest/config.go:25-35, reduced to what the tokenizer needs and
// encoded for the device: every rune a token can contain (alphabet members and the pad) has a
// small symbol code, a token is the concatenation of its codes.
struct TextConfig {
    int n = 0;
    std::string wrap_start, wrap_end;  // raw Wrap[0], Wrap[1]
    uint32_t pad_rune = 0;
    std::vector<std::string> alphabet; // description strings, kept for info
    // encoding
    int bits = 0;
    uint32_t pad_code = 0;
    uint32_t n_codes = 0;
    uint8_t ascii_code[128] = {0};
    std::vector<RuneRange> ranges;     // non-ASCII members
    std::vector<uint32_t> code_rune;   // code -> rune (code 0 unused)

    // returns "" or an error message (then the description is unsupported on the device)
    std::string init(int ngram, const char *wrap0, const char *wrap1, const char *pad, const char *const *alpha,
                     int n_alpha);
    uint32_t code_of(uint32_t rune) const;  // alphabet.Has(r) ? code(r) : pad_code
    // term bytes (as stored in a reference index header) -> packed key; 0 if it cannot be a token of this description
    uint64_t key_of_term(const uint8_t *term, size_t len) const;
};

// The tokenizer chain of pkg/suggest/tokenizer.go:9-20 producing packed keys.  Duplicates that
// appear after normalisation are kept (pkg/analysis/normalizer.go rewrites in place).
struct TokenScratch {
    std::string wrapped, lowered;
    std::vector<uint32_t> runes;
};
void tokenize_keys(const TextConfig &cfg, const uint8_t *text, size_t len, std::vector<uint64_t> *keys,
                   TokenScratch *scratch);

// ---------------- index in host memory, device layout (sg_device.h) ----------------
struct HostIndex {
    TextConfig text;
    uint32_t n_docs = 0, n_segments = 0, id_base = 0;
    uint64_t n_lists = 0;
    std::vector<uint64_t> term_keys;   // term id -> key
    std::vector<uint64_t> ht_keys;     // open addressing
    std::vector<uint32_t> ht_vals;
    std::vector<uint32_t> seg_start;   // S + 1
    std::vector<uint32_t> list_off;    // n_terms * (S + 1)
    std::vector<uint32_t> postings;    // padded to a multiple of 4 plus 4
    std::vector<uint32_t> perm;        // new id -> original id (n_ids entries, 0xFFFFFFFF in the alignment holes)
    uint64_t n_postings = 0;
    // bucket bitmaps (sg_device.h: DevIndex::bitmaps)
    int want_bshift = -1;              // in: < 0 lets the build choose
    uint64_t bitmap_budget = 1ull << 62;  // in: bytes the bitmaps may take; over it the index is built without them
    uint32_t n_ids = 0, bshift = 0, row_words = 0;
    std::vector<uint32_t> bitmaps;     // (n_terms + 1) * row_words, empty if over the budget

    void build_hash();
};

// bucket width, bitmap row length and aligned segment starts from the segment sizes and the terms' posting counts
std::string choose_layout(const std::vector<uint32_t> &seg_count, const std::vector<uint32_t> &freq, uint32_t n_docs, uint64_t n_postings,
                          int want_bshift, uint64_t bitmap_budget, uint32_t *bshift, uint32_t *row_words, std::vector<uint32_t> *seg_start,
                          uint32_t *n_ids);

// Index built on the device (sg_gpubuild.cu): device pointers, owned by the caller's allocation list.
struct GpuBuilt {
    const uint4 *term_table = nullptr;
    uint32_t term_mask = 0, n_terms = 0, n_segments = 0, n_ids = 0, bshift = 0, row_words = 0;
    const uint32_t *seg_start = nullptr, *list_off = nullptr, *postings = nullptr, *perm = nullptr, *bitmaps = nullptr;
    uint64_t n_postings = 0, n_lists = 0, device_bytes = 0;
    int kernel_launches = 0;
};
// `text`: a DevIndex whose tokenizer fields (and `ranges`, already on the device) are set.  "" = built; "fallback: ..." =
// not eligible (a document of more than 128 n-grams, more than 4 GB of text), build on the host instead; else a CUDA error.
std::string gpu_build(const DevIndex &text, const char *doc_bytes, const uint64_t *doc_off, uint32_t n_docs, int want_bshift,
                      uint64_t bitmap_budget, GpuBuilt *out, std::vector<void *> *allocs);

// sg_fine.cu: adds the exact level under the bucket bitmaps (DevIndex::rank4 / fine) to an index already in HBM.
// "" = built or not needed; "skip: ..." = over the budget (the index works without it); else a CUDA error text.
std::string build_fine_level(DevIndex *ix, uint64_t n_postings, uint64_t budget_bytes, std::vector<void *> *allocs,
                             uint64_t *device_bytes);

// suggest.Index (pkg/suggest/indexer.go:14-45) + index.Writer.AddDocument (pkg/index/indexer_writer.go:66-86)
std::string build_from_docs(HostIndex *ix, const char *doc_bytes, const uint64_t *doc_off, uint32_t n_docs);
// already decoded (segment, term) lists, original ids ascending (duplicates inside a list are dropped)
std::string build_from_lists(HostIndex *ix, uint32_t n_segments, uint64_t n_lists, const uint32_t *list_segment,
                             const char *term_bytes, const uint64_t *list_term_off, const uint32_t *ids,
                             const uint64_t *list_off);
// index.Reader.Read (pkg/index/index_reader.go:29-120) over `.hd` (gob) + `.dl` (VB / skipping / roaring)
std::string build_from_disk(HostIndex *ix, const char *hd_path, const char *dl_path);

}  // namespace sg
