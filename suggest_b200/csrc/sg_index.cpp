// sg_index.cpp — host construction of the HBM index layout (sg_device.h: DevIndex).
//
// The reference keeps one map[term][]docID per n-gram cardinality ("segment"), each list stored as
// VB / skipping / roaring bytes and decoded lazily on every query (pkg/index/indexer_writer.go:66-145,
// pkg/index/posting_list.go).  Here every list is decoded once and laid out so that a query reads
// one contiguous run per query token: documents are renumbered by (segment, original id), and the
// postings of a term are stored back to back over all segments, ascending in the new id.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <unordered_map>

#include "sg_host.h"

namespace sg {

void HostIndex::build_hash() {
    size_t cap = 16;
    while (cap < term_keys.size() * 2 + 2) cap <<= 1;
    ht_keys.assign(cap, 0);
    ht_vals.assign(cap, kNoTerm);
    for (size_t t = 0; t < term_keys.size(); t++) {
        size_t h = (size_t)mix64(term_keys[t]) & (cap - 1);
        while (ht_keys[h] != 0) h = (h + 1) & (cap - 1);
        ht_keys[h] = term_keys[t];
        ht_vals[h] = (uint32_t)t;
    }
}

// Bucket width of the bitmaps (2^bshift slots per bit), row length, and the segment starts aligned to it.
// Dictionaries of up to ~3M postings get one bit per document (the bit count is then the overlap itself).  Otherwise: a query's terms are
// drawn like the dictionary's own n-grams, so weigh every term by its number of postings f; W(s) = sum f * (1 - exp(-f * 2^s
// / n_docs)) / sum f is the fraction of buckets one list of a typical query hits at width 2^s, and the width is the one
// that minimises an estimate of the work per query (below).  Shared by the host build and the device build (sg_gpubuild.cu).
std::string choose_layout(const std::vector<uint32_t> &seg_count, const std::vector<uint32_t> &freq, uint32_t n_docs, uint64_t n_postings,
                          int want_bshift, uint64_t bitmap_budget, uint32_t *bshift, uint32_t *row_words, std::vector<uint32_t> *seg_start,
                          uint32_t *n_ids) {
    const uint32_t S = (uint32_t)seg_count.size();
    const size_t n_terms = freq.size();
    auto ids_at = [&](uint32_t s) {  // slots once every segment start is aligned to 2^s
        uint64_t n = 0;
        for (uint32_t b = 0; b < S; b++) n = ((n + ((1ull << s) - 1)) >> s << s) + seg_count[b];
        return n;
    };
    auto row_words_at = [&](uint32_t s) { return (((ids_at(s) + ((1ull << s) - 1)) >> s) + 2047) / 2048 * 64 + (ids_at(s) ? 0 : 64); };  // whole 64-word tiles
    uint32_t bs = 0;
    if (want_bshift >= 0) bs = std::min<uint32_t>((uint32_t)want_bshift, kMaxBucketShift);
    else if (n_docs > 0 && n_terms > 0 && n_postings > 0 && (double)n_postings * 0.0054 > 16000.0) {
        // Up to ~3M postings one bit per document stays: a query then adds ~c * n_docs / 32 words (c = mean n-grams per
        // document), which is cheap, and no bucket is ever resolved, whatever the metric and similarity ask for.  (Measured
        // on the reference's 235,887-word list: equal to 4 documents per bit for Jaccard 0.5, twice as fast for Cosine /
        // Dice 0.5 and Jaccard 0.3.)  Above that:
        const double c = (double)n_postings / (double)n_docs;
        const int T = std::max(2, (int)std::lround(0.55 * c));
        double best = 0.0;
        for (uint32_t s = 0; s <= kMaxBucketShift; s++) {
            double num = 0.0, den = 0.0;
            for (uint32_t f : freq) {
                num += (double)f * (1.0 - std::exp(-(double)f * (double)(1u << s) / (double)n_docs));
                den += (double)f;
            }
            const double lam = c * num / den, buckets = (double)((ids_at(s) + ((1ull << s) - 1)) >> s);
            double term = std::exp(-lam), tail = 1.0;  // P(Poisson(lam) >= T) = 1 - sum_{i<T} e^-lam lam^i / i!
            for (int i = 0; i < T; i++) { tail -= term; term *= lam / (double)(i + 1); }
            if (tail < 0.0) tail = 0.0;
            const double cost = c * buckets * 0.0054 + buckets * tail * 600.0;
            if (s == 0 || cost < best) { best = cost; bs = s; }
        }
    }
    while (bs < kMaxBucketShift && (n_terms + 1) * row_words_at(bs) * 4 > bitmap_budget) bs++;
    const bool with_bitmaps = (n_terms + 1) * row_words_at(bs) * 4 <= bitmap_budget && (n_terms + 1) * row_words_at(bs) < 0xFFFFFFF0ull;
    if (ids_at(bs) > 0xFFFFFFF0ull) return "more than 2^32 document ids";
    *bshift = bs;
    *row_words = with_bitmaps ? (uint32_t)row_words_at(bs) : 0u;
    // every segment starts at a multiple of the bucket width, so a bucket never holds documents of two segments
    seg_start->assign((size_t)S + 1, 0);
    for (uint32_t b = 0; b < S; b++) {
        const uint64_t a = ((uint64_t)(*seg_start)[b] + ((1ull << bs) - 1)) >> bs << bs;
        (*seg_start)[b] = (uint32_t)a;  // an empty segment moves with its successor's alignment
        (*seg_start)[b + 1] = (uint32_t)(a + seg_count[b]);
    }
    *n_ids = (*seg_start)[S];
    return "";
}

namespace {

// Common tail: given per-document (segment, distinct term ids) produce the CSR arrays.
// doc_terms[doc_term_off[d] .. doc_term_off[d+1]) are the distinct term ids of document d.
std::string finish(HostIndex *ix, const std::vector<uint32_t> &doc_seg, const std::vector<uint64_t> &doc_term_off,
                   const std::vector<uint32_t> &doc_terms, uint32_t n_segments) {
    const uint32_t n_docs = (uint32_t)doc_seg.size();
    const size_t n_terms = ix->term_keys.size();
    const uint32_t S = n_segments;
    ix->n_docs = n_docs;
    ix->n_segments = S;
    if ((uint64_t)n_terms * (S + 1) > 0xFFFFFFF0ull) return "term x segment offset table exceeds 32 bits";
    if (doc_terms.size() > 0xFFFFFFF0ull) return "more than 2^32 postings";
    std::vector<uint32_t> seg_count((size_t)S, 0);
    for (uint32_t d = 0; d < n_docs; d++) seg_count[doc_seg[d]]++;
    std::vector<uint32_t> freq(n_terms, 0);
    for (uint32_t t : doc_terms) freq[t]++;
    {
        std::string err = choose_layout(seg_count, freq, n_docs, doc_terms.size(), ix->want_bshift, ix->bitmap_budget, &ix->bshift,
                                        &ix->row_words, &ix->seg_start, &ix->n_ids);
        if (!err.empty()) return err;
    }
    const uint32_t bs = ix->bshift;
    ix->perm.assign(ix->n_ids, 0xFFFFFFFFu);
    {
        std::vector<uint32_t> cur(ix->seg_start.begin(), ix->seg_start.end() - 1);
        for (uint32_t d = 0; d < n_docs; d++) ix->perm[cur[doc_seg[d]]++] = d;
    }
    // list sizes -> offsets, row-major [term][segment]
    const size_t stride = (size_t)S + 1;
    ix->list_off.assign(n_terms * stride, 0);
    std::vector<uint32_t> &off = ix->list_off;
    for (uint32_t d = 0; d < n_docs; d++)
        for (uint64_t j = doc_term_off[d]; j < doc_term_off[d + 1]; j++) off[(size_t)doc_terms[j] * stride + doc_seg[d]]++;
    uint64_t run = 0, lists = 0;
    for (size_t t = 0; t < n_terms; t++) {
        for (uint32_t b = 0; b < S; b++) {
            uint32_t c = off[t * stride + b];
            off[t * stride + b] = (uint32_t)run;
            run += c;
            lists += c != 0;
        }
        off[t * stride + S] = (uint32_t)run;
    }
    ix->n_postings = run;
    ix->n_lists = lists;
    ix->postings.assign(((size_t)run + 3) / 4 * 4 + 4, 0xFFFFFFFFu);
    // fill in new-id order so that every list comes out ascending
    std::vector<uint32_t> cur(n_terms * stride);
    std::memcpy(cur.data(), off.data(), cur.size() * sizeof(uint32_t));
    const size_t rw = ix->row_words;
    ix->bitmaps.assign(rw ? (n_terms + 1) * rw : 0, 0u);
    for (uint32_t nid = 0; nid < ix->n_ids; nid++) {
        const uint32_t d = ix->perm[nid];
        if (d == 0xFFFFFFFFu) continue;  // alignment hole
        const uint32_t b = doc_seg[d], bucket = nid >> bs;
        for (uint64_t j = doc_term_off[d]; j < doc_term_off[d + 1]; j++) {
            const size_t t = doc_terms[j];
            ix->postings[cur[t * stride + b]++] = nid;
            if (rw) ix->bitmaps[t * rw + (bucket >> 5)] |= 1u << (bucket & 31);
        }
    }
    ix->build_hash();
    return "";
}

}  // namespace

std::string build_from_docs(HostIndex *ix, const char *doc_bytes, const uint64_t *doc_off, uint32_t n_docs) {
    std::unordered_map<uint64_t, uint32_t> term_of;
    term_of.reserve(1 << 16);
    std::vector<uint32_t> doc_seg(n_docs);
    std::vector<uint64_t> doc_term_off((size_t)n_docs + 1, 0);
    std::vector<uint32_t> doc_terms;
    doc_terms.reserve((size_t)n_docs * 16);
    std::vector<uint64_t> keys;
    std::vector<uint32_t> tids;
    TokenScratch sc;
    uint32_t max_card = 0;
    ix->term_keys.clear();
    for (uint32_t d = 0; d < n_docs; d++) {
        tokenize_keys(ix->text, (const uint8_t *)doc_bytes + doc_off[d], (size_t)(doc_off[d + 1] - doc_off[d]), &keys, &sc);
        // the segment counts duplicates (pkg/index/indexer_writer.go:67); a posting list holds the document once
        if (keys.size() > 0xFFFFu) return "a document has more than 65535 n-grams";
        uint32_t card = (uint32_t)keys.size();
        doc_seg[d] = card;
        max_card = std::max(max_card, card);
        tids.clear();
        for (uint64_t k : keys) {
            auto it = term_of.find(k);
            uint32_t t;
            if (it == term_of.end()) {
                t = (uint32_t)ix->term_keys.size();
                term_of.emplace(k, t);
                ix->term_keys.push_back(k);
            } else t = it->second;
            tids.push_back(t);
        }
        std::sort(tids.begin(), tids.end());
        tids.erase(std::unique(tids.begin(), tids.end()), tids.end());
        doc_terms.insert(doc_terms.end(), tids.begin(), tids.end());
        doc_term_off[d + 1] = doc_terms.size();
    }
    return finish(ix, doc_seg, doc_term_off, doc_terms, max_card + 1);
}

std::string build_from_lists(HostIndex *ix, uint32_t n_segments, uint64_t n_lists, const uint32_t *list_segment,
                             const char *term_bytes, const uint64_t *list_term_off, const uint32_t *ids,
                             const uint64_t *list_off) {
    // invert the lists back into per-document term sets; a document's segment is the segment of its lists
    std::unordered_map<uint64_t, uint32_t> term_of;
    ix->term_keys.clear();
    uint32_t max_id = 0;
    bool any = false;
    for (uint64_t l = 0; l < n_lists; l++)
        for (uint64_t j = list_off[l]; j < list_off[l + 1]; j++) { max_id = std::max(max_id, ids[j]); any = true; }
    const uint32_t n_docs = any ? max_id + 1 : 0;
    std::vector<uint32_t> doc_seg(n_docs, 0);
    std::vector<uint64_t> doc_term_off((size_t)n_docs + 1, 0);
    std::vector<uint32_t> list_tid(n_lists);
    for (uint64_t l = 0; l < n_lists; l++) {
        if (list_segment[l] >= n_segments) return "list segment out of range";
        uint64_t key = ix->text.key_of_term((const uint8_t *)term_bytes + list_term_off[l],
                                            (size_t)(list_term_off[l + 1] - list_term_off[l]));
        if (key == 0) return "a term of the index cannot be produced by this index description";
        auto it = term_of.find(key);
        if (it == term_of.end()) {
            list_tid[l] = (uint32_t)ix->term_keys.size();
            term_of.emplace(key, list_tid[l]);
            ix->term_keys.push_back(key);
        } else list_tid[l] = it->second;
        uint32_t prev = 0xFFFFFFFFu;
        for (uint64_t j = list_off[l]; j < list_off[l + 1]; j++) {
            if (ids[j] == prev) continue;  // an id repeated inside one list counts once (pkg/merger/scan_count.go:35-66)
            if (prev != 0xFFFFFFFFu && ids[j] < prev) return "posting list is not ascending";
            prev = ids[j];
            doc_seg[ids[j]] = list_segment[l];
            doc_term_off[(size_t)ids[j] + 1]++;
        }
    }
    for (uint32_t d = 0; d < n_docs; d++) doc_term_off[d + 1] += doc_term_off[d];
    std::vector<uint32_t> doc_terms(doc_term_off[n_docs]);
    std::vector<uint64_t> cur(doc_term_off.begin(), doc_term_off.end() - 1);
    for (uint64_t l = 0; l < n_lists; l++) {
        uint32_t prev = 0xFFFFFFFFu;
        for (uint64_t j = list_off[l]; j < list_off[l + 1]; j++) {
            if (ids[j] == prev) continue;
            prev = ids[j];
            if (doc_seg[ids[j]] != list_segment[l]) return "a document appears in two segments";
            doc_terms[cur[ids[j]]++] = list_tid[l];
        }
    }
    return finish(ix, doc_seg, doc_term_off, doc_terms, n_segments);
}

}  // namespace sg
