// sg_hostapi.cpp — host-only introspection entry points (no GPU needed): the tokenizer chain and the
// CSR build exactly as the library performs them before the upload.  The CPU test-suite checks them
// against the oracle; the search path itself has no host implementation.
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "../../include/suggest_b200.h"
#include "sg_host.h"

struct sg_host_index {
    sg::HostIndex h;
};

namespace {
thread_local std::string g_host_err;
}

extern "C" {

const char *sg_host_last_error(void) { return g_host_err.c_str(); }

int sg_host_tokenize(const sg_config *cfg, const char *text, uint32_t len, char *out_bytes, uint32_t cap,
                     uint32_t *tok_off, uint32_t max_tok) {
    if (!cfg || (!text && len) || !tok_off) { g_host_err = "null argument"; return SG_ERR_INVALID; }
    sg::TextConfig tc;
    std::string err = tc.init(cfg->ngram_size, cfg->wrap_start, cfg->wrap_end, cfg->pad, cfg->alphabet, cfg->n_alphabet);
    if (!err.empty()) { g_host_err = err; return SG_ERR_UNSUPPORTED; }
    std::vector<uint64_t> keys;
    sg::TokenScratch sc;
    sg::tokenize_keys(tc, (const uint8_t *)text, len, &keys, &sc);
    if (keys.size() > max_tok) { g_host_err = "token buffer too small"; return SG_ERR_INVALID; }
    std::string out;
    tok_off[0] = 0;
    const uint64_t mask = (1ull << tc.bits) - 1;
    for (size_t t = 0; t < keys.size(); t++) {
        for (int i = 0; i < tc.n; i++) {
            uint32_t code = (uint32_t)((keys[t] >> (tc.bits * i)) & mask);
            if (!code) break;
            uint8_t enc[4];
            int e = sg::utf8_encode(tc.code_rune[code], enc);
            out.append((const char *)enc, (size_t)e);
        }
        tok_off[t + 1] = (uint32_t)out.size();
    }
    if (out.size() > cap) { g_host_err = "byte buffer too small"; return SG_ERR_INVALID; }
    if (!out.empty()) std::memcpy(out_bytes, out.data(), out.size());
    return (int)keys.size();
}

int sg_host_to_lower(const char *text, uint32_t len, char *out, uint32_t cap) {
    std::string low;
    sg::to_lower((const uint8_t *)text, len, &low);
    if (low.size() > cap) return SG_ERR_INVALID;
    if (!low.empty()) std::memcpy(out, low.data(), low.size());
    return (int)low.size();
}

int sg_host_index_build(const sg_config *cfg, const char *doc_bytes, const uint64_t *doc_off, uint32_t n_docs,
                        sg_host_index **out) {
    if (!cfg || !out || (n_docs && (!doc_bytes || !doc_off))) { g_host_err = "null argument"; return SG_ERR_INVALID; }
    sg_host_index *hi = new (std::nothrow) sg_host_index();
    if (!hi) return SG_ERR_NOMEM;
    std::string err = hi->h.text.init(cfg->ngram_size, cfg->wrap_start, cfg->wrap_end, cfg->pad, cfg->alphabet, cfg->n_alphabet);
    static const uint64_t zero_off[1] = {0};
    if (const char *v = std::getenv("SG_BUCKET_SHIFT")) if (*v) hi->h.want_bshift = std::atoi(v);
    if (const char *v = std::getenv("SG_BITMAP_MAX_MB")) if (*v) hi->h.bitmap_budget = (uint64_t)std::atoll(v) << 20;
    if (err.empty()) err = sg::build_from_docs(&hi->h, doc_bytes, n_docs ? doc_off : zero_off, n_docs);
    if (!err.empty()) { g_host_err = err; delete hi; return SG_ERR_UNSUPPORTED; }
    *out = hi;
    return SG_OK;
}

int sg_host_index_open_disk(const sg_config *cfg, const char *hd_path, const char *dl_path, sg_host_index **out) {
    if (!cfg || !out || !hd_path || !dl_path) { g_host_err = "null argument"; return SG_ERR_INVALID; }
    sg_host_index *hi = new (std::nothrow) sg_host_index();
    if (!hi) return SG_ERR_NOMEM;
    std::string err = hi->h.text.init(cfg->ngram_size, cfg->wrap_start, cfg->wrap_end, cfg->pad, cfg->alphabet, cfg->n_alphabet);
    if (err.empty()) err = sg::build_from_disk(&hi->h, hd_path, dl_path);
    if (!err.empty()) { g_host_err = err; delete hi; return err.rfind("io:", 0) == 0 ? SG_ERR_IO : SG_ERR_FORMAT; }
    *out = hi;
    return SG_OK;
}

void sg_host_index_free(sg_host_index *hi) { delete hi; }

int sg_host_index_get_info(const sg_host_index *hi, sg_index_info *info) {
    if (!hi || !info) return SG_ERR_INVALID;
    info->n_docs = hi->h.n_docs;
    info->n_segments = hi->h.n_segments;
    info->n_terms = (uint32_t)hi->h.term_keys.size();
    info->n_lists = hi->h.n_lists;
    info->n_postings = hi->h.n_postings;
    info->device_bytes = 0;
    info->id_base = 0;
    info->device = -1;
    return SG_OK;
}

static int64_t term_id_of(const sg::HostIndex &h, const char *term, uint32_t term_len) {
    uint64_t key = h.text.key_of_term((const uint8_t *)term, term_len);
    if (key == 0) return -1;
    size_t m = h.ht_keys.size() - 1, s = (size_t)sg::mix64(key) & m;
    while (h.ht_keys[s] != 0 && h.ht_keys[s] != key) s = (s + 1) & m;
    if (h.ht_keys[s] == 0) return -1;
    return (int64_t)h.ht_vals[s];
}

static int64_t get_list(const sg_host_index *hi, uint32_t segment, const char *term, uint32_t term_len, uint32_t *out,
                        uint64_t cap, bool slots) {
    if (!hi || segment >= hi->h.n_segments) return -1;
    const sg::HostIndex &h = hi->h;
    const int64_t t = term_id_of(h, term, term_len);
    if (t < 0) return -1;
    const size_t stride = (size_t)h.n_segments + 1;
    const uint32_t a = h.list_off[(size_t)t * stride + segment], b = h.list_off[(size_t)t * stride + segment + 1];
    if (a == b) return -1;
    if (out) {
        if (cap < b - a) return -2;
        for (uint32_t i = a; i < b; i++) out[i - a] = slots ? h.postings[i] : h.perm[h.postings[i]];
    }
    return (int64_t)(b - a);
}

int64_t sg_host_index_get_list(const sg_host_index *hi, uint32_t segment, const char *term, uint32_t term_len,
                               uint32_t *out, uint64_t cap) {
    return get_list(hi, segment, term, term_len, out, cap, false);
}

int64_t sg_host_index_get_list_slots(const sg_host_index *hi, uint32_t segment, const char *term, uint32_t term_len,
                                     uint32_t *out, uint64_t cap) {
    return get_list(hi, segment, term, term_len, out, cap, true);
}

int sg_host_index_get_layout(const sg_host_index *hi, sg_index_layout *layout) {
    if (!hi || !layout) return SG_ERR_INVALID;
    layout->n_slots = hi->h.n_ids;
    layout->bucket_shift = hi->h.bshift;
    layout->row_words = hi->h.row_words;
    layout->engine = hi->h.row_words != 0;
    layout->built_on_device = 0;
    layout->pipeline = 0;
    layout->bitmap_bytes = (uint64_t)hi->h.bitmaps.size() * sizeof(uint32_t);
    return SG_OK;
}

int64_t sg_host_index_get_segments(const sg_host_index *hi, uint32_t *out, uint64_t cap) {
    if (!hi || !out) return -1;
    if (cap < hi->h.seg_start.size()) return -2;
    std::memcpy(out, hi->h.seg_start.data(), hi->h.seg_start.size() * sizeof(uint32_t));
    return (int64_t)hi->h.seg_start.size();
}

int64_t sg_host_index_get_bitmap(const sg_host_index *hi, const char *term, uint32_t term_len, uint32_t *out, uint64_t cap) {
    if (!hi || hi->h.row_words == 0) return -1;
    const int64_t t = term_id_of(hi->h, term, term_len);
    if (t < 0) return -1;
    if (cap < hi->h.row_words) return -2;
    if (out) std::memcpy(out, hi->h.bitmaps.data() + (size_t)t * hi->h.row_words, (size_t)hi->h.row_words * sizeof(uint32_t));
    return (int64_t)hi->h.row_words;
}

}  // extern "C"
