// sg_disk.cpp — reader for indexes written by the reference's `suggest indexer`: `<name>.hd` + `<name>.dl`.
//
// Replaces index.Reader.Read (pkg/index/index_reader.go:29-120).  The reference keeps the `.dl` bytes
// mapped and decodes a list on every query; here every list is decoded exactly once, at load:
//   header   gob stream of header{Version, Indices, Terms[]termDescription}  (pkg/index/indexer_writer.go:50-63,148-166)
//   <= 65    VB: LEB128 deltas from 0                                         (pkg/compression/varint.go:36-78)
//   66..256  skipping(64): per block a little-endian uint16 (block bytes + 2, bit 15 = last block)
//            followed by VB deltas, the first one relative to the first id of the previous block
//                                                                             (pkg/compression/skipping.go:67-146)
//   > 256    roaring bitmap, portable serialisation (RoaringBitmap/roaring v0.5.5 WriteTo, go.mod:6;
//            pkg/compression/bitmap.go:18-29): cookie 12346 / 12347, array, bitmap and run containers
// The length classes are pkg/index/codec.go:39-51.
#include <cstdio>
#include <cstring>

#include "sg_host.h"

namespace sg {

namespace {

bool read_file(const char *path, std::vector<uint8_t> *out) {
    FILE *f = std::fopen(path, "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (n < 0) { std::fclose(f); return false; }
    out->resize((size_t)n);
    size_t got = n ? std::fread(out->data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    return got == (size_t)n;
}

struct Cursor {
    const uint8_t *p;
    size_t n, i = 0;
    bool ok = true;
    Cursor(const uint8_t *p_, size_t n_) : p(p_), n(n_) {}
    uint8_t u8() { if (i >= n) { ok = false; return 0; } return p[i++]; }
    // encoding/gob unsigned integer: one byte < 128, else a negated byte count followed by big-endian bytes
    uint64_t gob_uint() {
        uint8_t c = u8();
        if (c < 0x80) return c;
        int cnt = 256 - (int)c;
        if (cnt > 8) { ok = false; return 0; }
        uint64_t v = 0;
        for (int k = 0; k < cnt; k++) v = (v << 8) | u8();
        return v;
    }
    int64_t gob_int() {
        uint64_t u = gob_uint();
        return (u & 1) ? ~(int64_t)(u >> 1) : (int64_t)(u >> 1);
    }
    uint32_t vu32() {  // store.Input.ReadVUInt32 (pkg/store/byte_input.go:130-193)
        uint32_t v = 0;
        for (int shift = 0; shift < 35; shift += 7) {
            uint8_t b = u8();
            v |= (uint32_t)(b & 0x7F) << shift;
            if (!(b & 0x80)) return v;
        }
        ok = false;
        return 0;
    }
    uint16_t le16() { uint16_t a = u8(); return (uint16_t)(a | ((uint16_t)u8() << 8)); }
    uint32_t le32() { uint32_t a = le16(); return a | ((uint32_t)le16() << 16); }
};

struct TermDescription {
    std::string term;
    uint32_t indice = 0, bytes = 0, position = 0, len = 0;
};

std::string parse_header(const std::vector<uint8_t> &hd, uint32_t *indices, std::vector<TermDescription> *terms) {
    Cursor c(hd.data(), hd.size());
    // messages: [length][type id][body]; negative ids carry type definitions, which are fixed here
    size_t end = 0;
    for (;;) {
        uint64_t len = c.gob_uint();
        if (!c.ok || len == 0 || c.i > hd.size() || len > hd.size() - c.i) return "corrupt gob header";  // no sum of a length from the file: it could wrap
        end = c.i + (size_t)len;
        int64_t tid = c.gob_int();
        if (!c.ok) return "corrupt gob header";
        if (tid >= 0) break;
        c.i = end;
    }
    std::string version;
    int field = -1;
    for (;;) {
        uint64_t d = c.gob_uint();
        if (!c.ok) return "corrupt gob header";
        if (d == 0) break;
        field += (int)d;
        if (field == 0) {
            uint64_t n = c.gob_uint();
            if (!c.ok || c.i > hd.size() || n > hd.size() - c.i) return "corrupt gob header";
            version.assign((const char *)hd.data() + c.i, (size_t)n);
            c.i += (size_t)n;
        } else if (field == 1) {
            *indices = (uint32_t)c.gob_uint();
        } else if (field == 2) {
            uint64_t cnt = c.gob_uint();
            if (!c.ok || cnt > hd.size()) return "corrupt gob header";
            terms->reserve((size_t)cnt);
            for (uint64_t t = 0; t < cnt; t++) {
                TermDescription td;
                int f = -1;
                for (;;) {
                    uint64_t d2 = c.gob_uint();
                    if (!c.ok) return "corrupt gob header";
                    if (d2 == 0) break;
                    f += (int)d2;
                    if (f == 0) {
                        uint64_t n = c.gob_uint();
                        if (!c.ok || c.i > hd.size() || n > hd.size() - c.i) return "corrupt gob header";
                        td.term.assign((const char *)hd.data() + c.i, (size_t)n);
                        c.i += (size_t)n;
                    } else {
                        uint32_t v = (uint32_t)c.gob_uint();
                        if (f == 1) td.indice = v; else if (f == 2) td.bytes = v; else if (f == 3) td.position = v;
                        else if (f == 4) td.len = v; else return "unknown field in termDescription";
                    }
                }
                terms->push_back(std::move(td));
            }
        } else return "unknown field in index header";
    }
    if (!c.ok || c.i != end) return "corrupt gob header";
    if (version != "v5.1") return "index version mismatch, expected v5.1 version";  // index_reader.go:71-73
    return "";
}

bool decode_vb(Cursor &c, uint32_t n, std::vector<uint32_t> *out) {
    uint32_t prev = 0;
    for (uint32_t i = 0; i < n; i++) { prev += c.vu32(); out->push_back(prev); }
    return c.ok;
}

bool decode_skipping(Cursor &c, uint32_t n, uint32_t gap, std::vector<uint32_t> *out) {
    uint32_t block_first = 0;
    for (uint32_t i = 0; i < n; i += gap) {
        c.le16();
        uint32_t j = i + gap < n ? i + gap : n;
        uint32_t prev = block_first;
        for (uint32_t k = i; k < j; k++) {
            prev += c.vu32();
            if (k == i) block_first = prev;
            out->push_back(prev);
        }
    }
    return c.ok;
}

// RoaringBitmap portable format
bool decode_roaring(Cursor &c, std::vector<uint32_t> *out) {
    const uint32_t cookie = c.le32();
    uint32_t size;
    std::vector<uint8_t> is_run;
    bool has_runs = false;
    if ((cookie & 0xFFFF) == 12347) {
        has_runs = true;
        size = (cookie >> 16) + 1;
        is_run.resize((size + 7) / 8);
        for (auto &b : is_run) b = c.u8();
    } else if (cookie == 12346) {
        size = c.le32();
    } else return false;
    if (!c.ok || size > 65536) return false;
    std::vector<uint16_t> keys(size);
    std::vector<uint32_t> cards(size);
    for (uint32_t k = 0; k < size; k++) { keys[k] = c.le16(); cards[k] = (uint32_t)c.le16() + 1; }
    if (!has_runs || size >= 4) for (uint32_t k = 0; k < size; k++) c.le32();  // offset header
    if (!c.ok) return false;
    for (uint32_t k = 0; k < size; k++) {
        const uint32_t hi = (uint32_t)keys[k] << 16;
        if (has_runs && (is_run[k / 8] >> (k % 8) & 1)) {
            uint32_t n_runs = c.le16();
            for (uint32_t r = 0; r < n_runs; r++) {
                uint32_t start = c.le16(), len = c.le16();
                if (!c.ok) return false;
                for (uint32_t v = start; v <= start + len; v++) out->push_back(hi | v);
            }
        } else if (cards[k] > 4096) {
            for (uint32_t w = 0; w < 1024; w++) {
                uint64_t bits = (uint64_t)c.le32();
                bits |= (uint64_t)c.le32() << 32;
                while (bits) {
                    out->push_back(hi | (w * 64 + (uint32_t)__builtin_ctzll(bits)));
                    bits &= bits - 1;
                }
            }
        } else {
            for (uint32_t v = 0; v < cards[k]; v++) out->push_back(hi | c.le16());
        }
        if (!c.ok) return false;
    }
    return true;
}

}  // namespace

std::string build_from_disk(HostIndex *ix, const char *hd_path, const char *dl_path) {
    std::vector<uint8_t> hd, dl;
    if (!read_file(hd_path, &hd)) return std::string("io: failed to open header: ") + hd_path;
    if (!read_file(dl_path, &dl)) return std::string("io: failed to open document list: ") + dl_path;
    uint32_t indices = 0;
    std::vector<TermDescription> terms;
    std::string err = parse_header(hd, &indices, &terms);
    if (!err.empty()) return err;
    std::vector<uint32_t> list_segment, ids;
    std::vector<uint64_t> list_term_off(1, 0), list_off(1, 0);
    std::string term_bytes;
    for (const TermDescription &td : terms) {
        if (td.bytes == 0) continue;  // never written for a non-empty list
        if ((uint64_t)td.position + td.bytes > dl.size()) return "posting list outside the document list file";
        Cursor c(dl.data() + td.position, td.bytes);
        const size_t before = ids.size();
        bool ok;
        if (td.len <= 65) ok = decode_vb(c, td.len, &ids);                    // pkg/index/codec.go:79-81
        else if (td.len <= 256) ok = decode_skipping(c, td.len, 64, &ids);    // :83-85
        else ok = decode_roaring(c, &ids);                                    // :87
        if (!ok || ids.size() - before != td.len) return "corrupt posting list for term '" + td.term + "'";
        list_segment.push_back(td.indice);
        term_bytes += td.term;
        list_term_off.push_back(term_bytes.size());
        list_off.push_back(ids.size());
    }
    return build_from_lists(ix, indices ? indices : 1, list_segment.size(), list_segment.data(), term_bytes.data(),
                            list_term_off.data(), ids.data(), list_off.data());
}

}  // namespace sg
