// sg_text.cpp — host text handling for libsuggest_b200: Go-compatible UTF-8 / ToLower, the alphabet
// description (pkg/alphabet) turned into symbol codes, and the tokenizer chain (pkg/suggest/tokenizer.go:9-20)
// emitting packed term keys.  Used when an index is built from dictionary text and to lower-case
// non-ASCII queries before they are copied to the device.
#include <algorithm>
#include <cstring>

#include "sg_host.h"

namespace sg {

// utf8.DecodeRune acceptance rules (Go `for range` over a string)
int utf8_decode(const uint8_t *s, size_t len, uint32_t *rune) {
    uint8_t b0 = s[0];
    if (b0 < 0x80) { *rune = b0; return 1; }
    int need;
    uint32_t r;
    uint8_t lo = 0x80, hi = 0xBF;
    if (b0 >= 0xC2 && b0 <= 0xDF) { need = 1; r = b0 & 0x1Fu; }
    else if (b0 >= 0xE0 && b0 <= 0xEF) { need = 2; r = b0 & 0x0Fu; if (b0 == 0xE0) lo = 0xA0; if (b0 == 0xED) hi = 0x9F; }
    else if (b0 >= 0xF0 && b0 <= 0xF4) { need = 3; r = b0 & 0x07u; if (b0 == 0xF0) lo = 0x90; if (b0 == 0xF4) hi = 0x8F; }
    else { *rune = kRuneError; return 1; }
    if (len < (size_t)need + 1) { *rune = kRuneError; return 1; }
    for (int i = 1; i <= need; i++) {
        uint8_t b = s[i];
        uint8_t l = i == 1 ? lo : 0x80, h = i == 1 ? hi : 0xBF;
        if (b < l || b > h) { *rune = kRuneError; return 1; }
        r = (r << 6) | (b & 0x3Fu);
    }
    *rune = r;
    return need + 1;
}

int utf8_encode(uint32_t r, uint8_t out[4]) {
    if (r > 0x10FFFF || (r >= 0xD800 && r <= 0xDFFF)) r = kRuneError;
    if (r < 0x80) { out[0] = (uint8_t)r; return 1; }
    if (r < 0x800) { out[0] = (uint8_t)(0xC0 | (r >> 6)); out[1] = (uint8_t)(0x80 | (r & 0x3F)); return 2; }
    if (r < 0x10000) {
        out[0] = (uint8_t)(0xE0 | (r >> 12)); out[1] = (uint8_t)(0x80 | ((r >> 6) & 0x3F));
        out[2] = (uint8_t)(0x80 | (r & 0x3F));
        return 3;
    }
    out[0] = (uint8_t)(0xF0 | (r >> 18)); out[1] = (uint8_t)(0x80 | ((r >> 12) & 0x3F));
    out[2] = (uint8_t)(0x80 | ((r >> 6) & 0x3F)); out[3] = (uint8_t)(0x80 | (r & 0x3F));
    return 4;
}

namespace {
struct CaseRange { uint32_t lo, hi; int32_t delta; uint32_t step; };
const CaseRange kLower[] = {
#include "unicode_lower.inc"
};
constexpr size_t kNLower = sizeof(kLower) / sizeof(kLower[0]);
}  // namespace

uint32_t rune_lower(uint32_t r) {
    if (r < 0x80) return (r >= 'A' && r <= 'Z') ? r + 32 : r;
    const CaseRange *it = std::lower_bound(kLower, kLower + kNLower, r,
                                           [](const CaseRange &c, uint32_t v) { return c.hi < v; });
    if (it != kLower + kNLower && r >= it->lo && (r - it->lo) % it->step == 0) return (uint32_t)((int32_t)r + it->delta);
    return r;
}

// strings.ToLower: an all-ASCII string is lowered bytewise; otherwise every rune is decoded
// (invalid byte -> U+FFFD), mapped and re-encoded.
void to_lower(const uint8_t *s, size_t len, std::string *out) {
    bool ascii = true;
    for (size_t i = 0; i < len; i++) if (s[i] >= 0x80) { ascii = false; break; }
    if (ascii) {
        for (size_t i = 0; i < len; i++) out->push_back((char)((s[i] >= 'A' && s[i] <= 'Z') ? s[i] + 32 : s[i]));
        return;
    }
    for (size_t i = 0; i < len;) {
        uint32_t r;
        i += (size_t)utf8_decode(s + i, len - i, &r);
        uint8_t enc[4];
        int e = utf8_encode(rune_lower(r), enc);
        out->append((const char *)enc, (size_t)e);
    }
}

static void decode_all(const std::string &s, std::vector<uint32_t> *out) {
    for (size_t i = 0; i < s.size();) {
        uint32_t r;
        i += (size_t)utf8_decode((const uint8_t *)s.data() + i, s.size() - i, &r);
        out->push_back(r);
    }
}

std::string TextConfig::init(int ngram, const char *wrap0, const char *wrap1, const char *pad_s,
                             const char *const *alpha, int n_alpha) {
    if (ngram < 1 || ngram > kMaxNgram) return "nGramSize must be in 1..8";
    if (!wrap0 || !wrap1 || !pad_s || (n_alpha > 0 && !alpha)) return "null string in index description";
    n = ngram;
    wrap_start = wrap0;
    wrap_end = wrap1;
    std::vector<uint32_t> pr;
    decode_all(pad_s, &pr);
    if (pr.size() != 1) return "pad must be exactly one character for the device tokenizer";
    pad_rune = pr[0];
    // every rune for which compositeAlphabet.Has is true (pkg/alphabet/composite_alphabet.go:35-45)
    std::vector<uint32_t> members;
    alphabet.clear();
    for (int i = 0; i < n_alpha; i++) {
        if (!alpha[i]) return "null string in alphabet";
        std::string a = alpha[i];
        alphabet.push_back(a);
        if (a == "english") for (uint32_t r = 'a'; r <= 'z'; r++) members.push_back(r);
        else if (a == "numbers") for (uint32_t r = '0'; r <= '9'; r++) members.push_back(r);
        else if (a == "russian") {  // а..я, and ё is tested as е (pkg/alphabet/russian_alphabet.go:16-22)
            for (uint32_t r = 0x430; r <= 0x44F; r++) members.push_back(r);
            members.push_back(0x451);
        } else decode_all(a, &members);  // NewSimpleAlphabet([]rune(symbols))
    }
    std::sort(members.begin(), members.end());
    members.erase(std::unique(members.begin(), members.end()), members.end());
    std::memset(ascii_code, 0, sizeof(ascii_code));
    ranges.clear();
    code_rune.assign(1, 0);
    uint32_t code = 0;
    for (uint32_t r : members) {
        code++;
        code_rune.push_back(r);
        if (r < 128) { ascii_code[r] = (uint8_t)code; continue; }
        if (!ranges.empty() && ranges.back().hi + 1 == r) ranges.back().hi = r;
        else ranges.push_back(RuneRange{r, r, code});
    }
    if (code > 250) return "alphabet has more than 250 characters";
    bool pad_member = std::binary_search(members.begin(), members.end(), pad_rune);
    n_codes = code;
    if (!pad_member) { n_codes = ++code; code_rune.push_back(pad_rune); }
    bits = 1;
    while ((1u << bits) <= n_codes) bits++;
    if (bits * n > 64) return "nGramSize * log2(alphabet size) exceeds 64 bits";
    pad_code = 0;
    pad_code = pad_member ? code_of(pad_rune) : n_codes;
    std::vector<uint32_t> w0, w1;
    std::string l0, l1;
    to_lower((const uint8_t *)wrap_start.data(), wrap_start.size(), &l0);
    to_lower((const uint8_t *)wrap_end.data(), wrap_end.size(), &l1);
    decode_all(l0, &w0);
    decode_all(l1, &w1);
    if (w0.size() > (size_t)kMaxWrapRunes || w1.size() > (size_t)kMaxWrapRunes) return "wrap longer than 8 characters";
    return "";
}

uint32_t TextConfig::code_of(uint32_t r) const {
    if (r < 128) return ascii_code[r] ? ascii_code[r] : pad_code;
    size_t lo = 0, hi = ranges.size();
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (ranges[mid].hi < r) lo = mid + 1; else hi = mid;
    }
    if (lo < ranges.size() && r >= ranges[lo].lo) return ranges[lo].base + (r - ranges[lo].lo);
    return pad_code;
}

uint64_t TextConfig::key_of_term(const uint8_t *term, size_t len) const {
    uint64_t key = 0;
    int i = 0;
    for (size_t p = 0; p < len; i++) {
        if (i >= n) return 0;
        uint32_t r;
        p += (size_t)utf8_decode(term + p, len - p, &r);
        // a stored term only holds alphabet members and the pad
        uint32_t c = code_of(r);
        if (c == pad_code && r != pad_rune) return 0;
        key |= (uint64_t)c << (bits * i);
    }
    return key;
}

void tokenize_keys(const TextConfig &cfg, const uint8_t *text, size_t len, std::vector<uint64_t> *keys,
                   TokenScratch *sc) {
    keys->clear();
    // wrapTokenizer (pkg/analysis/wrap_tokenizer.go:18-20)
    sc->wrapped.assign(cfg.wrap_start);
    sc->wrapped.append((const char *)text, len);
    sc->wrapped.append(cfg.wrap_end);
    // filterTokenizer (pkg/analysis/filter_tokenizer.go:20-27)
    sc->lowered.clear();
    to_lower((const uint8_t *)sc->wrapped.data(), sc->wrapped.size(), &sc->lowered);
    const uint8_t *p = (const uint8_t *)sc->lowered.data();
    size_t b = 0, e = sc->lowered.size();
    while (b < e && p[b] == ' ') b++;
    while (e > b && p[e - 1] == ' ') e--;
    // nGramTokenizer (pkg/analysis/ngram_tokenizer.go:17-43): the early-out is in bytes, windows are in runes
    const int n = cfg.n;
    if (e - b < (size_t)n) return;
    std::vector<uint32_t> &runes = sc->runes;
    runes.clear();
    for (size_t i = b; i < e;) {
        uint32_t r;
        i += (size_t)utf8_decode(p + i, e - i, &r);
        runes.push_back(r);
    }
    const size_t R = runes.size();
    const size_t n_win = R < (size_t)n ? 1 : R - (size_t)n + 1;
    const size_t wlen = R < (size_t)n ? R : (size_t)n;
    for (size_t i = 0; i < n_win; i++) {
        bool dup = false;  // appendUnique compares the raw windows (:46-54)
        for (size_t j = 0; j < i && !dup; j++) dup = std::equal(runes.begin() + i, runes.begin() + i + wlen, runes.begin() + j);
        if (dup) continue;
        uint64_t key = 0;  // normalizeFilter (pkg/analysis/normalizer.go:21-37)
        for (size_t c = 0; c < wlen; c++) key |= (uint64_t)cfg.code_of(runes[i + c]) << (cfg.bits * (int)c);
        keys->push_back(key);
    }
}

}  // namespace sg
