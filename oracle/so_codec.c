/*
 * so_codec.c — posting-list codecs and lazily decoding iterators, restated.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows:
 *   pkg/store/byte_output.go:26-38            WriteVUInt32 (7 bits per byte, LSB group first, 0x80 = more)
 *   pkg/store/byte_input.go:130-193           ReadVUInt32
 *   pkg/compression/varint.go:36-78           delta + varint
 *   pkg/compression/skipping.go:67-151        blocks of `gap`, uint16 header (len+2, bit15 = last block)
 *   pkg/compression/binary.go                 raw little-endian uint32
 *   pkg/merger/list_iterator.go:29-101        SliceIterator
 *   pkg/index/posting_list.go:16-108          VB iterator (decodes one varint per Next)
 *   pkg/index/skipping_posting_list.go:13-201 skipping iterator
 */
#include "so_internal.h"

/* ---- store ---- */
int so_read_vu32(so_input *in, uint32_t *v) {
    uint32_t r = 0;
    int shift = 0;
    for (int k = 0; k < 5; k++) {
        if (in->i >= in->len) return -1; /* ErrUInt32Overflow / unexpected EOF */
        uint8_t b = in->buf[in->i++];
        r |= (uint32_t)(b & 0x7F) << shift;
        if (!(b & 0x80)) { *v = r; return 0; }
        shift += 7;
    }
    return -1;
}

int so_read_u16(so_input *in, uint16_t *v) {
    if (in->i + 2 > in->len) return -1;
    *v = (uint16_t)(in->buf[in->i] | (in->buf[in->i + 1] << 8));
    in->i += 2;
    return 0;
}

static int64_t so_write_vu32(uint32_t v, uint8_t *out, uint64_t cap, uint64_t at) {
    int j = 0;
    uint8_t chunk[5];
    for (; v > 0x7F; j++) { chunk[j] = 0x80 | (uint8_t)(v & 0x7F); v >>= 7; }
    chunk[j] = (uint8_t)v;
    if (at + (uint64_t)j + 1 > cap) return -1;
    memcpy(out + at, chunk, (size_t)j + 1);
    return j + 1;
}

/* varIntEncode, varint.go:36-56 */
static int64_t so_varint_encode(const uint32_t *list, uint32_t n, uint32_t prev, uint8_t *out, uint64_t cap,
                                uint64_t at) {
    int64_t total = 0;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t delta = list[i] - prev;
        prev = list[i];
        int64_t w = so_write_vu32(delta, out, cap, at + (uint64_t)total);
        if (w < 0) return -1;
        total += w;
    }
    return total;
}

int64_t so_encode(int codec, int gap, const uint32_t *list, uint32_t n, uint8_t *out, uint64_t cap) {
    if (codec == SO_CODEC_VB) return so_varint_encode(list, n, 0, out, cap, 0);
    if (codec == SO_CODEC_BINARY) {
        if ((uint64_t)n * 4 > cap) return -1;
        for (uint32_t i = 0; i < n; i++) {
            out[4 * i] = (uint8_t)list[i]; out[4 * i + 1] = (uint8_t)(list[i] >> 8);
            out[4 * i + 2] = (uint8_t)(list[i] >> 16); out[4 * i + 3] = (uint8_t)(list[i] >> 24);
        }
        return (int64_t)n * 4;
    }
    if (codec == SO_CODEC_SKIPPING) { /* skippingEnc.Encode, skipping.go:67-117 */
        if ((int64_t)n < gap) return -1; /* ErrGapShouldBeGreaterThanListLen */
        uint32_t prev = 0;
        uint64_t total = 0;
        for (uint32_t i = 0; i < n; i += (uint32_t)gap) {
            uint32_t j = i + (uint32_t)gap;
            if (j > n) j = n;
            if (total + 2 > cap) return -1;
            int64_t w = so_varint_encode(list + i, j - i, prev, out, cap, total + 2);
            if (w < 0) return -1;
            prev = list[i]; /* the next block's first delta is taken against this block's first id */
            uint32_t pos = (uint32_t)w + 2;
            uint32_t packed = pos | (j == n ? 0x8000u : 0u);
            out[total] = (uint8_t)packed; out[total + 1] = (uint8_t)(packed >> 8);
            total += pos;
        }
        return (int64_t)total;
    }
    return -1;
}

int64_t so_decode(int codec, int gap, const uint8_t *in, uint64_t in_len, uint32_t *out, uint32_t n) {
    so_input r = {in, (int64_t)in_len, 0};
    if (codec == SO_CODEC_VB) { /* varIntDecode, varint.go:58-78 */
        uint32_t prev = 0, total = 0;
        while (total < n) {
            uint32_t v;
            if (so_read_vu32(&r, &v) < 0) return total;
            prev += v;
            out[total++] = prev;
        }
        return total;
    }
    if (codec == SO_CODEC_BINARY) {
        uint32_t i = 0;
        for (; i < n && r.i + 4 <= r.len; i++, r.i += 4)
            out[i] = (uint32_t)in[r.i] | ((uint32_t)in[r.i + 1] << 8) | ((uint32_t)in[r.i + 2] << 16) |
                     ((uint32_t)in[r.i + 3] << 24);
        return i;
    }
    if (codec == SO_CODEC_SKIPPING) { /* skippingEnc.Decode, skipping.go:121-146 */
        uint32_t prev = 0, i = 0;
        for (; i < n; i += (uint32_t)gap) {
            uint16_t hdr;
            if (so_read_u16(&r, &hdr) < 0) return -1;
            uint32_t j = i + (uint32_t)gap;
            if (j > n) j = n;
            uint32_t p = prev;
            for (uint32_t k = i; k < j; k++) {
                uint32_t v;
                if (so_read_vu32(&r, &v) < 0) return -1;
                p += v;
                out[k] = p;
            }
            prev = out[i];
        }
        return n;
    }
    return -1;
}

/* ---- SliceIterator ---- */
static int sl_valid(so_iter *it) { return it->index < it->size; }
static int sl_get(so_iter *it, uint32_t *v) {
    if (!sl_valid(it)) { *v = 0; return SO_IT_NOT_DEREF; }
    *v = it->slice[it->index];
    return SO_IT_OK;
}
static int sl_has_next(so_iter *it) { return it->index + 1 < it->size; }
static int sl_next(so_iter *it, uint32_t *v) {
    if (!sl_has_next(it)) { *v = 0; return SO_IT_NOT_DEREF; }
    it->index++;
    *v = it->slice[it->index];
    return SO_IT_OK;
}
static int sl_lower_bound(so_iter *it, uint32_t to, uint32_t *v) {
    if (!sl_valid(it)) { *v = 0; return SO_IT_NOT_DEREF; }
    int lo = it->index, hi = it->size; /* sort.Search over slice[index:] */
    while (lo < hi) {
        int mid = lo + (hi - lo) / 2;
        if (it->slice[mid] >= to) hi = mid; else lo = mid + 1;
    }
    if (lo >= it->size) { it->index = it->size; *v = 0; return SO_IT_NOT_DEREF; }
    it->index = lo;
    *v = it->slice[lo];
    return SO_IT_OK;
}
static int any_len(so_iter *it) { return it->size; }
static const so_iter_vt so_slice_vt = {sl_get, sl_has_next, sl_next, sl_lower_bound, any_len};

void so_iter_init_slice(so_iter *it, const uint32_t *ids, int n) {
    memset(it, 0, sizeof(*it));
    it->vt = &so_slice_vt;
    it->slice = ids;
    it->size = n;
}

/* ---- postingList (VB), posting_list.go ---- */
static int vb_get(so_iter *it, uint32_t *v) {
    if (!(it->index < it->size)) { *v = 0; return SO_IT_NOT_DEREF; }
    *v = it->current;
    return SO_IT_OK;
}
static int vb_has_next(so_iter *it) { return it->index + 1 < it->size; }
static int vb_next(so_iter *it, uint32_t *v) {
    if (!vb_has_next(it)) { *v = 0; return SO_IT_NOT_DEREF; }
    uint32_t cur;
    if (so_read_vu32(&it->in, &cur) < 0) return SO_IT_ERR;
    it->index++;
    it->current += cur;
    *v = it->current;
    return SO_IT_OK;
}
static int vb_lower_bound(so_iter *it, uint32_t to, uint32_t *v) {
    if (!(it->index < it->size)) { *v = 0; return SO_IT_NOT_DEREF; }
    if (it->current >= to) { *v = it->current; return SO_IT_OK; }
    while (vb_has_next(it)) {
        uint32_t cur;
        int rc = vb_next(it, &cur);
        if (rc != SO_IT_OK) return rc;
        if (cur >= to) { *v = cur; return SO_IT_OK; }
    }
    it->index = it->size;
    *v = 0;
    return SO_IT_NOT_DEREF;
}
static const so_iter_vt so_vb_vt = {vb_get, vb_has_next, vb_next, vb_lower_bound, any_len};

int so_iter_init_vb(so_iter *it, const uint8_t *buf, int64_t len, int list_size) {
    memset(it, 0, sizeof(*it));
    it->vt = &so_vb_vt;
    it->in.buf = buf; it->in.len = len; it->in.i = 0;
    it->size = list_size;
    it->index = 0;
    uint32_t cur;
    if (so_read_vu32(&it->in, &cur) < 0) return SO_IT_ERR;
    it->current = cur;
    return SO_IT_OK;
}

/* ---- skippingPostingList, skipping_posting_list.go ---- */
static int sk_read_skipping(so_iter *it) { /* :181-201 */
    uint16_t packed;
    if (so_read_u16(&it->in, &packed) < 0) return SO_IT_ERR;
    int position = packed & 0x7FFF, last = (packed & 0x8000) != 0; /* compression.UnpackPos */
    uint32_t cur;
    if (so_read_vu32(&it->in, &cur) < 0) return SO_IT_ERR;
    it->current = it->current_skip_value + cur;
    it->current_skip_value = it->current;
    it->next_skip_position += position;
    it->is_last_block = last;
    return SO_IT_OK;
}
static int sk_has_next(so_iter *it) { return it->index + 1 < it->size; }
static int sk_next(so_iter *it, uint32_t *v) { /* :38-67 */
    if (!sk_has_next(it)) { *v = 0; return SO_IT_NOT_DEREF; }
    if ((int)it->in.i == it->next_skip_position) {
        if (sk_read_skipping(it) != SO_IT_OK) return SO_IT_ERR;
    } else {
        uint32_t cur;
        if (so_read_vu32(&it->in, &cur) < 0) return SO_IT_ERR;
        it->current += cur;
    }
    it->index++;
    *v = it->current;
    return SO_IT_OK;
}
static int sk_lower_bound(so_iter *it, uint32_t to, uint32_t *v) { /* :71-145 */
    if (!(it->index < it->size)) { *v = 0; return SO_IT_NOT_DEREF; }
    if (it->current >= to) { *v = it->current; return SO_IT_OK; }
    int skips = 0;
    if (it->index > 0) skips = it->index / it->gap;
    while (!it->is_last_block && sk_has_next(it)) {
        so_iter prev = *it; /* includes the input cursor, i.e. prevPosition */
        it->in.i = it->next_skip_position;
        skips++;
        it->index = skips * it->gap - 1;
        if (it->index >= it->size) it->index = it->size - 2;
        uint32_t cur;
        int rc = sk_next(it, &cur);
        if (rc != SO_IT_OK) return rc == SO_IT_NOT_DEREF ? SO_IT_ERR : rc;
        if (cur < to && !it->is_last_block) continue;
        if (cur >= to) { *it = prev; break; }
    }
    while (sk_has_next(it)) {
        uint32_t cur;
        int rc = sk_next(it, &cur);
        if (rc != SO_IT_OK) return rc;
        if (cur >= to) { *v = cur; return SO_IT_OK; }
    }
    it->index = it->size;
    *v = 0;
    return SO_IT_NOT_DEREF;
}
static const so_iter_vt so_sk_vt = {vb_get, sk_has_next, sk_next, sk_lower_bound, any_len};

int so_iter_init_skipping(so_iter *it, const uint8_t *buf, int64_t len, int list_size, int gap) {
    memset(it, 0, sizeof(*it));
    it->vt = &so_sk_vt;
    it->in.buf = buf; it->in.len = len; it->in.i = 0;
    it->size = list_size;
    it->gap = gap;
    return sk_read_skipping(it);
}

/* posting_list_test.go:110-130 */
int so_posting_lower_bound_tail(int kind, int gap, const uint32_t *list, uint32_t n, uint32_t to, uint32_t *lb,
                                int *err, uint32_t *tail, uint32_t cap) {
    uint8_t *buf = (uint8_t *)malloc((size_t)n * 7 + 16);
    int64_t len = so_encode(kind, gap, list, n, buf, (uint64_t)n * 7 + 16);
    if (len < 0) { free(buf); return -1; }
    so_iter it;
    int rc = kind == SO_CODEC_VB ? so_iter_init_vb(&it, buf, len, (int)n)
                                 : so_iter_init_skipping(&it, buf, len, (int)n, gap);
    if (rc != SO_IT_OK) { free(buf); return -1; }
    rc = it.vt->lower_bound(&it, to, lb);
    *err = rc != SO_IT_OK;
    int cnt = 0;
    while (!*err) {
        uint32_t v;
        if (it.vt->get(&it, &v) != SO_IT_OK) { cnt = -1; break; }
        if ((uint32_t)cnt >= cap) { cnt = -1; break; }
        tail[cnt++] = v;
        if (!it.vt->has_next(&it)) break;
        if (it.vt->next(&it, &v) != SO_IT_OK) { cnt = -1; break; }
    }
    free(buf);
    return cnt;
}
