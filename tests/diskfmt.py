"""Test-side writer of the reference's on-disk index format (`<name>.hd` gob header + `<name>.dl` posting lists),
written from pkg/index/indexer_writer.go:50-63,148-166 and the published RoaringBitmap portable format.
Used to feed sg_index_open_disk / sg_host_index_open_disk with indexes the reference could have written."""
import os
import struct

import numpy as np


def gob_uint(v):
    if v < 128:
        return bytes([v])
    b = v.to_bytes((v.bit_length() + 7) // 8, "big")
    return bytes([256 - len(b)]) + b


def gob_header(terms, indices):
    """header{Version, Indices, Terms} body as encoding/gob lays it out (type definitions omitted: the reader skips them)"""
    body = gob_uint(1) + gob_uint(4) + b"v5.1" + gob_uint(1) + gob_uint(indices) + gob_uint(1) + gob_uint(len(terms))
    for term, indice, size, pos, length in terms:
        rec = gob_uint(1) + gob_uint(len(term)) + term
        last = 0
        for f, v in ((1, indice), (2, size), (3, pos), (4, length)):
            if v:
                rec += gob_uint(f - last) + gob_uint(v)
                last = f
        body += rec + gob_uint(0)
    body += gob_uint(0)
    msg = gob_uint(2 * 64) + body  # type id 64 as a gob int
    return gob_uint(len(msg)) + msg


def roaring_blob(values, runs=False):
    """RoaringBitmap portable serialisation written from the published format description"""
    values = np.unique(np.asarray(values, dtype=np.uint32))
    keys = np.unique(values >> 16)
    conts = [(int(k), (values[(values >> 16) == k] & 0xFFFF).astype(np.uint16)) for k in keys]
    size = len(conts)
    out = b""
    run_flags = []
    bodies = []
    for k, lows in conts:
        as_runs = []
        start = prev = int(lows[0])
        for v in lows[1:]:
            v = int(v)
            if v != prev + 1:
                as_runs.append((start, prev - start))
                start = v
            prev = v
        as_runs.append((start, prev - start))
        use_run = runs and 2 + 4 * len(as_runs) < min(2 * len(lows), 8192)
        run_flags.append(use_run)
        if use_run:
            bodies.append(struct.pack("<H", len(as_runs)) + b"".join(struct.pack("<HH", s, l) for s, l in as_runs))
        elif len(lows) > 4096:
            words = np.zeros(1024, dtype=np.uint64)
            for v in lows:
                words[int(v) >> 6] |= np.uint64(1) << np.uint64(int(v) & 63)
            bodies.append(words.tobytes())
        else:
            bodies.append(lows.astype("<u2").tobytes())
    if any(run_flags):
        out += struct.pack("<I", 12347 | ((size - 1) << 16))
        bm = bytearray((size + 7) // 8)
        for i, f in enumerate(run_flags):
            if f:
                bm[i // 8] |= 1 << (i % 8)
        out += bytes(bm)
    else:
        out += struct.pack("<II", 12346, size)
    for (k, lows) in conts:
        out += struct.pack("<HH", k, len(lows) - 1)
    if not any(run_flags) or size >= 4:
        pos = len(out) + 4 * size
        for b in bodies:
            out += struct.pack("<I", pos)
            pos += len(b)
    return out + b"".join(bodies)


def write_index(ox, directory, name):
    """Serialise every (segment, term) list of an OracleIndex the way Writer.Commit does (indexer_writer.go:88-145):
    codec by length (pkg/index/codec.go:76-88: VB up to 65, skipping(64) up to 256, roaring above), terms in one header.
    VB / skipping bytes come from the oracle's encoders (pinned to cars.dl), roaring from roaring_blob()."""
    from oracle import oracle as O
    terms, dl = [], b""
    for seg, term, ids in ox.iter_lists():
        ids = np.unique(np.asarray(ids, dtype=np.uint32))
        n = len(ids)
        if n <= 65:
            blob = O.encode(O.CODEC_VB, ids, 64)
        elif n <= 256:
            blob = O.encode(O.CODEC_SKIPPING, ids, 64)
        else:
            blob = roaring_blob(ids, runs=(len(terms) % 2 == 1))
        terms.append((term, seg, len(blob), len(dl), n))
        dl += bytes(blob)
    os.makedirs(directory, exist_ok=True)
    with open(os.path.join(directory, name + ".hd"), "wb") as f:
        f.write(gob_header(terms, ox.segments))
    with open(os.path.join(directory, name + ".dl"), "wb") as f:
        f.write(dl)
    return len(terms), sum(1 for t in terms if t[4] > 256)
