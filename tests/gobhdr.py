"""Minimal decoder for the gob-encoded index header `<name>.hd` (test-side helper).

Layout written by pkg/index/indexer_writer.go:50-63,148-166:
  header{Version string, Indices uint32, Terms []termDescription}
  termDescription{Term string, Indice, PostingListBytesSize, PostingListPosition, PostingListLen uint32}
gob stream = messages [uvarint length][int type id][body]; negative ids are type definitions and are
skipped.  Struct bodies are (field delta, value)* terminated by delta 0; zero-valued fields are omitted.
"""


def _uint(b, i):
    c = b[i]
    if c < 0x80:
        return c, i + 1
    n = 256 - c
    return int.from_bytes(b[i + 1:i + 1 + n], "big"), i + 1 + n


def _int(b, i):
    u, i = _uint(b, i)
    return (~(u >> 1) if u & 1 else (u >> 1)), i


def decode_header(data):
    i = 0
    while True:
        ln, i = _uint(data, i)
        end = i + ln
        tid, j = _int(data, i)
        if tid < 0:
            i = end
            continue
        break
    version, indices, terms = "", 0, []
    field = -1
    while True:
        d, j = _uint(data, j)
        if d == 0:
            break
        field += d
        if field == 0:
            n, j = _uint(data, j)
            version = data[j:j + n].decode()
            j += n
        elif field == 1:
            indices, j = _uint(data, j)
        elif field == 2:
            cnt, j = _uint(data, j)
            for _ in range(cnt):
                rec = [b"", 0, 0, 0, 0]
                f = -1
                while True:
                    d2, j = _uint(data, j)
                    if d2 == 0:
                        break
                    f += d2
                    if f == 0:
                        n, j = _uint(data, j)
                        rec[0] = bytes(data[j:j + n])
                        j += n
                    else:
                        rec[f], j = _uint(data, j)
                terms.append(tuple(rec))
    assert j == end, (j, end)
    return version, indices, terms
