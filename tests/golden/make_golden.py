#!/usr/bin/env python3
"""Regenerates the committed fixtures from the read-only reference checkout.

These are DATA files of the reference's own test-suite (pkg/suggest/testdata), copied verbatim so
that the GPU box (which has no /root/reference) can run configuration #1 of BASELINE.json and the
on-disk index checks:
  cars.dict        pkg/suggest/testdata/cars.dict          (5,066 lines, one entry per line)
  cars.hd, cars.dl pkg/suggest/testdata/db/cars.{hd,dl}    (index v5.1 written by the Go indexer)
  words.dict       pkg/suggest/testdata/words.dict         (235,886 English words, one per line: the real-language dictionary of
                   bench.py's config3 sweep and of the DESIGN.md measurements)
  roaring_samples.npz  twelve roaring-bitmap posting lists (> 256 ids) cut out of db/words.dl together with
                   the ids a rebuild of words.dict gives for the same (segment, term)
  lm/1-gm, 2-gm, 3-gm, test.lm   pkg/lm/testdata/fixtures (Google n-gram text of the "i am sam" corpus and the binary
                   model the reference's build-lm wrote from it); the expected scores live in tests/test_lm_oracle.py with
                   their pkg/lm/*_test.go line numbers
No reference source code is copied.
"""
import os
import shutil
import sys

REF = "/root/reference/pkg/suggest/testdata"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    if not os.path.isdir(REF):
        sys.exit("reference checkout not present; fixtures are already committed")
    shutil.copyfile(os.path.join(REF, "cars.dict"), os.path.join(HERE, "cars.dict"))
    shutil.copyfile(os.path.join(REF, "words.dict"), os.path.join(HERE, "words.dict"))
    for name in ("cars.hd", "cars.dl"):
        shutil.copyfile(os.path.join(REF, "db", name), os.path.join(HERE, name))
    os.makedirs(os.path.join(HERE, "lm"), exist_ok=True)
    for name in ("1-gm", "2-gm", "3-gm", "test.lm"):
        shutil.copyfile(os.path.join("/root/reference/pkg/lm/testdata/fixtures", name), os.path.join(HERE, "lm", name))
    roaring_samples()
    print("fixtures refreshed")


def roaring_samples():
    import numpy as np
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    sys.path.insert(0, os.path.dirname(HERE))
    from gobhdr import decode_header
    from oracle import oracle as O
    lines = open(os.path.join(REF, "words.dict"), "rb").read().split(b"\n")[:-1]
    ox = O.OracleIndex(3, ("^", "$"), "$", ("english", "numbers", "$^")).add_docs(lines)
    _, _, terms = decode_header(open(os.path.join(REF, "db", "words.hd"), "rb").read())
    dl = open(os.path.join(REF, "db", "words.dl"), "rb").read()
    big = sorted((t for t in terms if t[4] > 256), key=lambda t: t[4])
    pick = [big[0], big[1], big[len(big) // 4], big[len(big) // 2], big[3 * len(big) // 4], big[-3], big[-2], big[-1]]
    pick += [t for t in big if int.from_bytes(dl[t[3]:t[3] + 2], "little") == 12347][:4]
    out = {"n": np.array(len(pick))}
    for i, (term, indice, size, pos, length) in enumerate(pick):
        out[f"blob{i}"] = np.frombuffer(dl[pos:pos + size], dtype=np.uint8)
        out[f"ids{i}"] = ox.get_list(indice, term)
    np.savez_compressed(os.path.join(HERE, "roaring_samples.npz"), **out)


if __name__ == "__main__":
    main()
