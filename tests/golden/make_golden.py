#!/usr/bin/env python3
"""Regenerates the committed fixtures from the read-only reference checkout.

These are DATA files of the reference's own test-suite (pkg/suggest/testdata), copied verbatim so
that the GPU box (which has no /root/reference) can run configuration #1 of BASELINE.json and the
on-disk index checks:
  cars.dict        pkg/suggest/testdata/cars.dict          (5,066 lines, one entry per line)
  cars.hd, cars.dl pkg/suggest/testdata/db/cars.{hd,dl}    (index v5.1 written by the Go indexer)
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
    for name in ("cars.hd", "cars.dl"):
        shutil.copyfile(os.path.join(REF, "db", name), os.path.join(HERE, name))
    print("fixtures refreshed")


if __name__ == "__main__":
    main()
