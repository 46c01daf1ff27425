#!/usr/bin/env python3
"""Top CUDA source lines of an `ncu --page source --print-source cuda,sass --csv` export by executed instructions.
usage: ncu_top.py <source.csv> [n_queries] [top]"""
import csv
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main(path, nq=65536, top=45):
    rows = list(csv.reader(open(path)))
    hdr, cur, data = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if len(r) > 2 and r[0] == 'Line No':
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[2] == '-' and r[0].isdigit():
            data.append((cur, int(r[0]), r))
    isamp, iinst = hdr.index('# Samples'), hdr.index('Instructions Executed')
    tot = sum(num(r[isamp]) for f, l, r in data)
    toti = sum(num(r[iinst]) for f, l, r in data)
    print(f"samples {tot}  warp instructions {toti}  per query {toti / nq:.1f}")
    st = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {}
    for f, l, r in data:
        for i, h in st:
            agg[h[6:]] = agg.get(h[6:], 0) + num(r[i])
    print("stall samples:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    for f, l, r in sorted(data, key=lambda x: -num(x[2][iinst]))[:top]:
        s = num(r[isamp])
        stalls = sorted(((h[6:], num(r[i])) for i, h in st), key=lambda kv: -kv[1])[:2]
        print(f"{f}:{l:>4} inst/q {num(r[iinst]) / nq:7.1f} samples {100 * s / tot:5.1f}% {stalls} | {r[1].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 65536, int(sys.argv[3]) if len(sys.argv) > 3 else 45)
