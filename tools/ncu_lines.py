#!/usr/bin/env python3
"""Summarise an `ncu --page source --print-source cuda,sass --csv` export: samples / instructions / top stalls per CUDA line."""
import csv
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr, data, cur = None, [], None
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
    toti = sum(num(r[iinst]) for f, l, r in data if f == 'sg_kernels.cu')
    print("total samples", tot, "instructions (sg_kernels.cu lines)", toti)
    st = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {}
    for f, l, r in data:
        for i, h in st:
            agg[h] = agg.get(h, 0) + num(r[i])
    print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    for f, l, r in sorted(data, key=lambda x: -num(x[2][isamp]))[:top]:
        s = num(r[isamp])
        stalls = sorted(((h[6:], num(r[i])) for i, h in st), key=lambda kv: -kv[1])[:3]
        print(f"{f}:{l:>4} {100 * s / tot:5.1f}% inst {100 * num(r[iinst]) / toti:5.1f}%  {stalls}  | {r[1].strip()[:70]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
