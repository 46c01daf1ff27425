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


if __name__ == "__main__" and sys.argv[0].endswith("ncu_lines.py"):
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)


def regions(path, src_path, n_queries=65536, launches=1):
    n_queries *= launches
    """instruction / sample share of the kernel's phases (markers looked up in the source, kernel body only)"""
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
    src = open(src_path).read().split('\n')
    kstart = next(i + 1 for i, l in enumerate(src) if l.startswith('__global__') and 'sg_search_kernel' in l)

    def find(s, start=0):
        return next(i + 1 for i, l in enumerate(src) if i + 1 >= start and s in l)
    marks = [('metric/text helpers', 1), ('topk/emit/resolve', find('struct QueryCtx')), ('tma + count_group', find('shared-memory staging')),
             ('slice walker', find('struct SliceWalker')), ('(plan kernel)', find('sg_plan_kernel: steps')),
             ('load plan', kstart), ('chunk setup', find('for (uint64_t cs = 0', kstart)),
             ('count loop', find('---- count:', kstart)), ('scan', find('---- scan, segment', kstart)),
             ('results', find('5. results', kstart)), ('end', find('k best of n_parts', kstart))]
    ks = [d for d in data if d[0] == 'sg_kernels.cu']
    tot = sum(num(r[iinst]) for f, l, r in ks)
    tots = sum(num(r[isamp]) for f, l, r in ks)
    for i, (name, a) in enumerate(marks[:-1]):
        b = marks[i + 1][1] - 1
        ii = sum(num(r[iinst]) for f, l, r in ks if a <= l <= b)
        ss = sum(num(r[isamp]) for f, l, r in ks if a <= l <= b)
        print(f"{name:16s} lines {a:4d}-{b:4d}  inst {100 * ii / tot:5.1f}% ({ii / n_queries:7.0f}/query)  samples {100 * ss / tots:5.1f}%")


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[0].endswith("ncu_lines.py"):
    regions(sys.argv[1], sys.argv[3])
