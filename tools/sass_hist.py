#!/usr/bin/env python3
"""Execution-weighted SASS opcode histogram from an ncu report:
   ncu -i X.ncu-rep --page source --csv --print-source sass > x.csv ; python tools/sass_hist.py x.csv"""
import collections
import csv
import sys

ALU = {'IADD3', 'LOP3', 'SHF', 'ISETP', 'SEL', 'LEA', 'VIADD', 'VIMNMX', 'PRMT', 'MOV', 'PLOP3', 'BREV', 'FLO', 'POPC', 'IABS', 'P2R', 'R2P'}


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main(path, top=16):
    rows = list(csv.reader(open(path)))
    cur, ops, samples = None, {}, {}
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = r[1]; ops[cur] = collections.defaultdict(int); samples[cur] = collections.defaultdict(int); continue
        if cur and r and r[0].startswith('0x') and len(r) > 5:
            s = r[1].split()
            if not s:
                continue
            op = (s[1] if s[0].startswith('@') else s[0]).rstrip(';')
            ops[cur][op] += num(r[5]); samples[cur][op] += num(r[4])
    for fn, d in ops.items():
        tot = sum(d.values()); st = sum(samples[fn].values()) or 1
        if not tot:
            continue
        alu = sum(v for k, v in d.items() if k.split('.')[0] in ALU)
        imad = sum(v for k, v in d.items() if k.split('.')[0] == 'IMAD')
        wide = sum(v for k, v in d.items() if k.startswith('IMAD.WIDE'))
        print(f"\n{fn[:110]}\n  warp-inst {tot}  ALU-pipe {100*alu/tot:.1f}%  IMAD {100*imad/tot:.1f}% (WIDE {100*wide/tot:.1f}%)")
        for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:top]:
            print(f"    {k:24s} inst {100*v/tot:5.1f}%   stall-samples {100*samples[fn][k]/st:5.1f}%")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 16)
