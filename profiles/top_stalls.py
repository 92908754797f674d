#!/usr/bin/env python
"""Summarise an ncu source page (sass) CSV: samples per address bucket and the top instructions.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python profiles/top_stalls.py src.csv [bucket]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 32
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
ix = {h: i for i, h in enumerate(rows[hi])}
data = rows[hi + 1:]
S = lambda r: int(r[ix['# Samples']] or 0)
E = lambda r: int(r[ix['Instructions Executed']] or 0)
tot = sum(S(r) for r in data)
totE = sum(E(r) for r in data)
print(f"kernel: {rows[0][1][:90]}")
print(f"total samples {tot}, warp instructions executed {totE}, SASS instructions {len(data)}")
print("--- samples / executed per block of %d instructions (first opcode of interest shown)" % bucket)
for b in range(0, len(data), bucket):
    blk = data[b:b + bucket]
    s, e = sum(S(r) for r in blk), sum(E(r) for r in blk)
    ops = [r[ix['Source']].split()[0 if not r[ix['Source']].strip().startswith('@') else 1] for r in blk]
    key = [o for o in ops if any(k in o for k in ('LDG', 'STG', 'MUFU', 'CALL', 'BRA', 'DSETP', 'IMAD.HI', 'LDS', 'SHFL', 'BAR'))]
    print(f"{b:5d} {100.0 * s / max(tot, 1):6.1f}% samples {100.0 * e / max(totE, 1):6.1f}% instr  thr={blk[0][ix['Avg. Threads Executed']]:>4}  {' '.join(key[:10])}")
print("--- top instructions")
for r in sorted(data, key=lambda r: -S(r))[:25]:
    print(f"{S(r):6d} {E(r):9d} thr={r[ix['Avg. Threads Executed']]:>3} {r[ix['Address']][-5:]} {r[ix['Source']][:90]}")
