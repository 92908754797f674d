#!/usr/bin/env python
"""Attribute ncu warp-stall samples to CUDA source lines.

ncu's CLI source page carries metrics only per SASS instruction; nvdisasm --print-line-info gives the
source line of every SASS instruction of the same cubin.  Both list a kernel's instructions in order, so
they are joined by position.

usage: python profiles/samples_by_line.py <src.csv from `ncu --page source --csv`> <kernel substring> [top]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "abrsimulator_b200", "lib", "libabr_b200.so")
src_csv, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40

rows = list(csv.reader(open(src_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
ix = {h: i for i, h in enumerate(rows[hi])}
data = rows[hi + 1:]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=tmp, capture_output=True)
lines = None
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin") or "-" in f:
        continue
    out = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fn, cur_line, acc = None, None, []
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            if cur_fn and kern in cur_fn and len(acc) == len(data):
                lines = acc
            cur_fn, acc, cur_line = m.group(1), [], None
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
            acc.append(cur_line)
    if cur_fn and kern in cur_fn and len(acc) == len(data):
        lines = acc
if lines is None:
    sys.exit("could not match the kernel's SASS (%d instructions) in the library" % len(data))

samples, execd = defaultdict(int), defaultdict(int)
for r, l in zip(data, lines):
    samples[l] += int(r[ix["# Samples"]] or 0)
    execd[l] += int(r[ix["Instructions Executed"]] or 0)
tot, tote = sum(samples.values()), sum(execd.values())
srcs = {}
print(f"{tot} samples, {tote} warp instructions")
for l in sorted(samples, key=lambda l: -samples[l])[:top]:
    text = ""
    if l:
        path = os.path.join(ROOT, "abrsimulator_b200", "csrc", l[0])
        if path not in srcs and os.path.exists(path):
            srcs[path] = open(path).read().splitlines()
        if path in srcs and l[1] <= len(srcs[path]):
            text = srcs[path][l[1] - 1].strip()[:90]
    print(f"{100.0 * samples[l] / tot:5.1f}% samp {100.0 * execd[l] / tote:5.1f}% instr  {l[0] if l else '?'}:{l[1] if l else 0:<4} {text}")
