#!/usr/bin/env python3
"""Join `ncu --page source --csv` (per-SASS-instruction counters) with `nvdisasm -g` line info of the same cubin
and print executed warp-instructions / stall samples per CUDA source line.

usage: tools/ncu_hotlines.py <report.ncu-rep> <object-or-so containing the kernel> <kernel-name-substring> [top]
"""
import csv
import io
import re
import subprocess
import sys
import tempfile
import pathlib
import collections

rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
tmp = pathlib.Path(tempfile.mkdtemp())
subprocess.run(["cuobjdump", "-xelf", "all", str(pathlib.Path(obj).resolve())], cwd=tmp, capture_output=True)
linemap = {}
for cub in tmp.glob("*.cubin"):
    txt = subprocess.run(["nvdisasm", "-g", "-c", str(cub)], capture_output=True, text=True).stdout
    cur_fn, cur_line, active = None, None, False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            active = kname in m.group(1)
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (pathlib.Path(m.group(1)).name, int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
        if m:
            linemap[int(m.group(1), 16)] = (cur_line, m.group(2).strip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
base = None
per_line = collections.Counter()
samp_line = collections.Counter()
per_op = collections.Counter()
tot = 0
for r in rows[hi + 1:]:
    if len(r) <= ie:
        continue
    addr = int(r[ia], 16) if r[ia].startswith("0x") or re.fullmatch(r"[0-9a-fA-F]+", r[ia]) else None
    if addr is None:
        continue
    if base is None:
        base = addr
    off = addr - base
    n = int(float(r[ie] or 0))
    s = int(float(r[isamp] or 0))
    line, sass = linemap.get(off, (("?", 0), "?"))
    per_line[line] += n
    samp_line[line] += s
    per_op[sass.split()[0] if sass != "?" else "?"] += n
    tot += n
print(f"total warp instructions executed: {tot}")
tsamp = sum(samp_line.values()) or 1
print("--- by source line (inst%, stall-sample%)")
for line, n in per_line.most_common(top):
    print(f"{100 * n / tot:6.2f}%  {100 * samp_line[line] / tsamp:6.2f}%  {line[0] if line else '?'}:{line[1] if line else 0}")
print("--- by source line, sorted by stall samples (stall-sample%, inst%)")
for line, n in samp_line.most_common(top):
    print(f"{100 * n / tsamp:6.2f}%  {100 * per_line[line] / tot:6.2f}%  {line[0] if line else '?'}:{line[1] if line else 0}")
print("--- by opcode")
for op, n in per_op.most_common(25):
    print(f"{100 * n / tot:6.2f}%  {op}")
