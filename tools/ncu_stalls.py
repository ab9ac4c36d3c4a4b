#!/usr/bin/env python3
"""Per CUDA source line: stall samples by reason (ncu --page source --csv joined with nvdisasm -g line info).

usage: tools/ncu_stalls.py <report.ncu-rep> <object-or-so containing the kernel> <kernel-name-substring> [top]
"""
import collections
import csv
import io
import pathlib
import re
import subprocess
import sys
import tempfile

rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = pathlib.Path(tempfile.mkdtemp())
subprocess.run(["cuobjdump", "-xelf", "all", str(pathlib.Path(obj).resolve())], cwd=tmp, capture_output=True)
linemap = {}
for cub in tmp.glob("*.cubin"):
    txt = subprocess.run(["nvdisasm", "-g", "-c", str(cub)], capture_output=True, text=True).stdout
    cur_line, active = None, False
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
ia, ie = hdr.index("Address"), hdr.index("Instructions Executed")
reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
per_line = collections.defaultdict(collections.Counter)
inst_line = collections.Counter()
tot_reason = collections.Counter()
base = None
for r in rows[hi + 1:]:
    if len(r) <= ie or not re.fullmatch(r"(0x)?[0-9a-fA-F]+", r[ia]):
        continue
    addr = int(r[ia], 16)
    if base is None:
        base = addr
    line, _ = linemap.get(addr - base, (("?", 0), "?"))
    inst_line[line] += int(float(r[ie] or 0))
    for i, h in reasons:
        v = int(float(r[i] or 0))
        if v:
            per_line[line][h] += v
            tot_reason[h] += v
tot = sum(tot_reason.values()) or 1
print("stall samples by reason:", ", ".join(f"{h[6:]}={100 * v / tot:.1f}%" for h, v in tot_reason.most_common(10)))
byfile = collections.Counter()
for line, c in per_line.items():
    byfile[line[0] if line else "?"] += sum(c.values())
print("by file:", ", ".join(f"{f}={100 * v / tot:.1f}%" for f, v in byfile.most_common(8)))
print("--- lines by stall samples: share, inst share, reasons")
ti = sum(inst_line.values()) or 1
for line, c in sorted(per_line.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    s = sum(c.values())
    print(f"{100 * s / tot:6.2f}% {100 * inst_line[line] / ti:6.2f}%  {line[0]}:{line[1]:<5d} " +
          " ".join(f"{h[6:]}={100 * v / s:.0f}%" for h, v in c.most_common(4)))
