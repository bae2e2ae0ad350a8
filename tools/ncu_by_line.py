"""Attribute the warp-stall samples of an ncu source page to CUDA source lines and kernel phases.

    ncu -i x.ncu-rep --page source --csv > x_sass.csv
    cuobjdump -xelf all gecco_b200/libgecco_crf_b200.so; nvdisasm -g -c gcrf_stream.sm_100a.cubin > stream.sass
    python tools/ncu_by_line.py x_sass.csv stream.sass stream_kernelILi20ELi128ELi4EiE [min_pct]
"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, listing, kernel = sys.argv[1:4]
min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.8

# offset -> (file, line) from the nvdisasm listing of the chosen kernel
line_of = {}
cur = None
inside = False
for text in open(listing):
    if text.startswith("//---") and ".text." in text:
        inside = kernel in text
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', text)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", text)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())

rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][idx["Address"]], 16)
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
by_line = defaultdict(lambda: defaultdict(float))
tot = 0
for r in data:
    off = int(r[idx["Address"]], 16) - base
    n = int(r[idx["# Samples"]] or 0)
    tot += n
    key = line_of.get(off, ((None, 0), ""))[0]
    by_line[key]["samples"] += n
    by_line[key]["inst"] += int(r[idx["Instructions Executed"]] or 0)
    by_line[key]["wf"] += int(r[idx["L1 Wavefronts Shared"]] or 0)
    for h in reasons:
        by_line[key][h] += float(r[idx[h]] or 0)
print(f"total samples {tot}, instructions {sum(int(r[idx['Instructions Executed']] or 0) for r in data)}, "
      f"shared wavefronts {sum(int(r[idx['L1 Wavefronts Shared']] or 0) for r in data)}")
src = {}
for key in sorted(by_line, key=lambda k: (str(k[0]), k[1])):
    v = by_line[key]
    if v["samples"] < tot * min_pct / 100:
        continue
    f, ln = key
    if f not in src:
        try:
            src[f] = open(f"gecco_b200/csrc/{f}").read().split("\n")
        except OSError:
            src[f] = []
    text = src[f][ln - 1].strip()[:70] if 0 < ln <= len(src[f]) else ""
    top = sorted(((v[h], h[6:]) for h in reasons), reverse=True)[:2]
    print(f"{str(f)[:16]:16s}:{ln:4d} {100 * v['samples'] / tot:5.1f}% inst {int(v['inst']):8d} wf {int(v['wf']):8d}  "
          f"{top[0][1]}={top[0][0]:.0f} {top[1][1]}={top[1][0]:.0f} | {text}")
