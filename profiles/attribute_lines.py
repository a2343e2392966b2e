#!/usr/bin/env python
"""Attribute an ncu --page source (SASS) export to CUDA source lines / device functions.

usage: attribute_lines.py <ncu source csv> <nvdisasm -g -c listing> <mangled kernel name>
The nvdisasm listing carries `//## File "...", line N` markers; instruction order is the same
as in ncu's SASS view, so the two are joined positionally.
"""
import csv, re, sys, collections

src_csv, dis, kernel = sys.argv[1:4]
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(f".text.{kernel}:"))
tags = []
cur = ("?", 0, "")
for l in lines[start + 1:]:
    if l.startswith("//---") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), l)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        tags.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ie = hdr.index("Instructions Executed")
te = hdr.index("Thread Instructions Executed")
sm = hdr.index("# Samples")
body = rows[2:]
assert len(body) == len(tags), (len(body), len(tags))

def region(f, ln):
    if f == "tb_neighbor.cuh":
        for name, lo, hi in REGIONS:
            if lo <= ln <= hi:
                return name
    if f == "tb_int.cuh":
        return "tb_int.cuh (arith helpers)"
    return f

# function line ranges in tb_neighbor.cuh, found by scanning the source
import os
src = open(os.path.join(os.path.dirname(__file__), "..", "ternary_birch_b200", "csrc", "tb_neighbor.cuh")).read().split("\n")
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"__device__ __forceinline__ \S+.*?\b(\w+)\(", l)
    if m and not l.strip().startswith("//"):
        marks.append((m.group(1), i))
REGIONS = [(n, lo, (marks[j + 1][1] - 1 if j + 1 < len(marks) else 10 ** 9)) for j, (n, lo) in enumerate(marks)]

agg = collections.defaultdict(lambda: [0, 0, 0, 0])
sticky = "prologue"
for r, (f, ln, raw) in zip(body, tags):
    # -lineinfo records only the innermost inlined location; instructions from the arithmetic
    # helpers (tb_int.cuh, CUDA intrinsics headers) are charged to the last tb_neighbor.cuh /
    # tb_kernels.cuh region seen in instruction order (the code is laid out phase by phase)
    if f in ("tb_neighbor.cuh", "tb_kernels.cuh"):
        sticky = region(f, ln)
    key = sticky
    a = agg[key]
    a[0] += int(r[ie]); a[1] += int(r[te]); a[2] += int(r[sm]); a[3] += 1
tot = sum(a[0] for a in agg.values())
tots = sum(a[2] for a in agg.values())
print(f"{'region':40s} {'static':>7s} {'warp-inst':>14s} {'share':>6s} {'lanes':>6s} {'samples':>7s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:40s} {a[3]:7d} {a[0]:14d} {100*a[0]/tot:5.1f}% {a[1]/max(a[0],1):6.1f} {100*a[2]/max(tots,1):6.1f}%")
