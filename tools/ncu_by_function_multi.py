#!/usr/bin/env python
"""Like ncu_by_line.py, but for a kernel whose code comes from several source files: attributes every SASS instruction
to the device function of the FILE its line marker names.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
    nvdisasm -g -c kernels.cubin > k.sass
    python tools/ncu_by_function_multi.py sass.csv k.sass <mangled-kernel-substring> <src1> [<src2> ...]
"""
import bisect
import collections
import csv
import os
import re
import sys


def functions(src):
    res, lines = [], open(src).read().split("\n")
    for i, l in enumerate(lines, 1):
        m = re.search(r"\b([a-z_0-9]+)\s*\(", l)
        if ("__device__" in l or "__global__" in l or (i > 1 and ("__device__" in lines[i - 2] or "__global__" in lines[i - 2]) and "(" in l)) and m and "asm" not in l:
            res.append((i, m.group(1)))
    return res


def main():
    ncu_csv, k_sass, kernel = sys.argv[1:4]
    F = {os.path.basename(p): functions(p) for p in sys.argv[4:]}
    rows = list(csv.reader(open(ncu_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall = {h[6:]: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
    cur_file, cur, where, active = None, 0, [], False
    for ln in open(k_sass, errors="replace"):
        if ln.startswith(".text."):
            active = kernel in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]*)", line (\d+)', ln)
        if m:
            if "inlined at" not in ln:
                cur_file, cur = os.path.basename(m.group(1)), int(m.group(2))
            continue
        if re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln):
            where.append((cur_file, cur))
    print(f"ncu instructions: {len(body)}  nvdisasm instructions: {len(where)}")
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    for r, (f, l) in zip(body, where):
        name = str(f)
        if f in F:
            ks = [x[0] for x in F[f]]
            k = bisect.bisect_right(ks, l) - 1
            name = f.split(".")[0].replace("raceline_", "") + ":" + (F[f][k][1] if k >= 0 else "?")
        a = agg[name]
        a[0] += int(r[ci]); a[1] += int(r[cs])
        for k2, i2 in stall.items():
            a[2][k2] += int(r[i2] or 0)
    ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
    print(f"total warp-instructions {ti:.3e}  samples {ts}")
    print("function (file:name)                     inst%  samples%   top stalls")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
        tot = sum(v[2].values()) or 1
        top = ", ".join(f"{n}={c * 100 // tot}%" for n, c in v[2].most_common(3))
        print(f"  {k:38s} {v[0] / ti * 100:6.2f} {v[1] / ts * 100:7.2f}   {top}")


if __name__ == "__main__":
    main()
