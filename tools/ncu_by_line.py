#!/usr/bin/env python
"""Attribute an ncu capture to source lines / device functions.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
    nvdisasm -g -c kernels.cubin > k.sass
    python tools/ncu_by_line.py sass.csv k.sass <mangled-kernel-substring> <source.cu>

Joins the per-instruction counters of the ncu SASS page with nvdisasm's line markers (same instruction
order) and prints executed warp instructions and stall samples per device function and per hot line.
"""
import csv
import re
import sys
from collections import defaultdict


def sass_lines(path, kernel):
    """[(opcode text, source line)] of one kernel, in program order."""
    out, cur, active = [], 0, False
    for ln in open(path, errors="replace"):
        if ln.startswith(".text."):
            active = kernel in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "[^"]*", line (\d+)', ln)
        if m:
            if "inlined at" not in ln:
                cur = int(m.group(1))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((m.group(2).strip(), cur))
    return out


def function_ranges(src):
    """(first line, name) of every __device__/__global__ function in the .cu file."""
    fr, lines = [], open(src).read().split("\n")
    for i, l in enumerate(lines, 1):
        m = re.search(r"\b([a-z_0-9]+)\s*\(", l)
        if ("__device__" in l or "__global__" in l or (i > 1 and ("__device__" in lines[i - 2] or "__global__" in lines[i - 2]) and "(" in l)) and m:
            if "asm" in l:
                continue
            fr.append((i, m.group(1)))
    return fr


def main():
    ncu_csv, k_sass, kernel, src = sys.argv[1:5]
    rows = list(csv.reader(open(ncu_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ci, cs, cl = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    sl = sass_lines(k_sass, kernel)
    print(f"ncu instructions: {len(body)}  nvdisasm instructions: {len(sl)}")
    n = min(len(body), len(sl))
    fr = function_ranges(src)

    def fn_of(line):
        name = "?"
        for a, nm in fr:
            if a <= line:
                name = nm
        return name

    by_fn, by_line, by_op = defaultdict(lambda: [0, 0, 0]), defaultdict(lambda: [0, 0, 0]), defaultdict(int)
    stalls_fn = defaultdict(lambda: defaultdict(int))
    tot = [0, 0, 0]
    for i in range(n):
        r, (op, line) = body[i], sl[i]
        ex, sm, wf = int(r[ci] or 0), int(r[cs] or 0), int(r[cl] or 0)
        f = fn_of(line)
        for d, k in ((by_fn, f), (by_line, line)):
            d[k][0] += ex; d[k][1] += sm; d[k][2] += wf
        tot[0] += ex; tot[1] += sm; tot[2] += wf
        by_op[op.split()[0].split(".")[0] if not op.startswith("@") else op.split()[1].split(".")[0]] += ex
        for c in stall_cols:
            v = int(r[c] or 0)
            if v:
                stalls_fn[f][hdr[c]] += v
    print(f"total warp-instructions {tot[0]:.3e}  samples {tot[1]}  smem wavefronts {tot[2]:.3e}")
    print("\nper device function: inst%  samples%  smem-wavefront%   top stalls")
    for f, (ex, sm, wf) in sorted(by_fn.items(), key=lambda kv: -kv[1][1]):
        top = sorted(stalls_fn[f].items(), key=lambda kv: -kv[1])[:4]
        print(f"  {f:22s} {100*ex/tot[0]:6.2f} {100*sm/max(1,tot[1]):6.2f} {100*wf/max(1,tot[2]):6.2f}   " +
              ", ".join(f"{k[6:]}={100*v/max(1,sm):.0f}%" for k, v in top))
    print("\nhottest source lines: line inst% samples%")
    srcl = open(src).read().split("\n")
    for line, (ex, sm, wf) in sorted(by_line.items(), key=lambda kv: -kv[1][1])[:28]:
        print(f"  {line:5d} {100*ex/tot[0]:6.2f} {100*sm/max(1,tot[1]):6.2f}  {srcl[line-1].strip()[:110] if 0 < line <= len(srcl) else ''}")
    print("\nopcode mix (inst%):", ", ".join(f"{k}={100*v/tot[0]:.1f}" for k, v in sorted(by_op.items(), key=lambda kv: -kv[1])[:18]))


if __name__ == "__main__":
    main()
