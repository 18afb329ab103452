#!/usr/bin/env python
"""Aggregate an ncu `--page source --print-source sass --csv` export per CUDA source line.

ncu's CSV source page only shows the kernel's own file, not inlined headers, so this joins the
per-SASS-instruction rows of the export (in program order) with `nvdisasm --print-line-info`
of the same cubin (also in program order) and sums samples / executed instructions per
(file, line).  Usage:

    cuobjdump -xelf all wmix_b200/libwmix_b200.so          # -> wmixb.sm_100a.cubin
    nvdisasm --print-line-info wmixb.sm_100a.cubin > all.sass
    ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv
    python tools/ncu_by_line.py src.csv all.sass _Z9ns_kernelILi256E [--top 40] [--ranges a-b,c-d]
"""
import argparse
import collections
import csv
import re


def sass_lines(path, mangled_prefix):
    """[(offset, file, line, text)] for the first .text section whose name starts with the prefix."""
    out, on, cur = [], False, ("?", 0)
    for ln in open(path, errors="replace"):
        if ln.startswith(".text."):
            on = ln.startswith(".text." + mangled_prefix)
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            # the innermost location is printed first; "inlined at" lines follow it
            if "inlined at" not in ln:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), cur[0], cur[1], m.group(2).strip()))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("sass")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--file", default=None, help="only lines of this file in the range report")
    ap.add_argument("--ranges", default="", help="comma separated a-b line ranges to total (of --file)")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    # the export repeats the table per captured launch: keep the first one
    data = []
    for r in rows[hdr_i + 1:]:
        if r and r[0] in ("Address", "Kernel Name"):
            break
        if len(r) == len(hdr):
            data.append(r)
    sass = sass_lines(a.sass, a.kernel)
    if len(sass) != len(data):
        print("warning: %d SASS instructions in the cubin vs %d rows in the export" % (len(sass), len(data)))
    n = min(len(sass), len(data))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.defaultdict(lambda: collections.Counter())
    tot = collections.Counter()
    for k in range(n):
        r = data[k]
        key = (sass[k][1], sass[k][2])
        c = agg[key]
        c["samples"] += int(r[col["# Samples"]] or 0)
        c["inst"] += int(r[col["Instructions Executed"]] or 0)
        c["thread_inst"] += int(r[col["Thread Instructions Executed"]] or 0)
        c["sass"] += 1
        for s in stall_cols:
            c[s] += int(r[col[s]] or 0)
        tot["samples"] += int(r[col["# Samples"]] or 0)
        tot["inst"] += int(r[col["Instructions Executed"]] or 0)
    print("total samples %d, warp instructions %d, SASS lines %d" % (tot["samples"], tot["inst"], n))
    print("%-22s %8s %6s %10s %6s %5s  top stalls" % ("file:line", "samples", "%", "inst", "%", "sass"))
    for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[: a.top]:
        st = sorted(((c[s], s[6:]) for s in stall_cols), reverse=True)[:3]
        print("%-22s %8d %6.2f %10d %6.2f %5d  %s" % ("%s:%d" % key, c["samples"], 100.0 * c["samples"] / max(1, tot["samples"]),
                                                       c["inst"], 100.0 * c["inst"] / max(1, tot["inst"]), c["sass"],
                                                       " ".join("%s=%d" % (s, v) for v, s in st if v)))
    if a.ranges:
        print("\nranges of %s:" % a.file)
        for rg in a.ranges.split(","):
            lo, hi = [int(x) for x in rg.split("-")]
            c = collections.Counter()
            for (f, l), v in agg.items():
                if (a.file is None or f == a.file) and lo <= l <= hi:
                    c.update(v)
            st = sorted(((c[s], s[6:]) for s in stall_cols), reverse=True)[:4]
            print("  %5d-%-5d samples %6.2f%%  inst %6.2f%% (%d, sass %d)  %s" % (
                lo, hi, 100.0 * c["samples"] / max(1, tot["samples"]), 100.0 * c["inst"] / max(1, tot["inst"]), c["inst"], c["sass"],
                " ".join("%s=%.1f%%" % (s, 100.0 * v / max(1, tot["samples"])) for v, s in st if v)))


if __name__ == "__main__":
    main()
