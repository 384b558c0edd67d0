#!/usr/bin/env python
"""Per-source-line aggregation of an ncu source-page CSV (`ncu -i x.ncu-rep --page source --csv`).

The CSV has one row per SASS instruction (address, samples, stall reasons); the line of every instruction comes from
`nvdisasm -g` of the same cubin (`cuobjdump -xelf all file.o`), restricted to the kernel's `.text` section:

    python tools/ncu_lines.py sass.csv disasm_of_kernel.txt [--top 60] [--source-root torch-assimilate_b200/csrc]
"""
import argparse
import collections
import csv
import os
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("disasm")
    ap.add_argument("--top", type=int, default=60)
    ap.add_argument("--source-root", default=None)
    args = ap.parse_args()
    line_of, cur = {}, ("?", 0)
    for ln in open(args.disasm):
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(\S+)", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(args.csv)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    base = int(body[0][0], 16)
    ci = {c: hdr.index(c) for c in hdr}
    stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    agg = collections.defaultdict(lambda: collections.Counter())
    ops = collections.defaultdict(lambda: collections.Counter())
    tot = collections.Counter()
    for r in body:
        key = line_of.get(int(r[0], 16) - base, ("?", 0))
        smp, ins = float(r[ci["# Samples"]] or 0), float(r[ci["Instructions Executed"]] or 0)
        agg[key]["samples"] += smp; agg[key]["inst"] += ins
        tot["samples"] += smp; tot["inst"] += ins
        for s in stalls:
            v = float(r[ci[s]] or 0)
            agg[key][s] += v; tot[s] += v
        ops[key][r[ci["Source"]].split()[0] if r[ci["Source"]].split() else "?"] += smp
    print("total samples {0:.0f}  warp instructions {1:.0f}".format(tot["samples"], tot["inst"]))
    print("stall mix: " + ", ".join("{0} {1:.1f}%".format(s[6:], 100 * tot[s] / tot["samples"])
                                    for s in sorted(stalls, key=lambda s: -tot[s])[:8]))
    src = {}
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:args.top]:
        text = ""
        if args.source_root:
            p = os.path.join(args.source_root, key[0])
            if p not in src and os.path.exists(p):
                src[p] = open(p).read().splitlines()
            if p in src and 0 < key[1] <= len(src[p]):
                text = src[p][key[1] - 1].strip()[:90]
        top_st = ", ".join("{0} {1:.0f}%".format(s[6:], 100 * a[s] / max(a["samples"], 1))
                           for s in sorted(stalls, key=lambda s: -a[s])[:3])
        top_op = ", ".join("{0} {1:.0f}".format(o, n) for o, n in ops[key].most_common(2))
        print("{0:5.1f}% smp {1:5.1f}% inst  {2}:{3:<4d} [{4}] ({5}) | {6}".format(
            100 * a["samples"] / tot["samples"], 100 * a["inst"] / tot["inst"], key[0], key[1], top_st, top_op, text))


if __name__ == "__main__":
    main()
