#!/usr/bin/env python3
"""Summarise the source page of an ncu report per source file and per source line.

usage: ncu -i REPORT.ncu-rep --page source --csv --print-source cuda,sass > src.csv
       python tools/ncu_source_summary.py src.csv [--top 40] [--kernel k_propagate]

For every source line it adds up the warp-level instructions executed, the thread-level instructions and the
stall samples, and prints (a) totals per file with the average number of active lanes per issued instruction,
(b) the source lines that issue the most warp instructions / collect the most stall samples.
"""
import argparse
import csv
import collections
import sys


def parse(path, kernel):
    rows = []
    cur_file = None
    cur_func = None
    header = None
    with open(path, newline="") as fh:
        for rec in csv.reader(fh):
            if not rec:
                continue
            if rec[0] == "File Path":
                cur_file = rec[1]
                continue
            if rec[0] == "Function Name":
                cur_func = rec[1]
                continue
            if rec[0] == "Line No":
                header = rec
                continue
            if header is None or rec[0] == "":
                continue  # SASS rows under a source line (already aggregated into the line's row)
            if kernel and (cur_func is None or kernel not in cur_func):
                continue
            d = dict(zip(header[4:], rec[4:]))
            try:
                rows.append(
                    dict(
                        file=cur_file.split("/")[-1],
                        line=int(rec[0]),
                        src=rec[1].strip(),
                        samples=int(d["# Samples"]),
                        inst=int(d["Instructions Executed"]),
                        tinst=int(d["Thread Instructions Executed"]),
                        stalls={k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit()},
                    )
                )
            except (KeyError, ValueError):
                continue
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--kernel", default="k_propagate")
    a = ap.parse_args()
    rows = parse(a.csv, a.kernel)
    tot_inst = sum(r["inst"] for r in rows) or 1
    tot_samp = sum(r["samples"] for r in rows) or 1
    per_file = collections.defaultdict(lambda: [0, 0, 0])
    for r in rows:
        f = per_file[r["file"]]
        f[0] += r["inst"]
        f[1] += r["tinst"]
        f[2] += r["samples"]
    print(f"total warp instructions {tot_inst:.3e}, stall samples {tot_samp}")
    print(f"{'file':<16}{'warp inst %':>12}{'lanes/inst':>12}{'samples %':>11}")
    for name, (inst, tinst, samp) in sorted(per_file.items(), key=lambda kv: -kv[1][2]):
        print(f"{name:<16}{100 * inst / tot_inst:>12.1f}{(tinst / inst if inst else 0):>12.1f}{100 * samp / tot_samp:>11.1f}")
    stall_tot = collections.Counter()
    for r in rows:
        stall_tot.update(r["stalls"])
    print("stall reasons (all samples):", ", ".join(f"{k[6:]} {100 * v / tot_samp:.1f}%" for k, v in stall_tot.most_common(8)))
    print(f"\ntop {a.top} source lines by stall samples")
    for r in sorted(rows, key=lambda r: -r["samples"])[: a.top]:
        lanes = r["tinst"] / r["inst"] if r["inst"] else 0
        top_stall = max(r["stalls"].items(), key=lambda kv: kv[1])[0][6:] if r["stalls"] else "-"
        print(f"{r['file']:<14}{r['line']:>5} samp {100 * r['samples'] / tot_samp:5.1f}% inst {100 * r['inst'] / tot_inst:5.1f}% "
              f"lanes {lanes:5.1f} {top_stall:<10} | {r['src'][:90]}")


if __name__ == "__main__":
    sys.exit(main())
