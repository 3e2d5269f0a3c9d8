#!/usr/bin/env python3
"""DRAM traffic of every kernel family of one full-size update_packets step, from one ncu pass over ALL launches:

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \\
      --log-file gpurun_out/r2_dram.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --one-step
  python tools/ncu_dram_traffic.py gpurun_out/r2_dram.csv profiles/r2_dram_traffic.json

(three metrics = one replay pass per launch). bench.py reads the JSON for roofline.traffic."""
import collections
import csv
import json
import re
import sys

FAMILY = [("k_wf_stage<0>", "other"), ("k_wf_stage<(int)0>", "other"), ("k_wf_stage<1>", "rpkt_thin"), ("k_wf_stage<(int)1>", "rpkt_thin"),
          ("k_wf_stage<2>", "rpkt_thick"), ("k_wf_stage<(int)2>", "rpkt_thick"), ("k_wf_stage<3>", "macroatom"),
          ("k_wf_stage<(int)3>", "macroatom"), ("k_wf_refill<2>", "rpkt_thick"), ("k_wf_refill<(int)2>", "rpkt_thick"),
          ("k_wf_refill<3>", "macroatom"), ("k_wf_refill<(int)3>", "macroatom"), ("k_propagate", "history_tail"), ("k_build_", "table_build"),
          ("k_sort_", "sort_lists"), ("k_list_", "sort_lists"), ("k_wf_seed", "sort_lists"), ("k_wf_advance", "sort_lists"),
          ("k_wf_ma_swap", "sort_lists"), ("k_reset_work", "sort_lists"), ("k_aos_to_soa", "packet_convert"), ("k_soa_to_aos", "packet_convert")]


def family_of(name):
    for key, fam in FAMILY:
        if key in name:
            return fam
    return "other_kernels"


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = []
    with open(src, newline="") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    reader = csv.DictReader(lines)
    for rec in reader:
        rows.append(rec)
    agg = collections.defaultdict(lambda: {"dram_bytes_per_step": 0.0, "launches_per_step": 0, "time_ms": 0.0})
    ids = collections.defaultdict(set)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
             "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    for rec in rows:
        fam = family_of(rec["Kernel Name"])
        val = float(rec["Metric Value"].replace(",", ""))
        unit = rec["Metric Unit"]
        if rec["Metric Name"].startswith("dram__bytes"):
            agg[fam]["dram_bytes_per_step"] += val * scale.get(unit, 1.0)
        elif rec["Metric Name"].startswith("gpu__time_duration"):
            agg[fam]["time_ms"] += val * scale.get(unit, 1e-6)
            ids[fam].add(rec["ID"])
    for fam in agg:
        agg[fam]["launches_per_step"] = len(ids[fam])
    out = {"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every launch of one step ({src.split('/')[-1]}); "
                     "launch times under ncu are serialised and cold-cache", "kernels": dict(agg)}
    json.dump(out, open(dst, "w"), indent=1)
    for fam, v in sorted(agg.items(), key=lambda kv: -kv[1]["time_ms"]):
        print(f"{fam:16s} launches {v['launches_per_step']:5d} time {v['time_ms']:8.2f} ms  DRAM {v['dram_bytes_per_step'] / 1e9:8.2f} GB "
              f"({v['dram_bytes_per_step'] / max(v['time_ms'], 1e-9) / 1e6:7.1f} GB/s)")


if __name__ == "__main__":
    main()
