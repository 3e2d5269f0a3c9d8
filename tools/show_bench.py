#!/usr/bin/env python3
"""print the key numbers of bench.py JSON lines: python tools/show_bench.py gpurun_out/*.json"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
    except Exception as e:
        print(f"{f}: unreadable ({e})")
        continue
    sch = d.get("schedule", {})
    st = sch.get("stage_timing_pass", {}).get("stage_ms", {})
    print(f"{f.split('/')[-1]:<58} {d['value']:.3e}/s {d['ms_per_step']:7.1f} ms e2e {d['e2e']['value']:.3e} frac {d['roofline']['frac']:.3f} "
          f"it {sch.get('iterations')} tail {sch.get('tail_packets')}/{sch.get('tail_ms', 0):.1f}ms | "
          + " ".join(f"{k[:6]} {v:.0f}" for k, v in st.items()))
