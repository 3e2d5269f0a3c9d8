#!/usr/bin/env python3
"""one line per bench JSON: python tools/show_tune.py gpurun_out/r2c_tune_*.json"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable:", e)
        continue
    st = d["schedule"]["stage_timing_pass"]
    sm = st["stage_ms"]
    print(f"{f.split('/')[-1]:42s} ms/step {d['ms_per_step']:7.1f} e2e {d['e2e']['ms_per_step']:7.1f} frac {d['roofline']['frac']:.4f} | other {sm['other']:5.1f} "
          f"thin {sm['rpkt_thin']:6.1f} thick {sm['rpkt_thick']:5.1f} ma {sm['macroatom']:5.1f} tail {st['tail_ms']:4.1f} total {st['total_ms']:6.1f} "
          f"| it {d['schedule']['iterations']} launches {d['schedule']['launches']} int/pkt {d.get('workload_details', d['config'])['interactions_per_step_per_gpu'] / d['config']['packets_per_gpu']:.3f}")
