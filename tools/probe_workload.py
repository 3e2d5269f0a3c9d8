#!/usr/bin/env python3
"""Tuning aid (development container): run the few-packet reference build of a bench workload with overridden
run-time inputs and print interactions per packet / macro-atom steps per activation for each timestep.
  python tools/probe_workload.py --tmin 1 --tmax 20 --ngrey 0 --aperm 7 9.5 --forb 0.1 --nts 4 --timeout 120"""
import argparse
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import configs  # noqa: E402
import gen_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="kilonova_2d_probe")
ap.add_argument("--tmin", type=float)
ap.add_argument("--tmax", type=float)
ap.add_argument("--ngrey", type=int)
ap.add_argument("--thick", type=float)
ap.add_argument("--ntimesteps", type=int)
ap.add_argument("--aperm", type=float, nargs=2)
ap.add_argument("--forb", type=float)
ap.add_argument("--fperm", type=float, nargs=2)
ap.add_argument("--mass", type=float)
ap.add_argument("--nlevels", type=int)
ap.add_argument("--transfrac", type=float)
ap.add_argument("--nts", type=int, default=4)
ap.add_argument("--timeout", type=int, default=120)
a = ap.parse_args()

cfg = configs.get(a.config)
cfg["atomic"] = dict(cfg["atomic"])
cfg["run"] = dict(cfg["run"])
cfg["model"] = dict(cfg["model"])
for k, v in (("tmin", a.tmin), ("tmax", a.tmax), ("ngrey", a.ngrey), ("thick", a.thick), ("ntimesteps", a.ntimesteps)):
    if v is not None:
        cfg["run"][k] = v
cfg["run"]["nts_run"] = a.nts
if a.aperm:
    cfg["atomic"]["A_perm_log10"] = tuple(a.aperm)
if a.fperm:
    cfg["atomic"]["f_perm_log10"] = tuple(a.fperm)
if a.forb is not None:
    cfg["atomic"]["forb_frac"] = a.forb
if a.nlevels:
    cfg["atomic"]["nlevels"] = a.nlevels
if a.transfrac:
    cfg["atomic"]["trans_frac"] = a.transfrac
if a.mass:
    cfg["model"]["mass_msun"] = a.mass

odir = os.path.join(ROOT, "oracle", "_ref", a.config, "fast")
rundir = os.path.join(odir, "probe_run")
shutil.rmtree(rundir, ignore_errors=True)
os.makedirs(rundir)
gen_inputs.write_atomic(cfg, rundir)
{"1d": gen_inputs.write_model_1d, "2d": gen_inputs.write_model_2d, "3d": gen_inputs.write_model_3d}[cfg["model"]["kind"]](cfg, rundir)
gen_inputs.write_input(cfg, rundir)
os.symlink(os.path.join(ROOT, "oracle", "_ref", "data"), os.path.join(rundir, "data"))
try:
    subprocess.run([os.path.join(odir, "sn3d_ref")], cwd=rundir, env=dict(os.environ, ARTISB200_MODE="ref"),
                   stdout=open(os.path.join(rundir, "stdout.txt"), "w"), stderr=subprocess.STDOUT, timeout=a.timeout)
except subprocess.TimeoutExpired:
    print("(timeout)")
log = open(os.path.join(rundir, "output_0-0.txt")).read()
npk = 0
for m in re.finditer(r"ARTISB200_TIMING nts (\d+) mode \w+ npackets (\d+) wall_s ([\d.]+) interactions (\d+)", log):
    nts, npk, wall, nint = int(m[1]), int(m[2]), float(m[3]), int(m[4])
    act = re.search(rf"timestep {nts}: ma_stat_activation: collexc (\d+) collion (\d+) ntcollexc \d+ ntcollion \d+ bb (\d+) bf (\d+)", log)
    kst = re.search(rf"timestep {nts}: k_stat_to: ma_collexc (\d+) ma_collion (\d+) r_ff (\d+) r_fb (\d+) r_bb (\d+)", log)
    es = re.search(rf"timestep {nts}: electron_scatterings (\d+)", log)
    cc = re.search(rf"timestep {nts}: cellcrossings (\d+)", log)
    print(f"ts{nts}: {nint / npk:10.1f} int/pkt  wall {wall:7.2f}s  MA act: collexc {act[1]} collion {act[2]} bb {act[3]} bf {act[4]} | "
          f"k->: ff {kst[3]} fb {kst[4]} bb {kst[5]} | escat {es[1]} cellcross {cc[1]}")
