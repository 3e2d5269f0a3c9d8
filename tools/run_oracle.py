#!/usr/bin/env python3
"""Run a prebuilt oracle binary (oracle/_ref/<config>/<flavour>/sn3d_ref) on its synthetic inputs.

  python tools/run_oracle.py <config> [--flavor parity] [--mode ref_perpacket] [--dump-ts all] [--rundir DIR] [--seed N]
Returns/prints the run directory. Test infrastructure (used by tests/ and bench.py's CPU baseline leg)."""
import argparse
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_dir(config, flavor):
    return os.path.join(ROOT, "oracle", "_ref", config, flavor)


def run(config, flavor="parity", mode="ref_perpacket", dump_ts="all", rundir=None, seed=None, binary=None, env_extra=None,
        quiet=True):
    odir = oracle_dir(config, flavor)
    binary = binary or os.path.join(odir, "sn3d_ref")
    if not os.path.exists(binary):
        raise FileNotFoundError(f"{binary} not built (python __graft_entry__.py build, needs the reference source tree)")
    rundir = rundir or os.path.join(odir, "run")
    if os.path.isdir(rundir):
        shutil.rmtree(rundir)
    shutil.copytree(os.path.join(odir, "inputs"), rundir)
    datadir = os.path.join(ROOT, "oracle", "_ref", "data")
    os.symlink(datadir, os.path.join(rundir, "data"))
    import configs as _configs
    for extra in _configs.get(config).get("extra_files", []):  # e.g. xcom_photoion_data.txt, read from the run folder
        shutil.copy(os.path.join(datadir, extra), os.path.join(rundir, extra))
    if seed is not None:
        path = os.path.join(rundir, "input.txt")
        lines = open(path).read().split("\n")
        lines[0] = f"{int(seed):<24d} #  0"
        open(path, "w").write("\n".join(lines))
    env = dict(os.environ)
    env["ARTISB200_MODE"] = mode
    if dump_ts is not None:
        os.makedirs(os.path.join(rundir, "dump"), exist_ok=True)
        env["ARTISB200_DUMP_DIR"] = "dump"
        env["ARTISB200_DUMP_TS"] = str(dump_ts)
    if env_extra:
        env.update(env_extra)
    with open(os.path.join(rundir, "stdout.txt"), "w") as out:
        subprocess.run([binary], cwd=rundir, env=env, stdout=out if quiet else None, stderr=subprocess.STDOUT, check=True)
    return rundir


def timing_lines(rundir):
    """parse the ARTISB200_TIMING lines of output_0-0.txt -> list of dicts"""
    res = []
    with open(os.path.join(rundir, "output_0-0.txt")) as f:
        for line in f:
            if "ARTISB200_TIMING" in line:
                tok = line.split("ARTISB200_TIMING", 1)[1].split()
                d = {tok[i]: tok[i + 1] for i in range(0, len(tok) - 1, 2)}
                res.append({"nts": int(d["nts"]), "mode": d["mode"], "npackets": int(d["npackets"]),
                            "wall_s": float(d["wall_s"]), "interactions": int(d["interactions"])})
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--flavor", default="parity")
    ap.add_argument("--mode", default="ref_perpacket")
    ap.add_argument("--dump-ts", default="all")
    ap.add_argument("--rundir", default=None)
    ap.add_argument("--seed", type=int, default=None)
    a = ap.parse_args()
    rd = run(a.config, a.flavor, a.mode, a.dump_ts if a.dump_ts != "none" else None, a.rundir, a.seed)
    for t in timing_lines(rd):
        print(t)
    print(rd)
