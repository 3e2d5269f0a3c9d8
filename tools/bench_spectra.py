#!/usr/bin/env python3
"""Timing of the spectra / light-curve binning kernel on synthetic final packets, with the reference's own
binning timed beside it.

  python tools/bench_spectra.py [--packets 10000000] [--out gpurun_out/r2_spectra_bench.json] [--reference-replicas 700]

GPU: artisb200_bin_escaped_packets on N device-resident packets (static tables of the classic3d_toy fixture), CUDA-event
time of the kernel (median of 7 after 2 warm-up passes; the packet records, 2.3 GB at 1e7, are larger than L2) for
  angle-averaged only | + 100 direction bins | + emission/absorption of the angle-averaged set | + of every set.
Algorithmic bytes: 32-byte type sector per packet + 136 B (kinematics, energies, escape type and time) per escaped packet,
+ 28 B per escaped r-packet with the decomposition.
Reference: the compiled reference's own add_to_spec_res / add_to_lc_res (oracle/_ref/classic3d_toy, parity flavour, one
core), the fixture's packets replicated, dirbin -1 and 0..99 as write_partial_lightcurve_spectra does
(spectrum_lightcurve.cc:316-337)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from artis_b200 import snapshot as snap  # noqa: E402


def synthetic_packets(static, n, seed=11, stride=240):
    """random final packets over every branch of add_to_spec_res / add_to_lc_res: escaped r-packets and gamma packets, packets
    still in flight, arrival times and frequencies inside and outside the binned ranges, every kind of emission type"""
    rng = np.random.default_rng(seed)
    pk = np.zeros(n, dtype=snap.packet_dtype(stride))
    ts_start = static["timesteps.start"]
    tmin, tmax = float(static["scalar.tmin"][0]), float(ts_start[-1])
    rmax = float(static["scalar.rmax"][0])
    nlines = static["line.nu"].size
    nbf = static["cont.nu_edge"].size
    kind = rng.random(n)
    pk["type"] = np.where(kind < 0.8, 32, np.where(kind < 0.9, 11, 100))
    pk["escape_type"] = np.where(rng.random(n) < 0.85, 11, np.where(rng.random(n) < 0.8, 10, 12))
    v = rng.normal(size=(n, 3))
    pk["dir"] = v / np.linalg.norm(v, axis=1)[:, None] * (1. + 1e-9 * rng.normal(size=n))[:, None]  # not exactly normalised
    pk["pos"] = rng.normal(size=(n, 3)) * rmax * (tmax / tmin) * 0.3
    pk["escape_time"] = np.exp(rng.uniform(np.log(tmin * 0.8), np.log(tmax * 1.3), size=n)).astype(np.float32)
    pk["nu_rf"] = np.exp(rng.uniform(np.log(0.7e14), np.log(7e15), size=n))
    pk["e_rf"] = rng.uniform(0.5, 1.5, size=n) * 1e40
    pk["e_cmf"] = pk["e_rf"] * rng.uniform(0.9, 1.1, size=n)
    for field in ("emissiontype", "trueemissiontype"):
        sel = rng.random(n)
        et = rng.integers(0, nlines, size=n)
        et = np.where(sel < 0.5, et, np.where(sel < 0.75, -1 - rng.integers(0, max(nbf, 1), size=n), np.where(sel < 0.9, -9999999, -9999000)))
        pk[field] = et
    pk["absorptiontype"] = np.where(rng.random(n) < 0.6, rng.integers(0, nlines, size=n), -1)
    pk["absorptionfreq"] = np.exp(rng.uniform(np.log(0.7e14), np.log(7e15), size=n))
    pk["number"] = np.arange(n)
    return pk




def reference_timing(replicas):
    import run_oracle
    rundir = run_oracle.run("classic3d_toy", "parity", "ref_perpacket", "2", rundir="/tmp/artisb200_spectra_ref",
                            env_extra={"ARTISB200_DUMP_SPECTRA": "1", "ARTISB200_TIME_SPECTRA": str(replicas)})
    for line in open(os.path.join(rundir, "output_0-0.txt")):
        if "ARTISB200_SPECTRA_TIMING" in line:
            tok = line.split("ARTISB200_SPECTRA_TIMING", 1)[1].split()
            d = {tok[i]: tok[i + 1] for i in range(0, len(tok) - 1, 2)}
            return {"packets": int(d["npackets"]), "passes": int(d["passes"]), "wall_s": float(d["wall_s"]), "cores": 1,
                    "packets_per_s": int(d["npackets"]) / float(d["wall_s"]), "kind": "reference",
                    "sample": "classic3d_toy ts2 packets replicated, parity flavour (-O2, no fast-math)"}
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--packets", type=int, default=10_000_000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_spectra_bench.json"))
    ap.add_argument("--reference-replicas", type=int, default=700)
    ap.add_argument("--passes", type=int, default=7)
    a = ap.parse_args()
    from artis_b200 import lib as ablib
    from tests import fixtures
    static = fixtures.load_golden("classic3d_toy", 2)["static"]
    n = a.packets
    pk = synthetic_packets(static, n)
    escaped = int((pk["type"] == 32).sum())
    escaped_r = int(((pk["type"] == 32) & (pk["escape_type"] == 11)).sum())
    peak = 6553.0
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak = float(peaks.get("hbm_gbs_burst", peaks.get("hbm_gbs", peak)))
    except Exception:
        pass
    eng = ablib.ArtisB200(libpath=ablib.library_path("classic"))
    eng.set_arrays(static)
    eng.commit_static()
    eng.upload_packets(pk.view(np.uint8), n, pk.dtype.itemsize)
    variants = []
    for name, dirbins, emabs in (("angle-averaged", 0, 0), ("direction bins", 1, 0), ("direction bins + decomposition of set 0", 1, 1),
                                 ("direction bins + decomposition of every set", 1, 2)):
        times = []
        for k in range(a.passes + 2):
            eng.bin_escaped_packets(direction_bins=dirbins, emission_absorption=emabs, nprocs_exspec=1)
            if k >= 2:
                times.append(eng.last_binning_ms())
        ms = sorted(times)[len(times) // 2]
        b_alg = 32 * n + 136 * escaped + (28 * escaped_r if emabs else 0)
        variants.append({"variant": name, "ms": ms, "min_ms": min(times), "algorithmic_bytes": b_alg,
                         "achieved_GBs": b_alg / (ms * 1e-3) / 1e9, "frac_of_peak": b_alg / (ms * 1e-3) / 1e9 / peak,
                         "packets_per_s": n / (ms * 1e-3)})
        print(f"[spectra] {name}: {ms:.3f} ms, {variants[-1]['achieved_GBs']:.0f} GB/s algorithmic", flush=True)
    eng.close()
    out = {"kernel": "k_bin_escaped", "packets": n, "escaped": escaped, "escaped_rpkts": escaped_r, "peak_GBs": peak,
           "variants": variants}
    if a.reference_replicas > 0:
        try:
            out["reference"] = reference_timing(a.reference_replicas)
            if out["reference"]:
                out["speedup_direction_bins"] = variants[1]["packets_per_s"] / out["reference"]["packets_per_s"]
        except Exception as e:
            out["reference"] = {"error": str(e)}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
