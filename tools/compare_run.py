#!/usr/bin/env python3
"""Replay one timestep of an oracle dump through an artis_b200 library and compare packet by packet.

  python tools/compare_run.py <dumpdir> <nts> --lib <path to .so> [--rng xoshiro|philox]

<dumpdir> holds static.abt, ts<nts>_before.abt and ts<nts>_after.abt written by the oracle build
(oracle/_ref/<config>/parity/sn3d_ref with ARTISB200_MODE=ref_perpacket ARTISB200_DUMP_DIR=...).
With --rng xoshiro the library continues each packet's own Xoshiro128++ stream from the dumped state, so
every packet history should reproduce the oracle's up to floating-point rounding.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artis_b200 import lib as ablib  # noqa: E402
from artis_b200 import snapshot as snap  # noqa: E402

FLOAT_FIELDS = ["prop_time", "pos", "dir", "nu_cmf", "e_cmf", "nu_rf", "e_rf", "em_pos", "em_time", "absorptionfreq",
                "stokes_q", "stokes_u", "trueem_pos", "trueem_time", "escape_time", "tdecay"]
INT_FIELDS = ["next_trans", "nscatterings", "emissiontype", "absorptiontype", "trueemissiontype", "type", "cellindex",
              "escape_type", "number", "originated_from_particlenotgamma", "pellet_decaytype", "pellet_nucindex"]


def run_timestep(libpath, dumpdir, nts, rng="xoshiro", max_steps=0, device=0):
    static = snap.read_snapshot(os.path.join(dumpdir, "static.abt"))
    before = snap.read_snapshot(os.path.join(dumpdir, f"ts{nts}_before.abt"))
    eng = ablib.ArtisB200(libpath=libpath, device=device)
    eng.set_option("rng_mode", 1 if rng == "xoshiro" else 0)
    eng.set_option("max_steps_per_launch", max_steps)
    eng.set_arrays(static)
    eng.commit_static()
    eng.set_arrays(before)
    eng.begin_timestep(nts)
    n = int(before["packets.count"][0])
    stride = int(before["packets.stride"][0])
    aos = before["packets.aos"].copy()
    eng.update_packets_host(nts, aos, n, stride)
    est = eng.estimators()
    built = {k: eng.get_array(k) for k in ["built.levelpops", "built.maprocessrates", "built.matrans", "built.cooling_contrib",
                                           "built.cont_nnlevel", "built.chi_ff_nnionpart", "built.corrphotoioncoeff"]}
    built["built.cont_keepbits"] = eng.get_array("built.cont_keepbits")
    timing = eng.last_timing_ms()
    eng.close()
    return aos.view(snap.packet_dtype(stride)), est, built, timing


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    denom = np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300)
    err = np.abs(a - b) / denom
    err[both_nan] = 0.0
    err[np.isnan(err)] = np.inf
    return err


def compare(pk, est, after, tol=1e-9, verbose=True):
    ref = snap.packets_view(after)
    n = len(ref)
    bad = np.zeros(n, dtype=bool)
    worst = {}
    for f in INT_FIELDS:
        mism = ref[f] != pk[f]
        worst[f] = int(mism.sum())
        bad |= mism
    for f in FLOAT_FIELDS:
        e = rel_err(ref[f], pk[f])
        if e.ndim > 1:
            e = e.max(axis=1)
        worst[f] = float(e.max()) if n else 0.0
        bad |= e > tol
    frac_ok = 1.0 - bad.mean() if n else 1.0
    est_err = {}
    for name, val in est.items():
        if name in after and name not in ("diag",):
            r = after[name]
            if r.size == 0:
                continue
            m = min(r.size, val.size)  # the reference allocates some estimators with one spare element
            r, val = r[:m], val[:m]
            if np.issubdtype(r.dtype, np.integer):
                est_err[name] = int(np.abs(r.astype(np.int64) - val.astype(np.int64)).max())
            else:
                # per-element relative error: every entry (cell, ion, bin) is held to the tolerance on its own, so a
                # low-signal cell cannot hide behind the array's largest entry; the absolute floor (1e-20 of the
                # largest reference entry) only keeps 0 against 0 and denormal dust from dividing by nothing
                scale = np.abs(r).max()
                floor = 1e-20 * scale if scale > 0 else 1e-300
                denom = np.maximum(np.maximum(np.abs(r), np.abs(val)), floor)
                est_err[name] = float((np.abs(r - val) / denom).max())
    if verbose:
        print(f"packets: {n}  matching within {tol:g}: {frac_ok * 100:.4f}%  ({int(bad.sum())} differ)")
        print("  per-field worst (int: mismatches, float: max rel err):")
        for k, v in worst.items():
            if v:
                print(f"    {k:34s} {v}")
        print("  estimators (max per-element relative error; counters: max abs diff):")
        for k, v in est_err.items():
            print(f"    {k:20s} {v}")
        if bad.any():
            idx = np.nonzero(bad)[0][:5]
            for i in idx:
                print(f"  first differing packet #{i}: number {ref['number'][i]} ref type {ref['type'][i]} got {pk['type'][i]} "
                      f"ref cell {ref['cellindex'][i]} got {pk['cellindex'][i]} ref nu_cmf {ref['nu_cmf'][i]:.6e} got {pk['nu_cmf'][i]:.6e} "
                      f"ref t {ref['prop_time'][i]:.8e} got {pk['prop_time'][i]:.8e}")
    return frac_ok, worst, est_err


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("dumpdir")
    ap.add_argument("nts", type=int)
    ap.add_argument("--lib", required=True)
    ap.add_argument("--rng", default="xoshiro")
    ap.add_argument("--max-steps", type=int, default=0)
    ap.add_argument("--tol", type=float, default=1e-9)
    args = ap.parse_args()
    pk, est, built, timing = run_timestep(args.lib, args.dumpdir, args.nts, args.rng, args.max_steps)
    after = snap.read_snapshot(os.path.join(args.dumpdir, f"ts{args.nts}_after.abt"))
    print("timing ms (total, propagate, schedule):", timing)
    print("interactions:", est["counters"][26], "ref:", after["counters"][26])
    compare(pk, est, after, args.tol)
