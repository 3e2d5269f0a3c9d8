#!/usr/bin/env python3
"""Known answers for the spectra / light-curve binning, from the reference itself.

The oracle build (oracle/_ref/<config>/parity/sn3d_ref) is run as for make_golden.py with ARTISB200_DUMP_SPECTRA=1: after
update_packets of the chosen timestep the snapshot hook hands the packets to the reference's OWN add_to_spec_res /
add_to_lc_res (spectrum_lightcurve.cc:544-718, through integration/ref_access/ref_spectrum_lightcurve.cc) for the
angle-averaged bin with the emission / absorption decomposition and for each of the 100 direction bins, and to
get_escapedirectionbin (vectors.h:147). The packets are the "after" packets of the committed fixture
tests/golden/<config>_ts<N>.npz (checked byte for byte here), so only the binned arrays are stored:
    tests/golden/<config>_spectra_ts<N>.npz
Usage: python tests/golden/make_golden_spectra.py [config ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import run_oracle  # noqa: E402
from artis_b200 import snapshot as snap  # noqa: E402

# 3-D, 2-D and 1-D grids; classic (bound-free / free-free emission types) and kilonova (escaped gamma packets, line emission)
GOLDEN_SPECTRA = {"classic3d_toy": 2, "kilonova_toy": 4, "classic_toy_1d": 3, "classic_detailedbf_toy": 3}


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    only = sys.argv[1:]
    for config, nts in GOLDEN_SPECTRA.items():
        if only and config not in only:
            continue
        rundir = run_oracle.run(config, "parity", "ref_perpacket", str(nts), env_extra={"ARTISB200_DUMP_SPECTRA": "1"})
        after = snap.read_snapshot(os.path.join(rundir, "dump", f"ts{nts}_after.abt"))
        committed = np.load(os.path.join(here, f"{config}_ts{nts}.npz"))
        if not np.array_equal(after["packets.aos"], committed["after/packets.aos"]):
            raise RuntimeError(f"{config} ts{nts}: this run's packets differ from the committed fixture")
        arrays = {k: v for k, v in after.items() if k.startswith(("ref.spec.", "ref.lc."))}
        out = os.path.join(here, f"{config}_spectra_ts{nts}.npz")
        np.savez_compressed(out, **arrays)
        print(f"{out}: {os.path.getsize(out) / 1024:.0f} KiB; flux bins filled {np.count_nonzero(arrays['ref.spec.flux'])}, "
              f"emission {np.count_nonzero(arrays['ref.spec.emission'])}, absorption {np.count_nonzero(arrays['ref.spec.absorption'])}, "
              f"direction-resolved {np.count_nonzero(arrays['ref.spec.flux_res'])}, gamma lc {np.count_nonzero(arrays['ref.lc.gamma_lum'])}")


if __name__ == "__main__":
    main()
