#!/usr/bin/env python3
"""Known answers for the LTE part of update_grid_cell, from the reference itself.

The oracle build is run as for make_golden.py with ARTISB200_DUMP_GRID=1: before update_packets of the chosen timestep the
snapshot hook (integration/update_packets_b200.cc emit_reference_lte_gridupdate) calls, for every cell, the reference's OWN
    radfield::get_T_J_from_J                             radfield.cc:956-979   on a ladder of J values (clamps, non-finite J)
    calculate_cellpartfuncts + calculate_ion_balance_nne ltepop.cc:426-532     in Saha mode on the cell state as it is
and stores inputs it needs beyond the committed fixture (number densities of the elements) and the results:
    tests/golden/<config>_grid_ts<N>.npz
The cell state (temperatures, density, composition) is the "before" state of tests/golden/<config>_ts<N>.npz; the run's
packets are checked against that fixture, so the hook provably left the reference's state as it found it.
Usage: python tests/golden/make_golden_grid.py [config ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import run_oracle  # noqa: E402
from artis_b200 import snapshot as snap  # noqa: E402

# (classic_nlte_toy: partition functions that read the NLTE solver's level and superlevel populations, ltepop.cc:177-197)
GOLDEN_GRID = {"classic3d_toy": 2, "kilonova_toy": 4, "classic_toy_1d": 3, "kilonova_2d_kat": 2, "classic_nlte_toy": 4}
EXTRA_ENV = {"kilonova_2d_kat": {"ARTISB200_DUMP_CELLS": "6"}}


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    only = sys.argv[1:]
    for config, nts in GOLDEN_GRID.items():
        if only and config not in only:
            continue
        env = {"ARTISB200_DUMP_GRID": "1"}
        env.update(EXTRA_ENV.get(config, {}))
        rundir = run_oracle.run(config, "parity", "ref_perpacket", str(nts), env_extra=env)
        before = snap.read_snapshot(os.path.join(rundir, "dump", f"ts{nts}_before.abt"))
        after = snap.read_snapshot(os.path.join(rundir, "dump", f"ts{nts}_after.abt"))
        committed = np.load(os.path.join(here, f"{config}_ts{nts}.npz"))
        if not np.array_equal(after["packets.aos"], committed["after/packets.aos"]):
            raise RuntimeError(f"{config} ts{nts}: this run's packets differ from the committed fixture")
        for key in ("cell.Te", "cell.nne", "cell.ion_groundlevelpops", "cell.ion_partfuncts"):
            if not np.array_equal(before[key], committed["before/" + key]):
                raise RuntimeError(f"{config} ts{nts}: {key} differs from the committed fixture")
        arrays = {k: v for k, v in before.items() if k.startswith("ref.grid.") or k == "cell.elem_numberdens"}
        out = os.path.join(here, f"{config}_grid_ts{nts}.npz")
        np.savez_compressed(out, **arrays)
        same = np.mean(arrays["ref.grid.nne"] == before["cell.nne"])
        print(f"{out}: {os.path.getsize(out) / 1024:.0f} KiB; {arrays['ref.grid.nne'].size} cells, nne {arrays['ref.grid.nne'].min():.3g}.."
              f"{arrays['ref.grid.nne'].max():.3g}, cells whose Saha nne equals the run's own nne: {same * 100:.0f} %, "
              f"uppermost ions {np.bincount(arrays['ref.grid.uppermost_ion'] + 1)}")


if __name__ == "__main__":
    main()
