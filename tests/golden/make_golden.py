#!/usr/bin/env python3
"""Generate the committed oracle fixtures from the reference itself (development container only).

For each toy config the UNMODIFIED reference (oracle/_ref/<config>/parity/sn3d_ref: the reference's own
sources compiled with -DGPU_ON and its REPRODUCIBLE flags, plus the snapshot hooks of
integration/update_packets_b200.cc) is run on the synthetic inputs written by tools/gen_inputs.py, in the
order-independent per-packet schedule (every packet driven through the reference's do_packet() with its own
RNG stream). The inputs of update_packets() (static tables, cell state, packets) and its outputs (packets,
estimators, counters) are stored as
    tests/golden/<config>_static.npz, tests/golden/<config>_ts<N>.npz  (keys "before/<name>", "after/<name>")
Usage: python tests/golden/make_golden.py [config ...]   (needs `python __graft_entry__.py build` to have built oracle/_ref)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import run_oracle  # noqa: E402
from artis_b200 import snapshot as snap  # noqa: E402

GOLDEN = {
    "classic_toy": [0, 3],
    "classic_toy_1d": [0, 3],
    "classic3d_toy": [0, 2],
    "kilonova_toy": [1, 4],
    "classic_multibin_toy": [2, 4],
    "classic_nlte_toy": [2, 4],
    "classic_nt_toy": [2, 3],
    "classic_ntexc_toy": [2, 3],
    "classic_detailedbf_toy": [1, 3],
    "nltephot_toy": [1, 3],
    "kilonova_guttman_toy": [1],
    "kilonova_wollaeger_toy": [1],
    "kilonova_barnes_toy": [1],
    "classic3d_grey_toy": [0, 2],
    "kilonova_xcom_toy": [1, 4],
    # expansion-opacity / bound-bound thermalisation r-packet modes (rpkt.cc:221-320, 628-651, 964-981)
    "kilonova_expansionopac_toy": [2, 4],
    "kilonova_expopac_retrace_toy": [4],
    "kilonova_bbtherm_toy": [4],
    # BASELINE configs[1] at full atomic-data and grid size, 2000 packets: bench-scale KATs, histories, sampled cell tables
    "kilonova_2d_kat": [2],
}
# per config: extra environment of the oracle run
EXTRA_ENV = {"kilonova_2d_kat": {"ARTISB200_DUMP_CELLS": "6"}}


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    only = sys.argv[1:]  # optional: regenerate just these configs
    for config, timesteps in GOLDEN.items():
        if only and config not in only:
            continue
        rundir = run_oracle.run(config, "parity", "ref_perpacket", ",".join(str(t) for t in timesteps),
                                env_extra=EXTRA_ENV.get(config))
        dump = os.path.join(rundir, "dump")
        static = snap.read_snapshot(os.path.join(dump, "static.abt"))
        np.savez_compressed(os.path.join(here, f"{config}_static.npz"), **static)
        for nts in timesteps:
            before = snap.read_snapshot(os.path.join(dump, f"ts{nts}_before.abt"))
            after = snap.read_snapshot(os.path.join(dump, f"ts{nts}_after.abt"))
            arrays = {f"before/{k}": v for k, v in before.items()}
            arrays.update({f"after/{k}": v for k, v in after.items()})
            out = os.path.join(here, f"{config}_ts{nts}.npz")
            np.savez_compressed(out, **arrays)
            print(f"{out}: {os.path.getsize(out) / 1024:.0f} KiB, interactions {int(after['counters'][26])}")


if __name__ == "__main__":
    main()
