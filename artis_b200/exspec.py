"""The reference's post-processor exspec (exspec.cc:196-260), with the binning on the device.

exspec reads the packets<NN>_<rank>.out text files of a finished run (packet.cc:163-222) and bins ALL packets once for the
angle-averaged result and once more for each of the 100 direction bins (exspec.cc:28-150, one do_direction_bin call per bin).
Here the files are read by the library's reader (artisb200_read_text_packets: the reference's reader including its behaviour at
"nan" columns), uploaded once, binned in ONE pass (artisb200_bin_escaped_packets, emission / absorption decomposition for
every set, like exspec's do_emission_absorption = true) and written in the reference's formats by artis_b200/spectra.py:
    light_curve.out  spec.out  emission.out  emissiontrue.out  absorption.out  gamma_light_curve.out
    speclc_angle_res/{light_curve,spec,emission,emissiontrue,absorption}_res_NN.out      (2-D / 3-D models)
and, with POL_ON, specpol.out / emissionpol.out / absorptionpol.out (Stokes I, Q, U), and gamma_spec.out (escaped gamma rays).

    python -m artis_b200.exspec --preset classic --static dump/static.abt --rundir . [--nprocs 1] [--outdir .]
`static.abt` is the static-table snapshot the binding writes with ARTISB200_DUMP_DIR (integration/update_packets_b200.cc)."""
import argparse
import os

import numpy as np

from . import lib as ablib
from . import snapshot as snap
from . import spectra as spectra_mod


def exspec(engine, static, rundir, outdir=None, nprocs_exspec=1, stride=240, keep_escaped_gammas=True, only_dirbins=None, pol_on=False):
    """engine: ArtisB200 with the static tables committed; static: the named static arrays (grid type, timesteps);
    only_dirbins: write the files of these direction bins only (default: all 100); pol_on: the run's POL_ON (Stokes Q / U
    spectra: specpol.out, emissionpol.out, absorptionpol.out)"""
    outdir = outdir or rundir
    chunks = []
    for p in range(nprocs_exspec):  # exspec.cc:211-213
        raw, n = engine.read_text_packets(os.path.join(rundir, f"packets{0:02d}_{p:04d}.out"), stride)
        chunks.append(raw[:n * stride])
    aos = np.concatenate(chunks)
    npackets = aos.size // stride
    engine.upload_packets(aos, npackets, stride)
    multidimensional = int(static["scalar.grid_type"][0]) != 0  # GridType::SPHERICAL1D (exspec.cc:215)
    engine.set_option("spec_stokes", int(bool(pol_on)))
    engine.set_option("spec_gamma_spectrum", int(bool(keep_escaped_gammas)))
    engine.bin_escaped_packets(direction_bins=multidimensional, emission_absorption=(2 if multidimensional else 1),
                               nprocs_exspec=nprocs_exspec)
    b = spectra_mod.binned(engine)
    mid = static["timesteps.mid"]
    ntimesteps = mid.size - 1
    os.makedirs(outdir, exist_ok=True)
    for s in range(b["flux"].shape[0]):
        if s > 0 and only_dirbins is not None and (s - 1) not in only_dirbins:
            continue
        if s == 0:
            where, tag = outdir, ""
        else:
            where, tag = os.path.join(outdir, spectra_mod.OUTDIR_RESFILES), f"_res_{s - 1:02d}"
            os.makedirs(where, exist_ok=True)
        spectra_mod.write_light_curve(os.path.join(where, f"light_curve{tag}.out"), mid, b["lc_lum"][s], b["lc_lumcmf"][s], ntimesteps)
        spectra_mod.write_spectrum_file(os.path.join(where, f"spec{tag}.out"), mid, b["lower_freq"], b["delta_freq"], b["flux"][s], ntimesteps)
        spectra_mod.write_columns_file(os.path.join(where, f"emission{tag}.out"), b["emission"][s], ntimesteps)
        spectra_mod.write_columns_file(os.path.join(where, f"emissiontrue{tag}.out"), b["trueemission"][s], ntimesteps)
        spectra_mod.write_columns_file(os.path.join(where, f"absorption{tag}.out"), b["absorption"][s], ntimesteps)
        if pol_on:  # exspec.cc:100-103, 147-151
            spectra_mod.write_specpol(os.path.join(where, f"specpol{tag}.out"), os.path.join(where, f"emissionpol{tag}.out"),
                                      os.path.join(where, f"absorptionpol{tag}.out"), mid, b["lower_freq"], b["delta_freq"],
                                      [b["flux"][s], b["flux_q"][s], b["flux_u"][s]],
                                      [b["emission"][s], b["emission_q"][s], b["emission_u"][s]],
                                      [b["absorption"][s], b["absorption_q"][s], b["absorption_u"][s]])
        if s == 0 and keep_escaped_gammas:
            spectra_mod.write_light_curve(os.path.join(where, "gamma_light_curve.out"), mid, b["gamma_lc_lum"], b["gamma_lc_lumcmf"], ntimesteps)
            spectra_mod.write_spectrum_file(os.path.join(where, "gamma_spec.out"), mid, b["gamma_lower_freq"], b["gamma_delta_freq"],
                                            b["gamma_flux"], ntimesteps)
    return b


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--preset", required=True)
    ap.add_argument("--static", required=True, help="static.abt written by the binding (ARTISB200_DUMP_DIR)")
    ap.add_argument("--rundir", default=".")
    ap.add_argument("--outdir", default=None)
    ap.add_argument("--nprocs", type=int, default=1, help="nprocs_exspec: packets files to read (input.txt line 21)")
    ap.add_argument("--stride", type=int, default=240)
    ap.add_argument("--pol", action="store_true", help="the run was compiled with POL_ON: write the Stokes files")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    static = snap.read_snapshot(a.static)
    eng = ablib.ArtisB200(preset=a.preset, device=a.device)
    eng.set_arrays(static)
    eng.commit_static()
    exspec(eng, static, a.rundir, a.outdir, a.nprocs, a.stride, pol_on=a.pol)
    eng.close()


if __name__ == "__main__":
    main()
