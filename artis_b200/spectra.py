"""Host side of the device spectra / light-curve binning: the reference's spectrum_lightcurve.cc interface.

`write_partial_lightcurve_spectra(engine, nts, outdir, ...)` is the call sn3d.cc:806 makes after every timestep; here the
binning runs on the device-resident packets (artisb200_bin_escaped_packets, ONE pass for the angle-averaged result and all
MABINS direction bins) and this module only formats the arrays into the reference's output files:
    light_curve.out, gamma_light_curve.out          write_light_curve            spectrum_lightcurve.cc:339-354
    spec.out                                        write_spectrum_file          spectrum_lightcurve.cc:339-358 (anonymous namespace)
    emission.out, emissiontrue.out                  write_emission_spectrum_file spectrum_lightcurve.cc:362-379
    absorption.out                                  write_absorption_spectrum_file 381-396
    speclc_angle_res/{light_curve,spec,...}_res_NN.out  for the direction bins (constants.h:96)
Numbers are printed with the reference's "{:g}" (= printf %g)."""
import os

import numpy as np

MABINS = 100  # exspec.h:12
DAY = 86400.0  # constants.h:35
LSUN = 3.826e33  # constants.h:28
OUTDIR_RESFILES = "speclc_angle_res"  # constants.h:96


def binned(engine, ntimesteps=None):
    """the results of the last engine.bin_escaped_packets() as arrays shaped [set][nnu][nts]...; set 0 = angle-averaged"""
    lower = engine.get_array("spec.lower_freq")
    nnu = lower.size
    flux = engine.get_array("spec.flux")
    lum = engine.get_array("lc.lum")
    nts = engine.get_array("lc.gamma_lum").size
    nsets = lum.size // nts
    out = {"lower_freq": lower, "delta_freq": engine.get_array("spec.delta_freq"), "flux": flux.reshape(nsets, nnu, nts),
           "lc_lum": lum.reshape(nsets, nts), "lc_lumcmf": engine.get_array("lc.lumcmf").reshape(nsets, nts),
           "gamma_lc_lum": engine.get_array("lc.gamma_lum"), "gamma_lc_lumcmf": engine.get_array("lc.gamma_lumcmf")}
    n_em = engine.array_count("spec.emission")
    if n_em > 0:
        n_abs = engine.array_count("spec.absorption")
        # n_em = sets * nnu * nts * proccount, n_abs = sets * nnu * nts * ioncount, proccount = 2 * ioncount + 1: the
        # decomposition is kept for set 0 only (emission_absorption = 1) or for every set (2)
        sets_em = (n_em - 2 * n_abs) // (nnu * nts)
        em = engine.get_array("spec.emission")
        ab = engine.get_array("spec.absorption")
        out["emission"] = em.reshape(sets_em, nnu, nts, -1)
        out["trueemission"] = engine.get_array("spec.trueemission").reshape(sets_em, nnu, nts, -1)
        out["absorption"] = ab.reshape(sets_em, nnu, nts, -1)
    if engine.array_count("spec.dirbin") > 0:
        out["dirbin"] = engine.get_array("spec.dirbin")
    if engine.array_count("spec.flux_q") == flux.size:  # option spec_stokes
        for stokes in ("q", "u"):
            out[f"flux_{stokes}"] = engine.get_array(f"spec.flux_{stokes}").reshape(nsets, nnu, nts)
            if "emission" in out and engine.array_count(f"spec.emission_{stokes}") == out["emission"].size:
                out[f"emission_{stokes}"] = engine.get_array(f"spec.emission_{stokes}").reshape(out["emission"].shape)
                out[f"absorption_{stokes}"] = engine.get_array(f"spec.absorption_{stokes}").reshape(out["absorption"].shape)
    if engine.array_count("spec.gamma_flux") == nnu * nts:  # option spec_gamma_spectrum
        out["gamma_flux"] = engine.get_array("spec.gamma_flux").reshape(nnu, nts)
        out["gamma_lower_freq"] = engine.get_array("spec.gamma_lower_freq", dtype=np.float32)
        out["gamma_delta_freq"] = engine.get_array("spec.gamma_delta_freq", dtype=np.float32)
    return out


def _g(x):
    return "%g" % x


def write_light_curve(path, timesteps_mid, lum, lumcmf, numtimesteps):
    """spectrum_lightcurve.cc:339-354: one line per timestep: mid [d], L [Lsun], L_cmf [Lsun]"""
    with open(path, "w") as f:
        for nts in range(numtimesteps):
            f.write(f"{_g(timesteps_mid[nts] / DAY)} {_g(lum[nts] / LSUN)} {_g(lumcmf[nts] / LSUN)}\n")


def write_spectrum_file(path, timesteps_mid, lower_freq, delta_freq, flux, numtimesteps):
    """write_spectrum_file: header "0 t_mid..." then per frequency bin the bin centre (float arithmetic) and the fluxes"""
    with open(path, "w") as f:
        f.write("0 " + "".join(_g(timesteps_mid[p] / DAY) + " " for p in range(numtimesteps)) + "\n")
        centre = lower_freq.astype(np.float32) + (delta_freq.astype(np.float32) / np.float32(2))
        for nnu in range(flux.shape[0]):
            f.write(_g(centre[nnu]) + " " + "".join(_g(v) + " " for v in flux[nnu, :numtimesteps]) + "\n")


def write_columns_file(path, table, numtimesteps):
    """write_emission_spectrum_file / write_absorption_spectrum_file: one line per (frequency bin, timestep), one column per
    process / ion"""
    zero_line = "0 " * table.shape[-1] + "\n"
    nonzero = table.any(axis=-1)
    with open(path, "w") as f:
        for nnu in range(table.shape[0]):
            if not nonzero[nnu, :numtimesteps].any():
                f.write(zero_line * numtimesteps)
                continue
            for nts in range(numtimesteps):
                f.write("".join(_g(v) + " " for v in table[nnu, nts]) + "\n" if nonzero[nnu, nts] else zero_line)


def write_specpol(path, emission_path, absorption_path, timesteps_mid, lower_freq, delta_freq, fluxes, emissions=None, absorptions=None):
    """write_specpol (spectrum_lightcurve.cc:426-485): I, Q and U spectra side by side for ALL timesteps; with the
    decomposition, one line per (frequency bin, Stokes component, timestep) in the emission / absorption files.
    fluxes / emissions / absorptions: the (I, Q, U) arrays of one set"""
    ntimesteps = fluxes[0].shape[1]
    centre = lower_freq.astype(np.float32) + (delta_freq.astype(np.float32) / np.float32(2))
    em_file = open(emission_path, "w") if emissions is not None else None
    ab_file = open(absorption_path, "w") if absorptions is not None else None
    with open(path, "w") as f:
        f.write("0" + "".join(" " + _g(timesteps_mid[p] / DAY) for p in range(ntimesteps)) * 3 + "\n")
        for nnu in range(fluxes[0].shape[0]):
            row = [_g(centre[nnu])]
            for k in range(3):
                row.extend(_g(v) for v in fluxes[k][nnu, :ntimesteps])
                if em_file is not None:
                    for nts in range(ntimesteps):
                        em_file.write(" ".join(_g(v) for v in emissions[k][nnu, nts]) + "\n")
                        ab_file.write(" ".join(_g(v) for v in absorptions[k][nnu, nts]) + "\n")
            f.write(" ".join(row) + "\n")
    for extra in (em_file, ab_file):
        if extra is not None:
            extra.close()


def write_partial_lightcurve_spectra(engine, nts, outdir, timesteps_mid, ntimesteps_finish=None, multidimensional=True,
                                     write_emissionabsorption_at_end=False, keep_escaped_gammas=True, nprocs_exspec=1,
                                     reduce=None):
    """spectrum_lightcurve.cc:316-337 + 219-313: bin the engine's packets and write the reference's files for the timesteps
    up to nts. Direction-resolved files and the emission / absorption decomposition only when the simulation is complete,
    like the reference. `reduce(array) -> array` sums an array over the ranks (the reference's MPI_Allreduce calls)."""
    complete = ntimesteps_finish is not None and nts >= ntimesteps_finish - 1
    do_emabs = bool(write_emissionabsorption_at_end and complete)
    dirbins = bool(multidimensional and complete)
    engine.bin_escaped_packets(direction_bins=dirbins, emission_absorption=(2 if (do_emabs and dirbins) else int(do_emabs)),
                               nprocs_exspec=nprocs_exspec)
    b = binned(engine)
    if reduce is not None:
        for key in ("flux", "lc_lum", "lc_lumcmf", "gamma_lc_lum", "gamma_lc_lumcmf", "emission", "trueemission", "absorption"):
            if key in b:
                b[key] = reduce(b[key])
    n = nts + 1
    os.makedirs(outdir, exist_ok=True)
    for s in range(b["flux"].shape[0]):
        if s == 0:
            where, tag = outdir, ""
        else:
            where, tag = os.path.join(outdir, OUTDIR_RESFILES), f"_res_{s - 1:02d}"
            os.makedirs(where, exist_ok=True)
        write_light_curve(os.path.join(where, f"light_curve{tag}.out"), timesteps_mid, b["lc_lum"][s], b["lc_lumcmf"][s], n)
        if s == 0 and keep_escaped_gammas:
            write_light_curve(os.path.join(where, "gamma_light_curve.out"), timesteps_mid, b["gamma_lc_lum"], b["gamma_lc_lumcmf"], n)
        write_spectrum_file(os.path.join(where, f"spec{tag}.out"), timesteps_mid, b["lower_freq"], b["delta_freq"], b["flux"][s], n)
        if do_emabs and s < b["emission"].shape[0]:
            write_columns_file(os.path.join(where, f"emission{tag}.out"), b["emission"][s], n)
            write_columns_file(os.path.join(where, f"emissiontrue{tag}.out"), b["trueemission"][s], n)
            write_columns_file(os.path.join(where, f"absorption{tag}.out"), b["absorption"][s], n)
    return b
