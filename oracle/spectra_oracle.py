"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the reference's spectra / light-curve binning.

Never imported by the product (artis_b200/): only tests/ use it, as the checker of the device binning at sizes the
compiled reference's fixtures do not cover. PINNED: tests/test_spectra.py holds it to the arrays the reference's own
add_to_spec_res / add_to_lc_res / get_escapedirectionbin produced (tests/golden/*_spectra_ts*.npz) before it is trusted.

Follows, in the reference's operation order (numpy float64 = IEEE double, no fused operations):
    get_escapedirectionbin   vectors.h:147-175
    get_timestep             spectrum_lightcurve.cc:205-217
    get_logbinindex          sn3d.h:134-137
    init_spectra             spectrum_lightcurve.cc:485-504 (frequency grid in float)
    columnindex_from_emissiontype  spectrum_lightcurve.cc:169-203
    add_to_spec_res          spectrum_lightcurve.cc:544-661
    add_to_lc_res            spectrum_lightcurve.cc:691-718
"""
import numpy as np

NPHIBINS = 10  # exspec.h:10
NCOSTHETABINS = 10  # exspec.h:11
MABINS = NPHIBINS * NCOSTHETABINS
CLIGHT = 2.99792458e10  # constants.h
PARSEC = 3.0857e18  # constants.h:39
TYPE_GAMMA, TYPE_RPKT, TYPE_ESCAPE = 10, 11, 32  # packet.h
EMTYPE_NOTSET, EMTYPE_FREEFREE = -9999000, -9999999  # packet.h:79-80


def escape_direction_bin(dirs):
    """vectors.h:147-175 for an [n, 3] array of directions"""
    d = np.asarray(dirs, dtype=np.float64)
    mag = np.sqrt(((0. + d[:, 0] ** 2) + d[:, 1] ** 2) + d[:, 2] ** 2)
    d = d / mag[:, None]
    costheta = ((0. + d[:, 0] * 0.) + d[:, 1] * 0.) + d[:, 2] * 1.
    costhetabin = np.clip(((costheta + 1.0) * NCOSTHETABINS / 2.0).astype(np.int64), 0, NCOSTHETABINS - 1)
    # vec1 = dir x syn_dir, vec2 = xhat x syn_dir = (0, -1, 0), vec3 = vec2 x syn_dir = (-1, 0, 0)
    v1 = np.stack([d[:, 1] * 1. - 0. * d[:, 2], d[:, 2] * 0. - 1. * d[:, 0], d[:, 0] * 0. - 0. * d[:, 1]], axis=1)
    v1len = np.sqrt(((0. + v1[:, 0] ** 2) + v1[:, 1] ** 2) + v1[:, 2] ** 2)
    dot12 = ((0. + v1[:, 0] * 0.) + v1[:, 1] * -1.) + v1[:, 2] * 0.
    with np.errstate(divide="ignore", invalid="ignore"):
        cosphi = np.where(v1len > 1e-12, np.clip(dot12 / v1len, -1.0, 1.0), 1.0)
    testphi = ((0. + v1[:, 0] * -1.) + v1[:, 1] * 0.) + v1[:, 2] * 0.
    phi = np.where(testphi > 0, np.arccos(cosphi), np.arccos(cosphi) + np.pi)
    phibin = np.clip((phi / 2. / np.pi * NPHIBINS).astype(np.int64), 0, NPHIBINS - 1)
    return (costhetabin * NPHIBINS + phibin).astype(np.int32)


def frequency_grid(nu_min, nu_max, nnubins):
    dlognu = (np.log(nu_max) - np.log(nu_min)) / nnubins
    idx = np.arange(nnubins, dtype=np.float64)
    lower = np.exp(np.log(nu_min) + idx * dlognu).astype(np.float32)
    delta = (np.exp(np.log(nu_min) + (idx + 1.) * dlognu) - lower.astype(np.float64)).astype(np.float32)
    return dlognu, lower, delta


def _timestep(ts_start, ntimesteps, t):
    return np.clip(np.searchsorted(ts_start[:ntimesteps], t, side="right") - 1, 0, ntimesteps - 1)


def _logbin(value, nu_min, dlognu, nbins):
    return np.clip(np.floor((np.log(value) - np.log(nu_min)) / dlognu).astype(np.int64), 0, nbins - 1)


def bin_packets(pk, static, nu_min, nu_max, nnubins=1000, nprocs_exspec=1, direction_bins=True, emission_absorption=True):
    """pk: structured packet array (artis_b200.snapshot.packet_dtype); static: the named static tables.
    Returns the arrays of artisb200_bin_escaped_packets with set 0 = angle-averaged, sets 1.. = direction bins; the emission
    / absorption decomposition for set 0."""
    ts_start = np.asarray(static["timesteps.start"], dtype=np.float64)
    ts_width = np.asarray(static["timesteps.width"], dtype=np.float64)
    ntimesteps = ts_start.size - 1
    tmin = float(static["scalar.tmin"][0])
    tmax = float(ts_start[-1])
    vmax = float(static["scalar.vmax"][0])
    nelements = static["elem.anumber"].size
    max_nions = int(np.max(static["elem.nions"]))
    ioncount = nelements * max_nions
    proccount = 2 * ioncount + 1
    dlognu, lower, delta = frequency_grid(nu_min, nu_max, nnubins)
    nsets = 1 + MABINS if direction_bins else 1

    esc = pk["type"] == TYPE_ESCAPE
    dirbin_all = np.full(pk.size, -1, dtype=np.int32)
    dirbin_all[esc] = escape_direction_bin(pk["dir"][esc])
    out = {"lower_freq": lower, "delta_freq": delta, "dirbin": dirbin_all,
           "flux": np.zeros((nsets, nnubins, ntimesteps)), "lc_lum": np.zeros((nsets, ntimesteps)),
           "lc_lumcmf": np.zeros((nsets, ntimesteps)), "gamma_lc_lum": np.zeros(ntimesteps), "gamma_lc_lumcmf": np.zeros(ntimesteps)}
    if emission_absorption:
        out["emission"] = np.zeros((1, nnubins, ntimesteps, proccount))
        out["trueemission"] = np.zeros((1, nnubins, ntimesteps, proccount))
        out["absorption"] = np.zeros((1, nnubins, ntimesteps, ioncount))
    inverse_gamma = np.sqrt(1. - (vmax * vmax / (CLIGHT * CLIGHT)))

    for kind, sel in (("rpkt", esc & (pk["escape_type"] == TYPE_RPKT)), ("gamma", esc & (pk["escape_type"] == TYPE_GAMMA))):
        p = pk[sel]
        dirbin = dirbin_all[sel].astype(np.int64)
        escape_time = p["escape_time"].astype(np.float64)
        dot = ((0. + p["pos"][:, 0] * p["dir"][:, 0]) + p["pos"][:, 1] * p["dir"][:, 1]) + p["pos"][:, 2] * p["dir"][:, 2]
        t_arrive = escape_time - (dot / CLIGHT)
        arrives = (t_arrive > tmin) & (t_arrive < tmax)
        nts = _timestep(ts_start, ntimesteps, t_arrive)
        lum = p["e_rf"] / ts_width[nts]
        t_cmf = escape_time * inverse_gamma
        in_cmf = (t_cmf > tmin) & (t_cmf < tmax)
        nts_cmf = _timestep(ts_start, ntimesteps, t_cmf)
        lumcmf = p["e_cmf"] / ts_width[nts_cmf]
        if kind == "gamma":
            np.add.at(out["gamma_lc_lum"], nts[arrives], (lum * 1. / nprocs_exspec)[arrives])
            np.add.at(out["gamma_lc_lumcmf"], nts_cmf[in_cmf], (lumcmf * 1. / nprocs_exspec / inverse_gamma)[in_cmf])
            continue
        np.add.at(out["lc_lum"][0], nts[arrives], (lum * 1. / nprocs_exspec)[arrives])
        np.add.at(out["lc_lumcmf"][0], nts_cmf[in_cmf], (lumcmf * 1. / nprocs_exspec / inverse_gamma)[in_cmf])
        if direction_bins:
            np.add.at(out["lc_lum"], (1 + dirbin[arrives], nts[arrives]), (lum * float(MABINS) / nprocs_exspec)[arrives])
            np.add.at(out["lc_lumcmf"], (1 + dirbin[in_cmf], nts_cmf[in_cmf]),
                      (lumcmf * float(MABINS) / nprocs_exspec / inverse_gamma)[in_cmf])
        inspec = arrives & (p["nu_rf"] > nu_min) & (p["nu_rf"] < nu_max)
        q = p[inspec]
        qn = nts[inspec]
        qd = dirbin[inspec]
        nnu = _logbin(q["nu_rf"], nu_min, dlognu, nnubins)
        unit = q["e_rf"] / ts_width[qn] / delta[nnu].astype(np.float64) / 4.e12 / np.pi / PARSEC / PARSEC / nprocs_exspec
        np.add.at(out["flux"][0], (nnu, qn), unit * 1.)
        if direction_bins:
            np.add.at(out["flux"], (1 + qd, nnu, qn), unit * float(MABINS))
        if not emission_absorption:
            continue
        line_el, line_ion = static["line.elementindex"], static["line.ionindex"]
        bf_col = _bflist_columns(static, max_nions)
        for field, dest in (("trueemissiontype", "trueemission"), ("emissiontype", "emission")):
            et = q[field].astype(np.int64)
            col = np.full(et.size, -1, dtype=np.int64)
            bb = et >= 0
            col[bb] = line_el[et[bb]].astype(np.int64) * max_nions + line_ion[et[bb]]
            col[et == EMTYPE_FREEFREE] = 2 * ioncount
            bf = (et < 0) & (et != EMTYPE_FREEFREE) & (et != EMTYPE_NOTSET)
            col[bf] = (ioncount + bf_col[-1 - et[bf]]) if bf_col.size > 0 else 2 * ioncount
            ok = col >= 0
            np.add.at(out[dest][0], (nnu[ok], qn[ok], col[ok]), unit[ok] * 1.)
        absfreq = q["absorptionfreq"]
        at = q["absorptiontype"].astype(np.int64)
        with np.errstate(invalid="ignore", divide="ignore"):
            isabs = (absfreq > nu_min) & (absfreq < nu_max) & (at >= 0)
            nnu_abs = _logbin(np.where(isabs, absfreq, nu_min * 2.), nu_min, dlognu, nnubins)
        unit_abs = q["e_rf"] / ts_width[qn] / delta[nnu_abs].astype(np.float64) / 4.e12 / np.pi / PARSEC / PARSEC / nprocs_exspec
        col = line_el[at[isabs]].astype(np.int64) * max_nions + line_ion[at[isabs]]
        np.add.at(out["absorption"][0], (nnu_abs[isabs], qn[isabs], col), unit_abs[isabs] * 1.)
    return out


def _bflist_columns(static, max_nions):
    """element * max_nions + ion of every entry of globals::bflist (input.cc:1765-1793): indexed by -1 - emissiontype"""
    nbf = static["cont.nu_edge"].size
    col = np.zeros(nbf, dtype=np.int64)
    e_start, e_nions = static["elem.uniqueionindexstart"], static["elem.nions"]
    for e in range(e_nions.size):
        for i in range(int(e_nions[e])):
            u = int(e_start[e]) + i
            for lev in range(int(static["ion.nlevels_ionising"][u])):
                ulev = int(static["ion.uniquelevelindexstart"][u]) + lev
                start = int(static["level.bflist_start"][ulev])
                for t in range(int(static["level.nphixstargets"][ulev])):
                    if 0 <= start and start + t < nbf:
                        col[start + t] = e * max_nions + i
    return col
