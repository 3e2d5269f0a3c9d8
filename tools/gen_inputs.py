#!/usr/bin/env python3
"""Seeded generator of synthetic ARTIS run folders in the reference's own input file formats.

There is no network, so the reference's test atomic-data tarballs (tests/setup_*.sh) are not
available; every workload is synthetic but goes through the reference's real parsers:
  input.txt            24 positional lines          (reference input.cc:1862-2056)
  model.txt            1D / 2D / 3D ejecta model    (reference grid.cc:1931-2197, 749-927)
  abundances.txt       per-cell elemental fractions (reference grid.cc:669-747)
  compositiondata.txt  element / ion list           (reference input.cc:1306-1362)
  adata.txt, transitiondata.txt                     (reference input.cc:410-544, 1364-1482)
  phixsdata_v2.txt     photoionisation tables       (reference input.cc:296-408)

Usage: python tools/gen_inputs.py <config> <outdir> [--options-out artisoptions.h --reference /root/reference]
"""
import argparse
import math
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import configs  # noqa: E402

MSUN = 1.98855e33
CLIGHT = 2.99792458e10
DAY = 86400.0
NPHIXSPOINTS = 100
NPHIXSNUINCREMENT = 0.1


def ionpot_ev(Z, stage, dz=0.0):
    """smooth synthetic ionisation potential [eV] of ion `stage` (1 = neutral). `dz` (config atomic.ionpot_dz) adds a term
    that grows with Z: element sets in which two elements share Z % 9 (O and Fe) would otherwise get identical ionisation
    edges, which the reference's ground-continuum search rejects (input.cc:703-747)"""
    return 7.4 * stage**1.28 * (1.0 + 0.013 * (Z % 9) + dz * Z)


def write_atomic(cfg, outdir):
    a = cfg["atomic"]
    rng = random.Random(a["seed"])
    elements = a["elements"]
    nions = a["nions"]
    with open(os.path.join(outdir, "compositiondata.txt"), "w") as f:
        f.write(f"{len(elements)}\n0\n0\n")
        for Z, amu in elements:
            # Z nions lowest_ionstage highest_ionstage nlevelsmax(-1 = all) abundance(unused) mass_amu
            f.write(f"{Z} {nions} 1 {nions} -1 0.0 {amu}\n")
    ad = open(os.path.join(outdir, "adata.txt"), "w")
    tr = open(os.path.join(outdir, "transitiondata.txt"), "w")
    ph = open(os.path.join(outdir, "phixsdata_v2.txt"), "w")
    ph.write(f"{NPHIXSPOINTS}\n {NPHIXSNUINCREMENT:.1e}\n")
    nlines_total = 0
    for Z, _amu in elements:
        for ion in range(nions):
            stage = ion + 1
            ip = ionpot_ev(Z, stage, a.get("ionpot_dz", 0.0))
            # fewer levels in the higher ions, as in real data sets
            nlev = max(3, int(round(a["nlevels"] * (1.0 - 0.18 * ion))))
            en = [0.0] + sorted(rng.uniform(0.05, 0.82 * ip) for _ in range(nlev - 1))
            g = [rng.choice([1, 3, 5, 7, 9]) for _ in range(nlev)]
            ad.write(f"{Z} {stage} {nlev} {ip:.5f}\n")
            for lev in range(nlev):
                ad.write(f"{lev + 1} {en[lev]:.6f} {g[lev]} 0\n")
            ad.write("\n")
            pairs = [(lo, up) for lo in range(nlev) for up in range(lo + 1, nlev) if rng.random() < a["trans_frac"]]
            # every excited level keeps at least one downward transition so that it can deexcite
            have_down = {up for _, up in pairs}
            for up in range(1, nlev):
                if up not in have_down:
                    pairs.append((rng.randrange(0, up), up))
            pairs.sort()
            tr.write(f"{Z} {stage} {len(pairs)}\n")
            for lo, up in pairs:
                forb = rng.random() < a.get("forb_frac", 0.3)
                if "f_perm_log10" in a:
                    # physically consistent permitted lines: draw the absorption oscillator strength and derive
                    # A_ul = 6.670e15 f_lu (g_l / g_u) / lambda[Angstrom]^2, so that low-energy transitions get small A
                    # (drawing A independently of the wavelength gives f >> 1 for closely spaced levels, and with it
                    # absurd van Regemorter collision rates)
                    lo_f, hi_f = a["f_perm_log10"]
                    f_lu = 10 ** rng.uniform(lo_f, hi_f)
                    lam_angstrom = 12398.42 / max(en[up] - en[lo], 1e-4)
                    a_perm = 6.670e15 * f_lu * g[lo] / g[up] / lam_angstrom**2
                else:
                    lo_a, hi_a = a.get("A_perm_log10", (5.0, 8.5))
                    a_perm = 10 ** rng.uniform(lo_a, hi_a)
                A = 10 ** rng.uniform(-2, 1) if forb else a_perm
                tr.write(f"{lo + 1} {up + 1} {A:.4e} {-2.0 if forb else -1.0} {1 if forb else 0}\n")
            tr.write("\n")
            nlines_total += len(pairs)
            if ion < nions - 1:  # the top ion has no photoionisation data
                for lev in range(nlev):
                    # Z upperionstage upperlevel lowerionstage lowerlevel threshold_eV
                    ph.write(f"{Z} {stage + 1} 1 {stage} {lev + 1} {ip - en[lev]:.5f}\n")
                    s0 = rng.uniform(1.0, 10.0)
                    ph.write(" ".join(f"{s0 * (1 + NPHIXSNUINCREMENT * i) ** -3:.4e}" for i in range(NPHIXSPOINTS)) + "\n")
    ad.close()
    tr.close()
    ph.close()
    return nlines_total


def _abund_line(cellid, fracs, zmax):
    ab = [0.0] * zmax
    for Z, x in fracs.items():
        ab[Z - 1] = x
    return f"{cellid} " + " ".join(f"{x:.6e}" for x in ab) + "\n"


def write_model_1d(cfg, outdir):
    m = cfg["model"]
    elements = [Z for Z, _ in cfg["atomic"]["elements"]]
    ncell = m["ncell"]
    t_model = m["t_model_days"] * DAY
    vouts = [m["vmax_kmps"] * (i + 1) / ncell for i in range(ncell)]
    rho_shape = [math.exp(-v / m["v_e_kmps"]) for v in vouts]
    if m.get("rho0"):
        rhos = [m["rho0"] * s for s in rho_shape]
    else:
        mass = 0.0
        for i, v in enumerate(vouts):
            r_out = v * 1e5 * t_model
            r_in = (vouts[i - 1] * 1e5 * t_model) if i else 0.0
            mass += rho_shape[i] * 4.0 / 3.0 * math.pi * (r_out**3 - r_in**3)
        rhos = [s * m["mass_msun"] * MSUN / mass for s in rho_shape]
    with open(os.path.join(outdir, "model.txt"), "w") as f, open(os.path.join(outdir, "abundances.txt"), "w") as fa:
        f.write(f"{ncell}\n{m['t_model_days']}\n")
        for i in range(ncell):
            frac_in = i / ncell
            xni = 0.6 / (1.0 + math.exp((frac_in - 0.45) * 14.0)) + 0.02
            xfegrp = min(0.95, xni + 0.2)
            # id v_out[km/s] log10(rho) X_Fegroup X_Ni56 X_Co56 X_Fe52 X_Cr48
            f.write(f"{i + 1} {vouts[i]:.6e} {math.log10(rhos[i]):.6f} {xfegrp:.4f} {xni:.4f} 0.0 0.0 0.0\n")
            fracs = {}
            heavy = [Z for Z in elements if Z >= 26]
            light = [Z for Z in elements if Z < 26]
            # Ni must hold at least the Ni56 mass fraction (reference grid.cc:736)
            fracs[28] = xni + 0.03
            rest_heavy = max(0.0, xfegrp - fracs[28])
            others = [Z for Z in heavy if Z != 28]
            for Z in others:
                fracs[Z] = rest_heavy / len(others) + 0.01
            remaining = 1.0 - sum(fracs.values())
            if light:
                for Z in light:
                    fracs[Z] = remaining / len(light)
            else:
                fracs[14] = remaining  # silicon filler (not in the atomic data set; still counts towards mass)
            fa.write(_abund_line(i + 1, fracs, 30))


_KN_NUCLIDES = ["Sr92", "Ce143", "Nd147", "I131", "Te132", "Ra224", "Cf254"]


def write_model_2d(cfg, outdir):
    m = cfg["model"]
    rng = random.Random(m["seed"])
    elements = [Z for Z, _ in cfg["atomic"]["elements"]]
    nr, nz = m["nr"], m["nz"]
    t_model = m["t_model_days"] * DAY
    vmax = m["vmax_c"] * CLIGHT
    rmax = vmax * t_model
    cells = []
    mass = 0.0
    for iz in range(nz):
        for ir in range(nr):
            r_mid = (ir + 0.5) * rmax / nr
            z_mid = rmax * (-1.0 + 2.0 * (iz + 0.5) / nz)
            rad = math.hypot(r_mid, z_mid) / rmax
            if rad > 0.97:
                rho = 0.0
            else:
                torus = math.exp(-(((r_mid / rmax - 0.35) / 0.22) ** 2) - ((z_mid / rmax) / 0.18) ** 2)
                polar = 0.25 * math.exp(-((rad / 0.55) ** 2)) * (abs(z_mid) / (math.hypot(r_mid, z_mid) + 1e-30)) ** 2
                rho = torus + polar + 0.02 * math.exp(-rad / 0.3)
            vol = math.pi * (((ir + 1) * rmax / nr) ** 2 - (ir * rmax / nr) ** 2) * (2.0 * rmax / nz)
            mass += rho * vol
            cells.append((r_mid, z_mid, rho, rad))
    scale = m["mass_msun"] * MSUN / mass
    zmax = max(92, max(elements))
    with open(os.path.join(outdir, "model.txt"), "w") as f, open(os.path.join(outdir, "abundances.txt"), "w") as fa:
        f.write(f"{nr} {nz}\n{m['t_model_days']}\n{vmax:.10e}\n")
        f.write("#inputcellid pos_rcyl_mid pos_z_mid rho X_Fegroup X_Ni56 X_Co56 X_Fe52 X_Cr48 "
                + " ".join("X_" + n for n in _KN_NUCLIDES) + " Ye\n")
        for idx, (r_mid, z_mid, rho, rad) in enumerate(cells):
            polarness = abs(z_mid) / (math.hypot(r_mid, z_mid) + 1e-30)
            ye = 0.15 + 0.25 * polarness**2
            xr = {n: 0.0 for n in _KN_NUCLIDES}
            if rho > 0:
                xr["Sr92"] = 0.010 * (1 + 0.2 * rng.random())
                xr["Ce143"] = 0.012
                xr["Nd147"] = 0.015
                xr["I131"] = 0.010
                xr["Te132"] = 0.010
                xr["Ra224"] = 0.004 * (1.0 - polarness)
                xr["Cf254"] = 0.0015 * (1.0 - polarness)
            f.write(f"{idx + 1} {r_mid:.7e} {z_mid:.7e} {rho * scale:.6e} 0.02 0.005 0.0 0.0 0.0 "
                    + " ".join(f"{xr[n]:.5e}" for n in _KN_NUCLIDES) + f" {ye:.4f}\n")
            fracs = {26: 0.03, 38: 0.12 + 0.2 * polarness, 58: 0.10, 60: 0.12, 92: 0.03 * (1 - polarness) + 0.005}
            fracs = {Z: x for Z, x in fracs.items() if Z in elements}
            tot = sum(fracs.values())
            fracs[34] = max(0.0, 1.0 - tot)  # selenium filler: not in the atomic data set
            fa.write(_abund_line(idx + 1, fracs, zmax))


def write_model_3d(cfg, outdir):
    m = cfg["model"]
    elements = [Z for Z, _ in cfg["atomic"]["elements"]]
    n = m["n"]
    t_model = m["t_model_days"] * DAY
    vmax = m["vmax_kmps"] * 1e5
    xmax = vmax * t_model
    w = 2.0 * xmax / n
    cells = []
    mass = 0.0
    for iz in range(n):
        for iy in range(n):
            for ix in range(n):
                x0, y0, z0 = -xmax + ix * w, -xmax + iy * w, -xmax + iz * w
                xc, yc, zc = x0 + w / 2, y0 + w / 2, z0 + w / 2
                # ellipsoidal density with an off-centre Ni blob
                s = math.sqrt((xc / 1.0) ** 2 + (yc / 0.85) ** 2 + (zc / 0.7) ** 2) / xmax
                rho = math.exp(-s / 0.14) if s < 0.93 else 0.0
                blob = math.exp(-(((xc / xmax - 0.2) ** 2 + (yc / xmax) ** 2 + (zc / xmax + 0.1) ** 2) / 0.05))
                mass += rho * w**3
                cells.append((x0, y0, z0, rho, blob, s))
    scale = m["mass_msun"] * MSUN / mass
    with open(os.path.join(outdir, "model.txt"), "w") as f, open(os.path.join(outdir, "abundances.txt"), "w") as fa:
        f.write(f"{n ** 3}\n{m['t_model_days']}\n{vmax:.10e}\n")
        for idx, (x0, y0, z0, rho, blob, s) in enumerate(cells):
            xni = min(0.8, 0.05 + 0.7 * blob + 0.3 * math.exp(-s / 0.2))
            xfegrp = min(0.95, xni + 0.15)
            f.write(f"{idx + 1} {x0:.7e} {y0:.7e} {z0:.7e} {rho * scale:.6e} {xfegrp:.4f} {xni:.4f} 0.0 0.0 0.0\n")
            fracs = {28: xni + 0.02}
            others = [Z for Z in elements if Z >= 26 and Z != 28]
            for Z in others:
                fracs[Z] = max(0.0, xfegrp - fracs[28]) / len(others) + 0.01
            light = [Z for Z in elements if Z < 26]
            remaining = 1.0 - sum(fracs.values())
            if light:
                for Z in light:
                    fracs[Z] = remaining / len(light)
            else:
                fracs[14] = remaining
            fa.write(_abund_line(idx + 1, fracs, 30))


def write_input(cfg, outdir, resume=False):
    r = cfg["run"]
    dim = {"1d": 1, "2d": 2, "3d": 3}[cfg["model"]["kind"]]
    lines = [
        f"{r['seed']}", f"{r['ntimesteps']}", f"000 {r['nts_run']:03d}", f"{r['tmin']} {r['tmax']}",
        "1.33  1.330000001", "80", "3. 0.037", f"{dim}", "4", "1", "1.0", "-1", "0 0 1", "4", "1.0e-10", "-1",
        "1" if resume else "0", "1e-6", f"{r['nlte_ts']}", f"{r['thick']} {r['ngrey']}", "-1", "1", "1", "0.001 1000",
    ]
    assert len(lines) == 24
    with open(os.path.join(outdir, "input.txt"), "w") as f:
        for i, ln in enumerate(lines):
            f.write(f"{ln:<24s} # {i:2d}\n")


def write_options(cfg, reference_dir, outpath):
    """apply the config's compile-time overrides to the reference preset (same edits as tests/setup_*.sh)"""
    src = os.path.join(reference_dir, f"artisoptions_{cfg['preset']}.h")
    out = []
    todo = dict(cfg["opts"])
    with open(src) as f:
        for line in f:
            hit = None
            for prefix in todo:
                if line.startswith(prefix):
                    hit = prefix
                    break
            if hit:
                out.append(todo.pop(hit) + "\n")
            else:
                out.append(line)
    if todo:
        raise RuntimeError(f"options not found in {src}: {list(todo)}")
    with open(outpath, "w") as f:
        f.writelines(out)


# configs whose model files are too large to keep in the tree (hundreds of MB for 1e6 cells): their run folders are
# written where they are needed (bench.py on the GPU box, tools/run_oracle.py), not at build time
LARGE_MODELS = {"asym3d", "asym3d_cpu", "gamma_3d50", "gamma_3d50_cpu"}


def generate(name, outdir, reference_dir=None, options_out=None, data_link=None, options_only=False):
    cfg = configs.get(name)
    if options_only:
        os.makedirs(os.path.dirname(os.path.abspath(options_out)), exist_ok=True)
        write_options(cfg, reference_dir, options_out)
        return 0
    os.makedirs(outdir, exist_ok=True)
    nlines = write_atomic(cfg, outdir)
    {"1d": write_model_1d, "2d": write_model_2d, "3d": write_model_3d}[cfg["model"]["kind"]](cfg, outdir)
    write_input(cfg, outdir)
    if options_out:
        write_options(cfg, reference_dir, options_out)
    if data_link and not os.path.lexists(os.path.join(outdir, "data")):
        os.symlink(os.path.abspath(data_link), os.path.join(outdir, "data"))
    return nlines


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("outdir")
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--options-out", default=None)
    ap.add_argument("--data-link", default=None, help="directory to symlink as <outdir>/data (decay tables)")
    ap.add_argument("--options-only", action="store_true", help="write artisoptions.h only (no run folder)")
    args = ap.parse_args()
    n = generate(args.config, args.outdir, args.reference, args.options_out, args.data_link,
                 options_only=args.options_only or (args.config in LARGE_MODELS and args.options_out is not None))
    print(f"wrote {args.config} run folder to {args.outdir} ({n} lines in transitiondata)")
