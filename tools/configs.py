"""Named synthetic workloads (BASELINE.json `configs`, plus toy-sized variants for parity tests).

Every config names the reference preset (`artisoptions_<preset>.h`), the compile-time option
overrides applied to it (the same `sed` edits the reference's tests/setup_*.sh scripts make,
e.g. tests/setup_classicmode_1d_3dgrid.sh:27-34), and the parameters of the synthetic model and
atomic data that tools/gen_inputs.py writes in the reference's input file formats
(SURVEY.md Appendix B).
"""

# option override -> replacement source line (matched on the declaration prefix before '=' / ';')
def _opts(mpkts, grid_override=None, cuboid=None, extra=None):
    o = {"constexpr int MPKTS": f"constexpr int MPKTS = {int(mpkts)};"}
    if grid_override:
        o["constexpr std::optional<GridType> GRID_TYPE_OVERRIDE"] = (
            f"constexpr std::optional<GridType> GRID_TYPE_OVERRIDE = GridType::{grid_override};"
        )
    if cuboid:
        for ax in "XYZ":
            o[f"constexpr int CUBOID_NCOORDGRID_{ax}"] = f"constexpr int CUBOID_NCOORDGRID_{ax} = {int(cuboid)};"
    if extra:
        o.update(extra)
    return o


import os

_ASYM3D_N = int(os.environ.get("ARTISB200_ASYM3D_N", "100"))
_FEGROUP = [(26, 55.845), (27, 58.9332), (28, 58.6934)]
_CLASSIC_ELEMS = [(8, 15.999), (14, 28.085), (16, 32.06), (20, 40.078), (26, 55.845), (27, 58.9332), (28, 58.6934)]
_KN_ELEMS = [(26, 55.845), (38, 87.62), (58, 140.116), (60, 144.242), (92, 238.029)]
# same LUT temperature grid as the reference's own kilonova_2d test (tests/setup_kilonova_2d.sh:27-29)
_KN_LUT = {
    "constexpr int TABLESIZE": "constexpr int TABLESIZE = 20;",
    "constexpr double MINTEMP": "constexpr double MINTEMP = 1000.;",
    "constexpr double MAXTEMP": "constexpr double MAXTEMP = 20000.;",
}
# permitted lines get A from a drawn oscillator strength f in [1e-3, 0.3] (physically consistent collision rates)
_KN2D_ATOMIC = dict(elements=_KN_ELEMS, nions=4, nlevels=120, trans_frac=0.6, f_perm_log10=(-3.0, -0.5), seed=20260101)
_KN2D_MODEL = dict(kind="2d", nr=50, nz=100, vmax_c=0.3, t_model_days=0.1, mass_msun=0.02, seed=20260101)
# photospheric phase (2-12 d). Cells with grey optical depth >= 50 are treated in the grey approximation at all
# times (the reference's optical_depth_is_thick / num_grey_timesteps inputs, input.txt line 19; its classic test uses
# "8.0 999"); everywhere else packets get the detailed line-by-line / continuum / macro-atom treatment.
_KN2D_RUN = dict(seed=20260101, ntimesteps=30, tmin=2.0, tmax=12.0, nts_run=6, thick=50.0, ngrey=999, nlte_ts=999)

CONFIGS = {
    # ---- toy-sized parity cases (run in seconds on one CPU core) -------------------------------
    "classic_toy": dict(
        preset="classic",
        opts=_opts(1500, "CARTESIAN3D", 20),
        atomic=dict(elements=_FEGROUP, nions=3, nlevels=6, trans_frac=1.0, seed=1),
        model=dict(kind="1d", ncell=20, vmax_kmps=20000.0, t_model_days=2.0, rho0=2e-11, v_e_kmps=3000.0, seed=1),
        run=dict(seed=8, ntimesteps=8, tmin=4.0, tmax=40.0, nts_run=5, thick=8.0, ngrey=2, nlte_ts=5),
    ),
    "classic_toy_1d": dict(
        preset="classic",
        opts=_opts(1500),
        atomic=dict(elements=_FEGROUP, nions=3, nlevels=6, trans_frac=1.0, seed=1),
        model=dict(kind="1d", ncell=20, vmax_kmps=20000.0, t_model_days=2.0, rho0=2e-11, v_e_kmps=3000.0, seed=1),
        run=dict(seed=8, ntimesteps=8, tmin=4.0, tmax=40.0, nts_run=5, thick=8.0, ngrey=2, nlte_ts=5),
    ),
    # classic_toy_1d with the multi-bin radiation field model of the NLTE presets switched on from timestep 1
    "classic_multibin_toy": dict(
        preset="classic",
        opts=_opts(1500, None, None, {
            "constexpr bool MULTIBIN_RADFIELD_MODEL_ON": "constexpr bool MULTIBIN_RADFIELD_MODEL_ON = true;",
            "constexpr int FIRST_NLTE_RADFIELD_TIMESTEP": "constexpr int FIRST_NLTE_RADFIELD_TIMESTEP = 1;",
        }),
        atomic=dict(elements=_FEGROUP, nions=3, nlevels=6, trans_frac=1.0, seed=1),
        model=dict(kind="1d", ncell=20, vmax_kmps=20000.0, t_model_days=2.0, rho0=2e-11, v_e_kmps=3000.0, seed=1),
        run=dict(seed=8, ntimesteps=8, tmin=4.0, tmax=40.0, nts_run=5, thick=8.0, ngrey=2, nlte_ts=1),
    ),
    # classic_toy_1d with three NLTE excited levels (+ superlevel) in Fe II: the hot path reads the NLTE solver's populations
    "classic_nlte_toy": dict(
        preset="classic",
        opts=_opts(1500, None, None, {
            "constexpr int ION_NLEVELS_EXCITED_NLTE": "constexpr int ION_NLEVELS_EXCITED_NLTE(int element_z, int ionstage) "
                                                      "{ return (element_z == 26 && ionstage == 2) ? 3 : 0; }",
        }),
        atomic=dict(elements=_FEGROUP, nions=3, nlevels=6, trans_frac=1.0, seed=1),
        model=dict(kind="1d", ncell=20, vmax_kmps=20000.0, t_model_days=2.0, rho0=2e-11, v_e_kmps=3000.0, seed=1),
        run=dict(seed=8, ntimesteps=8, tmin=4.0, tmax=40.0, nts_run=5, thick=8.0, ngrey=2, nlte_ts=1),
    ),
    # classic_toy_1d with non-thermal deposition solved by Spencer-Fano (routing of deposited leptons to ionisation)
    "classic_nt_toy": dict(
        preset="classic",
        opts=_opts(1500, None, None, {
            "constexpr bool NT_ON": "constexpr bool NT_ON = true;",
            "constexpr bool NT_SOLVE_SPENCERFANO": "constexpr bool NT_SOLVE_SPENCERFANO = true;",
        }),
        atomic=dict(elements=_FEGROUP, nions=3, nlevels=6, trans_frac=1.0, seed=1),
        model=dict(kind="1d", ncell=20, vmax_kmps=20000.0, t_model_days=2.0, rho0=2e-11, v_e_kmps=3000.0, seed=1),
        run=dict(seed=8, ntimesteps=8, tmin=4.0, tmax=40.0, nts_run=4, thick=8.0, ngrey=2, nlte_ts=1),
    ),
    "classic_ntexc_toy": dict(
        preset="classic",
        opts=_opts(1500, None, None, {
            "constexpr bool NT_ON": "constexpr bool NT_ON = true;",
            "constexpr bool NT_SOLVE_SPENCERFANO": "constexpr bool NT_SOLVE_SPENCERFANO = true;",
            "constexpr bool NT_EXCITATION_ON": "constexpr bool NT_EXCITATION_ON = true;",
        }),
        atomic=dict(elements=_FEGROUP, nions=3, nlevels=6, trans_frac=1.0, seed=1),
        model=dict(kind="1d", ncell=20, vmax_kmps=20000.0, t_model_days=2.0, rho0=2e-11, v_e_kmps=3000.0, seed=1),
        run=dict(seed=8, ntimesteps=8, tmin=4.0, tmax=40.0, nts_run=4, thick=8.0, ngrey=2, nlte_ts=1),
    ),
    # classic_toy_1d with the detailed bound-free estimators of the NLTE presets (no photoionisation LUT)
    "classic_detailedbf_toy": dict(
        preset="classic",
        opts=_opts(1500, None, None, {
            "constexpr bool DETAILED_BF_ESTIMATORS_ON": "constexpr bool DETAILED_BF_ESTIMATORS_ON = true;",
            "constexpr bool USE_LUT_PHOTOION": "constexpr bool USE_LUT_PHOTOION = false;",
            "constexpr int DETAILED_BF_ESTIMATORS_USEFROMTIMESTEP": "constexpr int DETAILED_BF_ESTIMATORS_USEFROMTIMESTEP = 1;",
            "constexpr bool LEVEL_HAS_BFEST": "constexpr bool LEVEL_HAS_BFEST(int element_z, int ionstage, int level) { return level <= 2; }",
        }),
        atomic=dict(elements=_FEGROUP, nions=3, nlevels=6, trans_frac=1.0, seed=1),
        model=dict(kind="1d", ncell=20, vmax_kmps=20000.0, t_model_days=2.0, rho0=2e-11, v_e_kmps=3000.0, seed=1),
        run=dict(seed=8, ntimesteps=8, tmin=4.0, tmax=40.0, nts_run=4, thick=8.0, ngrey=2, nlte_ts=1),
    ),
    # BASELINE configs[4] in miniature: the reference's NLTE photospheric preset unchanged (NLTE levels, multi-bin radiation
    # field, detailed bound-free estimators, Spencer-Fano non-thermal deposition with excitation) on the 1D toy model
    "nltephot_toy": dict(
        preset="nltephotospheric_dynamic_ion_range",
        opts=_opts(1500),
        atomic=dict(elements=_FEGROUP, nions=3, nlevels=6, trans_frac=1.0, seed=1),
        model=dict(kind="1d", ncell=20, vmax_kmps=20000.0, t_model_days=2.0, rho0=2e-11, v_e_kmps=3000.0, seed=1),
        run=dict(seed=8, ntimesteps=8, tmin=4.0, tmax=40.0, nts_run=4, thick=8.0, ngrey=2, nlte_ts=1),
    ),
    "kilonova_toy": dict(
        preset="kilonova_lte",
        opts=_opts(1000, None, None, {
            "constexpr int TABLESIZE": "constexpr int TABLESIZE = 20;",
            "constexpr double MINTEMP": "constexpr double MINTEMP = 1000.;",
            "constexpr double MAXTEMP": "constexpr double MAXTEMP = 20000.;",
        }),
        atomic=dict(elements=_KN_ELEMS, nions=3, nlevels=8, trans_frac=0.6, seed=2),
        model=dict(kind="2d", nr=8, nz=16, vmax_c=0.3, t_model_days=0.1, mass_msun=0.01, seed=2),
        run=dict(seed=9, ntimesteps=10, tmin=0.2, tmax=6.0, nts_run=5, thick=0.0, ngrey=2, nlte_ts=999),
    ),
    # kilonova_toy with the parameterised thermalisation schemes (gammapkt.cc:752-858, update_packets.cc:57-84)
    **{f"kilonova_{tag}_toy": dict(
        preset="kilonova_lte",
        opts=_opts(1000, None, None, {
            "constexpr int TABLESIZE": "constexpr int TABLESIZE = 20;",
            "constexpr double MINTEMP": "constexpr double MINTEMP = 1000.;",
            "constexpr double MAXTEMP": "constexpr double MAXTEMP = 20000.;",
            "constexpr auto GAMMA_THERMALISATION_SCHEME": f"constexpr auto GAMMA_THERMALISATION_SCHEME = GammaThermalisationScheme::{gam};",
            "constexpr auto PARTICLE_THERMALISATION_SCHEME": f"constexpr auto PARTICLE_THERMALISATION_SCHEME = ParticleThermalisationScheme::{par};",
        }),
        atomic=dict(elements=_KN_ELEMS, nions=3, nlevels=8, trans_frac=0.6, seed=2),
        model=dict(kind="2d", nr=8, nz=16, vmax_c=0.3, t_model_days=0.1, mass_msun=0.01, seed=2),
        run=dict(seed=9, ntimesteps=10, tmin=0.2, tmax=6.0, nts_run=5, thick=0.0, ngrey=2, nlte_ts=999),
    ) for tag, gam, par in (("guttman", "GUTTMAN", "BARNES"), ("wollaeger", "WOLLAEGER", "WOLLAEGER"),
                            ("barnes", "BARNES", "INSTANTFULLDEPOSITION"))},
    # kilonova_toy with the expansion-opacity / bound-bound thermalisation r-packet modes (rpkt.cc:221-320, 628-651,
    # 964-981): the reference's CI variant (expansion opacities, every bound-bound event thermalises), expansion opacities
    # with the line-by-line re-trace, and a thermalisation probability on the line-by-line opacity
    **{f"kilonova_{tag}_toy": dict(
        preset="kilonova_lte",
        opts=_opts(1000, None, None, {
            "constexpr int TABLESIZE": "constexpr int TABLESIZE = 20;",
            "constexpr double MINTEMP": "constexpr double MINTEMP = 1000.;",
            "constexpr double MAXTEMP": "constexpr double MAXTEMP = 20000.;",
            "constexpr bool RPKT_USE_EXPANSION_OPACITIES": f"constexpr bool RPKT_USE_EXPANSION_OPACITIES = {expo};",
            "constexpr std::optional<float> RPKT_BOUNDBOUND_THERMALISATION_PROBABILITY":
                f"constexpr std::optional<float> RPKT_BOUNDBOUND_THERMALISATION_PROBABILITY{prob};",
        }),
        atomic=dict(elements=_KN_ELEMS, nions=3, nlevels=8, trans_frac=0.6, seed=2),
        model=dict(kind="2d", nr=8, nz=16, vmax_c=0.3, t_model_days=0.1, mass_msun=0.01, seed=2),
        run=dict(seed=9, ntimesteps=10, tmin=0.2, tmax=6.0, nts_run=5, thick=0.0, ngrey=2, nlte_ts=999),
    ) for tag, expo, prob in (("expansionopac", "true", " = 1."), ("expopac_retrace", "true", ""), ("bbtherm", "false", " = 0.5"))},
    # BASELINE configs[3] in miniature: 3-D model from 20 d with every cell grey for r-packets (optical_depth_is_thick = 0):
    # pellets, gamma-ray transport and deposition, k-packets and grey random walks
    "classic3d_grey_toy": dict(
        preset="classic",
        opts=_opts(1500),
        atomic=dict(elements=_FEGROUP, nions=3, nlevels=6, trans_frac=1.0, seed=1),
        model=dict(kind="3d", n=10, vmax_kmps=25000.0, t_model_days=2.0, mass_msun=1.4, seed=3),
        run=dict(seed=10, ntimesteps=30, tmin=20.0, tmax=80.0, nts_run=3, thick=0.0, ngrey=999, nlte_ts=999),
    ),
    # kilonova_toy with the XCOM photoionisation tables for the gamma-ray photoelectric opacity (the reference's CI variant
    # tests/setup_kilonova_2d_xcomgammaphotoion.sh)
    "kilonova_xcom_toy": dict(
        preset="kilonova_lte",
        opts=_opts(1000, None, None, {
            "constexpr int TABLESIZE": "constexpr int TABLESIZE = 20;",
            "constexpr double MINTEMP": "constexpr double MINTEMP = 1000.;",
            "constexpr double MAXTEMP": "constexpr double MAXTEMP = 20000.;",
            "constexpr bool USE_XCOM_GAMMAPHOTOION": "constexpr bool USE_XCOM_GAMMAPHOTOION = true;",
        }),
        atomic=dict(elements=_KN_ELEMS, nions=3, nlevels=8, trans_frac=0.6, seed=2),
        model=dict(kind="2d", nr=8, nz=16, vmax_c=0.3, t_model_days=0.1, mass_msun=0.01, seed=2),
        run=dict(seed=9, ntimesteps=10, tmin=0.2, tmax=6.0, nts_run=5, thick=0.0, ngrey=2, nlte_ts=999),
        extra_files=["xcom_photoion_data.txt"],
    ),
    "classic3d_toy": dict(
        preset="classic",
        opts=_opts(1500),
        atomic=dict(elements=_FEGROUP, nions=3, nlevels=6, trans_frac=1.0, seed=1),
        model=dict(kind="3d", n=10, vmax_kmps=20000.0, t_model_days=2.0, mass_msun=0.6, seed=3),
        run=dict(seed=10, ntimesteps=8, tmin=4.0, tmax=40.0, nts_run=5, thick=8.0, ngrey=2, nlte_ts=5),
    ),
    # ---- BASELINE.json configs (bench-sized) ----------------------------------------------------
    # configs[0]: classic LTE W7-like 1D model on a 3D grid, 1e5 packets
    "classic_1d3d": dict(
        preset="classic",
        opts=_opts(100000, "CARTESIAN3D", 100),
        atomic=dict(elements=_CLASSIC_ELEMS, nions=4, nlevels=40, trans_frac=0.15, ionpot_dz=0.002, seed=20260101),
        model=dict(kind="1d", ncell=100, vmax_kmps=25000.0, t_model_days=2.0, rho0=None, mass_msun=1.4,
                   v_e_kmps=2700.0, seed=20260101),
        run=dict(seed=20260101, ntimesteps=60, tmin=2.0, tmax=80.0, nts_run=12, thick=8.0, ngrey=3, nlte_ts=5),
    ),
    # configs[1]: kilonova LTE 2D cylindrical r-process ejecta, 1e7 packets  (the bench workload)
    "kilonova_2d": dict(preset="kilonova_lte", opts=_opts(10000000, None, None, _KN_LUT), atomic=_KN2D_ATOMIC, model=_KN2D_MODEL,
                      run=_KN2D_RUN),
    # reduced packet count variant of configs[1] used for CPU-side parity and the cpu_baseline sample
    # the bounded CPU sample of configs[1] timed by bench.py's cpu_baseline / reference arm
    "kilonova_2d_cpu": dict(preset="kilonova_lte", opts=_opts(100000, None, None, _KN_LUT), atomic=_KN2D_ATOMIC, model=_KN2D_MODEL,
                      run=_KN2D_RUN),
    # configs[1] at full atomic-data and grid size with few packets: bench-scale known-answer vectors (54 892 lines, 1 475
    # continua, 3 684 cells) and packet histories from the reference's parity build (tests/golden/kilonova_2d_kat_*)
    "kilonova_2d_kat": dict(preset="kilonova_lte", opts=_opts(2000, None, None, _KN_LUT), atomic=_KN2D_ATOMIC, model=_KN2D_MODEL,
                      run=_KN2D_RUN),
    # few-packet probe of configs[1] used while tuning the synthetic atomic data (interactions per packet per timestep)
    "kilonova_2d_probe": dict(preset="kilonova_lte", opts=_opts(2000, None, None, _KN_LUT), atomic=_KN2D_ATOMIC, model=_KN2D_MODEL,
                      run=_KN2D_RUN),
    # configs[2]: 3-D Cartesian 100^3 asymmetric SN Ia model (ellipsoidal density, off-centre Ni blob), classic macro-atom
    # mode. 1e6 packets per run here (BASELINE: 4e7): in this optically thick phase an active packet takes ~2e5
    # interactions per timestep (2.6e4 per packet averaged over all packets, most of which are still pellets), so 1e6 packets
    # are 2.6e10 interactions per step - 25 times the kilonova step. Measured timestep 4: the last of the LTE start-up phase (the host's per-cell temperature solve from timestep 5 on costs ~1 ms per cell and timestep).
    # ARTISB200_ASYM3D_N overrides the grid size (e.g. 40 for a quick look).
    "asym3d": dict(
        preset="classic",
        opts=_opts(1000000),
        atomic=dict(elements=_CLASSIC_ELEMS, nions=4, nlevels=40, trans_frac=0.15, ionpot_dz=0.002, seed=20260101),
        model=dict(kind="3d", n=_ASYM3D_N, vmax_kmps=25000.0, t_model_days=2.0, mass_msun=1.4, seed=20260101),
        run=dict(seed=20260101, ntimesteps=60, tmin=2.0, tmax=80.0, nts_run=12, thick=8.0, ngrey=3, nlte_ts=5),
    ),
    "asym3d_cpu": dict(
        preset="classic",
        opts=_opts(100000),
        atomic=dict(elements=_CLASSIC_ELEMS, nions=4, nlevels=40, trans_frac=0.15, ionpot_dz=0.002, seed=20260101),
        model=dict(kind="3d", n=_ASYM3D_N, vmax_kmps=25000.0, t_model_days=2.0, mass_msun=1.4, seed=20260101),
        run=dict(seed=20260101, ntimesteps=60, tmin=2.0, tmax=80.0, nts_run=12, thick=8.0, ngrey=3, nlte_ts=5),
    ),
    # configs[3]: gamma-packet-only Ni56/Co56 deposition run on a 3-D 50^3 grid: the packets are pellets and gamma rays
    # (Compton / photoelectric / pair transport and deposition); every cell grey for the r-packets (optical_depth_is_thick
    # = 0, num_grey_timesteps = 999) so that they stay cheap, as SURVEY.md 8d prescribes. Started at 20 d: at 2 d the ejecta
    # absorb every gamma ray at its first event and the timestep is all grey random walks (measured: 1.8e5 gamma events
    # against 7.9e8 grey scatterings); from 20 d on gamma rays Compton-scatter several times and some escape
    "gamma_3d50": dict(
        preset="classic",
        opts=_opts(10000000),
        atomic=dict(elements=_FEGROUP, nions=3, nlevels=6, trans_frac=1.0, seed=1),
        model=dict(kind="3d", n=50, vmax_kmps=25000.0, t_model_days=2.0, mass_msun=1.4, seed=20260101),
        run=dict(seed=20260101, ntimesteps=30, tmin=20.0, tmax=80.0, nts_run=6, thick=0.0, ngrey=999, nlte_ts=999),
    ),
    "gamma_3d50_cpu": dict(
        preset="classic",
        opts=_opts(100000),
        atomic=dict(elements=_FEGROUP, nions=3, nlevels=6, trans_frac=1.0, seed=1),
        model=dict(kind="3d", n=50, vmax_kmps=25000.0, t_model_days=2.0, mass_msun=1.4, seed=20260101),
        run=dict(seed=20260101, ntimesteps=30, tmin=20.0, tmax=80.0, nts_run=6, thick=0.0, ngrey=999, nlte_ts=999),
    ),
    # the stated stochastic test (tools/stochastic_ensemble.py): configs[0]'s model family and atomic data, 1e5 packets,
    # twelve timesteps over the photospheric phase in which a large part of the packets escapes (spectrum, light curve)
    "classic_spec": dict(
        preset="classic",
        opts=_opts(100000, "CARTESIAN3D", 40),
        atomic=dict(elements=_CLASSIC_ELEMS, nions=4, nlevels=40, trans_frac=0.15, ionpot_dz=0.002, A_perm_log10=(3.5, 7.0), seed=20260101),
        model=dict(kind="1d", ncell=40, vmax_kmps=25000.0, t_model_days=2.0, rho0=None, mass_msun=0.3,
                   v_e_kmps=3200.0, seed=20260101),
        run=dict(seed=20260101, ntimesteps=12, tmin=15.0, tmax=60.0, nts_run=12, thick=8.0, ngrey=2, nlte_ts=999),
    ),
    "classic_spec_probe": dict(
        preset="classic",
        opts=_opts(4000, "CARTESIAN3D", 40),
        atomic=dict(elements=_CLASSIC_ELEMS, nions=4, nlevels=40, trans_frac=0.15, ionpot_dz=0.002, A_perm_log10=(3.5, 7.0), seed=20260101),
        model=dict(kind="1d", ncell=40, vmax_kmps=25000.0, t_model_days=2.0, rho0=None, mass_msun=0.3,
                   v_e_kmps=3200.0, seed=20260101),
        run=dict(seed=20260101, ntimesteps=12, tmin=15.0, tmax=60.0, nts_run=12, thick=8.0, ngrey=2, nlte_ts=999),
    ),
    # configs[0] with the packet count of the CPU sample (the config itself is CPU-runnable: 1e5 packets)
    "classic_1d3d_cpu": dict(
        preset="classic",
        opts=_opts(100000, "CARTESIAN3D", 100),
        atomic=dict(elements=_CLASSIC_ELEMS, nions=4, nlevels=40, trans_frac=0.15, ionpot_dz=0.002, seed=20260101),
        model=dict(kind="1d", ncell=100, vmax_kmps=25000.0, t_model_days=2.0, rho0=None, mass_msun=1.4,
                   v_e_kmps=2700.0, seed=20260101),
        run=dict(seed=20260101, ntimesteps=60, tmin=2.0, tmax=80.0, nts_run=12, thick=8.0, ngrey=3, nlte_ts=5),
    ),
    "kilonova_2d_small": dict(preset="kilonova_lte", opts=_opts(200000, None, None, _KN_LUT), atomic=_KN2D_ATOMIC, model=_KN2D_MODEL,
                      run=_KN2D_RUN),
}


def get(name):
    if name not in CONFIGS:
        raise KeyError(f"unknown config {name!r}; known: {sorted(CONFIGS)}")
    c = dict(CONFIGS[name])
    c["name"] = name
    return c
