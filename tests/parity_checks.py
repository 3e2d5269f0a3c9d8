"""Parity assertions shared by the host-simulation tests (CPU, this container) and the GPU tests (B200, through
the C ABI). The checker is always the oracle fixture generated from the compiled reference (tests/golden)."""
import os

import numpy as np

import compare_run
from tests import fixtures

REL_TOL = 1e-12  # BASELINE.json north_star: deterministic doubles within 1e-12 relative, indices bit-exact


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300)
    return np.abs(a - b) / denom


def check_deterministic_kernels(libpath, config, nts, ks_draws=None):
    """boundary_distance / closest_transition / continuum opacity against the reference's golden vectors"""
    fx = fixtures.load_golden(config, nts)
    after = fx["after"]
    eng = fixtures.make_engine(libpath, fx)
    eng.set_array("scalar.max_path_step", after["kat.bd.max_path_step"])
    dist, nxt = eng.test_kernel("boundary_distance", after["kat.bd.in"], after["kat.bd.cell"])
    assert np.array_equal(nxt, after["kat.bd.next"]), "boundary_distance: next-cell index differs from the reference"
    assert _rel(dist, after["kat.bd.dist"]).max() <= REL_TOL
    _, ct = eng.test_kernel("closest_transition", after["kat.ct.nu"], after["kat.ct.next_trans"])
    assert np.array_equal(ct, after["kat.ct.out"]), "closest_transition: line index differs from the reference"
    n_chi = after["kat.chi.nu"].size
    if n_chi:
        eng.set_array("scalar.max_path_step", fx["before"]["scalar.max_path_step"])
        chi, _ = eng.test_kernel("chi_rpkt_cont", after["kat.chi.nu"], after["kat.chi.cell"])
        ref = after["kat.chi.out"]
        err = _rel(chi, ref)
        # chi_bf sums many terms whose stimulated-emission factors the reference caches lazily in evaluation order
        # (rpkt.cc:840-889): identical maths, last-bits summation differences -> same 1e-12 bound
        assert err.max() <= REL_TOL, f"continuum opacity differs from the reference by {err.max():.3e}"
    check_select_continuum_nu(eng, after, ks_draws)
    eng.close()
    return n_chi


def check_select_continuum_nu(eng, after, ks_draws=None):
    """free-bound emission frequencies (ratecoeff.cc:563-638) against reference-evaluated vectors: every vector within the
    reference's own integration accuracy (1e-3), and - because the device integrates with the same adaptive rule in the
    same order - nearly all of them to rounding; plus a two-sample KS test of the sampled distribution"""
    from scipy import stats as sps
    nu, _ = eng.test_kernel("select_continuum_nu", after["kat.sc.in"], after["kat.sc.cont"])
    ref = after["kat.sc.out"]
    err = _rel(nu, ref)
    assert err.max() <= 1e-3, f"select_continuum_nu differs from the reference by {err.max():.3e}"
    # (a subdivision decision of the adaptive rule can flip on another libm's last bit: allow a few at the 1e-3 level)
    assert np.count_nonzero(err > 1e-9) <= max(1, ref.size // 100), f"{np.count_nonzero(err > 1e-9)} of {ref.size} vectors not to rounding"
    setup = after["kat.scks.setup"]
    nks = after["kat.scks.out"].size // 2
    rng = np.random.default_rng(20260102)
    for c in range(2):
        nmine = min(nks, ks_draws or nks)  # (the single-threaded host build of the CPU suite draws fewer)
        zrand = 1. - rng.integers(0, 1 << 24, size=nmine).astype(np.float64) * 2.0 ** -24  # the packet draw: 24-bit uniform
        inp = np.stack([np.full(nmine, setup[2 * c + 1]), zrand], axis=1)
        mine, _ = eng.test_kernel("select_continuum_nu", inp, np.full(nmine, int(setup[2 * c]), dtype=np.int32))
        p = sps.ks_2samp(mine, after["kat.scks.out"][c * nks:(c + 1) * nks]).pvalue
        assert p > 0.01, f"select_continuum_nu: KS test against {nks} reference draws fails (p = {p:.2e}, case {c})"


def check_cell_tables(built, after):
    """device-built per-cell tables against the reference's own cell cache (update_packets.cc:397-464)"""
    pairs = [("built.levelpops", "ref.levelpops"), ("built.cont_nnlevel", "ref.cont_nnlevel"),
             ("built.chi_ff_nnionpart", "ref.chi_ff_nnionpart"), ("built.maprocessrates", "ref.maprocessrates"),
             ("built.matrans", "ref.matrans"), ("built.cooling_contrib", "ref.cooling_contrib")]
    for mine, ref in pairs:
        a, b = built[mine], after[ref]
        assert a.shape == b.shape, f"{mine}: shape {a.shape} vs reference {b.shape}"
        err = _rel(a, b)
        assert err.max() <= REL_TOL, f"{mine} differs from the reference cell cache by {err.max():.3e}"
    assert np.array_equal(built["built.cont_keepbits"], after["ref.cont_keepbits"]), "continuum keep-bitmaps differ"


def assert_aggregates(config, est, after, est_err, est_tol):
    """event counters, estimators, timestep scalars of a replayed timestep against the reference's"""
    assert int(est["counters"][fixtures.INTERACTIONS]) == int(after["counters"][fixtures.INTERACTIONS])
    # UPSCATTER / DOWNSCATTER classify a macro-atom emission by comparing the new comoving frequency with the one
    # the packet had when it was absorbed (macroatom.cc:227-232). After a resonance scattering the two agree to
    # the last bit or two, so the classification follows the rounding of the Doppler factor (device libm vs
    # glibc); every other counter, and the sum of the two, must be equal.
    UP, DOWN = 30, 31
    mine, ref_c = est["counters"].copy(), after["counters"].copy()
    assert int(mine[UP] + mine[DOWN]) == int(ref_c[UP] + ref_c[DOWN]), "up- plus down-scatterings differ from the reference"
    mine[[UP, DOWN]] = 0
    ref_c[[UP, DOWN]] = 0
    assert np.array_equal(mine, ref_c), f"event counters differ from the reference: {np.nonzero(mine != ref_c)[0]}"
    names = ["est.J", "est.nuJ", "est.ffheating", "est.colheating", "est.bfheating", "est.dep_gamma",
             "est.dep_positron", "est.dep_electron", "est.dep_alpha"]
    if "est.bfrate_raw" in after:  # DETAILED_BF_ESTIMATORS_ON fixtures
        assert "est.bfrate_raw" in est and after["est.bfrate_raw"].sum() > 0
        names += ["est.bfrate_raw"]
    if fixtures.PRESET_OF[config] not in fixtures.PRESETS_WITHOUT_LUT_PHOTOION:
        # (without the photoionisation LUT update_packets does not touch gammaestimator: the reference neither
        # zeroes it, sn3d.cc:729-732, nor adds to it, rpkt.cc:526-530; update_grid.cc:398-409 parks other values there)
        names += ["est.gamma"]
    if "est.bins_J_raw" in after:  # MULTIBIN_RADFIELD_MODEL_ON fixtures
        assert "est.bins_J_raw" in est and after["est.bins_J_raw"].sum() > 0
        names += ["est.bins_J_raw", "est.bins_nuJ_raw"]
    for name in names:
        assert est_err.get(name, 0.0) <= est_tol, f"{name}: {est_err[name]:.3e}"
    m = 9  # ts.scalars[9] (nt_energy_deposited) is file-static in the reference and not dumped
    ts_err = _rel(est["ts.scalars"][:m], after["ts.scalars"][:m])
    assert ts_err.max() <= est_tol, f"ts.scalars: {ts_err.max():.3e}"
    assert int(est["ts.pellet_decays"][0]) == int(after["ts.pellet_decays"][0])


def check_packet_histories(libpath, config, nts, max_steps=0, min_exact_fraction=1.0, tol=1e-9, est_tol=1e-9, options=None,
                           check_tables=True):
    """replay the timestep with every packet continuing its own reference RNG stream: histories must coincide"""
    fx = fixtures.load_golden(config, nts)
    pk, est, built, _ = fixtures.run_fixture(libpath, fx, rng="xoshiro", max_steps=max_steps, options=options)
    after = fx["after"]
    ref = fixtures.snap.packets_view(after)
    # Every packet must coincide, free-bound emissions included: select_continuum_nu integrates with the reference's own
    # adaptive Gauss-Kronrod rule in the reference's operation order (csrc/gk.h, emit.h), so the sampled frequency agrees
    # to rounding and the packet continues on the reference's history.
    n_fb_events = int(after["counters"][17]) + int(after["counters"][10])  # K_STAT_TO_R_FB + MA_STAT_DEACTIVATION_FB
    frac_ok, worst, est_err = compare_run.compare(pk, est, after, tol=tol, verbose=False)
    n_bad = int(round((1.0 - frac_ok) * len(ref)))
    allowed = int(np.ceil((1.0 - min_exact_fraction) * len(ref)))
    assert n_bad <= allowed, f"{n_bad} packets differ from the oracle ({n_fb_events} free-bound events) ({worst})"
    if min_exact_fraction == 1.0:
        if check_tables and not (options or {}).get("table_window_cells"):  # (windowed tables hold one group of cells)
            check_cell_tables(built, after)
            if "cell.corrphotoioncoeff" in fx["before"]:
                # USE_LUT_PHOTOION = false: the photoionisation coefficients evaluated on the device (previous timestep's
                # bound-free estimators / adaptive Gauss-Kronrod integral over the radiation field model) against the
                # reference's own get_corrphotoioncoeff (ratecoeff.cc:840-875) for every (cell, level, target)
                ref_gamma = fx["before"]["cell.corrphotoioncoeff"]
                assert ref_gamma.size > 0 and np.count_nonzero(ref_gamma) > 0
                err = _rel(built["built.corrphotoioncoeff"], ref_gamma)
                assert err.max() <= 1e-10, f"built.corrphotoioncoeff differs from the reference's get_corrphotoioncoeff by {err.max():.3e}"
        assert_aggregates(config, est, after, est_err, est_tol)
    n_fb = n_fb_events
    return frac_ok, n_fb, est


def check_table_windows(libpath, config, nts, window_cells, options=None):
    """cell-batched per-cell tables (the device form of the reference's cell-cache groups, update_packets.cc:468-524,
    574-612): with the tables of only `window_cells` cells resident at a time, packets wait for the pass that holds their
    cell - every packet history, counter and estimator must still coincide with the reference's"""
    opts = dict(options or {})
    opts["table_window_cells"] = window_cells
    _, _, est = check_packet_histories(libpath, config, nts, options=opts, check_tables=False)
    passes = int(est["diag"][12])  # ARTISB200_DIAG_TABLE_PASSES
    assert passes > 1, f"the run used {passes} table window pass(es): the windows were not exercised"
    return passes


def check_idempotence(libpath, config, nts, options=None):
    """update_packets on packets that have already been propagated to the end of the timestep changes nothing: not a
    byte of a packet (no step, no random number drawn), no estimator, no event counter (update_packets.cc:321-326: a
    packet is only processed while prop_time < ts_end and it has not escaped)"""
    fx = fixtures.load_golden(config, nts)
    eng = fixtures.make_engine(libpath, fx, rng="xoshiro", options=options)
    before = fx["before"]
    n = int(before["packets.count"][0])
    stride = int(before["packets.stride"][0])
    aos = before["packets.aos"].copy()
    eng.update_packets_host(nts, aos, n, stride)
    est1 = eng.estimators()
    again = aos.copy()
    eng.update_packets_host(nts, again, n, stride)
    est2 = eng.estimators()
    eng.close()
    assert np.array_equal(again, aos), "a second update_packets of the same timestep moved packets"
    for name, value in est1.items():
        if name == "diag":  # work counters: the second call still launches (empty) kernels
            continue
        assert np.array_equal(value, est2[name], equal_nan=True), f"{name} changed in a second update_packets of the same timestep"
    return n


def check_cell_tables_sampled(eng, static, after):
    """the same comparison for a bench-scale fixture, where the reference dumped its cell cache for a few cells only
    (after["ref.cells"]) and the device tables are read back cell by cell (artisb200_get_array_range)"""
    cells = after["ref.cells"]
    nlev = static["level.epsilon"].size
    nbf = static["cont.nu_edge"].size
    keepwords = (nbf + 63) // 64
    per_cell = {"levelpops": nlev, "maprocessrates": nlev * 9, "matrans": after["ref.matrans"].size // cells.size,
                "cooling_contrib": after["ref.cooling_contrib"].size // cells.size, "cont_nnlevel": nbf}
    for k, cell in enumerate(cells):
        for name, n in per_cell.items():
            mine = eng.get_array_range("built." + name, int(cell) * n, n)
            ref = after["ref." + name][k * n:(k + 1) * n]
            err = _rel(mine, ref)
            assert err.max() <= REL_TOL, f"built.{name} of cell {cell} differs from the reference cell cache by {err.max():.3e}"
        bits = eng.get_array_range("built.cont_keepbits", int(cell) * keepwords, keepwords, dtype=np.uint64)
        assert np.array_equal(bits, after["ref.cont_keepbits"][k * keepwords:(k + 1) * keepwords]), f"keep-bitmap of cell {cell} differs"
        chiff = eng.get_array_range("built.chi_ff_nnionpart", int(cell), 1)
        assert _rel(chiff, after["ref.chi_ff_nnionpart"][k:k + 1]).max() <= REL_TOL
    return cells.size


def check_bench_scale_histories(libpath, config, nts, options=None, tol=1e-9, est_tol=1e-9):
    """a fixture at the bench model's full atomic-data and grid size (few packets): sampled cell tables, every packet
    history, counters and estimators against the reference"""
    fx = fixtures.load_golden(config, nts)
    after, before = fx["after"], fx["before"]
    eng = fixtures.make_engine(libpath, fx, rng="xoshiro", options=options)
    n = int(before["packets.count"][0])
    stride = int(before["packets.stride"][0])
    aos = before["packets.aos"].copy()
    eng.update_packets_host(nts, aos, n, stride)
    est = eng.estimators()
    ncells = check_cell_tables_sampled(eng, fx["static"], after)
    eng.close()
    pk = aos.view(fixtures.snap.packet_dtype(stride))
    frac_ok, worst, est_err = compare_run.compare(pk, est, after, tol=tol, verbose=False)
    assert frac_ok == 1.0, f"{int(round((1 - frac_ok) * n))} of {n} packets differ from the oracle ({worst})"
    assert_aggregates(config, est, after, est_err, est_tol)
    return n, ncells


def check_spectra(libpath, config, nts, device=0, rel=1e-12):
    """spectra and light curves of the fixture's "after" packets binned by the library in ONE pass against the
    reference's own add_to_spec_res / add_to_lc_res run once per direction bin (tests/golden/make_golden_spectra.py).
    Bin indices (direction bin of every packet; which bins are filled) are exact; the sums agree to the rounding of the
    summation order: every addend is formed in the reference's operation order."""
    from artis_b200 import spectra as spectra_mod

    fx = fixtures.load_golden(config, nts)
    ref = dict(np.load(os.path.join(fixtures.GOLDEN_DIR, f"{config}_spectra_ts{nts}.npz")))
    eng = fixtures.ablib.ArtisB200(libpath=libpath, device=device)
    try:
        eng.set_arrays(fx["static"])
        eng.commit_static()
        after = fx["after"]
        n = int(after["packets.count"][0])
        stride = int(after["packets.stride"][0])
        eng.upload_packets(after["packets.aos"], n, stride)
        eng.set_option("spec_nnubins", int(ref["ref.spec.nnubins"][0]))
        eng.set_option("spec_record_dirbin", 1)
        nprocs = int(ref["ref.spec.nprocs_exspec"][0])
        mabins = int(ref["ref.spec.mabins"][0])
        assert mabins == spectra_mod.MABINS

        def close(name, got, want):
            got = np.asarray(got, dtype=np.float64).ravel()
            want = np.asarray(want, dtype=np.float64).ravel()
            assert got.shape == want.shape, f"{config} ts{nts} {name}: shape {got.shape} vs {want.shape}"
            assert np.array_equal(got != 0., want != 0.), f"{config} ts{nts} {name}: different bins are filled"
            err = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
            assert err.max(initial=0.) <= rel, f"{config} ts{nts} {name}: max relative error {err.max()}"

        # (1) everything at once: all direction bins + the decomposition of the angle-averaged set
        eng.bin_escaped_packets(direction_bins=True, emission_absorption=1, nprocs_exspec=nprocs)
        b = spectra_mod.binned(eng)
        assert np.array_equal(b["lower_freq"], ref["ref.spec.lower_freq"]) and np.array_equal(b["delta_freq"], ref["ref.spec.delta_freq"])
        assert np.array_equal(b["dirbin"], ref["ref.spec.dirbin"]), f"{config} ts{nts}: direction bins differ"
        assert b["flux"].shape[0] == 1 + mabins and b["emission"].shape[0] == 1
        close("flux", b["flux"][0], ref["ref.spec.flux"])
        close("flux_res", b["flux"][1:], ref["ref.spec.flux_res"])
        close("emission", b["emission"][0], ref["ref.spec.emission"])
        close("trueemission", b["trueemission"][0], ref["ref.spec.trueemission"])
        close("absorption", b["absorption"][0], ref["ref.spec.absorption"])
        close("lc_lum", b["lc_lum"][0], ref["ref.lc.lum"])
        close("lc_lumcmf", b["lc_lumcmf"][0], ref["ref.lc.lumcmf"])
        close("lc_lum_res", b["lc_lum"][1:], ref["ref.lc.lum_res"])
        close("lc_lumcmf_res", b["lc_lumcmf"][1:], ref["ref.lc.lumcmf_res"])
        close("gamma_lc_lum", b["gamma_lc_lum"], ref["ref.lc.gamma_lum"])
        close("gamma_lc_lumcmf", b["gamma_lc_lumcmf"], ref["ref.lc.gamma_lumcmf"])
        assert np.count_nonzero(b["flux"][0]) > 100
        # size-independent properties: the direction-resolved sets average to the angle-averaged one
        np.testing.assert_allclose(b["flux"][1:].sum(axis=0) / mabins, b["flux"][0], rtol=1e-12, atol=0)
        np.testing.assert_allclose(b["lc_lum"][1:].sum(axis=0) / mabins, b["lc_lum"][0], rtol=1e-12, atol=0)
        # the emission columns of a bin add up to its flux (every escaped r-packet of these fixtures has an emission type)
        np.testing.assert_allclose(b["emission"][0].sum(axis=-1), b["flux"][0], rtol=1e-12, atol=0)
        # (2) angle-averaged only, no decomposition: same flux, smaller outputs
        eng.bin_escaped_packets(direction_bins=False, emission_absorption=0, nprocs_exspec=nprocs)
        b0 = spectra_mod.binned(eng)
        assert b0["flux"].shape[0] == 1 and "emission" not in b0
        close("flux (angle-averaged only)", b0["flux"][0], ref["ref.spec.flux"])
        # (3) the decomposition for every direction bin: set 0 as before, the sets of the bins average to it
        eng.bin_escaped_packets(direction_bins=True, emission_absorption=2, nprocs_exspec=nprocs)
        b2 = spectra_mod.binned(eng)
        assert b2["emission"].shape[0] == 1 + mabins
        close("emission (all sets)", b2["emission"][0], ref["ref.spec.emission"])
        np.testing.assert_allclose(b2["absorption"][1:].sum(axis=0) / mabins, b2["absorption"][0], rtol=1e-12, atol=0)
        # (4) nprocs_exspec divides everything
        eng.bin_escaped_packets(direction_bins=False, emission_absorption=0, nprocs_exspec=4 * nprocs)
        np.testing.assert_allclose(spectra_mod.binned(eng)["flux"][0] * 4, b0["flux"][0], rtol=1e-15, atol=0)
        return b
    finally:
        eng.close()


def _float_ulps(a, b):
    """distance in float32 representable numbers"""
    ia = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    ib = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(ia - ib)


def _assert_grid_close(what, got, want, allowed_ulps, exact_algorithm, ncells):
    """float32 results of the grid update: within `allowed_ulps` float32 steps. With another libm (the device, or the host
    build with the +-1 ulp libm) the TOMS 748 iteration of a rare cell can accept its bracket one evaluation earlier or later
    than the reference's: both brackets hold the root to the reference's own accuracy of 1e-3, the accepted midpoints differ
    by less than that. Such cells (at most 1 in 1000) must agree to 1.5e-3; with the reference's libm none is allowed."""
    ulps = _float_ulps(got, want)
    outliers = ulps > allowed_ulps
    if exact_algorithm:
        assert not outliers.any(), f"{what}: {np.count_nonzero(outliers)} of {ulps.size} values differ by up to {ulps.max()} float32 steps"
        return
    cells = np.unique(np.nonzero(outliers)[0] // (got.size // ncells))  # the arrays are [cell] or [cell][ion]
    assert cells.size <= max(1, ncells // 1000), f"{what}: {cells.size} of {ncells} cells differ by more than {allowed_ulps} float32 steps"
    rel = np.abs(got.astype(np.float64) - want.astype(np.float64))[outliers] / np.maximum(np.abs(want.astype(np.float64))[outliers], 1e-300)
    assert rel.max(initial=0.) <= 1.5e-3, f"{what}: a value differs by {rel.max()} (the reference's root accuracy is 1e-3)"


def check_grid_update_lte(libpath, config, nts, device=0, max_ulps=0, max_ulps_balance=None):
    """partition functions, Saha ion balance, electron density (TOMS 748 to 1e-3) and the temperatures from J
    of every cell against the reference's own calculate_cellpartfuncts / calculate_ion_balance_nne / get_T_J_from_J
    (tests/golden/make_golden_grid.py). The results are float32 like the reference's grid arrays: `max_ulps` float32 steps
    (0 on the host build; the device's exp / pow differ from glibc's in the last bits of the double they are rounded from).
    `max_ulps_balance` (default: max_ulps) applies to the electron density and the ground-level populations that follow from
    it: the reference takes the root of the charge balance to 1e-3 only, and the interpolation steps of TOMS 748 divide by
    residuals that cancel to ~1e-9 of their terms near the root, so a last-bit difference in a Saha factor moves the accepted
    root by up to ~1e-7 relative (seen on 2 of 73 680 values of the bench-scale ladder on the B200, reproduced on the host
    with a +-1 ulp libm)."""
    if max_ulps_balance is None:
        max_ulps_balance = max_ulps
    fx = fixtures.load_golden(config, nts)
    ref = dict(np.load(os.path.join(fixtures.GOLDEN_DIR, f"{config}_grid_ts{nts}.npz")))
    eng = fixtures.ablib.ArtisB200(libpath=libpath, device=device)
    try:
        eng.set_arrays(fx["static"])
        eng.commit_static()
        before = dict(fx["before"])
        before["cell.elem_numberdens"] = ref["cell.elem_numberdens"]
        nc = before["cell.Te"].size
        # start from a state that the call has to overwrite: wrong electron densities and ground-level populations
        before["cell.nne"] = np.full(nc, 1.0, dtype=np.float32)
        eng.set_arrays(before)
        eng.update_grid_lte()
        got = {k: eng.get_array(k) for k in ("cell.nne", "cell.ion_partfuncts", "cell.ion_groundlevelpops", "cell.Te",
                                             "gridupdate.uppermost_ion", "gridupdate.status")}
        assert np.array_equal(got["cell.Te"], fx["before"]["cell.Te"])  # temperatures untouched in this mode
        assert np.array_equal(got["gridupdate.uppermost_ion"], ref["ref.grid.uppermost_ion"]), "uppermost ions differ"
        assert not np.any(got["gridupdate.status"] == 1)
        for key, refkey in (("cell.ion_partfuncts", "ref.grid.ion_partfuncts"), ("cell.nne", "ref.grid.nne"),
                            ("cell.ion_groundlevelpops", "ref.grid.ion_groundlevelpops")):
            _assert_grid_close(f"{config} ts{nts} {key}", got[key], ref[refkey], max_ulps if key == "cell.ion_partfuncts" else max_ulps_balance,
                               exact_algorithm=(max_ulps == 0), ncells=nc)
        # idempotence: the balance of a balanced state is the same state (the partition functions no longer change). Not with
        # NLTE populations: those are fixed numbers, so their share of a partition function moves with the ground population
        eng.update_grid_lte()
        for key in (() if "cell.nltepops" in before else ("cell.nne", "cell.ion_partfuncts", "cell.ion_groundlevelpops")):
            _assert_grid_close(f"{config} ts{nts} second pass {key}", eng.get_array(key), got[key], max(max_ulps_balance, 1),
                               exact_algorithm=False, ncells=nc)
        # charge conservation of the result: nne = sum over ions of charge x population
        nions = fx["static"]["ion.nlevels"].size
        g0 = fx["static"]["level.statweight"][fx["static"]["ion.uniquelevelindexstart"]].astype(np.float64)
        charge = np.concatenate([fx["static"]["elem.lowest_ionstage"][e] + np.arange(n) - 1
                                 for e, n in enumerate(fx["static"]["elem.nions"])]).astype(np.float64)
        pops = got["cell.ion_groundlevelpops"].astype(np.float64).reshape(nc, nions) * \
            got["cell.ion_partfuncts"].astype(np.float64).reshape(nc, nions) / g0
        np.testing.assert_allclose((pops * charge).sum(axis=1), got["cell.nne"], rtol=1e-6)

        # the same on the reference's temperature ladder 60 K .. 150 000 K: truncated ion lists, lowest-stage-only cells
        if "ref.grid.ladder_T" in ref:
            eng.set_arrays(before)
            # (the reference evaluated the ladder on the state its first pass left behind: with NLTE populations the
            # partition functions depend on the ground-level populations they start from)
            eng.set_array("cell.nne", ref["ref.grid.nne"])
            eng.set_array("cell.ion_partfuncts", ref["ref.grid.ion_partfuncts"])
            eng.set_array("cell.ion_groundlevelpops", ref["ref.grid.ion_groundlevelpops"])
            eng.set_array("cell.Te", ref["ref.grid.ladder_T"])
            eng.set_array("cell.TJ", ref["ref.grid.ladder_T"])
            eng.update_grid_lte()
            upper = eng.get_array("gridupdate.uppermost_ion")
            assert np.array_equal(upper, ref["ref.grid.ladder_uppermost_ion"]), "ladder: uppermost ions differ"
            assert upper.min() == 0 and len(np.unique(upper)) >= 2  # the truncation (and the neutral branch) is exercised
            for key, refkey in (("cell.ion_partfuncts", "ref.grid.ladder_ion_partfuncts"), ("cell.nne", "ref.grid.ladder_nne"),
                                ("cell.ion_groundlevelpops", "ref.grid.ladder_ion_groundlevelpops")):
                _assert_grid_close(f"{config} ts{nts} ladder {key}", eng.get_array(key), ref[refkey],
                                   max_ulps if key == "cell.ion_partfuncts" else max_ulps_balance, exact_algorithm=(max_ulps == 0),
                                   ncells=nc)

        # temperatures from the J estimator (get_T_J_from_J): needs the estimator buffer of a timestep (toy grids: the host
        # build takes most of a minute for the tables of the bench-scale model)
        if nc > 1000 and "hostsim" in os.path.basename(libpath):
            return eng.last_gridupdate_ms()
        eng.set_arrays(fx["before"])
        eng.set_array("cell.elem_numberdens", ref["cell.elem_numberdens"])
        eng.begin_timestep(nts)
        n = int(fx["before"]["packets.count"][0])
        stride = int(fx["before"]["packets.stride"][0])
        aos = fx["before"]["packets.aos"].copy()
        eng.set_option("rng_mode", 1)
        eng.update_packets_host(nts, aos, n, stride)
        J = eng.get_array("est.J")
        with np.errstate(divide="ignore", invalid="ignore"):
            factor = np.where(J > 0, ref["ref.grid.J_test"] / J, np.inf)  # so that J x factor is the reference's test value
        usable = np.isfinite(factor) | np.isinf(ref["ref.grid.J_test"])
        eng.set_array("cell.estimator_normfactor_over4pi", np.where(np.isfinite(factor), factor, np.inf))
        eng.update_grid_lte(temperatures_from_J=True, mintemp=float(ref["ref.grid.mintemp"][0]), maxtemp=float(ref["ref.grid.maxtemp"][0]))
        T_J = eng.get_array("cell.TJ")
        finite_target = np.isfinite(ref["ref.grid.J_test"]) & usable & (J > 0)
        # J x (J_test / J) reproduces J_test to an ulp: one float32 step more than the kernel's own tolerance
        ulps = _float_ulps(T_J[finite_target], ref["ref.grid.T_J_from_J"][finite_target])
        assert finite_target.sum() > 0 and ulps.max() <= max_ulps + 1, f"T_J: up to {ulps.max()} float32 steps"
        if finite_target.sum() > nc // 2:  # (the bench-scale fixture has 2000 packets for 3684 cells: most cells see no packet)
            clamped = ref["ref.grid.T_J_from_J"][finite_target]
            assert clamped.min() == ref["ref.grid.mintemp"][0] and clamped.max() == ref["ref.grid.maxtemp"][0]  # both clamps exercised
        keep = np.isinf(ref["ref.grid.J_test"]) & (J > 0)
        assert np.array_equal(T_J[keep], fx["before"]["cell.TJ"][keep])  # non-finite estimator: the old value is kept
        assert np.array_equal(eng.get_array("cell.Te"), T_J) and np.array_equal(eng.get_array("cell.TR"), T_J)
        assert np.all(eng.get_array("cell.W") == 1.)
        return eng.last_gridupdate_ms()
    finally:
        eng.close()


def check_device_cooling_contribs(libpath, config, nts, device=0, rel=REL_TOL, tol=1e-9, options=None):
    """kpkt::calculate_cooling_rates (kpkt.cc:281-303) evaluated by the per-cell table build (option device_cooling_contribs)
    instead of being handed over by the host: cell.ion_cooling_contribs against the reference's own values, and the packet
    histories of the timestep (k-packets pick their ion from it) unchanged. The host array is zeroed to prove who wrote it."""
    fx = fixtures.load_golden(config, nts)
    want = fx["before"]["cell.ion_cooling_contribs"]
    fx2 = dict(fx)
    fx2["before"] = dict(fx["before"])
    fx2["before"]["cell.ion_cooling_contribs"] = np.zeros_like(want)
    opts = {"device_cooling_contribs": 1}
    opts.update(options or {})
    eng = fixtures.make_engine(libpath, fx2, rng="xoshiro", device=device, options=opts)
    try:
        def check_totals():
            got = eng.get_array("cell.ion_cooling_contribs")
            err = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
            all_thick = bool(np.all(fx["before"]["cell.thick"] == 1))
            if windowed:
                # cell-batched tables: only the windows that packets wait for are built, the other cells are never read
                written = got != 0.
                assert (written.any() or all_thick) and err[written].max(initial=0.) <= rel, \
                    f"{config} ts{nts}: ion cooling contributions differ by {err[written].max(initial=0.)}"
                return
            assert (want.max() > 0 or all_thick) and err.max() <= rel, f"{config} ts{nts}: ion cooling contributions differ by {err.max()}"

        windowed = "table_window_cells" in opts  # then a cell's totals are written by the pass that builds its tables
        if not windowed:
            check_totals()
        n = int(fx["before"]["packets.count"][0])
        stride = int(fx["before"]["packets.stride"][0])
        aos = fx["before"]["packets.aos"].copy()
        eng.update_packets_host(nts, aos, n, stride)
        est = eng.estimators()
        check_totals()
        from artis_b200 import snapshot as snap
        frac_ok, worst, est_err = compare_run.compare(aos.view(snap.packet_dtype(stride)), est, fx["after"], tol=tol, verbose=False)
        assert frac_ok == 1.0, f"{config} ts{nts}: {100 * (1 - frac_ok):.3f} % of the packets differ with device-side cooling totals"
        assert int(est["counters"][fixtures.INTERACTIONS]) == int(fx["after"]["counters"][fixtures.INTERACTIONS])
    finally:
        eng.close()


def check_device_expansion_opacities(libpath, config, nts, device=0, tol=1e-9, options=None, max_ulps=0, rel=0.):
    """calculate_expansion_opacities (rpkt.cc:1071-1123) evaluated by the per-cell table build (option
    device_expansion_opacities) instead of being handed over by the host: cell.expansionopacities (float32) and
    cell.expopac_planck_cumulative against the reference's own arrays, and the packet histories of the timestep unchanged. The
    host arrays are zeroed to prove who wrote them."""
    from artis_b200 import snapshot as snap
    fx = fixtures.load_golden(config, nts)
    fx2 = dict(fx)
    fx2["before"] = dict(fx["before"])
    names = [k for k in ("cell.expansionopacities", "cell.expopac_planck_cumulative") if k in fx["before"]]
    assert names, f"{config}: the fixture has no expansion opacities"
    for name in names:
        fx2["before"][name] = np.zeros_like(fx["before"][name])
    opts = {"device_expansion_opacities": 1}
    opts.update(options or {})
    eng = fixtures.make_engine(libpath, fx2, rng="xoshiro", device=device, options=opts)
    try:
        n = int(fx["before"]["packets.count"][0])
        stride = int(fx["before"]["packets.stride"][0])
        aos = fx["before"]["packets.aos"].copy()
        eng.update_packets_host(nts, aos, n, stride)
        est = eng.estimators()
        if "cell.expansionopacities" in names:
            got = eng.get_array("cell.expansionopacities", dtype=np.float32)
            want = fx["before"]["cell.expansionopacities"]
            ulps = _float_ulps(got, want)
            assert np.count_nonzero(want) > 100 and ulps.max() <= max_ulps, \
                f"{config} ts{nts}: {np.count_nonzero(ulps > max_ulps)} of {ulps.size} bin opacities differ by up to {ulps.max()} float32 steps"
        if "cell.expopac_planck_cumulative" in names:
            got = eng.get_array("cell.expopac_planck_cumulative")
            want = fx["before"]["cell.expopac_planck_cumulative"]
            err = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
            assert want.max() > 0 and err.max() <= rel, f"{config} ts{nts}: Planck-weighted cumulative opacity differs by {err.max()}"
        frac_ok, worst, est_err = compare_run.compare(aos.view(snap.packet_dtype(stride)), est, fx["after"], tol=tol, verbose=False)
        assert frac_ok == 1.0, f"{config} ts{nts}: {100 * (1 - frac_ok):.3f} % of the packets differ with device-side expansion opacities"
        assert int(est["counters"][fixtures.INTERACTIONS]) == int(fx["after"]["counters"][fixtures.INTERACTIONS])
    finally:
        eng.close()
