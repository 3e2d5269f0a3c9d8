"""Parity assertions shared by the host-simulation tests (CPU, this container) and the GPU tests (B200, through
the C ABI). The checker is always the oracle fixture generated from the compiled reference (tests/golden)."""
import numpy as np

import compare_run
from tests import fixtures

REL_TOL = 1e-12  # BASELINE.json north_star: deterministic doubles within 1e-12 relative, indices bit-exact


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300)
    return np.abs(a - b) / denom


def check_deterministic_kernels(libpath, config, nts):
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
    eng.close()
    return n_chi


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


def check_packet_histories(libpath, config, nts, max_steps=0, min_exact_fraction=1.0, tol=1e-9, est_tol=1e-9, options=None):
    """replay the timestep with every packet continuing its own reference RNG stream: histories must coincide"""
    fx = fixtures.load_golden(config, nts)
    pk, est, built, _ = fixtures.run_fixture(libpath, fx, rng="xoshiro", max_steps=max_steps, options=options)
    after = fx["after"]
    ref = fixtures.snap.packets_view(after)
    # Free-bound emissions sample their frequency from select_continuum_nu, which the reference integrates
    # adaptively to a relative accuracy of 1e-3 (ratecoeff.cc:37, 583-610) and the device integrates with
    # fixed-order quadrature: those frequencies agree to the integration tolerance, not to rounding, and the rest of
    # that packet's history then differs. Every other history must coincide, so the number of differing packets is
    # bounded by the number of free-bound emission events the reference counted in this timestep.
    n_fb_events = int(after["counters"][17]) + int(after["counters"][10])  # K_STAT_TO_R_FB + MA_STAT_DEACTIVATION_FB
    frac_ok, worst, est_err = compare_run.compare(pk, est, after, tol=tol, verbose=False)
    n_bad = int(round((1.0 - frac_ok) * len(ref)))
    allowed = max(n_fb_events, int(np.ceil((1.0 - min_exact_fraction) * len(ref))))
    assert n_bad <= allowed, f"{n_bad} packets differ from the oracle, {n_fb_events} free-bound events ({worst})"
    fb = np.zeros(len(ref), dtype=bool)
    fb[:n_fb_events] = True
    n_fb = int(np.count_nonzero(fb))
    if min_exact_fraction == 1.0:
        check_cell_tables(built, after)
        if n_fb == 0:
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
            names = ["est.J", "est.nuJ", "est.ffheating", "est.colheating", "est.gamma", "est.bfheating", "est.dep_gamma",
                     "est.dep_positron", "est.dep_electron", "est.dep_alpha"]
            if "est.bins_J_raw" in after:  # MULTIBIN_RADFIELD_MODEL_ON fixtures
                assert "est.bins_J_raw" in est and after["est.bins_J_raw"].sum() > 0
                names += ["est.bins_J_raw", "est.bins_nuJ_raw"]
            if "est.bfrate_raw" in after:  # DETAILED_BF_ESTIMATORS_ON fixtures
                assert "est.bfrate_raw" in est and after["est.bfrate_raw"].sum() > 0
                names += ["est.bfrate_raw"]
            for name in names:
                assert est_err.get(name, 0.0) <= est_tol, f"{name}: {est_err[name]:.3e}"
            m = 9  # ts.scalars[9] (nt_energy_deposited) is file-static in the reference and not dumped
            scale = np.abs(after["ts.scalars"][:m]).max()
            if scale > 0:
                assert np.abs(est["ts.scalars"][:m] - after["ts.scalars"][:m]).max() / scale <= est_tol
            assert int(est["ts.pellet_decays"][0]) == int(after["ts.pellet_decays"][0])
    return frac_ok, n_fb, est


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
