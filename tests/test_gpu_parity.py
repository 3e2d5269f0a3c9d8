"""Parity of the CUDA path, through the C ABI, against the oracle (fixtures generated from the compiled reference).

  deterministic pieces  boundary_distance / closest_transition indices bit-exact, distances, continuum opacities
                        and every per-cell table within 1e-12 relative (BASELINE.json north_star)
  packet histories      every packet continues its own reference RNG stream (Xoshiro128++ state carried in the
                        256-byte GPU_ON Packet): final packet state, estimators and event counters must coincide
  properties            AoS->SoA->AoS round trip is the identity; Philox runs are bit-reproducible and independent
                        of launch segmentation for everything that has no per-launch cache; energy bookkeeping"""
import numpy as np
import pytest

from artis_b200 import lib as ablib
from tests import abi_checks, fixtures, parity_checks, stochastic_checks

pytestmark = pytest.mark.gpu
# the NLTE photospheric case has its own file (test_gpu_zz_nltephot.py), which pytest collects after this one
CASES = [(c, t) for c, ts in fixtures.GOLDEN_TIMESTEPS.items() for t in ts if c != "nltephot_toy"]


def _lib(config):
    return ablib.library_path(fixtures.PRESET_OF[config])


# how the packets are scheduled onto kernels must not change any packet's result (include/artis_b200.h, options)
SCHEDULES = {
    "wavefront": {"schedule": 1},                                       # default (toy sizes: mostly the tail kernel)
    "history": {"schedule": 0, "ma_record": 1},                                         # one whole-history kernel
    # chunked stage kernels (no lane refill), one step / one transition per visit
    "wavefront-notail": {"schedule": 1, "wf_tail": 0, "wf_sync_every": 3, "wf_rsteps_thick": 1, "wf_masteps": 1, "wf_ma_rounds": 1,
                         "wf_masteps_last": -1, "wf_refill_masteps": 0, "wf_refill_thicksteps": 0, "line_tau_table": 0, "ma_record": 0},
    "wavefront-walk": {"schedule": 1, "wf_tail": 0, "wf_masteps": 0, "wf_refill_masteps": 0},   # whole macro-atom walk per visit
    "wavefront-refill": {"schedule": 1, "wf_tail": 0, "wf_refill_masteps": 3, "wf_refill_thicksteps": 2},  # lane refill, short visits
    "wavefront-resort": {"schedule": 1, "wf_tail": 0, "wf_resort_every": 1, "wf_sync_every": 2, "ma_record": 1},  # lists re-sorted by cell
    "wavefront-rounds": {"schedule": 1, "wf_tail": 0, "wf_masteps": 1, "wf_ma_rounds": 3, "wf_masteps_last": 2, "wf_refill_masteps": 0},
    "wavefront-tail": {"schedule": 1, "wf_tail": 1000000, "wf_sync_every": 2, "wf_rsteps_thin": 3, "wf_masteps": 3},
    # per-cell tables of 11 cells at a time: packets wait for the pass that holds their cell (cell-batched tables)
    "wavefront-windows": {"schedule": 1, "wf_tail": 0, "table_window_cells": 11},
    # two wavefront instances over the two halves of the packets, side by side (CUDA backend; also with table windows)
    "wavefront-two": {"schedule": 1, "wf_tail": 300, "wf_instances": 2, "wf_sync_every": 2},
    "wavefront-two-windows": {"schedule": 1, "wf_tail": 0, "wf_instances": 2, "wf_grid_div": 2, "table_window_cells": 13},
}

@pytest.mark.parametrize("config,nts", CASES)
def test_deterministic_kernels(config, nts):
    parity_checks.check_deterministic_kernels(_lib(config), config, nts)


@pytest.mark.parametrize("schedule", sorted(SCHEDULES))
@pytest.mark.parametrize("config,nts", CASES)
def test_packet_histories_and_estimators(config, nts, schedule):
    # device libm differs from glibc in the last bits (<= 2 ulp): 1e-9 on packet state after up to ~1e3 interactions
    parity_checks.check_packet_histories(_lib(config), config, nts, tol=1e-9, est_tol=1e-9, options=SCHEDULES[schedule])


def test_bounded_launches_keep_histories():
    parity_checks.check_packet_histories(_lib("kilonova_toy"), "kilonova_toy", 4, max_steps=7, options={"schedule": 0})


def test_schedules_give_identical_packets():
    """Philox run: every schedule must return byte-identical packets (per-packet streams, per-packet caches)"""
    fx = fixtures.load_golden("kilonova_toy", 4)
    outs = {}
    for name, opts in SCHEDULES.items():
        pk, est, _, _ = fixtures.run_fixture(_lib("kilonova_toy"), fx, rng="philox", seed=5, options=opts)
        outs[name] = (pk.tobytes(), est["counters"].copy())
    first = outs["wavefront"]
    for name, (pkbytes, counters) in outs.items():
        assert pkbytes == first[0], name
        assert np.array_equal(counters, first[1]), name


@pytest.mark.parametrize("schedule", ["wavefront", "wavefront-notail", "wavefront-refill", "history"])
def test_streamed_download_returns_the_same_packets(schedule):
    """option stream_download: finished packets are copied back while the others are still being propagated, and the array
    comes back in completion order; sorted by packet number it must be byte for byte what the ordered download returns"""
    fx = fixtures.load_golden("kilonova_toy", 4)
    ordered, est_o, _, _ = fixtures.run_fixture(_lib("kilonova_toy"), fx, rng="philox", seed=11, options=SCHEDULES[schedule])
    opts = dict(SCHEDULES[schedule], stream_download=1)
    streamed, est_s, _, _ = fixtures.run_fixture(_lib("kilonova_toy"), fx, rng="philox", seed=11, options=opts)
    assert not np.array_equal(streamed["number"], ordered["number"]), "the streamed array is expected to be permuted"
    a = ordered[np.argsort(ordered["number"], kind="stable")]
    b = streamed[np.argsort(streamed["number"], kind="stable")]
    for name in a.dtype.names:  # every field of the reference's Packet, bit for bit (NaN == NaN)
        assert np.array_equal(np.ascontiguousarray(a[name]).view(np.uint8), np.ascontiguousarray(b[name]).view(np.uint8)), name
    assert np.array_equal(est_o["counters"], est_s["counters"])


@pytest.mark.parametrize("stride", [240, 256])
def test_packet_roundtrip_is_identity(stride):
    fx = fixtures.load_golden("classic_toy", 3)
    eng = fixtures.make_engine(_lib("classic_toy"), fx, rng="xoshiro" if stride == 256 else "philox")
    n = int(fx["before"]["packets.count"][0])
    src = fx["before"]["packets.aos"].reshape(n, 256)
    aos = np.ascontiguousarray(src[:, 16:] if stride == 240 else src).reshape(-1).copy()
    orig = aos.copy()
    eng.upload_packets(aos, n, stride)
    aos[:] = 0
    eng.download_packets(aos, n, stride)
    a = aos.view(fixtures.snap.packet_dtype(stride))
    b = orig.view(fixtures.snap.packet_dtype(stride))
    for name in a.dtype.names:
        if name == "rngstate" and stride == 256:
            continue
        x, y = a[name], b[name]
        assert np.array_equal(x, y) or np.array_equal(np.isnan(x), np.isnan(y)), name
    eng.close()


def _run_philox(config, nts, max_steps, seed=1234):
    fx = fixtures.load_golden(config, nts)
    return fixtures.run_fixture(_lib(config), fx, rng="philox", max_steps=max_steps, seed=seed)


def test_philox_is_reproducible_and_seed_dependent():
    pk1, est1, _, _ = _run_philox("classic3d_toy", 0, 0)
    pk2, est2, _, _ = _run_philox("classic3d_toy", 0, 0)
    assert pk1.tobytes() == pk2.tobytes()
    assert np.array_equal(est1["counters"], est2["counters"])
    pk3, est3, _, _ = _run_philox("classic3d_toy", 0, 0, seed=99)
    assert pk3.tobytes() != pk1.tobytes()


def test_philox_statistics_agree_with_reference_rng():
    """Same physics, different (counter-based) random numbers: the oracle's outcome must look like one more draw from
    the distribution of the Philox runs. K = 8 independent Philox seeds give mean and standard deviation of each
    aggregate; the oracle value x_ref has to satisfy |x_ref - mean| <= 4 * std * sqrt(1 + 1/K)  (two-sided 4 sigma for
    the difference of one sample and a K-sample mean)."""
    fx = fixtures.load_golden("classic3d_toy", 0)
    ref_pk = fixtures.snap.packets_view(fx["after"])
    after = fx["after"]

    def aggregates(pk, est):
        return {
            "escape_fraction": float(np.mean(pk["type"] == 32)),
            "rpkt_fraction": float(np.mean(pk["type"] == 11)),
            "pellet_fraction": float(np.mean(pk["type"] == 100)),
            "interactions": float(est["counters"][fixtures.INTERACTIONS]),
            "cellcrossings": float(est["counters"][29]),
            "J_sum": float(est["est.J"].sum()),
            "nuJ_sum": float(est["est.nuJ"].sum()),
            "dep_gamma_sum": float(est["est.dep_gamma"].sum()),
            "escaped_e_rf": float(pk["e_rf"][pk["type"] == 32].sum()),
        }

    ref = aggregates(ref_pk, {k: after[k] for k in ("counters", "est.J", "est.nuJ", "est.dep_gamma")})
    K = 8
    runs = []
    for seed in range(K):
        pk, est, _, _ = fixtures.run_fixture(_lib("classic3d_toy"), fx, rng="philox", seed=1000 + seed)
        runs.append(aggregates(pk, est))
    for name, x_ref in ref.items():
        xs = np.array([r[name] for r in runs])
        mean, std = xs.mean(), xs.std(ddof=1)
        assert abs(x_ref - mean) <= 4.0 * std * np.sqrt(1 + 1 / K) + 1e-12 * abs(mean), (name, x_ref, mean, std)
    # pellets that do not decay in this timestep are untouched by the random numbers: exact agreement
    assert np.isclose(runs[0]["pellet_fraction"], ref["pellet_fraction"], atol=4 * np.sqrt(0.25 * 2 / len(ref_pk)))


@pytest.mark.parametrize("config,nts", [("classic3d_toy", 2), ("kilonova_toy", 4), ("classic_toy", 3)])
def test_stochastic_parity_ks_and_estimators(config, nts):
    """Philox (production) random numbers against the reference's own run: KS test on the r-packet frequencies, energy
    budget per packet type, per-cell estimator z-scores (tests/stochastic_checks.py states the tests)."""
    stochastic_checks.check_stochastic_parity(_lib(config), config, nts, K=8)


def test_energy_bookkeeping():
    """rest-frame energy of a packet that did nothing but move is unchanged; total comoving energy emitted by pellet
    decays equals the timestep scalars (update_packets.cc:199-232)"""
    fx = fixtures.load_golden("kilonova_toy", 4)
    pk, est, _, _ = fixtures.run_fixture(_lib("kilonova_toy"), fx, rng="xoshiro")
    before = fixtures.snap.packets_view(fx["before"])
    still_pellet = pk["type"] == 100
    assert np.array_equal(pk["e_cmf"][still_pellet], before["e_cmf"][still_pellet])
    decayed = (before["type"] == 100) & ~still_pellet
    emitted = est["ts.scalars"][[2, 4, 6, 7, 8]].sum()  # positron, electron, alpha emission, spfission, gamma emission
    assert abs(emitted - before["e_cmf"][decayed].sum()) / max(emitted, 1e-300) < 1e-12
    assert int(est["ts.pellet_decays"][0]) == int(np.count_nonzero(decayed))


def test_abi_reports_misuse():
    abi_checks.check_abi_errors(_lib("classic_toy_1d"))


def test_errors_are_reported_not_swallowed():
    eng = ablib.ArtisB200(libpath=_lib("classic_toy"))
    with pytest.raises(ablib.ArtisB200Error):
        eng.commit_static()  # nothing set
    with pytest.raises(ablib.ArtisB200Error):
        eng.set_array("no.such.array", np.zeros(3))
    with pytest.raises(ablib.ArtisB200Error):
        eng.begin_timestep(0)
    eng.close()


@pytest.mark.parametrize("schedule", ["history", "wavefront-notail", "wavefront-resort", "wavefront-tail"])
@pytest.mark.parametrize("config,nts,window", [("classic3d_toy", 2, 37), ("kilonova_toy", 4, 9), ("classic_toy_1d", 3, 3),
                                               ("classic_toy", 0, 5)])
def test_cell_batched_tables_keep_histories(config, nts, window, schedule):
    parity_checks.check_table_windows(_lib(config), config, nts, window, options=SCHEDULES[schedule])


def test_failed_device_assertions_are_reported():
    abi_checks.check_device_error_record(_lib("kilonova_toy"))
