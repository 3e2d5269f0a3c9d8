"""CPU-side parity of the device physics headers, compiled for the host (tests/hostsim, TEST-ONLY build).

The development container has no GPU; these tests replay the oracle fixtures through a single-threaded host
compilation of exactly the headers the CUDA kernels are built from, so that every packet history, per-cell table
and deterministic helper is checked against the compiled reference before GPU time is spent. The same
assertions run against the real CUDA library on the B200 in tests/test_gpu_parity.py."""
import pytest

from tests import abi_checks, fixtures, parity_checks, stochastic_checks

CASES = [(c, t) for c, ts in fixtures.GOLDEN_TIMESTEPS.items() for t in ts]


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
    lib = fixtures.hostsim_library(fixtures.PRESET_OF[config])
    parity_checks.check_deterministic_kernels(lib, config, nts, ks_draws=500)


@pytest.mark.parametrize("schedule", sorted(SCHEDULES))
@pytest.mark.parametrize("config,nts", CASES)
def test_packet_histories_and_estimators(config, nts, schedule):
    lib = fixtures.hostsim_library(fixtures.PRESET_OF[config])
    parity_checks.check_packet_histories(lib, config, nts, options=SCHEDULES[schedule])


def test_bounded_launches_keep_histories():
    # cutting histories into bounded launches changes nothing: the per-packet continuum-opacity cache
    # (rpkt.cc:1023-1027) and pending macro-atom activations are part of the stored packet work state
    lib = fixtures.hostsim_library("kilonova_lte")
    parity_checks.check_packet_histories(lib, "kilonova_toy", 4, max_steps=7, options={"schedule": 0})


@pytest.mark.parametrize("config,nts", [("classic3d_toy", 2), ("kilonova_toy", 4)])
def test_stochastic_parity_ks_and_estimators(config, nts):
    # the same stated tests as on the GPU (tests/stochastic_checks.py)
    lib = fixtures.hostsim_library(fixtures.PRESET_OF[config])
    report = stochastic_checks.check_stochastic_parity(lib, config, nts, K=8)
    assert report["n_rpkt_ref"] > 0


def test_abi_reports_misuse():
    abi_checks.check_abi_errors(fixtures.hostsim_library("classic"))


@pytest.mark.parametrize("preset", ["classic", "kilonova_lte", "classic_multibin", "classic_nlte"])
def test_options_summary_names_the_preset(preset):
    abi_checks.check_options_summary(fixtures.hostsim_library(preset), preset)


# "Another libm": the device's exp/log/sin/cos/pow are correct to 1-2 ulp but not bit-equal to glibc's. This host build
# moves every transcendental result by -1/0/+1 ulp (tests/hostsim/hostsim.cc, ARTISB200_HOSTSIM_FUZZ_LIBM); the parity
# assertions, at the tolerances the GPU suite uses, must not depend on those bits. (It reproduces what the B200 showed:
# with nltephot_toy timestep 1 only the UPSCATTER/DOWNSCATTER split moves — parity_checks.py compares their sum.)
@pytest.mark.parametrize("config,nts", [("nltephot_toy", 1), ("kilonova_toy", 4), ("classic3d_toy", 2), ("classic_nt_toy", 3),
                                        ("classic_multibin_toy", 2)])
def test_parity_assertions_do_not_depend_on_libm_rounding(config, nts):
    lib = fixtures.hostsim_library(fixtures.PRESET_OF[config], defines=("ARTISB200_HOSTSIM_FUZZ_LIBM",), tag="_fuzz")
    parity_checks.check_packet_histories(lib, config, nts, tol=1e-9, est_tol=1e-9, options={"schedule": 0})
    parity_checks.check_packet_histories(lib, config, nts, tol=1e-9, est_tol=1e-9, options=SCHEDULES["wavefront-resort"])


def test_macroatom_searches_equal_upper_bound(tmp_path):
    # tests/hostsim/search_check.cc: the 8-way search against std::upper_bound (lengths 0..700)
    import os
    import subprocess
    csrc = os.path.join(fixtures.ROOT, "artis_b200", "csrc")
    exe = str(tmp_path / "search_check")
    subprocess.run(["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-Wno-unknown-pragmas", "-Wno-subobject-linkage", "-I" + csrc,
                    "-DARTISB200_PRESET_HEADER=\"options/preset_kilonova_lte.h\"",
                    os.path.join(fixtures.ROOT, "tests", "hostsim", "search_check.cc"), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert out.startswith("ok "), out


@pytest.mark.parametrize("schedule", ["history", "wavefront-notail", "wavefront-resort"])
@pytest.mark.parametrize("config,nts", [("kilonova_toy", 1), ("classic3d_toy", 0), ("classic_nt_toy", 3)])
def test_update_packets_is_idempotent_within_a_timestep(config, nts, schedule):
    lib = fixtures.hostsim_library(fixtures.PRESET_OF[config])
    assert parity_checks.check_idempotence(lib, config, nts, options=SCHEDULES[schedule]) > 0


def test_stream_download_option_is_accepted_by_a_backend_without_streams():
    # the host test build has no copy stream: the option falls back to the ordered download (same packets)
    lib = fixtures.hostsim_library("kilonova_lte")
    fx = fixtures.load_golden("kilonova_toy", 4)
    ordered, _, _, _ = fixtures.run_fixture(lib, fx, rng="philox", seed=11)
    streamed, _, _, _ = fixtures.run_fixture(lib, fx, rng="philox", seed=11, options={"stream_download": 1})
    assert ordered.tobytes() == streamed.tobytes()


# BASELINE configs[1] at its full atomic-data and grid size (54 892 lines, 1 475 continua, 3 684 cells), 2000 packets of
# the reference's parity build (tests/golden/kilonova_2d_kat_*): the same checks the B200 runs in test_gpu_bench_scale.py
def test_bench_scale_known_answer_vectors():
    lib = fixtures.hostsim_library("kilonova_lte")
    assert parity_checks.check_deterministic_kernels(lib, "kilonova_2d_kat", 2, ks_draws=100) > 1000


def test_bench_scale_histories_and_sampled_tables():
    lib = fixtures.hostsim_library("kilonova_lte")
    n, ncells = parity_checks.check_bench_scale_histories(lib, "kilonova_2d_kat", 2, options=SCHEDULES["wavefront-resort"])
    assert n == 2000 and ncells == 6


@pytest.mark.parametrize("schedule", ["history", "wavefront-notail", "wavefront-resort", "wavefront-tail"])
@pytest.mark.parametrize("config,nts,window", [("classic3d_toy", 2, 37), ("kilonova_toy", 4, 9), ("classic_toy_1d", 3, 3),
                                               ("nltephot_toy", 3, 1), ("classic_toy", 0, 5)])
def test_cell_batched_tables_keep_histories(config, nts, window, schedule):
    lib = fixtures.hostsim_library(fixtures.PRESET_OF[config])
    parity_checks.check_table_windows(lib, config, nts, window, options=SCHEDULES[schedule])


def test_failed_device_assertions_are_reported():
    abi_checks.check_device_error_record(fixtures.hostsim_library("kilonova_lte"))
