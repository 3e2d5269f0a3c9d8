"""GPU parity of the spectra / light-curve binning (artisb200_bin_escaped_packets, csrc/spectra.h).

  * the fixtures' packets against the reference's own add_to_spec_res / add_to_lc_res / get_escapedirectionbin
    (tests/golden/*_spectra_ts*.npz): bins exact, sums within 1e-12
  * two million synthetic packets against the numpy restatement oracle/spectra_oracle.py (pinned on the same reference
    arrays by tests/test_spectra.py) + size-independent properties (the direction-resolved sets average to the angle-averaged
    set; the emission columns of a bin add up to its flux; energy bookkeeping of the light curve)
Collected last (file name): a failure here must not hide the hot-path parity tests from a `pytest -x` run."""
import os
import sys

import numpy as np
import pytest

from artis_b200 import lib as ablib
from artis_b200 import snapshot as snap
from artis_b200 import spectra as spectra_mod
from tests import fixtures, parity_checks

sys.path.insert(0, os.path.join(fixtures.ROOT, "oracle"))
import spectra_oracle  # noqa: E402
from bench_spectra import synthetic_packets  # noqa: E402  (tools/)

pytestmark = pytest.mark.gpu
SPECTRA_CASES = [("classic3d_toy", 2), ("kilonova_toy", 4), ("classic_toy_1d", 3), ("classic_detailedbf_toy", 3)]


@pytest.mark.parametrize("config,nts", SPECTRA_CASES)
def test_binning_matches_the_reference(config, nts):
    parity_checks.check_spectra(ablib.library_path(fixtures.PRESET_OF[config]), config, nts)


def test_two_million_synthetic_packets_against_the_numpy_oracle():
    config = "classic3d_toy"
    static = fixtures.load_golden(config, 2)["static"]
    n = 2_000_000
    pk = synthetic_packets(static, n)
    eng = ablib.ArtisB200(libpath=ablib.library_path(fixtures.PRESET_OF[config]))
    try:
        eng.set_arrays(static)
        eng.commit_static()
        eng.upload_packets(pk.view(np.uint8), n, pk.dtype.itemsize)
        eng.set_option("spec_record_dirbin", 1)
        eng.bin_escaped_packets(direction_bins=True, emission_absorption=1, nprocs_exspec=3)
        got = spectra_mod.binned(eng)
        ms = eng.last_binning_ms()
        eng.set_option("spec_record_dirbin", 0)
        eng.bin_escaped_packets(direction_bins=True, emission_absorption=1, nprocs_exspec=3)
        ms_plain = eng.last_binning_ms()
    finally:
        eng.close()
    want = spectra_oracle.bin_packets(pk, static, 1e14, 5e15, nnubins=1000, nprocs_exspec=3)
    mismatched = int(np.count_nonzero(got["dirbin"] != want["dirbin"]))
    print(f"[spectra] {n} packets ({int((pk['type'] == 32).sum())} escaped) binned in {ms_plain:.3f} ms on the device "
          f"({ms:.3f} ms with the direction-bin record); direction bins differing from the numpy oracle: {mismatched}")
    # acos / log of the device against glibc's: a packet within an ulp of a bin edge may fall on the other side
    assert mismatched <= 2
    for key in ("flux", "emission", "trueemission", "absorption", "lc_lum", "lc_lumcmf", "gamma_lc_lum", "gamma_lc_lumcmf"):
        a, b = got[key].ravel(), want[key].ravel()
        differing = np.count_nonzero((a != 0.) != (b != 0.))
        assert differing <= 4, f"{key}: {differing} bins filled on one side only"
        scale = np.maximum(np.abs(b), np.abs(b).max() * 1e-6)
        bad = np.count_nonzero(np.abs(a - b) / scale > 1e-10)
        assert bad <= 4, f"{key}: {bad} bins differ by more than 1e-10"
    # properties that do not need an oracle
    np.testing.assert_allclose(got["flux"][1:].sum(axis=0) / 100., got["flux"][0], rtol=1e-11, atol=0)
    np.testing.assert_allclose(got["lc_lumcmf"][1:].sum(axis=0) / 100., got["lc_lumcmf"][0], rtol=1e-11, atol=0)
    width = static["timesteps.width"][:-1]
    esc_r = (pk["type"] == 32) & (pk["escape_type"] == 11)
    t_arrive = pk["escape_time"].astype(np.float64) - (pk["pos"] * pk["dir"]).sum(axis=1) / 2.99792458e10
    inside = esc_r & (t_arrive > static["scalar.tmin"][0]) & (t_arrive < static["timesteps.start"][-1])
    # energy bookkeeping: sum over timesteps of L * dt * nprocs = energy of the escaped r-packets that arrive in [tmin, tmax)
    np.testing.assert_allclose((got["lc_lum"][0] * width).sum() * 3, pk["e_rf"][inside].sum(), rtol=1e-10)
