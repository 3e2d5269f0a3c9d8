"""Property checks of the optional Stokes Q / U spectra and the gamma-ray spectrum of artisb200_bin_escaped_packets
(options spec_stokes, spec_gamma_spectrum), shared by the CPU (host build) and GPU tests. No oracle needed: with the same
Stokes parameters on every packet the Q / U arrays are multiples of the I arrays, and the gamma-ray spectrum integrates to the
energy of the escaped gamma packets that arrive inside the time and frequency ranges. (The exact arrays are held to the
reference's own exspec output in tests/test_exspec.py.)"""
import numpy as np

from artis_b200 import spectra as spectra_mod
from tests import fixtures


def check_stokes_and_gamma_spectrum(libpath, n=40_000, device=0):
    from bench_spectra import synthetic_packets  # tools/
    static = fixtures.load_golden("classic3d_toy", 2)["static"]
    pk = synthetic_packets(static, n, seed=5)
    pk["stokes_q"] = 0.25
    pk["stokes_u"] = -0.5
    gamma = (pk["type"] == 32) & (pk["escape_type"] == 10)
    rng = np.random.default_rng(9)
    pk["nu_rf"][gamma] = np.exp(rng.uniform(np.log(0.5e19), np.log(2e21), size=int(gamma.sum())))  # around 0.05 - 4 MeV
    eng = fixtures.ablib.ArtisB200(libpath=libpath, device=device)
    try:
        eng.set_arrays(static)
        eng.commit_static()
        eng.upload_packets(pk.view(np.uint8), n, pk.dtype.itemsize)
        eng.set_option("spec_stokes", 1)
        eng.set_option("spec_gamma_spectrum", 1)
        eng.bin_escaped_packets(direction_bins=True, emission_absorption=2, nprocs_exspec=2)
        b = spectra_mod.binned(eng)
        eng.set_option("spec_stokes", 0)
        eng.set_option("spec_gamma_spectrum", 0)
        eng.bin_escaped_packets(direction_bins=True, emission_absorption=2, nprocs_exspec=2)
        plain = spectra_mod.binned(eng)
    finally:
        eng.close()
    assert "flux_q" not in plain and "gamma_flux" not in plain
    for key in ("flux", "emission", "absorption"):
        np.testing.assert_allclose(plain[key], b[key], rtol=1e-12, atol=0)  # the I arrays do not depend on the options
        assert np.count_nonzero(b[key]) > 100
        np.testing.assert_allclose(b[f"{key}_q"], 0.25 * b[key], rtol=1e-12, atol=0)
        np.testing.assert_allclose(b[f"{key}_u"], -0.5 * b[key], rtol=1e-12, atol=0)
    # energy bookkeeping of the gamma-ray spectrum
    ts_start, width = static["timesteps.start"], static["timesteps.width"][:-1]
    t_arrive = pk["escape_time"].astype(np.float64) - (pk["pos"] * pk["dir"]).sum(axis=1) / 2.99792458e10
    nu_min, nu_max = 0.05 * 1.6021772e-6 / 6.6260755e-27, 4. * 1.6021772e-6 / 6.6260755e-27
    inside = gamma & (t_arrive > static["scalar.tmin"][0]) & (t_arrive < ts_start[-1]) & (pk["nu_rf"] > nu_min) & (pk["nu_rf"] < nu_max)
    assert inside.sum() > 100
    energy = (b["gamma_flux"] * b["gamma_delta_freq"].astype(np.float64)[:, None] * width[None, :]).sum() * 4.e12 * np.pi * 3.0857e18 ** 2 * 2
    np.testing.assert_allclose(energy, pk["e_rf"][inside].sum(), rtol=1e-10)
