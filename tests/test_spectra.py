"""Spectra and light-curve binning of the escaped packets, CPU side.

  * the device code (artis_b200/csrc/spectra.h), compiled for the host (tests/hostsim), against the arrays the reference's own
    add_to_spec_res / add_to_lc_res / get_escapedirectionbin produced for the fixtures' packets (tests/golden/*_spectra_ts*.npz)
  * the numpy restatement oracle/spectra_oracle.py against the same arrays: this pins the oracle that the GPU tests use at
    sizes the fixtures do not reach
  * the output files of artis_b200/spectra.py against the files the compiled reference itself wrote in a complete run
    (spec.out, light_curve.out, gamma_light_curve.out; development container only: needs oracle/_ref)"""
import os
import sys

import numpy as np
import pytest

from artis_b200 import snapshot as snap
from artis_b200 import spectra as spectra_mod
from tests import fixtures, parity_checks

sys.path.insert(0, os.path.join(fixtures.ROOT, "oracle"))
import spectra_oracle  # noqa: E402

SPECTRA_CASES = [("classic3d_toy", 2), ("kilonova_toy", 4), ("classic_toy_1d", 3), ("classic_detailedbf_toy", 3)]


@pytest.mark.parametrize("config,nts", SPECTRA_CASES)
def test_device_binning_code_on_the_host_matches_the_reference(config, nts):
    parity_checks.check_spectra(fixtures.hostsim_library(fixtures.PRESET_OF[config]), config, nts)


@pytest.mark.parametrize("config,nts", SPECTRA_CASES)
def test_numpy_oracle_is_pinned_on_the_reference(config, nts):
    fx = fixtures.load_golden(config, nts)
    ref = dict(np.load(os.path.join(fixtures.GOLDEN_DIR, f"{config}_spectra_ts{nts}.npz")))
    pk = snap.packets_view(fx["after"])
    got = spectra_oracle.bin_packets(pk, fx["static"], float(ref["ref.spec.nu_min"][0]), float(ref["ref.spec.nu_max"][0]),
                                     nnubins=int(ref["ref.spec.nnubins"][0]), nprocs_exspec=int(ref["ref.spec.nprocs_exspec"][0]))
    assert np.array_equal(got["dirbin"], ref["ref.spec.dirbin"])
    assert np.array_equal(got["lower_freq"], ref["ref.spec.lower_freq"]) and np.array_equal(got["delta_freq"], ref["ref.spec.delta_freq"])
    pairs = [("flux", 0, "ref.spec.flux"), ("emission", 0, "ref.spec.emission"), ("trueemission", 0, "ref.spec.trueemission"),
             ("absorption", 0, "ref.spec.absorption"), ("lc_lum", 0, "ref.lc.lum"), ("lc_lumcmf", 0, "ref.lc.lumcmf"),
             ("gamma_lc_lum", None, "ref.lc.gamma_lum"), ("gamma_lc_lumcmf", None, "ref.lc.gamma_lumcmf"),
             ("flux", slice(1, None), "ref.spec.flux_res"), ("lc_lum", slice(1, None), "ref.lc.lum_res"),
             ("lc_lumcmf", slice(1, None), "ref.lc.lumcmf_res")]
    for key, sel, refkey in pairs:
        mine = (got[key] if sel is None else got[key][sel]).ravel()
        # same addends in the same order as the reference's loop: equal to the last bit
        assert np.array_equal(mine, ref[refkey].ravel()), f"{config} ts{nts}: {key} differs from the reference"


def test_direction_bins_cover_the_sphere_like_the_reference():
    # poles, the phi = 0 seam and unnormalised directions (vectors.h:150-152)
    dirs = np.array([[0., 0., 1.], [0., 0., -1.], [1., 0., 0.], [-1., 0., 0.], [0., 1., 0.], [0., -1., 0.], [3., 0., 4.],
                     [1e-14, 0., 1.]])
    bins = spectra_oracle.escape_direction_bin(dirs)
    assert all(0 <= b < 100 for b in bins) and bins[0] // 10 == 9 and bins[1] // 10 == 0 and bins[2] // 10 == 5
    assert bins[6] == spectra_oracle.escape_direction_bin(np.array([[0.6, 0., 0.8]]))[0]
    rng = np.random.default_rng(5)
    v = rng.normal(size=(200000, 3))
    counts = np.bincount(spectra_oracle.escape_direction_bin(v), minlength=100)
    assert counts.min() > 1700 and counts.max() < 2300  # equal solid angles


def test_output_files_have_the_reference_format(tmp_path):
    mid = np.array([2., 3., 4.5]) * 86400.
    lower = np.array([1e14, 2e14], dtype=np.float32)
    delta = np.array([1e14, 2e14], dtype=np.float32)
    flux = np.array([[0., 1.5e-12, 3.25e-30], [1e-300, 123456789., 0.]])
    spectra_mod.write_spectrum_file(tmp_path / "spec.out", mid, lower, delta, flux, 2)
    assert (tmp_path / "spec.out").read_text() == "0 2 3 \n1.5e+14 0 1.5e-12 \n3e+14 1e-300 1.23457e+08 \n"
    spectra_mod.write_light_curve(tmp_path / "light_curve.out", mid, np.array([3.826e33, 0., 1.]), np.array([7.652e33, 1., 1.]), 2)
    assert (tmp_path / "light_curve.out").read_text() == "2 1 2\n3 0 2.6137e-34\n"
    spectra_mod.write_columns_file(tmp_path / "emission.out", np.arange(12, dtype=np.float64).reshape(2, 3, 2), 2)
    assert (tmp_path / "emission.out").read_text() == "0 1 \n2 3 \n6 7 \n8 9 \n"


def _numbers(path):
    return [[float(tok) for tok in line.split()] for line in open(path)]


@pytest.mark.parametrize("config", ["classic3d_toy", "kilonova_toy"])
def test_files_of_a_complete_reference_run(config, tmp_path):
    """the reference's own spec.out / light_curve.out / gamma_light_curve.out of a complete run against the files written
    from the same final packets by the device binning code (host build) + artis_b200/spectra.py"""
    import run_oracle
    binary = os.path.join(run_oracle.oracle_dir(config, "parity"), "sn3d_ref")
    if not os.path.exists(binary):
        pytest.skip("oracle/_ref not built (development container only)")
    static = fixtures.load_golden(config, fixtures.GOLDEN_TIMESTEPS[config][0])["static"]
    rundir = run_oracle.run(config, "parity", "ref_perpacket", "all", rundir=str(tmp_path / "run"))
    last = int(open(os.path.join(rundir, "input.txt")).read().split("\n")[2].split()[1]) - 1  # timestep_finish - 1
    after = snap.read_snapshot(os.path.join(rundir, "dump", f"ts{last}_after.abt"))
    eng = fixtures.ablib.ArtisB200(libpath=fixtures.hostsim_library(fixtures.PRESET_OF[config]))
    eng.set_arrays(static)
    eng.commit_static()
    eng.upload_packets(after["packets.aos"], int(after["packets.count"][0]), int(after["packets.stride"][0]))
    out = tmp_path / "mine"
    spectra_mod.write_partial_lightcurve_spectra(eng, last, str(out), static["timesteps.mid"], ntimesteps_finish=last + 1,
                                                 multidimensional=True)
    eng.close()
    files = ["spec.out", "light_curve.out", "gamma_light_curve.out"]
    resdir = os.path.join(rundir, spectra_mod.OUTDIR_RESFILES)
    if os.path.isdir(resdir):  # written by the reference for 2-D / 3-D models when the run is complete
        files += [os.path.join(spectra_mod.OUTDIR_RESFILES, f) for f in ("spec_res_00.out", "spec_res_57.out", "light_curve_res_99.out")]
    compared = 0
    for name in files:
        theirs = os.path.join(rundir, name)
        assert os.path.exists(theirs), f"the reference did not write {name}"
        a, b = _numbers(theirs), _numbers(os.path.join(out, name))
        assert len(a) == len(b), name
        for row_a, row_b in zip(a, b):
            # "{:g}" keeps six digits: a sum that differs in its last bits may round the sixth digit the other way
            np.testing.assert_allclose(row_b, row_a, rtol=2e-5, atol=0, err_msg=name)
        compared += sum(1 for row in a for v in row[1:] if v != 0.)
    assert compared > 100


def test_synthetic_packets_over_every_branch_against_the_numpy_oracle():
    """the device binning code (host build) on 60 000 random final packets - every kind of emission type, packets in flight,
    arrival times and frequencies inside and outside the binned ranges, escaped gamma packets - against the pinned numpy
    restatement; the same comparison runs with two million packets on the GPU (tests/test_gpu_zzz_spectra.py)"""
    from bench_spectra import synthetic_packets  # tools/
    static = fixtures.load_golden("classic3d_toy", 2)["static"]
    n = 60_000
    pk = synthetic_packets(static, n, seed=3)
    eng = fixtures.ablib.ArtisB200(libpath=fixtures.hostsim_library("classic"))
    eng.set_arrays(static)
    eng.commit_static()
    eng.upload_packets(pk.view(np.uint8), n, pk.dtype.itemsize)
    eng.set_option("spec_record_dirbin", 1)
    eng.bin_escaped_packets(direction_bins=True, emission_absorption=1, nprocs_exspec=2)
    got = spectra_mod.binned(eng)
    eng.close()
    want = spectra_oracle.bin_packets(pk, static, 1e14, 5e15, nnubins=1000, nprocs_exspec=2)
    assert np.array_equal(got["dirbin"], want["dirbin"])
    for key in ("flux", "emission", "trueemission", "absorption", "lc_lum", "lc_lumcmf", "gamma_lc_lum", "gamma_lc_lumcmf"):
        # one thread, packet order: the same additions in the same order as the oracle's
        assert np.array_equal(got[key], want[key]), key
    assert np.count_nonzero(got["emission"][0][:, :, -1]) > 0  # free-free column
    nions = static["elem.anumber"].size * int(static["elem.nions"].max())
    assert np.count_nonzero(got["emission"][0][:, :, nions:2 * nions]) > 0  # bound-free columns


def test_stokes_spectra_and_gamma_ray_spectrum_properties():
    from tests import stokes_gamma_checks
    stokes_gamma_checks.check_stokes_and_gamma_spectrum(fixtures.hostsim_library("classic"), n=12_000)
