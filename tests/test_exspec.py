"""exspec equivalence: the reference's own post-processor run on the packets file of a complete reference
run, against artis_b200/exspec.py (the library's text reader + the device binning code, host build here) on the same file.
Development container only: needs oracle/_ref/<config>/parity/{sn3d_ref, exspec}."""
import os
import subprocess

import numpy as np
import pytest

from artis_b200 import exspec as exspec_mod
from artis_b200 import snapshot as snap
from artis_b200 import spectra as spectra_mod
from tests import fixtures


def _numbers(path):
    return [[float(tok) for tok in line.split()] for line in open(path)]


@pytest.mark.parametrize("config", ["classic3d_toy"])  # 3-D, POL_ON (kilonova_toy passes as well: 2-D, escaped gamma packets)
def test_exspec_files_of_a_reference_run(config, tmp_path):
    import run_oracle
    odir = run_oracle.oracle_dir(config, "parity")
    if not (os.path.exists(os.path.join(odir, "sn3d_ref")) and os.path.exists(os.path.join(odir, "exspec"))):
        pytest.skip("oracle/_ref not built (development container only)")
    rundir = run_oracle.run(config, "parity", "ref_perpacket", "all", rundir=str(tmp_path / "run"))
    # the reference's exspec in the run folder (it re-reads the inputs and packets00_0000.out)
    for stale in ("spec.out", "light_curve.out", "emission.out", "absorption.out"):
        if os.path.exists(os.path.join(rundir, stale)):
            os.remove(os.path.join(rundir, stale))
    with open(os.path.join(rundir, "exspec_stdout.txt"), "w") as out:
        subprocess.run([os.path.join(odir, "exspec")], cwd=rundir, stdout=out, stderr=subprocess.STDOUT, check=True)
    static = snap.read_snapshot(os.path.join(rundir, "dump", "static.abt"))
    eng = fixtures.ablib.ArtisB200(libpath=fixtures.hostsim_library(fixtures.PRESET_OF[config]))
    eng.set_arrays(static)
    eng.commit_static()
    # the reader against the packets the run held in memory: the text keeps six digits
    raw, n = eng.read_text_packets(os.path.join(rundir, "packets00_0000.out"))
    pk = raw.view(snap.packet_dtype(240))
    assert n == len(pk) and n > 500
    last = int(open(os.path.join(rundir, "input.txt")).read().split("\n")[2].split()[1]) - 1
    mem = snap.packets_view(snap.read_snapshot(os.path.join(rundir, "dump", f"ts{last}_after.abt")))
    assert np.array_equal(pk["number"], mem["number"]) and np.array_equal(pk["type"], mem["type"])
    assert np.array_equal(pk["emissiontype"], mem["emissiontype"]) and np.array_equal(pk["escape_type"], mem["escape_type"])
    np.testing.assert_allclose(pk["nu_rf"], mem["nu_rf"], rtol=1e-5)
    # "nan" stops the reference's extraction: 0 is stored and the rest of the row keeps a default Packet's values
    no_trueem = np.isnan(mem["trueem_pos"][:, 0]) & ~np.isnan(mem["em_pos"][:, 0])
    if no_trueem.any():
        assert np.all(pk["trueem_pos"][no_trueem, 0] == 0.) and np.all(np.isnan(pk["trueem_pos"][no_trueem, 1]))
        assert np.all(pk["trueem_time"][no_trueem] == -1.) and np.all(pk["pellet_nucindex"][no_trueem] == -1)
        assert np.array_equal(pk["nscatterings"][no_trueem], mem["nscatterings"][no_trueem])  # columns before the nan are read
    no_em = np.isnan(mem["em_pos"][:, 0])  # packets that never emitted (pellets, packets still in flight as k-packets ...)
    assert no_em.any() and np.all(pk["em_pos"][no_em, 0] == 0.) and np.all(np.isnan(pk["em_pos"][no_em, 1]))
    assert np.all(pk["absorptiontype"][no_em] == 0) and np.all(pk["em_time"][no_em] == -1.) and np.all(pk["pellet_decaytype"][no_em] == -1)
    assert np.array_equal(pk["trueemissiontype"][no_em], mem["trueemissiontype"][no_em])  # columns before the nan are read

    out = tmp_path / "mine"
    exspec_mod.exspec(eng, static, rundir, outdir=str(out), nprocs_exspec=1, only_dirbins={0, 7, 18, 42, 63, 99}, pol_on=True)
    eng.close()
    files = ["spec.out", "light_curve.out", "gamma_light_curve.out", "emission.out", "emissiontrue.out", "absorption.out",
             "specpol.out", "emissionpol.out", "absorptionpol.out", "gamma_spec.out"]
    res = spectra_mod.OUTDIR_RESFILES
    files += [os.path.join(res, f) for f in ("spec_res_00.out", "spec_res_42.out", "spec_res_99.out", "light_curve_res_07.out",
                                             "emission_res_63.out", "absorption_res_18.out", "specpol_res_42.out", "emissionpol_res_07.out",
                                             "absorptionpol_res_99.out")]
    compared = 0
    for name in files:
        theirs = os.path.join(rundir, name)
        assert os.path.exists(theirs), f"the reference's exspec did not write {name}"
        a, b = _numbers(theirs), _numbers(os.path.join(out, name))
        assert len(a) == len(b), name
        for row_a, row_b in zip(a, b):
            assert len(row_a) == len(row_b), name
            # "{:g}" keeps six digits: a sum that differs in its last bits may round the sixth digit the other way
            np.testing.assert_allclose(row_b, row_a, rtol=2e-5, atol=0, err_msg=name)
        compared += sum(1 for row in a for v in row if v != 0.)
        # the files are expected to be identical text, not just numerically close
        assert open(theirs).read() == open(os.path.join(out, name)).read(), f"{name}: text differs"
    assert compared > 1000
