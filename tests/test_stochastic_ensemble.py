"""The stated stochastic parity test at matched packet counts (tools/stochastic_ensemble.py): a whole
twelve-timestep run of 1e5 packets, this library behind the reference's own driver against K = 8 seeds of the compiled
reference (tests/golden/stochastic_classic_spec.npz, generated in the development container), both post-processed by the
reference's exspec: spectrum (KS, chi^2), light curve, deposition, radiation-field estimators, interaction counts."""
import json
import os

import pytest

from tests import fixtures  # noqa: F401  (sets sys.path)
import stochastic_ensemble as se

CONFIG = "classic_spec"
ENSEMBLE = os.path.join(se.GOLDEN, f"stochastic_{CONFIG}.npz")


def test_thresholds_accept_the_reference_against_itself():
    """leave-one-out: one reference seed against the other seven must pass every stated threshold (a test that rejects two
    runs of the same code would be useless; one that cannot is toothless - see the biased case below)"""
    meta, refs = se.load_ensemble(ENSEMBLE)
    assert len(refs) == 8 and int(meta["nts"]) >= 10
    assert int(refs[0]["count"].sum()) > 50000  # most of the 1e5 packets escape within the run: a real spectrum
    report = se.compare(refs[0], refs[1:])
    failed = [name for name, t in report["tests"].items() if not t["passed"]]
    assert not failed, json.dumps({k: report["tests"][k] for k in failed}, indent=1)


def test_thresholds_reject_a_biased_run():
    """a 3 % shift of the escaping frequencies, or 5 % more energy, must fail"""
    _, refs = se.load_ensemble(ENSEMBLE)
    shifted = dict(refs[0], ks_nu=(refs[0]["ks_nu"] * 1.03).astype("float32"))
    assert not se.compare(shifted, refs[1:])["tests"]["spectrum_ks"]["passed"]
    brighter = dict(refs[0], lc=refs[0]["lc"] * 1.05, spec_sum=refs[0]["spec_sum"] * 1.05, spec_sum2=refs[0]["spec_sum2"] * 1.05**2)
    rep = se.compare(brighter, refs[1:])
    assert not (rep["tests"]["light_curve"]["passed"] and rep["tests"]["spectrum_chi2"]["passed"])


@pytest.mark.gpu
def test_whole_run_agrees_with_the_reference_ensemble():
    meta, refs = se.load_ensemble(ENSEMBLE)
    gpu = se.run_gpu(CONFIG, int(meta["nts"]), "classic")
    report = se.compare(gpu, refs)
    report["config"], report["reference"] = CONFIG, meta
    out = os.path.join(fixtures.ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "stochastic_report.json"), "w") as f:
            json.dump(report, f, indent=1)
    failed = [name for name, t in report["tests"].items() if not t["passed"]]
    assert not failed, json.dumps({k: report["tests"][k] for k in failed}, indent=1)
