"""Stochastic parity (BASELINE.json north_star: "stochastic outputs must agree with the reference CPU build on the same
synthetic model within Monte Carlo noise, checked by a stated chi-square/KS test at matched packet counts").

The oracle fixture holds ONE reference run (its own Xoshiro streams). The library is run K times with independent
Philox seeds on the same input packets. Stated tests:
  1. frequency distribution of the r-packets at the end of the timestep (in flight and escaped): two-sample
     Kolmogorov-Smirnov test of the pooled Philox runs against the reference run, p > 0.001; and each single run
     against the reference must pass p > 0.001 in at least K-1 of K runs;
  2. energy budget per packet type (escaped, r-packet, k-packet, gamma, pellet): reference within
     4 sigma sqrt(1 + 1/K) of the seed ensemble;
  3. per-cell estimators J, nuJ and the deposition estimator: z = (x_ref - mean_K) / (std_K sqrt(1 + 1/K)) over the
     cells that collect at least 1 % of the mean cell signal: |z| < 5 for >= 97 % of those cells, |mean z| < 0.75
     (a bias test; with K seeds z follows a t-distribution with K-1 degrees of freedom, hence the wide bounds)."""
import numpy as np
from scipy import stats as spstats

from tests import fixtures

TYPE_ESCAPE, TYPE_RPKT = 32, 11


def check_stochastic_parity(libpath, config, nts, K=8, options=None):
    fx = fixtures.load_golden(config, nts)
    after = fx["after"]
    ref = fixtures.snap.packets_view(after)
    runs = []
    for k in range(K):
        pk, est, _, _ = fixtures.run_fixture(libpath, fx, rng="philox", seed=4242 + k, options=options)
        runs.append((pk.copy(), {name: est[name].copy() for name in ("est.J", "est.nuJ", "est.dep_gamma", "counters")}))

    # 1. KS test on the r-packet frequencies
    def rpkt_nu(pk):
        sel = (pk["type"] == TYPE_RPKT) | (pk["type"] == TYPE_ESCAPE)
        return np.log(pk["nu_rf"][sel & (pk["nu_rf"] > 0)])

    nu_ref = rpkt_nu(ref)
    report = {"n_rpkt_ref": int(nu_ref.size)}
    if nu_ref.size >= 100:
        pooled = np.concatenate([rpkt_nu(pk) for pk, _ in runs])
        p_pooled = spstats.ks_2samp(nu_ref, pooled).pvalue
        p_single = [spstats.ks_2samp(nu_ref, rpkt_nu(pk)).pvalue for pk, _ in runs]
        report.update(ks_p_pooled=float(p_pooled), ks_p_single_min=float(min(p_single)))
        assert p_pooled > 1e-3, f"KS test of r-packet frequencies against the reference: p = {p_pooled:.2e}"
        assert sum(p > 1e-3 for p in p_single) >= K - 1, p_single

    # 2. energy per packet type
    widen = np.sqrt(1 + 1 / K)
    for t in (32, 11, 12, 10, 100):
        xs = np.array([pk["e_cmf"][pk["type"] == t].sum() for pk, _ in runs])
        x_ref = ref["e_cmf"][ref["type"] == t].sum()
        tol = 4.0 * xs.std(ddof=1) * widen + 1e-9 * abs(xs.mean())
        assert abs(x_ref - xs.mean()) <= tol, (f"energy in packets of type {t}", x_ref, xs.mean(), xs.std(ddof=1))

    # 3. per-cell estimators
    for name in ("est.J", "est.nuJ", "est.dep_gamma"):
        x = np.array([est[name] for _, est in runs])
        mean, std = x.mean(axis=0), x.std(axis=0, ddof=1)
        x_ref = after[name][: mean.size]  # the reference allocates some estimators with one spare element
        sel = (mean > 0.01 * mean.mean()) & (std > 0)
        if np.count_nonzero(sel) < 5:
            continue
        z = (x_ref[sel] - mean[sel]) / (std[sel] * widen)
        report[name] = {"cells": int(np.count_nonzero(sel)), "frac_within_5": float(np.mean(np.abs(z) < 5)), "mean_z": float(z.mean())}
        assert np.mean(np.abs(z) < 5) >= 0.97, (name, float(np.mean(np.abs(z) < 5)))
        assert abs(z.mean()) < 0.75, (name, float(z.mean()))
    return report
