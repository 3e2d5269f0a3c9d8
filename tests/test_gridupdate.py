"""LTE part of the per-cell grid update on the device (artisb200_update_grid_lte, csrc/gridupdate.h), CPU
side: the device code compiled for the host against the reference's own calculate_cellpartfuncts /
calculate_ion_balance_nne / get_T_J_from_J evaluated for every cell of the fixtures (tests/golden/*_grid_ts*.npz), on the
fixtures' own temperatures and on a ladder from 60 K to 150 000 K (truncated ion lists, cells with lowest ion stages only)."""
import os

import numpy as np
import pytest

from tests import fixtures, parity_checks

# (classic_nlte_toy: partition functions that read the NLTE solver's level and superlevel populations)
GRID_CASES = [("classic3d_toy", 2), ("kilonova_toy", 4), ("classic_toy_1d", 3), ("kilonova_2d_kat", 2), ("classic_nlte_toy", 4)]


@pytest.mark.parametrize("config,nts", GRID_CASES)
def test_lte_grid_update_matches_the_reference(config, nts):
    # float32 results identical to the reference's, cell by cell (same operations, same libm)
    parity_checks.check_grid_update_lte(fixtures.hostsim_library(fixtures.PRESET_OF[config]), config, nts, max_ulps=0)


@pytest.mark.parametrize("config,nts", [("classic3d_toy", 2), ("kilonova_toy", 4), ("kilonova_2d_kat", 2), ("classic_nlte_toy", 4)])
def test_lte_grid_update_with_another_libm(config, nts):
    # exp / pow moved by -1 / 0 / +1 ulp (tests/hostsim, ARTISB200_HOSTSIM_FUZZ_LIBM), as on the device: the partition functions
    # stay within one float32 step, the electron density root (TOMS 748 to 1e-3) and the populations within four: the
    # tolerances of the GPU test
    lib = fixtures.hostsim_library(fixtures.PRESET_OF[config], defines=("ARTISB200_HOSTSIM_FUZZ_LIBM",), tag="_fuzz")
    parity_checks.check_grid_update_lte(lib, config, nts, max_ulps=1, max_ulps_balance=4)


def test_lte_grid_update_reports_misuse():
    fx = fixtures.load_golden("classic3d_toy", 2)
    eng = fixtures.ablib.ArtisB200(libpath=fixtures.hostsim_library("classic"))
    eng.set_arrays(fx["static"])
    eng.commit_static()
    eng.set_arrays(fx["before"])
    with pytest.raises(fixtures.ablib.ArtisB200Error, match="cell.elem_numberdens"):
        eng.update_grid_lte()
    eng.set_array("cell.elem_numberdens", np.zeros(fx["before"]["cell.elem_massfracs"].size))
    with pytest.raises(fixtures.ablib.ArtisB200Error, match="estimator_normfactor_over4pi"):
        eng.update_grid_lte(temperatures_from_J=True, mintemp=3500., maxtemp=140000.)
    eng.close()
    nlte = fixtures.load_golden("classic_nlte_toy", 4)
    eng = fixtures.ablib.ArtisB200(libpath=fixtures.hostsim_library("classic_nlte"))
    eng.set_arrays(nlte["static"])
    eng.commit_static()
    eng.set_arrays({k: v for k, v in nlte["before"].items() if k != "cell.nltepops"})
    eng.set_array("cell.elem_numberdens", np.ones(nlte["before"]["cell.elem_massfracs"].size))
    with pytest.raises(fixtures.ablib.ArtisB200Error, match="cell.nltepops"):
        eng.update_grid_lte()
    eng.close()


COOLING_CASES = [("classic3d_toy", 0), ("classic3d_toy", 2), ("classic3d_grey_toy", 2), ("kilonova_toy", 4), ("classic_nt_toy", 3), ("nltephot_toy", 3), ("classic_detailedbf_toy", 3),
                 ("classic_multibin_toy", 4), ("kilonova_expansionopac_toy", 4)]


@pytest.mark.parametrize("windows", [False, True])
@pytest.mark.parametrize("config,nts", COOLING_CASES)
def test_ion_cooling_totals_from_the_table_build(config, nts, windows):
    # kpkt::calculate_cooling_rates (kpkt.cc:281-303) on the device: bit-identical to the reference's array on the host build,
    # every packet history of the timestep unchanged; also with cell-batched tables. Grey cells carry the reference's -1
    # (update_grid.cc:629-633): classic3d_toy timestep 0 mixes both kinds, classic3d_grey_toy is grey throughout
    options = {"schedule": 1, "wf_tail": 0, "table_window_cells": 11} if windows else None
    parity_checks.check_device_cooling_contribs(fixtures.hostsim_library(fixtures.PRESET_OF[config]), config, nts, rel=0., options=options)


EXPOPAC_CASES = [("kilonova_expansionopac_toy", 2), ("kilonova_expansionopac_toy", 4), ("kilonova_expopac_retrace_toy", 4),
                 ("kilonova_bbtherm_toy", 4)]


@pytest.mark.parametrize("variant", ["resident", "windows", "no-line-table"])
@pytest.mark.parametrize("config,nts", EXPOPAC_CASES)
def test_expansion_opacities_from_the_table_build(config, nts, variant):
    # calculate_expansion_opacities (rpkt.cc:1071-1123) on the device: the float32 bin opacities and the Planck-weighted
    # cumulative are bit-identical to the reference's arrays on the host build and every packet history is unchanged; with
    # resident and cell-batched tables, with and without the per-cell line table
    options = {"resident": None, "windows": {"schedule": 1, "wf_tail": 0, "table_window_cells": 11},
               "no-line-table": {"line_tau_table": 0}}[variant]
    parity_checks.check_device_expansion_opacities(fixtures.hostsim_library(fixtures.PRESET_OF[config]), config, nts, options=options)
