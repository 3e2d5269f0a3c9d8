"""GPU parity of the LTE grid update (artisb200_update_grid_lte): every cell's partition functions, Saha
ion balance, electron density and temperature from J against the reference's own functions (tests/golden/*_grid_ts*.npz),
toy grids and the bench-scale model (3 684 cells, 20 ions, 2 400 levels). One float32 step for the partition functions (the
device's exp / pow are not glibc's), four for the electron density and what follows from it (the reference takes that root to
1e-3 only; tests/parity_checks.py check_grid_update_lte says why; tests/test_gridupdate.py reproduces both on the host). Collected after the hot-path parity tests (file name)."""
import pytest

from artis_b200 import lib as ablib
from tests import fixtures, parity_checks

pytestmark = pytest.mark.gpu
GRID_CASES = [("classic3d_toy", 2), ("kilonova_toy", 4), ("classic_toy_1d", 3), ("kilonova_2d_kat", 2)]


@pytest.mark.parametrize("config,nts", GRID_CASES)
def test_lte_grid_update_matches_the_reference(config, nts):
    ms = parity_checks.check_grid_update_lte(ablib.library_path(fixtures.PRESET_OF[config]), config, nts, max_ulps=1, max_ulps_balance=4)
    print(f"[gridupdate] {config}: LTE update of all cells in {ms:.3f} ms on the device")


@pytest.mark.parametrize("config,nts", [("classic3d_toy", 2), ("kilonova_toy", 4), ("classic_nt_toy", 3), ("nltephot_toy", 3)])
def test_ion_cooling_totals_from_the_table_build(config, nts):
    # kpkt::calculate_cooling_rates (kpkt.cc:281-303) by the per-cell table build: 1e-12 against the reference's array, packet
    # histories unchanged (wavefront stage kernels)
    parity_checks.check_device_cooling_contribs(ablib.library_path(fixtures.PRESET_OF[config]), config, nts,
                                                options={"schedule": 1, "wf_tail": 0})


@pytest.mark.parametrize("config,nts", [("kilonova_expansionopac_toy", 4), ("kilonova_expopac_retrace_toy", 4), ("kilonova_bbtherm_toy", 4)])
def test_expansion_opacities_from_the_table_build(config, nts):
    # calculate_expansion_opacities (rpkt.cc:1071-1123) by the per-cell table build: float32 bin opacities within one step
    # (device expm1), the Planck-weighted cumulative within 1e-12, packet histories unchanged
    parity_checks.check_device_expansion_opacities(ablib.library_path(fixtures.PRESET_OF[config]), config, nts, max_ulps=1, rel=1e-12,
                                                   options={"schedule": 1, "wf_tail": 0})
