"""GPU parity of the LTE grid update for a preset with NLTE level populations (partition functions that read the NLTE solver's
level and superlevel populations, ltepop.cc:177-197) against the reference's own functions
(tests/golden/classic_nlte_toy_grid_ts4.npz), same tolerances as tests/test_gpu_zzz_gridupdate.py; and the optional outputs of
the binning kernel. The newest cases, collected last (file name)."""
import pytest

from artis_b200 import lib as ablib
from tests import fixtures, parity_checks

pytestmark = pytest.mark.gpu


def test_lte_grid_update_with_nlte_populations():
    parity_checks.check_grid_update_lte(ablib.library_path(fixtures.PRESET_OF["classic_nlte_toy"]), "classic_nlte_toy", 4,
                                        max_ulps=1, max_ulps_balance=4)


def test_stokes_spectra_and_gamma_ray_spectrum_properties():
    # the optional Stokes Q / U spectra and the gamma-ray spectrum of the binning kernel (exspec's specpol.out / gamma_spec.out)
    from tests import stokes_gamma_checks
    stokes_gamma_checks.check_stokes_and_gamma_spectrum(ablib.library_path("classic"), n=400_000)
