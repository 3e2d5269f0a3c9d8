"""GPU parity of the LTE grid update for a preset with NLTE level populations (partition functions that read the NLTE solver's
level and superlevel populations, ltepop.cc:177-197) against the reference's own functions
(tests/golden/classic_nlte_toy_grid_ts4.npz). Same tolerances as tests/test_gpu_zzz_gridupdate.py; the newest case, collected
last (file name)."""
import pytest

from artis_b200 import lib as ablib
from tests import fixtures, parity_checks

pytestmark = pytest.mark.gpu


def test_lte_grid_update_with_nlte_populations():
    parity_checks.check_grid_update_lte(ablib.library_path(fixtures.PRESET_OF["classic_nlte_toy"]), "classic_nlte_toy", 4,
                                        max_ulps=1, max_ulps_balance=4)
