"""GPU parity of the reference's NLTE photospheric preset, unmodified (BASELINE configs[4] in miniature: NLTE level
populations, multi-bin radiation field, detailed bound-free estimators without the photoionisation LUT, Spencer-Fano
non-thermal deposition with excitation), against the oracle fixture nltephot_toy. Same assertions as
tests/test_gpu_parity.py, in a file of its own (the newest preset; DESIGN.md section 8 has its GPU history)."""
import pytest

from artis_b200 import lib as ablib
from tests import fixtures, parity_checks
from tests.test_gpu_parity import SCHEDULES

pytestmark = pytest.mark.gpu
CONFIG = "nltephot_toy"
LIB = ablib.library_path(fixtures.PRESET_OF[CONFIG])


@pytest.mark.parametrize("nts", fixtures.GOLDEN_TIMESTEPS[CONFIG])
def test_deterministic_kernels(nts):
    parity_checks.check_deterministic_kernels(LIB, CONFIG, nts)


@pytest.mark.parametrize("schedule", sorted(SCHEDULES))
@pytest.mark.parametrize("nts", fixtures.GOLDEN_TIMESTEPS[CONFIG])
def test_packet_histories_and_estimators(nts, schedule):
    parity_checks.check_packet_histories(LIB, CONFIG, nts, tol=1e-9, est_tol=1e-9, options=SCHEDULES[schedule])
