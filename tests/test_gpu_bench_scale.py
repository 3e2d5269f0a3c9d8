"""Deterministic parity at the bench model's full size (B200, through the C ABI).

tests/golden/kilonova_2d_kat_* is BASELINE configs[1] with its full atomic data (54 892 lines, 1 750 levels, 1 475
bound-free continua) and grid (50 x 100 cells, 3 684 non-empty), evolved by the reference's parity build with 2000 packets:
known-answer vectors of boundary_distance, closest_transition (binary search over the full line list), the continuum
opacity (windows of hundreds of continua, multi-word keep-bitmaps) and select_continuum_nu evaluated by the reference
itself, its cell-cache tables of six cells, and all 2000 packet histories of timestep 2."""
import numpy as np
import pytest

from artis_b200 import lib as ablib
from tests import fixtures, parity_checks

pytestmark = pytest.mark.gpu
CONFIG, NTS = "kilonova_2d_kat", 2
LIB = ablib.library_path("kilonova_lte")
SCHEDULES = {
    "wavefront": {"schedule": 1},
    "wavefront-notail": {"schedule": 1, "wf_tail": 0, "wf_sync_every": 3},
    "wavefront-refill": {"schedule": 1, "wf_tail": 0, "wf_refill_masteps": 5, "wf_refill_thicksteps": 3},
    "history": {"schedule": 0},
}


def test_known_answer_vectors_at_bench_scale():
    n_chi = parity_checks.check_deterministic_kernels(LIB, CONFIG, NTS)
    assert n_chi > 1000  # the continuum-opacity vectors are not skipped here


@pytest.mark.parametrize("schedule", sorted(SCHEDULES))
def test_packet_histories_and_sampled_tables_at_bench_scale(schedule):
    n, ncells = parity_checks.check_bench_scale_histories(LIB, CONFIG, NTS, options=SCHEDULES[schedule])
    assert n == 2000 and ncells == 6


def test_bench_scale_fixture_is_the_bench_model():
    fx = fixtures.load_golden(CONFIG, NTS)
    assert fx["static"]["line.nu"].size == 54892 and fx["static"]["cont.nu_edge"].size == 1475
    assert fx["static"]["cell.ffegrp"].size == 3684
    assert np.count_nonzero(fx["after"]["counters"]) > 5
