"""Pin the oracle's C restatement (oracle/artis_oracle.c) against
 (i)  the reference's own known-answer tests: the static_asserts beside closest_transition (rpkt.h:179-185),
      get_linedistance (rpkt.h:137-139), index_upperbound/lowerbound (sn3d.h:105-109), get_linearbinindex
      (sn3d.h:124-128), lowest_set_bit (constants.h:183-198), get_sigma_pair_prod_factor continuity
      (gammapkt.cc:512-513), and the unit tests test_vector_geometry / test_compton / test_random_sampling /
      test_closest_transition_randomised (unittests.cc:118-175, 323-355, 225-253, 435-460);
 (ii) golden vectors produced by the reference's own compiled functions (tests/golden/*.npz, written by
      tests/golden/make_golden.py from oracle/_ref)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from tests import fixtures

ROOT = fixtures.ROOT
CLIGHT = 2.99792458e10
SIGMA_T = 6.6524e-25


@pytest.fixture(scope="module")
def ao():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "restatement"], check=True)
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libartis_oracle.so"))
    D, I, I64, P = ctypes.c_double, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p
    sigs = {
        "ao_closest_transition": (I, [D, I, P, I]), "ao_get_linedistance": (D, [D, D, D, D, I]),
        "ao_index_upperbound": (I64, [P, I64, D]), "ao_index_lowerbound": (I64, [P, I64, D]),
        "ao_get_linearbinindex": (I64, [D, D, D]), "ao_lowest_set_bit": (I, [ctypes.c_uint64]),
        "ao_angle_ab": (None, [P, P, P]), "ao_doppler_nucmf_on_nurf": (D, [P, P, D, I]),
        "ao_sigma_compton_partial": (D, [D, D]), "ao_choose_f": (D, [D, D]), "ao_meanf_sigma": (D, [D]),
        "ao_sigma_pair_prod_factor": (D, [D]), "ao_planck": (D, [D, D]),
        "ao_xoshiro_seed": (None, [ctypes.c_uint32, P]), "ao_rng_uniform": (ctypes.c_float, [P]),
        "ao_boundary_distance": (D, [P, P, P, D, I, P]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_closest_transition_reference_static_asserts(ao):
    nu = np.array([9., 7., 5., 3.])  # rpkt.h:179
    cases = [(10., -1, 0), (8., -1, 1), (5., -1, 2), (2., -1, -1), (8., 2, 2), (8., 4, -1)]  # rpkt.h:180-185
    for nu_cmf, next_trans, expected in cases:
        assert ao.ao_closest_transition(nu_cmf, next_trans, _p(nu), 4) == expected


def test_closest_transition_randomised_like_reference_unittest(ao):
    # unittests.cc:435-460: 1000 random nu against a linear scan over a 500-line list, seed 777
    rng = np.random.default_rng(777)
    nu = np.sort(rng.uniform(1e14, 5e15, 500))[::-1].copy()
    for x in rng.uniform(0.9e14, 5.2e15, 1000):
        expected = -1
        if x >= nu[-1]:
            expected = next(i for i in range(500) if nu[i] <= x)
        assert ao.ao_closest_transition(float(x), -1, _p(nu), 500) == expected


def test_linedistance_static_asserts(ao):
    assert ao.ao_get_linedistance(100., 1., 2., -0.5, 0) == 0.  # rpkt.h:137
    assert ao.ao_get_linedistance(2., 4., 2., -1., 0) == CLIGHT * 2. * 2. / 2.  # rpkt.h:138
    assert ao.ao_get_linedistance(2., 4., 2., -1., 1) == 2.  # rpkt.h:139


def test_index_bounds_static_asserts(ao):
    v = np.array([1., 2., 2., 3.])  # sn3d.h:105-109
    assert ao.ao_index_upperbound(_p(v), 4, 2.) == 3
    assert ao.ao_index_upperbound(_p(v), 4, 0.5) == 0
    assert ao.ao_index_upperbound(_p(v), 4, 3.) == 4
    assert ao.ao_index_lowerbound(_p(v), 4, 2.) == 1
    assert ao.ao_index_lowerbound(_p(v), 4, 4.) == 4


def test_linearbinindex_static_asserts(ao):
    for args, expected in [((1.5, 1., 1.), 0), ((3., 1., 1.), 2), ((1., 1., 1.), 0), ((0.5, 1., 1.), -1), ((-5., 1., 2.), -3)]:
        assert ao.ao_get_linearbinindex(*args) == expected  # sn3d.h:124-128


def test_lowest_set_bit_static_asserts(ao):
    cases = {1: 0, 2: 1, 0b1100: 2, 2**64 - 1: 0, 1 << 31: 31, 1 << 32: 32, 1 << 33: 33, 1 << 62: 62, 1 << 63: 63,
             3 << 40: 40, ((2**64 - 1) << 17) & (2**64 - 1): 17}  # constants.h:183-198
    for bits, expected in cases.items():
        assert ao.ao_lowest_set_bit(bits) == expected


def test_compton_like_reference_unittest(ao):
    # unittests.cc:323-355: total cross section against the closed-form Klein-Nishina formula, choose_f inversion,
    # meanf_sigma continuity at the Thomson limit
    for x in [0.05, 0.3, 1.0, 3.0, 10.0]:
        kn = (3. / 4.) * SIGMA_T * (((1 + x) / x**3) * ((2 * x * (1 + x) / (1 + 2 * x)) - np.log(1 + 2 * x))
                                     + np.log(1 + 2 * x) / (2 * x) - (1 + 3 * x) / (1 + 2 * x)**2)
        assert abs(ao.ao_sigma_compton_partial(x, 1 + 2 * x) / kn - 1) < 1e-10
        for z in [0.1, 0.5, 0.9]:
            f = ao.ao_choose_f(x, z)
            assert 1. <= f <= 1 + 2 * x
            assert abs(ao.ao_sigma_compton_partial(x, f) / (z * ao.ao_sigma_compton_partial(x, 1 + 2 * x)) - 1) < 2e-4
    lim = 1e-2
    assert abs(ao.ao_meanf_sigma(lim * (1 - 1e-9)) / ao.ao_meanf_sigma(lim * (1 + 1e-9)) - 1) < 1e-6


def test_pair_production_fit_is_continuous(ao):
    nu = 3.61990e+20  # gammapkt.cc:512-513
    assert abs(ao.ao_sigma_pair_prod_factor(nu * (1 + 1e-12)) - ao.ao_sigma_pair_prod_factor(nu)) < 1e-6


def test_vector_geometry_like_reference_unittest(ao):
    # unittests.cc:118-173: aberration round trip to 1e-12, Doppler factor against the direct formula
    rng = np.random.default_rng(5)
    for _ in range(50):
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        v = rng.normal(size=3)
        v *= rng.uniform(0, 0.5) * CLIGHT / np.linalg.norm(v)
        out = np.zeros(3)
        back = np.zeros(3)
        ao.ao_angle_ab(_p(d), _p(v), _p(out))
        ao.ao_angle_ab(_p(out), _p(-v), _p(back))
        assert np.allclose(back, d, atol=1e-12)
        t = 1e5
        pos = v * t
        assert abs(ao.ao_doppler_nucmf_on_nurf(_p(pos), _p(d), t, 0) - (1 - np.dot(d, v) / CLIGHT)) < 1e-14
        gamma = 1 / np.sqrt(1 - np.dot(v, v) / CLIGHT**2)
        assert abs(ao.ao_doppler_nucmf_on_nurf(_p(pos), _p(d), t, 1) - gamma * (1 - np.dot(d, v) / CLIGHT)) < 1e-14


def test_rng_uniform_range_and_mean(ao):
    state = np.zeros(4, dtype=np.uint32)  # unittests.cc:225-253
    ao.ao_xoshiro_seed(12345, _p(state))
    x = np.array([ao.ao_rng_uniform(_p(state)) for _ in range(20000)])
    assert x.min() >= 0. and x.max() < 1.
    assert abs(x.mean() - 0.5) < 0.01


def test_rng_stream_matches_reference_packet_states():
    """the packets of the golden fixtures carry the reference's own per-packet Xoshiro128++ state: seeded with
    pre_zseed + packet index (input.cc:1911-1916) and then advanced by the draws packet_init() made for that
    packet. The restatement, seeded the same way, must reach exactly that state after a small number of draws."""
    import ctypes as ct
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "restatement"], check=True)
    lib = ct.CDLL(os.path.join(ROOT, "oracle", "_ref", "libartis_oracle.so"))
    lib.ao_xoshiro_seed.argtypes = [ct.c_uint32, ct.c_void_p]
    lib.ao_xoshiro_next.argtypes = [ct.c_void_p]
    lib.ao_xoshiro_next.restype = ct.c_uint32
    fx = fixtures.load_golden("classic_toy", 0)
    pk = fixtures.snap.packets_view(fx["before"])
    seed0 = 8  # tools/configs.py classic_toy run.seed
    for i in [0, 1, 7, 100, 1499]:
        state = np.zeros(4, dtype=np.uint32)
        lib.ao_xoshiro_seed(ct.c_uint32(seed0 + i), state.ctypes.data_as(ct.c_void_p))
        for _ in range(400):
            if np.array_equal(state, pk["rngstate"][i]):
                break
            lib.ao_xoshiro_next(state.ctypes.data_as(ct.c_void_p))
        assert np.array_equal(state, pk["rngstate"][i]), f"packet {i}: seeded stream never reaches the reference state"


class _Grid(ctypes.Structure):
    _fields_ = [("grid_type", ctypes.c_int), ("ncoord", ctypes.c_int * 3), ("coords", ctypes.c_void_p * 3),
                ("tmin", ctypes.c_double), ("rmax", ctypes.c_double), ("max_path_step", ctypes.c_double)]


@pytest.mark.parametrize("config,nts", [("classic_toy", 3), ("classic_toy_1d", 3), ("classic3d_toy", 2), ("kilonova_toy", 4)])
def test_boundary_distance_and_closest_transition_vs_reference_golden(ao, config, nts):
    fx = fixtures.load_golden(config, nts)
    st, after = fx["static"], fx["after"]
    g = _Grid()
    g.grid_type = int(st["scalar.grid_type"][0])
    coords = [np.ascontiguousarray(st[f"grid.coord_pos_min_tmin{d}"]) for d in range(3)]
    for d in range(3):
        g.ncoord[d] = int(st["scalar.ncoordgrid"][d])
        g.coords[d] = coords[d].ctypes.data if coords[d].size else None
    g.tmin, g.rmax = float(st["scalar.tmin"][0]), float(st["scalar.rmax"][0])
    g.max_path_step = float(after["kat.bd.max_path_step"][0])
    rec = after["kat.bd.in"].reshape(-1, 7)
    nxt = ctypes.c_int()
    for k in range(len(rec)):
        pos, dr = np.ascontiguousarray(rec[k, :3]), np.ascontiguousarray(rec[k, 3:6])
        dist = ao.ao_boundary_distance(ctypes.byref(g), _p(dr), _p(pos), float(rec[k, 6]), int(after["kat.bd.cell"][k]), ctypes.byref(nxt))
        assert nxt.value == after["kat.bd.next"][k]
        ref = after["kat.bd.dist"][k]
        assert dist == ref or abs(dist - ref) <= 1e-12 * abs(ref)
    nu = np.ascontiguousarray(st["line.nu"])
    for k in range(len(after["kat.ct.nu"])):
        assert ao.ao_closest_transition(float(after["kat.ct.nu"][k]), int(after["kat.ct.next_trans"][k]), _p(nu), len(nu)) == after["kat.ct.out"][k]
