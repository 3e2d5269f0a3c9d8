"""Shared helpers of the test-suite: golden fixtures, library builds, replay of one timestep."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from artis_b200 import lib as ablib  # noqa: E402
from artis_b200 import snapshot as snap  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
PRESET_OF = {"classic_toy": "classic", "classic_toy_1d": "classic", "classic3d_toy": "classic", "kilonova_toy": "kilonova_lte",
             "classic_multibin_toy": "classic_multibin", "classic_nlte_toy": "classic_nlte",
             "kilonova_guttman_toy": "kilonova_guttman", "kilonova_wollaeger_toy": "kilonova_wollaeger",
             "kilonova_barnes_toy": "kilonova_barnes", "classic_nt_toy": "classic_nt",
             "classic_ntexc_toy": "classic_ntexc", "classic_detailedbf_toy": "classic_detailedbf",
             "nltephot_toy": "nltephotospheric", "kilonova_2d_kat": "kilonova_lte",
             "kilonova_expansionopac_toy": "kilonova_expansionopac", "kilonova_expopac_retrace_toy": "kilonova_expopac_retrace",
             "kilonova_bbtherm_toy": "kilonova_bbtherm", "classic3d_grey_toy": "classic", "kilonova_xcom_toy": "kilonova_xcom"}
GOLDEN_TIMESTEPS = {"classic_toy": [0, 3], "classic_toy_1d": [0, 3], "classic3d_toy": [0, 2], "kilonova_toy": [1, 4],
                    "classic_multibin_toy": [2, 4], "classic_nlte_toy": [2, 4],
                    "kilonova_guttman_toy": [1], "kilonova_wollaeger_toy": [1], "kilonova_barnes_toy": [1],
                    "classic_nt_toy": [2, 3], "classic_ntexc_toy": [2, 3],
                    "classic_detailedbf_toy": [1, 3], "nltephot_toy": [1, 3],
                    "kilonova_expansionopac_toy": [2, 4], "kilonova_expopac_retrace_toy": [4], "kilonova_bbtherm_toy": [4],
                    "classic3d_grey_toy": [0, 2], "kilonova_xcom_toy": [1, 4]}
# presets compiled with USE_LUT_PHOTOION = false (csrc/options/preset_*.h)
PRESETS_WITHOUT_LUT_PHOTOION = {"classic_detailedbf", "nltephotospheric"}
INTERACTIONS = 26  # stats::Counter::INTERACTIONS (reference stats.h:41)


def load_golden(config, nts):
    static = dict(np.load(os.path.join(GOLDEN_DIR, f"{config}_static.npz")))
    ts = np.load(os.path.join(GOLDEN_DIR, f"{config}_ts{nts}.npz"))
    before = {k.split("/", 1)[1]: ts[k] for k in ts.files if k.startswith("before/")}
    after = {k.split("/", 1)[1]: ts[k] for k in ts.files if k.startswith("after/")}
    return {"config": config, "nts": nts, "static": static, "before": before, "after": after}


def hostsim_library(preset, defines=(), tag=""):
    """test-only single-threaded host build of the device headers (tests/hostsim/hostsim.cc); built on demand.
    `defines` / `tag`: a variant with extra -D flags (e.g. ARTISB200_HOSTSIM_FUZZ_LIBM)"""
    out = os.path.join(ROOT, "tests", "_build", f"libartis_b200_hostsim_{preset}{tag}.so")
    csrc = os.path.join(ROOT, "artis_b200", "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".h")] + [os.path.join(ROOT, "tests", "hostsim", "hostsim.cc")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        tmp = f"{out}.{os.getpid()}.tmp"  # pytest-xdist workers may build the same library at the same time
        subprocess.run(["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                        "-Wno-subobject-linkage", "-I" + csrc, *[f"-D{d}" for d in defines], f"-DARTISB200_PRESET_HEADER=\"options/preset_{preset}.h\"",
                        os.path.join(ROOT, "tests", "hostsim", "hostsim.cc"), "-o", tmp], check=True)
        os.replace(tmp, out)
    os.environ["ARTISB200_ALLOW_HOSTSIM"] = "1"
    return out


def make_engine(libpath, fx, rng="xoshiro", max_steps=0, device=0, seed=None, options=None):
    eng = ablib.ArtisB200(libpath=libpath, device=device)
    eng.set_option("rng_mode", 1 if rng == "xoshiro" else 0)
    eng.set_option("max_steps_per_launch", max_steps)
    for name, value in (options or {}).items():
        eng.set_option(name, value)
    if seed is not None:
        eng.set_option("seed", seed)
    eng.set_arrays(fx["static"])
    eng.commit_static()
    eng.set_arrays(fx["before"])
    eng.begin_timestep(fx["nts"])
    return eng


def run_fixture(libpath, fx, rng="xoshiro", max_steps=0, device=0, seed=None, options=None):
    """replay the fixture's timestep through the C ABI -> (packets structured array, estimators, built tables, timing)"""
    eng = make_engine(libpath, fx, rng, max_steps, device, seed, options)
    before = fx["before"]
    n = int(before["packets.count"][0])
    stride = int(before["packets.stride"][0])
    aos = before["packets.aos"].copy()
    eng.update_packets_host(fx["nts"], aos, n, stride)
    est = eng.estimators()
    built = {k: eng.get_array(k) for k in ["built.levelpops", "built.maprocessrates", "built.matrans", "built.cooling_contrib",
                                           "built.cont_nnlevel", "built.chi_ff_nnionpart", "built.corrphotoioncoeff",
                                           "built.cont_keepbits"]}
    timing = eng.last_timing_ms()
    eng.close()
    return aos.view(snap.packet_dtype(stride)), est, built, timing
