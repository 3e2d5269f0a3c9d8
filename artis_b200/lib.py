"""ctypes binding of the C ABI in include/artis_b200.h.

The product library is artis_b200/_build/libartis_b200_<preset>.so (CUDA, sm_100a). There is no CPU
fallback: if the library is missing or no CUDA device is usable, construction raises."""
import ctypes
import os

import numpy as np

from . import snapshot as snap_mod

_HERE = os.path.dirname(os.path.abspath(__file__))


def library_path(preset):
    # ARTISB200_LIB_SUFFIX selects a side-by-side tuning build (see __graft_entry__.build_cuda); default: the product
    return os.path.join(_HERE, "_build", f"libartis_b200_{preset}{os.environ.get('ARTISB200_LIB_SUFFIX', '')}.so")


class ArtisB200Error(RuntimeError):
    pass


_SIGNATURES = {
    "artisb200_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
    "artisb200_destroy": (None, [ctypes.c_void_p]),
    "artisb200_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "artisb200_options_hash": (ctypes.c_uint64, []),
    "artisb200_options_summary": (ctypes.c_char_p, []),
    "artisb200_set_array": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char, ctypes.c_void_p, ctypes.c_int64]),
    "artisb200_get_array": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char, ctypes.c_void_p, ctypes.c_int64]),
    "artisb200_array_count": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_char_p]),
    "artisb200_get_array_range": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char, ctypes.c_void_p, ctypes.c_int64,
                                                 ctypes.c_int64]),
    "artisb200_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]),
    "artisb200_commit_static": (ctypes.c_int, [ctypes.c_void_p]),
    "artisb200_bin_escaped_packets": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "artisb200_last_binning_ms": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]),
    "artisb200_write_text_packets": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int]),
    "artisb200_read_text_packets": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                                   ctypes.POINTER(ctypes.c_int64)]),
    "artisb200_write_temp_packetsfile": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]),
    "artisb200_read_temp_packetsfile": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                                       ctypes.POINTER(ctypes.c_int64)]),
    "artisb200_update_grid_lte": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double]),
    "artisb200_last_gridupdate_ms": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]),
    "artisb200_begin_timestep": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "artisb200_upload_packets": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]),
    "artisb200_download_packets": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]),
    "artisb200_update_packets": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "artisb200_update_packets_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]),
    "artisb200_register_host_buffer": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]),
    "artisb200_unregister_host_buffer": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "artisb200_save_packets_device": (ctypes.c_int, [ctypes.c_void_p]),
    "artisb200_restore_packets_device": (ctypes.c_int, [ctypes.c_void_p]),
    "artisb200_estimator_device_buffer": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64)]),
    "artisb200_last_timing_ms": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "artisb200_last_schedule_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                                      ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64),
                                                      ctypes.POINTER(ctypes.c_int64)]),
    "artisb200_test_kernel": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p]),
    "artisb200_stream": (ctypes.c_void_p, [ctypes.c_void_p]),
}

EXPORTED_SYMBOLS = sorted(_SIGNATURES)

ESTIMATOR_NAMES = ["est.J", "est.nuJ", "est.ffheating", "est.colheating", "est.gamma", "est.bfheating", "est.dep_gamma",
                   "est.dep_positron", "est.dep_electron", "est.dep_alpha", "ts.scalars", "ts.pellet_decays", "counters", "diag"]
OPTIONAL_ESTIMATOR_NAMES = ["est.bins_J_raw", "est.bins_nuJ_raw", "est.bfrate_raw"]  # MULTIBIN / DETAILED_BF presets only
_OUT_DTYPES = {"ts.pellet_decays": np.int64, "counters": np.int64, "diag": np.int64, "diag_stage": np.int64, "dev_error": np.int64, "built.cont_keepbits": np.uint64,
               "gridupdate.uppermost_ion": np.int32, "gridupdate.status": np.int32,
               "cell.Te": np.float32, "cell.TJ": np.float32, "cell.TR": np.float32, "cell.W": np.float32, "cell.nne": np.float32,
               "cell.ion_partfuncts": np.float32, "cell.ion_groundlevelpops": np.float32,
               "spec.lower_freq": np.float32, "spec.delta_freq": np.float32, "spec.dirbin": np.int32}


def load_library(path):
    if not os.path.exists(path):
        raise ArtisB200Error(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


class ArtisB200:
    """One library context on one device. Mirrors the call sequence of the reference driver:
    static tables once -> per timestep (cell state, begin_timestep, update_packets, read estimators)."""

    def __init__(self, preset=None, device=0, libpath=None):
        self.libpath = libpath or library_path(preset)
        self.lib = load_library(self.libpath)
        handle = ctypes.c_void_p()
        if self.lib.artisb200_create(ctypes.byref(handle), device) != 0:
            raise ArtisB200Error(self.lib.artisb200_last_error(None).decode())
        self.ctx = handle

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.artisb200_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise ArtisB200Error(f"{what}: {self.lib.artisb200_last_error(self.ctx).decode()}")

    def options_summary(self):
        return self.lib.artisb200_options_summary().decode()

    def set_option(self, name, value):
        self._check(self.lib.artisb200_set_option(self.ctx, name.encode(), int(value)), f"set_option({name})")

    def set_array(self, name, arr):
        arr = np.ascontiguousarray(arr)
        code = snap_mod.dtype_code(arr).encode()
        self._check(self.lib.artisb200_set_array(self.ctx, name.encode(), code, arr.ctypes.data_as(ctypes.c_void_p), arr.size),
                    f"set_array({name})")

    def set_arrays(self, arrays, skip_prefixes=("packets.", "est.", "ts.", "counters", "diag")):
        for name, arr in arrays.items():
            if name.startswith(skip_prefixes):
                continue
            self.set_array(name, arr)

    def array_count(self, name):
        return int(self.lib.artisb200_array_count(self.ctx, name.encode()))

    def get_array(self, name, dtype=None):
        n = self.array_count(name)
        if n < 0:
            raise ArtisB200Error(f"array {name} not available")
        dt = np.dtype(dtype or _OUT_DTYPES.get(name, np.float64))
        out = np.empty(n, dtype=dt)
        code = snap_mod.dtype_code(out).encode()
        self._check(self.lib.artisb200_get_array(self.ctx, name.encode(), code, out.ctypes.data_as(ctypes.c_void_p), n),
                    f"get_array({name})")
        return out

    def get_array_range(self, name, offset, count, dtype=None):
        dt = np.dtype(dtype or _OUT_DTYPES.get(name, np.float64))
        out = np.empty(int(count), dtype=dt)
        code = snap_mod.dtype_code(out).encode()
        self._check(self.lib.artisb200_get_array_range(self.ctx, name.encode(), code, out.ctypes.data_as(ctypes.c_void_p), int(offset),
                                                       int(count)), f"get_array_range({name})")
        return out

    def commit_static(self):
        self._check(self.lib.artisb200_commit_static(self.ctx), "commit_static")

    def begin_timestep(self, nts):
        self._check(self.lib.artisb200_begin_timestep(self.ctx, int(nts)), "begin_timestep")

    def upload_packets(self, aos_bytes, npackets, stride):
        aos_bytes = np.ascontiguousarray(aos_bytes)
        self._check(self.lib.artisb200_upload_packets(self.ctx, aos_bytes.ctypes.data_as(ctypes.c_void_p), npackets, stride),
                    "upload_packets")

    def download_packets(self, aos_bytes, npackets, stride):
        self._check(self.lib.artisb200_download_packets(self.ctx, aos_bytes.ctypes.data_as(ctypes.c_void_p), npackets, stride),
                    "download_packets")

    def update_packets(self, nts):
        self._check(self.lib.artisb200_update_packets(self.ctx, int(nts)), "update_packets")

    def update_packets_host(self, nts, aos_bytes, npackets, stride):
        """the drop-in call: host AoS packets in, propagated host AoS packets out (in place)"""
        self._check(self.lib.artisb200_update_packets_host(self.ctx, int(nts), aos_bytes.ctypes.data_as(ctypes.c_void_p),
                                                           npackets, stride), "update_packets_host")

    def bin_escaped_packets(self, direction_bins=False, emission_absorption=0, nprocs_exspec=1):
        """spectra and light curves of the device-resident packets in one pass (include/artis_b200.h); read the results
        with get_array("spec.flux") ... or artis_b200.spectra.binned()"""
        self._check(self.lib.artisb200_bin_escaped_packets(self.ctx, int(bool(direction_bins)), int(emission_absorption),
                                                           int(nprocs_exspec)), "bin_escaped_packets")

    def last_binning_ms(self):
        ms = ctypes.c_double()
        self.lib.artisb200_last_binning_ms(self.ctx, ctypes.byref(ms))
        return ms.value

    def update_grid_lte(self, temperatures_from_J=False, mintemp=0., maxtemp=0.):
        """LTE part of update_grid_cell on the device copies of the cell state (include/artis_b200.h): partition functions,
        Saha ion balance, electron density; optionally the temperatures from the J estimator first. Read the results with
        get_array("cell.nne") / "cell.ion_groundlevelpops" / "cell.ion_partfuncts" / "cell.Te" / "gridupdate.uppermost_ion"."""
        self._check(self.lib.artisb200_update_grid_lte(self.ctx, int(bool(temperatures_from_J)), float(mintemp), float(maxtemp)),
                    "update_grid_lte")

    def last_gridupdate_ms(self):
        ms = ctypes.c_double()
        self.lib.artisb200_last_gridupdate_ms(self.ctx, ctypes.byref(ms))
        return ms.value

    def write_text_packets(self, filename, aos_bytes, npackets, stride, keep_escaped_gammas=True):
        """packets*.out of the reference (packet.cc:226-251) from the AoS packet array"""
        aos_bytes = np.ascontiguousarray(aos_bytes)
        self._check(self.lib.artisb200_write_text_packets(self.ctx, os.fsencode(filename), aos_bytes.ctypes.data_as(ctypes.c_void_p),
                                                          int(npackets), int(stride), int(bool(keep_escaped_gammas))), "write_text_packets")

    def read_text_packets(self, filename, stride=240):
        """-> (raw AoS bytes, packet count) of a packets*.out text file, read like exspec reads it (packet.cc:163-222)"""
        n = ctypes.c_int64()
        self._check(self.lib.artisb200_read_text_packets(self.ctx, os.fsencode(filename), None, 0, int(stride), ctypes.byref(n)),
                    "read_text_packets")
        out = np.zeros(n.value * int(stride), dtype=np.uint8)
        self._check(self.lib.artisb200_read_text_packets(self.ctx, os.fsencode(filename), out.ctypes.data_as(ctypes.c_void_p), n.value,
                                                         int(stride), ctypes.byref(n)), "read_text_packets")
        return out, n.value

    def write_temp_packetsfile(self, filename, aos_bytes, npackets, stride):
        """binary restart file packets_<rank>_ts<N>.tmp of the reference (packet.cc:273-311)"""
        aos_bytes = np.ascontiguousarray(aos_bytes)
        self._check(self.lib.artisb200_write_temp_packetsfile(self.ctx, os.fsencode(filename), aos_bytes.ctypes.data_as(ctypes.c_void_p),
                                                              int(npackets), int(stride)), "write_temp_packetsfile")

    def read_temp_packetsfile(self, filename, stride):
        """-> (raw AoS bytes, packet count) of a binary restart file (packet.cc:253-271)"""
        n = ctypes.c_int64()
        self._check(self.lib.artisb200_read_temp_packetsfile(self.ctx, os.fsencode(filename), None, 0, int(stride), ctypes.byref(n)),
                    "read_temp_packetsfile")
        out = np.empty(n.value * int(stride), dtype=np.uint8)
        self._check(self.lib.artisb200_read_temp_packetsfile(self.ctx, os.fsencode(filename), out.ctypes.data_as(ctypes.c_void_p), n.value,
                                                             int(stride), ctypes.byref(n)), "read_temp_packetsfile")
        return out, n.value

    def save_packets_device(self):
        self._check(self.lib.artisb200_save_packets_device(self.ctx), "save_packets_device")

    def restore_packets_device(self):
        self._check(self.lib.artisb200_restore_packets_device(self.ctx), "restore_packets_device")

    def estimator_device_buffer(self):
        ptr = ctypes.c_void_p()
        n = ctypes.c_int64()
        self._check(self.lib.artisb200_estimator_device_buffer(self.ctx, ctypes.byref(ptr), ctypes.byref(n)), "estimator buffer")
        return ptr.value, n.value

    def last_timing_ms(self):
        a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        self.lib.artisb200_last_timing_ms(self.ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return a.value, b.value, c.value

    def last_schedule_stats(self):
        stage = (ctypes.c_double * 4)()
        tail_ms = ctypes.c_double()
        tail_n, iters, launches = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        self.lib.artisb200_last_schedule_stats(self.ctx, stage, ctypes.byref(tail_ms), ctypes.byref(tail_n), ctypes.byref(iters),
                                               ctypes.byref(launches))
        return {"stage_ms": dict(zip(("other", "rpkt_thin", "rpkt_thick", "macroatom"), list(stage))), "tail_ms": tail_ms.value,
                "tail_packets": tail_n.value, "iterations": iters.value, "launches": launches.value}

    def stream(self):
        return self.lib.artisb200_stream(self.ctx)

    def test_kernel(self, which, in_f64, in_i32):
        """element-wise evaluation of a deterministic device function (see include/artis_b200.h)"""
        in_f64 = np.ascontiguousarray(in_f64, dtype=np.float64)
        in_i32 = np.ascontiguousarray(in_i32, dtype=np.int32)
        n = in_i32.size
        expect = {"boundary_distance": 7 * n, "select_continuum_nu": 2 * n}.get(which, n)
        if in_f64.size != expect:
            raise ValueError(f"test_kernel({which}): {in_f64.size} f64 inputs for {n} items, expected {expect}")
        out_f64 = np.zeros(3 * n if which == "chi_rpkt_cont" else n, dtype=np.float64)
        out_i32 = np.zeros(n, dtype=np.int32)
        self._check(self.lib.artisb200_test_kernel(self.ctx, which.encode(), n, in_f64.ctypes.data_as(ctypes.c_void_p),
                                                   in_i32.ctypes.data_as(ctypes.c_void_p), out_f64.ctypes.data_as(ctypes.c_void_p),
                                                   out_i32.ctypes.data_as(ctypes.c_void_p)), f"test_kernel({which})")
        return out_f64, out_i32

    def estimators(self):
        out = {name: self.get_array(name) for name in ESTIMATOR_NAMES}
        out.update({name: self.get_array(name) for name in OPTIONAL_ESTIMATOR_NAMES if self.array_count(name) >= 0})
        return out
