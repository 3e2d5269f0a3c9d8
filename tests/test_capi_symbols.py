"""The C-ABI library loads and exports every symbol include/artis_b200.h declares; without a GPU it refuses to
create a context (no CPU fallback). No compute calls here."""
import ctypes
import os
import re

import pytest

from artis_b200 import lib as ablib
from tests import fixtures

ROOT = fixtures.ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "artis_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(artisb200_[a-z0-9_]+)\s*\(", text)))


@pytest.mark.parametrize("preset", ["classic", "kilonova_lte"])
def test_library_exports_every_declared_symbol(preset):
    path = ablib.library_path(preset)
    assert os.path.exists(path), f"{path} missing: run `python __graft_entry__.py build`"
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/artis_b200.h but not exported by {path}"
    assert set(ablib.EXPORTED_SYMBOLS) == set(declared), "Python binding and header disagree on the symbol list"


def test_presets_differ_in_options_hash():
    a = ctypes.CDLL(ablib.library_path("classic"))
    b = ctypes.CDLL(ablib.library_path("kilonova_lte"))
    for lib in (a, b):
        lib.artisb200_options_hash.restype = ctypes.c_uint64
        lib.artisb200_options_summary.restype = ctypes.c_char_p
    assert a.artisb200_options_hash() != b.artisb200_options_hash()
    assert b"POL_ON=1" in a.artisb200_options_summary() and b"POL_ON=0" in b.artisb200_options_summary()


def test_no_cpu_fallback_without_a_device():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a CUDA device is present")
    with pytest.raises(ablib.ArtisB200Error) as err:
        ablib.ArtisB200(preset="classic")
    assert "CUDA" in str(err.value) or "cuda" in str(err.value)


def test_product_sources_do_not_touch_the_oracle():
    """oracle/ is test infrastructure: nothing shipped may include, import or execute it"""
    bad = []
    for sub in ("artis_b200", "include", "integration"):
        for dirpath, _dirs, files in os.walk(os.path.join(ROOT, sub)):
            if "_build" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                if not f.endswith((".py", ".h", ".cu", ".cc", ".c")):
                    continue
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for line in text.split("\n"):
                    if re.search(r"(#include|import|from|dlopen|CDLL|subprocess).*oracle[/_\.]", line) and "//" not in line.split("oracle")[0]:
                        bad.append((os.path.join(dirpath, f), line.strip()))
    assert not bad, bad
