"""CPU: the C-ABI library loads and exports every symbol include/*.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import glob
import os
import re

from visionllm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names.update(re.findall(r"\b(vllm_\w+)\s*\(", src))
    return sorted(names)


def test_header_declares_something():
    assert "vllm_msda_forward_f32" in declared_symbols()


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run `make` (or __graft_entry__.build()) first"
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    assert set(declared_symbols()) == set(_lib.exported_symbols())


def test_version_string():
    assert b"sm_100a" in _lib.lib().vllm_version()
