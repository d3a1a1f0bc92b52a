"""The C-ABI library builds, loads and exports every symbol include/bean_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

from crispr_bean_b200 import _lib
from crispr_bean_b200.build import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "bean_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bean_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_functions()
    assert "bean_ll_f32" in names and "bean_abi_version" in names
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/bean_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == names, "ctypes prototypes out of sync with the header"
    assert _lib.lib().bean_abi_version() == _lib.ABI_VERSION


def test_argument_validation_without_gpu():
    lib = _lib.lib()
    assert lib.bean_ll_f32(None, None, None) == -1
    assert b"screen is NULL" in lib.bean_last_error()
    assert lib.bean_ll_num_partials(129, 2) == 2      # thread per guide, 128 guides per CTA
    assert lib.bean_ll_num_partials(129, 231) == 33   # tiling: warp per guide, 4 guides per CTA
