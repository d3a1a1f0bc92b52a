"""The ctypes mirrors in crispr_bean_b200/_lib.py against the C structs of include/bean_b200.h: size and the offset of every
field, as gcc lays them out (no GPU needed; this is the check that the Python host and the .so agree on the ABI)."""
import ctypes as C
import os
import shutil
import subprocess

import pytest

from crispr_bean_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STRUCTS = [v for v in vars(_lib).values() if isinstance(v, type) and issubclass(v, C.Structure) and v is not C.Structure]


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_ctypes_structs_match_the_header(tmp_path):
    assert len(STRUCTS) >= 10
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT}/include/bean_b200.h"', "int main(void) {"]
    for s in STRUCTS:
        lines.append(f'  printf("{s.__name__} size %zu\\n", sizeof({s.__name__}));')
        for name, _ in s._fields_:
            lines.append(f'  printf("{s.__name__} {name} %zu\\n", offsetof({s.__name__}, {name}));')
    lines += ["  return 0;", "}"]
    src, exe = tmp_path / "layout.c", tmp_path / "layout"
    src.write_text("\n".join(lines))
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    want = {}
    for ln in out.splitlines():
        struct, field, val = ln.split()
        want[(struct, field)] = int(val)
    for s in STRUCTS:
        assert C.sizeof(s) == want[(s.__name__, "size")], s.__name__
        for name, _ in s._fields_:
            assert getattr(s, name).offset == want[(s.__name__, name)], (s.__name__, name)
