"""Measures (and records) the reference's own float32-vs-float64 error on every golden case: tests/fp32_floor.py."""
import json
import os

import pytest

from tests.fp32_floor import DTYPE_DEPENDENT, NORTH_STAR_FP32, fp32_tolerance, reference_fp32_floor, same_function
from tests.test_reference_golden import PROGRAMS


@pytest.mark.parametrize("name", PROGRAMS)
def test_reference_float32_floor(name):
    truth, floor = reference_fp32_floor(name)
    print(name, json.dumps({k: float(f"{v:.3g}") for k, v in floor.items()}))
    if same_function(name):
        # the float32 run of the reference is a float32 computation of the same thing: its own error stays below 1e-4
        # everywhere, and its loss meets the north-star tolerance (the waivers are about gradients)
        assert all(v < 1e-4 for v in floor.values()), floor
        assert floor["loss"] < NORTH_STAR_FP32
    else:
        assert max(floor.values()) > 1e-4, "no longer dtype-dependent: move it out of fp32_floor.DTYPE_DEPENDENT"


def test_dtype_dependent_list_is_exact():
    bad = [n for n in PROGRAMS if same_function(n) and max(reference_fp32_floor(n)[1].values()) >= 1e-4]
    assert not bad, bad
    assert set(DTYPE_DEPENDENT) <= set(PROGRAMS)


def test_floor_table(tmp_path):
    """One table of every case's floor and the tolerance the fp32 kernels are held to (printed with -s)."""
    rows = {}
    for name in PROGRAMS:
        _, floor = reference_fp32_floor(name)
        rows[name] = {k: {"floor": float(f"{v:.3g}"), "tol": float(f"{fp32_tolerance(v):.3g}")} for k, v in floor.items()}
    waived = {n: {k: r for k, r in row.items() if r["tol"] > NORTH_STAR_FP32} for n, row in rows.items()}
    waived = {n: r for n, r in waived.items() if r}
    print(json.dumps(waived, indent=1))
    out = os.environ.get("BEAN_FLOOR_TABLE")
    if out:
        json.dump(rows, open(out, "w"), indent=1)
