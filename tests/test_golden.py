"""The oracle against the committed golden vectors (made by the unmodified reference, tests/golden/make_golden.py)."""
import numpy as np
import pytest

import cases
from helpers import GOLDEN_SETS, assert_traces_equal, cmoracle, flat_for, load_golden


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_oracle_reproduces_golden(name):
    bsp, items = load_golden(name)
    o = cmoracle.OracleCM(flat_for(bsp))
    for case, want in items:
        got = cases.run_case(o, case)
        if case["kind"] == "points":
            assert np.array_equal(got, want), case["name"]
        else:
            assert_traces_equal(got, want, case["name"])


def test_golden_sets_are_not_trivial():
    for name in GOLDEN_SETS:
        _, items = load_golden(name)
        for case, want in items:
            if case["kind"] == "points":
                assert (want != 0).any(), case["name"]
            elif "position" in case["name"]:
                assert want["startsolid"].any(), case["name"]
            else:
                assert (want["fraction"] < 1).any() and (want["fraction"] == 1).any(), case["name"]
