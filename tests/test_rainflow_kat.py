"""Known-answer test for the rainflow restatements (oracle C + the Python stand-in used to run the reference).

The only externally pinned vector for this path is the ASTM E1049-85 example from the README of the third-party
`rainflow` package (pinned ==3.2.0 by the reference's requirements.txt); SURVEY.md §4.
"""
import os
import sys

import numpy as np
import pytest

from oracle import oracle

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle", "refshim"))
import rainflow as py_rainflow  # noqa: E402

ASTM = [-2, 1, -3, 5, -1, 3, -4, 4, -2]
ASTM_CYCLES = [(3, -0.5, 0.5, 0, 1), (4, -1.0, 0.5, 1, 2), (4, 1.0, 1.0, 4, 5), (8, 1.0, 0.5, 2, 3),
               (9, 0.5, 0.5, 3, 6), (8, 0.0, 0.5, 6, 7), (6, 1.0, 0.5, 7, 8)]


def test_astm_example_c():
    assert oracle.rainflow_cycles(ASTM) == ASTM_CYCLES


def test_astm_example_py():
    assert list(py_rainflow.extract_cycles(ASTM)) == ASTM_CYCLES


@pytest.mark.parametrize("series,expected", [
    ([0.5], []), ([0.5, 0.7], []),                       # < 3 samples: no closing reversal, no cycles
    ([1.0, 1.0, 1.0], [(0.0, 1.0, 0.5, 0, 2)]),          # constant: one zero-range half cycle
    ([0.2, 0.4, 0.4, 0.4, 0.1], [(0.2, 0.30000000000000004, 0.5, 0, 3), (0.30000000000000004, 0.25, 0.5, 3, 4)]),
])
def test_edge_cases(series, expected):
    assert oracle.rainflow_cycles(series) == expected
    assert list(py_rainflow.extract_cycles(series)) == expected


def test_c_matches_python_random():
    rng = np.random.default_rng(0)
    for trial in range(200):
        n = int(rng.integers(1, 200))
        x = rng.random(n)
        rep = rng.random(n) < 0.3            # plateaus: repeat the previous sample exactly
        for i in range(1, n):
            if rep[i]:
                x[i] = x[i - 1]
        if trial % 3 == 0:
            x = np.round(x, 1)
        assert oracle.rainflow_cycles(x) == [tuple(c) for c in py_rainflow.extract_cycles(list(x))]
