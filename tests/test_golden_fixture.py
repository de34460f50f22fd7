"""tests/golden/generator_rs_vectors.json (the reference's assert_eq! literals as data) against the
trees of tests/golden_cases.py and the CPU oracle."""
import json
import os

import numpy as np

from oracle.binding import OracleProgram
from tests.golden_cases import cases

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "generator_rs_vectors.json")


def test_fixture_matches_cases_and_oracle():
    doc = json.load(open(PATH))
    by_name = {v["name"]: v for v in doc["vectors"]}
    assert len(by_name) == len(cases()) >= 30
    for name, w, expected in cases():
        rec = by_name[name]
        assert rec["tree"] == repr(w), name
        want = np.asarray(rec["expected"], dtype=np.float32)
        np.testing.assert_array_equal(want, expected)
        for size in (1, 2, 4, 8):
            o = OracleProgram(w, 1)
            got = o.render(len(want), block=size) if len(want) else np.zeros(0, np.float32)
            np.testing.assert_array_equal(got, want, err_msg=f"{name} chunk {size}")
