"""Writes tests/golden/generator_rs_vectors.json: the known-answer vectors of the reference's own
generator tests (/root/reference/src/lib/generator.rs:1353-1925) as data, one record per case with
the citation of the Rust test it restates.  The trees that produce them are built by
tests/golden_cases.py (the reference cannot be run here — no cargo/rustc — so the vectors are the
literals of its `assert_eq!`s, not outputs of a run).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.golden_cases import cases, length_cases  # noqa: E402

CITES = {
    "time": "test_time generator.rs:1354-1357", "fixed": "test_fixed :1360-1372", "fin_marked": "test_fin :1375-1396",
    "reset": "test_reset :1543-1599", "append": "test_append :1602-1621", "add": "test_sum :1624-1675",
    "mul": "test_dot_product :1678-1736", "merge": "test_merge :1739-1777", "fir": "test_filter :1780-1903",
    "iir": "test_filter :1780-1903", "moving": "test_filter :1780-1903",
}


def cite(name):
    for k, v in CITES.items():
        if name.startswith(k):
            return v
    raise KeyError(name)


def main():
    doc = {
        "source": "/root/reference/src/lib/generator.rs:1353-1925 (assert_eq! literals; sample_rate 1; chunks 1/2/4/8)",
        "vectors": [{"name": n, "cite": cite(n), "tree": repr(w), "expected": [float(x) for x in e]}
                    for n, w, e in cases()],
        "lengths": [{"name": n, "position": p, "expected": e, "max": m, "tree": repr(w)}
                    for n, w, p, e, m in length_cases()],
    }
    with open(os.path.join(HERE, "generator_rs_vectors.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print(f"wrote {len(doc['vectors'])} vectors, {len(doc['lengths'])} length checks")


if __name__ == "__main__":
    main()
