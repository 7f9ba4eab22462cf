"""Replays a tests/golden case on any filter with the VSlamFilter method names and compares it with
the fixture (generated from the reference's own sources by oracle/gen_golden.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [p for p in (ROOT, os.path.join(ROOT, "oracle")) if p not in sys.path]
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def relerr(a, b):
    d = np.linalg.norm(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))
    s = np.linalg.norm(np.asarray(b, dtype=np.float64))
    return d / s if s > 0 else d


def load_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_golden_cases", os.path.join(ROOT, "oracle", "gen_golden.py"))
    src = open(spec.origin).read()
    # only the CASES table, run_case and controls_for are needed; importing the module would import refbind
    ns = {}
    start = src.index("INT_FIELDS = (")
    end = src.index("def main():")
    exec(compile("import hashlib\nimport numpy as np\n" + src[start:end], spec.origin, "exec"), ns)
    return ns["CASES"], ns["run_case"]


def compare(rec, gold, tol, ctx, int_exact=True):
    """Every key of the fixture against the replay.  Floats norm-wise relative, ints exact."""
    worst = 0.0
    for k in gold.files:
        g = gold[k]
        assert k in rec, f"{ctx}: replay has no {k}"
        r = np.asarray(rec[k])
        assert r.shape == g.shape, f"{ctx}: {k} shape {r.shape} vs golden {g.shape}"
        if g.dtype.kind in "iu" or k.endswith("_center"):
            if k.endswith("_tab"):
                # position_in_z is uninitialised in the reference until the first predict that sees the
                # feature in innovation (Patch.cpp:78-103): compare it only for in-innovation features
                cols = [c for c in range(g.shape[1]) if c != 1]
                assert np.array_equal(r[:, cols], g[:, cols]), f"{ctx}: {k} integer table differs"
                inn = g[:, 6] != 0
                assert np.array_equal(r[inn, 1], g[inn, 1]), f"{ctx}: {k} position_in_z differs"
            else:
                assert np.array_equal(r, g), f"{ctx}: {k} differs"
        else:
            e = relerr(r, g)
            worst = max(worst, e)
            assert e <= tol, f"{ctx}: {k} rel err {e:.3e} > {tol}"
    return worst
