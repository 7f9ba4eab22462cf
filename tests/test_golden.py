"""CPU: the oracle replays the golden fixtures of tests/golden/ — outputs of the REFERENCE'S OWN
sources (oracle/_ref/libref_f64.so, built by oracle/build_ref.py; generator oracle/gen_golden.py).
These fixtures are what pins the oracle on machines without /root/reference (the GPU box)."""
import os

import numpy as np
import pytest

from golden_replay import GOLDEN_DIR, compare, load_cases

CASES, run_case = load_cases()


@pytest.mark.parametrize("kind,tol", [(2, 1e-13), (0, 1e-12)], ids=["all-double", "parity-kind"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference_fixture(pkg, orc, name, kind, tol):
    gold = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    rec = run_case(pkg, name, CASES[name], lambda over: orc.OracleFilter(pkg.default_config(**over), kind=kind))
    worst = compare(rec, gold, tol, f"{name} kind={kind}")
    print(f"{name}: worst rel err vs reference fixture {worst:.2e}")


def test_fixture_inputs_are_reproducible(pkg):
    """The synthetic frames the fixtures were generated from regenerate bit-identically here."""
    for name, spec in CASES.items():
        gold = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        import hashlib
        sc = pkg.synth.Scene(**spec["scene"])
        h = hashlib.sha256()
        for t in range(1, sc.n_frames):
            h.update(sc.frame(t).tobytes())
        assert np.array_equal(np.frombuffer(h.digest(), dtype=np.uint8), gold["frames_sha256"]), name
