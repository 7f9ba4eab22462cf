"""-m gpu: the CUDA path (through the C ABI) replays the golden fixtures generated from the
reference's own sources (tests/golden/, oracle/gen_golden.py).  Integer tables and match
coordinates bit-exact, state / covariance within 1e-9 relative (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

from golden_replay import GOLDEN_DIR, compare, load_cases
from helpers import TOL

pytestmark = pytest.mark.gpu
CASES, run_case = load_cases()


@pytest.mark.parametrize("name", sorted(n for n in CASES if not CASES[n].get("xyz")))
def test_cuda_path_reproduces_reference_fixture(gpu_pkg, name):
    gold = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))

    def make(over):
        over = dict(over)
        over["xyz_conversion"] = 0   # no feature of these cases reaches the linearity threshold (checked below)
        return gpu_pkg.VSlamFilter(gpu_pkg.default_config(**over), feature_capacity=CASES[name]["scene"]["n_features"] + 4)

    for k in gold.files:
        if k.endswith("_tab"):
            assert not gold[k][:, 2].any(), "fixture converted a feature to XYZ"
    rec = run_case(gpu_pkg, name, CASES[name], make)
    worst = compare(rec, gold, TOL, name)
    print(f"{name}: CUDA path worst rel err vs reference fixture {worst:.2e}")
