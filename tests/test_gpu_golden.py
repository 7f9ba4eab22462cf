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


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_path_reproduces_reference_fixture(gpu_pkg, name):
    """Every fixture, including the XYZ-conversion case, with the reference's own configuration
    (convert2XYZ_ifLinearAll at the end of every update, vslamRansac.cpp:1317)."""
    gold = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))

    def make(over):
        return gpu_pkg.VSlamFilter(gpu_pkg.default_config(**over), feature_capacity=CASES[name]["scene"]["n_features"] + 4)

    rec = run_case(gpu_pkg, name, CASES[name], make)
    worst = compare(rec, gold, TOL, name)
    print(f"{name}: CUDA path worst rel err vs reference fixture {worst:.2e}")
