"""GPU: randomised differential test against the oracle (tests/fuzz_parity.py): random shapes,
centres inside and far outside the image, polynomial lengths, orders 0..5, all boundary modes,
six dtypes, perspective maps and row chunks of stacks -- every output bit-identical."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [11, 12])
def test_random_cases_bit_identical(seed):
    import fuzz_parity
    assert fuzz_parity.run(250, seed) == 0


@pytest.mark.gpu
def test_special_values():
    """NaN, +-Inf, signed zeros, denormals and near-overflow pixels (tests/special_values_check.py):
    every result equals the oracle's (NaN where it has NaN)."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    res = subprocess.run([sys.executable, os.path.join(here, "special_values_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    assert "0 cases with value mismatches" in res.stdout
