"""A short slice of tests/fuzz_parity.py inside the GPU suite (bit-exact assignment and NMS
against the oracle on adversarial random cases)."""
import os

import pytest

pytestmark = pytest.mark.gpu


def test_fuzz_slice():
    import fuzz_parity
    assert fuzz_parity.run(cases=24, seed=1, workers=min(8, os.cpu_count() or 4)) == 0
