"""Short slices of tests/fuzz_parity.py (bit-exact assignment and NMS against the oracle on adversarial
random cases) and tests/fuzz_loss.py (loss kernels on random shapes / label mixes / logit distributions)
inside the GPU suite."""
import os

import pytest

pytestmark = pytest.mark.gpu


def test_fuzz_slice():
    import fuzz_parity
    assert fuzz_parity.run(cases=24, seed=1, workers=min(8, os.cpu_count() or 4)) == 0


def test_fuzz_loss_slice():
    import fuzz_loss
    assert fuzz_loss.run(cases=40, seed=1) == 40


def test_fuzz_select_slice():
    import fuzz_select
    assert fuzz_select.run(cases=30, seed=1) == 30


def test_fuzz_tail_slice():
    import fuzz_tail
    assert fuzz_tail.run(cases=30, seed=1) == 30
