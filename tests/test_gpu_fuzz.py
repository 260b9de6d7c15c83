"""Randomised differential cases (tests/fuzz_util.py): random dataset layouts and sampler settings, the device sampler in
both RNG modes against the oracle.  Fixed seeds, so a failure reproduces."""

import numpy as np
import pytest

from tests.fuzz_util import run_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('seed', range(8))
def test_random_cases_match_the_oracle(seed):
    master = np.random.default_rng(1000 + seed)
    for _ in range(10):
        run_case(np.random.default_rng(int(master.integers(0, 2**31))))
