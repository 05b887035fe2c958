"""Randomised frame parity (tools/soak.py): random layout, bit depth, picture size and block mix
(OBMC, inter-intra, intrabc, palette, CfL, filter-intra, warp, dense / packed coefficients)
through the batched path, bit-exact against the reference-driven oracle."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.gpu
@pytest.mark.parametrize("first", [7000, 7012, 7024])
def test_random_frames_bit_exact(first):
    import soak
    for seed in range(first, first + 12):
        ok, what = soak.run_one(seed)
        assert ok, what
