"""Configs 1, 3 and 4 on float32 inputs at full size against the reference's float path (goldens:
tests/golden/ref_fullsize_c{1,3,4}f32.npz; the CPU oracle agrees with the reference on all three).  Added after the
round's GPU minutes were spent: this file sorts last so that its first GPU run cannot mask the verified suites."""
import pytest

import test_gpu_parity as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["c1f32", "c4f32", "c3f32"])
def test_full_size_float_twin_vs_reference_golden(name):
    P._check_full_size(name)
