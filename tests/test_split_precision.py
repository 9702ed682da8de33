"""CPU: the numerical contract of the two tensor-core operand splits, emulated in numpy (tools/split_precision.py).
Pins the design decisions of DESIGN.md sections 4 / 5c: both engines are fp32-grade for O(1) operands; the 3xFP16
split degrades for operands below fp16's normal range (gradients) while 3xTF32 does not — which is why the backward
GEMMs use the 3xTF32 engine — and an exact power-of-two pre-scale restores the 3xFP16 split at any magnitude."""
import os
import sys
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
import split_precision as sp  # noqa: E402


@pytest.mark.parametrize('mag', [100.0, 1.0, 1e-2])
def test_both_splits_are_fp32_grade_in_the_forward_range(mag):
    e = sp.rel_errors(mag)
    assert e['fp16x3'] <= 2e-7 and e['tf32x3'] <= 6e-7 and e['fp32'] <= 1e-6


@pytest.mark.parametrize('mag', [1e-6, 1e-8, 1e-10])
def test_small_operands_need_tf32_or_a_power_of_two_scale(mag):
    e = sp.rel_errors(mag)
    assert e['fp16x3'] > 5e-6                      # fp16 hi is subnormal: not fp32-grade any more
    assert e['tf32x3'] <= 6e-7                     # fp32 exponent range: unaffected
    assert e['fp16x3_scaled'] <= 2e-7              # exact 2^k pre-scale: fp32-grade again
