"""The numpy model behind the numbers DESIGN.md §3.3 quotes for the operand splits of the tensor-core gradient
(scripts/diag/mixed_split_model.py): the mixed split (fp16 product + 8-bit-float corrections) stays at ~1.5e-5 of a term
over wide amplitude distributions and down to populations 2^14 below the strongest visibility, dropping a low part costs
1.6e-4, and the expected round-toward-zero shrink of the accumulator is n * 2.07e-8."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("mixed_split_model", os.path.join(ROOT, "scripts", "diag", "mixed_split_model.py"))
model = importlib.util.module_from_spec(spec)
spec.loader.exec_module(model)


def test_8bit_float_grids():
    # E4M3: 3 mantissa bits, subnormal step 2^-9, max 448; E5M2: 2 mantissa bits, subnormal step 2^-16, max 57344
    assert model.e4m3(1.0625) == 1.0 and model.e4m3(1.1875) == 1.25 and model.e4m3(1000.0) == 448.0
    assert model.e4m3(2.0 ** -9) == 2.0 ** -9 and model.e4m3(2.0 ** -11) == 0.0
    assert model.e5m2(1.1) == 1.0 and model.e5m2(1.4) == 1.5 and model.e5m2(1e6) == 57344.0
    assert model.e5m2(2.0 ** -16) == 2.0 ** -16 and model.e5m2(-3.0 * 2.0 ** -16) == -3.0 * 2.0 ** -16


def test_mixed_split_error_over_amplitude_distributions():
    rng = np.random.default_rng(1)
    Z = 4000
    for amp in (np.full(Z, 1.0), rng.rayleigh(1.0, Z), np.exp(rng.normal(0, 3, Z)), np.r_[np.full(Z - 1, 1.0), 1000.0]):
        err = model.split_errors(amp / amp.max() * 2.0 ** 14, npix=32)
        assert err["fp16x3"] < 3e-7
        assert 5e-6 < err["mixed"] < 2.5e-5
        assert 1e-4 < err["two"] < 3e-4          # a second operand split is not optional at a 1e-5 target


def test_e5m2_keeps_the_correction_of_weak_populations():
    Z = 4000
    for R, lo, hi in ((1024, 0.0, 2.5e-5), (16384, 0.0, 2.5e-5), (2 ** 20, 1e-4, 3e-4)):
        err = model.split_errors(np.full(Z, 2.0 ** 14 / R), npix=32)
        assert lo <= err["mixed"] < hi, (R, err)
    # with E4M3 for the low part of the amplitude-carrying operand the correction is gone 2^12 below the maximum
    assert model.split_errors(np.full(Z, 2.0 ** 14 / 4096), npix=32)["mixed_e4m3_lo"] > 1e-4


def test_expected_truncation_shrink():
    # measured without the correction: 1.17e-5 at 512 instructions (mixed split, chunk 2048), 1.56e-5 at 768 (fp16x3)
    assert abs(model.truncation_shrink(512) - 1.17e-5) < 0.15e-5
    assert abs(model.truncation_shrink(768) - 1.56e-5) < 0.15e-5
    assert abs(model.truncation_shrink(1) - 2.1e-8) < 0.05e-8      # the constant in k_grad_umma's epilogue
