"""Numpy model of the operand splits of the tensor-core gradient: rounding of every operand element to
fp16 / E4M3 / E5M2 exactly as the generators do it, products and sums in fp64 (the tensor core forms the products of an
instruction exactly; its accumulator rounding is modelled separately, see `truncation_shrink`).

    A_k(i) = amp_k (cos, sin)(theta_ik)          B_k(j) = (cos, sin)(psi_jk)          d = sum_k A_k . B_k

    fp16x3 :  Ah Bh + Ah Bl16 + Al16 Bh                 (three fp16 products)
    mixed  :  Ah Bh + e5m2(Al) e4m3(Bh) + e5m2(Ah) e5m2(Bl)     (one fp16 product + one 8-bit product of twice the K)
    two    :  Ah Bh + Al16 Bh                           (what dropping one operand's low part would cost)

Run as a script for the table quoted in DESIGN.md; tests/test_mixed_split_model_cpu.py pins the bounds."""
import numpy as np


def fp8(x, mbits, emin, vmax):
    """Round to nearest even onto an 8-bit float grid: `mbits` mantissa bits, smallest normal exponent `emin`
    (subnormals below it), saturating at `vmax` (cvt.rn.satfinite)."""
    x = np.asarray(x, dtype=np.float64)
    s, a = np.sign(x), np.minimum(np.abs(x), vmax)
    e = np.maximum(np.floor(np.log2(np.maximum(a, 1e-300))), emin)
    step = 2.0 ** (e - mbits)
    return s * np.round(a / step) * step


def e4m3(x):
    return fp8(x, 3, -6, 448.0)


def e5m2(x):
    return fp8(x, 2, -14, 57344.0)


def f16(x):
    return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float64)


def split_errors(amp, npix=64, seed=0):
    """rel-L2 error of d over `npix` random pixel pairs for every split; amp: amplitudes scaled so that max = 2^14."""
    rng = np.random.default_rng(seed)
    Z = len(amp)
    th, ps = rng.uniform(0, 2 * np.pi, (npix, Z)), rng.uniform(0, 2 * np.pi, (npix, Z))
    A = (amp[None, :, None] * np.stack([np.cos(th), np.sin(th)], -1)).astype(np.float32).astype(np.float64)
    B = np.stack([np.cos(ps), np.sin(ps)], -1).astype(np.float32).astype(np.float64)
    exact = (A * B).sum((1, 2))
    Ah, Bh = f16(A), f16(B)
    Al, Bl = A - Ah, B - Bh
    out = {
        "fp16x3": Ah * Bh + Ah * f16(Bl) + f16(Al) * Bh,
        "mixed": Ah * Bh + e5m2(Al) * e4m3(Bh) + e5m2(Ah) * e5m2(Bl),
        "mixed_e4m3_lo": Ah * Bh + e4m3(Al) * e4m3(Bh) + e5m2(Ah) * e5m2(Bl),   # first version: E4M3 for Al
        "two": Ah * Bh + f16(Al) * Bh,
    }
    nrm = np.linalg.norm(exact)
    return {k: float(np.linalg.norm(v.sum((1, 2)) - exact) / nrm) for k, v in out.items()}


def truncation_shrink(n_instructions):
    """Expected relative shrink of an accumulator that grows from 0 over n round-toward-zero additions: each loses on
    average half an ulp, 0.5 * 2^-23 * E[1/m] of |acc| with the mantissa m uniform in [1, 2)."""
    return n_instructions * 0.5 * 0.5 * 2.0 ** -23 * np.log(2.0)


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    Z = 20000
    cases = [("constant", np.full(Z, 1.0)), ("Rayleigh", rng.rayleigh(1.0, Z)), ("log-normal s=2", np.exp(rng.normal(0, 2, Z))),
             ("log-normal s=3", np.exp(rng.normal(0, 3, Z))), ("one outlier x1000", np.r_[np.full(Z - 1, 1.0), 1000.0])]
    for name, amp in cases:
        print(f"{name:20s}", {k: f"{v:.2e}" for k, v in split_errors(amp / amp.max() * 2.0 ** 14).items()})
    print("a population at max / R alone (the fp16 scale is set by an outlier that is not part of the sum):")
    for R in (1, 64, 1024, 4096, 16384, 2 ** 17, 2 ** 20):
        print(f"R = {R:8d}", {k: f"{v:.2e}" for k, v in split_errors(np.full(Z, 2.0 ** 14 / R)).items()})
    for n, what in ((512, "mixed split, 2048 visibilities"), (768, "fp16x3, 2048 visibilities"), (1024, "mixed split, 4096 visibilities")):
        print(f"expected truncation shrink over {n} instructions ({what}): {truncation_shrink(n):.2e}")
