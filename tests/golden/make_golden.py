"""Generates tests/golden/ref_small.npz: outputs of the REFERENCE ITSELF (gpuvmem's own sources,
compiled unmodified for sm_100a into oracle/_ref/libgvref.so) on a small seeded synthetic problem.
Run on the GPU box (the reference's objective/gradient only exists as CUDA kernels):

    gpurun -- 'python tests/golden/make_golden.py'      # writes gpurun_out/golden/ref_small.npz
    cp gpurun_out/golden/ref_small.npz tests/golden/
    gpurun -- 'python tests/golden/make_golden.py ext'  # ref_small_ext.npz: error maps + degriddingGPU
    gpurun -- 'python tests/golden/make_golden.py priors'  # ref_small_priors.npz: every Fi kind on its own

The inputs are NOT stored: tests rebuild them from the same seed with gpuvmem_b200.synth.
tests/test_oracle_golden.py (CPU, no GPU needed) checks the C oracle against these vectors."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _checkers import GvRef  # noqa: E402
from gpuvmem_b200 import synth  # noqa: E402

PROBLEM = dict(N=128, nvis=3000, nchan=2, freq0=2.3e11, bandwidth=4e9, seed=101, grid_fill=1.02)
LAMBDAS = [0.01, 0.005, 0.002, 0.001]   # -Z: Entropy, L1-Norm, TSV, Laplacian (src/main.cu:193-197)
ARGS = "-X 16 -Y 16 -V 256 -z 0.001 -Z " + ",".join(map(str, LAMBDAS)) + " -t 3 -i synth.ms -o out.ms -m hdr.fits"


def golden_image(N, minpix, seed=3):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:N, 0:N]
    blob = np.exp(-((xx - N * 0.55) ** 2 + (yy - N * 0.45) ** 2) / (2 * (N / 16) ** 2))
    I = np.empty((2, N, N), np.float32)
    I[0] = (minpix * (1.0 + 40.0 * blob + 0.2 * rng.random((N, N)))).astype(np.float32)
    I[1] = (0.3 * blob + 0.05 * rng.standard_normal((N, N))).astype(np.float32)
    return I


def golden_grid(N, du, dv):
    """A closed-form Hermitian model grid, CENTRED (DC at [N/2][N/2]): the visibility function of three
    Gaussian components, G(u,v) = sum_s a_s exp(-(u^2+v^2)/(2 s_s^2)) exp(-2 pi i (u x_s + v y_s))."""
    k = np.arange(N) - N // 2
    u, v = np.meshgrid(k * abs(du), k * abs(dv))
    comps = [(1.0, 0.35, 0.0, 0.0), (0.6, 0.2, 0.21, -0.13), (0.3, 0.5, -0.4, 0.33)]   # (amp, width, x, y) in grid units
    g = np.zeros((N, N), np.complex128)
    umax = (N // 2) * abs(du)
    for a, wdt, x, y in comps:
        g += a * np.exp(-(u * u + v * v) / (2 * (wdt * umax) ** 2)) * np.exp(-2j * np.pi * (u * x + v * y) / abs(du) / N * 8)
    return g.astype(np.complex64)


PRIOR_GOLD = [("Entropy", 0), ("L1-Norm", 0), ("TotalVariation", 0), ("TotalSquaredVariation", 0), ("Laplacian", 0),
              ("Quadratic", 0), ("GEntropy", 0), ("GL1Norm", 0), ("TotalVariation", 1), ("Quadratic", 1)]
PRIOR_LAMBDA, PRIOR_EPS_B = 0.37, 1e-3


def prior_inputs(kind, index, N):
    """(image [2][N][N], prior image or None, epsilon_a) of one PRIOR_GOLD case."""
    I = golden_image(N, np.float32(0.001))
    if index == 1:
        I[1] = np.abs(I[1]) + np.float32(0.01)
    prior = (np.abs(I[index]) * 0.5 + 1e-4).astype(np.float32) if kind in ("GEntropy", "GL1Norm") else None
    return I, prior, (1e-12 if kind in ("L1-Norm", "GL1Norm") else 1e-6)


def main_priors():
    """Third fixture (ref_small_priors.npz): value and gradient of EVERY Fi kind the reference registers, each
    evaluated on its own by the reference build (gvref_prior_eval) — TV, Quadratic, GEntropy, GL1Norm are not
    wired by main.cu, so ref_small.npz does not cover them."""
    p = synth.make_problem(**PROBLEM)
    ref = GvRef()
    ref.set_problem(p)
    ref.init(ARGS)
    out = {"noise": ref.noise_image(), "noise_cut": np.float64(ref.scalars()["noise_cut"])}
    for kind, index in PRIOR_GOLD:
        I, prior, eps_a = prior_inputs(kind, index, p.N)
        v, dphi, after = ref.prior_eval(kind, I, PRIOR_LAMBDA, image_index=index, iteration=1, prior_image=prior,
                                        prior_value=0.001, eta=-1.0, eps_a=eps_a, eps_b=PRIOR_EPS_B)
        out[f"value_{kind}_{index}"] = np.float32(v)
        out[f"dphi_{kind}_{index}"] = dphi
        if after is not None:
            out[f"prior_after_{kind}_{index}"] = after
    d = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(d, exist_ok=True)
    np.savez_compressed(os.path.join(d, "ref_small_priors.npz"), **out)
    print("wrote", os.path.join(d, "ref_small_priors.npz"), sorted(out))


def main_ext():
    """Second fixture (ref_small_ext.npz): calculateErrors and the degriddingGPU kernel of the reference."""
    p = synth.make_problem(**PROBLEM)
    ref = GvRef()
    ref.set_problem(p)
    ref.init(ARGS)
    s = ref.scalars()
    I = golden_image(p.N, np.float32(0.001))
    ref.set_image(I)
    ref.calc_function(iteration=0)
    out = {"err_image": ref.error_image()}
    r = ref.get_vis(0)                                  # before cpu_ckernel: that call rebuilds the host datasets
    grid = golden_grid(p.N, s["deltau"], s["deltav"])
    # PSWF 9x9: the reference's Gaussian2D 7x7 table with its default w = 1 is a delta (src/gaussian2D.cu:27)
    table, support, _ = ref.cpu_ckernel("PSWF", 9, 9)
    out["degrid_table"] = table
    out["degrid_support"] = np.array(support)
    out["degrid_uvw"] = r["uvw"]
    out["degrid_Vm"] = ref.degridding(r["uvw"], grid, table, s["deltau"], s["deltav"], support[0], support[1])
    d = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(d, exist_ok=True)
    np.savez_compressed(os.path.join(d, "ref_small_ext.npz"), **out)
    print("wrote", os.path.join(d, "ref_small_ext.npz"), {k: v.shape for k, v in out.items()})


def main():
    p = synth.make_problem(**PROBLEM)
    ref = GvRef()
    ref.set_problem(p)
    ref.init(ARGS)
    s = ref.scalars()
    out = {"scalars_keys": np.array(sorted(s)), "scalars": np.array([s[k] for k in sorted(s)]),
           "noise": ref.noise_image(), "lambdas": np.array(LAMBDAS)}
    I = golden_image(p.N, np.float32(0.001))
    ref.set_image(I)
    v0, fi0 = ref.calc_function(iteration=0)          # priors gated off: 0.5*chi2
    out["half_chi2"] = np.float32(v0)
    out["image_after_clip"] = ref.get_image()
    for c in range(p.nchan):
        r = ref.get_vis(c)
        out[f"uvw_{c}"] = r["uvw"]; out[f"Vm_{c}"] = r["Vm"]; out[f"Vr_{c}"] = r["Vr"]; out[f"w_{c}"] = r["w"]
    out["grad_flag0"] = ref.calc_gradient(iteration=0, flag=0)
    out["grad_flag1"] = ref.calc_gradient(iteration=0, flag=1)
    v1, fi1 = ref.calc_function(iteration=1)          # priors active
    out["objective_it1"] = np.float32(v1)
    out["fi_it1"] = fi1
    out["grad_it1_flag0"] = ref.calc_gradient(iteration=1, flag=0)
    d = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(d, exist_ok=True)
    np.savez_compressed(os.path.join(d, "ref_small.npz"), **out)
    print("wrote", os.path.join(d, "ref_small.npz"), {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    {"ext": main_ext, "priors": main_priors}.get(sys.argv[1] if len(sys.argv) > 1 else "", main)()
