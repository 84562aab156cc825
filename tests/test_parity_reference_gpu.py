"""GPU parity against the REFERENCE ITSELF: gpuvmem's own sources compiled unmodified
for sm_100a (oracle/_ref/libgvref.so, built by oracle/Makefile from /root/reference
in the build container; the .so travels to the GPU box, the sources do not).

Everything the reference derives or computes on this path is compared on identical
synthetic input: the setup scalars and noise image (MFS::configure/setDevice), the
folded uvw in wavelengths (bit-exact), 0.5*chi2 (rel 1e-5), residuals, the chi2
gradient for both optimisation flags (rel-L2 1e-4), every prior value/gradient that
main.cu wires, and the assembled objective/gradient with priors active.

The reference library keeps its state in process globals, so ONE problem is
initialised per test process (module-scoped fixture)."""
import os

import numpy as np
import pytest

from gpuvmem_b200 import Engine, synth
from gpuvmem_b200.engine import GRAD_AUTO, GRAD_SIMT

pytestmark = pytest.mark.gpu

LAMBDAS = [0.01, 0.005, 0.002, 0.001]  # -Z: Entropy, L1-Norm, TSV, Laplacian (src/main.cu:193-197)
ARGS = "-X 16 -Y 16 -V 256 -z 0.001 -Z " + ",".join(map(str, LAMBDAS)) + " -t 3 -i synth.ms -o out.ms -m hdr.fits"


def _image(e, seed=3):
    rng = np.random.default_rng(seed)
    N = e.N
    yy, xx = np.mgrid[0:N, 0:N]
    I = e.initial_image()
    blob = np.exp(-((xx - N * 0.55) ** 2 + (yy - N * 0.45) ** 2) / (2 * (N / 16) ** 2))
    I[0] = (e.meta["minpix"] * (1.0 + 40.0 * blob + 0.2 * rng.random((N, N)))).astype(np.float32)
    I[1] = (0.3 * blob + 0.05 * rng.standard_normal((N, N))).astype(np.float32)
    return I


@pytest.fixture(scope="module")
def setup(gvref):
    import torch
    p = synth.make_problem(N=256, nvis=40000, nchan=2, freq0=2.3e11, bandwidth=4e9, seed=21, grid_fill=1.03)
    gvref.set_problem(p)
    gvref.init(ARGS)
    e = Engine.from_problem(p, keep_vm=True, grad_mode=GRAD_AUTO)
    yield p, e, gvref, torch
    e.close()


def test_setup_scalars_match_reference(setup):
    p, e, ref, _ = setup
    s = ref.scalars()
    m = e.meta
    assert s["deltau"] == m["deltau"] and s["deltav"] == m["deltav"]
    assert s["xpix"] == m["xpix"] and s["ypix"] == m["ypix"]
    assert np.float32(s["nu_0"]) == np.float32(m["nu_0"])
    assert np.float32(s["pb_cutoff"]) == np.float32(m["pb_cutoff"])
    assert np.float32(s["pb_factor"]) == np.float32(m["pb_factor"])
    assert abs(s["vis_noise"] - m["vis_noise"]) <= 1e-6 * s["vis_noise"]
    assert abs(s["noise_jypix"] - m["noise_jypix"]) <= 1e-5 * s["noise_jypix"]
    assert abs(s["fg_scale"] - m["fg_scale"]) <= 1e-5 * s["fg_scale"]
    assert abs(s["noise_cut"] - m["noise_cut"]) <= 1e-5 * s["noise_cut"]


def test_noise_image_matches_reference(setup):
    p, e, ref, _ = setup
    a, b = e.get_noise_image(), ref.noise_image()
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    np.testing.assert_allclose(a[fin], b[fin], rtol=2e-5)
    # same mask
    s = ref.scalars()
    assert np.array_equal(a < e.meta["noise_cut"], b < s["noise_cut"])


def test_uploaded_visibilities_bit_exact(setup):
    p, e, ref, _ = setup
    for c in range(p.nchan):
        r = ref.get_vis(c)
        g = e.get_vis(c, want=("uvw", "Vo"))
        assert np.array_equal(g["uvw"].view(np.uint64), r["uvw"].view(np.uint64)), "hermitianSymmetry + metres->lambda"
        assert np.array_equal(g["Vo"].view(np.uint32), r["Vo"].view(np.uint32))


def test_chi2_and_residuals_match_reference(setup):
    p, e, ref, torch = setup
    I = _image(e)
    ref.set_image(I)
    want, fi = ref.calc_function(iteration=0)   # priors gated off at iteration 0 -> 0.5*chi2
    I_dev = torch.from_numpy(I).cuda()
    got = e.chi2(I_dev)
    assert abs(got - want) <= 1e-5 * abs(want), (got, want)
    assert abs(fi[0] - want) <= 1e-6 * abs(want)
    assert np.array_equal(ref.get_image().view(np.uint32), I_dev.cpu().numpy().view(np.uint32)), "clip2IWNoise"
    for c in range(p.nchan):
        r = ref.get_vis(c)
        g = e.get_vis(c, want=("Vm", "Vr", "w"))
        assert np.array_equal(g["w"].view(np.uint32), r["w"].view(np.uint32)), "vis_mod zeroes off-grid weights"
        scale = np.abs(r["Vm"]).max()
        on = r["w"] > 0
        assert np.abs(g["Vm"][on] - r["Vm"][on]).max() <= 2e-5 * scale
        assert np.abs(g["Vr"][on] - r["Vr"][on]).max() <= 2e-5 * max(scale, np.abs(r["Vr"]).max())


@pytest.mark.parametrize("flag", [0, 1])
def test_chi2_gradient_matches_reference(setup, flag):
    p, e, ref, torch = setup
    I = _image(e)
    ref.set_image(I)
    ref.calc_function(iteration=0)
    want = ref.calc_gradient(iteration=0, flag=flag)   # priors gated -> pure chi2 gradient
    I_dev = torch.from_numpy(I).cuda()
    e.set_flag_opt(flag)
    e.chi2(I_dev)
    g = torch.zeros_like(I_dev)
    e.dchi2(I_dev, g, flag_opt=flag)
    got = g.cpu().numpy()
    err = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert err <= 1e-4, (flag, err, e.last_grad_mode())
    assert np.array_equal(got == 0, want == 0) or np.count_nonzero((got == 0) != (want == 0)) < 4


def test_priors_match_reference(setup):
    p, e, ref, torch = setup
    I = _image(e)
    ref.set_image(I)
    total, fi = ref.calc_function(iteration=1)   # priors active
    I_dev = torch.from_numpy(I).cuda()
    chi2 = e.chi2(I_dev)
    vals = [e.prior_value("Entropy", I_dev, 0, prior_value=0.001, eta=-1.0),
            e.prior_value("L1-Norm", I_dev, 0, epsilon=1e-12),
            e.prior_value("TotalSquaredVariation", I_dev, 0),
            e.prior_value("Laplacian", I_dev, 0)]
    for k, v in enumerate(vals):
        assert abs(v - fi[k + 1]) <= 2e-5 * abs(fi[k + 1]), (k, v, fi[k + 1])
    mine = chi2 + sum(l * v for l, v in zip(LAMBDAS, vals))
    assert abs(mine - total) <= 2e-5 * abs(total)


def test_full_gradient_with_priors_matches_reference(setup):
    p, e, ref, torch = setup
    I = _image(e)
    ref.set_image(I)
    ref.calc_function(iteration=1)
    want = ref.calc_gradient(iteration=1, flag=0)
    I_dev = torch.from_numpy(I).cuda()
    e.chi2(I_dev)
    dphi = torch.zeros_like(I_dev)
    e.dchi2(I_dev, dphi, flag_opt=0)               # Chi2::addToDphi overwrites dphi with result_dchi2
    dgi = torch.empty(p.N, p.N, device="cuda")
    for kind, lam, kw in [("Entropy", LAMBDAS[0], dict(prior_value=0.001, eta=-1.0)),
                          ("L1-Norm", LAMBDAS[1], dict(epsilon=1e-12)),
                          ("TotalSquaredVariation", LAMBDAS[2], {}), ("Laplacian", LAMBDAS[3], {})]:
        e.prior_grad(kind, I_dev, dgi, lam, 0, **kw)
        e.add_to_dphi(dphi, dgi, 0)
    got = dphi.cpu().numpy()
    err = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert err <= 1e-4, err


def test_reference_distance_to_fp64_truth(setup, oracle):
    """Reports (and bounds) how far the reference's own fp32 gradient is from the fp64
    oracle next to how far the engine is: the engine must be at least as close."""
    p, e, ref, torch = setup
    from test_parity_gpu import _grad_oracle_sample
    I = _image(e)
    ref.set_image(I)
    ref.calc_function(iteration=0)
    gref = ref.calc_gradient(iteration=0, flag=0)[0].reshape(-1)
    I_dev = torch.from_numpy(I).cuda()
    e.chi2(I_dev)
    g = torch.zeros_like(I_dev)
    e.dchi2(I_dev, g, flag_opt=0)
    gmine = g.cpu().numpy()[0].reshape(-1)
    pix = np.arange(0, p.N * p.N, 97)
    truth = _grad_oracle_sample(oracle, p, e, I_dev.cpu().numpy(), pix, 0)
    err_ref = np.linalg.norm(gref[pix] - truth) / np.linalg.norm(truth)
    err_mine = np.linalg.norm(gmine[pix] - truth) / np.linalg.norm(truth)
    print(f"\n[parity] rel-L2 to fp64 oracle: reference={err_ref:.3e} engine={err_mine:.3e}")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_distances.txt"), "a") as f:
        f.write(f"N={p.N} Z={p.total_vis()} mode={e.last_grad_mode()} ref_vs_fp64={err_ref:.4e} engine_vs_fp64={err_mine:.4e}\n")
    assert err_ref <= 1e-4
    assert err_mine <= max(2e-5, err_ref)


def test_error_maps_match_reference(setup, oracle):
    """calculateErrors of the reference build (src/functions.cu:4966-5040) on the same image and
    residuals. sigma(I_nu0) is image-sized fp32 arithmetic (rel 2e-5). In alpha_Noise the reference
    rounds x, y and the products x*u, y*v to FLOAT (:4133-4160; ~1e-5 turns at |x u| ~ 100), and sums
    Z terms sequentially in fp32, so it is compared at the level its own rounding allows and both
    are placed against the fp64 oracle (which can mimic those float roundings)."""
    p, e, ref, torch = setup
    if not hasattr(ref.lib, "gvref_error_image"):
        pytest.skip("oracle/_ref/libgvref.so predates gvref_error_image")
    I = _image(e)
    ref.set_image(I)
    ref.calc_function(iteration=0)
    want = ref.error_image().reshape(2, -1)
    I_dev = torch.from_numpy(I).cuda()
    e.chi2(I_dev)
    err = torch.empty_like(I_dev)
    e.error_maps(I_dev, err)
    got = err.cpu().numpy().reshape(2, -1)
    assert np.array_equal(got[0] == 0, want[0] == 0)
    nz = want[0] > 0
    np.testing.assert_allclose(got[0][nz], want[0][nz], rtol=2e-5)
    both = (got[1] > 0) & (want[1] > 0)
    assert both.sum() > 0.5 * max((want[1] > 0).sum(), 1)
    assert np.count_nonzero((got[1] > 0) != (want[1] > 0)) <= 0.02 * both.sum()
    rel = np.abs(got[1][both] - want[1][both]) / want[1][both]
    from test_parity_gpu import _cfg, _error_blocks
    pix = np.flatnonzero(both)[::53]
    Ic = I_dev.cpu().numpy()
    t0, t1 = oracle.error_maps(pix, p.N, _error_blocks(p, e), e.get_noise_image(), Ic, e.meta, _cfg(p))
    ok = t1 > 0
    r_ref = np.abs(want[1][pix][ok] - t1[ok]) / t1[ok]
    r_me = np.abs(got[1][pix][ok] - t1[ok]) / t1[ok]
    print(f"\n[error maps] sigma(alpha) median rel. distance: engine-reference {np.median(rel):.2e}, "
          f"reference-fp64 {np.median(r_ref):.2e}, engine-fp64 {np.median(r_me):.2e}")
    assert np.median(rel) <= 1e-3
    assert np.median(r_me) <= max(2e-5, np.median(r_ref))


def test_conv_degridding_matches_reference_kernel(setup, oracle):
    """Forward-model option gvm_set_degrid_kernel against the reference's degriddingGPU kernel
    (src/functions.cu:2205-2254, launched by the harness on the engine's own model grid, centred
    with fftshift) and the C oracle. The reference reads the left half of the grid through the
    Hermitian twin, the engine reads the cell itself: equal up to the fp32 asymmetry of the FFT."""
    p, e, ref, torch = setup
    if not hasattr(ref.lib, "gvref_degridding"):
        pytest.skip("oracle/_ref/libgvref.so predates gvref_degridding")
    from gpuvmem_b200 import host
    du, dv = e.meta["deltau"], e.meta["deltav"]
    for name, m, n in (("Gaussian2D", 7, 7), ("PSWF", 9, 9), ("PillBox2D", 1, 1)):
        table, (sx, sy) = host.ckernel_table(name, m, n, np.float32(abs(du)), np.float32(abs(dv)))
        e.set_degrid_kernel(table, (sx, sy))
        try:
            I_dev = torch.from_numpy(_image(e)).cuda()
            chi2 = e.chi2(I_dev)
            c = p.nchan - 1                                  # the model grid left behind is the last channel's
            V = e.get_model_grid()
            g = e.get_vis(c, want=("uvw", "Vo", "Vm", "Vr", "w"))
            Vc = np.fft.fftshift(V)
            want_ref = ref.degridding(g["uvw"], Vc, table, du, dv, sx, sy)
            want_orc = oracle.degrid_conv(g["uvw"], Vc, table, du, dv, sx, sy)
            N = p.N
            j = (g["uvw"][:, 0] / du + N // 2 + 0.5).astype(np.int64)
            k = (g["uvw"][:, 1] / dv + N // 2 + 0.5).astype(np.int64)
            inner = (j - sx > 0) & (j + sx < N) & (k - sy > 0) & (k + sy < N)   # no tap on row/column 0 (reference OOB)
            assert inner.sum() > 0.9 * len(j)
            scale = np.abs(want_ref[inner]).max()
            assert np.abs(want_orc[inner] - want_ref[inner]).max() <= 1e-6 * scale, name
            assert np.abs(g["Vm"][inner] - want_ref[inner]).max() <= 2e-5 * scale, name
            assert np.allclose(g["Vr"], g["Vo"] - g["Vm"], atol=1e-6 * scale)
            half = 0.5 * float(np.sum(g["w"].astype(np.float64) * (g["Vr"].astype(np.float64) ** 2).sum(1)))
            tot = 0.0
            for cc in range(p.nchan):
                gg = e.get_vis(cc, want=("Vr", "w"))
                tot += 0.5 * float(np.sum(gg["w"].astype(np.float64) * (gg["Vr"].astype(np.float64) ** 2).sum(1)))
            assert abs(chi2 - tot) <= 1e-5 * tot
        finally:
            e.set_degrid_kernel(None)
    # PillBox 1x1 = nearest-cell sampling of the grid
    assert sx == 0 and sy == 0
    # and the bilinear model is back
    I_dev = torch.from_numpy(_image(e)).cuda()
    ref.set_image(_image(e))
    want, _ = ref.calc_function(iteration=0)
    assert abs(e.chi2(I_dev) - want) <= 1e-5 * abs(want)


# Every Fi kind the reference registers (factory names, src/*.cu registerCreationFunction), evaluated ON ITS OWN
# by the reference build (gvref_prior_eval: Fi::configure + calcFi + restartDGi + calcGi + addToDphi) — TV,
# Quadratic, GEntropy and GL1Norm are not wired by main.cu, so the objective-level tests above never reach them.
PRIOR_CASES = [("Entropy", 0), ("L1-Norm", 0), ("TotalVariation", 0), ("TotalSquaredVariation", 0), ("Laplacian", 0),
               ("Quadratic", 0), ("GEntropy", 0), ("GL1Norm", 0), ("TotalVariation", 1), ("Quadratic", 1), ("L1-Norm", 1)]
LAM, EPS_B = 0.37, 1e-3


def _prior_case(e, kind, index):
    I = _image(e)
    if index == 1:
        I[1] = np.abs(I[1]) + np.float32(0.01)      # a positive plane (the terms take logs / square roots of it)
    prior_img = (np.abs(I[index]) * 0.5 + 1e-4).astype(np.float32) if kind in ("GEntropy", "GL1Norm") else None
    eps_a = 1e-12 if kind in ("L1-Norm", "GL1Norm") else 1e-6
    return I, prior_img, eps_a


def _close(who, g, want_g, kind, index):
    rel = np.linalg.norm(g - want_g) / np.linalg.norm(want_g)
    worst = float(np.abs(g - want_g).max()) / float(np.abs(want_g).max())
    assert rel <= 1e-5 and worst <= 1e-4, (who, kind, index, rel, worst)
    assert np.count_nonzero((g == 0) != (want_g == 0)) <= 2, (who, kind, index)


@pytest.mark.parametrize("kind,index", PRIOR_CASES)
def test_every_prior_kind_matches_reference(setup, oracle, kind, index):
    """Kernel level: gvm_prior_value / gvm_prior_grad and the C oracle's gvo_prior_* against the reference's Fi."""
    p, e, ref, torch = setup
    if not hasattr(ref.lib, "gvref_prior_eval"):
        pytest.skip("oracle/_ref/libgvref.so predates gvref_prior_eval")
    from gpuvmem_b200.engine import PRIOR
    I, prior_img, eps_a = _prior_case(e, kind, index)
    want_v, want_dphi, prior_after = ref.prior_eval(kind, I, LAM, image_index=index, iteration=1, prior_image=prior_img,
                                                    prior_value=0.001, eta=-1.0, eps_a=eps_a, eps_b=EPS_B)
    # TVariation::addToDphi always adds to image 0 (src/totalvariation.cu:46); every other Fi uses imageToAdd
    target = 0 if kind == "TotalVariation" else index
    assert not want_dphi[1 - target].any(), "the reference adds the gradient to one image only"
    want_g = want_dphi[target]
    grad_prior = prior_img
    if kind == "GL1Norm":
        # reference quirk: GL1Norm::calcGi swaps DGL1Norm's image arguments (src/gl1norm.cu:145-148 vs
        # src/functions.cu:4700): the gradient is computed against a ZERO prior (the freshly reset device_DS) and lands
        # in the prior image; dphi receives nothing
        assert not want_dphi.any()
        assert not np.array_equal(prior_after, prior_img)
        want_g, grad_prior = prior_after, np.zeros_like(prior_img)
    else:
        assert np.abs(want_g).max() > 0
        if prior_img is not None:
            assert np.array_equal(prior_after, prior_img)
    assert np.isfinite(want_v)
    noise = e.get_noise_image()
    k = PRIOR[kind]
    okw = dict(G=0.001, eta=-1.0, eps=eps_a, eps_b=EPS_B)
    ov = oracle.prior_value(k, I[index], noise, e.meta["noise_cut"], prior_image=prior_img, **okw)
    og = oracle.prior_grad(k, I[index], noise, e.meta["noise_cut"], LAM, prior_image=grad_prior, **okw)
    I_dev = torch.from_numpy(I).cuda()
    kw = dict(prior_value=0.001, eta=-1.0, epsilon=eps_a, epsilon_b=EPS_B)
    vkw, gkw = dict(kw), dict(kw)
    if prior_img is not None:
        vkw["prior_image"] = torch.from_numpy(prior_img).cuda()
        gkw["prior_image"] = torch.from_numpy(grad_prior).cuda()
    gv = e.prior_value(kind, I_dev, index, **vkw)
    dgi = torch.empty(p.N, p.N, device="cuda")
    e.prior_grad(kind, I_dev, dgi, LAM, index, **gkw)
    gg = dgi.cpu().numpy()
    for who, v, g in (("oracle", ov, og), ("engine", gv, gg)):
        assert abs(v - want_v) <= 2e-5 * abs(want_v), (who, kind, index, v, want_v)
        _close(who, g, want_g, kind, index)


@pytest.fixture(scope="module")
def session(setup):
    from gpuvmem_b200 import host
    p = setup[0]
    host.set_quiet(True)
    s = host.Session(p, args=ARGS.replace("-X 16 -Y 16 -V 256 ", "").replace(" -i synth.ms -o out.ms -m hdr.fits", ""))
    yield s
    s.close()


@pytest.mark.parametrize("kind,index", PRIOR_CASES)
def test_host_fi_terms_match_reference(setup, session, kind, index):
    """Class level: the host layer's Fi adapters (calcFi, restartDGi, calcGi, addToDphi incl. the image each term adds
    to, the flag_opt gate and GL1Norm's swapped-argument quirk) against the reference's Fi objects."""
    p, e, ref, torch = setup
    if not hasattr(ref.lib, "gvref_prior_eval"):
        pytest.skip("oracle/_ref/libgvref.so predates gvref_prior_eval")
    I, prior_img, eps_a = _prior_case(e, kind, index)
    kw = dict(image_index=index, iteration=1, prior_image=prior_img, prior_value=0.001, eta=-1.0, eps_a=eps_a, eps_b=EPS_B)
    want_v, want_dphi, want_after = ref.prior_eval(kind, I, LAM, **kw)
    got_v, got_dphi, got_after = session.fi_eval(kind, I, LAM, **kw)
    assert abs(got_v - want_v) <= 2e-5 * abs(want_v), (kind, index, got_v, want_v)
    for img in range(2):
        if want_dphi[img].any():
            _close("host", got_dphi[img], want_dphi[img], kind, index)
        else:
            assert not got_dphi[img].any(), (kind, index, img)
    if prior_img is not None:
        if np.array_equal(want_after, prior_img):
            assert np.array_equal(got_after, prior_img)
        else:
            _close("host prior-after", got_after, want_after, kind, index)
    # gate closed: flag_opt % 2 != imageIndex -> no gradient at all; iteration 0 -> value 0 too
    _, dphi_closed, _ = session.fi_eval(kind, I, LAM, **dict(kw, flag=1 - index))
    _, ref_closed, _ = ref.prior_eval(kind, I, LAM, **dict(kw, flag=1 - index))
    assert not dphi_closed.any() and not ref_closed.any()
    v0, dphi0, _ = session.fi_eval(kind, I, LAM, **dict(kw, iteration=0))
    r0, rdphi0, _ = ref.prior_eval(kind, I, LAM, **dict(kw, iteration=0))
    assert v0 == 0.0 and r0 == 0.0 and not dphi0.any() and not rdphi0.any()
