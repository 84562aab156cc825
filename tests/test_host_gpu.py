"""GPU parity of the C++ host layer against the REFERENCE ITSELF, scenario by scenario:
MFS::configure/setDevice scalars, the weighted (and gridded) visibilities (bit-exact),
ObjectiveFunction::calcFunction / calcGradient at a probe image, and the image after the
optimizer has run N iterations (north-star tolerance: chi2 rel 1e-5, gradient rel-L2 1e-4;
final image: stated per scenario below, it inherits the amplification of the line search).

The reference runs in its own process per scenario (tests/_ref_runner.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from _ref_runner import probe_image
from _scenarios import SCENARIOS, problem

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# rel-L2 tolerance on the final image (unmasked pixels) after the scenario's iterations
FINAL_TOL = {"cg_natural": 2e-3, "lbfgs_natural": 2e-3, "cg_mfs_briggs": 2e-3, "cg_gridded_gaussian": 2e-3,
             "cg_gridded_pswf": 2e-3, "cg_nopositivity_eta": 2e-3, "lbfgs_mfs_threshold_radial": 2e-3,
             "cg_offset_field": 2e-3}


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def refdir(tmp_path_factory):
    from _checkers import GVREF_SO
    if not os.path.exists(GVREF_SO):
        pytest.skip("oracle/_ref/libgvref.so not built")
    return tmp_path_factory.mktemp("ref")


def _reference(name, refdir):
    out = os.path.join(str(refdir), name + ".npz")
    if not os.path.exists(out):
        # one OpenMP thread: the reference's Briggs/uniform grids are fp32 sums whose order follows
        # the thread interleaving (omp critical); only the 1-thread order is deterministic
        env = dict(os.environ, OMP_THREAD_LIMIT="1", OMP_NUM_THREADS="1")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_ref_runner.py"), name, out],
                           capture_output=True, text=True, timeout=900, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return np.load(out)


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_scenario_matches_reference(name, refdir):
    from gpuvmem_b200 import host
    ref = _reference(name, refdir)
    kw, args, optimizer, scheme, ck, ck_size, K = SCENARIOS[name]
    p = problem(name)
    s = host.Session(p, args=args, optimizer=optimizer, scheme=scheme, ckernel=ck, ck_size=ck_size)
    try:
        if K:
            s.set_lbfgs_k(K)
        sc = s.scalars()
        # -- MFS::configure / setDevice ----------------------------------------------------
        assert sc["deltau"] == ref["s_deltau"] and sc["deltav"] == ref["s_deltav"]
        assert np.float32(sc["xobs_pix"]) == np.float32(ref["s_xpix"]) and np.float32(sc["yobs_pix"]) == np.float32(ref["s_ypix"])
        assert np.float32(sc["nu_0"]) == np.float32(ref["s_nu_0"])
        if name == "cg_offset_field":
            assert abs(sc["xobs_pix"] - p.N / 2) > 3 and abs(sc["yobs_pix"] - p.N / 2) > 3, "the field must be off-centre"
        for mine, theirs, tol in (("vis_noise", "s_vis_noise", 1e-6), ("noise_jypix", "s_noise_jypix", 1e-5),
                                  ("fg_scale", "s_fg_scale", 1e-5), ("noise_cut", "s_noise_cut", 1e-5)):
            assert abs(sc[mine] - float(ref[theirs])) <= tol * abs(float(ref[theirs])), (mine, sc[mine], float(ref[theirs]))
        # -- weights (+ gridding): bit-exact -----------------------------------------------
        for c in range(p.nchan):
            uvw, Vo, w = s.host_vis(c)
            assert len(w) == len(ref[f"w{c}"]), (c, len(w), len(ref[f"w{c}"]))
            assert np.array_equal(w.view(np.uint32), ref[f"w{c}"].view(np.uint32)), f"weights, channel {c}"
            assert np.array_equal(uvw.view(np.uint64), ref[f"uvw{c}"].view(np.uint64)), f"uvw, channel {c}"
            assert np.array_equal(Vo.view(np.uint32), ref[f"Vo{c}"].view(np.uint32)), f"Vo, channel {c}"
        assert np.array_equal(s.get_image(), ref["I_start"])
        # -- objective + gradient at a probe image -----------------------------------------
        z = [float(t) for t in args.split("-z")[1].split()[0].split(",")]
        probe = probe_image(p.N, np.float32(z[0]), z[1] if len(z) > 1 else 0.0)
        s.set_image(probe)
        s.set_iteration(1)
        v, fi = s.calc_function()
        rfi = ref["probe_fi"][:len(fi)]
        assert abs(fi[0] - rfi[0]) <= 1e-5 * abs(rfi[0]), ("chi2", fi[0], rfi[0])
        assert np.allclose(fi[1:], rfi[1:], rtol=2e-5, atol=0), (fi, rfi)
        assert abs(v - float(ref["probe_value"])) <= 1e-5 * abs(float(ref["probe_value"]))
        g = s.calc_gradient(1)
        assert np.array_equal(s.get_image(), ref["probe_image_after"]), "clip2IWNoise side effect"
        assert _rel(g[0], ref["probe_grad"][0]) <= 1e-4, _rel(g[0], ref["probe_grad"][0])
        assert np.array_equal(g[0] == 0, ref["probe_grad"][0] == 0), "masked pixels must be exactly 0"
        if "probe_grad_flag1" in ref.files:
            s.set_flag(1)
            g1 = s.calc_gradient(1)
            s.set_flag(0)
            r1 = ref["probe_grad_flag1"]
            assert _rel(g1[1], r1[1]) <= 1e-4, _rel(g1[1], r1[1])
            assert np.array_equal(g1[1] == 0, r1[1] == 0), "alpha gradient: masked / below-threshold pixels exactly 0"
            assert np.array_equal(g1[0] == 0, r1[0] == 0)
        # -- Error "SecondDerivateError" (calculateErrors) on the same residuals ------------
        if "probe_err" in ref.files:
            err, rerr = s.error_image(), ref["probe_err"]
            assert np.array_equal(err[0] == 0, rerr[0] == 0)
            assert _rel(err[0], rerr[0]) <= 2e-5, _rel(err[0], rerr[0])
            both = (err[1] > 0) & (rerr[1] > 0)
            assert np.count_nonzero((err[1] > 0) != (rerr[1] > 0)) <= 0.02 * max(both.sum(), 50)
            if both.any():   # single-channel scenarios have ln(nu/nu0) = 0: sigma(alpha) is 0 everywhere
                med = float(np.median(np.abs(err[1][both] - rerr[1][both]) / rerr[1][both]))
                print(f"\n[{name}] error maps: sigma(I) rel-L2 {_rel(err[0], rerr[0]):.2e}, sigma(alpha) median rel {med:.2e}")
                assert med <= 1e-3, med
        # -- the optimizer: image after N iterations ---------------------------------------
        s.set_image(ref["I_start"])
        s.set_iteration(0)
        img, seconds = s.run()
        it = int(s.scalars()["iterations_done"])
        assert it == int(ref["iterations"]), (it, int(ref["iterations"]), s.exit_reason())
        err0 = _rel(img[0], ref["final_image"][0])
        print(f"\n[{name}] iterations={it} final-image rel-L2={err0:.3e} reference {float(ref['run_ms']):.0f} ms, "
              f"here {seconds * 1e3:.0f} ms, exit={s.exit_reason()}")
        assert err0 <= FINAL_TOL[name], err0
        if len(z) > 1:
            m = ref["final_image"][1] != np.float32(z[1])
            assert _rel(img[1][m], ref["final_image"][1][m]) <= 10 * FINAL_TOL[name]
        v2, fi2 = s.calc_function()
        assert abs(v2 - float(ref["final_value"])) <= 1e-3 * abs(float(ref["final_value"]))
        # -- residual / model write-back (MFS::writeResiduals + modelToHost) on the REFERENCE's final image ----
        if "wb_chi2" in ref.files:
            s.set_image(ref["final_image"])
            s.calc_function()
            wb_chi2, blocks = s.write_residuals()
            gridded = "-g" in args
            assert len(blocks) == p.nchan
            for c, b in enumerate(blocks):
                # the samples that go to the file are the ORIGINAL ones (ungridded again after a gridded run)
                assert np.array_equal(b["uvw"].view(np.uint64), ref[f"wb_uvw{c}"].view(np.uint64)), f"write-back uvw {c}"
                assert np.array_equal(b["uvw"], np.ascontiguousarray(p.uvw[c], np.float64))
                assert np.array_equal(b["Vo"].view(np.uint32), ref[f"wb_Vo{c}"].view(np.uint32)), f"write-back Vo {c}"
                assert np.array_equal(b["w"].view(np.uint32), ref[f"wb_w{c}"].view(np.uint32)), f"write-back weights {c}"
                rVm = ref[f"wb_Vm{c}"]
                scale = float(np.abs(rVm).max())
                assert scale > 0
                assert np.abs(b["Vm"] - rVm).max() <= 2e-5 * scale, (c, np.abs(b["Vm"] - rVm).max() / scale)
                assert np.abs(b["Vr"] - (ref[f"wb_Vo{c}"] - rVm)).max() <= 2e-5 * max(scale, float(np.abs(b["Vo"]).max()))
            if gridded:
                want = float(ref["wb_chi2"])
                assert abs(wb_chi2 - want) <= 1e-5 * abs(want), ("non-gridded chi2", wb_chi2, want)
                print(f"\n[{name}] write-back: non-gridded 0.5*chi2 {wb_chi2:.6e} (reference {want:.6e})")
    finally:
        s.close()
