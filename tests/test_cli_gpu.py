"""The shipped command-line program (gpuvmem_b200/bin/gpuvmem = src/main.cu:100-229 on the new classes)
end to end on a GPU: GVMS container in, image + alpha + error maps (-E -P) + residual file out, and the
result must equal the same reconstruction driven through the C entry points of the host layer."""
import json
import os
import subprocess

import numpy as np
import pytest

from gpuvmem_b200 import fits, host, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "gpuvmem_b200", "bin", "gpuvmem")


def test_command_line_reconstruction_with_error_maps(tmp_path):
    assert os.path.exists(BIN), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    p = synth.make_problem(N=128, nvis=12000, nchan=3, freq0=1.0e11, bandwidth=6e9, seed=77, grid_fill=0.9)
    gv = str(tmp_path / "in.gvms")
    synth.write_gvms(p, gv)
    mem = str(tmp_path / "mem") + "/"
    os.makedirs(mem)
    img = str(tmp_path / "image.fits")
    # -m: a FITS model image whose header carries the astrometry (readFITSHeader, src/MSFITSIO.cu:262-325)
    model = str(tmp_path / "mod_in.fits")
    fits.write_fits(model, np.zeros((p.N, p.M), np.float32),
                    {"CTYPE1": "RA---SIN", "CRVAL1": p.ra, "CDELT1": p.DELTAX, "CRPIX1": p.crpix1, "CTYPE2": "DEC--SIN",
                     "CRVAL2": p.dec, "CDELT2": p.DELTAY, "CRPIX2": p.crpix2, "TELESCOP": p.telescope, "OBJECT": "synthetic"})
    args = ["-i", gv, "-m", model, "-o", str(tmp_path / "out.gvmr"), "-O", img, "-p", mem, "-z", "0.001,0.2",
            "-Z", "0.01,0.0,0.001", "-t", "4", "-E", "-P"]
    r = subprocess.run([BIN] + args, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, GVM_OPTIMIZER="CG-FRPRMN"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "Calculating Error Images" in r.stdout
    N = p.N
    hdr, out0 = fits.read_fits(img)                      # written as FITS: the model's header copied (OCopyFITS)
    out0 = out0.astype(np.float32)
    assert hdr["NAXIS1"] == N and hdr["NAXIS2"] == N and hdr["BUNIT"] == "JY/PIXEL" and hdr["NITER"] == 4
    assert hdr["CDELT1"] == p.DELTAX and hdr["CRPIX2"] == p.crpix2 and hdr["OBJECT"] == "synthetic"
    assert abs(hdr["CRVAL1"] - p.ra) < 1e-9 and abs(hdr["CRVAL2"] - p.dec) < 1e-9
    alpha = fits.read_fits(mem + "alpha.fits")[1].astype(np.float32)
    e0 = fits.read_fits(mem + "error_Inu_0.fits")[1].astype(np.float32)
    e1 = fits.read_fits(mem + "error_alpha_0.fits")[1].astype(np.float32)
    assert os.path.getsize(str(tmp_path / "out.gvmr")) > 8 + 12 + p.total_vis() * 20

    # the same run through the C entry points
    host.set_quiet(True)
    s = host.Session(p, args="-z 0.001,0.2 -Z 0.01,0.0,0.001 -t 4", optimizer="CG-FRPRMN",
                     fi_spec="Chi2:-1:0:0,Entropy:0:0:0,L1-Norm:1:0:0,TotalSquaredVariation:2:0:0,Laplacian:3:0:0")
    try:
        want, _ = s.run()
        sc = s.scalars()
        np.testing.assert_allclose(out0, want[0] * np.float32(sc["fg_scale"]), rtol=1e-6, atol=0)
        assert np.array_equal(alpha, want[1])
        err = s.error_image()
        assert np.array_equal(e0, err[0]) and np.array_equal(e1, err[1])
        assert (e0 > 0).any() and (e1 > 0).any()
    finally:
        s.close()


def _noise_of(s):
    out = np.empty((s.M, s.N), np.float32)
    assert s.eng.gvm_get_noise_image(s.engine_handle(), out.ctypes.data) == 0
    return out


def test_user_mask_and_radius_mask(tmp_path):
    """-U file: the plane replaces the noise image after fg_scale / noise_cut were derived, noise_cut = 1 x min
    noise (src/functions.cu:296-298, src/mfs.cu:916-927); -M: distance_image (src/functions.cu:2360-2380)."""
    p = synth.make_problem(N=128, nvis=8000, nchan=1, seed=78, grid_fill=0.9)
    N = p.N
    host.set_quiet(True)
    base = host.Session(p, args="-z 0.001 -Z 0.01 -t 3")
    sc0 = base.scalars()
    base.close()
    mask = np.full((N, N), 1e30, np.float32)
    mask[40:90, 30:100] = 0.0
    path = str(tmp_path / "mask.fits")
    fits.write_fits(path, mask)                          # -U takes a FITS plane (read_data_float_FITS)
    s = host.Session(p, args=f"-z 0.001 -Z 0.01 -t 3 -U {path}")
    try:
        sc = s.scalars()
        assert np.array_equal(_noise_of(s), mask)
        assert abs(sc["fg_scale"] - sc0["fg_scale"]) <= 1e-6 * sc0["fg_scale"]
        assert abs(sc["noise_cut"] - sc["fg_scale"]) <= 1e-6 * sc["fg_scale"]      # 1 x min(noise)
        s.set_iteration(1)
        s.calc_function()
        g = s.calc_gradient(1)
        assert not g[0][mask > 0].any()
        assert np.count_nonzero(g[0][mask == 0]) > 0.99 * np.count_nonzero(mask == 0)
        img = s.get_image()                       # clip2IWNoise: masked pixels pinned at -eta*MINPIX, alpha 0
        assert (img[0][mask > 0] == np.float32(0.001)).all() and not img[1][mask > 0].any()
    finally:
        s.close()
    s = host.Session(p, args="-z 0.001 -Z 0.01 -t 3 -M")
    try:
        d = _noise_of(s)
        x0, y0 = int(s.scalars()["xobs_pix"]), int(s.scalars()["yobs_pix"])
        want = np.ones((N, N), np.float32)
        want[y0, x0] = 0.0
        assert np.array_equal(d, want)
    finally:
        s.close()


def test_gridding_filter_equals_the_gridded_run():
    """Filter "Gridding" (src/gridding.cu) applied to the Visibilities of an ungridded session gives the very
    samples a `-g` session grids in MFS::configure (same kernel, natural weights): bit-identical."""
    p = synth.make_problem(N=128, nvis=15000, nchan=2, freq0=1.0e11, bandwidth=2e9, seed=91, grid_fill=0.9)
    host.set_quiet(True)
    a = host.Session(p, args="-z 0.001 -Z 0.01 -t 2 -g 1", ckernel="Gaussian2D", ck_size=(7, 7))
    try:
        want = [a.host_vis(c) for c in range(p.nchan)]
    finally:
        a.close()
    b = host.Session(p, args="-z 0.001 -Z 0.01 -t 2")
    try:
        before = [len(b.host_vis(c)[2]) for c in range(p.nchan)]
        b.filter_gridding("Gaussian2D", (7, 7))
        for c in range(p.nchan):
            uvw, Vo, w = b.host_vis(c)
            assert len(w) < before[c] and len(w) == len(want[c][2])
            assert np.array_equal(uvw.view(np.uint64), want[c][0].view(np.uint64))
            assert np.array_equal(Vo.view(np.uint32), want[c][1].view(np.uint32))
            assert np.array_equal(w.view(np.uint32), want[c][2].view(np.uint32))
    finally:
        b.close()


def test_mosaic_container_two_fields_and_an_unused_correlation(tmp_path):
    """A GVMS container with two fields (one off the image centre) and two correlations (XX, XY) through the
    file-reading path of the host layer (MFS::configure -> Io::read), against an engine assembled by hand:
    only LL/RR/XX/YY blocks are used (src/functions.cu:4378-4381), every field keeps its own pointing / phase
    centre from direccos (src/mfs.cu:660-691), the noise image adds the beams of all fields (:850-916)."""
    import math
    import torch
    from gpuvmem_b200 import Engine
    from gpuvmem_b200.engine import RPDEG_D, beam_model, direccos
    kw = dict(N=128, nchan=2, freq0=1.0e11, bandwidth=2e9, grid_fill=0.9)
    p = synth.make_problem(nvis=8000, seed=201, **kw)
    q = synth.make_problem(nvis=6000, seed=202, **kw)
    q.field_centre = (p.ra + 9.2 * p.DELTAX / math.cos(math.radians(p.dec)), p.dec - 6.1 * p.DELTAY)
    path = str(tmp_path / "mosaic.gvms")
    synth.write_gvms(p, path, fields=[p, q], corr_types=(9, 10))     # XX, XY
    host.set_quiet(True)
    s = host.Session(None, args=f"-i {path} -m {path} -o {tmp_path / 'out.gvmr'} -z 0.001,0.1 -Z 0.01 -t 3",
                     shape=(p.M, p.N))
    try:
        sc = s.scalars()
        assert int(sc["total_visibilities"]) == p.total_vis() + q.total_vis(), "the XY blocks must not count"
        I = s.get_image()
        s.set_iteration(0)
        v, fi = s.calc_function()
        g = s.calc_gradient(0)
        pbf, pbc, pb = beam_model(p.telescope, p.antenna_diameter, float(p.freqs.min()))
        e = Engine(p.M, p.N, p.DELTAX, p.DELTAY, sc["nu_0"], eta=-1.0, minpix=0.001, noise_cut=1e30, threshold=0.0,
                   fg_scale=1.0)
        try:
            dx, dy = RPDEG_D * p.DELTAX, RPDEG_D * p.DELTAY
            for fld in (p, q):
                fc = getattr(fld, "field_centre", None) or (p.ra, p.dec)
                l, m = direccos(math.radians(fc[0]), math.radians(fc[1]), math.radians(p.ra), math.radians(p.dec))
                xp = float(np.float32(l / dx + float(np.float32(p.crpix1) - np.float32(1.0))))
                yp = float(np.float32(m / dy + float(np.float32(p.crpix2) - np.float32(1.0))))
                for c in range(p.nchan):
                    e.add_channel(float(p.freqs[c]), fld.uvw[c], fld.Vo[c], fld.w[c], p.antenna_diameter, pbf, pbc, pb,
                                  (xp, yp), (xp, yp))
            assert e.num_channels() == 4
            fg = e.build_noise_image(sc["noise_jypix"])
            assert abs(fg - sc["fg_scale"]) <= 1e-6 * fg
            e.set_scalars(fg, float(np.float32(10.0) * np.float32(fg)), 0.0)
            I_dev = torch.from_numpy(I.copy()).cuda()
            chi2 = e.chi2(I_dev)
            assert abs(chi2 - fi[0]) <= 1e-6 * abs(chi2), (chi2, fi[0])
            gd = torch.zeros_like(I_dev)
            e.dchi2(I_dev, gd, flag_opt=0)
            want = gd[0].cpu().numpy()
            assert np.linalg.norm(g[0] - want) <= 1e-6 * np.linalg.norm(want)
            assert abs(yp - p.N / 2) > 3, "the second field sits off the image centre"
        finally:
            e.close()
        img, _ = s.run()                      # and the whole reconstruction runs on the mosaic
        assert np.isfinite(img).all() and int(s.scalars()["iterations_done"]) == 3
    finally:
        s.close()


def test_single_synchronisation_objective_equals_the_per_term_loop(monkeypatch):
    """ObjectiveFunction::calcFunction's fast path (every Fi value launched into an engine slot, ONE stream
    synchronisation) against the reference's per-term loop (GVM_SINGLE_SYNC=0): identical values, identical
    reconstruction."""
    from _ref_runner import probe_image
    p = synth.make_problem(N=128, nvis=12000, nchan=2, freq0=1.0e11, bandwidth=4e9, seed=55, grid_fill=0.9)
    host.set_quiet(True)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("GVM_SINGLE_SYNC", mode)
        s = host.Session(p, args="-z 0.001,0.1 -Z 0.01,0.005,0.002,0.001 -t 4")
        try:
            start = s.get_image()
            s.set_image(probe_image(p.N, np.float32(0.001), 0.1))
            s.set_iteration(0)
            v0, fi0 = s.calc_function()            # priors gated off: their values are 0
            s.set_iteration(1)
            v1, fi1 = s.calc_function()
            clipped = s.get_image()
            s.set_image(start)
            s.set_iteration(0)
            img, _ = s.run()
            out[mode] = (v0, fi0, v1, fi1, clipped, img, s.stats()["function_evals"])
        finally:
            s.close()
    a, b = out["1"], out["0"]
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and not a[1][1:].any()
    assert a[2] == b[2] and np.array_equal(a[3], b[3]) and a[3][1:].all()
    assert np.array_equal(a[4], b[4]) and np.array_equal(a[5], b[5]) and a[6] == b[6]


@pytest.mark.parametrize("fi_spec", [None, "Chi2:-1:0:0,TotalVariation:0:0:0,Quadratic:1:0:0,L1-Norm:2:1:1",
                                     "Entropy:0:0:0,Chi2:-1:0:0"])
def test_fused_gradient_and_graph_replay_equal_the_reference_loop(monkeypatch, fi_spec):
    """ObjectiveFunction's fast paths against the reference's loops, bit for bit: (a) calcGradient with every term
    writing straight into xi (Fi::gradInto: chi2 accumulated in place, priors added by gvm_prior_grad_add) vs
    restartDGi / calcGi / addToDphi through per-term buffers (GVM_FUSED_GRADIENT=0); (b) calcFunction replaying a
    captured CUDA graph vs plain launches (GVM_GRAPHS=0). The third spec puts a prior BEFORE Chi2: Chi2's addToDphi
    overwrites dphi (src/chi2.cu:60-70), so the fused path must step aside and the result still be the loop's."""
    from _ref_runner import probe_image
    p = synth.make_problem(N=128, nvis=12000, nchan=2, freq0=1.0e11, bandwidth=4e9, seed=56, grid_fill=0.9)
    host.set_quiet(True)
    out = {}
    for mode in ("fast", "loop"):
        monkeypatch.setenv("GVM_FUSED_GRADIENT", "1" if mode == "fast" else "0")
        monkeypatch.setenv("GVM_GRAPHS", "1" if mode == "fast" else "0")
        s = host.Session(p, args="-z 0.001,0.1 -Z 0.01,0.005,0.002,0.001 -t 4", fi_spec=fi_spec)
        try:
            start = s.get_image()
            s.set_image(probe_image(p.N, np.float32(0.001), 0.1))
            s.set_iteration(1)
            vals = [s.calc_function() for _ in range(4)]       # 1st plain, 2nd captured, 3rd and 4th replayed
            g0 = s.calc_gradient(1)
            s.set_flag(1)
            g1 = s.calc_gradient(1)
            s.set_flag(0)
            s.set_image(start)
            s.set_iteration(0)
            img, _ = s.run()
            out[mode] = (vals, g0, g1, img)
        finally:
            s.close()
    a, b = out["fast"], out["loop"]
    for (va, fa), (vb, fb) in zip(a[0], b[0]):
        assert va == vb and np.array_equal(fa, fb)
    assert a[0][0][0] == a[0][3][0], "replaying the graph must reproduce the plain evaluation"
    assert np.array_equal(a[1], b[1]) and a[1][0].any()
    assert np.array_equal(a[2], b[2])
    assert np.array_equal(a[3], b[3])
