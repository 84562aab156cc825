"""The shipped command-line program (gpuvmem_b200/bin/gpuvmem = src/main.cu:100-229 on the new classes)
end to end on a GPU: GVMS container in, image + alpha + error maps (-E -P) + residual file out, and the
result must equal the same reconstruction driven through the C entry points of the host layer."""
import json
import os
import subprocess

import numpy as np
import pytest

from gpuvmem_b200 import host, synth

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
    img = str(tmp_path / "image.f32")
    args = ["-i", gv, "-m", gv, "-o", str(tmp_path / "out.gvmr"), "-O", img, "-p", mem, "-z", "0.001,0.2",
            "-Z", "0.01,0.0,0.001", "-t", "4", "-E", "-P"]
    r = subprocess.run([BIN] + args, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, GVM_OPTIMIZER="CG-FRPRMN"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "Calculating Error Images" in r.stdout
    N = p.N
    out0 = np.fromfile(img, np.float32).reshape(N, N)
    side = json.load(open(img + ".json"))
    assert side["shape"] == [N, N] and side["bunit"] == "JY/PIXEL"
    alpha = np.fromfile(mem + "alpha.fits", np.float32).reshape(N, N)
    e0 = np.fromfile(mem + "error_Inu_0.fits", np.float32).reshape(N, N)
    e1 = np.fromfile(mem + "error_alpha_0.fits", np.float32).reshape(N, N)
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
    path = str(tmp_path / "mask.f32")
    mask.tofile(path)
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
