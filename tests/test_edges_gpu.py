"""GPU edge cases and full-size properties of the objective/gradient path (through the C ABI).

Small sizes: ragged and degenerate visibility counts (1, KV-1, KV+1, one TMEM chunk + 1), an empty
block between two populated ones, all weights zero, every pixel masked, image sizes below and off
the 256-row tile — each against the fp64 oracle. BASELINE.json's full configuration (configs[1]:
2048^2 x 10 M) through size-independent properties: Vm + Vr = Vo, chi2 = 1/2 sum w |Vr|^2 recomputed
in fp64, determinism, accumulation, masked pixels exactly zero, and the tensor-core gradient
against the fp64 oracle at sampled pixels over ALL 10 M visibilities."""
import os
import subprocess
import sys

import numpy as np
import pytest

from gpuvmem_b200 import Engine, synth
from gpuvmem_b200.engine import GRAD_SIMT, GRAD_SIMT_EXACT, GRAD_UMMA

from test_parity_gpu import _cfg, _grad_oracle_sample, _test_image, _torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("nvis", [3, 31, 33, 2049, 4097])   # 1-2 samples: the beam fit of calculateNoiseAndBeam is singular
@pytest.mark.parametrize("mode", [GRAD_SIMT, GRAD_SIMT_EXACT, GRAD_UMMA])
def test_ragged_visibility_counts(oracle, nvis, mode):
    torch = _torch()
    p = synth.make_problem(N=64, nvis=nvis, nchan=1, seed=100 + nvis)
    e = Engine.from_problem(p, grad_mode=mode, keep_vm=True)
    try:
        I_dev = torch.from_numpy(_test_image(e)).cuda()
        chi2 = e.chi2(I_dev)
        v = e.get_vis(0, want=("Vo", "Vm", "Vr", "w"))
        assert np.array_equal(v["Vr"], v["Vo"] - v["Vm"])
        want = 0.5 * float(np.sum(v["w"].astype(np.float64) * (v["Vr"].astype(np.float64) ** 2).sum(1)))
        assert abs(chi2 - want) <= 1e-6 * max(want, 1e-30)
        g = torch.zeros_like(I_dev)
        e.dchi2(I_dev, g, flag_opt=0)
        assert e.last_grad_mode() == mode
        pix = np.arange(0, p.N * p.N, 7)
        truth = _grad_oracle_sample(oracle, p, e, I_dev.cpu().numpy(), pix, 0)
        got = g[0].cpu().numpy().reshape(-1)[pix]
        assert _rel(got, truth) <= 3e-5, (nvis, mode, _rel(got, truth))
    finally:
        e.close()


def test_empty_block_between_populated_blocks(oracle):
    torch = _torch()
    p = synth.make_problem(N=64, nvis=3000, nchan=3, freq0=2.3e11, bandwidth=4e9, seed=5)
    p.uvw[1], p.Vo[1], p.w[1] = p.uvw[1][:0], p.Vo[1][:0], p.w[1][:0]      # channel 1 has no samples
    e = Engine.from_problem(p, grad_mode=GRAD_UMMA)
    try:
        assert e.lib.gvm_channel_nvis(e.h, 1) == 0
        I_dev = torch.from_numpy(_test_image(e)).cuda()
        chi2 = e.chi2(I_dev)
        g = torch.zeros_like(I_dev)
        e.dchi2(I_dev, g, flag_opt=0)
        q = synth.make_problem(N=64, nvis=3000, nchan=3, freq0=2.3e11, bandwidth=4e9, seed=5)
        pix = np.arange(0, p.N * p.N, 11)
        tot = np.zeros(len(pix))
        half = 0.0
        for c in (0, 2):
            v = e.get_vis(c, want=("uvw", "Vr", "w"))
            half += 0.5 * float(np.sum(v["w"].astype(np.float64) * (v["Vr"].astype(np.float64) ** 2).sum(1)))
            d = oracle.dchi2(pix, p.N, v["uvw"], v["Vr"], v["w"], e.get_noise_image(), None, float(q.freqs[c]), e.meta, _cfg(p))
            tot += d * oracle.chain(I_dev.cpu().numpy(), pix, float(q.freqs[c]), e.meta, e.cfg.threshold, 0)
        assert abs(chi2 - half) <= 1e-5 * half
        assert _rel(g[0].cpu().numpy().reshape(-1)[pix], tot) <= 3e-5
        err = torch.empty_like(I_dev)
        e.error_maps(I_dev, err)                     # the empty block contributes nothing
        assert torch.isfinite(err).all()
    finally:
        e.close()


@pytest.mark.parametrize("mode", [GRAD_SIMT, GRAD_UMMA])
def test_zero_weights_and_fully_masked_image(mode):
    torch = _torch()
    p = synth.make_problem(N=64, nvis=2000, nchan=1, seed=8)
    e = Engine.from_problem(p, grad_mode=mode)
    try:
        I_dev = torch.from_numpy(_test_image(e)).cuda()
        # every pixel masked: DChi2 returns early everywhere (src/functions.cu:3723-3726)
        e.set_scalars(e.meta["fg_scale"], 0.0, e.cfg.threshold)
        e.chi2(I_dev)
        g = torch.full_like(I_dev, 3.0)
        e.dchi2(I_dev, g, flag_opt=0)
        torch.cuda.synchronize()
        assert (g == 3.0).all(), "nothing is added for masked pixels"
        err = torch.full_like(I_dev, 5.0)
        e.error_maps(I_dev, err)
        assert (err == 0).all()
    finally:
        e.close()
    # zero-weight samples (flagged data) contribute nothing: identical to dropping them
    p.w[0][::2] = 0.0
    e = Engine.from_problem(p, grad_mode=mode)
    try:
        I_dev = torch.from_numpy(_test_image(e)).cuda()
        chi2 = e.chi2(I_dev)
        g = torch.zeros_like(I_dev)
        e.dchi2(I_dev, g, flag_opt=0)
        v = e.get_vis(0, want=("Vr", "w"))
        assert (v["w"][::2] == 0).all()
        keep = v["w"] > 0
        half = 0.5 * float(np.sum(v["w"][keep].astype(np.float64) * (v["Vr"][keep].astype(np.float64) ** 2).sum(1)))
        assert abs(chi2 - half) <= 1e-6 * half
        q = synth.make_problem(N=64, nvis=2000, nchan=1, seed=8)
        q.uvw[0], q.Vo[0], q.w[0] = q.uvw[0][1::2], q.Vo[0][1::2], q.w[0][1::2]
        e2 = Engine.from_problem(q, grad_mode=mode)
        try:
            # same scalars as the full problem, so that only the sample set differs
            e2.set_noise_image(e.get_noise_image())
            e2.set_scalars(e.meta["fg_scale"], e.meta["noise_cut"], e.cfg.threshold)
            J_dev = torch.from_numpy(_test_image(e)).cuda()
            chi2b = e2.chi2(J_dev)
            g2 = torch.zeros_like(J_dev)
            e2.dchi2(J_dev, g2, flag_opt=0)
            assert abs(chi2b - chi2) <= 2e-6 * chi2
            assert _rel(g2[0].cpu().numpy(), g[0].cpu().numpy().astype(np.float64)) <= 2e-5
        finally:
            e2.close()
    finally:
        e.close()


@pytest.mark.parametrize("N", [32, 96, 320])
def test_image_sizes_off_the_tile(oracle, N):
    """N below one 256-row band, and N = 320 (one full band + a ragged one)."""
    torch = _torch()
    p = synth.make_problem(N=N, nvis=5000, nchan=1, seed=30 + N)
    e = Engine.from_problem(p, grad_mode=GRAD_UMMA)
    try:
        I_dev = torch.from_numpy(_test_image(e)).cuda()
        e.chi2(I_dev)
        g = torch.zeros_like(I_dev)
        e.dchi2(I_dev, g, flag_opt=0)
        assert e.last_grad_mode() == GRAD_UMMA
        e.set_grad_mode(GRAD_SIMT)
        g2 = torch.zeros_like(I_dev)
        e.dchi2(I_dev, g2, flag_opt=0)
        a, b = g[0].cpu().numpy(), g2[0].cpu().numpy()
        assert np.array_equal(a == 0, b == 0)
        assert _rel(a, b.astype(np.float64)) <= 4e-5
        pix = np.arange(0, N * N, 13)
        truth = _grad_oracle_sample(oracle, p, e, I_dev.cpu().numpy(), pix, 0)
        assert _rel(a.reshape(-1)[pix], truth) <= 3e-5
    finally:
        e.close()


def test_full_size_c2_properties(oracle):
    """BASELINE.json configs[1] at full size: 2048 x 2048 image, 10 M visibilities."""
    torch = _torch()
    p = synth.config_c2()
    assert p.N == 2048 and p.total_vis() == 10_000_000
    e = Engine.from_problem(p, keep_vm=True)
    try:
        I = _test_image(e)
        I_dev = torch.from_numpy(I).cuda()
        chi2 = e.chi2(I_dev)
        v = e.get_vis(0, want=("uvw", "Vo", "Vm", "Vr", "w"))
        assert np.array_equal(v["Vr"], v["Vo"] - v["Vm"])
        half = 0.5 * float(np.sum(v["w"].astype(np.float64) * (v["Vr"].astype(np.float64) ** 2).sum(1)))
        assert abs(chi2 - half) <= 1e-6 * half, (chi2, half)
        assert e.chi2(I_dev) == chi2, "the forward pass is deterministic"
        # the forward model itself at the benchmarked size: fp64 oracle (clip is already applied to I_dev; FFT of
        # the 2048^2 model image, bilinear degridding of all 10 M samples) vs the engine's Vm, Vr and 0.5*chi2
        Ic = I_dev.cpu().numpy()
        prep = oracle.prep(p.uvw[0], p.Vo[0], p.w[0], float(p.freqs[0]), e.meta["deltau"], e.meta["deltav"], p.N)
        assert np.array_equal(prep["uvw"].view(np.uint64), v["uvw"].view(np.uint64)), "fold + metres->lambda, 10 M samples"
        assert np.array_equal(prep["w"].view(np.uint32), v["w"].view(np.uint32)), "off-grid weights"
        Vre, Vim = oracle.model_grid(Ic, None, float(p.freqs[0]), e.meta, _cfg(p))
        s_or, Vm_or, Vr_or = oracle.degrid_chi2(Vre, Vim, prep, p.N)
        scale = float(np.abs(Vm_or).max())
        on = v["w"] > 0
        dvm = float(np.abs(v["Vm"][on] - Vm_or[on]).max()) / scale
        print(f"\n[C2 full size] 0.5*chi2 engine {chi2:.8e} oracle {0.5 * s_or:.8e} (rel {abs(chi2 - 0.5 * s_or) / (0.5 * s_or):.2e}); "
              f"max |Vm - Vm_oracle| / max |Vm| = {dvm:.2e} over {int(on.sum())} samples")
        assert abs(chi2 - 0.5 * s_or) <= 1e-5 * 0.5 * s_or, (chi2, 0.5 * s_or)       # north-star tolerance on chi2
        assert dvm <= 2e-5, dvm
        g = torch.zeros_like(I_dev)
        e.dchi2(I_dev, g, flag_opt=0)
        assert e.last_grad_mode() == GRAD_UMMA
        g_again = torch.zeros_like(I_dev)
        e.dchi2(I_dev, g_again, flag_opt=0)
        assert torch.equal(g, g_again), "the tensor-core gradient is deterministic (no unordered atomics)"
        e.dchi2(I_dev, g_again, flag_opt=0)             # accumulates
        torch.testing.assert_close(g_again, 2 * g, rtol=2e-6, atol=0)
        assert not g[1].any()
        noise = e.get_noise_image()
        masked = torch.from_numpy(noise >= e.meta["noise_cut"]).cuda()
        assert not g[0][masked].any()
        # fp64 oracle over all 10 M visibilities at 192 pixels spread over the unmasked image
        rng = np.random.default_rng(12)
        cand = np.flatnonzero(noise.reshape(-1) < e.meta["noise_cut"])
        pix = np.sort(rng.choice(cand, 192, replace=False))
        d = oracle.dchi2(pix, p.N, v["uvw"], v["Vr"], v["w"], noise, None, float(p.freqs[0]), e.meta, _cfg(p))
        truth = d * oracle.chain(I_dev.cpu().numpy(), pix, float(p.freqs[0]), e.meta, e.cfg.threshold, 0)
        got = g[0].cpu().numpy().reshape(-1)[pix]
        err = _rel(got, truth)
        print(f"\n[C2 full size] gradient rel-L2 vs fp64 oracle over 10 M visibilities at {len(pix)} pixels: {err:.3e}")
        assert err <= 2e-5, err          # north-star tolerance 1e-4; measured 5.7e-6 (1.2e-5 before the truncation correction)
    finally:
        e.close()


def test_mosaic_blocks_with_their_own_pointing_and_phase_centres(oracle):
    """Two blocks (mosaic fields) whose pointing centre (attenuation, Field::ref_xobs_pix) and phase centre
    (phase_rotate / DChi2, Field::phs_xobs_pix) differ from each other and from the image centre —
    the reference keeps them per field (src/mfs.cu:660-691, src/functions.cu:4371-4376, 3729-3733)."""
    torch = _torch()
    p = synth.make_problem(N=128, nvis=9000, nchan=2, freq0=2.3e11, bandwidth=2e9, seed=61, grid_fill=0.9)
    e = Engine.from_problem(p, grad_mode=GRAD_UMMA, keep_vm=True)
    try:
        m = e.meta
        centres = [((m["xpix"] + 9.0, m["ypix"] - 6.0), (m["xpix"] + 9.0, m["ypix"] - 6.0)),     # offset field
                   ((m["xpix"] - 11.0, m["ypix"] + 4.0), (m["xpix"] - 3.0, m["ypix"] + 5.0))]    # pointing != phase
        e._ck(e.lib.gvm_clear_channels(e.h))
        for c in range(2):
            e.add_channel(float(p.freqs[c]), p.uvw[c], p.Vo[c], p.w[c], p.antenna_diameter, m["pb_factor"],
                          m["pb_cutoff"], m["primary_beam"], centres[c][0], centres[c][1])
        noise_min = e.build_noise_image(m["noise_jypix"])
        m = dict(m, fg_scale=noise_min, noise_cut=float(np.float32(10.0) * np.float32(noise_min)))
        e.set_scalars(m["fg_scale"], m["noise_cut"], e.cfg.threshold)
        noise = e.get_noise_image()
        I = _test_image(e)
        I_dev = torch.from_numpy(I).cuda()
        chi2 = e.chi2(I_dev)
        Ic = I_dev.cpu().numpy()
        cfg = _cfg(p)
        total = 0.0
        pix = np.arange(0, p.N * p.N, 17)
        grad = np.zeros(len(pix))
        for c in range(2):
            ref_pix, phs_pix = centres[c]
            prep = oracle.prep(p.uvw[c], p.Vo[c], p.w[c], float(p.freqs[c]), m["deltau"], m["deltav"], p.N)
            Vre, Vim = oracle.model_grid(Ic, None, float(p.freqs[c]), m, cfg, ref_pix=ref_pix, phs_pix=phs_pix)
            ssum, Vm, Vr = oracle.degrid_chi2(Vre, Vim, prep, p.N)
            total += 0.5 * ssum
            v = e.get_vis(c, want=("uvw", "Vm", "Vr", "w"))
            scale = np.abs(Vm).max()
            assert np.abs(v["Vm"] - Vm).max() <= 3e-5 * scale, c
            d = oracle.dchi2(pix, p.N, v["uvw"], v["Vr"], v["w"], noise, None, float(p.freqs[c]), m, cfg,
                             ref_pix=ref_pix, phs_pix=phs_pix)
            grad += d * oracle.chain(Ic, pix, float(p.freqs[c]), m, e.cfg.threshold, 0)
        assert abs(chi2 - total) <= 1e-5 * total, (chi2, total)
        g = torch.zeros_like(I_dev)
        e.dchi2(I_dev, g, flag_opt=0)
        assert e.last_grad_mode() == GRAD_UMMA
        assert _rel(g[0].cpu().numpy().reshape(-1)[pix], grad) <= 3e-5
        e.set_grad_mode(GRAD_SIMT)
        g2 = torch.zeros_like(I_dev)
        e.dchi2(I_dev, g2, flag_opt=0)
        assert _rel(g2[0].cpu().numpy().reshape(-1)[pix], grad) <= 3e-5
    finally:
        e.close()


@pytest.mark.parametrize("n,bits", [(1, 8), (31, 3), (4096, 8), (4097, 13), (1_000_003, 20), (3_000_000, 32)])
def test_radix_sort_is_stable_and_sorted(n, bits):
    """csrc/sort.cu: hand-written LSD radix sort used by the tile-sorted upload and the gridding path."""
    from gpuvmem_b200 import lib
    L = lib.load_library()
    rng = np.random.default_rng(n)
    keys = (rng.integers(0, 1 << min(bits, 31), n, dtype=np.uint64) if bits < 32
            else rng.integers(0, 1 << 32, n, dtype=np.uint64)).astype(np.uint32)
    if n > 1000:
        keys[: n // 3] = keys[0]                       # long runs of equal keys
    vals = np.arange(n, dtype=np.uint32)
    k2, v2 = keys.copy(), vals.copy()
    assert L.gvm_sort_pairs_host(0, k2.ctypes.data, v2.ctypes.data, n, bits) == 0, L.gvm_last_error()
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k2, keys[order])
    assert np.array_equal(v2, vals[order]), "equal keys must keep their input order"


def test_tile_sorted_upload_is_invisible_to_the_caller(oracle):
    """gvm_add_channel stores the samples in uv-tile order; gvm_get_vis gives every array back in the caller's order,
    and the tiled degridder (shared-memory grid tile + bulk-copied streams) equals the untiled kernel sample for sample."""
    torch = _torch()
    p = synth.make_problem(N=256, nvis=70001, nchan=1, seed=77, grid_fill=2.3)    # |u| / deltau > N for some samples: they fall off the grid
    outs = []
    for untiled in ("0", "1"):
        os.environ["GVM_FORWARD_UNTILED"] = untiled
        code = ("import os, sys, numpy as np, torch; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
                "from gpuvmem_b200 import Engine, synth\n"
                "from test_edges_gpu import _test_image\n"
                "p = synth.make_problem(N=256, nvis=70001, nchan=1, seed=77, grid_fill=2.3)\n"
                "e = Engine.from_problem(p, keep_vm=True)\n"
                "I = torch.from_numpy(_test_image(e)).cuda(); chi2 = e.chi2(I)\n"
                "v = e.get_vis(0, want=('uvw', 'cell', 'Vo', 'Vm', 'Vr', 'w'))\n"
                "np.savez(sys.argv[1], chi2=chi2, **v); e.close()\n") % (ROOT, os.path.join(ROOT, "tests"))
        out = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"gvm_tiled_{untiled}.npz")
        r = subprocess.run([sys.executable, "-c", code, out], capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, GVM_FORWARD_UNTILED=untiled))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs.append(np.load(out))
    os.environ.pop("GVM_FORWARD_UNTILED", None)
    tiled, plain = outs
    assert (plain["w"] == 0).sum() > 10, "no off-grid samples in the test problem"
    for k in ("uvw", "cell", "Vo", "w", "Vm", "Vr"):
        assert np.array_equal(tiled[k], plain[k]), k
    assert abs(float(tiled["chi2"]) - float(plain["chi2"])) <= 2e-6 * float(plain["chi2"])   # same terms, other sum order
    prep = oracle.prep(p.uvw[0], p.Vo[0], p.w[0], float(p.freqs[0]), 1.0 / (p.M * np.deg2rad(p.DELTAX)),
                       1.0 / (p.N * np.deg2rad(p.DELTAY)), p.N)
    assert np.array_equal(prep["uvw"].view(np.uint64), tiled["uvw"].view(np.uint64))
    assert np.array_equal(prep["cell"], tiled["cell"]) or np.array_equal(prep["cell"][prep["w"] > 0], tiled["cell"][prep["w"] > 0])
