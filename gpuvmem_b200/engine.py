"""Python mirror of the reference's per-run setup + a thin object over the C-ABI.

``Engine.from_problem`` restates, on the host, what ``MFS::configure`` /
``MFS::setDevice`` derive before the optimizer starts (reference ``src/mfs.cu:79-916``):
uv cell sizes, reference frequency, antenna beam model, visibility noise and
synthesized-beam estimate, noise per pixel, phase-centre pixel, the starting image,
and (on the GPU, through ``gvm_build_noise_image``) the noise image, ``fg_scale``
and the scaled ``noise_cut``. The hot path itself is only ever executed by
``libgvmb200.so``; device buffers are torch CUDA tensors passed by pointer.
"""
import ctypes as C
import math

import numpy as np

from . import lib as _lib

GRAD_AUTO, GRAD_UMMA, GRAD_SIMT, GRAD_SIMT_EXACT = 0, 1, 2, 3
PRIOR = {"Entropy": 0, "L1-Norm": 1, "TotalVariation": 2, "TotalSquaredVariation": 3,
         "Laplacian": 4, "Quadratic": 5, "GEntropy": 6, "GL1Norm": 7}
WEIGHTING = {"Natural": 0, "Uniform": 1, "Briggs": 2, "Radial": 3}
RPDEG_D = math.pi / 180.0
LIGHTSPEED = np.float32(2.99792458e8)


class EngineError(RuntimeError):
    pass


def _ptr(t):
    """Device/host pointer of a torch tensor or numpy array (must be contiguous)."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        assert t.flags["C_CONTIGUOUS"]
        return t.ctypes.data
    assert t.is_contiguous()
    return t.data_ptr()


def beam_model(telescope, antenna_diameter, min_freq):
    """Per-telescope primary-beam model, reference src/MSFITSIO.cu:510-551."""
    f32 = np.float32
    max_wavelength = f32(LIGHTSPEED / f32(min_freq))
    if telescope == "ALMA":
        pb_factor, pb = f32(1.13), 0
    elif telescope == "EVLA":
        pb_factor, pb = f32(1.25), 1
    else:
        pb_factor, pb = f32(f32(3.8317059702075125) / f32(math.pi)), 1  # cyl_bessel_j_zero(1,1)/pi
    pb_cutoff = f32(pb_factor * f32(max_wavelength / f32(antenna_diameter)))
    return float(pb_factor), float(pb_cutoff), pb


def noise_and_beam(problem):
    """calculateNoiseAndBeam (reference src/functions.cu:1700-1840): sum of weights
    (sequential fp32), weighted second moments (fp64), noise = 0.5*sqrt(1/sum w)."""
    f32 = np.float32
    s_uu = s_vv = s_uv = 0.0
    sum_w = f32(0.0)
    for c in range(problem.nchan):
        w = problem.w[c]
        if len(w) == 0:
            continue
        lam = f32(LIGHTSPEED / f32(problem.freqs[c]))
        u = problem.uvw[c][:, 0] / np.float64(lam)
        v = problem.uvw[c][:, 1] / np.float64(lam)
        wd = w.astype(np.float64)
        s_uu += float(np.sum(u * u * wd))
        s_vv += float(np.sum(v * v * wd))
        s_uv += float(np.sum(u * v * wd))
        sum_w = f32(sum_w + np.cumsum(w, dtype=np.float32)[-1])  # reduceCPU: running float sum
    s_uu /= float(sum_w)
    s_vv /= float(sum_w)
    s_uv /= float(sum_w)
    variance = f32(f32(1.0) / sum_w)
    sq = math.sqrt((s_uu - s_vv) ** 2 + 4.0 * s_uv * s_uv)
    bmaj = 1.0 / math.sqrt(2.0) / math.pi / math.sqrt((s_uu + s_vv) - sq) / RPDEG_D
    bmin = 1.0 / math.sqrt(2.0) / math.pi / math.sqrt((s_uu + s_vv) + sq) / RPDEG_D
    bpa = -0.5 * math.atan2(2.0 * s_uv, s_uu - s_vv) / RPDEG_D
    vis_noise = f32(f32(0.5) * np.sqrt(variance))
    return float(sum_w), float(vis_noise), bmaj, bmin, bpa


def direccos(ra, dec, ra0, dec0):
    """reference src/directioncosines.cu:40-59"""
    dra = ra - ra0
    l = math.cos(dec) * math.sin(dra)
    m = math.sin(dec) * math.cos(dec0) - math.cos(dec) * math.sin(dec0) * math.cos(dra)
    return l, m


class Engine:
    """One engine = one GPU's share of the visibility blocks + the image-sized state."""

    def __init__(self, M, N, DELTAX, DELTAY, nu_0, eta=-1.0, minpix=1e-3, noise_cut=1e30,
                 threshold=0.0, fg_scale=1.0, device=0, grad_mode=GRAD_AUTO, keep_vm=False):
        self.lib = _lib.load_library()
        self.cfg = _lib.gvm_config(M, N, DELTAX, DELTAY, nu_0, eta, minpix, noise_cut, threshold,
                                   fg_scale, device, grad_mode, 1 if keep_vm else 0)
        h = C.c_void_p()
        if self.lib.gvm_create(C.byref(self.cfg), C.byref(h)) != 0:
            raise EngineError(self.lib.gvm_last_error().decode())
        self.h = h
        self.M, self.N = M, N
        self.device = device
        self.meta = {}

    # -- plumbing ---------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise EngineError(self.lib.gvm_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.gvm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.gvm_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def use_torch_stream(self):
        import torch
        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def synchronize(self):
        self._ck(self.lib.gvm_synchronize(self.h))

    def set_scalars(self, fg_scale, noise_cut, threshold):
        self.cfg.fg_scale, self.cfg.noise_cut, self.cfg.threshold = fg_scale, noise_cut, threshold
        self._ck(self.lib.gvm_set_scalars(self.h, fg_scale, noise_cut, threshold))

    def set_grad_mode(self, mode):
        self._ck(self.lib.gvm_set_grad_mode(self.h, mode))

    def set_flag_opt(self, flag):
        self._ck(self.lib.gvm_set_flag_opt(self.h, flag))

    # -- static inputs ----------------------------------------------------------
    def set_noise_image(self, noise):
        if isinstance(noise, np.ndarray):
            noise = np.ascontiguousarray(noise, dtype=np.float32)
            self._ck(self.lib.gvm_set_noise_image(self.h, noise.ctypes.data, 0))
        else:
            self._ck(self.lib.gvm_set_noise_image(self.h, _ptr(noise), 1))

    def build_noise_image(self, noise_jypix):
        out = C.c_float()
        self._ck(self.lib.gvm_build_noise_image(self.h, noise_jypix, C.byref(out)))
        return out.value

    def get_noise_image(self):
        out = np.empty((self.M, self.N), dtype=np.float32)
        self._ck(self.lib.gvm_get_noise_image(self.h, out.ctypes.data))
        return out

    def set_gcf(self, gcf):
        if gcf is None:
            self._ck(self.lib.gvm_set_gcf(self.h, None))
        else:
            gcf = np.ascontiguousarray(gcf, dtype=np.float32)
            self._ck(self.lib.gvm_set_gcf(self.h, gcf.ctypes.data))

    def set_degrid_kernel(self, table, support=None):
        """Convolutional degridding in the forward model (degriddingGPU, src/functions.cu:2205-2254);
        ``table`` [m][n] CKernel table, ``support`` (sx, sy); None restores the bilinear vis_mod."""
        if table is None:
            self._ck(self.lib.gvm_set_degrid_kernel(self.h, None, 0, 0, 0, 0))
            return
        table = np.ascontiguousarray(table, dtype=np.float32)
        m, n = table.shape
        sx, sy = support if support is not None else (n // 2, m // 2)
        self._ck(self.lib.gvm_set_degrid_kernel(self.h, table.ctypes.data, m, n, int(sx), int(sy)))

    def set_forward_mode(self, mode):
        """0 auto, 1 full plane (C2C + phase_rotate, the reference's pipeline), 2 half plane (R2C + per-tap rotation)."""
        self._ck(self.lib.gvm_set_forward_mode(self.h, mode))

    def last_forward_mode(self):
        return self.lib.gvm_last_forward_mode(self.h)

    def get_model_grid(self):
        out = np.empty((self.N, self.N, 2), np.float32)
        self._ck(self.lib.gvm_get_model_grid(self.h, out.ctypes.data))
        return out[..., 0] + 1j * out[..., 1]

    def add_channel(self, freq, uvw_m, Vo, w, antenna_diameter, pb_factor, pb_cutoff,
                    primary_beam, ref_pix, phs_pix):
        d = _lib.gvm_channel_desc(freq, antenna_diameter, pb_factor, pb_cutoff, primary_beam,
                                  ref_pix[0], ref_pix[1], phs_pix[0], phs_pix[1])
        uvw_m = np.ascontiguousarray(uvw_m, dtype=np.float64)
        Vo = np.ascontiguousarray(Vo, dtype=np.float32)
        w = np.ascontiguousarray(w, dtype=np.float32)
        chan = C.c_int()
        self._ck(self.lib.gvm_add_channel(self.h, C.byref(d), len(w), uvw_m.ctypes.data,
                                          Vo.ctypes.data, w.ctypes.data, C.byref(chan)))
        return chan.value

    def num_channels(self):
        return self.lib.gvm_num_channels(self.h)

    def nvis(self, chan):
        return self.lib.gvm_channel_nvis(self.h, chan)

    def get_vis(self, chan, want=("uvw", "cell", "Vo", "Vr", "w")):
        Z = self.nvis(chan)
        out = {}
        bufs = {"uvw": np.empty((Z, 3), np.float64), "cell": np.empty((Z, 2), np.int32),
                "Vo": np.empty((Z, 2), np.float32), "Vm": np.empty((Z, 2), np.float32),
                "Vr": np.empty((Z, 2), np.float32), "w": np.empty(Z, np.float32)}
        args = [bufs[k].ctypes.data if k in want else None for k in ("uvw", "cell", "Vo", "Vm", "Vr", "w")]
        self._ck(self.lib.gvm_get_vis(self.h, chan, *args))
        for k in want:
            out[k] = bufs[k]
        return out

    # -- hot path ---------------------------------------------------------------
    def chi2(self, I_dev, normalize=False):
        out = C.c_float()
        self._ck(self.lib.gvm_chi2(self.h, _ptr(I_dev), int(normalize), C.byref(out)))
        return out.value

    def chi2_async(self, I_dev, normalize=False, out_dev=None):
        self._ck(self.lib.gvm_chi2_async(self.h, _ptr(I_dev), int(normalize), _ptr(out_dev)))

    def dchi2(self, I_dev, result_dev, flag_opt=0, normalize=False):
        self._ck(self.lib.gvm_dchi2(self.h, _ptr(I_dev), flag_opt, int(normalize), _ptr(result_dev)))

    def error_maps(self, I_dev, errors_dev, dist_mode=0):
        """calculateErrors (src/functions.cu:4966-5040): errors_dev [2][M][N] <- (sigma I_nu0, sigma alpha)."""
        self._ck(self.lib.gvm_error_maps(self.h, _ptr(I_dev), dist_mode, _ptr(errors_dev)))

    def eval_host(self, I_host, grad_host, flag_opt=0, normalize=False):
        out = C.c_float()
        self._ck(self.lib.gvm_eval_host(self.h, _ptr(I_host), flag_opt, int(normalize), C.byref(out),
                                        _ptr(grad_host)))
        return out.value

    # -- priors / vector ops ----------------------------------------------------
    def _pp(self, prior_value=0.001, eta=None, epsilon=1e-12, epsilon_b=0.0, prior_image=None):
        return _lib.gvm_prior_params(prior_value, self.cfg.eta if eta is None else eta, epsilon,
                                     epsilon_b, _ptr(prior_image))

    def prior_value(self, kind, I_dev, image_index=0, **kw):
        out = C.c_float()
        pp = self._pp(**kw)
        self._ck(self.lib.gvm_prior_value(self.h, PRIOR.get(kind, kind), _ptr(I_dev), image_index,
                                          C.byref(pp), C.byref(out)))
        return out.value

    def prior_grad(self, kind, I_dev, dgi_dev, lam, image_index=0, **kw):
        pp = self._pp(**kw)
        self._ck(self.lib.gvm_prior_grad(self.h, PRIOR.get(kind, kind), _ptr(I_dev), image_index,
                                         C.byref(pp), lam, _ptr(dgi_dev)))

    def add_to_dphi(self, dphi_dev, dgi_dev, index=0):
        self._ck(self.lib.gvm_add_to_dphi(self.h, _ptr(dphi_dev), _ptr(dgi_dev), index))

    def vec_evaluate_xt(self, xt, pcom, xicom, x, image_count=2, nopositivity=False):
        self._ck(self.lib.gvm_vec_evaluate_xt(self.h, _ptr(xt), _ptr(pcom), _ptr(xicom), x,
                                              image_count, int(nopositivity)))

    def vec_new_p(self, p, xi, xmin, image_count=2, nopositivity=False):
        self._ck(self.lib.gvm_vec_new_p(self.h, _ptr(p), _ptr(xi), xmin, image_count, int(nopositivity)))

    def vec_dot(self, a, b, n):
        out = C.c_float()
        self._ck(self.lib.gvm_vec_dot(self.h, _ptr(a), _ptr(b), n, C.byref(out)))
        return out.value

    def vec_gg_dgg(self, xi, g, image_count=2):
        a, b = C.c_float(), C.c_float()
        self._ck(self.lib.gvm_vec_gg_dgg(self.h, _ptr(xi), _ptr(g), image_count, C.byref(a), C.byref(b)))
        return a.value, b.value

    def vec_grad_condition(self, xi, p, den, image_count=2):
        out = C.c_float()
        self._ck(self.lib.gvm_vec_grad_condition(self.h, _ptr(xi), _ptr(p), den, image_count, C.byref(out)))
        return out.value

    def vec_new_xi(self, g, xi, h, gam, image_count=2):
        self._ck(self.lib.gvm_vec_new_xi(self.h, _ptr(g), _ptr(xi), _ptr(h), gam, image_count))

    def vec_axpby(self, a, x, b, y, n):
        self._ck(self.lib.gvm_vec_axpby(self.h, a, _ptr(x), b, _ptr(y), n))

    # -- telemetry --------------------------------------------------------------
    def launch_count(self):
        return self.lib.gvm_launch_count(self.h)

    def last_grad_kernel_ms(self):
        ms, n = C.c_float(), C.c_int()
        self._ck(self.lib.gvm_last_grad_kernel_ms(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def last_grad_mode(self):
        return self.lib.gvm_last_grad_mode(self.h)

    def grad_plan(self):
        """(tiles, output pixels) of the tensor-core gradient's plan over the unmasked pixels."""
        nt, px = C.c_int(), C.c_int64()
        self._ck(self.lib.gvm_grad_plan(self.h, C.byref(nt), C.byref(px)))
        return nt.value, px.value

    # -- MFS::configure / setDevice on the host --------------------------------
    @classmethod
    def from_problem(cls, p, device=0, z0=0.001, alpha0=0.0, eta=-1.0, noise_cut=10.0,
                     threshold_sigmas=0.0, nu_0=-1.0, grad_mode=GRAD_AUTO, keep_vm=False,
                     channels=None, vis_slice=None, normalize=False):
        """``channels``: indices of the channels this engine owns (channel sharding);
        ``vis_slice``: (start, stop) fraction-free sample range within each channel
        (visibility-chunk sharding). The derived scalars always come from the WHOLE
        problem, so every rank computes identical values."""
        f32 = np.float32
        min_f, max_f = f32(p.freqs.min()), f32(p.freqs.max())
        nu0 = f32(nu_0) if nu_0 > 0 else f32(f32(0.5) * f32(max_f + min_f))     # src/mfs.cu:329-334
        pb_factor, pb_cutoff, pb = beam_model(p.telescope, p.antenna_diameter, min_f)
        sum_w, vis_noise, bmaj, bmin, bpa = noise_and_beam(p)
        bmaj_pix, bmin_pix = bmaj / abs(p.DELTAX), bmin / abs(p.DELTAX)           # src/mfs.cu:637-638
        noise_jypix = f32(vis_noise / (math.pi * bmaj_pix * bmin_pix / (4.0 * float(np.log(f32(2.0))))))
        deltax, deltay = RPDEG_D * p.DELTAX, RPDEG_D * p.DELTAY
        l, m = direccos(p.ra * RPDEG_D, p.dec * RPDEG_D, p.ra * RPDEG_D, p.dec * RPDEG_D)
        xpix = f32(l / deltax + float(f32(p.crpix1) - f32(1.0)))                  # src/mfs.cu:681-691
        ypix = f32(m / deltay + float(f32(p.crpix2) - f32(1.0)))
        minpix = f32(f32(z0) * f32(-1.0) * f32(eta))                              # src/mfs.cu:169
        e = cls(p.M, p.N, p.DELTAX, p.DELTAY, float(nu0), eta=eta, minpix=float(minpix),
                noise_cut=1e30, threshold=float(f32(threshold_sigmas) * f32(5.0)), fg_scale=1.0,
                device=device, grad_mode=grad_mode, keep_vm=keep_vm)
        chans = range(p.nchan) if channels is None else channels
        for c in chans:
            sl = slice(None) if vis_slice is None else slice(vis_slice[0], vis_slice[1])
            e.add_channel(float(p.freqs[c]), p.uvw[c][sl], p.Vo[c][sl], p.w[c][sl],
                          p.antenna_diameter, pb_factor, pb_cutoff, pb, (float(xpix), float(ypix)),
                          (float(xpix), float(ypix)))
        noise_min = e.build_noise_image(float(noise_jypix))
        fg_scale = 1.0 if normalize else noise_min                                # src/mfs.cu:912, 983-984
        e.set_scalars(fg_scale, float(f32(noise_cut) * f32(noise_min)), e.cfg.threshold)
        e.meta = dict(nu_0=float(nu0), pb_factor=pb_factor, pb_cutoff=pb_cutoff, primary_beam=pb,
                      sum_weights=sum_w, vis_noise=vis_noise, bmaj_deg=bmaj, bmin_deg=bmin, bpa_deg=bpa,
                      noise_jypix=float(noise_jypix), xpix=float(xpix), ypix=float(ypix),
                      minpix=float(minpix), alpha0=float(alpha0), fg_scale=fg_scale,
                      noise_cut=float(f32(noise_cut) * f32(noise_min)),
                      deltau=1.0 / (p.M * deltax), deltav=1.0 / (p.N * deltay))
        return e

    def initial_image(self):
        """host_I of MFS::setDevice (src/mfs.cu:742-750): constant images."""
        I = np.empty((2, self.M, self.N), dtype=np.float32)
        I[0] = self.meta["minpix"]
        I[1] = self.meta["alpha0"]
        return I


# -- weights and gridding (stateless entry points of the C-ABI) ------------------------------
def weights(scheme, robust, M, N, deltau, deltav, uvw_list, freqs, w_list, device=0, taper=None):
    """WeightingScheme::apply for one dataset (reference src/*weightingscheme.cu). ``uvw_list[b]``:
    [Z_b][3] float64 metres; ``w_list[b]``: float32 weights, returned as NEW arrays."""
    lib = _lib.load_library()
    nb = len(w_list)
    uvw = [np.ascontiguousarray(u, dtype=np.float64) for u in uvw_list]
    w = [np.array(x, dtype=np.float32, copy=True) for x in w_list]
    Z = (C.c_int64 * nb)(*[len(x) for x in w])
    up = (C.c_void_p * nb)(*[u.ctypes.data for u in uvw])
    wp = (C.c_void_p * nb)(*[x.ctypes.data for x in w])
    fr = np.ascontiguousarray(freqs, dtype=np.float32)
    tp = None
    if taper is not None:
        tp = C.byref(_lib.gvm_taper(1, *taper))
    rc = lib.gvm_weights(device, WEIGHTING.get(scheme, scheme), robust, M, N, deltau, deltav, nb,
                         C.cast(Z, C.c_void_p), C.cast(up, C.c_void_p), fr.ctypes.data,
                         C.cast(wp, C.c_void_p), tp)
    if rc != 0:
        raise EngineError(lib.gvm_last_error().decode())
    return w


def grid_block(M, N, deltau, deltav, freq, uvw_m, Vo, w, table, support, device=0):
    """do_gridding for one (field, channel, stokes) block (reference src/functions.cu:1339-1653).
    Returns (uvw_out [n][3] metres, Vo_out [n][2], w_out [n]) in the reference's row-major order."""
    lib = _lib.load_library()
    uvw_m = np.ascontiguousarray(uvw_m, dtype=np.float64)
    Vo = np.ascontiguousarray(Vo, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    table = np.ascontiguousarray(table, dtype=np.float32)
    uo = np.empty((M * N, 3), np.float64)
    vo = np.empty((M * N, 2), np.float32)
    wo = np.empty(M * N, np.float32)
    n = C.c_int64()
    rc = lib.gvm_grid_block(device, M, N, deltau, deltav, freq, len(w), uvw_m.ctypes.data, Vo.ctypes.data,
                            w.ctypes.data, table.ctypes.data, table.shape[0], table.shape[1],
                            support[0], support[1], uo.ctypes.data, vo.ctypes.data, wo.ctypes.data, C.byref(n))
    if rc != 0:
        raise EngineError(lib.gvm_last_error().decode())
    k = n.value
    return uo[:k].copy(), vo[:k].copy(), wo[:k].copy()
