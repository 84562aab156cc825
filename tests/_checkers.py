"""ctypes wrappers of the two CHECKERS (test infrastructure, never the product):

* ``Oracle``  — oracle/liboracle.so, the plain-C CPU restatement (oracle/gvm_oracle.c)
* ``GvRef``   — oracle/_ref/libgvref.so, the reference's own sources compiled
  unmodified with stub third-party headers (oracle/Makefile, oracle/ref_harness.cu).
  Its host-only entry points run anywhere; its CUDA entry points need a GPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
GVREF_SO = os.path.join(ROOT, "oracle", "_ref", "libgvref.so")

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_V = C.c_void_p


def _arr_of_ptrs(arrays):
    return (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])


def ensure_oracle_built():
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return ORACLE_SO


class Oracle:
    KIND = {"PillBox2D": 0, "Gaussian2D": 1, "GaussianSinc2D": 2, "Sinc2D": 3, "PSWF": 4}

    def __init__(self):
        self.lib = C.CDLL(ensure_oracle_built())
        L = self.lib
        L.gvo_prep.argtypes = [C.c_long, f64p, f32p, f32p, C.c_float, C.c_double, C.c_double, C.c_long,
                               f64p, i32p, f64p, f32p, f32p]
        L.gvo_attenuation.restype = C.c_float
        L.gvo_clip.argtypes = [f32p, f32p, C.c_long, C.c_long, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]
        L.gvo_model_grid.argtypes = [f32p, _V, C.c_long] + [C.c_float] * 10 + [C.c_double] * 4 + [C.c_int, f64p, f64p]
        L.gvo_model_grid.restype = C.c_int
        L.gvo_degrid_chi2.argtypes = [C.c_long, C.c_long, f64p, f64p, i32p, f64p, f32p, f32p, _V, _V]
        L.gvo_degrid_chi2.restype = C.c_double
        L.gvo_dchi2.argtypes = [C.c_long, i64p, C.c_long, C.c_long, f64p, f32p, f32p, f32p, _V] + \
                               [C.c_float] * 10 + [C.c_double] * 2 + [C.c_int] * 3 + [f64p]
        L.gvo_error_accumulate.argtypes = [C.c_long, i64p, C.c_long, C.c_long, f64p, f32p, f32p, f32p, f32p] + \
                                          [C.c_float] * 8 + [C.c_double] * 2 + [C.c_int] * 2 + [f64p, f64p]
        L.gvo_error_reduce.argtypes = [C.c_long, f64p]
        L.gvo_degrid_conv.argtypes = [C.c_long, f64p, f32p, f32p, C.c_double, C.c_double] + [C.c_int] * 5 + [f32p]
        L.gvo_chain.argtypes = [f32p, C.c_long, C.c_long, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]
        L.gvo_chain.restype = C.c_double
        L.gvo_prior_value.argtypes = [C.c_int, f32p, f32p, _V, C.c_long] + [C.c_float] * 5
        L.gvo_prior_value.restype = C.c_double
        L.gvo_prior_grad.argtypes = [C.c_int, f32p, f32p, _V, C.c_long] + [C.c_float] * 6 + [f32p]
        L.gvo_noise_image.argtypes = [C.c_long] + [C.c_float] * 6 + [C.c_double] * 2 + [C.c_int, C.c_float, f32p]
        L.gvo_noise_image.restype = C.c_float
        L.gvo_weights.argtypes = [C.c_int, C.c_float, C.c_long, C.c_long, C.c_double, C.c_double, C.c_int,
                                  _V, _V, f32p, _V, _V]
        L.gvo_weight_cells.argtypes = [C.c_long, f64p, C.c_float, C.c_double, C.c_double, C.c_long, C.c_long, i64p]
        L.gvo_gridding.argtypes = [C.c_long, C.c_long, C.c_double, C.c_double, C.c_float, C.c_long, f64p,
                                   f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, f64p, f32p, f32p]
        L.gvo_gridding.restype = C.c_long
        L.gvo_ckernel.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, f32p]
        L.gvo_ckernel_default_w.restype = C.c_float
        L.gvo_ckernel_default_w.argtypes = [C.c_int]
        L.gvo_num_threads.restype = C.c_int

    def threads(self):
        return self.lib.gvo_num_threads()

    def set_threads(self, n):
        self.lib.gvo_set_threads(n)

    def prep(self, uvw_m, Vo, w, freq, deltau, deltav, N):
        Z = len(w)
        uvw_l = np.empty((Z, 3)); cell = np.empty((Z, 2), np.int32); frac = np.empty((Z, 2))
        Vo2 = np.empty((Z, 2), np.float32); w2 = np.empty(Z, np.float32)
        self.lib.gvo_prep(Z, np.ascontiguousarray(uvw_m, np.float64), np.ascontiguousarray(Vo, np.float32),
                          np.ascontiguousarray(w, np.float32), freq, deltau, deltav, N, uvw_l, cell, frac, Vo2, w2)
        return dict(uvw=uvw_l, cell=cell, frac=frac, Vo=Vo2, w=w2)

    def clip(self, I, noise, noise_cut, minpix, eta, threshold, schedule):
        M, N = I.shape[1], I.shape[2]
        self.lib.gvo_clip(I.reshape(-1), noise.reshape(-1), M, N, noise_cut, minpix, eta, threshold, schedule)

    def model_grid(self, I, gcf, nu, meta, cfg, ref_pix=None, phs_pix=None):
        """ref_pix / phs_pix: (x, y) pointing-centre and phase-centre pixels of the block (mosaics:
        Field::ref_xobs_pix / phs_xobs_pix, include/MSFITSIO.cuh:102-121); default: the image's."""
        N = I.shape[2]
        rx, ry = ref_pix if ref_pix is not None else (meta["xpix"], meta["ypix"])
        px, py = phs_pix if phs_pix is not None else (meta["xpix"], meta["ypix"])
        Vre = np.empty(N * N); Vim = np.empty(N * N)
        g = None if gcf is None else np.ascontiguousarray(gcf, np.float32).ctypes.data
        rc = self.lib.gvo_model_grid(np.ascontiguousarray(I.reshape(-1)), g, N, nu, meta["nu_0"], meta["minpix"],
                                     cfg["eta"], meta["fg_scale"], cfg["D"], meta["pb_factor"], meta["pb_cutoff"],
                                     rx, ry, float(np.float32(px)), float(np.float32(py)),
                                     cfg["DELTAX"], cfg["DELTAY"], meta["primary_beam"], Vre, Vim)
        assert rc == 0, "oracle FFT needs a power-of-two image"
        return Vre, Vim

    def degrid_chi2(self, Vre, Vim, prep, N):
        Z = len(prep["w"])
        Vm = np.empty((Z, 2), np.float32); Vr = np.empty((Z, 2), np.float32)
        s = self.lib.gvo_degrid_chi2(Z, N, Vre, Vim, prep["cell"], prep["frac"], prep["Vo"], prep["w"],
                                     Vm.ctypes.data, Vr.ctypes.data)
        return s, Vm, Vr

    def dchi2(self, pix, N, uvw_l, Vr, w, noise, gcf, nu, meta, cfg, normalize=0, fp32_phase=0, ref_pix=None,
              phs_pix=None):
        rx, ry = ref_pix if ref_pix is not None else (meta["xpix"], meta["ypix"])
        px, py = phs_pix if phs_pix is not None else (meta["xpix"], meta["ypix"])
        pix = np.ascontiguousarray(pix, np.int64)
        out = np.empty(len(pix))
        g = None if gcf is None else np.ascontiguousarray(gcf, np.float32).ctypes.data
        self.lib.gvo_dchi2(len(pix), pix, N, len(w), np.ascontiguousarray(uvw_l), np.ascontiguousarray(Vr),
                           np.ascontiguousarray(w), np.ascontiguousarray(noise.reshape(-1)), g,
                           meta["noise_cut"], meta["fg_scale"], cfg["D"], meta["pb_factor"], meta["pb_cutoff"],
                           nu, rx, ry, px, py,
                           cfg["DELTAX"], cfg["DELTAY"], meta["primary_beam"], normalize, fp32_phase, out)
        return out

    def degrid_conv(self, uvw_lambda, Vg_centred, table, du, dv, sx, sy):
        """degriddingGPU (src/functions.cu:2205-2254) on a centred complex grid [M][N]."""
        M, N = Vg_centred.shape
        g = np.ascontiguousarray(np.stack([Vg_centred.real, Vg_centred.imag], -1), np.float32).reshape(-1)
        t = np.ascontiguousarray(table, np.float32)
        out = np.zeros(2 * len(uvw_lambda), np.float32)
        self.lib.gvo_degrid_conv(len(uvw_lambda), np.ascontiguousarray(uvw_lambda, np.float64), g, t.reshape(-1), du, dv,
                                 M, N, t.shape[1], sx, sy, out)
        return out.reshape(-1, 2)

    def error_maps(self, pix, N, blocks, noise, I, meta, cfg, fp32_xy=0):
        """calculateErrors at the pixels `pix`: blocks = [(uvw_lambda, Vr, w, nu), ...]; returns (err_I, err_alpha)."""
        pix = np.ascontiguousarray(pix, np.int64)
        a0, a1 = np.zeros(len(pix)), np.zeros(len(pix))
        for uvw_l, Vr, w, nu in blocks:
            self.lib.gvo_error_accumulate(len(pix), pix, N, len(w), np.ascontiguousarray(uvw_l), np.ascontiguousarray(Vr),
                                          np.ascontiguousarray(w), np.ascontiguousarray(noise.reshape(-1), np.float32),
                                          np.ascontiguousarray(I.reshape(-1), np.float32), meta["noise_cut"], cfg["D"],
                                          meta["pb_factor"], meta["pb_cutoff"], nu, meta["nu_0"], meta["xpix"], meta["ypix"],
                                          cfg["DELTAX"], cfg["DELTAY"], int(meta["primary_beam"]), fp32_xy, a0, a1)
        self.lib.gvo_error_reduce(len(pix), a0)
        self.lib.gvo_error_reduce(len(pix), a1)
        return a0, a1

    def chain(self, I, idx, nu, meta, threshold, flag_opt):
        MN = I.shape[1] * I.shape[2]
        flat = np.ascontiguousarray(I.reshape(-1))
        return np.array([self.lib.gvo_chain(flat, MN, int(i), nu, meta["nu_0"], meta["fg_scale"], threshold,
                                            flag_opt) for i in idx])

    def prior_value(self, kind, img, noise, noise_cut, G=0.001, eta=-1.0, eps=1e-12, eps_b=0.0, prior_image=None):
        N = img.shape[0]
        P = None if prior_image is None else np.ascontiguousarray(prior_image, np.float32).ctypes.data
        return self.lib.gvo_prior_value(kind, np.ascontiguousarray(img.reshape(-1)),
                                        np.ascontiguousarray(noise.reshape(-1)), P, N, noise_cut, G, eta, eps, eps_b)

    def prior_grad(self, kind, img, noise, noise_cut, lam, G=0.001, eta=-1.0, eps=1e-12, eps_b=0.0, prior_image=None):
        N = img.shape[0]
        out = np.empty(N * N, np.float32)
        P = None if prior_image is None else np.ascontiguousarray(prior_image, np.float32).ctypes.data
        self.lib.gvo_prior_grad(kind, np.ascontiguousarray(img.reshape(-1)), np.ascontiguousarray(noise.reshape(-1)),
                                P, N, noise_cut, G, eta, eps, eps_b, lam, out)
        return out.reshape(N, N)

    def noise_image(self, N, cfg, meta):
        out = np.empty(N * N, np.float32)
        mn = self.lib.gvo_noise_image(N, cfg["D"], meta["pb_factor"], meta["pb_cutoff"], meta["nu_0"],
                                      meta["xpix"], meta["ypix"], cfg["DELTAX"], cfg["DELTAY"],
                                      meta["primary_beam"], meta["noise_jypix"], out)
        return mn, out.reshape(N, N)

    def weights(self, scheme, robust, M, N, deltau, deltav, uvw_list, freqs, w_list, taper=None):
        Z = np.array([len(w) for w in w_list], dtype=np.int64)
        uv = [np.ascontiguousarray(u, np.float64) for u in uvw_list]
        ws = [np.array(w, dtype=np.float32, copy=True) for w in w_list]
        t = None if taper is None else np.ascontiguousarray(taper, np.float32).ctypes.data
        self.lib.gvo_weights(scheme, robust, M, N, deltau, deltav, len(ws), Z.ctypes.data, _arr_of_ptrs(uv),
                             np.ascontiguousarray(freqs, np.float32), _arr_of_ptrs(ws), t)
        return ws

    def weight_cells(self, uvw_m, freq, deltau, deltav, M, N):
        out = np.empty(len(uvw_m), np.int64)
        self.lib.gvo_weight_cells(len(uvw_m), np.ascontiguousarray(uvw_m, np.float64), freq, deltau, deltav, M, N, out)
        return out

    def ckernel(self, name, m, n, sx, sy, w=None, gcf=False):
        kind = self.KIND[name]
        if w is None:
            w = self.lib.gvo_ckernel_default_w(kind)
        t = np.empty(m * n, np.float32)
        self.lib.gvo_ckernel(kind, m, n, sx, sy, w, int(gcf), t)
        return t.reshape(m, n)

    def gridding(self, M, N, deltau, deltav, freq, uvw_m, Vo, w, table, support):
        cap = M * N
        uo = np.empty((cap, 3)); vo = np.empty((cap, 2), np.float32); wo = np.empty(cap, np.float32)
        n = self.lib.gvo_gridding(M, N, deltau, deltav, freq, len(w), np.ascontiguousarray(uvw_m, np.float64),
                                  np.ascontiguousarray(Vo, np.float32), np.ascontiguousarray(w, np.float32),
                                  np.ascontiguousarray(table.reshape(-1), np.float32), table.shape[0], table.shape[1],
                                  support[0], support[1], uo, vo, wo)
        return uo[:n].copy(), vo[:n].copy(), wo[:n].copy()


class GvRef:
    """The reference itself (oracle/_ref/libgvref.so)."""

    def __init__(self):
        if not os.path.exists(GVREF_SO):
            raise FileNotFoundError(GVREF_SO)
        self.lib = C.CDLL(GVREF_SO)
        L = self.lib
        L.gvref_problem_begin.argtypes = [C.c_long, C.c_long] + [C.c_double] * 6 + [C.c_char_p, C.c_float, C.c_int, f32p]
        L.gvref_problem_channel.argtypes = [C.c_int, C.c_long, f64p, f32p, f32p]
        L.gvref_cpu_weights.argtypes = [C.c_char_p, C.c_float, C.c_int, _V]
        L.gvref_cpu_ckernel.argtypes = [C.c_char_p, C.c_int, C.c_int, _V, _V, i32p]
        L.gvref_cpu_gridding.argtypes = [C.c_char_p, C.c_float, C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.gvref_cpu_gridded_count.restype = C.c_long
        L.gvref_cpu_gridded_count.argtypes = [C.c_int]
        L.gvref_cpu_gridded_fetch.argtypes = [C.c_int, f64p, f32p, f32p]
        L.gvref_init.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.gvref_scalars.argtypes = [f64p]
        L.gvref_noise_image.argtypes = [f32p]
        L.gvref_nvis.restype = C.c_long
        L.gvref_nvis.argtypes = [C.c_int]
        L.gvref_get_vis.argtypes = [C.c_int, _V, _V, _V, _V, _V]
        L.gvref_get_host_vis.argtypes = [C.c_int, _V, _V, _V]
        L.gvref_set_image.argtypes = [f32p]
        L.gvref_get_image.argtypes = [f32p]
        L.gvref_calc_function.restype = C.c_float
        L.gvref_calc_function.argtypes = [C.c_int, f32p, C.c_int]
        L.gvref_calc_gradient.argtypes = [C.c_int, C.c_int, f32p]
        L.gvref_time_evals.restype = C.c_float
        L.gvref_time_evals.argtypes = [C.c_int, C.c_int, C.c_int]
        L.gvref_run.argtypes = [_V, C.POINTER(C.c_float)]
        L.gvref_set_lbfgs_k.argtypes = [C.c_int]
        if hasattr(L, "gvref_set_field_centre"):
            L.gvref_set_field_centre.argtypes = [C.c_double, C.c_double]
        if hasattr(L, "gvref_cpu_last_seconds"):
            L.gvref_cpu_last_seconds.argtypes = [f64p]
        if hasattr(L, "gvref_degridding"):
            L.gvref_degridding.argtypes = [C.c_long, f64p, _V, f32p, C.c_double, C.c_double] + [C.c_int] * 6 + [_V]
        if hasattr(L, "gvref_error_image"):
            L.gvref_error_image.argtypes = [f32p]
        L.gvref_set_verbose.argtypes = [C.c_int]
        if hasattr(L, "gvref_prior_eval"):
            L.gvref_prior_eval.argtypes = [C.c_char_p, f32p, _V] + [C.c_float] * 5 + [C.c_int, C.c_int, C.c_int,
                                                                                     C.POINTER(C.c_float), f32p, _V]
        if hasattr(L, "gvref_write_residuals"):
            L.gvref_write_residuals.argtypes = [C.POINTER(C.c_float)]
            L.gvref_get_host_model.argtypes = [C.c_int, _V, _V, _V, _V]
        self.problem = None

    def set_problem(self, p):
        self.problem = p
        self.lib.gvref_problem_begin(p.M, p.N, p.DELTAX, p.DELTAY, p.ra, p.dec, p.crpix1, p.crpix2,
                                     p.telescope.encode(), p.antenna_diameter, p.nchan,
                                     np.ascontiguousarray(p.freqs, np.float32))
        if getattr(p, "field_centre", None) is not None:
            self.lib.gvref_set_field_centre(float(p.field_centre[0]), float(p.field_centre[1]))
        for c in range(p.nchan):
            self.lib.gvref_problem_channel(c, len(p.w[c]), np.ascontiguousarray(p.uvw[c], np.float64),
                                           np.ascontiguousarray(p.Vo[c], np.float32),
                                           np.ascontiguousarray(p.w[c], np.float32))

    # host-only reference code
    def cpu_weights(self, scheme, robust=0.0, threads=1):
        p = self.problem
        outs = [np.empty(len(w), np.float32) for w in p.w]
        rc = self.lib.gvref_cpu_weights(scheme.encode(), robust, threads, _arr_of_ptrs(outs))
        assert rc == 0
        return outs

    def cpu_ckernel(self, name, m, n, want_gcf=False):
        p = self.problem
        info = np.zeros(4, np.int32)
        table = np.zeros(max(m * n, 1), np.float32)
        gcf = np.zeros(p.M * p.N, np.float32) if want_gcf else None
        rc = self.lib.gvref_cpu_ckernel(name.encode(), m, n, table.ctypes.data,
                                        None if gcf is None else gcf.ctypes.data, info)
        assert rc == 0
        mm, nn = int(info[2]), int(info[3])
        return table[:mm * nn].reshape(mm, nn), (int(info[0]), int(info[1])), \
            (None if gcf is None else gcf.reshape(p.M, p.N))

    def cpu_gridding(self, ckname, m, n, scheme="", robust=0.0, threads=1):
        rc = self.lib.gvref_cpu_gridding(scheme.encode(), robust, ckname.encode(), m, n, threads)
        assert rc == 0
        out = []
        for c in range(self.problem.nchan):
            cnt = self.lib.gvref_cpu_gridded_count(c)
            u = np.empty((cnt, 3)); v = np.empty((cnt, 2), np.float32); w = np.empty(cnt, np.float32)
            self.lib.gvref_cpu_gridded_fetch(c, u, v, w)
            out.append((u, v, w))
        return out

    def cpu_last_seconds(self):
        """(WeightingScheme::apply, do_gridding) wall seconds of the last cpu_gridding call."""
        out = np.zeros(2)
        self.lib.gvref_cpu_last_seconds(out)
        return float(out[0]), float(out[1])

    # CUDA reference path (GPU box only)
    def init(self, args, optimizer="CG-FRPRMN", scheme="Natural", ckernel="PillBox2D", ck_m=1, ck_n=1, with_tv=0):
        rc = self.lib.gvref_init(args.encode(), optimizer.encode(), scheme.encode(), ckernel.encode(), ck_m, ck_n, with_tv)
        if rc != 0:
            raise RuntimeError(f"gvref_init failed ({rc})")

    def scalars(self):
        out = np.zeros(16)
        self.lib.gvref_scalars(out)
        keys = ["fg_scale", "noise_cut", "noise_jypix", "deltau", "deltav", "xpix", "ypix", "nu_0", "bmaj_pix",
                "bmin_pix", "bpa", "vis_noise", "pb_cutoff", "pb_factor", "eta", "threshold"]
        return dict(zip(keys, out.tolist()))

    def noise_image(self):
        p = self.problem
        out = np.empty(p.M * p.N, np.float32)
        self.lib.gvref_noise_image(out)
        return out.reshape(p.M, p.N)

    def get_vis(self, chan):
        Z = self.lib.gvref_nvis(chan)
        uvw = np.empty((Z, 3)); Vo = np.empty((Z, 2), np.float32); Vm = np.empty((Z, 2), np.float32)
        Vr = np.empty((Z, 2), np.float32); w = np.empty(Z, np.float32)
        self.lib.gvref_get_vis(chan, uvw.ctypes.data, Vo.ctypes.data, Vm.ctypes.data, Vr.ctypes.data, w.ctypes.data)
        return dict(uvw=uvw, Vo=Vo, Vm=Vm, Vr=Vr, w=w)

    def get_host_vis(self, chan):
        Z = self.lib.gvref_nvis(chan)
        uvw = np.empty((Z, 3)); Vo = np.empty((Z, 2), np.float32); w = np.empty(Z, np.float32)
        self.lib.gvref_get_host_vis(chan, uvw.ctypes.data, Vo.ctypes.data, w.ctypes.data)
        return dict(uvw=uvw, Vo=Vo, w=w)

    def set_image(self, I):
        self.lib.gvref_set_image(np.ascontiguousarray(I.reshape(-1), np.float32))

    def get_image(self):
        p = self.problem
        out = np.empty(2 * p.M * p.N, np.float32)
        self.lib.gvref_get_image(out)
        return out.reshape(2, p.M, p.N)

    def calc_function(self, iteration=0):
        fi = np.zeros(8, np.float32)
        v = self.lib.gvref_calc_function(iteration, fi, 8)
        return v, fi

    def calc_gradient(self, iteration=0, flag=0):
        p = self.problem
        out = np.empty(2 * p.M * p.N, np.float32)
        self.lib.gvref_calc_gradient(iteration, flag, out)
        return out.reshape(2, p.M, p.N)

    def degridding(self, uvw_lambda, Vg_centred, table, du, dv, sx, sy):
        """degriddingGPU (src/functions.cu:2205-2254) on a centred complex grid."""
        Z = len(uvw_lambda)
        M, N = Vg_centred.shape
        g = np.ascontiguousarray(np.stack([Vg_centred.real, Vg_centred.imag], -1), np.float32)
        out = np.zeros((Z, 2), np.float32)
        t = np.ascontiguousarray(table, np.float32)
        rc = self.lib.gvref_degridding(Z, np.ascontiguousarray(uvw_lambda, np.float64), g.ctypes.data, t.reshape(-1),
                                       du, dv, M, N, t.shape[0], t.shape[1], sx, sy, out.ctypes.data)
        assert rc == 0, rc
        return out

    def error_image(self):
        p = self.problem
        out = np.empty(2 * p.M * p.N, np.float32)
        self.lib.gvref_error_image(out)
        return out.reshape(2, p.M, p.N)

    def prior_eval(self, name, I, lam, image_index=0, iteration=1, prior_image=None, prior_value=0.001, eta=-1.0,
                   eps_a=1e-12, eps_b=1e-12, flag=None):
        """One Fi of the reference on its own: (get_fivalue(), dphi [2][M][N] after restartDGi + calcGi + addToDphi,
        the term's prior image after calcGi or None). flag: flag_opt, default = image_index (gradient gate open)."""
        p = self.problem
        val = C.c_float()
        dphi = np.zeros(2 * p.M * p.N, np.float32)
        pr = None if prior_image is None else np.ascontiguousarray(prior_image, np.float32)
        after = None if pr is None else np.zeros(p.M * p.N, np.float32)
        rc = self.lib.gvref_prior_eval(name.encode(), np.ascontiguousarray(I.reshape(-1), np.float32),
                                       None if pr is None else pr.ctypes.data, lam, prior_value, eta, eps_a, eps_b,
                                       image_index, iteration, image_index if flag is None else flag, C.byref(val), dphi,
                                       None if after is None else after.ctypes.data)
        assert rc == 0, rc
        return val.value, dphi.reshape(2, p.M, p.N), (None if after is None else after.reshape(p.M, p.N))

    def write_residuals(self):
        """MFS::writeResiduals; returns (non-gridded 0.5*chi2 of the last Chi2::calcFi, [per channel dict])."""
        v = C.c_float()
        rc = self.lib.gvref_write_residuals(C.byref(v))
        assert rc == 0, rc
        out = []
        for c in range(self.problem.nchan):
            Z = self.lib.gvref_nvis(c)
            uvw = np.empty((Z, 3)); Vo = np.empty((Z, 2), np.float32); Vm = np.empty((Z, 2), np.float32)
            w = np.empty(Z, np.float32)
            self.lib.gvref_get_host_model(c, uvw.ctypes.data, Vo.ctypes.data, Vm.ctypes.data, w.ctypes.data)
            out.append(dict(uvw=uvw, Vo=Vo, Vm=Vm, w=w))
        return v.value, out

    def time_evals(self, n, iteration=0, flag=0):
        return self.lib.gvref_time_evals(n, iteration, flag)

    def run(self):
        p = self.problem
        out = np.empty(2 * p.M * p.N, np.float32)
        ms = C.c_float()
        it = self.lib.gvref_run(out.ctypes.data, C.byref(ms))
        return out.reshape(2, p.M, p.N), it, ms.value
