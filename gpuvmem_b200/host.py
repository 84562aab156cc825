"""ctypes binding of the C++ host layer (``libgvmhost.so``, ``include/gvm_host.h``).

The host layer is gpuvmem's plugin surface restated in C++ on top of the engine's C ABI
(``gpuvmem_b200/csrc/host``): MFS synthesizer, CG / L-BFGS optimizers with the NR line
search, ObjectiveFunction + Fi terms, CKernel and WeightingScheme families, factories and
the reference's command line. This module only marshals arrays; there is no Python
arithmetic on the path and no CPU fallback.
"""
import ctypes as C
import json
import os

import numpy as np

from . import lib as _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_P = C.c_void_p
_HOST = None

DEFAULT_FI_SPEC = "Chi2:-1:0:0,Entropy:0:0:0,L1-Norm:1:0:0,TotalSquaredVariation:2:0:0,Laplacian:3:0:0"


def host_lib_path():
    return os.path.join(_HERE, "libgvmhost.so")


class gvmh_problem(C.Structure):
    _fields_ = [("M", C.c_int64), ("N", C.c_int64), ("DELTAX", C.c_double), ("DELTAY", C.c_double),
                ("ra", C.c_double), ("dec", C.c_double), ("crpix1", C.c_double), ("crpix2", C.c_double),
                ("telescope", C.c_char_p), ("antenna_diameter", C.c_float), ("beam_noise", C.c_float),
                ("nchan", C.c_int), ("freqs", _P), ("Z", _P), ("uvw_m", _P), ("Vo", _P), ("w", _P),
                ("has_field_centre", C.c_int), ("field_ra", C.c_double), ("field_dec", C.c_double)]


FN1D = C.CFUNCTYPE(C.c_float, C.c_float, _P)

# every symbol include/gvm_host.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gvmh_create": (C.c_int, [C.POINTER(gvmh_problem), C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int,
                              C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.POINTER(_P)]),
    "gvmh_destroy": (C.c_int, [_P]),
    "gvmh_set_quiet": (C.c_int, [C.c_int]),
    "gvmh_run": (C.c_int, [_P, _P, C.POINTER(C.c_double)]),
    "gvmh_clear_run": (C.c_int, [_P]),
    "gvmh_set_lbfgs_k": (C.c_int, [_P, C.c_int]),
    "gvmh_write_outputs": (C.c_int, [_P]),
    "gvmh_use_ckernel_degridding": (C.c_int, [_P, C.c_int]),
    "gvmh_write_residuals": (C.c_int, [_P, _P]),
    "gvmh_get_host_model": (C.c_int, [_P, C.c_int, _P, _P]),
    "gvmh_fits_read": (C.c_int, [C.c_char_p, _P, _P, C.c_int64]),
    "gvmh_fits_write": (C.c_int, [C.c_char_p, _P, C.c_int64, C.c_int64, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p,
                                  C.c_float, C.c_double, C.c_double]),
    "gvmh_error_image": (C.c_int, [_P, _P]),
    "gvmh_fi_eval": (C.c_int, [_P, C.c_char_p, _P, _P] + [C.c_float] * 5 + [C.c_int] * 3 + [C.POINTER(C.c_float), _P, _P]),
    "gvmh_filter_gridding": (C.c_int, [_P, C.c_char_p, C.c_int, C.c_int]),
    "gvmh_set_image": (C.c_int, [_P, _P]),
    "gvmh_get_image": (C.c_int, [_P, _P]),
    "gvmh_set_iteration": (C.c_int, [_P, C.c_int]),
    "gvmh_set_flag": (C.c_int, [_P, C.c_int]),
    "gvmh_calc_function": (C.c_int, [_P, C.POINTER(C.c_float), _P, C.c_int]),
    "gvmh_calc_gradient": (C.c_int, [_P, C.c_int, _P]),
    "gvmh_eval_device": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float)]),
    "gvmh_eval_host": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_float), _P]),
    "gvmh_engine": (_P, [_P]),
    "gvmh_scalars": (C.c_int, [_P, _P]),
    "gvmh_stats": (C.c_int, [_P, _P, _P]),
    "gvmh_nvis": (C.c_int64, [_P, C.c_int]),
    "gvmh_get_host_vis": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "gvmh_exit_reason": (C.c_char_p, [_P]),
    "gvmh_history": (C.c_int, [_P, _P, C.c_int]),
    "gvmh_ckernel_table": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, _P,
                                     C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gvmh_ckernel_gcf": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _P]),
    "gvmh_factory_has": (C.c_int, [C.c_char_p, C.c_char_p]),
    "gvmh_parse_args": (C.c_int, [C.c_char_p, C.c_char_p, C.c_size_t]),
    "gvmh_linmin_1d": (C.c_int, [FN1D, _P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "gvmh_read_gvms": (C.c_int, [C.c_char_p, _P]),
    "gvmh_shard_plan": (C.c_int, [C.c_int, _P, C.c_int, C.c_int, _P, _P]),
}


def load_host_library():
    """Load libgvmhost.so (which links libgvmb200.so). Raises if either is not built."""
    global _HOST
    if _HOST is not None:
        return _HOST
    _lib.load_library()
    path = host_lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run python -c \"import __graft_entry__ as g; g.build()\"")
    h = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(h, name)
        fn.restype = res
        fn.argtypes = args
    _HOST = h
    return h


SCALAR_NAMES = ("fg_scale", "noise_cut", "noise_jypix", "nu_0", "vis_noise", "sum_weights", "bmaj_deg",
                "bmin_deg", "bpa_deg", "deltau", "deltav", "xobs_pix", "yobs_pix", "total_visibilities",
                "iterations_done", "n_fi")


class Session:
    """One reconstruction set-up: what ``main()`` of the reference builds before ``sy->run()``
    (``src/main.cu:147-212``), from a synthetic :class:`gpuvmem_b200.synth.Problem`."""

    def __init__(self, problem, args="-z 0.001 -Z 0.0 -t 10", optimizer="CG-FRPRMN", scheme="Natural",
                 ckernel="PillBox2D", ck_size=(0, 0), fi_spec=None, rank=0, world=1, nccl_id=None,
                 channels=None, shape=None):
        """``problem`` None: the datasets and the image header are read from the -i / -m files named in ``args``
        (GVMS container, FITS model image); ``shape`` = (M, N) of that image is then required."""
        self.h = load_host_library()
        p = problem
        if p is None:      # datasets and header come from the -i / -m files named in args
            prob_ref = None
        else:
            chans = list(range(p.nchan)) if channels is None else list(channels)
            self._keep = dict(
                freqs=np.ascontiguousarray(p.freqs[chans], dtype=np.float32),
                Z=np.array([len(p.w[c]) for c in chans], dtype=np.int64),
                uvw=[np.ascontiguousarray(p.uvw[c], dtype=np.float64) for c in chans],
                Vo=[np.ascontiguousarray(p.Vo[c], dtype=np.float32) for c in chans],
                w=[np.ascontiguousarray(p.w[c], dtype=np.float32) for c in chans])
            k = self._keep
            n = len(chans)
            k["uvw_p"] = (C.c_void_p * n)(*[a.ctypes.data for a in k["uvw"]])
            k["Vo_p"] = (C.c_void_p * n)(*[a.ctypes.data for a in k["Vo"]])
            k["w_p"] = (C.c_void_p * n)(*[a.ctypes.data for a in k["w"]])
            prob = gvmh_problem(p.M, p.N, p.DELTAX, p.DELTAY, p.ra, p.dec, p.crpix1, p.crpix2,
                                p.telescope.encode(), p.antenna_diameter, -1.0, n, k["freqs"].ctypes.data,
                                k["Z"].ctypes.data, C.cast(k["uvw_p"], _P), C.cast(k["Vo_p"], _P), C.cast(k["w_p"], _P),
                                0, 0.0, 0.0)
            fc = getattr(p, "field_centre", None)     # (ra, dec) in degrees of the field when it is not the image centre
            if fc is not None:
                prob.has_field_centre, prob.field_ra, prob.field_dec = 1, float(fc[0]), float(fc[1])
            prob_ref = C.byref(prob)
        s = _P()
        rc = self.h.gvmh_create(prob_ref, args.encode(), optimizer.encode(), scheme.encode(), ckernel.encode(),
                                ck_size[0], ck_size[1], (fi_spec or DEFAULT_FI_SPEC).encode(), rank, world,
                                nccl_id, C.byref(s))
        if rc != 0:
            raise RuntimeError("gvmh_create failed: " + _lib.load_library().gvm_last_error().decode())
        self.s = s
        if p is not None:
            self.M, self.N = p.M, p.N
        else:
            self.M, self.N = shape
        self.eng = _lib.load_library()

    # -- lifecycle ------------------------------------------------------------------------
    def close(self):
        if getattr(self, "s", None):
            self.h.gvmh_destroy(self.s)
            self.s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reconstruction -------------------------------------------------------------------
    def run(self):
        img = np.empty((2, self.M, self.N), np.float32)
        sec = C.c_double()
        self.h.gvmh_run(self.s, img.ctypes.data, C.byref(sec))
        return img, sec.value

    def clear_run(self):
        self.h.gvmh_clear_run(self.s)

    def set_lbfgs_k(self, k):
        self.h.gvmh_set_lbfgs_k(self.s, k)

    def write_outputs(self):
        self.h.gvmh_write_outputs(self.s)

    def filter_gridding(self, ckernel="", ck_size=(0, 0)):
        """Filter "Gridding" on the session's visibilities (host side, in place)."""
        self.h.gvmh_filter_gridding(self.s, ckernel.encode(), ck_size[0], ck_size[1])

    def error_image(self):
        """Error "SecondDerivateError": (sigma I_nu0, sigma alpha) as [2][M][N]."""
        out = np.empty((2, self.M, self.N), np.float32)
        self.h.gvmh_error_image(self.s, out.ctypes.data)
        return out

    # -- objective function ---------------------------------------------------------------
    def set_image(self, I):
        I = np.ascontiguousarray(I, dtype=np.float32)
        self.h.gvmh_set_image(self.s, I.ctypes.data)

    def get_image(self):
        I = np.empty((2, self.M, self.N), np.float32)
        self.h.gvmh_get_image(self.s, I.ctypes.data)
        return I

    def set_iteration(self, it):
        self.h.gvmh_set_iteration(self.s, it)

    def set_flag(self, flag):
        self.h.gvmh_set_flag(self.s, flag)

    def calc_function(self):
        v = C.c_float()
        n = int(self.scalars()["n_fi"])
        fi = np.zeros(max(n, 1), np.float32)
        self.h.gvmh_calc_function(self.s, C.byref(v), fi.ctypes.data, n)
        return v.value, fi[:n]

    def calc_gradient(self, iteration, fetch=True):
        g = np.empty((2, self.M, self.N), np.float32) if fetch else None
        self.h.gvmh_calc_gradient(self.s, iteration, g.ctypes.data if fetch else None)
        return g

    def eval_device(self, iteration=1):
        v = C.c_float()
        self.h.gvmh_eval_device(self.s, iteration, C.byref(v))
        return v.value

    def eval_host(self, I_host_ptr, grad_host_ptr, iteration=1):
        v = C.c_float()
        self.h.gvmh_eval_host(self.s, I_host_ptr, iteration, C.byref(v), grad_host_ptr)
        return v.value

    # -- introspection --------------------------------------------------------------------
    def engine_handle(self):
        return self.h.gvmh_engine(self.s)

    def scalars(self):
        out = np.zeros(16, np.float64)
        self.h.gvmh_scalars(self.s, out.ctypes.data)
        return dict(zip(SCALAR_NAMES, out.tolist()))

    def stats(self):
        sec = np.zeros(6, np.float64)
        cnt = np.zeros(2, np.int64)
        self.h.gvmh_stats(self.s, sec.ctypes.data, cnt.ctypes.data)
        return dict(setup_s=sec[0], weighting_s=sec[1], gridding_s=sec[2], optimize_s=sec[3],
                    function_s=sec[4], gradient_s=sec[5], function_evals=int(cnt[0]), gradient_evals=int(cnt[1]))

    def host_vis(self, chan=0):
        n = self.h.gvmh_nvis(self.s, chan)
        uvw = np.empty((n, 3), np.float64)
        Vo = np.empty((n, 2), np.float32)
        w = np.empty(n, np.float32)
        self.h.gvmh_get_host_vis(self.s, chan, uvw.ctypes.data, Vo.ctypes.data, w.ctypes.data)
        return uvw, Vo, w

    def fi_eval(self, name, I, lam, image_index=0, iteration=1, prior_image=None, prior_value=0.001, eta=-1.0,
                eps_a=1e-12, eps_b=1e-12, flag=None):
        """One Fi term on its own through the host layer's classes: (value, dphi [2][M][N], prior image after calcGi)."""
        N = I.shape[-1]
        val = C.c_float()
        dphi = np.zeros(I.size, np.float32)
        Ic = np.ascontiguousarray(I.reshape(-1), np.float32)
        pr = None if prior_image is None else np.ascontiguousarray(prior_image, np.float32)
        after = None if pr is None else np.zeros(pr.size, np.float32)
        rc = self.h.gvmh_fi_eval(self.s, name.encode(), Ic.ctypes.data, None if pr is None else pr.ctypes.data, lam,
                                 prior_value, eta, eps_a, eps_b, image_index, iteration,
                                 image_index if flag is None else flag, C.byref(val), dphi.ctypes.data,
                                 None if after is None else after.ctypes.data)
        assert rc == 0, rc
        return val.value, dphi.reshape(I.shape), (None if after is None else after.reshape(N, N))

    def use_ckernel_degridding(self, on=True):
        self.h.gvmh_use_ckernel_degridding(self.s, int(on))

    def write_residuals(self):
        """MFS::writeResiduals; returns (non-gridded 0.5*chi2 or 0, [per channel dict(uvw, Vo, w, Vm, Vr)])."""
        v = C.c_float()
        self.h.gvmh_write_residuals(self.s, C.byref(v))
        out = []
        c = 0
        while self.h.gvmh_nvis(self.s, c) >= 0:
            uvw, Vo, w = self.host_vis(c)
            Vm = np.empty_like(Vo); Vr = np.empty_like(Vo)
            self.h.gvmh_get_host_model(self.s, c, Vm.ctypes.data, Vr.ctypes.data)
            out.append(dict(uvw=uvw, Vo=Vo, w=w, Vm=Vm, Vr=Vr))
            c += 1
        return v.value, out

    def exit_reason(self):
        return self.h.gvmh_exit_reason(self.s).decode()

    def history(self):
        buf = np.zeros(4096, np.float32)
        n = self.h.gvmh_history(self.s, buf.ctypes.data, len(buf))
        return buf[:min(n, len(buf))].copy()

    def launch_count(self):
        return self.eng.gvm_launch_count(self.engine_handle())

    def last_grad_kernel_ms(self):
        ms, n = C.c_float(), C.c_int()
        self.eng.gvm_last_grad_kernel_ms(self.engine_handle(), C.byref(ms), C.byref(n))
        return ms.value, n.value

    def last_grad_mode(self):
        return self.eng.gvm_last_grad_mode(self.engine_handle())

    def grad_plan(self):
        nt, px = C.c_int(), C.c_int64()
        self.eng.gvm_grad_plan(self.engine_handle(), C.byref(nt), C.byref(px))
        return nt.value, px.value

    def local_nvis(self):
        e = self.engine_handle()
        return sum(self.eng.gvm_channel_nvis(e, c) for c in range(self.eng.gvm_num_channels(e)))

    def collectives(self):
        return self.eng.gvm_dist_collectives(self.engine_handle())


# -- stateless helpers (no GPU) --------------------------------------------------------------
def ckernel_table(name, m, n, sx, sy, w=-1.0):
    h = load_host_library()
    t = np.zeros((m, n), np.float32)
    a, b = C.c_int(), C.c_int()
    h.gvmh_ckernel_table(name.encode(), m, n, sx, sy, w, t.ctypes.data, C.byref(a), C.byref(b))
    return t, (a.value, b.value)


def ckernel_gcf(name, m, n, M, N, dx, dy):
    h = load_host_library()
    g = np.zeros((M, N), np.float32)
    h.gvmh_ckernel_gcf(name.encode(), m, n, M, N, dx, dy, g.ctypes.data)
    return g


FITS_HEADER_NAMES = ["naxis1", "naxis2", "bitpix", "has_wcs", "cdelt1", "cdelt2", "crval1", "crval2", "crpix1", "crpix2",
                     "bmaj", "bmin", "bpa", "noise", "equinox", "ncards"]


def fits_read(path, want_data=True):
    """The host layer's FITS reader (csrc/host/fits.cpp): (header dict, data [naxis2][naxis1] float32 or None)."""
    h = load_host_library()
    hdr = np.zeros(16)
    if h.gvmh_fits_read(path.encode(), hdr.ctypes.data, None, 0) != 0:
        raise RuntimeError(f"cannot read {path} as FITS")
    d = dict(zip(FITS_HEADER_NAMES, hdr.tolist()))
    data = None
    if want_data:
        data = np.empty((int(d["naxis2"]), int(d["naxis1"])), np.float32)
        if h.gvmh_fits_read(path.encode(), hdr.ctypes.data, data.ctypes.data, data.size) != 0:
            raise RuntimeError(f"cannot read the image of {path}")
    return d, data


def fits_write(path, data, template=None, bunit="JY/PIXEL", niter=0, radesys="ICRS", equinox=2000.0, crval1=0.0,
               crval2=0.0):
    """The host layer's FITS writer (OCopyFITS semantics: template header copied, a few keys replaced)."""
    data = np.ascontiguousarray(data, dtype=np.float32)
    rc = load_host_library().gvmh_fits_write(path.encode(), data.ctypes.data, data.shape[1], data.shape[0],
                                             template.encode() if template else None, bunit.encode(), niter,
                                             radesys.encode(), equinox, crval1, crval2)
    if rc != 0:
        raise RuntimeError(f"cannot write {path}")


def factory_has(kind, name):
    return bool(load_host_library().gvmh_factory_has(kind.encode(), name.encode()))


def parse_args(args):
    buf = C.create_string_buffer(4096)
    load_host_library().gvmh_parse_args(args.encode(), buf, len(buf))
    return json.loads(buf.value.decode())


def linmin_1d(f):
    h = load_host_library()
    cb = FN1D(lambda x, _u: float(f(x)))
    xm, fm, n = C.c_float(), C.c_float(), C.c_int()
    h.gvmh_linmin_1d(cb, None, C.byref(xm), C.byref(fm), C.byref(n))
    return xm.value, fm.value, n.value


def read_gvms(path):
    out = np.zeros(8, np.float64)
    rc = load_host_library().gvmh_read_gvms(path.encode(), out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"cannot read {path}")
    return dict(M=int(out[0]), N=int(out[1]), nchan=int(out[2]), total_vis=int(out[3]), min_freq=out[4],
                max_freq=out[5], max_blength=out[6], uvmax_wavelength=out[7])


def set_quiet(quiet=True):
    load_host_library().gvmh_set_quiet(1 if quiet else 0)


def shard_plan(Z, world, rank):
    """[(lo, hi)] per channel: what MFS::setDevice uploads on ``rank`` of ``world``."""
    Z = np.ascontiguousarray(Z, dtype=np.int64)
    lo, hi = np.zeros_like(Z), np.zeros_like(Z)
    load_host_library().gvmh_shard_plan(len(Z), Z.ctypes.data, world, rank, lo.ctypes.data, hi.ctypes.data)
    return list(zip(lo.tolist(), hi.tolist()))


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    e = _lib.load_library()
    if e.gvm_dist_unique_id(buf, 128) != 0:
        raise RuntimeError(e.gvm_last_error().decode())
    return buf.raw
