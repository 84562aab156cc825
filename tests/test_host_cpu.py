"""CPU tests of the C++ host layer (gpuvmem_b200/csrc/host, libgvmhost.so): the library loads
and exports every symbol of include/gvm_host.h; CKernel tables / GCF images are bit-identical
to the reference's own host code and to the oracle; factory keys, command-line parsing, the
GVMS reader; and the NR line search follows the oracle's restatement probe for probe."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

from gpuvmem_b200 import host, synth
from gpuvmem_b200.engine import RPDEG_D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_library_exports_every_declared_symbol():
    src = open(os.path.join(ROOT, "include", "gvm_host.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(gvmh_[a-z0-9_]+)\s*\(", src)))
    assert len(names) >= 25
    so = ctypes.CDLL(host.host_lib_path())
    for n in names:
        assert hasattr(so, n), f"{n} declared in include/gvm_host.h but not exported"
    assert set(host.SIGNATURES) == set(names), set(host.SIGNATURES) ^ set(names)


def test_factory_keys_match_the_reference():
    # the string ids the reference registers
    keys = {"Synthesizer": ["MFS"], "Optimizer": ["CG-FRPRMN", "CG-LBFGS"], "ObjectiveFunction": ["ObjectiveFunction"],
            "Fi": ["Chi2", "Entropy", "L1-Norm", "TotalVariation", "TotalSquaredVariation", "Laplacian", "Quadratic",
                   "GEntropy", "GL1Norm"],
            "CKernel": ["PillBox2D", "Gaussian2D", "GaussianSinc2D", "Sinc2D", "PSWF"],
            "WeightingScheme": ["Natural", "Uniform", "Briggs", "Radial"], "Io": ["IoMS", "IoFITS"],
            "Filter": ["Gridding"], "Error": ["SecondDerivateError"]}
    for kind, names in keys.items():
        for n in names:
            assert host.factory_has(kind, n), (kind, n)
    assert not host.factory_has("CKernel", "PSWF_12D")   # the class name is not the key (pswf_12D.cu:287)
    assert not host.factory_has("Fi", "nope")


def test_command_line_defaults_and_flags():
    d = host.parse_args("-i in.ms -o out.ms -z 0.001")
    # defaults of getOptions, src/functions.cu:185-262
    assert d["ok"] and d["noise_cut"] == 10 and d["eta"] == -1 and d["robust_param"] == 2 and d["it_max"] == 500
    assert d["gpus"] == "0" and d["modin"] == "mod_in_0.fits" and d["output_image"] == "mod_out.fits" and d["gridding"] == 0
    d = host.parse_args("--input a,b --output c,d -m hdr -O img -z 0.001,2.5 -Z 0.01,0,1e-4 -t 50 -R -0.5 -g 8 "
                        "-G 0,1 -N 3.5 -e -2 -F 2.3e11 -T 3 -p out/ -f stats.txt -X 16 -Y 16 -V 256 -v -x -P -W")
    assert d["input"] == "a,b" and d["output"] == "c,d" and d["modin"] == "hdr" and d["output_image"] == "img"
    assert d["initial_values"] == "0.001,2.5" and d["penalization_factors"] == "0.01,0,1e-4"
    assert d["it_max"] == 50 and d["robust_param"] == -0.5 and d["gridding"] == 8 and d["gpus"] == "0,1"
    assert d["noise_cut"] == 3.5 and d["eta"] == -2 and abs(d["nu_0"] - 2.3e11) < 1e5 and d["threshold"] == 3
    assert d["blockSizeX"] == 16 and d["blockSizeV"] == 256
    assert d["verbose"] == 1 and d["nopositivity"] == 1 and d["print_images"] == 1 and d["modify_weights"] == 1
    # the reference prints the help and exits on -h, a negative -g or -r outside [0,1]
    for bad in ("-h", "-g -1 -z 1", "-r 1.5 -z 1", "--no-such-flag"):
        assert not host.parse_args(bad)["ok"], bad


@pytest.fixture(scope="module")
def prob():
    return synth.make_problem(N=128, nvis=6000, nchan=2, freq0=1.0e11, bandwidth=8e9, seed=7, grid_fill=1.15)


def _sigmas(p):
    du, dv = 1.0 / (p.M * RPDEG_D * p.DELTAX), 1.0 / (p.N * RPDEG_D * p.DELTAY)
    return np.float32(abs(du)), np.float32(abs(dv))


CK = [("PillBox2D", 1, 1), ("Gaussian2D", 7, 7), ("GaussianSinc2D", 7, 7), ("Sinc2D", 7, 7), ("PSWF", 9, 9),
      ("Gaussian2D", 5, 5), ("PSWF", 7, 7)]


@pytest.mark.parametrize("name,m,n", CK)
def test_ckernel_tables_bit_exact_vs_oracle(oracle, prob, name, m, n):
    sx, sy = _sigmas(prob)
    mine, support = host.ckernel_table(name, m, n, sx, sy)
    want = oracle.ckernel(name, m, n, sx, sy)
    assert np.array_equal(mine.view(np.uint32), want.view(np.uint32))
    assert support == (m // 2, m // 2)   # both supports come from m (ckernel.cuh:508-511)
    dx, dy = np.float32(abs(RPDEG_D * prob.DELTAX)), np.float32(abs(RPDEG_D * prob.DELTAY))
    gcf = host.ckernel_gcf(name, m, n, prob.M, prob.N, dx, dy)
    want_gcf = oracle.ckernel(name, prob.M, prob.N, dx, dy, w=float(prob.M), gcf=True)
    assert np.array_equal(gcf.view(np.uint32), want_gcf.view(np.uint32))


@pytest.mark.parametrize("name,m,n", CK[:6])
def test_ckernel_tables_bit_exact_vs_reference(gvref, prob, name, m, n):
    gvref.set_problem(prob)
    table, support, gcf = gvref.cpu_ckernel(name, m, n, want_gcf=True)
    sx, sy = _sigmas(prob)
    mine, sup = host.ckernel_table(name, table.shape[0], table.shape[1], sx, sy)
    assert np.array_equal(table.view(np.uint32), mine.view(np.uint32))
    assert sup == support
    if name in ("GaussianSinc2D", "Sinc2D"):
        return  # the reference never builds a GCF image for these (see test_oracle_vs_reference_cpu.py)
    dx, dy = np.float32(abs(RPDEG_D * prob.DELTAX)), np.float32(abs(RPDEG_D * prob.DELTAY))
    mine_gcf = host.ckernel_gcf(name, m, n, prob.M, prob.N, dx, dy)
    assert np.array_equal(gcf.view(np.uint32), mine_gcf.view(np.uint32))


def test_gvms_container_round_trip(tmp_path, prob):
    path = str(tmp_path / "p.gvms")
    synth.write_gvms(prob, path)
    d = host.read_gvms(path)
    assert d["M"] == prob.M and d["N"] == prob.N and d["nchan"] == prob.nchan and d["total_vis"] == prob.total_vis()
    assert np.float32(d["min_freq"]) == prob.freqs.min() and np.float32(d["max_freq"]) == prob.freqs.max()
    bl = np.sqrt(prob.uvw[0][:, 0] ** 2 + prob.uvw[0][:, 1] ** 2).max()
    assert abs(d["max_blength"] - bl) <= 1e-6 * bl
    with open(path, "r+b") as f:
        f.write(b"XXXX")
    with pytest.raises(RuntimeError):
        host.read_gvms(path)


FUNCS = [
    lambda x: (x - 3.3) ** 2 + 1.0,                    # minimum beyond the first bracket
    lambda x: (x + 0.4) ** 2,                          # minimum behind the start
    lambda x: math.cosh(0.3 * x - 2.0),                # slow growth: parabolic steps hit the limit
    lambda x: abs(x - 0.25) + 0.1 * (x - 0.25) ** 2,   # kink: golden-section steps
    lambda x: 1e6 * (x - 1e-3) ** 2 + 5.0,             # very narrow
    lambda x: (x - 250.0) ** 2 * 1e-4,                 # far away (GLIMIT branch)
    lambda x: float(np.float32(x) ** 4 - 3 * np.float32(x) ** 3 + 2),
]


@pytest.mark.parametrize("k", range(len(FUNCS)))
def test_line_search_follows_the_oracle_probe_for_probe(oracle, k):
    f = FUNCS[k]
    seen_a, seen_b = [], []

    def fa(x):
        seen_a.append(np.float32(x))
        return float(np.float32(f(float(np.float32(x)))))

    def fb(x):
        seen_b.append(np.float32(x))
        return float(np.float32(f(float(np.float32(x)))))

    xm, fm, n = host.linmin_1d(fa)
    FN = ctypes.CFUNCTYPE(ctypes.c_float, ctypes.c_float, ctypes.c_void_p)
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    lib.gvo_linmin_1d.argtypes = [FN, ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float),
                                  ctypes.POINTER(ctypes.c_int)]
    oxm, ofm, on = ctypes.c_float(), ctypes.c_float(), ctypes.c_int()
    lib.gvo_linmin_1d(FN(lambda x, _u: fb(x)), None, ctypes.byref(oxm), ctypes.byref(ofm), ctypes.byref(on))
    assert n == on.value == len(seen_a) == len(seen_b)
    assert np.array_equal(np.array(seen_a).view(np.uint32), np.array(seen_b).view(np.uint32))
    assert np.float32(xm) == np.float32(oxm.value) and np.float32(fm) == np.float32(ofm.value)
    # and it is a minimum
    assert f(xm) <= min(f(xm * (1 + 1e-3) + 1e-4), f(xm * (1 - 1e-3) - 1e-4)) + 1e-6 * abs(f(xm))


@pytest.mark.parametrize("k", range(len(FUNCS)))
def test_line_search_matches_the_reference_binary(gvref, k):
    """The reference's OWN mnbrak + brent (host functions inside oracle/_ref/libgvref.so, called
    through their C++ symbols exactly as linmin does, src/linmin.cu:78-84) against the host
    layer's LineSearch: same probes, same minimum."""
    f = FUNCS[k]
    ref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libgvref.so"))
    F1 = ctypes.CFUNCTYPE(ctypes.c_float, ctypes.c_float)
    PF = ctypes.POINTER(ctypes.c_float)
    mnbrak = getattr(ref, "_Z6mnbrakPfS_S_S_S_S_PFffE")
    mnbrak.argtypes = [PF, PF, PF, PF, PF, PF, F1]
    mnbrak.restype = None
    brent = getattr(ref, "_Z5brentffffPfPFffE")
    brent.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, PF, F1]
    brent.restype = ctypes.c_float
    seen_r, seen_h = [], []

    def fr(x):
        seen_r.append(np.float32(x))
        return float(np.float32(f(float(np.float32(x)))))

    def fh(x):
        seen_h.append(np.float32(x))
        return float(np.float32(f(float(np.float32(x)))))

    cb = F1(fr)
    ax, xx, bx = ctypes.c_float(0.0), ctypes.c_float(1.0), ctypes.c_float()
    fa, fx, fb = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
    mnbrak(ctypes.byref(ax), ctypes.byref(xx), ctypes.byref(bx), ctypes.byref(fa), ctypes.byref(fx), ctypes.byref(fb), cb)
    xmin = ctypes.c_float()
    fret = brent(ax, xx, bx, ctypes.c_float(1.0e-7), ctypes.byref(xmin), cb)
    xm, fm, n = host.linmin_1d(fh)
    assert n == len(seen_r)
    assert np.array_equal(np.array(seen_r).view(np.uint32), np.array(seen_h).view(np.uint32))
    assert np.float32(xm) == np.float32(xmin.value) and np.float32(fm) == np.float32(fret)


def test_session_fails_loudly_without_a_gpu(prob):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA device"):
        host.Session(prob, args="-z 0.001 -Z 0.01 -t 2")


def test_libraries_do_not_leak_runtime_symbols():
    """Both libraries export only their own API: a second copy of C++ runtime symbols in the
    process (some toolchains link libstdc++ statically) breaks iostream in the host program."""
    import subprocess
    from gpuvmem_b200 import lib
    for path, allowed in ((lib.lib_path(), ("gvm_",)), (host.host_lib_path(), ("gvmh_", "gpuvmem::"))):
        out = subprocess.run(["nm", "-DC", "--defined-only", path], capture_output=True, text=True).stdout
        leaked = [l for l in out.splitlines() if l and not any(a in l for a in allowed)]
        assert not leaked, leaked[:5]


def test_shard_plan_covers_every_visibility_exactly_once():
    # >= world channels: whole channels, rank = chan % world (the reference's rule)
    Z = [1000 + 7 * c for c in range(64)]
    for world in (2, 4, 8):
        owners = np.zeros(64, int)
        for r in range(world):
            for c, (lo, hi) in enumerate(host.shard_plan(Z, world, r)):
                if hi > lo:
                    assert (lo, hi) == (0, Z[c]) and c % world == r
                    owners[c] += 1
        assert (owners == 1).all()
    # fewer channels than ranks: contiguous visibility chunks
    for Zs, world in [([10_000_000], 8), ([1_048_576, 999_999], 4), ([7], 2), ([5], 8)]:
        for c, z in enumerate(Zs):
            covered = np.zeros(z, int)
            for r in range(world):
                lo, hi = host.shard_plan(Zs, world, r)[c]
                covered[lo:hi] += 1
            assert (covered == 1).all()
    assert host.shard_plan([5, 6], 1, 0) == [(0, 5), (0, 6)]


def test_option_parser_is_reentrant():
    """getOptions is called once per Synthesizer::configure; glibc's getopt keeps a pointer into the previous
    argv unless it is fully re-initialised (optind = 0) — a trailing boolean flag used to poison the next parse."""
    for _ in range(3):
        a = host.parse_args("-z 0.001 -Z 0.01 -t 3 -M")
        b = host.parse_args("-z 0.002,0.5 -Z 0.01,0.005 -t 6 -e -0.5 -x")
        c = host.parse_args("-x -v -P")
        assert a["ok"] and a["it_max"] == 3 and a["initial_values"] == "0.001"
        assert b["ok"] and b["it_max"] == 6 and b["initial_values"] == "0.002,0.5" and b["eta"] == -0.5 and b["nopositivity"] == 1
        assert c["ok"] and c["nopositivity"] == 1 and c["verbose"] == 1 and c["print_images"] == 1 and c["it_max"] == 500
