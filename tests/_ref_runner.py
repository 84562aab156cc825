"""Runs ONE scenario through the reference itself (oracle/_ref/libgvref.so: gpuvmem's own
sources, unmodified, compiled for sm_100a) in a fresh process and stores what it computed:
setup scalars, the visibilities after weighting/gridding, objective + gradient at a probe
image, and the image after the optimizer ran. TEST INFRASTRUCTURE (GPU box only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def probe_image(N, minpix, alpha0, seed=3):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:N, 0:N]
    blob = np.exp(-((xx - N * 0.55) ** 2 + (yy - N * 0.45) ** 2) / (2 * (N / 16) ** 2))
    I = np.empty((2, N, N), np.float32)
    I[0] = (minpix * (1.0 + 40.0 * blob + 0.2 * rng.random((N, N)))).astype(np.float32)
    I[1] = (alpha0 + 0.3 * blob + 0.05 * rng.standard_normal((N, N))).astype(np.float32)
    return I


def main():
    from _checkers import GvRef
    from _scenarios import REF_EXTRA, SCENARIOS, problem
    name, out = sys.argv[1], sys.argv[2]
    kw, args, optimizer, scheme, ck, (m, n), K = SCENARIOS[name]
    p = problem(name)
    ref = GvRef()
    ref.set_problem(p)
    ref.lib.gvref_set_verbose(0)
    ref.init(args + REF_EXTRA, optimizer=optimizer, scheme=scheme, ckernel=ck, ck_m=m, ck_n=n)
    if K:
        ref.lib.gvref_set_lbfgs_k(K)
    sc = ref.scalars()
    res = {f"s_{k}": v for k, v in sc.items()}
    for c in range(p.nchan):
        hv = ref.get_host_vis(c)
        res[f"uvw{c}"], res[f"Vo{c}"], res[f"w{c}"] = hv["uvw"], hv["Vo"], hv["w"]
    I0 = ref.get_image()
    res["I_start"] = I0
    z = [float(t) for t in args.split("-z")[1].split()[0].split(",")]
    probe = probe_image(p.N, np.float32(z[0]), z[1] if len(z) > 1 else 0.0)
    ref.set_image(probe)
    v, fi = ref.calc_function(iteration=1)
    res["probe_value"], res["probe_fi"] = v, fi
    res["probe_grad"] = ref.calc_gradient(iteration=1, flag=0)
    res["probe_image_after"] = ref.get_image()   # the clip mutates the image
    if len(z) > 1:                               # spectral-index gradient (flag_opt odd, DChi2_total_alpha + threshold)
        res["probe_grad_flag1"] = ref.calc_gradient(iteration=1, flag=1)
    if hasattr(ref.lib, "gvref_error_image"):
        res["probe_err"] = ref.error_image()     # calculateErrors on the probe image's residuals
    ref.set_image(I0)
    img, iters, ms = ref.run()
    res["final_image"], res["iterations"], res["run_ms"] = img, iters, ms
    v, fi = ref.calc_function(iteration=max(iters, 1))
    res["final_value"], res["final_fi"] = v, fi
    # MFS::writeResiduals (src/mfs.cu:1115-1155) on the final image: what would go to the output Measurement Set
    if hasattr(ref.lib, "gvref_write_residuals"):
        wb_chi2, blocks = ref.write_residuals()
        res["wb_chi2"] = np.float32(wb_chi2)
        for c, b in enumerate(blocks):
            res[f"wb_uvw{c}"], res[f"wb_Vo{c}"], res[f"wb_Vm{c}"], res[f"wb_w{c}"] = b["uvw"], b["Vo"], b["Vm"], b["w"]
    np.savez(out, **res)
    print("reference scenario", name, "iterations", iters, "value", v, "ms", ms)


if __name__ == "__main__":
    main()
