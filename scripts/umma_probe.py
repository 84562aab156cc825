"""Probe of the tensor-core gradient kernel: accuracy vs the CUDA-core separable kernel and
kernel time, for several TMEM chunk lengths (env GVM_UMMA_CHUNK). Run on the GPU box."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpuvmem_b200 import Engine, synth  # noqa: E402
from gpuvmem_b200.engine import GRAD_SIMT, GRAD_UMMA  # noqa: E402

N = int(os.environ.get("PROBE_N", "2048"))
Z = int(os.environ.get("PROBE_Z", "1000000"))
chunks = [int(x) for x in os.environ.get("PROBE_CHUNKS", "512,2048,8192,32768").split(",")]
p = synth.make_problem(N=N, nvis=Z, nchan=1, seed=3, bmax=float(os.environ.get("PROBE_BMAX", "1000")),
                       bmin=float(os.environ.get("PROBE_BMIN", "15")))
e = Engine.from_problem(p, grad_mode=GRAD_SIMT, noise_cut=float(os.environ.get("PROBE_NOISE_CUT", "10")))
e.use_torch_stream()
I_dev = torch.from_numpy(e.initial_image()).cuda()
I_dev[0] *= 1.0 + torch.rand(N, N, device="cuda")
chi2 = e.chi2(I_dev)
ref = torch.zeros_like(I_dev)
e.dchi2(I_dev, ref, 0)
torch.cuda.synchronize()
ms_simt, _ = e.last_grad_kernel_ms()
print(f"N={N} Z={Z} chi2={chi2:.6e} SIMT kernel {ms_simt:.2f} ms "
      f"({4.0*N*N*Z/ms_simt/1e9:.1f} TFLOP/s algorithmic)", flush=True)
# fp64 truth at sampled pixels (CPU oracle; test infrastructure)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _checkers import Oracle  # noqa: E402
o = Oracle()
cfg = dict(D=p.antenna_diameter, DELTAX=p.DELTAX, DELTAY=p.DELTAY, eta=-1.0)
npix = int(os.environ.get("PROBE_NPIX", "400"))
pix = np.linspace(0, N * N - 1, npix).astype(np.int64)
v = e.get_vis(0, want=("uvw", "Vr", "w"))
Ic = I_dev.cpu().numpy()
t0 = time.time()
truth = o.dchi2(pix, N, v["uvw"], v["Vr"], v["w"], e.get_noise_image(), None, float(p.freqs[0]), e.meta, cfg)
truth = truth * o.chain(Ic, pix, float(p.freqs[0]), e.meta, e.cfg.threshold, 0)
print(f"oracle: {npix} pixels in {time.time()-t0:.1f} s", flush=True)
def err64(g):
    got = g[0].reshape(-1)[torch.from_numpy(pix).cuda()].double().cpu().numpy()
    return float(np.linalg.norm(got - truth) / np.linalg.norm(truth))
print(f"SIMT rel-L2 vs fp64 oracle: {err64(ref):.3e}", flush=True)
e.set_grad_mode(GRAD_UMMA)
for ch in chunks:
    os.environ["GVM_UMMA_CHUNK"] = str(ch)
    for rep in range(2):
        g = torch.zeros_like(I_dev)
        e.dchi2(I_dev, g, 0)
        torch.cuda.synchronize()
    ms, n = e.last_grad_kernel_ms()
    err = float((g[0] - ref[0]).norm() / ref[0].norm())
    nt, px = e.grad_plan()
    print(f"chunk={ch:6d}: UMMA kernel {ms:.3f} ms, plan {nt} tiles / {px} px of {N*N} "
          f"({4.0*px*Z/ms/1e9:.1f} TFLOP/s algorithmic on computed pixels; {N*N*Z/ms/1e9:.2f} Mvis*Mpix/s whole image) "
          f"rel-L2 vs SIMT {err:.3e} vs fp64 {err64(g):.3e}", flush=True)
e.close()
