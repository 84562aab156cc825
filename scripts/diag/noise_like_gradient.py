"""Accuracy of the tensor-core gradient when the residuals are pure noise (every pixel's sum is a random walk, the regime
near convergence) against a coherent sky, as a function of the TMEM accumulation chunk (GVM_UMMA_CHUNK): rel-L2 against
the fp64 oracle at sampled pixels. 4 M visibilities on a 256^2 image: one tile, 592 K slices of 6757 visibilities."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from _checkers import Oracle
from gpuvmem_b200 import Engine, synth
from gpuvmem_b200.engine import GRAD_UMMA

o = Oracle()
nvis = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
for kind in ("coherent", "noise"):
    p = synth.make_problem(N=256, nvis=nvis, nchan=1, seed=23)
    if kind == "noise":
        rng = np.random.default_rng(1)
        p.Vo[0] = (100.0 * rng.standard_normal(p.Vo[0].shape)).astype(np.float32)
    e = Engine.from_problem(p, grad_mode=GRAD_UMMA)
    I = torch.from_numpy(e.initial_image()).cuda()
    cfg = dict(D=p.antenna_diameter, DELTAX=p.DELTAX, DELTAY=p.DELTAY, eta=-1.0)
    e.chi2(I)
    v = e.get_vis(0, want=("uvw", "Vr", "w"))
    noise = e.get_noise_image()
    pix = np.flatnonzero(noise.reshape(-1) < e.meta["noise_cut"])[::97][:160]
    want = o.dchi2(pix, p.N, v["uvw"], v["Vr"], v["w"], noise, None, float(p.freqs[0]), e.meta, cfg)
    want = want * o.chain(I.cpu().numpy(), pix, float(p.freqs[0]), e.meta, e.cfg.threshold, 0)
    for chunk in (1024, 2048, 4096, 8192):
        os.environ["GVM_UMMA_CHUNK"] = str(chunk)
        g = torch.zeros_like(I)
        e.dchi2(I, g, flag_opt=0)
        torch.cuda.synchronize()
        got = g.cpu().numpy()[0].reshape(-1)[pix]
        print(f"{kind:9s} chunk {chunk:5d}: rel-L2 vs fp64 oracle {np.linalg.norm(got - want) / np.linalg.norm(want):.3e}", flush=True)
    e.close()
