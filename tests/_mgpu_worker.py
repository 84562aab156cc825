"""One rank of a multi-GPU parity run (launched by torchrun from tests/test_multi_gpu.py):
the same scenario through the C++ host layer with the visibilities sharded over the ranks and
the engine's NCCL all-reduces; rank 0 stores objective, gradient and the image after the
optimizer ran."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from _ref_runner import probe_image
    from gpuvmem_b200 import dist as gdist
    from gpuvmem_b200 import host, synth
    nchan, out = int(sys.argv[1]), sys.argv[2]
    mode = sys.argv[3] if len(sys.argv) > 3 else ""
    normalize = mode == "normalize"
    rank, world, local = gdist.init_from_env(0)
    torch.cuda.set_device(local)
    nccl_id = None
    if world > 1:
        box = [host.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]
    p = synth.make_problem(N=128, nvis=15000, nchan=nchan, freq0=1.0e11, bandwidth=4e9 if nchan > 1 else 0.0,
                           seed=41, grid_fill=0.9)
    host.set_quiet(True)
    # normalize: Chi2 configured with normalize = true (chi2 and its gradient divided by the block's visibility count)
    fi_spec = "Chi2:-1:0:0:1,Entropy:0:0:0,L1-Norm:1:0:0,TotalSquaredVariation:2:0:0" if normalize else None
    # gridded_*: weighting scheme + convolutional gridding (-g) are DISTRIBUTED over the ranks (gvm_weights_dist,
    # gvm_grid_block_dist) and must reproduce the single-rank samples bit for bit
    extra, scheme, ck, ck_size = "", "Natural", "PillBox2D", (0, 0)
    if mode == "gridded_briggs":
        extra, scheme, ck, ck_size = " -g 1 -R 0.0", "Briggs", "Gaussian2D", (7, 7)
    elif mode == "gridded_uniform":
        extra, scheme, ck, ck_size = " -g 1", "Uniform", "PSWF", (9, 9)
    elif mode == "radial":
        scheme = "Radial"
    s = host.Session(p, args=f"-z 0.001,0.1 -Z 0.01,0.005,0.002 -t 4 -G {local}" + extra, optimizer="CG-FRPRMN",
                     scheme=scheme, ckernel=ck, ck_size=ck_size, fi_spec=fi_spec, rank=rank, world=world, nccl_id=nccl_id)
    vis = {}
    for c in range(p.nchan):
        u, V, wt = s.host_vis(c)
        vis[f"uvw{c}"], vis[f"Vo{c}"], vis[f"w{c}"] = u, V, wt
    start = s.get_image()
    s.set_image(probe_image(p.N, np.float32(0.001), 0.1))
    s.set_iteration(1)
    v, fi = s.calc_function()
    g = s.calc_gradient(1)
    err = s.error_image()      # SecondDerivateError on the residuals of that evaluation
    s.set_image(start)
    s.set_iteration(0)
    img, sec = s.run()
    if rank == 0:
        np.savez(out, value=v, fi=fi, grad=g, image=img, err=err, local_nvis=s.local_nvis(), collectives=s.collectives(),
                 world=world, **vis)
    s.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
