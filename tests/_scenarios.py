"""Reconstruction scenarios shared by the reference runner (tests/_ref_runner.py, one fresh
process per scenario because the reference keeps its state in process globals) and the host
layer's GPU parity tests (tests/test_host_gpu.py)."""
from gpuvmem_b200 import synth

LAMBDAS = "0.01,0.005,0.002,0.001"   # -Z: Entropy, L1-Norm, TSV, Laplacian (src/main.cu:193-197)

SCENARIOS = {
    # name: (problem kwargs, reference command line, optimizer, scheme, ckernel, (m, n), lbfgs K)
    "cg_natural": (dict(N=128, nvis=20000, nchan=1, freq0=2.3e11, seed=31, grid_fill=0.9),
                   f"-z 0.001 -Z {LAMBDAS} -t 6", "CG-FRPRMN", "Natural", "PillBox2D", (1, 1), 0),
    "lbfgs_natural": (dict(N=128, nvis=20000, nchan=1, freq0=2.3e11, seed=32, grid_fill=0.9),
                      f"-z 0.001 -Z {LAMBDAS} -t 6", "CG-LBFGS", "Natural", "PillBox2D", (1, 1), 4),
    "cg_mfs_briggs": (dict(N=128, nvis=12000, nchan=3, freq0=1.0e11, bandwidth=6e9, seed=33, grid_fill=0.9),
                      "-z 0.001,0.2 -Z 0.01,0.0,0.002 -R 0.5 -t 4", "CG-FRPRMN", "Briggs", "PillBox2D", (1, 1), 0),
    "cg_gridded_gaussian": (dict(N=128, nvis=20000, nchan=1, freq0=2.3e11, seed=34, grid_fill=0.9),
                            f"-z 0.001 -Z {LAMBDAS} -g 1 -R 0.0 -t 4", "CG-FRPRMN", "Briggs", "Gaussian2D", (7, 7), 0),
    "cg_gridded_pswf": (dict(N=128, nvis=20000, nchan=1, freq0=2.3e11, seed=35, grid_fill=0.9),
                        "-z 0.001 -Z 0.01 -g 1 -t 3", "CG-FRPRMN", "Uniform", "PSWF", (9, 9), 0),
    # -x: no positivity projection (so no entropy term: ln of a negative pixel is NaN in the reference too);
    # eta != -1 (MINPIX = -eta * z, the value masked pixels are pinned to); a tighter mask (-N 5)
    "cg_nopositivity_eta": (dict(N=128, nvis=16000, nchan=1, freq0=2.3e11, seed=36, grid_fill=0.9),
                            "-z 0.002 -Z 0.0,0.005,0.002,0.001 -x -e -0.5 -N 5 -t 4", "CG-FRPRMN", "Natural", "PillBox2D", (1, 1), 0),
    # MFS with a spectral-index threshold (-T sigma: alpha frozen where I0 < 5 T) and radial weighting, L-BFGS
    "lbfgs_mfs_threshold_radial": (dict(N=128, nvis=12000, nchan=3, freq0=1.0e11, bandwidth=6e9, seed=37, grid_fill=0.9),
                                   "-z 0.001,0.1 -Z 0.01,0.0,0.002 -T 0.001 -t 4", "CG-LBFGS", "Radial", "PillBox2D", (1, 1), 3),
    # the field's phase / pointing centre is NOT the image centre: direccos -> phs_xobs_pix / ref_xobs_pix
    # (src/mfs.cu:660-691) off the central pixel, phase_rotate and the beam follow it
    "cg_offset_field": (dict(N=128, nvis=16000, nchan=1, freq0=2.3e11, seed=38, grid_fill=0.9),
                        f"-z 0.001 -Z {LAMBDAS} -t 4", "CG-FRPRMN", "Natural", "PillBox2D", (1, 1), 0),
}
FIELD_OFFSET_PIX = {"cg_offset_field": (7.3, -4.6)}   # field centre relative to the image centre, in pixels
REF_EXTRA = " -X 16 -Y 16 -V 256 -i synth.ms -o out.ms -m hdr.fits"


def problem(name):
    p = synth.make_problem(**SCENARIOS[name][0])
    if name in FIELD_OFFSET_PIX:
        import math
        dx, dy = FIELD_OFFSET_PIX[name]
        p.field_centre = (p.ra + dx * p.DELTAX / math.cos(math.radians(p.dec)), p.dec + dy * p.DELTAY)
    return p
