"""Pins the CPU oracle (oracle/gvm_oracle.c) against the reference's OWN host code
(WeightingScheme::apply, do_gridding, CKernel tables) compiled unmodified into
oracle/_ref/libgvref.so. Bit-for-bit, single-thread order.
Runs without a GPU."""
import numpy as np
import pytest

from gpuvmem_b200 import synth
from gpuvmem_b200.engine import RPDEG_D, WEIGHTING


def _deltas(p):
    return 1.0 / (p.M * RPDEG_D * p.DELTAX), 1.0 / (p.N * RPDEG_D * p.DELTAY)


@pytest.fixture(scope="module")
def prob():
    # 2 channels so that Briggs' never-cleared first-pass grid matters; grid_fill > 1 puts
    # some samples outside the grid (weight -> 0 edge case)
    return synth.make_problem(N=128, nvis=6000, nchan=2, freq0=1.0e11, bandwidth=8e9, seed=7, grid_fill=1.15)


@pytest.mark.parametrize("scheme,robust", [("Natural", 0.0), ("Uniform", 0.0), ("Briggs", 0.0),
                                            ("Briggs", -2.0), ("Briggs", 2.0), ("Radial", 0.0)])
def test_weights_bit_exact(oracle, gvref, prob, scheme, robust):
    gvref.set_problem(prob)
    ref = gvref.cpu_weights(scheme, robust, threads=1)
    du, dv = _deltas(prob)
    mine = oracle.weights(WEIGHTING[scheme], robust, prob.M, prob.N, du, dv, prob.uvw, prob.freqs, prob.w)
    for c in range(prob.nchan):
        assert np.array_equal(ref[c].view(np.uint32), mine[c].view(np.uint32)), (scheme, c)
    if scheme in ("Uniform", "Briggs"):
        assert any((r == 0).any() for r in ref), "edge case (out-of-grid sample) not exercised"


@pytest.mark.parametrize("name,m,n", [("PillBox2D", 1, 1), ("Gaussian2D", 7, 7), ("GaussianSinc2D", 7, 7),
                                      ("Sinc2D", 7, 7), ("PSWF", 9, 9), ("Gaussian2D", 5, 5)])
def test_ckernel_tables_bit_exact(oracle, gvref, prob, name, m, n):
    gvref.set_problem(prob)
    table, support, gcf = gvref.cpu_ckernel(name, m, n, want_gcf=True)
    du, dv = _deltas(prob)
    sx, sy = np.float32(abs(du)), np.float32(abs(dv))
    mine = oracle.ckernel(name, table.shape[0], table.shape[1], sx, sy)
    assert np.array_equal(table.view(np.uint32), mine.view(np.uint32))
    assert support == (table.shape[0] // 2, table.shape[0] // 2)
    if name in ("GaussianSinc2D", "Sinc2D"):
        # Reference quirk: these classes never override buildGCF (only PillBox2D, Gaussian2D and
        # PSWF_12D do), so initializeGCF leaves a 7x7 clone behind and the "GCF image" the
        # reference would read is out-of-bounds memory. The engine defines it as GCF() == 1.
        return
    dx, dy = np.float32(abs(RPDEG_D * prob.DELTAX)), np.float32(abs(RPDEG_D * prob.DELTAY))
    mine_gcf = oracle.ckernel(name, prob.M, prob.N, dx, dy, w=float(prob.M), gcf=True)
    assert np.array_equal(gcf.view(np.uint32), mine_gcf.view(np.uint32))


@pytest.mark.parametrize("name,m,n", [("PillBox2D", 1, 1), ("Gaussian2D", 7, 7), ("GaussianSinc2D", 7, 7),
                                      ("PSWF", 9, 9)])
def test_gridding_bit_exact(oracle, gvref, prob, name, m, n):
    gvref.set_problem(prob)
    ref = gvref.cpu_gridding(name, m, n, threads=1)
    table, support, _ = gvref.cpu_ckernel(name, m, n)
    du, dv = _deltas(prob)
    for c in range(prob.nchan):
        u, v, w = oracle.gridding(prob.M, prob.N, du, dv, float(prob.freqs[c]), prob.uvw[c], prob.Vo[c],
                                  prob.w[c], table, support)
        ru, rv, rw = ref[c]
        assert len(w) == len(rw) > 0
        assert np.array_equal(u.view(np.uint64), ru.view(np.uint64))
        assert np.array_equal(v.view(np.uint32), rv.view(np.uint32))
        assert np.array_equal(w.view(np.uint32), rw.view(np.uint32))


def test_weight_cell_index_edge_cases(oracle, prob):
    du, dv = _deltas(prob)
    lam = float(np.float32(2.99792458e8) / prob.freqs[0])
    # samples exactly on cell borders, on the grid edge, and mirrored pairs
    g = np.array([[0.0, 0.0], [0.5, -0.5], [-0.5, 0.5], [prob.N / 2 - 0.5, 0.0], [prob.N / 2 + 0.49, 1.0],
                  [-(prob.N / 2), -(prob.M / 2)], [3.25, -7.75], [-3.25, 7.75]])
    uvw = np.zeros((len(g), 3))
    uvw[:, 0] = g[:, 0] * abs(du) * lam
    uvw[:, 1] = g[:, 1] * abs(dv) * lam
    cells = oracle.weight_cells(uvw, float(prob.freqs[0]), du, dv, prob.M, prob.N)
    assert cells[6] == cells[7], "Hermitian twins must land in the same weighting cell"
    assert cells[0] == prob.N * (prob.M // 2) + prob.N // 2
    assert (cells >= -1).all() and (cells < prob.M * prob.N).all()
