"""CPU tests of the oracle itself: the reference-derived known answer, cell-list == all-pairs sets,
and the committed golden vectors (tests/golden/, made by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from prestige_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_eq1_known_answer():
    """prestige/src/lib.rs:7-12 through the loop of simple_cpu.rs:7-16: force[i] = sum_j mass[j], self included."""
    mass = np.arange(1.0, 11.0)
    assert np.array_equal(orc.eq1_allpairs(mass, np.zeros(10)), np.full(10, 55.0))
    assert np.array_equal(orc.eq1_allpairs(mass, np.ones(10)), np.full(10, 56.0))      # += keeps the initial value
    assert np.array_equal(orc.eq1_allpairs(mass.astype(np.float32), np.zeros(10, np.float32)), np.full(10, 55.0, np.float32))


@pytest.mark.parametrize("real", [np.float64, np.float32])
def test_cell_list_reproduces_allpairs_sets(real):
    b = synth.wcsph_block_3d(14, 12, 13).shuffled().astype(real)
    a = b.arrays
    g = orc.make_grid(3, b.lo, b.hi, b.cell_size)
    pa, margin = orc.pairs(3, a["x"], a["y"], a["z"], a["h"])
    pc, _ = orc.pairs(3, a["x"], a["y"], a["z"], a["h"], grid=g)
    assert np.array_equal(pa, pc)
    if real == np.float64:
        assert margin > 1e-12, "a pair sits on the cutoff: the synthetic jitter must prevent that"
    d = synth.dem_column_3d(9).shuffled().astype(real)
    a = d.arrays
    g = orc.make_grid(3, d.lo, d.hi, d.cell_size)
    pa, _ = orc.pairs(3, a["x"], a["y"], a["z"], a["rad"], mode=1)
    pc, _ = orc.pairs(3, a["x"], a["y"], a["z"], a["rad"], mode=1, grid=g)
    assert np.array_equal(pa, pc)


def test_wcsph_cells_match_allpairs():
    for b in (synth.wcsph_block_3d(12, 11, 10).shuffled(), synth.wcsph_dambreak_2d(dx=0.04).shuffled()):
        g = orc.make_grid(b.dim, b.lo, b.hi, b.cell_size)
        ra = orc.wcsph(b.dim, b.params, b.arrays)
        for mode in (False, True):
            rc = orc.wcsph(b.dim, b.params, b.arrays, grid=g, sorted_step=mode)
            for k in ra:
                scale = max(np.abs(ra[k]).max(), 1e-300)
                assert np.abs(ra[k] - rc[k]).max() <= 1e-12 * scale, k


def test_wcsph_momentum_antisymmetry():
    b = synth.wcsph_block_3d(10, 10, 10)
    b.params["gz"] = 0.0
    r = orc.wcsph(3, b.params, b.arrays)
    m = b.arrays["m"]
    for k in ("au", "av", "aw"):
        assert abs((m * r[k]).sum()) <= 1e-12 * (m * np.abs(r[k])).sum()


def test_dem_history_and_antisymmetry():
    d = synth.dem_column_3d(8).shuffled()
    g = orc.make_grid(3, d.lo, d.hi, d.cell_size)
    hist = None
    for _ in range(3):
        fa, ha, ov = orc.dem(d.params, 12, d.arrays, hist=hist)
        fc, hc, _ = orc.dem(d.params, 12, d.arrays, hist=hist, grid=g)
        assert ov == 0
        da, dc = orc.history_as_dict(ha), orc.history_as_dict(hc)
        assert da.keys() == dc.keys()
        for k in da:
            assert np.allclose(da[k], dc[k], rtol=0, atol=1e-20)
            assert tuple(-np.array(da[(k[1], k[0])])) == tuple(np.array(da[k])), "xi_ji = -xi_ij bit-exactly"
        for k in fa:
            assert np.abs(fa[k] - fc[k]).max() <= 1e-12 * np.abs(fa[k]).max()
        hist = ha
    assert ha["hist_n"].max() == 6
    _, _, ov = orc.dem(d.params, 3, d.arrays)
    assert ov == 1                                  # more contacts than slots is reported, never truncated silently


def test_golden_vectors():
    """Committed fixtures pin the oracle (and, on the GPU box, the CUDA path) against drift."""
    z = np.load(os.path.join(GOLD, "wcsph3d_small.npz"))
    b = synth.wcsph_block_3d(9, 8, 7).shuffled()
    assert np.array_equal(z["x"], b.arrays["x"]), "synthetic generator changed"
    r = orc.wcsph(3, b.params, b.arrays)
    for k in ("p", "au", "av", "aw", "arho"):
        assert np.array_equal(z[k], r[k]), k
    pr, _ = orc.pairs(3, b.arrays["x"], b.arrays["y"], b.arrays["z"], b.arrays["h"])
    assert np.array_equal(z["pairs"], pr)
    z = np.load(os.path.join(GOLD, "dem3d_small.npz"))
    d = synth.dem_column_3d(6).shuffled()
    f1, h1, _ = orc.dem(d.params, 12, d.arrays)
    f2, h2, _ = orc.dem(d.params, 12, d.arrays, hist=h1)
    for k in ("fx", "fy", "fz", "tx", "ty", "tz"):
        assert np.array_equal(z[k], f2[k]), k
    assert np.array_equal(z["hist_n"], h2["hist_n"])
    z = np.load(os.path.join(GOLD, "wcsph2d_small.npz"))
    c = synth.wcsph_dambreak_2d(dx=0.05).shuffled()
    r = orc.wcsph(2, c.params, c.arrays)
    for k in ("p", "au", "av", "arho"):
        assert np.array_equal(z[k], r[k]), k
