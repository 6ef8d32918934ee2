"""Coupled SPH-DEM (BASELINE.json configs[4]): rigid spheres in fluid, one cell grid.

CPU part: the oracle's coupled restatement (cell list == all pairs, pairwise antisymmetry, golden fixture).
GPU part (-m gpu): the CUDA path through the C ABI against that oracle -- neighbour and contact sets bit-exact as
sets, rates / contact forces / history within 1e-10 (f64) or 1e-5 (f32), pst_step against the documented integrator,
momentum conservation over 1000 steps.  "Parity unpinned": the reference has no such code.
"""
import os

import numpy as np
import pytest

from prestige_b200 import synth
from oracle import oracle as orc
from util import assert_close, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RATES = ("p", "au", "av", "aw", "arho")
FORCES = ("fx", "fy", "fz", "tx", "ty", "tz")


def _sets(b):
    """(SPH neighbour set, contact set) of a coupled block from the all-pairs oracle, filtered by the tag rules."""
    a = b.arrays
    tag = a["tag"]
    nb, margin = orc.pairs(3, a["x"], a["y"], a["z"], a["h"])
    nb = nb[(tag[nb[:, 0]] == 0) | (tag[nb[:, 1]] == 0)]
    ct, _ = orc.pairs(3, a["x"], a["y"], a["z"], a["rad"], mode=1)
    ct = ct[(tag[ct[:, 0]] == 2) & (tag[ct[:, 1]] != 0)]
    return nb, ct, margin


# ------------------------------------------------------------------------------------------------
# CPU: the oracle itself
# ------------------------------------------------------------------------------------------------
def test_coupled_block_shape():
    b = synth.coupled_block_3d(12, 10, 12)
    tag = b.arrays["tag"]
    assert b.physics == "wcsph+dem" and b.n == 12 * 10 * 12 + 3 * 12 * 10
    assert (tag == 1).sum() == 3 * 12 * 10 and 0.05 * b.n < (tag == 2).sum() < 0.2 * b.n
    assert (b.arrays["z"][tag == 2] < 4 * b.meta["dx"]).all(), "solids live in the lower third"
    assert (b.arrays["rad"][tag == 0] == 0).all() and (b.arrays["rad"][tag != 0] > 0).all()
    # slabs of a wider block reproduce the global block's particles (counter-based generator)
    s = synth.coupled_block_3d(5, 10, 12, ix0=4, nx_total=12)
    m = (b.meta["ids"][:, None] == s.meta["ids"][None, :]).argmax(0)
    for k in ("x", "z", "u", "wx", "rho", "m"):
        assert np.array_equal(b.arrays[k][m], s.arrays[k]), k
    assert np.array_equal(b.arrays["tag"][m], s.arrays["tag"])


def test_coupled_cells_match_allpairs():
    b = synth.coupled_block_3d(12, 10, 12).shuffled()
    g = orc.make_grid(3, b.lo, b.hi, b.cell_size)
    hist = None
    for _ in range(2):
        ra, ha, ov = orc.coupled(b.params, b.max_contacts, b.arrays, hist=hist)
        rc, hc, _ = orc.coupled(b.params, b.max_contacts, b.arrays, hist=hist, grid=g)
        assert ov == 0
        for k in RATES + FORCES:
            assert rel_err(rc[k], ra[k]) <= 1e-12, k
        assert np.array_equal(ha["hist_n"], hc["hist_n"])
        da, dc = orc.history_as_dict(ha), orc.history_as_dict(hc)
        assert da.keys() == dc.keys()
        hist = ha
    tag = b.arrays["tag"]
    _, ct, _ = _sets(b)
    assert ha["hist_n"].sum() == len(ct) > 0
    assert (ha["hist_n"][tag != 2] == 0).all(), "only solids carry contacts"
    assert (ra["fx"][tag != 2] == 0).all()


def test_coupled_reduces_to_wcsph_and_dem():
    """Without solids the coupled sums are the WCSPH ones (boundary-boundary pairs aside); the solids' contact forces
    are the DEM ones of the non-fluid sub-system."""
    b = synth.coupled_block_3d(10, 9, 12).shuffled()
    a = b.arrays
    tag = a["tag"]
    r, h, _ = orc.coupled(b.params, b.max_contacts, a)
    sub = tag != 0
    d = {k: np.ascontiguousarray(v[..., sub]) for k, v in a.items()}
    fd, hd, _ = orc.dem(b.params, b.max_contacts, d)
    sol = tag[sub] == 2
    for k in FORCES:
        assert np.array_equal(r[k][sub][sol], fd[k][sol]), k
    fl = dict(a)
    fl["m"] = orc.sph_mass(a, b.params)
    w = orc.wcsph(3, b.params, fl)
    # a fluid particle's sums are exactly the WCSPH ones with the SPH masses
    for k in RATES:
        assert np.array_equal(r[k][tag == 0], w[k][tag == 0]), k


def test_coupled_momentum_antisymmetry():
    b = synth.coupled_block_3d(12, 10, 12, floor=False).shuffled()
    P = dict(b.params, gz=0.0)
    r, _, _ = orc.coupled(P, b.max_contacts, b.arrays)
    ms = orc.sph_mass(b.arrays, P)
    for acc, f in (("au", "fx"), ("av", "fy"), ("aw", "fz")):
        tot = (ms * r[acc]).sum() + r[f].sum()
        scale = (ms * np.abs(r[acc])).sum() + np.abs(r[f]).sum()
        assert abs(tot) <= 1e-12 * scale, f"{acc}: net force {tot / scale:.3e}"


def test_coupled_golden_vectors():
    z = np.load(os.path.join(GOLD, "coupled3d_small.npz"))
    b = synth.coupled_block_3d(9, 8, 9).shuffled()
    assert np.array_equal(z["x"], b.arrays["x"]) and np.array_equal(z["tag"], b.arrays["tag"]), "synthetic generator changed"
    r1, h1, _ = orc.coupled(b.params, b.max_contacts, b.arrays)
    r2, h2, _ = orc.coupled(b.params, b.max_contacts, b.arrays, hist=h1)
    for k in RATES + FORCES:
        assert np.array_equal(z[k], r2[k]), k
    assert np.array_equal(z["hist_n"], h2["hist_n"])
    nb, ct, _ = _sets(b)
    assert np.array_equal(z["neighbours"], nb) and np.array_equal(z["contacts"], ct)


# ------------------------------------------------------------------------------------------------
# GPU: the CUDA path through the C ABI
# ------------------------------------------------------------------------------------------------
def _ctx(block, real, **kw):
    import prestige_b200 as pb
    ctx = pb.context_for_block(block, real=real, **kw)
    ctx.load_block(block)
    return ctx


@pytest.mark.gpu
@pytest.mark.parametrize("real", [np.float64, np.float32])
@pytest.mark.parametrize("variant", [3, 2, 0])
def test_coupled_gpu_one_evaluation(real, variant):
    b = synth.coupled_block_3d(14, 12, 15).shuffled().astype(real)
    nb, ct, margin = _sets(b)
    assert margin > 1e-12 or real == np.float32
    hist = None
    with _ctx(b, real) as ctx:
        if variant == 2:
            ctx.set_option("zsub", 1)             # the list kernel cuts its tiles in whole cells
        ctx.set_option("force_kernel", variant)
        for ev in range(3):                       # history carries over; the second and third pass re-sort first
            ref, hist, ov = orc.coupled(b.params, b.max_contacts, b.arrays, hist=hist)
            assert ov == 0
            ctx.build_neighbours()
            ctx.apply(["tait_eos", "continuity", "momentum", "dem_contact"])
            ctx.sync()
            for k in RATES + FORCES:
                assert_close(ctx.download(k), ref[k], f"coupled {k} eval {ev} variant {variant}")
            got = {k: ctx.download(k) for k in ("hist_n", "hist_id", "hist_x", "hist_y", "hist_z")}
            assert np.array_equal(got["hist_n"], hist["hist_n"])
            dg, dr = orc.history_as_dict(got), orc.history_as_dict(hist)
            assert dg.keys() == dr.keys()
            tol = 1e-10 if real == np.float64 else 1e-5
            sc = max(np.abs(np.array(list(dr.values()))).max(), 1e-300)
            assert max(np.abs(np.array(dg[k]) - np.array(dr[k])).max() for k in dr) <= tol * sc
        assert np.array_equal(ctx.dump_pairs(0), nb), "SPH neighbour set (pairs with a fluid member) must be bit-exact"
        assert np.array_equal(ctx.dump_pairs(1), ct), "contact set (solid i, non-fluid j) must be bit-exact"


@pytest.mark.gpu
def test_coupled_gpu_golden_and_edge_tiles():
    import prestige_b200 as pb
    z = np.load(os.path.join(GOLD, "coupled3d_small.npz"))
    b = synth.coupled_block_3d(9, 8, 9).shuffled()
    for opts in ({}, {"tile_g": 1}, {"tile_g": 4, "tile_jcap": 100, "tile_lcap": 16}, {"dem_kernel": 0}):
        with _ctx(b, np.float64) as ctx:
            for k, v in opts.items():
                ctx.set_option(k, v)
            ctx.build_neighbours()
            ctx.apply(["tait_eos", "continuity", "momentum", "dem_contact"])
            ctx.apply(["tait_eos", "continuity", "momentum", "dem_contact"])
            for k in RATES + FORCES:
                assert_close(ctx.download(k), z[k], f"golden coupled {k} {opts}")
            assert np.array_equal(ctx.download("hist_n"), z["hist_n"])
            assert np.array_equal(ctx.dump_pairs(0), z["neighbours"]) and np.array_equal(ctx.dump_pairs(1), z["contacts"])
    with _ctx(b, np.float64) as ctx:              # the fused pair kernel computes continuity and momentum together
        ctx.build_neighbours()
        with pytest.raises(pb.PstError):
            ctx.apply(["tait_eos", "momentum"])


@pytest.mark.gpu
def test_coupled_step_matches_host_integration():
    """pst_step == build + EOS + fused SPH pass + contact pass + the documented integrator, three steps in a row."""
    b = synth.coupled_block_3d(10, 9, 12).shuffled()
    dt = 2e-6
    P = dict(b.params, dt=dt)
    a, hist = b.arrays, None
    with _ctx(b, np.float64) as ctx:
        for step in range(3):
            r, hist, _ = orc.coupled(P, b.max_contacts, a, hist=hist)
            a = orc.coupled_integrate(a, r, P, dt)
            ctx.step(dt, 1)
            for k in ("x", "y", "z", "u", "v", "w", "rho", "wx", "wy", "wz"):
                assert_close(ctx.download(k), a[k], f"step {step} {k}", tol=1e-11)
        tag = b.arrays["tag"]
        assert np.array_equal(ctx.download("x")[tag == 1], b.arrays["x"][tag == 1]), "boundaries do not move"


@pytest.mark.gpu
def test_coupled_conservation_1000_steps():
    """No gravity, no boundaries: total linear momentum of fluid + spheres stays within 1e-10 over 1000 steps, no
    particle is lost, the spheres' contact history stays bounded."""
    b = synth.coupled_block_3d(12, 10, 12, floor=False).shuffled()
    b.params["gz"] = 0.0
    a = b.arrays
    m = a["m"]
    with _ctx(b, np.float64) as ctx:
        p0 = np.array([(m * a[k]).sum() for k in ("u", "v", "w")])
        ctx.step(2e-6, 1000)
        ctx.sync()
        p1 = np.array([(m * ctx.download(k)).sum() for k in ("u", "v", "w")])
        x1 = ctx.download("x")
        assert np.array_equal(ctx.download("m"), m) and np.array_equal(ctx.download("tag"), a["tag"])
        assert ctx.download("hist_n").max() <= b.max_contacts
    assert np.isfinite(x1).all()
    scale = (m * np.abs(a["u"])).sum()
    assert np.abs(p1 - p0).max() <= 1e-10 * scale, f"momentum drift {np.abs(p1 - p0).max() / scale:.3e}"


@pytest.mark.gpu
def test_coupled_300k_against_oracle():
    b = synth.coupled_block_3d(64, 64, 72)
    g = orc.make_grid(3, b.lo, b.hi, b.cell_size)
    r1, h1, ov = orc.coupled(b.params, b.max_contacts, b.arrays, grid=g)
    r2, h2, _ = orc.coupled(b.params, b.max_contacts, b.arrays, hist=h1, grid=g)
    assert ov == 0
    with _ctx(b, np.float64) as ctx:
        for _ in range(2):
            ctx.build_neighbours()
            ctx.apply(["tait_eos", "continuity", "momentum", "dem_contact"])
        for k in RATES + FORCES:
            assert_close(ctx.download(k), r2[k], f"coupled 300k {k}")
        assert np.array_equal(ctx.download("hist_n"), h2["hist_n"])


@pytest.mark.gpu
def test_coupled_20m_against_oracle():
    """BASELINE configs[4] at FULL size (250 x 250 x 320 lattice + floor, 20.2 M particles, 10 % spheres) on one GPU against the
    oracle's cell-list mode with all host threads: SPH rates, contact forces and torques, stored-contact counts; plus the
    size-independent property sum m (a - g) = 0 over fluid + spheres for the SPH part (pairwise antisymmetry)."""
    import os
    b = synth.coupled_block_3d(250, 250, 320)
    orc.set_num_threads(os.cpu_count() or 1)
    ref, hist, ov = orc.coupled(b.params, b.max_contacts, b.arrays, grid=orc.make_grid(3, b.lo, b.hi, b.cell_size))
    assert ov == 0
    with _ctx(b, np.float64) as ctx:
        ctx.build_neighbours()
        ctx.apply(["tait_eos", "continuity", "momentum", "dem_contact"])
        for k in RATES + FORCES:
            assert_close(ctx.download(k), ref[k], f"coupled 20M {k}")
        assert np.array_equal(ctx.download("hist_n"), hist["hist_n"])
