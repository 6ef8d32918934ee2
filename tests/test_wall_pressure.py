"""Dummy-particle wall pressure: every non-fluid particle takes the pressure
extrapolated from its fluid neighbours (after Adami, Hu & Adams 2012) and the density the EOS maps to it.

CPU part: the oracle's restatement (cell list == all pairs, the hydrostatic known answer, fluid rows untouched,
golden fixture).  GPU part (-m gpu): equation "wall_pressure" through the C ABI against that oracle -- p and rho of the
dummy particles and the rates of the following pair kernel within 1e-10 (f64) / 1e-5 (f32), pst_step with the
parameter boundary_model = 1 against the documented stage on the host.  "Parity unpinned": the reference has no such
code.
"""
import os

import numpy as np
import pytest

from prestige_b200 import synth
from oracle import oracle as orc
from util import assert_close, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RATES2, RATES3 = ("p", "rho", "au", "av", "arho"), ("p", "rho", "au", "av", "aw", "arho")


def hydrostatic_tank_2d(nx=40, ny=40, dx=0.01, layers=3):
    """A resting column over a floor: fluid rows carry the hydrostatic density, floor rows rho0."""
    h, H = 1.2 * dx, ny * dx
    P = synth.wcsph_params(2, h, H)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(-layers, ny), indexing="ij")
    x, y = (ix.ravel() + 0.5) * dx, (iy.ravel() + 0.5) * dx
    tag = (iy.ravel() < 0).astype(np.int32)
    n = len(x)
    B = P["rho0"] * P["c0"] ** 2 / P["gamma"]
    rho = P["rho0"] * (P["rho0"] * 9.81 * np.maximum(H - y, 0.0) / B + 1.0) ** (1.0 / P["gamma"])
    rho[tag == 1] = P["rho0"]
    a = {"x": x, "y": y, "u": np.zeros(n), "v": np.zeros(n), "rho": rho, "m": np.full(n, P["rho0"] * dx * dx), "h": np.full(n, h), "tag": tag}
    return P, a, H


def walled_block_3d(nx=12, ny=11, nz=13, layers=3):
    """The jittered WCSPH block with its lowest lattice planes turned into wall particles (tag 1)."""
    b = synth.wcsph_block_3d(nx, ny, nz)
    b.arrays["tag"] = (b.arrays["z"] < (layers + 0.25) * b.meta["dx"]).astype(np.int32)
    b.params["boundary_model"] = 1.0
    return b


# ------------------------------------------------------------------------------------------------
# CPU: the oracle itself
# ------------------------------------------------------------------------------------------------
def test_hydrostatic_known_answer():
    """p_w = p_f + rho g (depth difference): the two wall layers that see fluid continue the hydrostatic line; the
    third layer (no fluid inside its support) gets 0; the resting fluid above the floor stays (nearly) at rest."""
    P, a, H = hydrostatic_tank_2d()
    p = orc.eos(2, P, a["rho"])
    rho_w, p_w = orc.wall_pressure(2, P, a, p)
    fl, y = a["tag"] == 0, a["y"]
    assert np.array_equal(p_w[fl], p[fl]) and np.array_equal(rho_w[fl], a["rho"][fl]), "fluid rows are untouched"
    mid = (~fl) & (np.abs(a["x"] - 0.2) < 0.1)
    exact = P["rho0"] * 9.81 * (H - y)
    seen = mid & (y > -0.02)
    assert np.abs(p_w[seen] / exact[seen] - 1.0).max() < 2e-3
    assert (p_w[mid & (y < -0.02)] == 0).all()
    # EOS consistency of the wall density
    assert np.abs(orc.eos(2, P, rho_w)[~fl] - p_w[~fl]).max() < 1e-9 * p_w.max()
    core = fl & (np.abs(a["x"] - 0.2) < 0.1) & (y < 0.3)
    with_model = orc.wcsph(2, dict(P, boundary_model=1), a)
    without = orc.wcsph(2, P, a)
    assert np.abs(with_model["av"][core]).max() < 0.01 * np.abs(without["av"][core]).max(), "the extrapolated floor holds the column"


@pytest.mark.parametrize("real", [np.float64, np.float32])
def test_wall_pressure_cells_match_allpairs(real):
    for b in (synth.wcsph_dambreak_2d(dx=0.04).shuffled().astype(real), walled_block_3d().shuffled().astype(real)):
        P = dict(b.params, boundary_model=1)
        g = orc.make_grid(b.dim, b.lo, b.hi, b.cell_size)
        ra = orc.wcsph(b.dim, P, b.arrays)
        assert (ra["rho"][b.arrays["tag"] != 0] != b.arrays["rho"][b.arrays["tag"] != 0]).any()
        for mode in (False, True):
            rc = orc.wcsph(b.dim, P, b.arrays, grid=g, sorted_step=mode)
            for k in ra:
                assert rel_err(rc[k], ra[k]) <= (1e-12 if real == np.float64 else 2e-5), k
        r0 = orc.wcsph(b.dim, b.params if b.dim == 2 else dict(b.params, boundary_model=0), b.arrays)
        assert "rho" not in r0, "boundary_model = 0 leaves the state density alone"


def test_coupled_wall_pressure_oracle():
    b = synth.coupled_block_3d(12, 10, 12).shuffled()
    P = dict(b.params, boundary_model=1)
    g = orc.make_grid(3, b.lo, b.hi, b.cell_size)
    ra, ha, _ = orc.coupled(P, b.max_contacts, b.arrays)
    rc, hc, _ = orc.coupled(P, b.max_contacts, b.arrays, grid=g)
    r0, h0, _ = orc.coupled(b.params, b.max_contacts, b.arrays)
    tag = b.arrays["tag"]
    for k in RATES3 + ("fx", "fy", "fz"):
        assert rel_err(rc[k], ra[k]) <= 1e-12, k
    for k in ("fx", "fy", "fz", "tx", "ty", "tz"):
        assert np.array_equal(ra[k], r0[k]), "contacts do not depend on the boundary model"
    assert np.array_equal(ra["p"][tag == 0], r0["p"][tag == 0])
    assert (ra["p"][tag != 0] != r0["p"][tag != 0]).any()
    # pairwise antisymmetry survives: both sides of a fluid-dummy pair read the same extrapolated pressure
    P0 = dict(P, gz=0.0)
    bb = synth.coupled_block_3d(12, 10, 12, floor=False).shuffled()
    r, _, _ = orc.coupled(P0, bb.max_contacts, bb.arrays)
    ms = orc.sph_mass(bb.arrays, P0)
    for acc, f in (("au", "fx"), ("av", "fy"), ("aw", "fz")):
        tot = (ms * r[acc]).sum() + r[f].sum()
        assert abs(tot) <= 1e-12 * ((ms * np.abs(r[acc])).sum() + np.abs(r[f]).sum())


def test_wall_pressure_golden_vectors():
    z = np.load(os.path.join(GOLD, "wall2d_small.npz"))
    c = synth.wcsph_dambreak_2d(dx=0.05).shuffled()
    r = orc.wcsph(2, dict(c.params, boundary_model=1), c.arrays)
    for k in RATES2:
        assert np.array_equal(z[k], r[k]), k


# ------------------------------------------------------------------------------------------------
# GPU: the CUDA path through the C ABI
# ------------------------------------------------------------------------------------------------
def _ctx(block, real, **kw):
    import prestige_b200 as pb
    ctx = pb.context_for_block(block, real=real, **kw)
    ctx.load_block(block)
    return ctx


EQS = ["tait_eos", "wall_pressure", "continuity", "momentum"]


@pytest.mark.gpu
@pytest.mark.parametrize("real", [np.float64, np.float32])
@pytest.mark.parametrize("key", ["linear", "morton"])
def test_wall_pressure_gpu_wcsph(real, key):
    for b in (synth.wcsph_dambreak_2d(dx=0.02).shuffled().astype(real), walled_block_3d(14, 12, 15).shuffled().astype(real)):
        b.params["boundary_model"] = 1.0
        ref = orc.wcsph(b.dim, b.params, b.arrays)
        with _ctx(b, real, key=key) as ctx:
            for ev in range(2):                  # the second pass re-sorts and starts from the slaved wall density
                ctx.build_neighbours()
                ctx.apply(EQS)
                for k in (RATES3 if b.dim == 3 else RATES2):
                    assert_close(ctx.download(k), ref[k], f"wall pressure {b.dim}D {k} eval {ev} {key}")
                if ev == 0:                      # what the oracle sees on the second pass: the state the device now holds
                    ref = orc.wcsph(b.dim, b.params, dict(b.arrays, rho=ref["rho"]))
    if real == np.float64 and key == "linear":   # golden fixture (f64, the small dam break)
        z = np.load(os.path.join(GOLD, "wall2d_small.npz"))
        c = synth.wcsph_dambreak_2d(dx=0.05).shuffled()
        c.params["boundary_model"] = 1.0
        with _ctx(c, np.float64) as ctx:
            ctx.build_neighbours()
            ctx.apply(EQS)
            for k in RATES2:
                assert_close(ctx.download(k), z[k], f"golden wall pressure {k}")


@pytest.mark.gpu
@pytest.mark.parametrize("real", [np.float64, np.float32])
def test_wall_pressure_gpu_coupled(real):
    b = synth.coupled_block_3d(14, 12, 15).shuffled().astype(real)
    b.params["boundary_model"] = 1.0
    ref, hist, _ = orc.coupled(b.params, b.max_contacts, b.arrays)
    with _ctx(b, real) as ctx:
        ctx.build_neighbours()
        ctx.apply(EQS + ["dem_contact"])
        for k in RATES3 + ("fx", "fy", "fz", "tx", "ty", "tz"):
            assert_close(ctx.download(k), ref[k], f"coupled wall pressure {k}")
        assert np.array_equal(ctx.download("hist_n"), hist["hist_n"])


@pytest.mark.gpu
def test_wall_pressure_step_and_errors():
    """pst_step with boundary_model = 1 == build + EOS + wall_pressure + pair kernel + the documented stage on the host
    (dummy density slaved, not integrated); the equation is refused without neighbours / EOS."""
    import prestige_b200 as pb
    b = walled_block_3d(10, 10, 12).shuffled()
    dt = 1e-5
    a = b.arrays
    ref = orc.wcsph(3, b.params, a)
    fl = a["tag"] == 0
    exp = {"rho": np.where(fl, a["rho"] + ref["arho"] * dt, ref["rho"])}
    for pos, vel, acc in (("x", "u", "au"), ("y", "v", "av"), ("z", "w", "aw")):
        exp[vel] = np.where(fl, a[vel] + ref[acc] * dt, a[vel])
        exp[pos] = np.where(fl, a[pos] + exp[vel] * dt, a[pos])
    with _ctx(b, np.float64) as ctx:
        ctx.step(dt, 1)
        for k, v in exp.items():
            assert_close(ctx.download(k), v, f"step {k}", tol=1e-12)
    c = synth.coupled_block_3d(10, 9, 12).shuffled()
    c.params["boundary_model"] = 1.0
    Pc = dict(c.params, dt=dt)                              # pst_step hands its dt to the contact model (history increment)
    r, _, _ = orc.coupled(Pc, c.max_contacts, c.arrays)
    new = orc.coupled_integrate(c.arrays, r, Pc, dt)
    with _ctx(c, np.float64) as ctx:
        ctx.step(dt, 1)
        for k in ("x", "y", "z", "u", "v", "w", "rho", "wx"):
            assert_close(ctx.download(k), new[k], f"coupled step {k}", tol=1e-11)
    with _ctx(b, np.float64) as ctx:
        with pytest.raises(pb.PstError):
            ctx.apply(["wall_pressure"])                    # no neighbours yet
        ctx.build_neighbours()
        ctx.upload("rho", b.arrays["rho"])                  # (the re-sort of a single-GPU context evaluates the EOS: touch rho)
        with pytest.raises(pb.PstError):
            ctx.apply(["wall_pressure"])                    # p of the fluid is not current
        ctx.apply(["tait_eos"])
        ctx.apply(["wall_pressure"])                        # separate sets are fine once the EOS has run
        assert_close(ctx.download("p"), ref["p"], "separate application p")
