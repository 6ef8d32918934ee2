"""Multi-particle rigid bodies in a coupled context.

CPU part: the per-body arithmetic the device runs (prestige_b200/csrc/rigid_core.h, compiled for the host by
tests/cpp/rigid_core_harness.cpp) against the numpy oracle, and the oracle's own invariants.
GPU part: setup, same-body contact exclusion, reduction, stage and conservation through the C ABI.
No reference code exists for this physics: parity is against oracle/, "parity unpinned".
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc
from prestige_b200 import synth

from util import assert_close, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "cpp", "librigid_core_harness.so")


def _harness():
    src = os.path.join(ROOT, "tests", "cpp", "rigid_core_harness.cpp")
    hdr = os.path.join(ROOT, "prestige_b200", "csrc", "rigid_core.h")
    if not os.path.exists(HARNESS) or os.path.getmtime(HARNESS) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", HARNESS], check=True, capture_output=True)
    lib = C.CDLL(HARNESS)
    lib.rbh_fields.restype = C.c_int
    return lib


def _random_bodies(nb, rng):
    """Plausible body records: SPD inertia, random rotation, random state."""
    A = rng.normal(size=(nb, 3, 3))
    I0 = A @ np.transpose(A, (0, 2, 1)) + 3.0 * np.eye(3)
    Q, _ = np.linalg.qr(rng.normal(size=(nb, 3, 3)))
    Q = Q * np.sign(np.linalg.det(Q))[:, None, None]
    return {"M": rng.uniform(0.5, 2.0, nb), "X": rng.normal(size=(nb, 3)), "V": rng.normal(size=(nb, 3)),
            "W": rng.normal(size=(nb, 3)) * 5.0, "R": Q, "I0": I0 * 1e-3, "F": rng.normal(size=(nb, 3)), "T": rng.normal(size=(nb, 3)) * 1e-2}


def _pack(lib, bodies, b):
    lay = (C.c_int * 8)()
    lib.rbh_layout(lay)
    s = np.zeros(lib.rbh_fields())
    s[lay[0]] = bodies["M"][b]
    for k, off in (("X", lay[1]), ("V", lay[2]), ("W", lay[3]), ("F", lay[6]), ("T", lay[7])):
        s[off:off + 3] = bodies[k][b]
    s[lay[4]:lay[4] + 9] = bodies["R"][b].ravel()
    I = bodies["I0"][b]
    s[lay[5]:lay[5] + 6] = [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]
    return s, lay


def test_core_integrate_matches_oracle():
    lib = _harness()
    rng = np.random.default_rng(7)
    nb = 64
    bodies = _random_bodies(nb, rng)
    bodies["W"][0] = 0.0; bodies["T"][0] = 0.0          # the vanishing-angle branch
    for dt in (1e-4, 0.05):
        ref = orc.rigid_integrate(bodies, dt)
        for b in range(nb):
            s, lay = _pack(lib, bodies, b)
            lib.rbh_integrate(s.ctypes.data_as(C.c_void_p), C.c_double(dt))
            for k, off, w in (("X", lay[1], 3), ("V", lay[2], 3), ("W", lay[3], 3), ("R", lay[4], 9)):
                got, want = s[off:off + w], ref[k][b].ravel()
                assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, np.max(np.abs(want))), (k, b, dt)


def test_core_member_and_force_match_oracle():
    lib = _harness()
    rng = np.random.default_rng(11)
    nb, n = 8, 200
    bodies = _random_bodies(nb, rng)
    body = rng.integers(0, nb, n).astype(np.int32)
    r0 = rng.normal(size=(n, 3)) * 1e-2
    _, x, v, w = orc.rigid_members(bodies, body, r0)
    for i in range(n):
        s, _ = _pack(lib, bodies, body[i])
        xo, vo = np.zeros(3), np.zeros(3)
        lib.rbh_member(s.ctypes.data_as(C.c_void_p), r0[i].ctypes.data_as(C.c_void_p), xo.ctypes.data_as(C.c_void_p), vo.ctypes.data_as(C.c_void_p))
        assert np.allclose(xo, x[i], rtol=0, atol=1e-14) and np.allclose(vo, v[i], rtol=0, atol=1e-13)
    # particle force: the single-sphere expression of coupled_integrate times m
    m, ratio = 0.37, 0.4
    f, a, g = rng.normal(size=3), rng.normal(size=3) * 10, np.array([0.0, 0.0, -9.81])
    out = np.zeros(3)
    lib.rbh_particle_force(C.c_double(m), C.c_double(ratio), f.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, m * ((f / m + ratio * (a - g)) + g))


def test_oracle_rigid_invariants():
    """Free body (no force): linear and angular momentum R I0 R^T W are conserved, R stays a rotation, members keep
    their mutual distances; the torque-free symmetric top keeps |W|."""
    rng = np.random.default_rng(3)
    nb = 5
    bodies = _random_bodies(nb, rng)
    bodies["F"][:] = 0.0; bodies["T"][:] = 0.0
    body = rng.integers(0, nb, 60).astype(np.int32)
    r0 = rng.normal(size=(60, 3)) * 1e-2
    L0 = np.einsum("npq,nq->np", np.einsum("npq,nqr,nsr->nps", bodies["R"], bodies["I0"], bodies["R"]), bodies["W"])
    _, x0, _, _ = orc.rigid_members(bodies, body, r0)
    ang = lambda q: np.einsum("npq,nq->np", np.einsum("npq,nqr,nsr->nps", q["R"], q["I0"], q["R"]), q["W"])
    dt = 2e-5
    b = bodies
    for _ in range(500):
        b = orc.rigid_integrate(b, dt)
    c = bodies
    for _ in range(1000):
        c = orc.rigid_integrate(c, 0.5 * dt)
    d1, d2 = np.max(np.abs(ang(b) - L0)), np.max(np.abs(ang(c) - L0))
    assert d1 <= 1e-3 * np.max(np.abs(L0))                # first-order scheme: O(dt) drift of R I0 R^T W, not conservation to rounding
    assert 0.4 < d2 / d1 < 0.6                            # ... and it halves with the step
    assert np.max(np.abs(b["R"] @ np.transpose(b["R"], (0, 2, 1)) - np.eye(3))) < 1e-12
    assert np.allclose(b["V"], bodies["V"]) and np.allclose(b["X"], bodies["X"] + bodies["V"] * dt * 500)
    _, x1, _, _ = orc.rigid_members(b, body, r0)
    same = body[:, None] == body[None, :]
    d0 = np.linalg.norm(x0[:, None] - x0[None, :], axis=2)[same]
    d1 = np.linalg.norm(x1[:, None] - x1[None, :], axis=2)[same]
    assert np.max(np.abs(d1 - d0)) < 1e-13


def test_oracle_setup_and_reduce_cpu():
    """Setup: masses, centres and inertia of the synthetic bodies; reduce: internal contact forces are absent
    (same-body exclusion), so the body forces sum with the fluid's to the antisymmetric-pair total."""
    blk = synth.rigid_block_3d(12, 10, 12).shuffled()
    a, P, nb = blk.arrays, blk.params, blk.meta["n_bodies"]
    bodies, r0 = orc.rigid_setup(a, nb)
    cnt = np.bincount(a["body"][a["body"] >= 0], minlength=nb)
    m_s = a["m"][a["tag"] == 2][0]
    assert np.allclose(bodies["M"], cnt * m_s, rtol=1e-14)
    assert np.all(np.linalg.eigvalsh(bodies["I0"]) > 0)
    mem = a["body"] >= 0
    assert np.max(np.abs(np.stack([np.bincount(a["body"][mem], a["m"][mem] * r0[mem][:, k], nb) for k in range(3)], 1))) < 1e-18
    # contact set with and without the exclusion
    with_b, _, _ = orc.coupled(P, blk.max_contacts, a)
    no_body = {k: v for k, v in a.items() if k != "body"}
    without_b, _, _ = orc.coupled(P, blk.max_contacts, no_body)
    hn_w = orc.coupled(P, blk.max_contacts, a)[1]["hist_n"]
    hn_o = orc.coupled(P, blk.max_contacts, no_body)[1]["hist_n"]
    assert hn_w.sum() < hn_o.sum() and np.all(hn_w <= hn_o)               # members of one body do overlap in this block
    red = orc.rigid_reduce(a, with_b, P, bodies)
    # per-body force = sum over members of the per-particle expression
    Ft = a["m"][:, None] * ((np.stack([with_b["fx"], with_b["fy"], with_b["fz"]], 1) / a["m"][:, None]
                             + P["rho0"] / P["rho_solid"] * (np.stack([with_b["au"], with_b["av"], with_b["aw"]], 1) - np.array([P["gx"], P["gy"], P["gz"]])))
                            + np.array([P["gx"], P["gy"], P["gz"]]))
    assert np.allclose(red["F"].sum(0), Ft[mem].sum(0), rtol=1e-12, atol=1e-18)
    assert without_b is not None


# ---------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------
def _state(ctx, names):
    return {k: ctx.download(k) for k in names}


def _bodies_from_ctx(ctx):
    """Every per-body quantity as the device holds it, in the oracle's record layout."""
    nb = ctx.n_bodies
    i6 = ctx.body_get("inertia0")
    I0 = np.empty((nb, 3, 3))
    I0[:, 0, 0], I0[:, 1, 1], I0[:, 2, 2] = i6[:, 0], i6[:, 1], i6[:, 2]
    I0[:, 0, 1] = I0[:, 1, 0] = i6[:, 3]
    I0[:, 0, 2] = I0[:, 2, 0] = i6[:, 4]
    I0[:, 1, 2] = I0[:, 2, 1] = i6[:, 5]
    return {"M": ctx.body_get("mass"), "X": ctx.body_get("cm"), "V": ctx.body_get("vel"), "W": ctx.body_get("omega"),
            "R": ctx.body_get("rot").reshape(nb, 3, 3), "I0": I0, "F": ctx.body_get("force"), "T": ctx.body_get("torque")}


STATE = ("x", "y", "z", "u", "v", "w", "rho", "m", "h", "tag", "wx", "wy", "wz", "rad", "inertia", "body")


@pytest.mark.gpu
@pytest.mark.parametrize("real", [np.float64, np.float32])
def test_rigid_setup_reduce_parity_gpu(real):
    import prestige_b200 as pb
    tol = 1e-10 if real == np.float64 else 1e-5
    blk = synth.rigid_block_3d(14, 12, 14).shuffled().astype(real)
    nb = blk.meta["n_bodies"]
    with pb.context_for_block(blk, real=real) as ctx:
        ctx.load_block(blk)                       # bodies_create + uploads + bodies_setup
        bodies, r0 = orc.rigid_setup(blk.arrays, nb)
        for k, name in (("M", "mass"), ("X", "cm"), ("V", "vel")):
            assert rel_err(ctx.body_get(name), bodies[k]) <= tol, name
        I0 = ctx.body_get("inertia0")
        want = np.stack([bodies["I0"][:, 0, 0], bodies["I0"][:, 1, 1], bodies["I0"][:, 2, 2], bodies["I0"][:, 0, 1], bodies["I0"][:, 0, 2], bodies["I0"][:, 1, 2]], 1)
        assert rel_err(I0, want) <= tol
        assert rel_err(np.stack([ctx.download(k) for k in ("bx0", "by0", "bz0")], 1).astype(np.float64), r0) <= tol
        # members now carry the rigid motion: the oracle evaluates the SAME state
        a = _state(ctx, STATE)
        mem, x, v, w = orc.rigid_members(bodies, a["body"], r0)
        assert rel_err(np.stack([a["u"], a["v"], a["w"]], 1)[mem].astype(np.float64), v) <= tol
        ctx.build_neighbours()
        ctx.apply(["tait_eos", "continuity", "momentum", "dem_contact", "body_reduce"])
        ctx.sync()
        ref, hist, ov = orc.coupled(blk.params, blk.max_contacts, a)
        assert not ov
        # contact set: solid i, non-fluid j, not the same body -- bit-exact as a set
        refc, _ = orc.pairs(3, a["x"], a["y"], a["z"], a["rad"], mode=1)
        ti, tj = a["tag"][refc[:, 0]], a["tag"][refc[:, 1]]
        bi, bj = a["body"][refc[:, 0]], a["body"][refc[:, 1]]
        same = (bi >= 0) & (bi == bj)
        assert same.any()                         # the block does hold overlapping members of one body
        keep = (ti == 2) & (tj != 0) & ~same
        assert np.array_equal(ctx.dump_pairs(1), refc[keep])
        assert np.array_equal(ctx.download("hist_n"), hist["hist_n"])
        for k in ("fx", "fy", "fz", "tx", "ty", "tz", "au", "av", "aw"):
            assert_close(ctx.download(k), ref[k], k, tol)
        # body sums of per-particle values that agree to `tol`; the torque about the centre of mass cancels further
        bref = orc.rigid_reduce(a, ref, blk.params, bodies)
        assert rel_err(ctx.body_get("force"), bref["F"]) <= 10 * tol
        assert rel_err(ctx.body_get("torque"), bref["T"]) <= 100 * tol
        # ... and exactly (to summation order) the sums of the device's own per-particle values
        own = {k: ctx.download(k) for k in ("fx", "fy", "fz", "tx", "ty", "tz", "au", "av", "aw")}
        bown = orc.rigid_reduce(a, own, blk.params, _bodies_from_ctx(ctx))
        assert rel_err(ctx.body_get("force"), bown["F"]) <= 1e-12
        assert rel_err(ctx.body_get("torque"), bown["T"]) <= 1e-11


@pytest.mark.gpu
def test_rigid_step_matches_host_stage_gpu():
    """pst_step against the documented stage evaluated on the host from the GPU's own forces, three steps."""
    import prestige_b200 as pb
    blk = synth.rigid_block_3d(12, 10, 12).shuffled()
    nb = blk.meta["n_bodies"]
    dt = 2e-6
    with pb.context_for_block(blk) as ctx:
        ctx.load_block(blk)
        ctx.set_params(dt=dt)
        a = _state(ctx, STATE)
        r0 = np.stack([ctx.download(k) for k in ("bx0", "by0", "bz0")], 1)
        for step in range(3):
            ctx.build_neighbours()
            ctx.apply(["tait_eos", "continuity", "momentum", "dem_contact", "body_reduce"])
            r = {k: ctx.download(k) for k in ("au", "av", "aw", "arho", "fx", "fy", "fz", "tx", "ty", "tz")}
            bodies = _bodies_from_ctx(ctx)        # state before the stage, force and torque of this evaluation
            ctx.integrate(dt)
            new = orc.coupled_integrate(a, r, blk.params, dt)
            nbod = orc.rigid_integrate(bodies, dt)
            mem, x, v, w = orc.rigid_members(nbod, a["body"], r0)
            for c, (p, q, o) in enumerate((("x", "u", "wx"), ("y", "v", "wy"), ("z", "w", "wz"))):
                new[p][mem], new[q][mem], new[o][mem] = x[:, c], v[:, c], w[:, c]
            got = _state(ctx, STATE)
            for k in ("x", "y", "z", "u", "v", "w", "rho", "wx", "wy", "wz"):
                assert_close(got[k], new[k], f"{k} after step {step}", 1e-10)
            for k, name in (("X", "cm"), ("V", "vel"), ("W", "omega")):
                assert rel_err(ctx.body_get(name), nbod[k]) <= 1e-10, name
            assert rel_err(ctx.body_get("rot").reshape(nb, 3, 3), nbod["R"]) <= 1e-12
            a = got


@pytest.mark.gpu
def test_rigid_conservation_and_rigidity_gpu():
    """200 steps, g = 0, no floor: fluid + body linear momentum is conserved to summation-order error, members of a
    body keep their mutual distances, bodies keep their mass."""
    import prestige_b200 as pb
    blk = synth.rigid_block_3d(12, 10, 12, floor=False).shuffled()
    blk.params.update(gx=0.0, gy=0.0, gz=0.0)
    nb = blk.meta["n_bodies"]
    dt = 2e-6
    with pb.context_for_block(blk) as ctx:
        ctx.load_block(blk)
        ctx.set_params(dt=dt)

        def momentum():
            s = _state(ctx, ("u", "v", "w", "m", "tag", "body"))
            free = s["body"] < 0
            pf = np.stack([np.sum((s["m"] * s[k])[free]) for k in ("u", "v", "w")])
            return pf + (ctx.body_get("mass")[:, None] * ctx.body_get("vel")).sum(0), s

        p0, s0 = momentum()
        x0 = np.stack([ctx.download(k) for k in ("x", "y", "z")], 1)
        scale = np.sum(s0["m"] * np.sqrt(s0["u"] ** 2 + s0["v"] ** 2 + s0["w"] ** 2))
        ctx.step(dt, 200)
        ctx.sync()
        p1, s1 = momentum()
        assert np.max(np.abs(p1 - p0)) <= 1e-10 * scale
        x1 = np.stack([ctx.download(k) for k in ("x", "y", "z")], 1)
        body = s1["body"]
        for b in range(nb):
            idx = np.nonzero(body == b)[0]
            if len(idx) > 1:
                d0 = np.linalg.norm(x0[idx][:, None] - x0[idx][None, :], axis=2)
                d1 = np.linalg.norm(x1[idx][:, None] - x1[idx][None, :], axis=2)
                assert np.max(np.abs(d1 - d0)) <= 1e-12 * blk.meta["dx"] * 100
        R = ctx.body_get("rot").reshape(nb, 3, 3)
        assert np.max(np.abs(R @ np.transpose(R, (0, 2, 1)) - np.eye(3))) < 1e-11


@pytest.mark.gpu
def test_rigid_api_errors_gpu():
    import prestige_b200 as pb
    from prestige_b200 import _lib as L
    blk = synth.rigid_block_3d(8, 8, 8).shuffled()
    with pb.context_for_block(blk) as ctx:
        with pytest.raises(L.PstError) as e:                     # body_reduce without bodies
            ctx.load_block(blk, arrays=[k for k in blk.arrays if k != "body"])
            ctx.build_neighbours()
            ctx.apply(["body_reduce"])
        assert e.value.status == L.PST_ESTATE
        ctx.bodies_create(blk.meta["n_bodies"])
        with pytest.raises(L.PstError):                          # twice
            ctx.bodies_create(3)
        bad = blk.arrays["body"].copy()
        bad[np.argmax(bad >= 0)] = blk.meta["n_bodies"] + 5      # index out of range -> PST_EINVAL at setup
        ctx.upload("body", bad)
        with pytest.raises(L.PstError) as e:
            ctx.bodies_setup()
        assert e.value.status == L.PST_EINVAL
        ctx.upload("body", blk.arrays["body"])
        ctx.bodies_setup()
        with pytest.raises(L.PstError):                          # read-only quantity
            ctx.body_set("force", np.ones((blk.meta["n_bodies"], 3)))
    w = synth.wcsph_block_3d(8, 8, 8)
    with pb.context_for_block(w) as ctx:
        with pytest.raises(L.PstError) as e:                     # bodies need a coupled context
            ctx.bodies_create(2)
        assert e.value.status == L.PST_ESTATE


@pytest.mark.gpu
def test_rigid_checkpoint_resume_bit_exact_gpu(tmp_path):
    """A run split by a checkpoint (particles, history, body records) continues bit-identically."""
    import prestige_b200 as pb
    blk = synth.rigid_block_3d(12, 10, 12).shuffled()
    dt = 2e-6
    names = ("x", "y", "z", "u", "v", "w", "rho", "wx", "wy", "wz")
    with pb.context_for_block(blk) as ctx:
        ctx.load_block(blk)
        ctx.step(dt, 5)
        pb.io.save_checkpoint(ctx, str(tmp_path / "rigid.npz"))
        ctx.step(dt, 5)
        want = {k: ctx.download(k) for k in names}
        want_b = {k: ctx.body_get(k) for k in ("cm", "vel", "omega", "rot")}
    with pb.context_for_block(blk) as ctx:
        ctx.set_params(**blk.params)
        pb.io.load_checkpoint(ctx, str(tmp_path / "rigid.npz"))
        ctx.step(dt, 5)
        for k in names:
            assert np.array_equal(ctx.download(k), want[k]), k
        for k, v in want_b.items():
            assert np.array_equal(ctx.body_get(k), v), k             # body sums run in a fixed order (bpos): deterministic
