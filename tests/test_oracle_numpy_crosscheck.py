"""A second, independent restatement of SURVEY.md Appendix A in vectorised numpy (all pairs, small N), written from the
formulas rather than from oracle.cpp, cross-checked against the C++ oracle.  The reference pins nothing on this path
("parity unpinned"); two independently written restatements agreeing is the strongest pin available."""
import numpy as np

from oracle import oracle as orc
from prestige_b200 import synth
from util import rel_err


def numpy_wcsph(dim, P, a, coupled=False, wall_model=False):
    """Appendix A.1 (neighbour rule) + A.2 (Wendland C2, Tait EOS, continuity, momentum with artificial viscosity).
    coupled: a pair counts iff i or j is fluid; a solid's SPH mass is the displaced fluid mass.
    wall_model: non-fluid rows take the pressure extrapolated from their fluid neighbours
    p_w = sum (p_f + rho_f g . x_wf) W_wf / sum W_wf and the density rho0 (p_w / B + 1)^(1/gamma)."""
    pos = np.stack([a["x"], a["y"]] + ([a["z"]] if dim == 3 else []), axis=1)
    vel = np.stack([a["u"], a["v"]] + ([a["w"]] if dim == 3 else []), axis=1)
    rho, m, h = a["rho"], a["m"], a["h"]
    tag = a.get("tag", np.zeros(len(rho), np.int32))
    if coupled:
        m = np.where(tag == 2, m * P["rho0"] / P["rho_solid"], m)
    xij = pos[:, None, :] - pos[None, :, :]
    vij = vel[:, None, :] - vel[None, :, :]
    r2 = np.sum(xij * xij, axis=2)
    hi = h[:, None]
    nb = (r2 < (P.get("kfac", 2.0) * hi) ** 2) & (r2 > 0)
    r = np.sqrt(np.where(nb, r2, 1.0))
    q = r / hi
    ad = 21.0 / (16.0 * np.pi * hi ** 3) if dim == 3 else 7.0 / (4.0 * np.pi * hi ** 2)
    dwdq = -5.0 * ad * q * (1.0 - 0.5 * q) ** 3
    gw = np.where(nb, dwdq / (hi * r), 0.0)[:, :, None] * xij             # grad_i W_ij
    B = P["rho0"] * P["c0"] ** 2 / P["gamma"]
    p = B * ((rho / P["rho0"]) ** P["gamma"] - 1.0)
    g = np.array([P.get("gx", 0.0), P.get("gy", 0.0), P.get("gz", 0.0)])[:dim]
    if wall_model:
        W = np.where(nb & (tag[None, :] == 0), ad * (1.0 - 0.5 * q) ** 4 * (2.0 * q + 1.0), 0.0)
        S0 = W.sum(axis=1)
        num = np.sum(W * (p[None, :] + rho[None, :] * np.sum(xij * g, axis=2)), axis=1)
        pw = np.where(S0 > 0, num / np.where(S0 > 0, S0, 1.0), 0.0)
        p = np.where(tag != 0, pw, p)
        rho = np.where(tag != 0, P["rho0"] * (np.maximum(pw / B, -0.5) + 1.0) ** (1.0 / P["gamma"]), rho)
    if coupled:
        nb = nb & ((tag[:, None] == 0) | (tag[None, :] == 0))
        gw = np.where(nb[:, :, None], gw, 0.0)
    vx = np.sum(vij * xij, axis=2)
    arho = np.sum(m[None, :] * np.sum(vij * gw, axis=2), axis=1)
    mu = hi * vx / (r2 + 0.01 * hi ** 2)
    rhob = 0.5 * (rho[:, None] + rho[None, :])
    Pi = np.where(vx < 0, (-P["alpha"] * P["c0"] * mu + P["beta"] * mu * mu) / rhob, 0.0)
    coef = -(m[None, :] * ((p / rho ** 2)[:, None] + (p / rho ** 2)[None, :] + Pi))
    acc = np.sum(coef[:, :, None] * gw, axis=1) + g
    return {"p": p, "arho": arho, "au": acc[:, 0], "av": acc[:, 1], **({"aw": acc[:, 2]} if dim == 3 else {})}, nb


def numpy_dem_pass(P, a, xi0=None):
    """Appendix A.3, one evaluation, linear or Hertz-Mindlin law.  xi0[i, j] = tangential spring stored at i for partner j
    (None: no stored history, xi = 0 before the update).  Returns (forces, contact mask, capped count, new xi)."""
    pos = np.stack([a["x"], a["y"], a["z"]], axis=1)
    vel = np.stack([a["u"], a["v"], a["w"]], axis=1)
    om = np.stack([a["wx"], a["wy"], a["wz"]], axis=1)
    R, m = a["rad"], a["m"]
    xij = pos[:, None, :] - pos[None, :, :]
    r2 = np.sum(xij * xij, axis=2)
    rs = R[:, None] + R[None, :]
    ct = (r2 < rs * rs) & (r2 > 0)
    r = np.sqrt(np.where(ct, r2, 1.0))
    n = xij / r[:, :, None]
    delta = rs - r
    Rw = R[:, None, None] * om[:, None, :] + R[None, :, None] * om[None, :, :]
    vc = (vel[:, None, :] - vel[None, :, :]) - np.cross(Rw, n)
    vn = np.sum(vc * n, axis=2)
    vt = vc - vn[:, :, None] * n
    kn, gn, kt, gt = P["kn"], P["gn"], P["kt"], P["gt"]
    if int(P.get("dem_model", 0)) == 1:
        Rs = R[:, None] * R[None, :] / rs
        ms = m[:, None] * m[None, :] / (m[:, None] + m[None, :])
        sq = np.sqrt(np.where(ct, Rs * delta, 1.0))
        Sn, St = 2.0 * P["Estar"] * sq, 8.0 * P["Gstar"] * sq
        le = np.log(P["erest"])
        be = -le / np.sqrt(le * le + np.pi ** 2)
        kn, kt = 4.0 / 3.0 * P["Estar"] * sq, St
        gn, gt = 2.0 * np.sqrt(5.0 / 6.0) * be * np.sqrt(Sn * ms), 2.0 * np.sqrt(5.0 / 6.0) * be * np.sqrt(St * ms)
        kt = kt[:, :, None]; gt = gt[:, :, None]
    fn = kn * delta - gn * vn
    xi = np.zeros_like(vt) if xi0 is None else xi0
    xi = xi - np.sum(xi * n, axis=2)[:, :, None] * n + vt * P["dt"]       # rotate into the current tangent plane, then stretch
    ft = -kt * xi - gt * vt
    ftm = np.sqrt(np.sum(ft * ft, axis=2))
    cap = P["mu"] * np.abs(fn)
    sc = np.where(ftm > cap, cap / np.where(ftm > 0, ftm, 1.0), 1.0)
    ft = ft * sc[:, :, None]
    xi = np.where((ftm > cap)[:, :, None], -(ft + gt * vt) / kt, xi)      # a sliding contact keeps the spring at the Coulomb limit
    F = np.sum(np.where(ct[:, :, None], fn[:, :, None] * n + ft, 0.0), axis=1)
    T = np.sum(np.where(ct[:, :, None], np.cross(-R[:, None, None] * n, ft), 0.0), axis=1)
    capped = int(((ftm > cap) & ct).sum())
    return {"fx": F[:, 0], "fy": F[:, 1], "fz": F[:, 2], "tx": T[:, 0], "ty": T[:, 1], "tz": T[:, 2]}, ct, capped, xi


def _pairs_of(mask):
    i, j = np.nonzero(mask)
    pr = np.stack([i, j], axis=1).astype(np.uint32)
    return pr[np.lexsort((pr[:, 1], pr[:, 0]))]


def test_wcsph_numpy_vs_cpp():
    for b in (synth.wcsph_block_3d(8, 7, 9).shuffled(), synth.wcsph_dambreak_2d(dx=0.08).shuffled()):
        P = dict(b.params, beta=0.2)
        ref, nb = numpy_wcsph(b.dim, P, b.arrays)
        got = orc.wcsph(b.dim, P, b.arrays)
        prs, _ = orc.pairs(b.dim, b.arrays["x"], b.arrays["y"], b.arrays.get("z"), b.arrays["h"])
        assert np.array_equal(prs, _pairs_of(nb)), "neighbour sets of the two restatements differ"
        for k in ref:
            assert rel_err(got[k], ref[k]) <= 1e-10, (b.name, k, rel_err(got[k], ref[k]))


def test_dem_numpy_vs_cpp():
    d = synth.dem_column_3d(7).shuffled()
    rng = np.random.default_rng(7)
    spin = {k: rng.uniform(-1.0, 1.0, d.n) ** 3 for k in ("wx", "wy", "wz")}
    for extra, amp in (({}, 2.5e4), ({"dem_model": 1, "Estar": 1e7, "Gstar": 4e6, "erest": 0.8}, 1e3)):
        for k in spin:                                    # spin: exercises the torque and the R w x n term, and makes the surfaces
            d.arrays[k] = spin[k] * amp                   # slide fast enough for the Coulomb cap to bite on some of the contacts
        P = dict(d.params, **extra)
        ref, ct, capped, xi = numpy_dem_pass(P, d.arrays)
        assert 0.1 * ct.sum() < capped < 0.9 * ct.sum(), "both branches of the Coulomb cap must be exercised"
        got, hist, ov = orc.dem(P, 12, d.arrays)
        assert ov == 0
        prs, _ = orc.pairs(3, d.arrays["x"], d.arrays["y"], d.arrays["z"], d.arrays["rad"], mode=1)
        assert np.array_equal(prs, _pairs_of(ct))
        assert hist["hist_n"].sum() == ct.sum()
        stored = orc.history_as_dict(hist)                # {(i, j): xi} -- the tangential springs the pass leaves behind
        assert set(stored) == {(int(i), int(j)) for i, j in zip(*np.nonzero(ct))}
        xs = max(float(np.abs(xi[ct]).max()), 1e-300)
        assert max(float(np.abs(np.array(v) - xi[i, j]).max()) for (i, j), v in stored.items()) <= 1e-10 * xs
        for k in ref:
            assert rel_err(got[k], ref[k]) <= 1e-10, (extra, k, rel_err(got[k], ref[k]))
        # second evaluation: the stored springs are projected, stretched and (where sliding) reset again
        moved = dict(d.arrays)
        for k, amp in (("x", 2e-6), ("y", 2e-6), ("z", 2e-6)):
            moved[k] = d.arrays[k] + rng.uniform(-amp, amp, d.n)
        ref2, ct2, _, xi2 = numpy_dem_pass(P, moved, np.where(ct[:, :, None], xi, 0.0))
        got2, hist2, _ = orc.dem(P, 12, moved, hist=hist)
        for k in ref2:
            assert rel_err(got2[k], ref2[k]) <= 1e-10, ("second pass", extra, k, rel_err(got2[k], ref2[k]))
        stored2 = orc.history_as_dict(hist2)
        assert set(stored2) == {(int(i), int(j)) for i, j in zip(*np.nonzero(ct2))}
        assert max(float(np.abs(np.array(v) - xi2[i, j]).max()) for (i, j), v in stored2.items()) <= 1e-10 * max(float(np.abs(xi2[ct2]).max()), 1e-300)


def test_coupled_rule_and_wall_pressure_numpy_vs_cpp():
    """The coupled SPH pair rule with displaced-fluid masses and the dummy-particle wall pressure (4d)."""
    c = synth.coupled_block_3d(8, 7, 8).shuffled()
    tag = c.arrays["tag"]
    for wall in (False, True):
        P = dict(c.params, boundary_model=1 if wall else 0)
        ref, nb = numpy_wcsph(3, P, c.arrays, coupled=True, wall_model=wall)
        got, _, _ = orc.coupled(P, c.max_contacts, c.arrays)
        for k in ref:
            assert rel_err(got[k], ref[k]) <= 1e-10, (wall, k, rel_err(got[k], ref[k]))
        if wall:
            assert (got["p"][tag != 0] != orc.coupled(c.params, c.max_contacts, c.arrays)[0]["p"][tag != 0]).any()
    w = synth.wcsph_dambreak_2d(dx=0.08).shuffled()                       # WCSPH-only context with tag-1 walls
    P = dict(w.params, boundary_model=1)
    ref, _ = numpy_wcsph(2, P, w.arrays, wall_model=True)
    got = orc.wcsph(2, P, w.arrays)
    for k in ref:
        assert rel_err(got[k], ref[k]) <= 1e-10, (k, rel_err(got[k], ref[k]))
