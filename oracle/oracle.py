"""ctypes front-end of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see oracle.cpp header).  May be
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by prestige_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class WcsphParams(C.Structure):
    _fields_ = [("dim", C.c_int32), ("pad", C.c_int32), ("kfac", C.c_double), ("rho0", C.c_double),
                ("c0", C.c_double), ("gamma", C.c_double), ("alpha", C.c_double), ("beta", C.c_double),
                ("g", C.c_double * 3), ("p_given", C.c_int32), ("pad2", C.c_int32)]


class DemParams(C.Structure):
    _fields_ = [("model", C.c_int32), ("K", C.c_int32), ("kn", C.c_double), ("gn", C.c_double),
                ("kt", C.c_double), ("gt", C.c_double), ("mu", C.c_double), ("dt", C.c_double),
                ("Estar", C.c_double), ("Gstar", C.c_double), ("erest", C.c_double)]


class Grid(C.Structure):
    _fields_ = [("dim", C.c_int32), ("n", C.c_int32 * 3), ("lo", C.c_double * 3), ("cell", C.c_double)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_num_threads.restype = C.c_int
        for sfx in ("f64", "f32"):
            getattr(_LIB, f"orc_pairs_allpairs_{sfx}").restype = C.c_int64
            getattr(_LIB, f"orc_pairs_cells_{sfx}").restype = C.c_int64
            getattr(_LIB, f"orc_dem_forces_{sfx}").restype = C.c_int
            getattr(_LIB, f"orc_coupled_forces_{sfx}").restype = C.c_int
    return _LIB


def _sfx(a: np.ndarray) -> str:
    return {np.dtype(np.float64): "f64", np.dtype(np.float32): "f32"}[a.dtype]


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


def make_grid(dim, lo, hi, cell) -> Grid:
    import math
    g = Grid()
    g.dim = dim
    for a in range(3):
        g.n[a] = max(1, int(math.ceil((hi[a] - lo[a]) / cell))) if a < dim else 1
        g.lo[a] = lo[a]
    g.cell = cell
    return g


def wcsph_params(dim: int, P: dict) -> WcsphParams:
    w = WcsphParams()
    w.dim = dim
    w.kfac = P.get("kfac", 2.0)
    w.rho0, w.c0, w.gamma, w.alpha, w.beta = P["rho0"], P["c0"], P["gamma"], P["alpha"], P["beta"]
    w.g[0], w.g[1], w.g[2] = P.get("gx", 0.0), P.get("gy", 0.0), P.get("gz", 0.0)
    w.p_given = 1 if P.get("_p_given") is not None else 0
    return w


def dem_params(P: dict, K: int) -> DemParams:
    d = DemParams()
    d.model = int(P.get("dem_model", 0))
    d.K = K
    d.kn, d.gn, d.kt, d.gt, d.mu, d.dt = P["kn"], P["gn"], P["kt"], P["gt"], P["mu"], P["dt"]
    d.Estar, d.Gstar, d.erest = P.get("Estar", 0.0), P.get("Gstar", 0.0), P.get("erest", 1.0)
    return d


def eq1_allpairs(mass: np.ndarray, force: np.ndarray) -> np.ndarray:
    """force[i] += sum_j mass[j], the loop simple_cpu.rs:7-16 emits for lib.rs:10."""
    force = np.ascontiguousarray(force.copy())
    getattr(lib(), f"orc_eq1_allpairs_{_sfx(mass)}")(C.c_int64(len(mass)), _p(np.ascontiguousarray(mass)), _p(force))
    return force


def pairs(dim, x, y, z, s, mode=0, kfac=2.0, grid: Grid | None = None, cap=None):
    """Neighbour (mode 0, s = h) or contact (mode 1, s = radius) set as an (n_pairs, 2) uint32 array
    sorted lexicographically.  grid=None -> all-pairs (truth); else cell list.  Also returns the
    minimum relative distance of any pair from the cutoff (all-pairs mode only)."""
    n = len(x)
    z = z if z is not None else np.zeros_like(x)
    cap = cap or max(1024, 128 * n)
    oi = np.empty(cap, np.uint32); oj = np.empty(cap, np.uint32)
    sfx = _sfx(x)
    margin = C.c_double(0.0)
    if grid is None:
        cnt = getattr(lib(), f"orc_pairs_allpairs_{sfx}")(C.c_int(dim), C.c_int(mode), C.c_double(kfac), C.c_int64(n),
                                                          _p(x), _p(y), _p(z), _p(s), _p(oi), _p(oj), C.c_int64(cap),
                                                          C.byref(margin))
    else:
        cnt = getattr(lib(), f"orc_pairs_cells_{sfx}")(C.byref(grid), C.c_int(mode), C.c_double(kfac), C.c_int64(n),
                                                       _p(x), _p(y), _p(z), _p(s), _p(oi), _p(oj), C.c_int64(cap))
    assert cnt <= cap, "pair buffer too small"
    pr = np.stack([oi[:cnt], oj[:cnt]], axis=1)
    pr = pr[np.lexsort((pr[:, 1], pr[:, 0]))]
    return pr, margin.value


def wall_pressure(dim, P: dict, a: dict, p: np.ndarray, grid: Grid | None = None):
    """Dummy-particle wall pressure (oracle.cpp wall_pressure): returns (rho, p) copies in which every
    non-fluid row (tag != 0) holds the pressure extrapolated from its fluid neighbours and the density that the EOS
    maps to it; fluid rows are unchanged."""
    x = np.ascontiguousarray(a["x"]); n = len(x)
    z = np.ascontiguousarray(a["z"]) if dim == 3 else np.zeros_like(x)
    rho = np.array(a["rho"], dtype=x.dtype, copy=True); p = np.array(p, dtype=x.dtype, copy=True)
    wp = wcsph_params(dim, P)
    getattr(lib(), f"orc_wall_pressure_{_sfx(x)}")(C.byref(wp), C.byref(grid) if grid is not None else None, C.c_int64(n), _p(x),
                                                   _p(np.ascontiguousarray(a["y"])), _p(z), _p(np.ascontiguousarray(a["h"])),
                                                   _p(np.ascontiguousarray(a["tag"], np.int32)), _p(rho), _p(p))
    return rho, p


def _with_wall_pressure(dim, P: dict, a: dict, grid):
    """boundary_model = 1: EOS, then wall_pressure; the pair loops get the modified rho and p as inputs."""
    rho, p = wall_pressure(dim, P, a, eos(dim, P, np.ascontiguousarray(a["rho"])), grid)
    a = dict(a)
    a["rho"] = rho
    return a, p


def wcsph(dim, P: dict, a: dict, grid: Grid | None = None, sorted_step: bool = False):
    """EOS + continuity + momentum.  Returns dict p, au, av, aw, arho in the input order (+ "rho": the density the
    pair loop saw, when P["boundary_model"] == 1 replaced the dummy particles' by the extrapolated one)."""
    if int(P.get("boundary_model", 0)) == 1 and P.get("_p_given") is None:
        a2, p = _with_wall_pressure(dim, P, a, grid)
        out = wcsph(dim, dict(P, _p_given=p), a2, grid, sorted_step)
        out["rho"] = a2["rho"]
        return out
    x = np.ascontiguousarray(a["x"]); n = len(x); dt = x.dtype
    z = np.ascontiguousarray(a["z"]) if dim == 3 else np.zeros_like(x)
    w = np.ascontiguousarray(a["w"]) if dim == 3 else np.zeros_like(x)
    out = {k: np.zeros(n, dt) for k in ("p", "au", "av", "aw", "arho")}
    if P.get("_p_given") is not None:
        out["p"][:] = P["_p_given"]
    args = [_p(x), _p(np.ascontiguousarray(a["y"])), _p(z), _p(np.ascontiguousarray(a["u"])),
            _p(np.ascontiguousarray(a["v"])), _p(w), _p(np.ascontiguousarray(a["rho"])),
            _p(np.ascontiguousarray(a["m"])), _p(np.ascontiguousarray(a["h"])),
            _p(out["p"]), _p(out["au"]), _p(out["av"]), _p(out["aw"]), _p(out["arho"])]
    wp = wcsph_params(dim, P)
    sfx = _sfx(x)
    if grid is None:
        getattr(lib(), f"orc_wcsph_allpairs_{sfx}")(C.byref(wp), C.c_int64(n), *args)
    elif sorted_step:
        getattr(lib(), f"orc_wcsph_step_sorted_{sfx}")(C.byref(wp), C.byref(grid), C.c_int64(n), *args)
    else:
        getattr(lib(), f"orc_wcsph_cells_{sfx}")(C.byref(wp), C.byref(grid), C.c_int64(n), *args)
    return out


def eos(dim, P: dict, rho: np.ndarray) -> np.ndarray:
    p = np.zeros_like(rho)
    wp = wcsph_params(dim, P)
    getattr(lib(), f"orc_wcsph_eos_{_sfx(rho)}")(C.byref(wp), C.c_int64(len(rho)), _p(np.ascontiguousarray(rho)), _p(p))
    return p


def empty_history(n: int, K: int, dt=np.float64):
    return {"hist_n": np.zeros(n, np.int32), "hist_id": np.zeros((K, n), np.uint32),
            "hist_x": np.zeros((K, n), dt), "hist_y": np.zeros((K, n), dt), "hist_z": np.zeros((K, n), dt)}


def dem(P: dict, K: int, a: dict, hist: dict | None = None, ids: np.ndarray | None = None, grid: Grid | None = None):
    """One DEM force evaluation.  Returns (forces dict, new history dict, overflow flag)."""
    x = np.ascontiguousarray(a["x"]); n = len(x); dt = x.dtype
    hist = hist or empty_history(n, K, dt)
    ids = np.arange(n, dtype=np.uint32) if ids is None else np.ascontiguousarray(ids, np.uint32)
    new = empty_history(n, K, dt)
    out = {k: np.zeros(n, dt) for k in ("fx", "fy", "fz", "tx", "ty", "tz")}
    dp = dem_params(P, K)
    c = lambda k: _p(np.ascontiguousarray(a[k]))
    ov = getattr(lib(), f"orc_dem_forces_{_sfx(x)}")(
        C.byref(dp), C.byref(grid) if grid is not None else None, C.c_int64(n),
        _p(x), c("y"), c("z"), c("u"), c("v"), c("w"), c("wx"), c("wy"), c("wz"), c("rad"), c("m"), _p(ids),
        _p(np.ascontiguousarray(hist["hist_n"])), _p(np.ascontiguousarray(hist["hist_id"])),
        _p(np.ascontiguousarray(hist["hist_x"])), _p(np.ascontiguousarray(hist["hist_y"])),
        _p(np.ascontiguousarray(hist["hist_z"])),
        _p(new["hist_n"]), _p(new["hist_id"]), _p(new["hist_x"]), _p(new["hist_y"]), _p(new["hist_z"]),
        _p(out["fx"]), _p(out["fy"]), _p(out["fz"]), _p(out["tx"]), _p(out["ty"]), _p(out["tz"]))
    return out, new, int(ov)


def sph_mass(a: dict, P: dict) -> np.ndarray:
    """SPH mass of every particle in a coupled block: m for fluid (tag 0) and boundary (tag 1), the displaced fluid
    mass m * rho0 / rho_solid for solids (tag 2)."""
    m = np.ascontiguousarray(a["m"])
    ratio = m.dtype.type(P["rho0"]) / m.dtype.type(P["rho_solid"])
    return np.where(a["tag"] == 2, m * ratio, m).astype(m.dtype)


def coupled(P: dict, K: int, a: dict, hist: dict | None = None, ids: np.ndarray | None = None, grid: Grid | None = None):
    """One coupled SPH-DEM force evaluation (3D).  Returns (rates + forces dict, new history dict, overflow flag).
    P["boundary_model"] == 1: walls and solids carry the extrapolated fluid pressure (wall_pressure); out["rho"] is
    the density the pair loop saw."""
    if int(P.get("boundary_model", 0)) == 1 and P.get("_p_given") is None:
        a2, p = _with_wall_pressure(3, P, a, grid)
        out, new, ov = coupled(dict(P, _p_given=p), K, a2, hist, ids, grid)
        out["rho"] = a2["rho"]
        return out, new, ov
    x = np.ascontiguousarray(a["x"]); n = len(x); dt = x.dtype
    hist = hist or empty_history(n, K, dt)
    ids = np.arange(n, dtype=np.uint32) if ids is None else np.ascontiguousarray(ids, np.uint32)
    new = empty_history(n, K, dt)
    names_out = ("p", "au", "av", "aw", "arho", "fx", "fy", "fz", "tx", "ty", "tz")
    out = {k: np.zeros(n, dt) for k in names_out}
    if P.get("_p_given") is not None:
        out["p"][:] = P["_p_given"]
    ins = {k: np.ascontiguousarray(a[k], dt) for k in ("x", "y", "z", "u", "v", "w", "rho", "m", "h", "wx", "wy", "wz", "rad")}
    ins["ms"] = sph_mass(a, P)
    order = ("x", "y", "z", "u", "v", "w", "rho", "ms", "h", "wx", "wy", "wz", "rad", "m")
    in_ptrs = (C.c_void_p * len(order))(*[ins[k].ctypes.data for k in order])
    out_ptrs = (C.c_void_p * len(names_out))(*[out[k].ctypes.data for k in names_out])
    wp, dp = wcsph_params(3, P), dem_params(P, K)
    tag = np.ascontiguousarray(a["tag"], np.int32)
    body = np.ascontiguousarray(a["body"], np.int32) if "body" in a else None   # rigid bodies: same-body contacts are skipped
    ov = getattr(lib(), f"orc_coupled_forces_{_sfx(x)}")(
        C.byref(wp), C.byref(dp), C.byref(grid) if grid is not None else None, C.c_int64(n), in_ptrs, _p(tag), _p(body), _p(ids),
        _p(np.ascontiguousarray(hist["hist_n"])), _p(np.ascontiguousarray(hist["hist_id"])),
        _p(np.ascontiguousarray(hist["hist_x"])), _p(np.ascontiguousarray(hist["hist_y"])),
        _p(np.ascontiguousarray(hist["hist_z"])), out_ptrs,
        _p(new["hist_n"]), _p(new["hist_id"]), _p(new["hist_x"]), _p(new["hist_y"]), _p(new["hist_z"]))
    return out, new, int(ov)


def coupled_integrate(a: dict, r: dict, P: dict, dt: float) -> dict:
    """The documented semi-implicit Euler stage of a coupled context, on the host (numpy, same operation order
    as k_coupled_integrate): every particle rho += arho dt; fluid v += a dt, x += v dt; solids
    v += (F/m + (rho0/rho_solid)(a - g) + g) dt, x += v dt, omega += T/I dt; boundaries keep x, v."""
    T = a["x"].dtype.type
    dt = T(dt)
    new = {k: np.array(v, copy=True) for k, v in a.items()}
    tag = a["tag"]
    fl, so = tag == 0, tag == 2
    g = [T(P.get("gx", 0.0)), T(P.get("gy", 0.0)), T(P.get("gz", 0.0))]
    ratio = T(P["rho0"]) / T(P["rho_solid"])
    new["rho"] = a["rho"] + r["arho"] * dt
    if int(P.get("boundary_model", 0)) == 1:        # dummy-particle density is slaved to the extrapolated pressure, not integrated
        new["rho"] = np.where(fl, new["rho"], r["rho"] if "rho" in r else a["rho"])
    im = T(1) / a["m"]
    ii = T(1) / np.where(so, a["inertia"], T(1))
    for ax, (pos, vel, acc, frc, gk) in enumerate((("x", "u", "au", "fx", g[0]), ("y", "v", "av", "fy", g[1]), ("z", "w", "aw", "fz", g[2]))):
        a_s = (r[frc] * im + ratio * (r[acc] - gk)) + gk
        vn = np.where(fl, a[vel] + r[acc] * dt, np.where(so, a[vel] + a_s * dt, a[vel]))
        new[vel] = vn
        new[pos] = np.where(fl | so, a[pos] + vn * dt, a[pos])
    for om, tq in (("wx", "tx"), ("wy", "ty"), ("wz", "tz")):
        new[om] = np.where(so, a[om] + r[tq] * ii * dt, a[om])
    return new


# ---------------------------------------------------------------------------------------------------------------
# Multi-particle rigid bodies.  numpy restatement, always float64; no reference code
# exists for this physics -- this is the repo's own written contract, like Appendix A.
# A body record: dict M [nb], X V W F T [nb,3], R [nb,3,3], I0 [nb,3,3].
# ---------------------------------------------------------------------------------------------------------------
def rigid_setup(a: dict, nb: int):
    """Mass, centre of mass, mass-weighted velocity, body-frame offsets r0 = x - X (rounded to the particle dtype) and
    body-frame inertia I0 = sum m (|r0|^2 1 - r0 r0^T) + inertia_i 1 of every body; R = identity, omega = 0.
    Returns (bodies, r0 [n,3]); r0 rows of non-members are zero."""
    body = np.asarray(a["body"])
    mem = body >= 0
    bi = body[mem]
    f8 = lambda k: np.asarray(a[k], np.float64)[mem]
    m = f8("m")
    pos = np.stack([f8("x"), f8("y"), f8("z")], axis=1)
    vel = np.stack([f8("u"), f8("v"), f8("w")], axis=1)
    M = np.bincount(bi, m, nb)
    safe = np.where(M > 0, M, 1.0)
    X = np.stack([np.bincount(bi, m * pos[:, k], nb) for k in range(3)], axis=1) / safe[:, None]
    V = np.stack([np.bincount(bi, m * vel[:, k], nb) for k in range(3)], axis=1) / safe[:, None]
    r0m = (pos - X[bi]).astype(np.asarray(a["x"]).dtype).astype(np.float64)
    r2 = np.sum(r0m * r0m, axis=1)
    I0 = np.zeros((nb, 3, 3))
    ins = f8("inertia")
    for p in range(3):
        for q in range(3):
            term = -m * r0m[:, p] * r0m[:, q]
            if p == q:
                term = term + m * r2 + ins
            I0[:, p, q] = np.bincount(bi, term, nb)
    r0 = np.zeros((len(body), 3))
    r0[mem] = r0m
    bodies = {"M": M, "X": X, "V": V, "W": np.zeros((nb, 3)), "R": np.tile(np.eye(3), (nb, 1, 1)), "I0": I0,
              "F": np.zeros((nb, 3)), "T": np.zeros((nb, 3))}
    return bodies, r0


def rigid_members(bodies: dict, body: np.ndarray, r0: np.ndarray):
    """Member particles from the body state: x = X + R r0, v = V + W x (x - X), spin = W.  Returns (mask, x, v, w)."""
    mem = body >= 0
    bi = body[mem]
    r = np.einsum("npq,nq->np", bodies["R"][bi], r0[mem])
    x = bodies["X"][bi] + r
    v = bodies["V"][bi] + np.cross(bodies["W"][bi], r)
    return mem, x, v, bodies["W"][bi]


def rigid_reduce(a: dict, r: dict, P: dict, bodies: dict) -> dict:
    """F_b = sum Ft_i, T_b = sum (x_i - X_b) x Ft_i + t_i over the members, Ft_i = m_i ((f_i / m_i + ratio (a_i - g)) + g)
    (the single-sphere expression of coupled_integrate times m_i).  Constants are rounded to the particle dtype first,
    as the device does."""
    T = np.asarray(a["x"]).dtype.type
    body = np.asarray(a["body"])
    nb = len(bodies["M"])
    mem = body >= 0
    bi = body[mem]
    f8 = lambda d, k: np.asarray(d[k], np.float64)[mem]
    ratio = float(T(P["rho0"]) / T(P["rho_solid"]))
    g = [float(T(P.get("gx", 0.0))), float(T(P.get("gy", 0.0))), float(T(P.get("gz", 0.0)))]
    m = f8(a, "m")
    Ft = np.stack([m * ((f8(r, fk) / m + ratio * (f8(r, ak) - gk)) + gk)
                   for fk, ak, gk in (("fx", "au", g[0]), ("fy", "av", g[1]), ("fz", "aw", g[2]))], axis=1)
    pos = np.stack([f8(a, "x"), f8(a, "y"), f8(a, "z")], axis=1)
    tq = np.cross(pos - bodies["X"][bi], Ft) + np.stack([f8(r, "tx"), f8(r, "ty"), f8(r, "tz")], axis=1)
    out = dict(bodies)
    out["F"] = np.stack([np.bincount(bi, Ft[:, k], nb) for k in range(3)], axis=1)
    out["T"] = np.stack([np.bincount(bi, tq[:, k], nb) for k in range(3)], axis=1)
    return out


def rigid_integrate(bodies: dict, dt: float) -> dict:
    """Semi-implicit Euler stage of the bodies: V += F/M dt; X += V dt; W += I^-1 (T - W x (I W)) dt with
    I = R I0 R^T; R <- exp([W dt]x) R (Rodrigues).  Bodies without mass are left alone."""
    b = {k: np.array(v, copy=True) for k, v in bodies.items()}
    ok = b["M"] > 0
    M = np.where(ok, b["M"], 1.0)
    V = b["V"] + b["F"] / M[:, None] * dt
    X = b["X"] + V * dt
    I = np.einsum("npq,nqr,nsr->nps", b["R"], b["I0"], b["R"])
    I = np.where(ok[:, None, None], I, np.eye(3))
    Iw = np.einsum("npq,nq->np", I, b["W"])
    rhs = b["T"] - np.cross(b["W"], Iw)
    W = b["W"] + np.linalg.solve(I, rhs[:, :, None])[:, :, 0] * dt
    ang = W * dt
    th = np.linalg.norm(ang, axis=1)
    small = th < 1e-12
    ths = np.where(small, 1.0, th)
    s = np.where(small, 1.0, np.sin(ths) / ths)
    c = np.where(small, 0.5, (1.0 - np.cos(ths)) / (ths * ths))
    K = np.zeros((len(th), 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -ang[:, 2], ang[:, 1]
    K[:, 1, 0], K[:, 1, 2] = ang[:, 2], -ang[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -ang[:, 1], ang[:, 0]
    E = np.eye(3) + s[:, None, None] * K + c[:, None, None] * (K @ K)
    R = E @ b["R"]
    for k, v in (("V", V), ("X", X), ("W", W), ("R", R)):
        b[k] = np.where(ok.reshape((-1,) + (1,) * (v.ndim - 1)), v, bodies[k])
    return b


def history_as_dict(hist: dict, ids: np.ndarray | None = None) -> dict:
    """{(id_i, id_j): (xi_x, xi_y, xi_z)} -- history compared as a keyed set, slot order is free."""
    n = len(hist["hist_n"])
    ids = np.arange(n) if ids is None else ids
    d = {}
    for i in np.nonzero(hist["hist_n"])[0]:
        for k in range(hist["hist_n"][i]):
            d[(int(ids[i]), int(hist["hist_id"][k, i]))] = (hist["hist_x"][k, i], hist["hist_y"][k, i], hist["hist_z"][k, i])
    return d
