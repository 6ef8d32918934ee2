"""Synthetic particle blocks for the configs BASELINE.json names.

Counter-based generator: every random number is SplitMix64(seed, particle id,
field), so any rank can generate any slab of a block without communication and
the CPU oracle and the GPU path see bit-identical inputs.  seed = 42 everywhere.

Nothing here touches the GPU or the oracle; it only builds numpy arrays in
*id order* (lattice order, x slowest) plus the parameter dict a context needs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SEED = 42
_U64 = np.uint64
_MASK = (1 << 64) - 1

# field numbers (stable: they are part of the input definition)
F_JX, F_JY, F_JZ, F_U, F_V, F_W, F_RHO, F_PERM, F_SOLID, F_WX, F_WY, F_WZ = range(12)


def splitmix64(z: np.ndarray) -> np.ndarray:
    """SplitMix64 finaliser on a uint64 array (wrap-around arithmetic)."""
    z = z.astype(_U64, copy=True)
    with np.errstate(over="ignore"):
        z += _U64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> _U64(27))) * _U64(0x94D049BB133111EB)
        z = z ^ (z >> _U64(31))
    return z


def uniform01(ids: np.ndarray, fld: int, seed: int = SEED) -> np.ndarray:
    """U[0,1) as float64 from (seed, id, field); 53 random bits."""
    with np.errstate(over="ignore"):
        k = ids.astype(_U64) * _U64(0xD1342543DE82EF95) + _U64((seed * 0x2545F4914F6CDD1D + fld * 0x9E3779B97F4A7C15 + 1) & _MASK)
    bits = splitmix64(splitmix64(k))
    return (bits >> _U64(11)).astype(np.float64) * (1.0 / (1 << 53))


def usym(ids: np.ndarray, fld: int, amp: float, seed: int = SEED) -> np.ndarray:
    """U(-amp, amp)."""
    return (2.0 * uniform01(ids, fld, seed) - 1.0) * amp


@dataclass
class Block:
    """A synthetic particle block in id order plus everything a context needs."""
    name: str
    dim: int
    physics: str                      # "wcsph" | "dem" | "wcsph+dem" (coupled)
    arrays: dict                      # name -> np.ndarray (float64 / uint32 / int32)
    params: dict                      # pst_set_param names -> float
    lo: tuple
    hi: tuple
    cell_size: float
    max_contacts: int = 0
    meta: dict = field(default_factory=dict)

    @property
    def n(self) -> int:
        return len(self.arrays["x"])

    def astype(self, real) -> "Block":
        arrs = {k: (v.astype(real) if v.dtype.kind == "f" else v) for k, v in self.arrays.items()}
        return Block(self.name, self.dim, self.physics, arrs, dict(self.params), self.lo, self.hi,
                     self.cell_size, self.max_contacts, dict(self.meta))

    def shuffled(self, seed: int = SEED) -> "Block":
        """Same particles, hash-permuted storage order (ids move with them)."""
        n = self.n
        order = np.argsort(splitmix64(np.arange(n, dtype=_U64) + _U64(seed * 7919 + F_PERM)), kind="stable")
        arrs = {k: (v[..., order] if v.shape[-1] == n else v) for k, v in self.arrays.items()}
        return Block(self.name, self.dim, self.physics, arrs, dict(self.params), self.lo, self.hi,
                     self.cell_size, self.max_contacts, dict(self.meta))


CELL_MARGIN = 1.0 + 2.0 ** -20   # cell edge = cutoff * margin, so rounding can never lose a neighbour


def wcsph_params(dim: int, h: float, H: float, g=9.81, alpha=0.1, beta=0.0, rho0=1000.0, gamma=7.0):
    c0 = 10.0 * math.sqrt(2.0 * g * H)
    p = {"rho0": rho0, "c0": c0, "gamma": gamma, "alpha": alpha, "beta": beta, "kfac": 2.0,
         "gx": 0.0, "gy": 0.0, "gz": 0.0}
    p["gy" if dim == 2 else "gz"] = -g
    return p


def _lattice_ids(ix0: int, nx: int, ny: int, nz: int):
    """ids and integer lattice coords of the x-slab [ix0, ix0+nx) of a ny*nz cross-section."""
    i, j, k = np.meshgrid(np.arange(ix0, ix0 + nx, dtype=np.int64), np.arange(ny, dtype=np.int64),
                          np.arange(nz, dtype=np.int64), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    ids = (i * ny + j) * nz + k
    return ids, i, j, k


def wcsph_block_3d(nx: int, ny: int, nz: int, dx: float = 0.005, ix0: int = 0, nx_total: int | None = None,
                   seed: int = SEED, name: str = "wcsph3d") -> Block:
    """C3 / C4: jittered lattice block of fluid, 3D WCSPH.

    `ix0`/`nx_total` select an x-slab of a wider block (multi-GPU weak scaling):
    ids and random fields are those of the global block.
    """
    nx_total = nx if nx_total is None else nx_total
    ids, i, j, k = _lattice_ids(ix0, nx, ny, nz)
    rho0 = 1000.0
    h = 1.2 * dx
    H = nz * dx
    P = wcsph_params(3, h, H)
    jit = 0.05 * dx
    arr = {
        "x": (i + 0.5) * dx + usym(ids, F_JX, jit, seed),
        "y": (j + 0.5) * dx + usym(ids, F_JY, jit, seed),
        "z": (k + 0.5) * dx + usym(ids, F_JZ, jit, seed),
        "u": usym(ids, F_U, 0.1, seed) * P["c0"] * 0.01,
        "v": usym(ids, F_V, 0.1, seed) * P["c0"] * 0.01,
        "w": usym(ids, F_W, 0.1, seed) * P["c0"] * 0.01,
        "rho": rho0 * (1.0 + usym(ids, F_RHO, 1e-3, seed)),
        "m": np.full(len(ids), rho0 * dx ** 3),
        "h": np.full(len(ids), h),
        "tag": np.zeros(len(ids), dtype=np.int32),
    }
    cutoff = 2.0 * h
    lo = (0.0, 0.0, 0.0)
    hi = (nx_total * dx, ny * dx, nz * dx)
    return Block(name, 3, "wcsph", arr, P, lo, hi, cutoff * CELL_MARGIN,
                 meta={"dx": dx, "h": h, "lattice": (nx_total, ny, nz), "ix0": ix0, "ids": ids.astype(np.uint32)})


def wcsph_dambreak_2d(dx: float = 0.01, seed: int = SEED, jitter: float = 0.05) -> Block:
    """C1: 2D dam break, fluid column 1.0 x 2.0 in a 4.0 x 3.0 tank with 3 wall layers (tag 1)."""
    nfx, nfy = int(round(1.0 / dx)), int(round(2.0 / dx))
    i, j = np.meshgrid(np.arange(nfx), np.arange(nfy), indexing="ij")
    fx, fy = (i.ravel() + 0.5) * dx, (j.ravel() + 0.5) * dx
    nf = len(fx)
    # walls: floor and two side walls, 3 layers, outside the tank interior [0,4] x [0,3]
    nl = 3
    ntx, nty = int(round(4.0 / dx)), int(round(3.0 / dx))
    wi, wj = [], []
    for l in range(nl):
        ii = np.arange(-nl, ntx + nl)
        wi.append(ii); wj.append(np.full_like(ii, -1 - l))           # floor
        jj = np.arange(0, nty)
        wi.append(np.full_like(jj, -1 - l)); wj.append(jj)           # left
        wi.append(np.full_like(jj, ntx + l)); wj.append(jj)          # right
    wi, wj = np.concatenate(wi), np.concatenate(wj)
    bx, by = (wi + 0.5) * dx, (wj + 0.5) * dx
    n = nf + len(bx)
    ids = np.arange(n, dtype=np.int64)
    rho0 = 1000.0
    h = 1.2 * dx
    P = wcsph_params(2, h, 2.0)
    x = np.concatenate([fx, bx]); y = np.concatenate([fy, by])
    tag = np.concatenate([np.zeros(nf, np.int32), np.ones(len(bx), np.int32)])
    fluid = tag == 0
    x = x + np.where(fluid, usym(ids, F_JX, jitter * dx, seed), 0.0)
    y = y + np.where(fluid, usym(ids, F_JY, jitter * dx, seed), 0.0)
    arr = {
        "x": x, "y": y,
        "u": np.where(fluid, usym(ids, F_U, 0.1, seed) * P["c0"] * 0.01, 0.0),
        "v": np.where(fluid, usym(ids, F_V, 0.1, seed) * P["c0"] * 0.01, 0.0),
        "rho": rho0 * (1.0 + usym(ids, F_RHO, 1e-3, seed)),
        "m": np.full(n, rho0 * dx * dx),
        "h": np.full(n, h),
        "tag": tag,
    }
    cutoff = 2.0 * h
    pad = (nl + 1) * dx
    lo = (-pad, -pad, 0.0)
    hi = (4.0 + pad, 3.0 + pad, 0.0)
    return Block("wcsph2d_dambreak", 2, "wcsph", arr, P, lo, hi, cutoff * CELL_MARGIN,
                 meta={"dx": dx, "h": h, "n_fluid": nf, "n_boundary": len(bx)})


def dem_params(R: float = 1e-3, rho_s: float = 2500.0, kn: float = 1e5, e: float = 0.8, mu: float = 0.5,
               dt: float = 1e-6):
    m = rho_s * 4.0 / 3.0 * math.pi * R ** 3
    le = math.log(e)
    meff = 0.5 * m
    gn = -2.0 * le * math.sqrt(meff * kn) / math.sqrt(le * le + math.pi ** 2)
    return {"dem_model": 0.0, "kn": kn, "gn": gn, "kt": 2.0 / 7.0 * kn, "gt": 0.5 * gn, "mu": mu, "dt": dt,
            "Estar": 1e7, "Gstar": 4e6, "erest": e}


def dem_column_3d(n_side: int = 100, R: float = 1e-3, floor: bool = True, seed: int = SEED,
                  nx: int | None = None, ix0: int = 0, nx_total: int | None = None) -> Block:
    """C2: n^3 monodisperse spheres on a lattice at 2R(1-0.01) (1 % overlap) + jitter, plus a floor
    plane of wall spheres (tag 1) one lattice step below."""
    ny = nz = n_side
    nx = n_side if nx is None else nx
    nx_total = nx if nx_total is None else nx_total
    ids, i, j, k = _lattice_ids(ix0, nx, ny, nz)
    sp = 2.0 * R * (1.0 - 0.01)
    jit = 0.005 * R
    P = dem_params(R)
    x = (i + 0.5) * sp + usym(ids, F_JX, jit, seed)
    y = (j + 0.5) * sp + usym(ids, F_JY, jit, seed)
    z = (k + 0.5) * sp + usym(ids, F_JZ, jit, seed)
    tag = np.zeros(len(ids), np.int32)
    u = usym(ids, F_U, 0.01, seed); v = usym(ids, F_V, 0.01, seed); w = usym(ids, F_W, 0.01, seed)
    gid = ids
    if floor:
        fi, fj = np.meshgrid(np.arange(ix0, ix0 + nx, dtype=np.int64), np.arange(ny, dtype=np.int64), indexing="ij")
        fi, fj = fi.ravel(), fj.ravel()
        fid = nx_total * ny * nz + fi * ny + fj
        x = np.concatenate([x, (fi + 0.5) * sp]); y = np.concatenate([y, (fj + 0.5) * sp])
        z = np.concatenate([z, np.full(len(fi), -0.5 * sp)])
        tag = np.concatenate([tag, np.ones(len(fi), np.int32)])
        u = np.concatenate([u, np.zeros(len(fi))]); v = np.concatenate([v, np.zeros(len(fi))])
        w = np.concatenate([w, np.zeros(len(fi))])
        gid = np.concatenate([ids, fid])
    n = len(x)
    m = 2500.0 * 4.0 / 3.0 * math.pi * R ** 3
    arr = {
        "x": x, "y": y, "z": z, "u": u, "v": v, "w": w,
        "wx": np.zeros(n), "wy": np.zeros(n), "wz": np.zeros(n),
        "rad": np.full(n, R), "m": np.full(n, m), "inertia": np.full(n, 0.4 * m * R * R),
        "tag": tag,
    }
    cutoff = 2.0 * R
    lo = (0.0, 0.0, -sp)
    hi = (nx_total * sp, ny * sp, nz * sp)
    return Block("dem3d_column", 3, "dem", arr, P, lo, hi, cutoff * CELL_MARGIN, max_contacts=12,
                 meta={"R": R, "spacing": sp, "lattice": (nx_total, ny, nz), "ix0": ix0, "ids": gid.astype(np.uint32)})


def coupled_block_3d(nx: int, ny: int, nz: int, dx: float = 0.005, solid_fraction: float = 0.3, floor: bool = True,
                     ix0: int = 0, nx_total: int | None = None, seed: int = SEED, rho_solid: float = 2500.0) -> Block:
    """C5 (BASELINE configs[4]): rigid spheres in fluid.  A jittered lattice block like C3 whose lower third holds
    solid spheres (tag 2) on a hash-selected `solid_fraction` of the sites (10 % of all particles at 0.3), radius
    0.505 dx so lattice-adjacent spheres overlap by ~1 % (contacts exist at step 0, as in C2), small random spin;
    optional floor of 3 static boundary layers (tag 1) that are SPH dummy particles and DEM wall spheres at once.
    `ix0`/`nx_total` select an x-slab of a wider block (ids and random fields are those of the global block)."""
    nx_total = nx if nx_total is None else nx_total
    ids, i, j, k = _lattice_ids(ix0, nx, ny, nz)
    rho0 = 1000.0
    h = 1.2 * dx
    R = 0.505 * dx
    P = wcsph_params(3, h, nz * dx)
    P.update(dem_params(R, rho_s=rho_solid))
    P["rho_solid"] = rho_solid
    solid = (k < nz // 3) & (uniform01(ids, F_SOLID, seed) < solid_fraction)
    jit = np.where(solid, 0.005 * R, 0.05 * dx)
    x = (i + 0.5) * dx + (2.0 * uniform01(ids, F_JX, seed) - 1.0) * jit
    y = (j + 0.5) * dx + (2.0 * uniform01(ids, F_JY, seed) - 1.0) * jit
    z = (k + 0.5) * dx + (2.0 * uniform01(ids, F_JZ, seed) - 1.0) * jit
    vs = np.where(solid, 0.01, 0.1 * P["c0"] * 0.01)
    u = (2.0 * uniform01(ids, F_U, seed) - 1.0) * vs
    v = (2.0 * uniform01(ids, F_V, seed) - 1.0) * vs
    w = (2.0 * uniform01(ids, F_W, seed) - 1.0) * vs
    wx = np.where(solid, usym(ids, F_WX, 1.0, seed), 0.0)
    wy = np.where(solid, usym(ids, F_WY, 1.0, seed), 0.0)
    wz = np.where(solid, usym(ids, F_WZ, 1.0, seed), 0.0)
    tag = np.where(solid, 2, 0).astype(np.int32)
    gid = ids
    nl = 3 if floor else 0
    if floor:
        l, fi, fj = np.meshgrid(np.arange(nl, dtype=np.int64), np.arange(ix0, ix0 + nx, dtype=np.int64), np.arange(ny, dtype=np.int64), indexing="ij")
        l, fi, fj = l.ravel(), fi.ravel(), fj.ravel()
        fid = nx_total * ny * nz + (l * nx_total + fi) * ny + fj
        nfl = len(fid)
        x = np.concatenate([x, (fi + 0.5) * dx]); y = np.concatenate([y, (fj + 0.5) * dx]); z = np.concatenate([z, -(l + 0.5) * dx])
        zero = np.zeros(nfl)
        u, v, w = (np.concatenate([q, zero]) for q in (u, v, w))
        wx, wy, wz = (np.concatenate([q, zero]) for q in (wx, wy, wz))
        tag = np.concatenate([tag, np.ones(nfl, np.int32)])
        gid = np.concatenate([ids, fid])
    n = len(x)
    is_solid = tag == 2
    m_s = rho_solid * 4.0 / 3.0 * math.pi * R ** 3
    m = np.where(is_solid, m_s, rho0 * dx ** 3)
    arr = {
        "x": x, "y": y, "z": z, "u": u, "v": v, "w": w,
        "rho": rho0 * (1.0 + usym(gid, F_RHO, 1e-3, seed)),
        "m": m, "h": np.full(n, h), "tag": tag,
        "wx": wx, "wy": wy, "wz": wz,
        "rad": np.where(tag == 0, 0.0, R), "inertia": np.where(is_solid, 0.4 * m_s * R * R, 1.0),
    }
    cutoff = 2.0 * h
    lo = (0.0, 0.0, -nl * dx)
    hi = (nx_total * dx, ny * dx, nz * dx)
    return Block("coupled3d", 3, "wcsph+dem", arr, P, lo, hi, cutoff * CELL_MARGIN, max_contacts=12,
                 meta={"dx": dx, "h": h, "R": R, "lattice": (nx_total, ny, nz), "ix0": ix0, "ids": gid.astype(np.uint32),
                       "n_solid": int(is_solid.sum()), "n_boundary": int((tag == 1).sum())})


def rigid_block_3d(nx: int, ny: int, nz: int, group: int = 2, **kw) -> Block:
    """Rigid bodies in fluid: the coupled block whose solid spheres are joined into multi-particle
    rigid bodies, one per `group`^3 cube of lattice sites that holds at least one solid.  Members of a body overlap
    like any lattice-adjacent spheres (radius 0.505 dx), so the same-body contact exclusion is exercised, and bodies
    touch their neighbours' members.  Adds the i32 array `body` (-1 = not a member) and meta["n_bodies"]."""
    b = coupled_block_3d(nx, ny, nz, **kw)
    ix0 = kw.get("ix0", 0)
    _, i, j, k = _lattice_ids(ix0, nx, ny, nz)
    n_lat = len(i)
    solid = b.arrays["tag"][:n_lat] == 2
    gy, gz = (ny + group - 1) // group, (nz + group - 1) // group
    key = ((i // group) * gy + j // group) * gz + k // group
    uniq, inv = np.unique(key[solid], return_inverse=True)
    body = np.full(b.n, -1, np.int32)
    body[np.nonzero(solid)[0]] = inv.astype(np.int32)
    b.arrays["body"] = body
    b.meta["n_bodies"] = int(len(uniq))
    b.name = "rigid3d"
    return b


def grid_dims(block: Block):
    """Cell-grid extents the way the library computes them: ceil((hi-lo)/cell), at least 1."""
    n = []
    for a in range(3):
        ext = block.hi[a] - block.lo[a]
        n.append(max(1, int(math.ceil(ext / block.cell_size))) if a < block.dim else 1)
    return tuple(n)
