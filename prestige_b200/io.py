"""Snapshot I/O: the formats the reference's .gitignore anticipates (*.vtk, *.csv;
/root/reference/.gitignore:4,6) plus a raw checkpoint that includes DEM contact history, so a long run can be split
and resumed bit-exactly.  Host-side only: everything goes through Context.download / Context.upload (id order).
"""
from __future__ import annotations

import numpy as np

_STATE_WCSPH = ["x", "y", "z", "u", "v", "w", "rho", "m", "h", "tag"]
_STATE_DEM = ["x", "y", "z", "u", "v", "w", "wx", "wy", "wz", "rad", "m", "inertia", "tag"]
_HISTORY = ["hist_n", "hist_id", "hist_x", "hist_y", "hist_z"]
_RIGID = ["body", "bpos", "bx0", "by0", "bz0"]                                   # particle side of the rigid bodies
_BODY_STATE = ["mass", "inertia0", "cm", "vel", "omega", "rot"]          # per-body records, restored in this order


def state_names(ctx) -> list:
    names = []
    for k in _STATE_WCSPH + _STATE_DEM + _HISTORY + _RIGID:
        if k not in names and ctx.has_array(k):
            names.append(k)
    return names


def save_checkpoint(ctx, path: str, extra: dict | None = None) -> None:
    """Every persistent array (incl. contact history) plus the DEVICE ORDER -> one .npz.

    Arrays are stored in id order together with `__order` (the stable id held by each device slot).  Restoring the
    device order matters for bit-exact resumption: the radix sort is stable, so the order inside a cell -- and with
    it the floating-point summation order of every pair loop -- depends on the order before the sort."""
    data = {k: ctx.download(k) for k in state_names(ctx)}
    data["__n"] = np.array([ctx.n], np.int64)
    data["__order"] = ctx.download("id")
    if ctx.has_array("body"):
        for k in _BODY_STATE:
            data["__body_" + k] = ctx.body_get(k)
    for k, v in (extra or {}).items():
        data["__extra_" + k] = np.asarray(v)
    np.savez(path, **data)


def load_checkpoint(ctx, path: str) -> dict:
    """Restore a checkpoint into a context created with the same configuration.  Returns the `extra` dict."""
    z = np.load(path)
    n = int(z["__n"][0])
    order = z["__order"].astype(np.uint32)
    ctx.set_count(n)                                  # identity order: uploads land slot by slot
    rigid = "__body_mass" in z.files
    if rigid and not ctx.has_array("body"):
        ctx.bodies_create(len(z["__body_mass"]))
    for k in z.files:
        if k.startswith("__"):
            continue
        if not ctx.has_array(k):
            raise KeyError(f"checkpoint array '{k}' does not exist in this context")
        ctx.upload(k, np.ascontiguousarray(z[k][..., order]))     # id order -> saved device order
    ctx.upload("id", order)                           # declare the ids; host arrays are id-ordered again from here on
    if rigid:
        # member-list ranges from `body`; the saved positions `bpos` (the order of every body sum) are kept.
        ctx.bodies_restore()
        # every write scatters x = X + R r0, v = V + w x r onto the members; after the last one the records are
        # complete and the scatter repeats the arithmetic that produced the saved particle state: bit-identical resume
        for k in _BODY_STATE:
            ctx.body_set(k, z["__body_" + k])
    return {k[len("__extra_"):]: z[k] for k in z.files if k.startswith("__extra_")}


def write_csv(ctx, path: str, fields=None) -> None:
    fields = fields or [k for k in state_names(ctx) if not k.startswith("hist_")]
    cols = [ctx.download(k) for k in fields]
    np.savetxt(path, np.column_stack(cols), delimiter=",", header=",".join(fields), comments="")


def write_vtk(ctx, path: str, scalars=("rho", "p", "m", "tag", "rad"), vectors=(("velocity", ("u", "v", "w")),)) -> None:
    """Legacy ASCII VTK polydata: one vertex per particle, point data for the requested fields that exist."""
    n = ctx.n
    x = ctx.download("x"); y = ctx.download("y")
    z = ctx.download("z") if ctx.has_array("z") else np.zeros(n, x.dtype)
    with open(path, "w") as f:
        f.write("# vtk DataFile Version 3.0\nprestige_b200 particles\nASCII\nDATASET POLYDATA\n")
        f.write(f"POINTS {n} double\n")
        np.savetxt(f, np.column_stack([x, y, z]), fmt="%.17g")
        f.write(f"VERTICES {n} {2 * n}\n")
        np.savetxt(f, np.column_stack([np.ones(n, np.int64), np.arange(n, dtype=np.int64)]), fmt="%d")
        f.write(f"POINT_DATA {n}\n")
        for s in scalars:
            if ctx.has_array(s):
                a = ctx.download(s)
                f.write(f"SCALARS {s} {'int' if a.dtype.kind in 'iu' else 'double'} 1\nLOOKUP_TABLE default\n")
                np.savetxt(f, a, fmt="%d" if a.dtype.kind in "iu" else "%.17g")
        for name, comps in vectors:
            if all(ctx.has_array(c) for c in comps[:ctx.dim]):
                cols = [ctx.download(c) if ctx.has_array(c) else np.zeros(n) for c in comps]
                f.write(f"VECTORS {name} double\n")
                np.savetxt(f, np.column_stack(cols), fmt="%.17g")
