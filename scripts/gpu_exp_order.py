"""Experiment (one GPU shot, torch-free): how sensitive is the fused pair kernel to the ORDER OF PARTICLES INSIDE A CELL?
The stable cell sort keeps the previous order inside a cell, so pre-ordering the host arrays selects it:
lattice (the bench's order), z-sorted, Morton sub-cell order, random.  Also sweeps the existing tile options.
Wall-clock timing around pst_sync over several kernel runs (kernel ~ms, host overhead ~10 us)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import prestige_b200 as pb  # noqa: E402
from prestige_b200 import synth  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "exp_order.txt"), "w")


def say(*a):
    msg = " ".join(str(x) for x in a)
    print(msg, flush=True)
    out.write(msg + "\n"); out.flush()


nx, ny, nz = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (160, 160, 200)))
t0 = time.time()
blk = synth.wcsph_block_3d(nx, ny, nz)
n = blk.n
say(f"block {nx}x{ny}x{nz} = {n} particles, generated in {time.time() - t0:.1f} s")
a = blk.arrays
cell = blk.cell_size
cx = np.floor(a["x"] / cell).astype(np.int64); cy = np.floor(a["y"] / cell).astype(np.int64); cz = np.floor(a["z"] / cell).astype(np.int64)
key = (cx * (cy.max() + 1) + cy) * (cz.max() + 1) + cz
fx, fy, fz = a["x"] / cell - cx, a["y"] / cell - cy, a["z"] / cell - cz


def spread(v):   # 2 bits -> every 3rd bit
    return (v & 1) | ((v & 2) << 2)


q = lambda f: np.minimum((f * 4).astype(np.int64), 3)
sub = (spread(q(fx)) << 2) | (spread(q(fy)) << 1) | spread(q(fz))        # 4x4x4 Morton sub-cell
rng = np.random.default_rng(1)
orders = {
    "lattice (bench order)": None,
    "z inside cell": np.lexsort((a["z"], key)),
    "morton 4x4x4 inside cell": np.lexsort((sub, key)),
    "x then z inside cell": np.lexsort((a["z"], q(fx), key)),
    "random inside cell": np.lexsort((rng.random(n), key)),
}


def timed(ctx, names, reps=6):
    ctx.apply(names); ctx.sync()
    t = time.perf_counter()
    for _ in range(reps):
        ctx.apply(names)
    ctx.sync()
    return (time.perf_counter() - t) / reps * 1e3


ref = None
for name, order in orders.items():
    b = blk if order is None else synth.Block(blk.name, 3, blk.physics, {k: v[order] for k, v in a.items()}, blk.params, blk.lo, blk.hi, blk.cell_size, 0, blk.meta)
    with pb.context_for_block(b) as ctx:
        ctx.load_block(b)
        ctx.build_neighbours()
        ctx.apply(["tait_eos", "continuity", "momentum"]); ctx.sync()
        ms = timed(ctx, ["continuity", "momentum"])
        au = ctx.download("au")
        if order is not None:
            inv = np.empty(n, np.int64); inv[order] = np.arange(n)
            au = au[inv]
        if ref is None:
            ref = au
        err = float(np.max(np.abs(au - ref)) / np.sqrt(np.mean(ref * ref)))
        say(f"{name:28s} pair kernel {ms:7.3f} ms  ({n / ms * 1e-6:.3f} G particles/s)  max|au - au_lattice|/rms = {err:.1e}")
        if order is None:
            base = ms
            for opt, vals in (("tile_lcap", (64, 96, 112)), ("tile_ta", (3,)), ("tile_g", (3, 4, 5, 6)), ("force_kernel", (1,))):
                for v in vals:
                    ctx.set_option(opt, v)
                    try:
                        say(f"    option {opt} = {v}: {timed(ctx, ['continuity', 'momentum'], 4):7.3f} ms")
                    except Exception as e:   # noqa: BLE001
                        say(f"    option {opt} = {v}: {e}")
                    ctx.set_option(opt, {"tile_lcap": 80, "tile_ta": 2, "tile_g": 0, "force_kernel": 2}[opt])
say("done")
