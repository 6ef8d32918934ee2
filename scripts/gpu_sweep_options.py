"""Torch-free sweep of the fused pair kernel's run-time knobs (pst_set_option): hit-list capacity `tile_lcap` (smaller lists
leave more L1 for the phase-2 gathers and balance the lanes of a warp -- every lane drains at exactly lcap hits -- at the
price of more drains), tile depth `tile_g` and the 2x2 / 2x3 column shape `tile_ta`.  Results are checked against the
default configuration (same neighbour sums up to summation order).
usage: python scripts/gpu_sweep_options.py [nx ny nz]     (default 128^3 = 2.1 M particles)  -> gpurun_out/sweep_options.txt
NOT YET RUN (written after the round's GPU minutes were spent)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import prestige_b200 as pb  # noqa: E402
from prestige_b200 import synth  # noqa: E402

nx, ny, nz = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (128, 128, 128)))
blk = synth.wcsph_block_3d(nx, ny, nz)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "sweep_options.txt"), "a")


def say(msg):
    print(msg, flush=True)
    out.write(msg + "\n"); out.flush()


def timed(ctx, reps=10):
    names = ["continuity", "momentum"]
    ctx.apply(names); ctx.sync()
    t = time.perf_counter()
    for _ in range(reps):
        ctx.apply(names)
    ctx.sync()
    return (time.perf_counter() - t) / reps * 1e3


with pb.context_for_block(blk) as ctx:
    ctx.load_block(blk)
    ctx.build_neighbours()
    ctx.apply(["tait_eos", "continuity", "momentum"]); ctx.sync()
    base_ms = timed(ctx)
    base = ctx.download("au")
    scale = float(np.sqrt(np.mean(base ** 2)))
    say(f"{blk.n} particles f64, default options: pair kernel {base_ms:.3f} ms")
    sweeps = [("tile_lcap", v, {}) for v in (24, 32, 40, 48, 64, 96, 112)] + [("tile_g", v, {}) for v in (2, 3, 4, 6, 8)] + \
             [("tile_ta", 3, {}), ("tile_ta", 3, {"tile_lcap": 48})]
    for name, v, extra in sweeps:
        for k, x in {name: v, **extra}.items():
            ctx.set_option(k, x)
        try:
            ms = timed(ctx)
            err = float(np.max(np.abs(ctx.download("au") - base))) / scale
            say(f"  {name} = {v} {extra or ''}: {ms:.3f} ms ({(ms / base_ms - 1) * 100:+.1f} %), max |d au| / rms = {err:.1e}")
        except pb.PstError as e:
            say(f"  {name} = {v} {extra or ''}: refused ({e})")
        for k in {name: v, **extra}:
            ctx.set_option(k, 0 if k in ("tile_g",) else {"tile_lcap": 80, "tile_ta": 2}.get(k, 0))
