"""A/B of two builds of the library on the fused pair kernel (torch-free; PRESTIGE_B200_LIB selects the build).
usage: python scripts/gpu_exp_ab.py TAG [nx ny nz]  -> prints ms per launch, saves au/arho to /tmp/ab_TAG.npz"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import prestige_b200 as pb  # noqa: E402
from prestige_b200 import synth  # noqa: E402

tag = sys.argv[1]
nx, ny, nz = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (160, 160, 160)))
blk = synth.wcsph_block_3d(nx, ny, nz)


def timed(ctx, names, reps=8):
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
    line = f"[{tag}] {blk.n} particles: pair kernel {timed(ctx, ['continuity', 'momentum']):.3f} ms"
    np.savez(f"/tmp/ab_{tag}.npz", au=ctx.download("au"), aw=ctx.download("aw"), arho=ctx.download("arho"))
    if tag != "base":
        ctx.set_option("uniform_mass", 0)
        line += f" | uniform_mass=0: {timed(ctx, ['continuity', 'momentum']):.3f} ms"
        ctx.set_option("uniform_mass", 1)
    t = time.perf_counter()
    for _ in range(5):
        ctx.build_neighbours(); ctx.apply(["tait_eos", "continuity", "momentum"])
    ctx.sync()
    line += f" | full step {(time.perf_counter() - t) / 5 * 1e3:.3f} ms"
    print(line, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "exp_ab.txt"), "a").write(line + "\n")
if tag != "base" and os.path.exists("/tmp/ab_base.npz"):
    a, b = np.load("/tmp/ab_base.npz"), np.load(f"/tmp/ab_{tag}.npz")
    msg = "results vs base: " + ", ".join(f"{k} {'bit-identical' if np.array_equal(a[k], b[k]) else 'DIFFERENT max rel %.2e' % (np.max(np.abs(a[k] - b[k])) / np.sqrt(np.mean(a[k] ** 2)))}" for k in a.files)
    print(msg, flush=True)
    open(os.path.join(ROOT, "gpurun_out", "exp_ab.txt"), "a").write(msg + "\n")
