"""A/B of the fused pair kernel's phase 1 (torch-free, one process, both builds dlopen'ed side by side):
  base   build_ab/libprestige_b200_aos.so  (make -C prestige_b200/csrc variant NAME=aos DEFS=-DPST_P1_AOS)
         float4 (x, y, z, index) per staged candidate, scalar FADD/FMUL/FFMA -- the layout that shipped until this A/B
  p1soa  prestige_b200/libprestige_b200.so  (the default build)  SoA staging, packed FADD2/FMUL2/FFMA2, two candidates at a time
(When profiles/r1_exp_p1soa_ab.txt was recorded the roles were reversed: the default build was the float4 one and the SoA
form was built with a define.)  Checks whether the results are bit-identical (f64 and f32, 3D, 2D and a coupled block)
and prints ms per launch.
Also times the stages of a coupled step with the dummy-particle wall pressure (default build).
Output: gpurun_out/exp_p1soa.txt"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import prestige_b200 as pb  # noqa: E402
from prestige_b200 import _lib as L, synth  # noqa: E402

LIBS = [("base", os.path.join(ROOT, "build_ab", "libprestige_b200_aos.so")), ("p1soa", os.path.join(ROOT, "prestige_b200", "libprestige_b200.so"))]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "exp_p1soa.txt"), "w")


def say(msg):
    print(msg, flush=True)
    out.write(msg + "\n"); out.flush()


def use(path):
    L._lib = None
    L.LIB_PATH = path
    L.load()


def timed(ctx, names, reps):
    ctx.apply(names); ctx.sync()
    t = time.perf_counter()
    for _ in range(reps):
        ctx.apply(names)
    ctx.sync()
    return (time.perf_counter() - t) / reps * 1e3


def run(block, real, names, outs, reps=0):
    b = block.astype(real)
    with pb.context_for_block(b, real=real) as ctx:
        ctx.load_block(b)
        ctx.build_neighbours()
        ctx.apply(names); ctx.sync()
        res = {k: ctx.download(k) for k in outs}
        ms = timed(ctx, [n for n in names if n in ("continuity", "momentum")], reps) if reps else 0.0
    return res, ms


big = synth.wcsph_block_3d(128, 128, 128)
cases = [("wcsph3d 2.1M f64", big, np.float64, 12), ("wcsph3d 2.1M f32", big, np.float32, 12),
         ("wcsph3d small f64 (ragged tiles)", synth.wcsph_block_3d(23, 17, 29).shuffled(), np.float64, 0),
         ("dam break 2D f64", synth.wcsph_dambreak_2d(dx=0.02).shuffled(), np.float64, 0),
         ("dam break 2D f32", synth.wcsph_dambreak_2d(dx=0.02).shuffled(), np.float32, 0),
         ("coupled small f64", synth.coupled_block_3d(14, 12, 15).shuffled(), np.float64, 0)]
results = {}
for tag, path in LIBS:
    if not os.path.exists(path):
        say(f"[{tag}] {path} missing, skipped")
        continue
    use(path)
    for name, blk, real, reps in cases:
        eqs = ["tait_eos", "continuity", "momentum"]
        outs = ["au", "av", "arho"] + (["aw"] if blk.dim == 3 else [])
        res, ms = run(blk, real, eqs, outs, reps)
        results[(tag, name)] = res
        if reps:
            say(f"[{tag}] {name}: pair kernel {ms:.3f} ms  ({blk.n / ms / 1e6:.3f} G particles/s)")
for name, *_ in cases:
    if ("base", name) in results and ("p1soa", name) in results:
        a, b = results[("base", name)], results[("p1soa", name)]
        same = all(np.array_equal(a[k], b[k]) for k in a)
        worst = max(float(np.max(np.abs(a[k].astype(np.float64) - b[k].astype(np.float64))) / (np.sqrt(np.mean(a[k].astype(np.float64) ** 2)) + 1e-300)) for k in a)
        say(f"{name}: p1soa vs base {'bit-identical' if same else 'DIFFERENT, max rel %.2e' % worst}")

# ---- stage times of a coupled step with the dummy-particle wall pressure (default build)
use(LIBS[1][1])
c = synth.coupled_block_3d(125, 125, 128)
c.params["boundary_model"] = 1.0
with pb.context_for_block(c) as ctx:
    ctx.load_block(c)
    ctx.build_neighbours()
    ctx.apply(["tait_eos", "wall_pressure", "continuity", "momentum", "dem_contact"]); ctx.sync()
    n_dummy = int((c.arrays["tag"] != 0).sum())
    say(f"coupled {c.n} particles ({n_dummy} dummy): eos {timed(ctx, ['tait_eos'], 10):.3f} ms, eos+wall_pressure "
        f"{timed(ctx, ['tait_eos', 'wall_pressure'], 10):.3f} ms, eos+pair {timed(ctx, ['tait_eos', 'continuity', 'momentum'], 5):.3f} ms, "
        f"contact {timed(ctx, ['dem_contact'], 5):.3f} ms")
