"""Round 2: variant 3 (k_wcsph_zrun) -- parity on small blocks for every zsub, then stage timings at bench size.
Torch-free (ctypes + numpy only).  Output: gpurun_out/r2_v3_quick.txt"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import prestige_b200 as pb  # noqa: E402
from prestige_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

OUT = open(os.path.join(ROOT, "gpurun_out", os.environ.get("R2_OUT", "r2_v3_quick.txt")), "w")


def say(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    OUT.write(s + "\n"); OUT.flush()


def rel_err(got, ref):
    scale = np.sqrt(np.mean(ref.astype(np.float64) ** 2))
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), scale))) if ref.size else 0.0


def run(block, real, variant, zsub, opts=None, names=("tait_eos", "continuity", "momentum")):
    b = block.astype(real)
    ctx = pb.context_for_block(b, real=real)
    try:
        ctx.set_option("zsub", zsub)
        ctx.set_option("force_kernel", variant)
        for k, v in (opts or {}).items():
            ctx.set_option(k, v)
        ctx.load_block(b)
        ctx.build_neighbours()
        ctx.apply(list(names))
        out = {k: ctx.download(k) for k in (["p", "au", "av", "arho"] + (["aw"] if b.dim == 3 else []))}
        pairs = ctx.dump_pairs(0)
    finally:
        ctx.close()
    return out, pairs


def parity():
    ok = True
    cases = [("3d", synth.wcsph_block_3d(20, 18, 22).shuffled(), 3), ("3d_big", synth.wcsph_block_3d(40, 36, 50).shuffled(), 3),
             ("2d", synth.wcsph_dambreak_2d(dx=0.02).shuffled(), 2)]
    for name, b, dim in cases:
        for real in (np.float64, np.float32):
            br = b.astype(real)
            a = br.arrays
            g = orc.make_grid(dim, br.lo, br.hi, br.cell_size)
            ref = orc.wcsph(dim, br.params, a, grid=g)
            refp, _ = orc.pairs(dim, a["x"], a["y"], a.get("z"), a["h"], grid=g)
            tol = 1e-10 if real == np.float64 else 1e-5
            for variant, zsub, opts in [(0, 4, None), (3, 1, None), (3, 4, None), (3, 8, {"rec_impl": 0}), (3, 4, {"rec_impl": 0}),
                                        (3, 4, {"tile_words": 4}), (3, 4, {"tile_g": 1}), (3, 4, {"tile_jcap": 100}), (3, 4, {"uniform_mass": 0}),
                                        (3, 2, {"uniform_mass": 0, "rec_impl": 0})]:
                try:
                    got, pairs = run(b, real, variant, zsub, opts)
                    errs = {k: rel_err(got[k], ref[k]) for k in got}
                    good = np.array_equal(pairs, refp) and max(errs.values()) < tol
                    ok &= good
                    say(f"{'ok  ' if good else 'FAIL'} {name} {np.dtype(real).name} variant {variant} zsub {zsub} {opts or ''} pairs_equal={np.array_equal(pairs, refp)} max_err={max(errs.values()):.2e}")
                except Exception as e:
                    ok = False
                    say(f"FAIL {name} {np.dtype(real).name} variant {variant} zsub {zsub} {opts}: {e}")
                    traceback.print_exc()
    return ok


def timing(shape=(200, 200, 250)):
    import ctypes as C
    from prestige_b200 import _lib
    lib = _lib.load()
    b = synth.wcsph_block_3d(*shape)
    say(f"timing block {shape}: {b.n} particles")
    base = None
    for variant, zsub, opts in [(3, 4, {}), (3, 4, {"tile_cta": 1}), (3, 4, {"tile_cta": 1, "tile_words": 16}), (3, 4, {"tile_cta": 2}), (3, 4, {"tile_cta": 3, "tile_words": 16})]:
        ctx = pb.context_for_block(b)
        try:
            ctx.set_option("zsub", zsub)
            ctx.set_option("force_kernel", variant)
            for k, v in opts.items():
                ctx.set_option(k, v)
            ctx.load_block(b)

            def stage(f, reps=5):
                ctx.sync()
                t0 = time.perf_counter()
                for _ in range(reps):
                    f()
                ctx.sync()
                return (time.perf_counter() - t0) / reps * 1e3
            ctx.build_neighbours(); ctx.apply(["tait_eos"]); ctx.apply(["continuity", "momentum"]); ctx.sync()
            t_n = stage(ctx.build_neighbours)
            t_e = stage(lambda: ctx.apply(["tait_eos"]))
            t_f = stage(lambda: ctx.apply(["continuity", "momentum"]))
            au = ctx.download("au")
            if base is None:
                base = au
            d = float(np.max(np.abs(au - base)) / np.sqrt(np.mean(base ** 2)))
            say(f"variant {variant} zsub {zsub} {opts}: nnps {t_n:.3f} ms  eos {t_e:.3f} ms  pair kernel {t_f:.3f} ms  (au vs variant 2: {d:.1e})")
        finally:
            ctx.close()


if __name__ == "__main__":
    what = sys.argv[1:] or ["parity", "timing"]
    good = True
    if "parity" in what:
        good = parity()
        say("PARITY", "ALL OK" if good else "FAILURES")
    if "timing" in what:
        timing()
    sys.exit(0 if good else 1)
