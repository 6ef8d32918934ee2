"""Round 2 A/B harness: for every library build given on the command line (paths to libprestige_b200*.so; "default" = the
in-tree one) run, in a fresh process, (1) parity of the default pair kernel against the oracle on two small 3D blocks and
(2) stage timings at 10 M particles (CUDA events on the context's stream would need torch; wall clock around pst_sync over
10 launches is within 1 % at these durations).  Torch-free.
    python scripts/r2_ab.py build_ab/libprestige_b200_dev.so build_ab/libprestige_b200_devold.so [--opt name=v ...] [--shape 200,200,250]
Output: gpurun_out/r2_ab.txt (appended)"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(shape, opts):
    sys.path.insert(0, ROOT)
    import numpy as np
    import prestige_b200 as pb
    from prestige_b200 import synth
    from oracle import oracle as orc

    def rel_err(got, ref):
        scale = np.sqrt(np.mean(ref.astype(np.float64) ** 2))
        return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), scale))) if ref.size else 0.0
    out = []
    for name, b in [("3d_small", synth.wcsph_block_3d(20, 18, 22).shuffled()), ("3d_mid", synth.wcsph_block_3d(40, 36, 50).shuffled())]:
        g = orc.make_grid(3, b.lo, b.hi, b.cell_size)
        ref = orc.wcsph(3, b.params, b.arrays, grid=g)
        refp, _ = orc.pairs(3, b.arrays["x"], b.arrays["y"], b.arrays["z"], b.arrays["h"], grid=g)
        ctx = pb.context_for_block(b)
        try:
            for k, v in opts.items():
                ctx.set_option(k, v)
            ctx.load_block(b)
            ctx.build_neighbours()
            ctx.apply(["tait_eos", "continuity", "momentum"])
            errs = {k: rel_err(ctx.download(k), ref[k]) for k in ("au", "av", "aw", "arho")}
            same = np.array_equal(ctx.dump_pairs(0), refp)
        finally:
            ctx.close()
        out.append(f"parity {name}: pairs_equal={same} max_err={max(errs.values()):.2e}")
    b = synth.wcsph_block_3d(*shape)
    ctx = pb.context_for_block(b)
    try:
        for k, v in opts.items():
            ctx.set_option(k, v)
        ctx.load_block(b)

        def stage(f, reps=10):
            ctx.sync()
            t0 = time.perf_counter()
            for _ in range(reps):
                f()
            ctx.sync()
            return (time.perf_counter() - t0) / reps * 1e3
        for _ in range(2):
            ctx.build_neighbours(); ctx.apply(["tait_eos"]); ctx.apply(["continuity", "momentum"])
        ctx.sync()
        t_n = stage(ctx.build_neighbours)
        t_e = stage(lambda: ctx.apply(["tait_eos"]))
        t_f = stage(lambda: ctx.apply(["continuity", "momentum"]))
        au = ctx.download("au")
        out.append(f"timing {b.n} particles {opts}: nnps {t_n:.3f} ms  eos {t_e:.3f} ms  pair kernel {t_f:.3f} ms  (sum au {float(np.sum(au)):.6e}, |au| {float(np.sqrt(np.mean(au ** 2))):.6e})")
    finally:
        ctx.close()
    print("\n".join(out), flush=True)


def main():
    args = sys.argv[1:]
    if args and args[0] == "--child":
        shape = tuple(int(v) for v in args[1].split(","))
        opts = {a.split("=")[0]: int(a.split("=")[1]) for a in args[2:]}
        child(shape, opts)
        return
    shape, opts, libs = "200,200,250", [], []
    it = iter(args)
    for a in it:
        if a == "--shape": shape = next(it)
        elif a == "--opt": opts.append(next(it))
        else: libs.append(a)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r2_ab.txt"), "a") as f:
        for lib in libs or ["default"]:
            env = dict(os.environ)
            if lib != "default":
                env["PRESTIGE_B200_LIB"] = os.path.join(ROOT, lib)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", shape, *opts], env=env, capture_output=True, text=True)
            txt = f"=== {lib} {' '.join(opts)}\n{r.stdout}{r.stderr[-2000:] if r.returncode else ''}"
            print(txt, flush=True)
            f.write(txt + "\n")


if __name__ == "__main__":
    main()
