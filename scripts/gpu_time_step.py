"""Torch-free timing of the WCSPH hot path (wall clock around a stream sync; bench.py with CUDA events is the reference measurement).
usage: python scripts/gpu_time_step.py [nx ny nz]   (default 200 200 250 = the 10 M configuration)  -> gpurun_out/time_step.txt"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import prestige_b200 as pb  # noqa: E402
from prestige_b200 import synth  # noqa: E402

nx, ny, nz = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (200, 200, 250)))
t0 = time.perf_counter()
blk = synth.wcsph_block_3d(nx, ny, nz)
t_gen = time.perf_counter() - t0


def timed(fn, reps):
    fn(); ctx.sync()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    return (time.perf_counter() - t) / reps * 1e3


with pb.context_for_block(blk) as ctx:
    ctx.load_block(blk)
    ctx.build_neighbours()
    ctx.apply(["tait_eos", "continuity", "momentum"]); ctx.sync()
    pair = timed(lambda: ctx.apply(["continuity", "momentum"]), 6)

    def step():
        ctx.build_neighbours()
        ctx.apply(["tait_eos", "continuity", "momentum"])
    full = timed(step, 6)
    line = (f"{blk.n} particles f64: fused pair kernel {pair:.3f} ms, full step (keys+sort+table+permute+EOS+pair) {full:.3f} ms = "
            f"{blk.n / full / 1e6:.3f} G particle-steps/s (block generated in {t_gen:.1f} s)")
print(line, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "time_step.txt"), "a").write(line + "\n")
