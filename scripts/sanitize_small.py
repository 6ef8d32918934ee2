"""Small end-to-end pass of every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import prestige_b200 as pb
from prestige_b200 import synth

for real in (np.float64, np.float32):
    b = synth.wcsph_block_3d(14, 11, 13).shuffled().astype(real)
    for variant in (2, 1, 0):
        for sort_impl in (1, 0):
            with pb.context_for_block(b, real=real) as ctx:
                ctx.load_block(b)
                ctx.set_option("force_kernel", variant); ctx.set_option("sort_impl", sort_impl)
                ctx.build_neighbours(); ctx.apply(["tait_eos", "continuity", "momentum"])
                ctx.dump_pairs(0)
                ctx.step(1e-5, 2)
                ctx.download("au")
    c = synth.wcsph_dambreak_2d(dx=0.05).shuffled().astype(real)
    for variant in (2, 1, 0):
        with pb.context_for_block(c, real=real) as ctx:
            ctx.load_block(c); ctx.set_option("force_kernel", variant)
            ctx.build_neighbours(); ctx.apply(["tait_eos", "continuity", "momentum"]); ctx.download("au")
    d = synth.dem_column_3d(7).shuffled().astype(real)
    for key in ("linear", "morton"):
        with pb.context_for_block(d, real=real, key=key) as ctx:
            ctx.load_block(d)
            ctx.step(2e-6, 3)
            ctx.download("hist_x"); ctx.build_neighbours(); ctx.dump_pairs(1)
with pb.Context(dim=3, lo=(0, 0, 0), hi=(1, 1, 1), cell_size=0.5, capacity=300) as ctx:
    ctx.set_count(300); ctx.array_create("force"); ctx.array_create("mass")
    ctx.upload("mass", np.ones(300)); ctx.upload("force", np.zeros(300)); ctx.apply(["eq1"]); ctx.download("force")
print("sanitize_small done")
