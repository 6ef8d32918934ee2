"""Small end-to-end pass of every kernel for compute-sanitizer (memcheck / racecheck / initcheck): the round-2 defaults (pair-kernel
variant 3 with TMA staging, the fused permute + EOS, pst_step as a CUDA graph) and every selectable variant beside them."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import prestige_b200 as pb
from prestige_b200 import synth

light = "--light" in sys.argv          # racecheck is ~100x slower: the defaults only
for real in (np.float64, np.float32):
    b = synth.wcsph_block_3d(14, 11, 13).shuffled().astype(real)
    variants = [(3, {}), (3, {"rec_impl": 0}), (3, {"fuse_eos": 0}), (3, {"tile_jcap": 60}), (3, {"tile_words": 4, "uniform_mass": 0})]
    if not light:
        variants += [(3, {"zsub": 1}), (3, {"zsub": 8}), (2, {"zsub": 1}), (1, {"zsub": 1}), (0, {})]
    for variant, opts in variants:
        for sort_impl in ((1,) if light else (1, 0)):
            with pb.context_for_block(b, real=real) as ctx:
                for k, v in opts.items():
                    ctx.set_option(k, v)
                ctx.load_block(b)
                ctx.set_option("force_kernel", variant); ctx.set_option("sort_impl", sort_impl)
                ctx.build_neighbours(); ctx.apply(["tait_eos", "continuity", "momentum"])
                ctx.dump_pairs(0)
                ctx.step(1e-5, 2)
                ctx.download("au")
    with pb.context_for_block(b, real=real) as ctx:         # pst_step as a CUDA graph
        ctx.load_block(b); ctx.set_option("graph", 1); ctx.step(1e-5, 7); ctx.download("x")
    c = synth.wcsph_dambreak_2d(dx=0.05).shuffled().astype(real)
    for variant, opts in ((3, {}), (3, {"rec_impl": 0})) if light else ((3, {}), (3, {"rec_impl": 0}), (2, {"zsub": 1}), (1, {"zsub": 1}), (0, {})):
        with pb.context_for_block(c, real=real) as ctx:
            for k, v in opts.items():
                ctx.set_option(k, v)
            ctx.load_block(c); ctx.set_option("force_kernel", variant)
            ctx.set_params(boundary_model=1)
            ctx.build_neighbours(); ctx.apply(["tait_eos", "wall_pressure", "continuity", "momentum"]); ctx.download("au")
    cp = synth.coupled_block_3d(10, 9, 10).shuffled().astype(real)
    with pb.context_for_block(cp, real=real) as ctx:
        ctx.load_block(cp); ctx.step(1e-5, 3); ctx.download("fx")
    d = synth.dem_column_3d(7).shuffled().astype(real)
    for key in ("linear",) if light else ("linear", "morton"):
        with pb.context_for_block(d, real=real, key=key) as ctx:
            ctx.load_block(d)
            ctx.step(2e-6, 3)
            ctx.download("hist_x"); ctx.build_neighbours(); ctx.dump_pairs(1)
with pb.Context(dim=3, lo=(0, 0, 0), hi=(1, 1, 1), cell_size=0.5, capacity=300) as ctx:
    ctx.set_count(300); ctx.array_create("force"); ctx.array_create("mass")
    ctx.upload("mass", np.ones(300)); ctx.upload("force", np.zeros(300)); ctx.apply(["eq1"]); ctx.download("force")
print("sanitize_small done")
