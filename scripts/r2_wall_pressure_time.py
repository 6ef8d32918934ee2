import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import prestige_b200 as pb
from prestige_b200 import synth
b = synth.coupled_block_3d(125, 125, 128)
ctx = pb.context_for_block(b)
ctx.load_block(b)
ctx.set_params(boundary_model=1)
ctx.build_neighbours(); ctx.apply(["tait_eos", "wall_pressure", "continuity", "momentum"]); ctx.sync()
def t(f, reps=10):
    ctx.sync(); t0 = time.perf_counter()
    for _ in range(reps): f()
    ctx.sync(); return (time.perf_counter() - t0) / reps * 1e3
te = t(lambda: ctx.apply(["tait_eos"]))
tw = t(lambda: ctx.apply(["tait_eos", "wall_pressure"]))
print(f"coupled 2M: eos {te:.3f} ms, eos + wall_pressure {tw:.3f} ms -> wall pressure {tw - te:.3f} ms; dummies {(b.arrays['tag'] != 0).sum()}")
