#!/usr/bin/env python
"""2D WCSPH dam break (BASELINE.json configs[0]) through the public API: the equation front-end of the reference
(prestige/src/lib.rs:32-52 builds the IRs and fuses them) feeding the b200 back-end instead of generate_simple_cpu.

    python examples/dambreak2d.py [--steps 2000] [--dx 0.01] [--dt 2e-5] [--wall-pressure] [--out out/]

Needs a CUDA device (there is no CPU path).  Writes VTK snapshots and a checkpoint, prints mass / momentum / front position.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prestige_b200 as pb                      # noqa: E402
from prestige_b200 import io, synth             # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--dx", type=float, default=0.01)
    ap.add_argument("--dt", type=float, default=2e-5)
    ap.add_argument("--every", type=int, default=500, help="snapshot interval in steps")
    ap.add_argument("--wall-pressure", action="store_true", help="dummy-particle wall pressure (boundary_model = 1, DESIGN.md 4d)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel of pst_step from the host instead of replaying a CUDA graph")
    ap.add_argument("--out", default="dambreak2d_out")
    args = ap.parse_args()

    block = synth.wcsph_dambreak_2d(dx=args.dx)
    if args.wall_pressure:
        block.params["boundary_model"] = 1.0
    os.makedirs(args.out, exist_ok=True)

    # what the reference's test_fusion does (lib.rs:32-52): equation IRs -> fuse -> back-end
    fused = pb.fuse([pb.tait_eos.ir()] + ([pb.wall_pressure.ir()] if args.wall_pressure else []) + [pb.continuity.ir(), pb.momentum.ir()])
    print("fused set reads ", sorted(fused.reads))
    print("fused set writes", sorted(fused.writes))
    print("simple_cpu back-end would emit:\n" + pb.codegen.generate_simple_cpu(fused))

    fluid = block.arrays["tag"] == 0
    m = block.arrays["m"]
    with pb.context_for_block(block) as ctx:
        ctx.load_block(block)
        # one explicit evaluation through the back-end object, as a maintainer would call it ...
        ctx.build_neighbours()
        pb.codegen.b200.run(ctx, fused)
        print(f"{block.n} particles ({int(fluid.sum())} fluid), {len(ctx.dump_pairs(0))} neighbour pairs, "
              f"max |a| = {np.abs(ctx.download('av')).max():.3g} m/s^2")
        # ... then the stepping loop (re-sort -> EOS -> [wall pressure] -> fused pair kernel -> semi-implicit Euler); at this size a
        # step is launch-bound, so pst_step replays it as a CUDA graph (two steps per replay; bit-identical to eager launches)
        if not args.no_graph:
            ctx.set_option("graph", 1)
        t0 = time.perf_counter()
        for done in range(0, args.steps, args.every):
            k = min(args.every, args.steps - done)
            ctx.step(args.dt, k)
            x, u, v, rho = (ctx.download(c) for c in ("x", "u", "v", "rho"))
            print(f"step {done + k:6d}  t = {(done + k) * args.dt:.4f} s  front x = {x[fluid].max():.3f} m  "
                  f"mass = {m[fluid].sum():.6g}  p_x = {(m * u)[fluid].sum():+.4e}  p_y = {(m * v)[fluid].sum():+.4e}  "
                  f"rho in [{rho[fluid].min():.1f}, {rho[fluid].max():.1f}]")
            io.write_vtk(ctx, os.path.join(args.out, f"dambreak_{done + k:06d}.vtk"))
        ctx.sync()
        wall = time.perf_counter() - t0
        io.save_checkpoint(ctx, os.path.join(args.out, "checkpoint.npz"), extra={"steps": args.steps, "dt": args.dt})
        print(f"{args.steps} steps in {wall:.2f} s = {block.n * args.steps / wall / 1e6:.1f} M particle-steps/s (snapshots included); "
              f"{int(ctx.stat('launches'))} kernel launches")


if __name__ == "__main__":
    main()
