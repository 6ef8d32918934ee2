"""torchrun worker for the multi-GPU parity test: x-slab decomposition + NCCL halo exchange on real GPUs.
Every rank owns a whole number of cell layers of one jittered WCSPH block, runs
build_neighbours -> halo_exchange -> EOS -> fused pair kernel, and compares its owned particles with the
single-domain oracle.  Usage: torchrun --nproc-per-node N tests/multi_gpu_worker.py OUTDIR"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import prestige_b200 as pb                      # noqa: E402
from prestige_b200 import decomp, synth         # noqa: E402
from oracle import oracle as orc                # noqa: E402
from util import rel_err                        # noqa: E402


def main():
    out = sys.argv[1]
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    whole = synth.wcsph_block_3d(60, 24, 20)
    cell = whole.cell_size
    n_layers = int(np.ceil((whole.hi[0] - whole.lo[0]) / cell))
    first, k = decomp.split_layers(n_layers, world)[rank]
    lo, hi = whole.lo[0] + first * cell, whole.lo[0] + (first + k) * cell
    own = decomp.owner_mask(whole.arrays["x"], lo, hi, rank == 0, rank == world - 1)
    mine = {c: np.ascontiguousarray(v[own]) for c, v in whole.arrays.items()}
    n = int(own.sum())
    res = {}
    for variant in (1, 0):
        ctx = pb.Context(dim=3, lo=(lo, whole.lo[1], whole.lo[2]), hi=(hi, whole.hi[1], whole.hi[2]), cell_size=cell, capacity=n + 16,
                         physics="wcsph", device=lr, ghost_capacity=decomp.ghost_capacity(24, 20, cell, whole.meta["dx"]))
        uid = [pb.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
        ctx.set_option("force_kernel", variant)
        ctx.set_count(n)
        ctx.set_params(**whole.params)
        for c, v in mine.items():
            ctx.upload(c, v)
        for rep in range(2):                     # second pass: identity re-sort + fresh halo
            ctx.build_neighbours()
            ctx.halo_exchange()
            ctx.apply(["tait_eos", "continuity", "momentum"])
        got = {c: ctx.download(c) for c in ("au", "av", "aw", "arho", "p")}
        ghosts = (ctx.stat("n_ghost_l"), ctx.stat("n_ghost_r"))
        ctx.close()
        ref = orc.wcsph(3, whole.params, whole.arrays, grid=orc.make_grid(3, whole.lo, whole.hi, cell))
        res[variant] = {c: rel_err(got[c], ref[c][own]) for c in got}
        res[variant]["ghosts"] = ghosts
    tot = torch.tensor([n], device="cuda")
    dist.all_reduce(tot)
    with open(os.path.join(out, f"rank{rank}.json"), "w") as f:
        json.dump({"rank": rank, "world": world, "n": n, "n_total": int(tot[0]), "n_whole": whole.n, "err": {str(k): v for k, v in res.items()}}, f)
    dist.barrier(device_ids=[lr])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
