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


def _slab_context(whole, rank, world, lr, extra_cap=0, physics="wcsph", max_contacts=0, zsub=None):
    cell = whole.cell_size
    n_layers = int(np.ceil((whole.hi[0] - whole.lo[0]) / cell))
    first, k = decomp.split_layers(n_layers, world)[rank]
    lo, hi = whole.lo[0] + first * cell, whole.lo[0] + (first + k) * cell
    own = decomp.owner_mask(whole.arrays["x"], lo, hi, rank == 0, rank == world - 1)
    n = int(own.sum())
    ctx = pb.Context(dim=3, lo=(lo, whole.lo[1], whole.lo[2]), hi=(hi, whole.hi[1], whole.hi[2]), cell_size=cell, capacity=n + extra_cap + 16,
                     physics=physics, max_contacts=max_contacts, device=lr,
                     ghost_capacity=decomp.ghost_capacity(24, 24, cell, whole.meta["dx"]))
    if zsub is not None:
        ctx.set_option("zsub", zsub)             # before the communicator: the halo windows are sized by the cell layer
    uid = [pb.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], rank, world)
    ctx.set_count(n)
    ctx.set_params(**whole.params)
    for c, v in whole.arrays.items():
        ctx.upload(c, np.ascontiguousarray(v[own]))
    ctx.upload("id", np.nonzero(own)[0].astype(np.uint32))      # distributed mode: global labels, device-order transfers
    return ctx, own


def halo_parity(out, rank, world, lr):
    whole = synth.wcsph_block_3d(60, 24, 20)
    res = {}
    ref = orc.wcsph(3, whole.params, whole.arrays, grid=orc.make_grid(3, whole.lo, whole.hi, whole.cell_size))
    # "3g" / "2g": the tiled kernels with m (and h) as constants (every rank uploaded the same values); 1 and 2 cut whole cells
    for variant in (3, "3g", 2, "2g", 1, 0):
        fk = int(str(variant)[0])
        ctx, own = _slab_context(whole, rank, world, lr, zsub=1 if fk in (1, 2) else None)
        ctx.set_option("force_kernel", fk)
        if str(variant).endswith("g"):
            ctx.set_option("uniform_mass_global", 1)
        for rep in range(2):                     # second pass: identity re-sort + fresh halo
            ctx.build_neighbours()
            ctx.halo_exchange()
            ctx.apply(["tait_eos", "continuity", "momentum"])
        gid = ctx.download("id").astype(np.int64)
        assert np.array_equal(np.sort(gid), np.nonzero(own)[0])
        res[variant] = {c: rel_err(ctx.download(c), ref[c][gid]) for c in ("au", "av", "aw", "arho", "p")}
        res[variant]["ghosts"] = (ctx.stat("n_ghost_l"), ctx.stat("n_ghost_r"))
        n = ctx.n
        ctx.close()
    tot = torch.tensor([n], device="cuda")
    dist.all_reduce(tot)
    with open(os.path.join(out, f"rank{rank}.json"), "w") as f:
        json.dump({"rank": rank, "world": world, "n": n, "n_total": int(tot[0]), "n_whole": whole.n, "err": {str(k): v for k, v in res.items()}}, f)


def migration(out, rank, world, lr):
    """A block drifting along +x through the slab faces: pst_step on `world` ranks (re-sort, MIGRATION, halo, forces,
    integrate every step) against the same steps on one GPU; particles are matched by global id."""
    whole = synth.wcsph_block_3d(48, 24, 20)
    whole.params["gz"] = 0.0
    whole.arrays["u"] = whole.arrays["u"] + 3.0
    dt, steps = 4e-5, 120                        # drift 3 m/s * 4.8 ms = 1.2 cells
    ctx, own = _slab_context(whole, rank, world, lr, extra_cap=whole.n // 2)
    n0 = ctx.n
    ctx.step(dt, steps)
    ctx.sync()
    cnt = ctx.refresh_count()
    gid = ctx.download("id").astype(np.int64)
    got = {c: ctx.download(c) for c in ("x", "y", "z", "u", "rho")}
    ctx.close()
    # single-GPU truth on this rank's device
    with pb.context_for_block(whole, device=lr) as one:
        one.load_block(whole)
        one.step(dt, steps)
        ref = {c: one.download(c) for c in got}
    err = {c: float(np.max(np.abs(got[c] - ref[c][gid]) / max(np.abs(ref[c]).max(), 1e-300))) if len(gid) else 0.0 for c in got}
    allg = [None] * world
    dist.all_gather_object(allg, gid.tolist())
    flat = np.sort(np.concatenate([np.array(g, dtype=np.int64) for g in allg]))
    with open(os.path.join(out, f"rank{rank}.json"), "w") as f:
        json.dump({"rank": rank, "world": world, "n0": n0, "n1": int(cnt), "moved": int(n0 != cnt), "err": err,
                   "all_ids_once": bool(np.array_equal(flat, np.arange(whole.n)))}, f)


def coupled(out, rank, world, lr):
    """Coupled SPH-DEM across slab faces: spheres and fluid drift along +x, so contacts (with their history rows),
    the signed SPH mass and the tag mask all have to work on ghosts and survive migration.  One evaluation against
    the oracle, then pst_step on `world` ranks against the same steps on one GPU, matched by global id."""
    whole = synth.coupled_block_3d(48, 16, 18)
    whole.params["gz"] = 0.0
    mob = whole.arrays["tag"] != 1
    whole.arrays["u"] = whole.arrays["u"] + np.where(mob, 3.0, 0.0)
    fields = ("au", "av", "aw", "arho", "fx", "fy", "fz", "tx", "ty", "tz")
    ref, _, _ = orc.coupled(whole.params, whole.max_contacts, whole.arrays, grid=orc.make_grid(3, whole.lo, whole.hi, whole.cell_size))
    ctx, own = _slab_context(whole, rank, world, lr, extra_cap=whole.n // 2, physics="wcsph+dem", max_contacts=whole.max_contacts)
    ctx.build_neighbours()
    ctx.halo_exchange()
    ctx.apply(["tait_eos", "continuity", "momentum", "dem_contact"])
    gid = ctx.download("id").astype(np.int64)
    err1 = {c: rel_err(ctx.download(c), ref[c][gid]) for c in fields}
    n0 = ctx.n
    dt, steps = 1e-5, 200                        # drift 3 m/s * 2 ms = half a cell
    ctx.step(dt, steps)
    ctx.sync()
    cnt = ctx.refresh_count()
    gid = ctx.download("id").astype(np.int64)
    state = ("x", "y", "z", "u", "rho", "wx", "wz")
    got = {c: ctx.download(c) for c in state}
    hn = ctx.download("hist_n")
    ctx.close()
    with pb.context_for_block(whole, device=lr) as one:
        one.load_block(whole)
        one.build_neighbours()
        one.apply(["tait_eos", "continuity", "momentum", "dem_contact"])   # same call sequence as the slab ranks
        one.step(dt, steps)
        ref2 = {c: one.download(c) for c in state}
        hn_ref = one.download("hist_n")
    err = {c: float(np.max(np.abs(got[c] - ref2[c][gid]) / max(np.abs(ref2[c]).max(), 1e-300))) if len(gid) else 0.0 for c in state}
    allg = [None] * world
    dist.all_gather_object(allg, gid.tolist())
    flat = np.sort(np.concatenate([np.array(g, dtype=np.int64) for g in allg]))
    with open(os.path.join(out, f"rank{rank}.json"), "w") as f:
        json.dump({"rank": rank, "world": world, "n0": n0, "n1": int(cnt), "moved": int(n0 != cnt), "err1": err1, "err": err,
                   "hist_equal": bool(np.array_equal(hn, hn_ref[gid])), "contacts": int(hn.sum()),
                   "all_ids_once": bool(np.array_equal(flat, np.arange(whole.n)))}, f)


def main():
    out, mode = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "halo")
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    {"halo": halo_parity, "migration": migration, "coupled": coupled}[mode](out, rank, world, lr)
    dist.barrier(device_ids=[lr])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
