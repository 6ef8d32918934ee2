"""world_size = 2 on CPU (gloo): the slab decomposition + ghost-layer protocol of the multi-GPU path.

Each rank owns a whole number of cell layers, ships its outermost owned layer to its neighbour (what
csrc/halo.cu does with ncclSend/ncclRecv), and evaluates forces for its owned particles over
owned + ghost candidates.  The oracle is the checker: per-rank results must equal the single-domain
evaluation, and the neighbour set must be identical.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _exchange(rank, world, send_left, send_right):
    """Send edge layers to the slab neighbours, return (from_left, from_right).  Counts first, like halo.cu."""
    def xfer(peer, payload):
        cnt = torch.tensor([payload.shape[0]], dtype=torch.int64)
        rcnt = torch.zeros(1, dtype=torch.int64)
        ops = [dist.P2POp(dist.isend, cnt, peer), dist.P2POp(dist.irecv, rcnt, peer)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        buf = torch.zeros((int(rcnt[0]), payload.shape[1]), dtype=torch.float64)
        ops = [dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(payload)), peer), dist.P2POp(dist.irecv, buf, peer)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return buf.numpy()
    fl = xfer(rank - 1, send_left) if rank > 0 else np.zeros((0, send_left.shape[1]))
    fr = xfer(rank + 1, send_right) if rank + 1 < world else np.zeros((0, send_right.shape[1]))
    return fl, fr


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from prestige_b200 import decomp, synth
    try:
        orc.set_num_threads(2)
        whole = synth.wcsph_block_3d(24, 10, 9)
        cell = whole.cell_size
        n_layers = int(np.ceil((whole.hi[0] - whole.lo[0]) / cell))
        first, k = decomp.split_layers(n_layers, world)[rank]
        lo, hi = whole.lo[0] + first * cell, whole.lo[0] + (first + k) * cell
        own = decomp.owner_mask(whole.arrays["x"], lo, hi, rank == 0, rank == world - 1)
        names = ["x", "y", "z", "u", "v", "w", "rho", "m", "h"]
        mine = np.stack([whole.arrays[c][own] for c in names], axis=1)
        ids = np.nonzero(own)[0]
        to_l, to_r = decomp.edge_layers(mine[:, 0], lo, cell, k)
        idcol = ids.astype(np.float64)[:, None]
        fl, fr = _exchange(rank, world, np.hstack([mine[to_l], idcol[to_l]]), np.hstack([mine[to_r], idcol[to_r]]))
        allp = np.vstack([mine, fl[:, :-1], fr[:, :-1]])
        gids = np.concatenate([ids, fl[:, -1].astype(np.int64), fr[:, -1].astype(np.int64)])
        a = {c: np.ascontiguousarray(allp[:, i]) for i, c in enumerate(names)}
        g = orc.make_grid(3, (lo - cell, whole.lo[1], whole.lo[2]), (hi + cell, whole.hi[1], whole.hi[2]), cell)
        res = orc.wcsph(3, whole.params, a, grid=g)
        pr, _ = orc.pairs(3, a["x"], a["y"], a["z"], a["h"], grid=g)
        n_own = len(ids)
        pr = pr[pr[:, 0] < n_own]                               # only owned i
        pr = np.stack([gids[pr[:, 0]], gids[pr[:, 1]]], axis=1)
        # single-domain truth
        ref = orc.wcsph(3, whole.params, whole.arrays)
        refp, _ = orc.pairs(3, whole.arrays["x"], whole.arrays["y"], whole.arrays["z"], whole.arrays["h"])
        refp = refp[own[refp[:, 0]]]
        ok_pairs = np.array_equal(pr[np.lexsort((pr[:, 1], pr[:, 0]))], refp)
        err = max(np.abs(res[c][:n_own] - ref[c][own]).max() / np.abs(ref[c]).max() for c in ("au", "av", "aw", "arho"))
        tot = torch.tensor([n_own], dtype=torch.int64)
        dist.all_reduce(tot)
        # what bench.py asks before it lets every rank skip the m[j] gather: one mass value on ALL ranks
        uni = (decomp.uniform_across_ranks(mine[:, 7], dist),                                   # the block's masses: all equal
               decomp.uniform_across_ranks(mine[:, 7] * (1.0 + (rank == world - 1)), dist),     # one rank differs: every rank must see it
               decomp.uniform_across_ranks(np.where(np.arange(n_own) == 0, 2.0, 1.0) if rank == 0 else np.ones(n_own), dist))
        q.put((rank, ok_pairs, float(err), int(tot[0]), whole.n, len(fl), len(fr), uni))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_decomposition_with_ghost_layers(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_pairs, err, tot, n, nl, nr, uni in res:
        assert uni == (True, False, False), f"rank {rank}: cross-rank mass uniformity check {uni}"
        assert ok_pairs, f"rank {rank}: neighbour set of owned particles differs from the single-domain set"
        assert err < 1e-12, f"rank {rank}: {err:.3e}"
        assert tot == n, "every particle is owned by exactly one rank"
        assert (nl > 0) == (rank > 0) and (nr > 0) == (rank < world - 1)


def test_split_layers_and_bounds():
    from prestige_b200 import decomp
    assert decomp.split_layers(10, 3) == [(0, 4), (4, 3), (7, 3)]
    assert sum(k for _, k in decomp.split_layers(83 * 8, 8)) == 83 * 8
    lo, hi = decomp.slab_bounds(0.0, 0.012, 83, 2)
    assert abs(lo - 2 * 83 * 0.012) < 1e-15 and abs(hi - 3 * 83 * 0.012) < 1e-15


@pytest.mark.parametrize("world", [2, 3, 8])
def test_bench_slabs_partition_the_block(world):
    """bench.py cuts the SAME lattice block into `world` x-slabs of whole cell layers (strong scaling): every particle of the
    block belongs to exactly one slab, slabs are contiguous in x and tile the box, and each rank generates its slab alone
    (counter-based generator) with the values the whole block has."""
    sys.path.insert(0, ROOT)
    import bench
    from prestige_b200 import synth
    nx, ny, nz = 57, 9, 11
    whole = synth.wcsph_block_3d(nx, ny, nz, dx=bench.DX)
    seen = np.zeros(whole.n, dtype=np.int64)
    prev_hi = 0.0
    for rank in range(world):
        b, (lo_x, hi_x) = bench._slab_of(synth.wcsph_block_3d, nx, ny, nz, rank, world)
        assert lo_x == pytest.approx(prev_hi)
        prev_hi = hi_x
        ids = b.meta["ids"].astype(np.int64)
        seen[ids] += 1
        for k in ("x", "u", "rho"):
            assert np.array_equal(b.arrays[k], whole.arrays[k][ids])
        inside = (b.arrays["x"] >= lo_x) | (rank == 0)
        inside &= (b.arrays["x"] < hi_x) | (rank == world - 1)
        assert inside.all()
        cells = (hi_x - lo_x) / whole.cell_size
        assert abs(cells - round(cells)) < 1e-9          # whole cell layers: what pst_comm_init requires
    assert np.all(seen == 1)
    assert prev_hi >= nx * bench.DX
