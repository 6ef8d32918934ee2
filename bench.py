#!/usr/bin/env python
"""bench.py -- particle-steps/sec of the NNPS + pair-force hot path on B200 (BASELINE.json metric).

One "step" = cell keys -> radix sort -> cell table -> permute state (-> history remap) -> EOS ->
fused continuity+momentum pair kernel (or the DEM contact kernel) over one block of synthetic
particles.  Integrator excluded.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload wcsph3d_10m|dem3d_1m|wcsph2d_20k|coupled3d_20m] [--real f64|f32]
    python bench.py --impl reference ...      # the CPU restatement (oracle/) on the host cores, same metric
    torchrun --nproc-per-node N bench.py --gpus N ...   # weak scaling: one ~10M-particle x-slab per rank
                                                        # (coupled3d_20m: STRONG scaling, the 20M block cut into N slabs)

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_steps_per_sec"
UNIT = "particle-steps/s"

# algorithmic (compulsory) bytes per particle-step, SURVEY.md 8d / BASELINE.md 4
B_ALG = {("wcsph", 3, "f64"): 350, ("wcsph", 3, "f32"): 202, ("wcsph", 2, "f64"): 286, ("wcsph", 2, "f32"): 166}
B_FORCE = {("wcsph", 3, "f64"): 117, ("wcsph", 3, "f32"): 61, ("wcsph", 2, "f64"): 93, ("wcsph", 2, "f32"): 49}


# algorithmic f64 flop per particle of the fused pair kernel (SURVEY.md 8d "FLOP figure for the co-bound"): ~8 flop per
# candidate of the 27 (9) cell stencil + ~70 flop per in-range pair, at h = 1.2 dx, cutoff 2h, cell = cutoff:
# 3D 373 candidates / 58 neighbours, 2D 52 / 18.  Peak: scripts/ubench/fp64_peak.cu, profiles/ubench_fp64_peak.txt.
FLOP_FORCE = {3: 8 * 373 + 70 * 58, 2: 8 * 52 + 70 * 18}
FP64_PEAK_TFLOPS = 36.5


def dem_bytes(real: str, zbar: float):
    """(step, force-pass) algorithmic bytes per particle for 3D DEM with mean stored contacts zbar."""
    if real == "f64":
        return 446 + 112 * zbar, 159 + 2 * (28 * zbar + 4)
    return 250 + 64 * zbar, 87 + 2 * (16 * zbar + 4)


def coupled_bytes(real: str, zbar_solid: float, f_solid: float):
    """Coupled SPH-DEM: particle-weighted sum of the WCSPH and DEM columns; boundaries count as SPH."""
    ds, df = dem_bytes(real, zbar_solid)
    ws, wf = B_ALG[("wcsph", 3, real)], B_FORCE[("wcsph", 3, real)]
    return (1 - f_solid) * ws + f_solid * ds, (1 - f_solid) * wf + f_solid * df


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
# what pst_halo_exchange does by default (csrc/halo.cu: halo_impl option / PST_HALO_IMPL): 2 = peer memory, 1 = packed NCCL, 0 = per-array NCCL
HALO_DESC = {"2": "peer-memory halo (pack kernel stores into the neighbour's cudaIpc buffer over NVLink, epoch flag, wait+unpack kernel)",
             "1": "one packed NCCL send/recv message per neighbour", "0": "per-array NCCL send/recv halo with count hand-shake"}.get(
                 os.environ.get("PST_HALO_IMPL", "2"), "peer-memory halo")
SLAB_CELLS = 83   # multi-GPU: every rank owns 83 cell layers (0.996 m, ~199.2 lattice planes, ~9.96 M particles)


def make_block(workload: str, rank: int = 0, world: int = 1):
    from prestige_b200 import synth
    if workload == "wcsph3d_10m":
        if world == 1:
            return synth.wcsph_block_3d(200, 200, 250, name="wcsph3d_10m"), None
        # weak scaling (configs[3] family): global tank of world*83 cell layers along x, slab per rank
        dx = 0.005
        cell = 2.0 * 1.2 * dx * synth.CELL_MARGIN
        lo_x, hi_x = rank * SLAB_CELLS * cell, (rank + 1) * SLAB_CELLS * cell
        nx_total = int(math.floor(world * SLAB_CELLS * cell / dx))
        p0 = max(0, int(math.floor(lo_x / dx)) - 1)
        p1 = min(nx_total, int(math.ceil(hi_x / dx)) + 1)
        b = synth.wcsph_block_3d(p1 - p0, 200, 250, dx=dx, ix0=p0, nx_total=nx_total, name="wcsph3d_10m_slab")
        keep = (b.arrays["x"] >= lo_x) & (b.arrays["x"] < hi_x)
        b.arrays = {k: np.ascontiguousarray(v[keep]) for k, v in b.arrays.items()}
        b.meta["ids"] = b.meta["ids"][keep]
        return b, (lo_x, hi_x)
    if workload.startswith("coupled3d_"):   # configs[4]: rigid spheres in fluid; N > 1 cuts the SAME block into x-slabs (strong scaling)
        from prestige_b200 import decomp
        nx, ny, nz = {"20m": (250, 250, 320), "2m": (125, 125, 128), "300k": (64, 64, 72)}[workload.split("_")[1]]
        if world == 1:
            return synth.coupled_block_3d(nx, ny, nz), None
        dx = 0.005
        cell = 2.0 * 1.2 * dx * synth.CELL_MARGIN
        first, k = decomp.split_layers(int(math.ceil(nx * dx / cell)), world)[rank]
        lo_x, hi_x = first * cell, (first + k) * cell
        p0 = max(0, int(math.floor(lo_x / dx)) - 1)
        p1 = min(nx, int(math.ceil(hi_x / dx)) + 1)
        b = synth.coupled_block_3d(p1 - p0, ny, nz, dx=dx, ix0=p0, nx_total=nx)
        keep = decomp.owner_mask(b.arrays["x"], lo_x, hi_x, rank == 0, rank == world - 1)
        b.arrays = {k_: np.ascontiguousarray(v[keep]) for k_, v in b.arrays.items()}
        b.meta["ids"] = b.meta["ids"][keep]
        return b, (lo_x, hi_x)
    if workload == "wcsph3d_80m":          # configs[3] on ONE GPU (strong-scaling reference point): 800 x 400 x 250
        return synth.wcsph_block_3d(800, 400, 250, name="wcsph3d_80m"), None
    if workload == "dem3d_1m":
        return synth.dem_column_3d(100), None
    if workload == "dem3d_8m":
        return synth.dem_column_3d(200), None
    if workload == "wcsph2d_20k":
        return synth.wcsph_dambreak_2d(dx=0.01), None
    if workload.startswith("wcsph3d_"):     # e.g. wcsph3d_1m: cubes for quick runs
        n = {"1m": (100, 100, 100), "2m": (100, 100, 200), "500k": (100, 100, 50)}[workload.split("_")[1]]
        return synth.wcsph_block_3d(*n, name=workload), None
    raise SystemExit(f"unknown workload {workload}")


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, exact pid, killed after the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's timed step on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_sample_block(workload: str):
    from prestige_b200 import synth
    if workload.startswith("wcsph3d"):
        return synth.wcsph_block_3d(100, 100, 100), "first 100x100x100 lattice planes (1.0 M particles) of the same generator"
    if workload.startswith("dem3d"):
        return synth.dem_column_3d(64), "64^3 spheres + floor (0.27 M particles) of the same generator"
    if workload.startswith("coupled3d"):
        return synth.coupled_block_3d(100, 100, 96), "100x100x96 lattice + floor (0.99 M particles, 10 % spheres) of the same generator"
    return synth.wcsph_dambreak_2d(dx=0.01), "the full 2D dam break (23 k particles)"


def cpu_step_time(block, reps: int):
    from oracle import oracle as orc
    g = orc.make_grid(block.dim, block.lo, block.hi, block.cell_size)
    best = float("inf")
    hist = None
    for _ in range(reps):
        t0 = time.perf_counter()
        if block.physics == "wcsph":
            orc.wcsph(block.dim, block.params, block.arrays, grid=g, sorted_step=True)
        elif block.physics == "dem":
            _, hist, _ = orc.dem(block.params, block.max_contacts, block.arrays, hist=hist, grid=g)
        else:
            _, hist, _ = orc.coupled(block.params, block.max_contacts, block.arrays, hist=hist, grid=g)
        best = min(best, time.perf_counter() - t0)
    return best, orc.num_threads()


def cpu_baseline(workload: str, budget_s: float = 20.0):
    block, desc = cpu_sample_block(workload)
    t1, cores = cpu_step_time(block, 1)
    reps = int(max(1, min(5, budget_s / max(t1, 1e-3) - 1)))
    t, cores = cpu_step_time(block, reps) if reps > 1 else (t1, cores)
    t = min(t, t1)
    out = {"value": block.n / t, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"{desc}; oracle cell-list step (key+sort+permute+EOS+pair loop), best of {reps + 1}, {t * 1e3:.1f} ms/step"}
    if workload == "wcsph2d_20k":
        # for configs[0] also the loop the reference's back-end actually emits (simple_cpu.rs:7-8): all pairs,
        # one thread -- "reference loop semantics as written"
        from oracle import oracle as orc
        orc.set_num_threads(1)
        try:
            t0 = time.perf_counter()
            orc.wcsph(block.dim, block.params, block.arrays)
            ta = time.perf_counter() - t0
        finally:
            orc.set_num_threads(cores)
        out["allpairs_literal"] = {"value": block.n / ta, "unit": UNIT, "cores": 1,
                                   "sample": f"same block, literal for i {{ for j {{ bodies }} }} loop over all {block.n}^2 pairs, one evaluation, {ta * 1e3:.0f} ms"}
    return out


def run_reference(args, rank: int):
    if rank != 0:
        return
    block, desc = cpu_sample_block(args.workload)
    from oracle import oracle as orc
    g = orc.make_grid(block.dim, block.lo, block.hi, block.cell_size)
    hist = None

    def one():
        nonlocal hist
        if block.physics == "wcsph":
            orc.wcsph(block.dim, block.params, block.arrays, grid=g, sorted_step=True)
        elif block.physics == "dem":
            _, hist, _ = orc.dem(block.params, block.max_contacts, block.arrays, hist=hist, grid=g)
        else:
            _, hist, _ = orc.coupled(block.params, block.max_contacts, block.arrays, hist=hist, grid=g)
    for _ in range(args.warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t0
    val = block.n * args.steps / dt
    real = args.real
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong" if args.workload.startswith("coupled3d") else "weak",
            "vs_baseline": None, "dtype": real, "data": "synthetic",
            "config": {"workload": args.workload, "sample": desc, "note": "CPU restatement (oracle/), not reference code: the reference has no runnable path (SURVEY.md 0.1)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": orc.num_threads(), "kind": "port", "sample": desc},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Pinned:
    """Pinned host buffers from pst_host_alloc, viewed as numpy arrays."""

    def __init__(self, lib):
        self.lib, self.ptrs = lib, []

    def array(self, n: int, dtype) -> np.ndarray:
        dt = np.dtype(dtype)
        p = self.lib.pst_host_alloc(max(1, n) * dt.itemsize)
        if not p:
            raise MemoryError("pst_host_alloc failed")
        self.ptrs.append(p)
        return np.frombuffer((C.c_char * (n * dt.itemsize)).from_address(p), dtype=dt, count=n)

    def free(self):
        for p in self.ptrs:
            self.lib.pst_host_free(p)
        self.ptrs = []


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="wcsph3d_10m")
    ap.add_argument("--real", default="f64", choices=["f64", "f32"])
    ap.add_argument("--force-kernel", type=int, default=2)
    ap.add_argument("--key", default="linear", choices=["linear", "morton"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="name=int kernel option (pst_set_option)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
        os.environ["NCCL_DEBUG"] = "WARN"          # NCCL's version banner goes to stdout: keep stdout to the one JSON line
    import torch
    import torch.distributed as dist
    import prestige_b200 as pb
    from prestige_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback; use --impl reference for the CPU restatement)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    real = np.float64 if args.real == "f64" else np.float32
    block, slab = make_block(args.workload, rank, world)
    block = block.astype(real)
    n = block.n
    ghost_cap = 0
    lo, hi = list(block.lo), list(block.hi)
    if world > 1:
        if slab is None:
            raise SystemExit("multi-GPU is implemented for the wcsph3d_10m (weak) and coupled3d_* (strong) slab workloads")
        lo[0], hi[0] = slab
        ny_p, nz_p = block.meta["lattice"][1], block.meta["lattice"][2]
        ghost_cap = int(1.1 * ny_p * (nz_p + 3) * 3)    # a cell layer holds at most 3 lattice planes; also the halo window size
    cap = int(n * 1.02) + 1024
    ctx = pb.Context(dim=block.dim, lo=lo, hi=hi, cell_size=block.cell_size, capacity=cap, real=real, physics=block.physics,
                     key=args.key, max_contacts=block.max_contacts, device=local_rank, ghost_capacity=ghost_cap)
    ctx.set_option("force_kernel", args.force_kernel)
    for o in args.opt:
        k, v = o.split("=")
        ctx.set_option(k, int(v))
    if world > 1:
        uid = [pb.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
    ctx.load_block(block)
    if world > 1 and "m" in block.arrays:
        # the single-rank run skips the m[j] gather because every uploaded mass is equal; across ranks only the caller can know
        # that: tell the library when min == max over ALL ranks' particles, so every N times the same kernel
        from prestige_b200 import decomp
        if decomp.uniform_across_ranks(block.arrays["m"], dist, device="cuda"):
            ctx.set_option("uniform_mass_global", 1)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    coupled = block.physics == "wcsph+dem"
    pair_eqs = ["dem_contact"] if block.physics == "dem" else ["continuity", "momentum"]

    def step(ev=None):
        if ev: ev[0].record(stream)
        ctx.build_neighbours()
        if world > 1:
            ctx.halo_exchange()
        if ev: ev[1].record(stream)
        if block.physics != "dem":
            ctx.apply(["tait_eos"])
        if ev: ev[2].record(stream)
        ctx.apply(pair_eqs)
        if ev: ev[3].record(stream)
        if coupled:
            ctx.apply(["dem_contact"])
        if ev: ev[4].record(stream)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()
        ctx.sync()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    l0 = ctx.stat("launches")
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        step(evs[k])
    barrier()
    wall = time.perf_counter() - t0
    launches = ctx.stat("launches") - l0
    clocks = sampler.stop() if rank == 0 else None
    t_dev = evs[0][0].elapsed_time(evs[-1][4]) * 1e-3         # device time of the whole timed region
    t_nnps = sum(e[0].elapsed_time(e[1]) for e in evs) * 1e-3 / args.steps
    t_eos = sum(e[1].elapsed_time(e[2]) for e in evs) * 1e-3 / args.steps
    t_force = sum(e[2].elapsed_time(e[3]) for e in evs) * 1e-3 / args.steps
    t_dem = sum(e[3].elapsed_time(e[4]) for e in evs) * 1e-3 / args.steps
    zbar, n_solid = 0.0, 0
    if block.physics == "dem":
        zbar = ctx.stat("contacts_total") / n
    elif coupled:
        n_solid = int((block.arrays["tag"] == 2).sum())
        zbar = ctx.stat("contacts_total") / max(n_solid, 1)     # mean stored contacts per SPHERE
    n_total, t_max = n, t_dev
    if world > 1:
        tt = torch.tensor([t_dev, wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_max = float(tt[0])
        nn = torch.tensor([n, n_solid, int(round(zbar * n_solid))], device="cuda", dtype=torch.int64)
        dist.all_reduce(nn)
        n_total, n_solid_total = int(nn[0]), int(nn[1])
        zbar_total = float(nn[2]) / max(n_solid_total, 1)
    else:
        n_solid_total, zbar_total = n_solid, zbar
    value = n_total * args.steps / t_max

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        lib = _lib.load()
        pin = Pinned(lib)
        ins = [k for k in block.arrays if ctx.has_array(k) and k not in ("tag", "m", "h", "rad", "inertia")]
        outs = (["au", "av", "aw", "arho"] if block.dim == 3 else ["au", "av", "arho"]) if block.physics != "dem" else []
        outs += ["fx", "fy", "fz", "tx", "ty", "tz"] if block.physics != "wcsph" else []
        hin = {k: pin.array(n, block.arrays[k].dtype) for k in ins}
        for k in ins:
            hin[k][:] = block.arrays[k]
        hout = {k: pin.array(n, real) for k in outs}
        h2d = sum(v.nbytes for v in hin.values()); d2h = sum(v.nbytes for v in hout.values())

        def e2e_step():
            # the public asynchronous path: H2D of this step's state and D2H of the previous step's rates ride the two
            # copy engines (PCIe is full duplex) while the compute stream works; every byte still crosses every step
            if world > 1:
                ctx.set_count(n)                 # distributed mode: host arrays are in device order, so restart from the host order
            for k in ins:
                ctx.upload_async(k, hin[k].ctypes.data)
            step()
            ctx.wait_transfers()                 # previous step's rates have landed: the host may consume them now
            for k in outs:
                ctx.download_async(k, hout[k].ctypes.data)
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(3):                       # warm-up: staging ring allocated, pinned pages touched
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()                                # pst_sync: all uploads consumed, all rates of the last step on the host
        te = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([te], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            te = float(tt[0])
        e2e = {"value": n_total * e2e_steps / te, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": e2e_steps, "ms_per_step": te / e2e_steps * 1e3,
               "path": f"pst_upload_async({len(ins)} state arrays, pinned host) -> pst_build_neighbours -> pst_apply -> pst_download_async({len(outs)} rate arrays, pinned host), pst_sync at the end"}
        pin.free()

    if rank == 0:
        peak, peak_src = measured_peaks()
        key = (block.physics, block.dim, args.real)
        if block.physics == "dem":
            b_step, b_force = dem_bytes(args.real, zbar)
        elif coupled:
            # the SPH pass of a coupled step touches every particle's SPH columns; the contact pass is reported beside it
            b_step, _ = coupled_bytes(args.real, zbar_total, n_solid_total / n_total)
            b_force = B_FORCE[("wcsph", 3, args.real)]
        else:
            b_step, b_force = B_ALG[key], B_FORCE[key]
        n_local = n
        ach = b_force * n_local / t_force / 1e9
        kern = ({1: "k_wcsph_cellwarp", 2: "k_wcsph_tiled"}.get(args.force_kernel, "k_wcsph_gather") if args.key == "linear" else "k_wcsph_gather") if block.physics != "dem" else "k_dem_forces"
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(f"{kern}:{args.workload}:{args.real}")
        roofline = {"bound": "hbm", "kernel": kern, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "peak_source": peak_src, "alg_bytes_per_particle": b_force, "particles_per_launch": n_local,
                    "kernel_ms": t_force * 1e3,
                    "step": {"alg_bytes_per_particle": b_step, "achieved": b_step * n_total / (t_max / args.steps) / 1e9 / world,
                             "frac": b_step * n_total / (t_max / args.steps) / 1e9 / world / peak},
                    "stage_ms": {"nnps(keys+sort+table+permute" + ("+halo)" if world > 1 else ")"): t_nnps * 1e3, "eos": t_eos * 1e3, "pair_kernel": t_force * 1e3}}
        if coupled:
            roofline["stage_ms"]["contact_kernel"] = t_dem * 1e3
        if args.real == "f64" and block.physics != "dem":
            # the f64 pair kernel is bound on-chip, not by HBM: its FP64 fraction beside the HBM one
            tf = FLOP_FORCE[block.dim] * n_local / t_force / 1e12
            roofline["fp64_cobound"] = {"alg_flop_per_particle": FLOP_FORCE[block.dim], "achieved": tf, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                                        "frac": tf / FP64_PEAK_TFLOPS, "peak_source": "measured DFMA rate (profiles/ubench_fp64_peak.txt)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t_max / args.steps * 1e3, "higher_is_better": True, "scaling": "strong" if coupled else "weak", "vs_baseline": None,
                "dtype": args.real, "data": "synthetic",
                "config": {"workload": args.workload, "particles": n_total, "particles_per_gpu": n_local, "dim": block.dim,
                           "physics": block.physics, "key": args.key, "force_kernel": kern,
                           "decomposition": (f"{world} x-slabs of the same block, {HALO_DESC}" if coupled else f"{world} x-slabs of {SLAB_CELLS} cell layers, {HALO_DESC}") if world > 1 else "single GPU",
                           "l2": f"no flush needed: state + outputs = {n_local * (b_step) / 1e6:.0f} MB touched per step >> 126 MB L2",
                           "timed": "keys+sort+cell table+permute" + ("+history remap" if block.physics != "wcsph" else "") + ("+halo exchange" if world > 1 else "") + ("+contact kernel" if block.physics == "dem" else "+EOS+fused pair kernel" + ("+contact kernel" if coupled else "")) + "; integrator excluded",
                           "mean_contacts": zbar_total if block.physics != "wcsph" else None,
                           "spheres": n_solid_total if coupled else None},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "wall_ms_per_step": wall / args.steps * 1e3,
                "roofline": roofline}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload)
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
