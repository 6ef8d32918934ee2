#!/usr/bin/env python
"""bench.py -- particle-steps/sec of the NNPS + pair-force hot path on B200 (BASELINE.json metric).

One "step" = cell keys -> counting sort -> cell table -> permute state (-> history remap) (-> halo exchange) -> EOS ->
fused continuity+momentum pair kernel (or the DEM contact kernel) over one block of synthetic particles.  Integrator
excluded from the headline; `moving` in the JSON line times the same step WITH the integrator, so the sort
works on a permutation that is near the identity but not the identity.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload wcsph3d_80m|wcsph3d_10m|dem3d_1m|wcsph2d_20k|coupled3d_20m] [--real f64|f32]
    python bench.py --impl reference ...      # the CPU restatement (oracle/) on ALL host cores, the SAME workload config
    torchrun --nproc-per-node N bench.py --gpus N ...

Headline workload = BASELINE configs[3], the 3D WCSPH tank of 80 M particles (800 x 400 x 250 lattice): it fits one B200,
and --gpus N cuts the SAME tank into N x-slabs, so the per-N values form the STRONG-scaling curve north_star asks for.
The line also carries `extra_configs` (N = 1 only): short runs of configs[2] (10 M, the size the single-GPU roofline target
is quoted on), configs[1] (DEM 1 M), configs[0] (2D dam break) and a 2 M coupled block, each with its own stage split.

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
DEFAULT_WORKLOAD = "wcsph3d_80m"

# algorithmic (compulsory) bytes per particle-step, SURVEY.md 8d / BASELINE.md 4
B_ALG = {("wcsph", 3, "f64"): 350, ("wcsph", 3, "f32"): 202, ("wcsph", 2, "f64"): 286, ("wcsph", 2, "f32"): 166}
B_FORCE = {("wcsph", 3, "f64"): 117, ("wcsph", 3, "f32"): 61, ("wcsph", 2, "f64"): 93, ("wcsph", 2, "f32"): 49}

# algorithmic f64 flop per particle of the fused pair kernel (SURVEY.md 8d "FLOP figure for the co-bound"): ~8 flop per
# candidate of the 27 (9) cell stencil + ~70 flop per in-range pair, at h = 1.2 dx, cutoff 2h, cell = cutoff:
# 3D 373 candidates / 58 neighbours, 2D 52 / 18.  Peak: scripts/ubench/fp64_peak.cu, profiles/ubench_fp64_peak.txt.
FLOP_FORCE = {3: 8 * 373 + 70 * 58, 2: 8 * 52 + 70 * 18}
FP64_PEAK_TFLOPS = 36.5

TANK = (800, 400, 250)      # configs[3]: lattice planes of the 80 M tank (dx = 0.005)
DX = 0.005


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


def host_threads() -> int:
    """Host threads the CPU arm may use: every core this process is allowed on (torchrun's OMP_NUM_THREADS=1 must not decide)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
# what pst_halo_exchange does by default (csrc/halo.cu: halo_impl option / PST_HALO_IMPL): 2 = peer memory, 1 = packed NCCL, 0 = per-array NCCL
HALO_DESC = {"2": "peer-memory halo (pack kernel stores into the neighbour's cudaIpc buffer over NVLink, epoch flag, wait+unpack kernel)",
             "1": "one packed NCCL send/recv message per neighbour", "0": "per-array NCCL send/recv halo with count hand-shake"}.get(
                 os.environ.get("PST_HALO_IMPL", "2"), "peer-memory halo")
SLAB_CELLS = 83   # weak-scaling family: every rank owns 83 cell layers (0.996 m, ~199.2 lattice planes, ~9.96 M particles)


def _slab_of(gen, nx, ny, nz, rank, world, **kw):
    """x-slab `rank` of `world` of the lattice block gen(nx, ny, nz): whole cell layers, near-equal split; the slab is generated
    directly (counter-based generator), never the whole block."""
    from prestige_b200 import decomp, synth
    cell = 2.0 * 1.2 * DX * synth.CELL_MARGIN
    first, k = decomp.split_layers(int(math.ceil(nx * DX / cell)), world)[rank]
    lo_x, hi_x = first * cell, (first + k) * cell
    p0 = max(0, int(math.floor(lo_x / DX)) - 1)
    p1 = min(nx, int(math.ceil(hi_x / DX)) + 1)
    b = gen(p1 - p0, ny, nz, dx=DX, ix0=p0, nx_total=nx, **kw)
    keep = decomp.owner_mask(b.arrays["x"], lo_x, hi_x, rank == 0, rank == world - 1)
    b.arrays = {k_: np.ascontiguousarray(v[keep]) for k_, v in b.arrays.items()}
    b.meta["ids"] = b.meta["ids"][keep]
    return b, (lo_x, hi_x)


def make_block(workload: str, rank: int = 0, world: int = 1):
    """-> (block, slab or None, scaling).  world > 1: the x-slab of this rank."""
    from prestige_b200 import synth
    if workload == "wcsph3d_80m":              # configs[3]; N > 1 cuts the SAME tank into x-slabs: strong scaling
        if world == 1:
            return synth.wcsph_block_3d(*TANK, name="wcsph3d_80m"), None, "strong"
        b, slab = _slab_of(synth.wcsph_block_3d, *TANK, rank, world, name="wcsph3d_80m_slab")
        return b, slab, "strong"
    if workload == "wcsph3d_10m":              # configs[2]; N > 1: WEAK scaling, one ~10 M slab of 83 cell layers per rank
        if world == 1:
            return synth.wcsph_block_3d(200, 200, 250, name="wcsph3d_10m"), None, "weak"
        cell = 2.0 * 1.2 * DX * synth.CELL_MARGIN
        lo_x, hi_x = rank * SLAB_CELLS * cell, (rank + 1) * SLAB_CELLS * cell
        nx_total = int(math.floor(world * SLAB_CELLS * cell / DX))
        p0 = max(0, int(math.floor(lo_x / DX)) - 1)
        p1 = min(nx_total, int(math.ceil(hi_x / DX)) + 1)
        b = synth.wcsph_block_3d(p1 - p0, 200, 250, dx=DX, ix0=p0, nx_total=nx_total, name="wcsph3d_10m_slab")
        keep = (b.arrays["x"] >= lo_x) & (b.arrays["x"] < hi_x)
        b.arrays = {k: np.ascontiguousarray(v[keep]) for k, v in b.arrays.items()}
        b.meta["ids"] = b.meta["ids"][keep]
        return b, (lo_x, hi_x), "weak"
    if workload.startswith("coupled3d_"):      # configs[4]: rigid spheres in fluid; N > 1 cuts the SAME block (strong scaling)
        nx, ny, nz = {"20m": (250, 250, 320), "2m": (125, 125, 128), "300k": (64, 64, 72)}[workload.split("_")[1]]
        if world == 1:
            return synth.coupled_block_3d(nx, ny, nz), None, "strong"
        b, slab = _slab_of(synth.coupled_block_3d, nx, ny, nz, rank, world)
        return b, slab, "strong"
    if world > 1:
        raise SystemExit("multi-GPU is implemented for the slab workloads wcsph3d_80m (strong), wcsph3d_10m (weak) and coupled3d_* (strong)")
    if workload == "dem3d_1m":
        return synth.dem_column_3d(100), None, "strong"
    if workload == "dem3d_8m":
        return synth.dem_column_3d(200), None, "strong"
    if workload == "wcsph2d_20k":
        return synth.wcsph_dambreak_2d(dx=0.01), None, "strong"
    if workload.startswith("wcsph3d_"):        # e.g. wcsph3d_1m: cubes for quick runs
        n = {"1m": (100, 100, 100), "2m": (100, 100, 200), "500k": (100, 100, 50)}[workload.split("_")[1]]
        return synth.wcsph_block_3d(*n, name=workload), None, "strong"
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
# CPU arm: the oracle's cell-list step on the SAME workload config, all host threads
# ------------------------------------------------------------------------------------------------
def oracle_step_fn(block):
    """One CPU step of the same path (key + sort + permute + EOS + pair loop, cell-list mode) on `block`; returns a callable."""
    from oracle import oracle as orc
    orc.set_num_threads(host_threads())
    g = orc.make_grid(block.dim, block.lo, block.hi, block.cell_size)
    state = {"hist": None}

    def one():
        if block.physics == "wcsph":
            orc.wcsph(block.dim, block.params, block.arrays, grid=g, sorted_step=True)
        elif block.physics == "dem":
            _, state["hist"], _ = orc.dem(block.params, block.max_contacts, block.arrays, hist=state["hist"], grid=g)
        else:
            _, state["hist"], _ = orc.coupled(block.params, block.max_contacts, block.arrays, hist=state["hist"], grid=g)
    return one, orc.num_threads()


def cpu_baseline(workload: str, block=None, budget_s: float = 25.0):
    """The CPU restatement timed on rank 0's host cores on the FULL workload config (the whole block, not a slab): one warm-up
    step when the budget allows, then as many timed steps as fit ~budget_s (at least one), best-of."""
    if block is None:
        block, _, _ = make_block(workload)
    one, cores = oracle_step_fn(block)
    t0 = time.perf_counter(); one(); t1 = time.perf_counter() - t0
    times = [t1]
    spent = t1
    while spent + min(times) < budget_s and len(times) < 6:
        t0 = time.perf_counter(); one(); dt = time.perf_counter() - t0
        times.append(dt); spent += dt
    t = min(times)
    out = {"value": block.n / t, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"the full {workload} block ({block.n} particles), same generator and seed; oracle cell-list step "
                     f"(key+sort+permute+EOS+pair loop), best of {len(times)}, {t * 1e3:.1f} ms/step"}
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
    """--impl reference: the reference has no runnable path, so this arm times the CPU restatement (oracle/) of the
    same path on ALL host cores, on the SAME workload config as the GPU arm (the whole block: the CPU does not decompose).  Steps are
    full-size; the run is bounded in time instead (steps are cut, never particles), and the line says how many steps ran."""
    if rank != 0:
        return
    block, _, scaling = make_block(args.workload)
    one, cores = oracle_step_fn(block)
    budget = float(os.environ.get("PST_REF_BUDGET_S", "150"))
    t_start = time.perf_counter()
    t_w = []
    for _ in range(args.warmup):
        t0 = time.perf_counter(); one(); t_w.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > 0.25 * budget:
            break
    est = min(t_w) if t_w else None
    done, t_total = 0, 0.0
    while done < args.steps:
        if done >= 1 and est is not None and (time.perf_counter() - t_start) + est > budget:
            break
        t0 = time.perf_counter(); one(); dt = time.perf_counter() - t0
        t_total += dt; done += 1
        est = dt if est is None else min(est, dt)
    val = block.n * done / t_total
    desc = (f"the full {args.workload} block ({block.n} particles), same generator and seed; {done} of the requested {args.steps} steps "
            f"(time budget {budget:.0f} s), {len(t_w)} warm-up")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
            "warmup": len(t_w), "ms_per_step": t_total / done * 1e3, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": args.real, "data": "synthetic",
            "config": workload_config(args.workload, block, 1, args.real, args.key),
            "note": "CPU restatement (oracle/), not reference code: the reference has no runnable path (SURVEY.md 0.1)",
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(workload, block, world, real, key, n_total=None):
    """The config keys both arms print (so the driver's same-config check compares like with like)."""
    n_total = block.n if n_total is None else n_total
    return {"workload": workload, "particles": int(n_total), "dim": block.dim, "physics": block.physics, "real": real, "key": key,
            "geometry": {"wcsph3d_80m": "800 x 400 x 250 jittered lattice, dx = 0.005 (BASELINE configs[3])",
                         "wcsph3d_10m": "200 x 200 x 250 jittered lattice, dx = 0.005 (BASELINE configs[2])",
                         "dem3d_1m": "100^3 spheres + floor (BASELINE configs[1])", "wcsph2d_20k": "2D dam break, dx = 0.01 (BASELINE configs[0])",
                         "coupled3d_20m": "250 x 250 x 320 lattice + floor, 10 % spheres (BASELINE configs[4])"}.get(workload, workload),
            "timed": "keys+sort+cell table+permute" + ("+history remap" if block.physics != "wcsph" else "") +
                     ("+contact kernel" if block.physics == "dem" else "+EOS+fused pair kernel" + ("+contact kernel" if block.physics == "wcsph+dem" else "")) +
                     " (+migration+halo exchange between slabs when decomposed); integrator excluded",
            "l2": l2_note(block, n_total, real)}


def l2_note(block, n_total, real):
    if block.physics == "dem":
        b_step = dem_bytes(real, 6.0)[0]
    elif block.physics == "wcsph+dem":
        b_step = coupled_bytes(real, 2.0, 0.1)[0]
    else:
        b_step = B_ALG[(block.physics, block.dim, real)]
    mb = n_total * b_step / 1e6
    if mb > 4 * 126:
        return f"no flush needed: ~{mb:.0f} MB of state + outputs touched per step >> 126 MB L2"
    return f"inputs (~{mb:.0f} MB per step) fit the 126 MB L2: launch-bound configuration, reported as measured"


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


def traffic_for(kernel_mangled: str, workload: str, real: str):
    """dram bytes per launch of the dominant kernel from profiles/traffic.json (written by scripts/ncu_traffic.py from an
    `ncu --set full` capture).  Keyed by the MANGLED kernel symbol: an entry for another instantiation never matches."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tp):
        return None, "no profiles/traffic.json", None
    with open(tp) as f:
        t = json.load(f)
    import re

    def norm(sym):      # the anonymous-namespace tag hashes the build path: everything else (name, every template argument) must match
        return re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+", "_ANON_", sym or "")
    for e in t.get("entries", []):
        if norm(e.get("kernel_mangled")) == norm(kernel_mangled) and e.get("workload") == workload and e.get("real") == real:
            return e.get("dram_bytes_per_launch"), e.get("source"), {k: e.get(k) for k in ("l1_data_pipe_pct_of_peak", "fp64_pipe_pct_of_peak", "dram_pct_of_peak",
                                                                                             "l1_hit_rate_pct", "l2_hit_rate_pct", "kernel_ms_under_ncu")}
    return None, "no ncu capture of this kernel instantiation on this workload (profiles/traffic.json holds others)", None


def parity_check(ctx, block, slab, rank, world, dist, torch):
    """In-line correctness check of the timed configuration (WCSPH slabs / blocks): this rank's rates against the CPU oracle on a
    band two cells either side of each of its slab faces (single GPU: of the mid plane), matched by global id; plus
    sum m (a - g) = 0 (pairwise antisymmetry) and particle-count conservation over all ranks."""
    from oracle import oracle as orc
    from prestige_b200 import synth
    if block.physics != "wcsph" or block.dim != 3 or "lattice" not in block.meta:
        return None
    orc.set_num_threads(max(1, host_threads() // world))
    cell = block.cell_size
    ny, nz = block.meta["lattice"][1], block.meta["lattice"][2]
    nx_total = block.meta.get("nx_total", block.meta["lattice"][0])
    dx = block.meta.get("dx", DX)
    faces = []
    if slab is None:
        faces.append(0.5 * (block.lo[0] + block.hi[0]))
    else:
        if rank > 0: faces.append(slab[0])
        if rank < world - 1: faces.append(slab[1])
    gid = ctx.download("id").astype(np.int64) if world > 1 else None
    x = ctx.download("x")
    got = {c: ctx.download(c) for c in ("au", "av", "aw", "arho")}
    if world == 1:
        gid = block.meta["ids"].astype(np.int64)       # host arrays are in id order = generator order
    worst, checked = 0.0, 0
    for xf in faces:
        p0 = max(0, int(math.floor((xf - 2.02 * cell) / dx)))
        p1 = min(nx_total, int(math.ceil((xf + 2.02 * cell) / dx)) + 1)
        sub = synth.wcsph_block_3d(p1 - p0, ny, nz, dx=dx, ix0=p0, nx_total=nx_total)
        ref = orc.wcsph(3, sub.params, sub.arrays, grid=orc.make_grid(3, (p0 * dx - 1e-9, 0.0, 0.0), (p1 * dx + 1e-9, sub.hi[1], sub.hi[2]), cell))
        sid = sub.meta["ids"].astype(np.int64)
        inner = np.abs(sub.arrays["x"] - xf) < 0.98 * cell           # complete neighbourhoods inside the sub-block
        order = np.argsort(sid[inner]); sid_in = sid[inner][order]
        mine = np.abs(x - xf) < 0.98 * cell
        pos = np.searchsorted(sid_in, gid[mine])
        ok = (pos < len(sid_in)) & (sid_in[np.minimum(pos, len(sid_in) - 1)] == gid[mine])
        if not ok.all():
            return {"ok": False, "error": "a particle of the band is missing from the oracle's sub-block"}
        for c in got:
            r = ref[c][inner][order][pos]
            scale = float(np.sqrt(np.mean(ref[c] ** 2)))
            worst = max(worst, float(np.max(np.abs(got[c][mine] - r) / np.maximum(np.abs(r), scale))))
        checked += int(mine.sum())
    m = ctx.download("m")
    g = [block.params.get("gx", 0.0), block.params.get("gy", 0.0), block.params.get("gz", 0.0)]
    mom = np.array([np.sum(m * (got["au"] - g[0])), np.sum(m * (got["av"] - g[1])), np.sum(m * (got["aw"] - g[2])),
                    np.sum(m * np.abs(got["au"])), float(ctx.n), float(checked), worst], dtype=np.float64)
    if world > 1:
        t = torch.tensor(mom, device="cuda")
        tmax = t.clone()
        dist.all_reduce(t)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        mom = t.cpu().numpy(); worst = float(tmax[6])
    n_expected = nx_total * ny * nz
    res = {"ok": bool(worst < 1e-10 and int(mom[4]) == n_expected and abs(mom[0]) + abs(mom[1]) + abs(mom[2]) < 1e-9 * mom[3]),
           "max_rel_err_vs_oracle": worst, "tolerance": 1e-10, "particles_checked": int(mom[5]),
           "bands": "two-cell bands around every slab face" if slab is not None else "two-cell band around the mid plane",
           "sum_m_a_over_sum_m_abs_a": float((abs(mom[0]) + abs(mom[1]) + abs(mom[2])) / max(mom[3], 1e-300)),
           "particles_total": int(mom[4]), "particles_expected": int(n_expected)}
    return res


def run_gpu(workload, args, rank, world, local_rank, torch, dist, *, steps, warmup, do_e2e, do_parity, do_moving):
    """One timed GPU run of `workload`; returns the pieces of the JSON line (rank 0) or None (other ranks)."""
    import prestige_b200 as pb
    from prestige_b200 import _lib, decomp
    real = np.float64 if args.real == "f64" else np.float32
    t_gen = time.perf_counter()
    block, slab, scaling = make_block(workload, rank, world)
    block = block.astype(real)
    t_gen = time.perf_counter() - t_gen
    n = block.n
    ghost_cap = 0
    lo, hi = list(block.lo), list(block.hi)
    if world > 1:
        lo[0], hi[0] = slab
        ny_p, nz_p = block.meta["lattice"][1], block.meta["lattice"][2]
        ghost_cap = int(1.1 * ny_p * (nz_p + 3) * 3)    # a cell layer holds at most 3 lattice planes; also the halo window size
    cap = int(n * 1.02) + 1024
    ctx = pb.Context(dim=block.dim, lo=lo, hi=hi, cell_size=block.cell_size, capacity=cap, real=real, physics=block.physics,
                     key=args.key, max_contacts=block.max_contacts, device=local_rank, ghost_capacity=ghost_cap)
    if args.force_kernel is not None:
        ctx.set_option("force_kernel", args.force_kernel)
    for o in args.opt:
        k, v = o.split("=")
        ctx.set_option(k, int(v))
    if world > 1:
        uid = [pb.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
    ctx.load_block(block)
    if world > 1:
        ctx.upload("id", block.meta["ids"].astype(np.uint32))      # distributed mode: global labels, device-order transfers
        if "m" in block.arrays and decomp.uniform_across_ranks(block.arrays["m"], dist, device="cuda") and \
                decomp.uniform_across_ranks(block.arrays["h"], dist, device="cuda"):
            # the single-rank run skips the m[j] gather because every uploaded mass is equal; across ranks only the caller can
            # know that: tell the library when min == max over ALL ranks' particles, so every N times the same kernel
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

    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(steps)]
    l0 = ctx.stat("launches")
    barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        step(evs[k])
    barrier()
    wall = time.perf_counter() - t0
    launches = ctx.stat("launches") - l0
    clocks = sampler.stop() if rank == 0 else None
    t_dev = evs[0][0].elapsed_time(evs[-1][4]) * 1e-3         # device time of the whole timed region
    t_nnps = sum(e[0].elapsed_time(e[1]) for e in evs) * 1e-3 / steps
    t_eos = sum(e[1].elapsed_time(e[2]) for e in evs) * 1e-3 / steps
    t_force = sum(e[2].elapsed_time(e[3]) for e in evs) * 1e-3 / steps
    t_dem = sum(e[3].elapsed_time(e[4]) for e in evs) * 1e-3 / steps
    kern_mangled = ctx.kernel_name("contact" if block.physics == "dem" else "pair", demangle=False)
    kern = ctx.kernel_name("contact" if block.physics == "dem" else "pair")
    zbar, n_solid = 0.0, 0
    if block.physics == "dem":
        zbar = ctx.stat("contacts_total") / n
    elif coupled:
        n_solid = int((block.arrays["tag"] == 2).sum())
        zbar = ctx.stat("contacts_total") / max(n_solid, 1)     # mean stored contacts per SPHERE
    n_total, t_max, t_force_max = n, t_dev, t_force
    if world > 1:
        tt = torch.tensor([t_dev, wall, t_force], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_max, t_force_max = float(tt[0]), float(tt[2])
        nn = torch.tensor([n, n_solid, int(round(zbar * n_solid))], device="cuda", dtype=torch.int64)
        dist.all_reduce(nn)
        n_total, n_solid_total = int(nn[0]), int(nn[1])
        zbar_total = float(nn[2]) / max(n_solid_total, 1)
    else:
        n_solid_total, zbar_total = n_solid, zbar
    value = n_total * steps / t_max

    # ---- in-line correctness of exactly this configuration (every N)
    parity = None
    if do_parity:
        step()
        parity = parity_check(ctx, block, slab, rank, world, dist, torch)

    # ---- the same step WITH the integrator: particles move, the re-sort permutes (near-identity, not identity)
    moving = None
    if do_moving:
        dt_s = 0.1 * (1.2 * DX if block.dim == 3 and block.physics != "dem" else 1e-3) / max(block.params.get("c0", 1.0), 1.0) if block.physics != "dem" else 1e-6
        m_steps = max(4, min(steps, 10)) & ~1 if n > 1000000 else 200
        graph = world == 1
        if graph:
            ctx.set_option("graph", 1)           # pst_step as a CUDA graph (two steps per replay): what the launch-bound configs need
        ctx.step(dt_s, 6)                        # warm-up (and, with graph = 1, the capture)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.step(dt_s, m_steps)
        e1.record(stream)
        barrier()
        tm = e0.elapsed_time(e1) * 1e-3
        if world > 1:
            tt = torch.tensor([tm], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            tm = float(tt[0])
        moving = {"ms_per_step": tm / m_steps * 1e3, "value": n_total * m_steps / tm, "unit": UNIT, "steps": m_steps, "dt": dt_s,
                  "cuda_graph": graph,
                  "path": "pst_step: re-sort (+ migration + halo) -> EOS -> pair kernel -> integrator; the sort sees moved particles"}
        if graph:
            ctx.set_option("graph", 0)
        ctx.load_block(block)                    # back to the synthetic state for the end-to-end run
        if world > 1:
            ctx.upload("id", block.meta["ids"].astype(np.uint32))

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if do_e2e:
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
        step(); barrier()
        n_now = ctx.refresh_count()
        assert n_now == n, "the synthetic block does not move: no particle may have migrated"
        if world > 1:
            # distributed mode: host arrays are in DEVICE order (ids are global labels), so the host copy of the state is the
            # one a coupled code would hold after its last download -- fetch it once, then every step ships it back unchanged
            for k in ins:
                hin[k][:] = ctx.download(k)

        def e2e_step():
            # the public asynchronous path: H2D of this step's state and D2H of the previous step's rates ride the two
            # copy engines (PCIe is full duplex) while the compute stream works; every byte still crosses every step
            for k in ins:
                ctx.upload_async(k, hin[k].ctypes.data)
            step()
            ctx.wait_transfers()                 # previous step's rates have landed: the host may consume them now
            for k in outs:
                ctx.download_async(k, hout[k].ctypes.data)
        e2e_steps = max(3, min(steps, 10))
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
        e2e = {"value": n_total * e2e_steps / te, "unit": UNIT, "h2d_bytes_per_step": h2d * world if world == 1 else int(h2d * n_total / max(n, 1)),
               "d2h_bytes_per_step": d2h * world if world == 1 else int(d2h * n_total / max(n, 1)),
               "steps": e2e_steps, "ms_per_step": te / e2e_steps * 1e3,
               "path": f"pst_upload_async({len(ins)} state arrays, pinned host) -> pst_build_neighbours -> pst_apply -> pst_download_async({len(outs)} rate arrays, pinned host), pst_sync at the end; bytes are the job's total over all ranks"}
        pin.free()
    ctx.close()
    if rank != 0:
        return None

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
    t_kernel = t_force if block.physics != "dem" else t_force
    ach = b_force * n_local / t_kernel / 1e9
    traffic, traffic_src, ncu_units = traffic_for(kern_mangled, workload, args.real)
    roofline = {"bound": "hbm", "kernel": kern, "kernel_mangled": kern_mangled, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                # the unit this kernel actually sits on (recorded ncu capture of the SAME instantiation, not a live measurement)
                "ncu_units": ncu_units,
                "peak_source": peak_src, "alg_bytes_per_particle": b_force, "particles_per_launch": n_local,
                "kernel_ms": t_kernel * 1e3,
                "step": {"alg_bytes_per_particle": b_step, "achieved": b_step * n_total / (t_max / steps) / 1e9 / world,
                         "frac": b_step * n_total / (t_max / steps) / 1e9 / world / peak},
                "stage_ms": {"nnps(keys+sort+table+permute" + ("+migration+halo)" if world > 1 else ")"): t_nnps * 1e3, "eos": t_eos * 1e3, "pair_kernel": t_force * 1e3}}
    if coupled:
        roofline["stage_ms"]["contact_kernel"] = t_dem * 1e3
    if args.real == "f64" and block.physics != "dem":
        # the f64 pair kernel is bound on-chip, not by HBM: its FP64 fraction beside the HBM one
        tf = FLOP_FORCE[block.dim] * n_local / t_force / 1e12
        roofline["fp64_cobound"] = {"alg_flop_per_particle": FLOP_FORCE[block.dim], "achieved": tf, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                                    "frac": tf / FP64_PEAK_TFLOPS, "peak_source": "measured DFMA rate (profiles/ubench_fp64_peak.txt)"}
    cfg = workload_config(workload, block, world, args.real, args.key, n_total)        # identical keys and values in the CPU arm
    import re
    km = re.search(r"(k_\w+)", kern)
    run = {"particles_per_gpu": n_local, "force_kernel": km.group(1) if km else kern,
           "decomposition": f"{world} x-slabs of the SAME block (whole cell layers, near-equal split), {HALO_DESC}" if world > 1 and scaling == "strong" else
                            (f"{world} x-slabs of {SLAB_CELLS} cell layers, {HALO_DESC}" if world > 1 else "single GPU"),
           "per_gpu_mb_per_step": n_local * b_step / 1e6,
           "mean_contacts": zbar_total if block.physics != "wcsph" else None,
           "spheres": n_solid_total if coupled else None}
    return {"value": value, "ms_per_step": t_max / steps * 1e3, "scaling": scaling, "config": cfg, "run": run, "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "wall_ms_per_step": wall / steps * 1e3, "roofline": roofline, "parity_check": parity,
            "moving": moving, "block": block, "host_s": {"generate_block": t_gen}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--real", default="f64", choices=["f64", "f32"])
    ap.add_argument("--force-kernel", type=int, default=None, help="pair-kernel variant (default: the library's, 3)")
    ap.add_argument("--key", default="linear", choices=["linear", "morton"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip extra_configs (short runs of the other BASELINE configs, N = 1)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-moving", action="store_true")
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

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback; use --impl reference for the CPU restatement)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    r = run_gpu(args.workload, args, rank, world, local_rank, torch, dist, steps=args.steps, warmup=args.warmup,
                do_e2e=not args.no_e2e, do_parity=not args.no_parity, do_moving=not args.no_moving)
    extras = {}
    if not args.no_extra and args.workload == DEFAULT_WORKLOAD:
        # N = 1: the other BASELINE configs; every N: configs[4], the coupled SPH-DEM block of 20 M particles, cut into N slabs like the
        # headline (its strong-scaling curve comes out of the same driver runs)
        todo = [("wcsph3d_10m", 20), ("dem3d_1m", 20), ("wcsph2d_20k", 50), ("coupled3d_2m", 10)] if world == 1 else []
        todo.append(("coupled3d_20m", 10))
        for wl, k in todo:
            try:
                x = run_gpu(wl, args, rank, world, local_rank, torch, dist, steps=k, warmup=3, do_e2e=False,
                            do_parity=wl == "wcsph3d_10m", do_moving=True)
                if rank == 0:
                    extras[wl] = {kk: x[kk] for kk in ("value", "ms_per_step", "scaling", "config", "run", "roofline", "gpu_launches", "parity_check", "moving")}
                    extras[wl]["steps"] = k
            except Exception as e:          # an extra must never take the headline down
                if world > 1:
                    raise                   # (with other ranks waiting in a collective there is nothing to salvage: fail loudly)
                extras[wl] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": r["scaling"], "vs_baseline": None,
                "dtype": args.real, "data": "synthetic", "config": r["config"], "run": r["run"], "clocks": r["clocks"], "e2e": r["e2e"],
                "gpu_launches": r["gpu_launches"], "wall_ms_per_step": r["wall_ms_per_step"], "roofline": r["roofline"],
                "parity_check": r["parity_check"], "moving": r["moving"]}
        if extras:
            line["extra_configs"] = extras
        if not args.no_cpu_baseline and world == 1:       # rank 0 at N = 1 only (at N > 1 the other ranks would idle in a barrier)
            line["cpu_baseline"] = cpu_baseline(args.workload, r["block"])
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
