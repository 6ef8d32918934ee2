"""CPU tests: host-side mirror of the reference interface, and the C-ABI library's exports."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

import prestige_b200 as pb
from prestige_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- the reference's two tests (prestige/src/lib.rs:20-27, :32-52), with the assertions it lacks ----------
def test_equation_ir():
    ir = pb.eq1.ir()
    pb.debug_equation(ir)
    assert ir.name == "eq1"
    assert set(ir.writes) == {"force"}
    assert set(ir.reads) == {"force", "mass"}          # written arrays are also reads (lib.rs:84,119)
    assert ir.body.replace(" ", "") == "{force[i]+=mass[j];}"


def test_fusion():
    eqs = [pb.eq1.ir()]
    fused = pb.fuse(eqs)
    assert set(fused.reads) == {"force", "mass"} and set(fused.writes) == {"force"}
    assert len(fused.bodies) == 1 and fused.names == ["eq1"]
    code = pb.codegen.generate_simple_cpu(fused)
    lines = code.split("\n")
    assert lines[0] == "for i in 0..n {" and lines[1] == "    for j in 0..n {"       # simple_cpu.rs:7-8
    assert lines[2].startswith("        {") and lines[3] == "    }" and lines[4] == "}" and lines[5] == ""


def test_macro_semantics():
    @pb.equation
    def eq2(i, j, a, b, c, self_like):
        a[i] = b[j] * 2.0
        c[i] -= a[i] + b[i]
        self_like.x[i] = 1.0        # not a bare path: not detected (lib.rs:73-75)

    ir = eq2.ir()
    assert set(ir.writes) == {"a", "c"}
    assert set(ir.reads) == {"a", "b", "c"}
    with pytest.raises(TypeError):
        eq2(0, 0)                   # the macro drops the function (lib.rs:36-55)
    f = pb.fuse([pb.eq1.ir(), ir])
    assert f.bodies[0].startswith("{ force") and f.names == ["eq1", "eq2"]      # bodies keep input order (fuse.rs:18,30)
    assert set(f.reads) == {"force", "mass", "a", "b", "c"}


def test_b200_backend_plan():
    f = pb.fuse([pb.tait_eos.ir(), pb.continuity.ir(), pb.momentum.ir()])
    assert pb.codegen.b200.generate_b200(f) == ["tait_eos", "continuity", "momentum"]
    w = pb.fuse([pb.tait_eos.ir(), pb.wall_pressure.ir(), pb.continuity.ir(), pb.momentum.ir()])
    assert pb.codegen.b200.generate_b200(w)[1] == "wall_pressure" and {"p", "rho"} <= set(w.writes)
    # every name the back-ends list has a kernel behind pst_apply (api.cu) and vice versa
    api = open(os.path.join(ROOT, "prestige_b200", "csrc", "api.cu")).read()
    known = re.search(r"known = \{([^}]*)\}", api).group(1)
    assert sorted(re.findall(r'"(\w+)"', known)) == sorted(pb.codegen.b200.KERNELS)
    hpp = open(os.path.join(ROOT, "include", "prestige.hpp")).read()
    assert sorted(re.findall(r'"(\w+)"', re.search(r"kernels = \{([^}]*)\}", hpp).group(1))) == sorted(pb.codegen.b200.KERNELS)
    rs = open(os.path.join(ROOT, "rust", "prestige", "src", "codegen", "b200.rs")).read()
    assert sorted(re.findall(r'"(\w+)"', re.search(r"KERNELS: \[&str; \d+\] = \[([^\]]*)\]", rs).group(1))) == sorted(pb.codegen.b200.KERNELS)
    with pytest.raises(ValueError):
        @pb.equation
        def unknown(i, j, q):
            q[i] += q[j]
        pb.codegen.b200.generate_b200(pb.fuse([unknown.ir()]))


# ---- the C ABI ---------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "prestige_b200.h")).read()
    declared = re.findall(r"PST_API\s+[\w\s\*]+?\b(pst_\w+)\s*\(", hdr)
    assert len(declared) >= 25
    assert sorted(declared) == sorted(_lib.SYMBOLS)
    # the Rust FFI crate (source only, never compiled here) declares exactly the same entry points
    rs = open(os.path.join(ROOT, "rust", "prestige_b200_sys", "src", "lib.rs")).read()
    assert sorted(re.findall(r"pub fn (pst_\w+)", rs)) == sorted(declared)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), f"{s} not exported"
    assert b"sm_100a" in _lib.load().pst_version()


def test_config_struct_layout():
    assert ctypes.sizeof(_lib.PstConfig) == 4 * 8 + 8 * 2 + 8 * 7


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pb.PstError) as e:
        pb.Context(dim=3, lo=(0, 0, 0), hi=(1, 1, 1), cell_size=0.1, capacity=10)
    assert e.value.status == _lib.PST_ECUDA and "no CPU path" in str(e.value)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under prestige_b200/ may reference it."""
    for dp, _, fs in os.walk(os.path.join(ROOT, "prestige_b200")):
        if "build" in dp:
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


# ---- synthetic generators ----------------------------------------------------------------------
def test_synth_is_counter_based():
    a = synth.wcsph_block_3d(8, 6, 5)
    whole = synth.wcsph_block_3d(16, 6, 5)
    slab = synth.wcsph_block_3d(8, 6, 5, ix0=8, nx_total=16)
    assert np.array_equal(whole.arrays["x"][8 * 30:], slab.arrays["x"])
    assert np.array_equal(whole.arrays["rho"][:8 * 30], a.arrays["rho"])
    assert np.array_equal(whole.meta["ids"][8 * 30:], slab.meta["ids"])
    u = synth.uniform01(np.arange(100000), 3)
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 5e-3


def test_synth_configs():
    c1 = synth.wcsph_dambreak_2d(dx=0.01)
    assert c1.meta["n_fluid"] == 20000 and c1.dim == 2
    d = synth.dem_column_3d(10)
    assert d.n == 1000 + 100 and d.max_contacts == 12
    s = d.shuffled()
    assert sorted(s.arrays["x"]) == sorted(d.arrays["x"])


# ---- bench.py plumbing that needs no GPU -------------------------------------------------------
def test_bench_traffic_lookup_matches_only_the_same_instantiation():
    """roofline.traffic / ncu_units come from profiles/traffic.json and are attached only to the kernel instantiation that was
    profiled: same mangled symbol (the anonymous-namespace tag, which hashes the build path, aside), same workload, same real."""
    sys.path.insert(0, ROOT)
    import bench
    import json
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    e = next(x for x in t["entries"] if "k_wcsph_zrun" in x["kernel_mangled"] and x["workload"] == "wcsph3d_10m")
    sym = e["kernel_mangled"]
    got, src, units = bench.traffic_for(sym, "wcsph3d_10m", "f64")
    assert got == e["dram_bytes_per_launch"] and units["l1_data_pipe_pct_of_peak"] > 50
    moved = re.sub(r"_GLOBAL__N__[0-9a-f]+_", "_GLOBAL__N__deadbeef_", sym)          # the same kernel built under another path
    assert moved != sym and bench.traffic_for(moved, "wcsph3d_10m", "f64")[0] == got
    other = sym.replace("ELb1ELb1ELb0ELb1ELb1E", "ELb1ELb1ELb0ELb0ELb1E")              # another instantiation (UNI off)
    assert other != sym and bench.traffic_for(other, "wcsph3d_10m", "f64")[0] is None
    assert bench.traffic_for(sym, "wcsph3d_10m", "f32")[0] is None
    assert bench.traffic_for(sym, "dem3d_1m", "f64")[0] is None


def test_bench_config_is_the_same_object_in_both_arms():
    """The CPU arm (--impl reference) and the GPU arm describe the workload with the same `config` keys and values (the GPU arm's
    run-specific details live under `run`), so the driver compares like with like."""
    sys.path.insert(0, ROOT)
    import bench
    b = synth.wcsph_block_3d(6, 6, 6)
    c1 = bench.workload_config("wcsph3d_80m", b, 1, "f64", "linear")
    c8 = bench.workload_config("wcsph3d_80m", b, 8, "f64", "linear", n_total=b.n)
    assert c1 == c8 and set(c1) == {"workload", "particles", "dim", "physics", "real", "key", "geometry", "timed", "l2"}
    assert "BASELINE configs[3]" in c1["geometry"]
