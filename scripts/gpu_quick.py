"""Torch-free regression runner for a GPU box with little time: smoke() plus a curated list of the GPU parity tests,
called directly (no pytest collection, no `import torch`: a fresh box pays ~1 min for that).  The full suite stays
`python -m pytest tests -m gpu`.  Output: gpurun_out/quick.txt"""
import os
import sys
import tempfile
import time
import traceback
from pathlib import Path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import __graft_entry__ as G  # noqa: E402
import test_coupled as TC  # noqa: E402
import test_gpu_parity as TP  # noqa: E402
import test_rigid as TR  # noqa: E402
import test_wall_pressure as TW  # noqa: E402

f64, f32 = np.float64, np.float32
tmp = lambda: Path(tempfile.mkdtemp())
CASES = [
    ("smoke", G.smoke),
    ("mass_paths_f64_3d", lambda: TP.test_wcsph_nonuniform_and_uniform_mass_paths(f64, 3)),
    ("mass_paths_f32_3d", lambda: TP.test_wcsph_nonuniform_and_uniform_mass_paths(f32, 3)),
    ("mass_paths_f64_2d", lambda: TP.test_wcsph_nonuniform_and_uniform_mass_paths(f64, 2)),
    ("wcsph_3d_f64_v2", lambda: TP.test_wcsph_3d(f64, 2)),
    ("wcsph_2d_f64_v2", lambda: TP.test_wcsph_2d_dambreak(f64, 2)),
    ("tiled_edge_shapes", lambda: [TP.test_wcsph_tiled_edge_shapes(o, 2) for o in ({"tile_g": 1}, {"tile_g": 3, "tile_lcap": 8}, {"tile_g": 4, "tile_jcap": 100, "tile_lcap": 16}, {"tile_g": 64})]),
    ("separate_eq_equal_fused", TP.test_wcsph_separate_equations_equal_fused),
    ("wcsph_step_vs_host", TP.test_wcsph_step_matches_host_integration),
    ("dem_history_f64_linear", lambda: TP.test_dem_linear_history(f64, "linear")),
    ("dem_history_f64_morton", lambda: TP.test_dem_linear_history(f64, "morton")),
    ("dem_antisymmetry", TP.test_dem_antisymmetry_bit_exact),
    ("dem_overflow", TP.test_dem_overflow_is_reported),
    ("checkpoint_resume", lambda: TP.test_checkpoint_resume_is_bit_exact(tmp())),
    ("golden_vectors", TP.test_golden_vectors_on_gpu),
    ("errors_are_loud", TP.test_errors_are_loud),
    ("async_round_trip", TP.test_async_transfers_round_trip),
    ("counting_sort_dem", lambda: TP.test_counting_sort_equals_stable_radix_sort("dem")),
    ("coupled_f64_v2", lambda: TC.test_coupled_gpu_one_evaluation(f64, 2)),
    ("coupled_f64_v0", lambda: TC.test_coupled_gpu_one_evaluation(f64, 0)),
    ("coupled_f32_v2", lambda: TC.test_coupled_gpu_one_evaluation(f32, 2)),
    ("coupled_golden_edge", TC.test_coupled_gpu_golden_and_edge_tiles),
    ("coupled_step_vs_host", TC.test_coupled_step_matches_host_integration),
    ("rigid_setup_reduce", lambda: TR.test_rigid_setup_reduce_parity_gpu(f64)),
    ("rigid_checkpoint", lambda: TR.test_rigid_checkpoint_resume_bit_exact_gpu(tmp())),
    ("wall_wcsph_f64_linear", lambda: TW.test_wall_pressure_gpu_wcsph(f64, "linear")),
    ("wall_wcsph_f32_morton", lambda: TW.test_wall_pressure_gpu_wcsph(f32, "morton")),
    ("wall_coupled_f64", lambda: TW.test_wall_pressure_gpu_coupled(f64)),
    ("wall_coupled_f32", lambda: TW.test_wall_pressure_gpu_coupled(f32)),
    ("wall_step_errors", TW.test_wall_pressure_step_and_errors),
]
if len(sys.argv) > 1:
    CASES = [c for c in CASES if c[0] in sys.argv[1:]]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "quick.txt"), "w")


def say(msg):
    print(msg, flush=True)
    out.write(msg + "\n"); out.flush()


bad = 0
t_all = time.time()
for name, fn in CASES:
    t0 = time.time()
    try:
        fn()
        say(f"PASS {name} ({time.time() - t0:.2f} s)")
    except BaseException:
        bad += 1
        say(f"FAIL {name} ({time.time() - t0:.2f} s)\n" + traceback.format_exc()[-2500:])
say(f"{len(CASES) - bad} of {len(CASES)} cases passed in {time.time() - t_all:.1f} s")
sys.exit(1 if bad else 0)
