"""Run the GPU tests of tests/test_rigid.py without pytest/torch start-up cost (a fresh box pays ~1 min for `import torch`).
Usage on the GPU box: python scripts/gpu_rigid_quick.py  -> gpurun_out/rigid_quick.txt"""
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

import test_rigid as T  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "rigid_quick.txt"), "w")


def say(*a):
    msg = " ".join(str(x) for x in a)
    print(msg, flush=True)
    out.write(msg + "\n")
    out.flush()


cases = [("setup_reduce_f64", lambda: T.test_rigid_setup_reduce_parity_gpu(np.float64)),
         ("step_vs_host_stage", T.test_rigid_step_matches_host_stage_gpu),
         ("conservation_rigidity", T.test_rigid_conservation_and_rigidity_gpu),
         ("checkpoint_resume", lambda: T.test_rigid_checkpoint_resume_bit_exact_gpu(Path(tempfile.mkdtemp()))),
         ("api_errors", T.test_rigid_api_errors_gpu),
         ("setup_reduce_f32", lambda: T.test_rigid_setup_reduce_parity_gpu(np.float32))]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] in sys.argv[1:]]
bad = 0
for name, fn in cases:
    t0 = time.time()
    try:
        fn()
        say(f"PASS {name} ({time.time() - t0:.1f} s)")
    except BaseException:
        bad += 1
        say(f"FAIL {name} ({time.time() - t0:.1f} s)")
        say(traceback.format_exc()[-3000:])
say(f"{len(cases) - bad} of {len(cases)} rigid-body GPU cases passed")
sys.exit(1 if bad else 0)
