"""The CPU oracle under AddressSanitizer + UndefinedBehaviorSanitizer ("-fsanitize=address,undefined on
the CPU oracle"): every entry point on small blocks, in a subprocess with libasan preloaded.  The oracle is the checker
of every parity claim, so an out-of-bounds read in it would be a silent wrong answer."""
import os
import subprocess
import sys
import tempfile
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent("""
    import sys
    sys.path.insert(0, {root!r})
    from oracle import oracle as orc
    orc.build = lambda force=False: {so!r}              # the instrumented build instead of oracle/liboracle.so
    import numpy as np
    from prestige_b200 import synth
    b = synth.wcsph_block_3d(9, 8, 7).shuffled()
    g = orc.make_grid(3, b.lo, b.hi, b.cell_size)
    orc.wcsph(3, b.params, b.arrays); orc.wcsph(3, b.params, b.arrays, grid=g); orc.wcsph(3, b.params, b.arrays, grid=g, sorted_step=True)
    orc.pairs(3, b.arrays["x"], b.arrays["y"], b.arrays["z"], b.arrays["h"]); orc.pairs(3, b.arrays["x"], b.arrays["y"], b.arrays["z"], b.arrays["h"], grid=g)
    orc.eq1_allpairs(np.arange(1.0, 11.0), np.zeros(10))
    d = synth.dem_column_3d(6).shuffled()
    gd = orc.make_grid(3, d.lo, d.hi, d.cell_size)
    f, h, _ = orc.dem(d.params, 12, d.arrays); orc.dem(d.params, 12, d.arrays, hist=h, grid=gd)
    assert orc.dem(d.params, 3, d.arrays)[2] == 1       # slot overflow path
    orc.dem(dict(d.params, dem_model=1, Estar=1e7, Gstar=4e6), 12, d.arrays, grid=gd)
    c = synth.coupled_block_3d(9, 8, 9).shuffled()
    gc = orc.make_grid(3, c.lo, c.hi, c.cell_size)
    r, hc, _ = orc.coupled(c.params, c.max_contacts, c.arrays); orc.coupled(dict(c.params, boundary_model=1), c.max_contacts, c.arrays, hist=hc, grid=gc)
    w = synth.wcsph_dambreak_2d(dx=0.05).shuffled().astype(np.float32)
    orc.wcsph(2, dict(w.params, boundary_model=1), w.arrays, grid=orc.make_grid(2, w.lo, w.hi, w.cell_size))
    print("sanitized oracle ok")
""")


def test_oracle_under_asan_ubsan():
    asan = subprocess.run(["/usr/bin/gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not installed")
    with tempfile.TemporaryDirectory() as tmp:
        so = os.path.join(tmp, "liboracle_asan.so")
        subprocess.run(["/usr/bin/g++", "-O1", "-g", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fsanitize=address,undefined",
                        "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer", "-shared", "-o", so, os.path.join(ROOT, "oracle", "oracle.cpp")],
                       check=True, capture_output=True)
        env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
        r = subprocess.run([sys.executable, "-c", SCRIPT.format(root=ROOT, so=so)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "sanitized oracle ok" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
