"""CPU check of the device arithmetic: prestige_b200/csrc/wcsph_core.h -- the pair body (continuity + momentum with the
regrouped viscosity term, branch-free dummy pairs, signed SPH mass of coupled contexts) and the wall-pressure sums that
the kernels of wcsph.cu inline -- compiled for the host (tests/cpp/wcsph_core_harness.cpp) and compared with the oracle.
The hardware rsqrt / rcp seeds are the only device-side pieces the host build replaces."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from prestige_b200 import synth
from oracle import oracle as orc
from util import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "cpp", "libwcsph_core_harness.so")


def _lib():
    src = os.path.join(ROOT, "tests", "cpp", "wcsph_core_harness.cpp")
    hdr = os.path.join(ROOT, "prestige_b200", "csrc", "wcsph_core.h")
    if not os.path.exists(HARNESS) or os.path.getmtime(HARNESS) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-o", HARNESS, src],
                       check=True, capture_output=True)
    return C.CDLL(HARNESS)


def _params(P):
    return (C.c_double * 9)(P.get("kfac", 2.0), P["rho0"], P["c0"], P["gamma"], P["alpha"], P["beta"], P.get("gx", 0.0), P.get("gy", 0.0), P.get("gz", 0.0))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _forces(dim, P, a, ms, p):
    real = a["x"].dtype
    sfx = "f64" if real == np.float64 else "f32"
    n = len(a["x"])
    z = a["z"] if dim == 3 else np.zeros(n, real)
    w = a["w"] if dim == 3 else np.zeros(n, real)
    out = {k: np.zeros(n, real) for k in ("au", "av", "aw", "arho")}
    c = lambda v: _ptr(np.ascontiguousarray(v, real))
    getattr(_lib(), f"wch_forces_{sfx}")(C.c_int(dim), _params(P), C.c_int64(n), c(a["x"]), c(a["y"]), c(z), c(a["u"]), c(a["v"]), c(w),
                                        c(a["rho"]), c(ms), c(a["h"]), c(p), _ptr(out["au"]), _ptr(out["av"]), _ptr(out["aw"]), _ptr(out["arho"]))
    return out


@pytest.mark.parametrize("real", [np.float64, np.float32])
def test_pair_body_matches_oracle(real):
    tol = 1e-12 if real == np.float64 else 2e-5
    for b in (synth.wcsph_block_3d(11, 10, 9).shuffled().astype(real), synth.wcsph_dambreak_2d(dx=0.05).shuffled().astype(real)):
        P = dict(b.params, beta=0.3)                      # both viscosity coefficients in play
        ref = orc.wcsph(b.dim, P, b.arrays)
        got = _forces(b.dim, P, b.arrays, b.arrays["m"], ref["p"])
        for k in ("au", "av", "arho") + (("aw",) if b.dim == 3 else ()):
            assert rel_err(got[k], ref[k]) <= tol, (b.name, k, rel_err(got[k], ref[k]))


def test_pair_body_coupled_signed_mass():
    """k_eos' signed SPH mass (+m fluid, -m boundary, -m rho0/rho_solid solid): a pair counts iff i or j is fluid."""
    b = synth.coupled_block_3d(9, 8, 9).shuffled()
    ref, _, _ = orc.coupled(b.params, b.max_contacts, b.arrays)
    tag = b.arrays["tag"]
    ms = np.where(tag == 0, 1.0, -1.0) * orc.sph_mass(b.arrays, b.params)
    got = _forces(3, b.params, b.arrays, ms, ref["p"])
    for k in ("au", "av", "aw", "arho"):
        assert rel_err(got[k], ref[k]) <= 1e-12, (k, rel_err(got[k], ref[k]))


@pytest.mark.parametrize("real", [np.float64, np.float32])
def test_wall_sums_match_oracle(real):
    tol = 1e-12 if real == np.float64 else 2e-5
    blocks = [synth.wcsph_dambreak_2d(dx=0.04).shuffled().astype(real), synth.coupled_block_3d(9, 8, 9).shuffled().astype(real)]
    for b in blocks:
        a = b.arrays
        n = len(a["x"])
        p0 = orc.eos(b.dim, b.params, np.ascontiguousarray(a["rho"]))
        rho_ref, p_ref = orc.wall_pressure(b.dim, b.params, a, p0)
        rho, p = np.array(a["rho"], copy=True), np.array(p0, copy=True)
        z = a["z"] if b.dim == 3 else np.zeros(n, real)
        sfx = "f64" if real == np.float64 else "f32"
        getattr(_lib(), f"wch_wall_{sfx}")(C.c_int(b.dim), _params(b.params), C.c_int64(n), _ptr(np.ascontiguousarray(a["x"])),
                                          _ptr(np.ascontiguousarray(a["y"])), _ptr(np.ascontiguousarray(z)), _ptr(np.ascontiguousarray(a["h"])),
                                          _ptr(np.ascontiguousarray(a["tag"], np.int32)), _ptr(rho), _ptr(p))
        assert rel_err(p, p_ref) <= tol and rel_err(rho, rho_ref) <= tol, (b.name, rel_err(p, p_ref), rel_err(rho, rho_ref))
        fl = a["tag"] == 0
        assert np.array_equal(p[fl], p0[fl]) and np.array_equal(rho[fl], a["rho"][fl])
        assert (p[~fl] != p0[~fl]).any()
