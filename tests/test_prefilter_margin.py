"""Property test of the f32 pre-filter of the fused pair kernel (k_wcsph_tiled, f64 contexts).

The kernel decides neighbourhood with the exact FMA-free f64 test, but only for candidates that survive a cheaper f32
test on tile-local coordinates:  r2f = |float(x_i - o) - float(x_j - o)|^2  <=  round_up(rc2 (1 + 2^-15)).
The neighbour set stays bit-exact only if that filter NEVER rejects a true neighbour.  This emulates the kernel's f32
arithmetic in numpy (plain and fused evaluation order) on millions of pairs placed within a hair of the cutoff, at
coordinates up to the kernel's `far` limit of (G + 6) cells from the tile origin (beyond it the kernel switches the
filter off), for tile depths up to the launcher's cap kMaxTileG, and asserts: no false negative, and a measured worst-case
f32 error below half the margin.  (Without the cap the test fails at G = 64: one f32 ulp at 70 cells from the origin is
worth 1.4e-5 of r2, too close to the 3.05e-5 margin -- which is why the cap exists.)"""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MARGIN = 1.0 / 32768.0
MAX_G = 26
FAR_CELLS = lambda G: G + 6


def test_constants_match_the_kernel():
    src = open(os.path.join(ROOT, "prestige_b200", "csrc", "wcsph.cu")).read()
    assert "(double)I.rc2 * (1.0 + 1.0 / 32768.0)" in src and "__double2float_ru" in src
    assert re.search(r"far_lim = \(float\)\(G \+ 6\) \* \(float\)g\.cell", src)
    assert f"constexpr int kMaxTileG = {MAX_G};" in src and "G = std::min(G, kMaxTileG);" in src


def _round_up_f32(v):
    f = np.float32(v)
    return np.where(f.astype(np.float64) < v, np.nextafter(f, np.float32(np.inf)), f).astype(np.float32)


@pytest.mark.parametrize("G", [1, 4, MAX_G])
@pytest.mark.parametrize("fused", [False, True])
def test_no_false_negative_near_the_cutoff(G, fused):
    rng = np.random.default_rng(1234 + G + fused)
    n = 1_500_000
    dx = 0.005
    h = 1.2 * dx
    rc = 2.0 * h
    rc2 = rc * rc                                                # mul_rn(kfac h, kfac h) in the kernel; same double here
    cell = rc * 1.0001
    far = FAR_CELLS(G) * cell
    origin = rng.uniform(0.0, 4.0, 3)                            # tile origin: some particle's f64 position
    xi = origin + rng.uniform(-far, far, (n, 3)) * 0.999
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    eps = 10.0 ** rng.uniform(-14, -4, n) * rng.choice([-1.0, 1.0], n)      # both sides of the cutoff, down to the last bits
    xj = xi - d * (rc * (1.0 - eps))[:, None]
    inside_far = np.all(np.abs(xj - origin) <= far, axis=1)     # the kernel only filters tiles whose candidates are all within `far`
    xi, xj = xi[inside_far], xj[inside_far]
    # exact test, as the kernel and the oracle evaluate it (left to right, no FMA, f64)
    dd = xi - xj
    r2 = dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1]
    r2 = r2 + dd[:, 2] * dd[:, 2]
    truth = (r2 < rc2) & (r2 > 0)
    # the pre-filter, in f32 on tile-local coordinates
    li = (xi - origin).astype(np.float32)
    lj = (xj - origin).astype(np.float32)
    df = li - lj                                                 # f32 subtraction
    if fused:   # r2f = fma(dz, dz, fma(dy, dy, dx * dx)): each fma rounds once (products of two f32 are exact in f64)
        t = (df[:, 0] * df[:, 0]).astype(np.float32)
        t = (df[:, 1].astype(np.float64) * df[:, 1].astype(np.float64) + t.astype(np.float64)).astype(np.float32)
        r2f = (df[:, 2].astype(np.float64) * df[:, 2].astype(np.float64) + t.astype(np.float64)).astype(np.float32)
    else:
        r2f = df[:, 0] * df[:, 0] + df[:, 1] * df[:, 1]
        r2f = r2f + df[:, 2] * df[:, 2]
    rc2f = _round_up_f32(np.float64(rc2) * (1.0 + MARGIN))
    passed = r2f <= rc2f
    assert truth.sum() > 0.3 * len(truth) and (~truth).sum() > 0.3 * len(truth)
    assert not np.any(truth & ~passed), "the f32 pre-filter rejected a true neighbour: the neighbour set would not be bit-exact"
    worst = float(np.max(np.abs(r2f.astype(np.float64) - r2) / r2))
    assert worst < MARGIN / 2, f"worst-case f32 error {worst:.2e} is too close to the margin {MARGIN:.2e}"
    false_pos = float(np.mean(passed & ~truth & (np.abs(r2 / rc2 - 1.0) > 2 * MARGIN)))
    assert false_pos == 0.0, "pairs farther than twice the margin outside the cutoff must not survive the filter"


@pytest.mark.parametrize("ratio", [20, 333, 5000, 200000])
@pytest.mark.parametrize("fused", [False, True])
def test_grid_relative_coordinates(ratio, fused):
    """Variant 3 with TMA staging (wcsph_zrun.cuh) runs the pre-filter on f32 coordinates relative to the GRID origin, clamped to
    the box + 2 cells, with a margin that grows with the box: max(2^-15, 8 * 2^-23 * E / rc), E = the largest clamped
    coordinate.  Same property: never a false negative, for boxes from 20 to 200 000 cutoffs across."""
    src = open(os.path.join(ROOT, "prestige_b200", "csrc", "wcsph_zrun.cuh")).read()
    assert "8.0 * std::ldexp(1.0, -23) * E / rc" in src and "1.0 / 32768.0" in src
    rng = np.random.default_rng(77 + ratio + fused)
    n = 1_000_000
    rc = 0.012
    rc2 = rc * rc
    E = ratio * rc
    margin = max(1.0 / 32768.0, 8.0 * 2.0 ** -23 * E / rc)
    lo = rng.uniform(-3.0, 3.0, 3)
    xi = lo + rng.uniform(0.0, E, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    eps = 10.0 ** rng.uniform(-14, -4, n) * rng.choice([-1.0, 1.0], n)
    xj = xi - d * (rc * (1.0 - eps))[:, None]
    keep = np.all((xj - lo >= 0) & (xj - lo <= E), axis=1)
    xi, xj = xi[keep], xj[keep]
    dd = xi - xj
    r2 = dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1]
    r2 = r2 + dd[:, 2] * dd[:, 2]
    truth = (r2 < rc2) & (r2 > 0)
    gi = (xi - lo).astype(np.float32)                # pos_f32: (float)(x - lo), within the clamp window here
    gj = (xj - lo).astype(np.float32)
    rc2f = _round_up_f32(rc2 * (1.0 + margin))
    df = (gi - gj).astype(np.float32)
    if fused:       # the kernel: d = fma(dz, dz, fma(dy, dy, fma(dx, dx, -rc2f))), sign bit decides
        acc = (df[:, 0].astype(np.float64) * df[:, 0].astype(np.float64) - np.float64(rc2f)).astype(np.float32)
        acc = (df[:, 1].astype(np.float64) * df[:, 1].astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
        acc = (df[:, 2].astype(np.float64) * df[:, 2].astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
        passed = np.signbit(acc)
    else:
        r2f = (df[:, 0] * df[:, 0] + df[:, 1] * df[:, 1]).astype(np.float32)
        r2f = (r2f + df[:, 2] * df[:, 2]).astype(np.float32)
        passed = r2f < rc2f
    assert truth.sum() > 100000
    assert not np.any(truth & ~passed), f"{int(np.sum(truth & ~passed))} true neighbours rejected at E / rc = {ratio}"
    # the margin is not wasteful either: nothing farther than (1 + 3 margin) rc^2 passes
    assert not np.any(passed & (r2 > rc2 * (1.0 + 3.0 * margin)))
