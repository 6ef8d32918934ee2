"""Shared helpers for the parity tests: tolerances and oracle/GPU comparison."""
import numpy as np

# north_star: 1e-10 relative in f64, 1e-5 in f32, after one step.  "Relative" is taken against
# max(|ref|, scale) with scale = RMS of the reference field (interior accelerations
# cancel to ~0, so a pure per-element relative error is meaningless there).
TOL = {np.dtype(np.float64): 1e-10, np.dtype(np.float32): 1e-5}


def rel_err(got: np.ndarray, ref: np.ndarray) -> float:
    ref64 = ref.astype(np.float64); got64 = got.astype(np.float64)
    scale = float(np.sqrt(np.mean(ref64 * ref64))) if ref64.size else 0.0
    den = np.maximum(np.abs(ref64), scale)
    den[den == 0] = 1.0
    return float(np.max(np.abs(got64 - ref64) / den)) if ref64.size else 0.0


def assert_close(got, ref, what, tol=None):
    tol = TOL[np.dtype(ref.dtype)] if tol is None else tol
    e = rel_err(got, ref)
    if e > tol:    # say how loose the normalisation is for this field: the RMS floor next to the largest reference value
        r = ref.astype(np.float64)
        rms, big = float(np.sqrt(np.mean(r * r))) if r.size else 0.0, float(np.max(np.abs(r))) if r.size else 0.0
        k = int(np.argmax(np.abs(got.astype(np.float64) - r) / np.maximum(np.abs(r), rms if rms > 0 else 1.0))) if r.size else 0
        raise AssertionError(f"{what}: relative error {e:.3e} > {tol:.1e}  (error / max(|ref|, RMS); RMS(ref) = {rms:.3e}, max|ref| = {big:.3e}; "
                             f"worst element {k}: got {got.flat[k]!r}, ref {ref.flat[k]!r})")
    return e
