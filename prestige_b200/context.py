"""Host-side particle context: a thin object over the pst_* C ABI.

Mirrors what the reference's pair loop is given (prestige/src/lib.rs:8): contiguous,
caller-owned slices identified by *name* (prestige/src/equations/fuse.rs:6-8) and a
loop bound n (prestige/src/codegen/simple_cpu.rs:7-8).  Device buffers are SoA and
cell-ordered; host arrays are always in id order.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

_NP = {L.PST_F32: np.float32, L.PST_F64: np.float64, L.PST_U32: np.uint32, L.PST_I32: np.int32}


class Context:
    def __init__(self, *, dim: int, lo, hi, cell_size: float, capacity: int, real=np.float64, physics: str | int = 0,
                 key: str = "linear", max_contacts: int = 0, device: int = 0, ghost_capacity: int = 0):
        self._lib = L.load()
        self._h = C.c_void_p()
        cfg = L.PstConfig()
        cfg.struct_size = C.sizeof(L.PstConfig)
        cfg.device, cfg.dim = device, dim
        self.real = np.dtype(real)
        cfg.real = L.PST_F64 if self.real == np.float64 else L.PST_F32
        cfg.key = L.PST_KEY_MORTON if key == "morton" else L.PST_KEY_LINEAR
        cfg.max_contacts = max_contacts
        if isinstance(physics, str):
            physics = sum({"wcsph": L.PST_PHYS_WCSPH, "dem": L.PST_PHYS_DEM, "none": 0, "": 0}[p] for p in physics.split("+"))
        cfg.physics = physics
        cfg.capacity, cfg.ghost_capacity = capacity, ghost_capacity
        for a in range(3):
            cfg.lo[a] = lo[a] if a < len(lo) else 0.0
            cfg.hi[a] = hi[a] if a < len(hi) else 0.0
        cfg.cell_size = cell_size
        self.dim, self.max_contacts, self.n = dim, max_contacts, 0
        st = self._lib.pst_create(C.byref(cfg), C.byref(self._h))
        if st != L.PST_OK:
            raise L.PstError(st, self._lib.pst_last_error(None).decode())

    # -- plumbing ---------------------------------------------------------------------------
    def _ck(self, st: int):
        if st != L.PST_OK:
            raise L.PstError(st, self._lib.pst_last_error(self._h).decode())

    def close(self):
        if self._h:
            self._lib.pst_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        return self._h

    @property
    def stream(self) -> int:
        return self._lib.pst_stream(self._h) or 0

    def sync(self):
        self._ck(self._lib.pst_sync(self._h))

    # -- parameters, counts, arrays ---------------------------------------------------------
    def set_params(self, **kw):
        for k, v in kw.items():
            self._ck(self._lib.pst_set_param(self._h, k.encode(), float(v)))

    def get_param(self, name: str) -> float:
        v = C.c_double()
        self._ck(self._lib.pst_get_param(self._h, name.encode(), C.byref(v)))
        return v.value

    def set_option(self, name: str, value: int):
        self._ck(self._lib.pst_set_option(self._h, name.encode(), int(value)))

    def set_count(self, n: int):
        self._ck(self._lib.pst_set_count(self._h, n))
        self.n = n

    def array_create(self, name: str, dtype="real", persistent=True):
        dt = L.PST_REAL if dtype == "real" else {np.dtype(np.float32): L.PST_F32, np.dtype(np.float64): L.PST_F64,
                                                   np.dtype(np.uint32): L.PST_U32, np.dtype(np.int32): L.PST_I32}[np.dtype(dtype)]
        self._ck(self._lib.pst_array_create(self._h, name.encode(), dt, L.PST_ARRAY_PERSISTENT if persistent else L.PST_ARRAY_OUTPUT))

    def array_info(self, name: str):
        """(None, n, dtype, rows): shape only.  The device pointer is NOT requested: handing it out tells the library the
        caller may write through it (packed copies and the uniform-mass decision are then refreshed); see array_ptr."""
        n, dt, rows = C.c_size_t(), C.c_int(), C.c_int()
        self._ck(self._lib.pst_array(self._h, name.encode(), None, C.byref(n), C.byref(dt), C.byref(rows)))
        return None, n.value, _NP[dt.value], rows.value

    def array_ptr(self, name: str) -> int:
        """Device pointer of the current buffer (cell order), for chaining device work."""
        p, n, dt, rows = C.c_void_p(), C.c_size_t(), C.c_int(), C.c_int()
        self._ck(self._lib.pst_array(self._h, name.encode(), C.byref(p), C.byref(n), C.byref(dt), C.byref(rows)))
        return p.value

    def has_array(self, name: str) -> bool:
        n, dt, rows = C.c_size_t(), C.c_int(), C.c_int()
        return self._lib.pst_array(self._h, name.encode(), None, C.byref(n), C.byref(dt), C.byref(rows)) == L.PST_OK

    def upload(self, name: str, host: np.ndarray):
        _, _, dt, rows = self.array_info(name)
        a = np.ascontiguousarray(host, dtype=dt)
        assert a.size == rows * self.n, f"{name}: expected {rows}x{self.n} elements, got {a.shape}"
        self._ck(self._lib.pst_upload(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), self.n))

    def upload_ptr(self, name: str, ptr: int):
        """Upload from a raw host pointer (e.g. pinned memory) holding rows*n elements."""
        self._ck(self._lib.pst_upload(self._h, name.encode(), C.c_void_p(ptr), self.n))

    def download(self, name: str, out: np.ndarray | None = None) -> np.ndarray:
        _, _, dt, rows = self.array_info(name)
        if out is None:
            out = np.empty((rows, self.n) if rows > 1 else self.n, dt)
        assert out.dtype == dt and out.size == rows * self.n and out.flags["C_CONTIGUOUS"]
        self._ck(self._lib.pst_download(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), self.n))
        return out

    def download_ptr(self, name: str, ptr: int):
        self._ck(self._lib.pst_download(self._h, name.encode(), C.c_void_p(ptr), self.n))

    def upload_async(self, name: str, ptr: int):
        """Asynchronous upload from PINNED host memory (pst_host_alloc); valid data required until sync()."""
        self._ck(self._lib.pst_upload_async(self._h, name.encode(), C.c_void_p(ptr), self.n))

    def download_async(self, name: str, ptr: int):
        """Asynchronous download into PINNED host memory; complete after wait_transfers() or sync()."""
        self._ck(self._lib.pst_download_async(self._h, name.encode(), C.c_void_p(ptr), self.n))

    def wait_transfers(self):
        self._ck(self._lib.pst_wait_transfers(self._h))

    def load_block(self, block, arrays=None):
        """set_count + params + upload of every array of a synth.Block this context knows."""
        self.set_count(block.n)
        known = {k: v for k, v in block.params.items()}
        self.set_params(**known)
        if "body" in block.arrays and (arrays is None or "body" in arrays) and not self.has_array("body"):
            self.bodies_create(block.meta["n_bodies"])
        for k, v in block.arrays.items():
            if arrays is not None and k not in arrays:
                continue
            if self.has_array(k):
                self.upload(k, v)
        if "body" in block.arrays and (arrays is None or "body" in arrays):
            self.bodies_setup()

    # -- the hot path -----------------------------------------------------------------------
    def refresh_count(self) -> int:
        """Owned-particle count as the library sees it (changes when particles migrate between ranks)."""
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self._lib.pst_get_count(self._h, C.byref(a), C.byref(b)))
        self.n = a.value
        return self.n

    def build_neighbours(self):
        self._ck(self._lib.pst_build_neighbours(self._h))
        self.refresh_count()

    def apply(self, names):
        arr = (C.c_char_p * len(names))(*[s.encode() for s in names])
        self._ck(self._lib.pst_apply(self._h, arr, len(names)))

    def dump_pairs(self, mode: int = 0, cap: int | None = None) -> np.ndarray:
        cap = cap or max(1024, 128 * self.n)
        i = np.empty(cap, np.uint32); j = np.empty(cap, np.uint32)
        cnt = C.c_size_t()
        self._ck(self._lib.pst_dump_pairs(self._h, mode, i.ctypes.data_as(C.c_void_p), j.ctypes.data_as(C.c_void_p), cap, C.byref(cnt)))
        pr = np.stack([i[:cnt.value], j[:cnt.value]], axis=1)
        return pr[np.lexsort((pr[:, 1], pr[:, 0]))]

    def step(self, dt: float, n_steps: int = 1):
        self._ck(self._lib.pst_step(self._h, dt, n_steps))
        self.refresh_count()

    def integrate(self, dt: float):
        self._ck(self._lib.pst_integrate(self._h, dt))

    def stat(self, name: str) -> float:
        v = C.c_double()
        self._ck(self._lib.pst_get_stat(self._h, name.encode(), C.byref(v)))
        return v.value

    def kernel_name(self, stage: str = "pair", demangle: bool = True) -> str:
        """Symbol of the kernel the last apply() launched for `stage` ("pair" | "contact"); demangled with c++filt when available."""
        buf = C.create_string_buffer(1024)
        self._ck(self._lib.pst_kernel_name(self._h, stage.encode(), buf, len(buf)))
        name = buf.value.decode()
        if demangle:
            import subprocess
            try:
                name = subprocess.run(["c++filt", name], capture_output=True, text=True, timeout=10).stdout.strip() or name
            except Exception:
                pass
        return name

    # -- multi-particle rigid bodies (coupled contexts) ---------------------------------------
    _BODY_WIDTH = {"mass": 1, "cm": 3, "vel": 3, "omega": 3, "rot": 9, "inertia0": 6, "force": 3, "torque": 3}

    def bodies_create(self, n_bodies: int):
        self._ck(self._lib.pst_bodies_create(self._h, int(n_bodies)))
        self.n_bodies = int(n_bodies)

    def bodies_setup(self):
        """After `body`, positions, velocities, m and inertia are uploaded: mass, centre of mass, velocity, offsets and
        inertia tensor of every body; members take the rigid motion."""
        self._ck(self._lib.pst_bodies_setup(self._h))

    def bodies_restore(self):
        """Checkpoint restore: body, bpos, bx0 by0 bz0 uploaded as saved; the records follow through body_set."""
        self._ck(self._lib.pst_bodies_restore(self._h))

    def body_get(self, name: str) -> np.ndarray:
        w = self._BODY_WIDTH[name]
        out = np.empty((self.n_bodies, w) if w > 1 else self.n_bodies, np.float64)
        self._ck(self._lib.pst_bodies_state(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), out.size, 0))
        return out

    def body_set(self, name: str, value: np.ndarray):
        a = np.ascontiguousarray(value, np.float64)
        assert a.size == self.n_bodies * self._BODY_WIDTH[name]
        self._ck(self._lib.pst_bodies_state(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), a.size, 1))

    # -- multi-GPU --------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        lib = L.load()
        buf = C.create_string_buffer(L.PST_COMM_ID_BYTES)
        st = lib.pst_comm_unique_id(buf)
        if st != L.PST_OK:
            raise L.PstError(st, lib.pst_last_error(None).decode())
        return buf.raw

    def comm_init(self, uid: bytes, rank: int, n_ranks: int):
        buf = C.create_string_buffer(uid, L.PST_COMM_ID_BYTES)
        self._ck(self._lib.pst_comm_init(self._h, buf, rank, n_ranks))

    def halo_exchange(self):
        self._ck(self._lib.pst_halo_exchange(self._h))


def context_for_block(block, real=np.float64, key="linear", capacity=None, device=0, ghost_capacity=0, lo=None, hi=None) -> Context:
    """A context sized for a synth.Block (its box, cell size, physics and history depth)."""
    return Context(dim=block.dim, lo=lo or block.lo, hi=hi or block.hi, cell_size=block.cell_size,
                   capacity=capacity or block.n, real=real, physics=block.physics, key=key,
                   max_contacts=block.max_contacts, device=device, ghost_capacity=ghost_capacity)
