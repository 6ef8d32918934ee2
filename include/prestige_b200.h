/* prestige_b200.h -- C ABI of the B200-native particle hot path.
 *
 * Drop-in boundary for dineshadepu/prestige (reference mounted at /root/reference).
 * The reference today has ONE back-end, prestige::codegen::simple_cpu
 * (prestige/src/codegen/mod.rs:1), a free function taking &FusedEquations
 * (prestige/src/codegen/simple_cpu.rs:3) and emitting the all-pairs gather loop
 *     for i in 0..n { for j in 0..n { <bodies> } }      (simple_cpu.rs:7-16)
 * over caller-owned, contiguous, name-identified f64 slices (prestige/src/lib.rs:8,
 * prestige/src/equations/fuse.rs:4-12).  It defines no FFI.  The entry points
 * below are what a sibling back-end `prestige::codegen::b200` binds (see
 * INTEGRATION.md for the Rust `extern "C"` block and the ctypes stub):
 *
 *   reference concept                         (file:line)                 -> entry point
 *   ----------------------------------------------------------------------------------------
 *   slices named by identifier string         fuse.rs:6-8, lib.rs:8       -> pst_array_create / pst_array /
 *                                                                            pst_upload / pst_download
 *   loop bound `n`                            simple_cpu.rs:7-8           -> pst_set_count
 *   FusedEquations{reads,writes,bodies}       fuse.rs:4-12, fuse():14-40  -> pst_apply(eq_names[], n_eq)
 *   generate_simple_cpu(&FusedEquations)      simple_cpu.rs:3             -> pst_apply (runs the fused CUDA kernel
 *                                                                            instead of returning loop text)
 *   `for j in 0..n` (all pairs, j == i incl.) simple_cpu.rs:8             -> pst_build_neighbours (+ cutoff inside the
 *                                                                            bodies; eq1 keeps the literal all-pairs loop)
 *   eq1: force[i] += mass[j]                  lib.rs:7-12                 -> pst_apply({"eq1"})
 *   (no reference code: integrator)           --                          -> pst_step
 *
 * Conventions: every call returns pst_status and never throws or aborts.  Device
 * memory is owned by the context; host buffers are caller-owned and borrowed for
 * the duration of the call.  Host arrays are always in *id order* (the order of
 * the first upload); the device keeps particles in cell order and un-permutes in
 * pst_download.  Calls are asynchronous on the context's stream except
 * pst_download / pst_sync / pst_dump_pairs / pst_get_stat.  A context is not
 * thread-safe; distinct contexts may be used from distinct threads.  One context
 * drives one GPU; multi-GPU = one context per rank + pst_comm_init.
 */
#ifndef PRESTIGE_B200_H
#define PRESTIGE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define PST_API
#else
#define PST_API __attribute__((visibility("default")))
#endif

typedef struct pst_ctx pst_ctx;

typedef enum pst_status {
    PST_OK = 0,
    PST_EINVAL = 1,     /* bad argument / unknown name / wrong dtype */
    PST_ENOMEM = 2,     /* device or host allocation failed, or capacity exceeded */
    PST_ECUDA = 3,      /* CUDA runtime error (pst_last_error has the text) */
    PST_ENCCL = 4,      /* NCCL error, or NCCL not loadable */
    PST_EOVERFLOW = 5,  /* more than max_contacts contacts on a particle / pair buffer too small */
    PST_ESTATE = 6      /* call sequence error (e.g. pst_apply before pst_build_neighbours) */
} pst_status;

typedef enum pst_dtype { PST_F32 = 0, PST_F64 = 1, PST_U32 = 2, PST_I32 = 3, PST_REAL = 15 /* = the context's real */ } pst_dtype;
typedef enum pst_key { PST_KEY_LINEAR = 0, PST_KEY_MORTON = 1 } pst_key;

/* physics bit mask: which standard arrays a context registers at creation */
#define PST_PHYS_NONE 0u
#define PST_PHYS_WCSPH 1u /* x y [z] u v [w] rho m h tag | p au av [aw] arho */
#define PST_PHYS_DEM 2u   /* x y z u v w wx wy wz rad m inertia tag | fx fy fz tx ty tz | hist_* */
/* PST_PHYS_WCSPH | PST_PHYS_DEM = coupled SPH-DEM, rigid spheres in fluid (3D): the union of both array sets in ONE
 * cell grid (cell_size >= max(kfac h, 2 rad)).  tag 0 = fluid, 1 = static boundary (SPH dummy particle and, with
 * rad > 0, DEM wall sphere), 2 = solid sphere.  SPH pairs need a fluid member and see a solid through its displaced
 * fluid mass m rho0 / rho_solid; contacts are evaluated for solid i against non-fluid j; pst_step moves a solid by
 * F_contact + m rho0/rho_solid (a_sph - g) + m g.  DESIGN.md "Coupled formulation". */

/* array flags */
#define PST_ARRAY_PERSISTENT 1u /* state: follows its particle through every re-sort */
#define PST_ARRAY_OUTPUT 2u     /* per-step result: overwritten by pst_apply, not permuted */

typedef struct pst_config {
    uint32_t struct_size;  /* = sizeof(pst_config), ABI guard */
    int32_t device;        /* CUDA device ordinal */
    int32_t dim;           /* 2 | 3 */
    int32_t real;          /* PST_F32 | PST_F64: type of every PST_REAL array */
    int32_t key;           /* pst_key */
    int32_t max_contacts;  /* K history slots per particle (DEM); 0 = none */
    uint32_t physics;      /* PST_PHYS_* mask */
    uint32_t reserved;
    uint64_t capacity;     /* max owned particles */
    uint64_t ghost_capacity; /* max ghost particles per face (multi-GPU only; 0 otherwise) */
    double lo[3], hi[3];   /* domain box; particles outside are clamped into the edge cells */
    double cell_size;      /* cell edge, must be >= the largest cutoff */
} pst_config;

PST_API const char* pst_version(void);

/* ---- context ---------------------------------------------------------------------------- */
PST_API pst_status pst_create(const pst_config* cfg, pst_ctx** out);
PST_API void pst_destroy(pst_ctx* ctx);
/* ctx-owned string, valid until the next call on ctx; ctx == NULL returns the last create error */
PST_API const char* pst_last_error(const pst_ctx* ctx);
/* the context's cudaStream_t (as void*) so a host can time or chain work on it */
PST_API void* pst_stream(pst_ctx* ctx);
PST_API pst_status pst_sync(pst_ctx* ctx);

/* named scalar parameters: rho0 c0 gamma alpha beta kfac gx gy gz | dem_model kn gn kt gt mu dt Estar Gstar erest | rho_solid |
 * boundary_model (0: boundaries and solids are dynamic SPH particles; 1: dummy-particle wall pressure, see pst_apply) */
PST_API pst_status pst_set_param(pst_ctx* ctx, const char* name, double value);
PST_API pst_status pst_get_param(pst_ctx* ctx, const char* name, double* value);

/* ---- particle arrays (name = the identifier the reference's EquationIR carries) ---------- */
/* number of owned particles; resets ids to 0..n-1, the order to identity, and clears contact history */
PST_API pst_status pst_set_count(pst_ctx* ctx, uint64_t n);
PST_API pst_status pst_get_count(pst_ctx* ctx, uint64_t* n_owned, uint64_t* n_ghost);
PST_API pst_status pst_array_create(pst_ctx* ctx, const char* name, int dtype, uint32_t flags);
/* device pointer (current buffer, cell order), element count per row, dtype, rows.  The pointer is for reading and for
 * chaining device work; the library tracks one property of uploaded data -- whether all masses `m` are equal, which lets the
 * fused pair kernel skip the m[j] gather -- so a caller that WRITES `m` through this pointer must say so with
 * pst_set_option(ctx, "uniform_mass", 0) (or upload `m` again). */
PST_API pst_status pst_array(pst_ctx* ctx, const char* name, void** dev_ptr, size_t* n, int* dtype, int* rows);
/* host buffers hold rows * n elements, row-major, particle index = id */
PST_API pst_status pst_upload(pst_ctx* ctx, const char* name, const void* host, size_t n);
PST_API pst_status pst_download(pst_ctx* ctx, const char* name, void* host, size_t n);
/* Asynchronous variants: copies run on dedicated H2D / D2H streams and overlap the compute stream (and each other:
 * PCIe is full duplex).  `host` must be pinned (pst_host_alloc) and must stay untouched until pst_sync (uploads) or
 * pst_wait_transfers / pst_sync (downloads).  Fall back to the synchronous path for multi-row arrays. */
PST_API pst_status pst_upload_async(pst_ctx* ctx, const char* name, const void* host, size_t n);
PST_API pst_status pst_download_async(pst_ctx* ctx, const char* name, void* host, size_t n);
/* wait until every pst_download_async issued so far has landed in host memory (does not wait for compute) */
PST_API pst_status pst_wait_transfers(pst_ctx* ctx);
/* pinned host memory for fast transfers (optional for the synchronous calls; any host memory works there) */
PST_API void* pst_host_alloc(size_t bytes);
PST_API void pst_host_free(void* p);

/* ---- the hot path ------------------------------------------------------------------------ */
/* cell keys -> radix sort -> cell start table -> permute persistent state (+ history remap) */
PST_API pst_status pst_build_neighbours(pst_ctx* ctx);
/* fuse(): run the hand-written fused kernels for this equation SET (order and duplicates in eq_names do not matter: the
 * library runs tait_eos, wall_pressure, the fused continuity / momentum pair kernel, dem_contact, body_reduce in that order).
 * Known names: "eq1" | "tait_eos" "wall_pressure" "continuity" "momentum" | "dem_contact" | "body_reduce"
 * "wall_pressure" (no reference code): every non-fluid particle (tag != 0) takes the pressure
 * extrapolated from its fluid neighbours, p_w = sum (p_f + rho_f g . x_wf) W_wf / sum W_wf, and the density the EOS
 * maps to it (p, por2 and the state array rho are overwritten for those rows); runs after tait_eos and before the
 * pair kernel.  With the parameter boundary_model = 1 pst_step includes it and the integrator no longer advances the
 * density of non-fluid particles.  One GPU only (no communicator).                                              */
PST_API pst_status pst_apply(pst_ctx* ctx, const char* const* eq_names, int n_eq);
/* parity hook: neighbour set (mode 0: r2 < (kfac h_i)^2) or contact set (mode 1: r2 < (R_i+R_j)^2)
 * as stable ids; order unspecified.  *n_pairs is always the true count; PST_EOVERFLOW if > cap. */
PST_API pst_status pst_dump_pairs(pst_ctx* ctx, int mode, uint32_t* i, uint32_t* j, size_t cap, size_t* n_pairs);
/* n_steps of: build_neighbours -> forces -> integrate (semi-implicit Euler) */
PST_API pst_status pst_step(pst_ctx* ctx, double dt, int n_steps);
/* integrator stage alone (forces must be current) */
PST_API pst_status pst_integrate(pst_ctx* ctx, double dt);

/* counters: n_cells max_cell_count launches pairs_tested contacts_total key_bits ... */
PST_API pst_status pst_get_stat(pst_ctx* ctx, const char* name, double* value);
/* mangled symbol of the kernel the last pst_apply launched for `stage` ("pair": the fused continuity / momentum kernel,
 * "contact": the DEM contact kernel), copied into buf (NUL-terminated, truncated to cap).  For measurement scripts: it is
 * the name a profiler reports, so a profile can be tied to the kernel a run actually used.  PST_ESTATE before the first launch. */
PST_API pst_status pst_kernel_name(pst_ctx* ctx, const char* stage, char* buf, size_t cap);
/* select a kernel variant (A/B measurement): name "force_kernel" value 0 = per-particle gather, 1 = warp per cell, 2 = tiled
 * lists, 3 = tiled z-runs + bit masks (default).  "uniform_mass" 0 = always gather m[j] (default 1: skip the gather when every uploaded mass is equal);
 * "uniform_mass_global" 1 = multi-GPU: the caller guarantees that EVERY rank uploaded the same single mass value, so the
 * skip also applies to ghosts and migrants (default 0: with a communicator m[j] is always gathered). */
PST_API pst_status pst_set_option(pst_ctx* ctx, const char* name, int value);

/* ---- multi-particle rigid bodies in a coupled context (no reference code) ---------------------
 * A body is the set of tag-2 particles whose persistent i32 array `body` holds its index (-1 = none: the particle is
 * its own body).  Members keep taking part in the force loop one by one (SPH pairs with fluid, contacts with
 * non-members); pst_apply({"body_reduce"}) sums their forces to one force and one torque about the centre of mass per
 * body, and pst_integrate / pst_step advance the bodies (semi-implicit Euler, rotation matrix by Rodrigues' formula)
 * and set x = X + R r0, v = V + omega x (x - X), spin = omega on every member.  One GPU only (no communicator). */
/* registers `n_bodies` bodies and the particle arrays body (i32, all -1), bx0 by0 bz0 (body-frame offsets) and bpos
 * (i32, position in the body-grouped member list: fixes the order of every body sum, so they are deterministic) */
PST_API pst_status pst_bodies_create(pst_ctx* ctx, uint32_t n_bodies);
/* after uploading body, x y z, u v w, m, inertia: mass, centre of mass, mass-weighted velocity, offsets and inertia
 * tensor of every body; orientation = identity, omega = 0; members take the rigid motion */
PST_API pst_status pst_bodies_setup(pst_ctx* ctx);
/* checkpoint restore: body, bpos, bx0 by0 bz0 were uploaded as saved; rebuilds the member-list ranges only (the saved
 * member order, and with it the bit pattern of every body sum, is kept); the records follow through pst_bodies_state */
PST_API pst_status pst_bodies_restore(pst_ctx* ctx);
/* read (write = 0) / write (write = 1) a per-body quantity, host data always double, body-major:
 * "mass" [nb] | "cm" "vel" "omega" "force" "torque" [nb][3] | "rot" [nb][9] row-major | "inertia0" [nb][6] xx yy zz xy xz yz.
 * Writable: everything but force and torque (members are moved accordingly; mass and inertia0 are normally left to
 * pst_bodies_setup and only written when a checkpoint is restored). */
PST_API pst_status pst_bodies_state(pst_ctx* ctx, const char* name, double* host, size_t n, int write);

/* ---- multi-GPU: slab decomposition along x ------------------------------------------------------------------
 * One context per rank.  Ghost layers travel by peer-memory stores over NVLink (cudaIpc windows + an epoch word; option
 * halo_impl = 2, the default) where the GPUs can map each other's memory -- probed once and agreed over all ranks -- and as
 * one packed ncclSend/ncclRecv message per neighbour (halo_impl = 1) otherwise; halo_impl = 0 is the exact per-array NCCL
 * exchange.  Particles that leave the slab migrate to the neighbour rank inside pst_build_neighbours (NCCL).  With a
 * communicator attached host transfers are in DEVICE order and `id` is a caller-supplied global label. */
#define PST_COMM_ID_BYTES 128
PST_API pst_status pst_comm_unique_id(void* id_bytes /* PST_COMM_ID_BYTES */);
/* rank r owns global cell layers [ix_lo, ix_hi) of the cfg box; neighbours are r-1 and r+1 */
PST_API pst_status pst_comm_init(pst_ctx* ctx, const void* id_bytes, int rank, int n_ranks);
/* after pst_build_neighbours: send edge cell layers, receive ghosts, extend the cell table */
PST_API pst_status pst_halo_exchange(pst_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PRESTIGE_B200_H */
