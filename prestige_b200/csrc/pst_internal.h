// pst_internal.h -- context layout and device helpers shared by the kernels.
// Product code: never includes or links anything under oracle/.
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only (v3): a no-op unless a profiler injects itself
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/prestige_b200.h"

// ---------------------------------------------------------------------------------------------
// host-side context
// ---------------------------------------------------------------------------------------------
struct PstArray {
    std::string name;
    int dtype = PST_F64;     // resolved (never PST_REAL)
    uint32_t flags = 0;
    int rows = 1;
    size_t esize = 8;
    char* buf[2] = {nullptr, nullptr};  // allocation base; owned view starts ghost_cap elements in
    int cur = 0;
};

struct PstGrid {
    int dim = 3;
    int n[3] = {1, 1, 1};    // cells per axis (x includes the two ghost layers when a communicator is attached)
    double lo[3] = {0, 0, 0};
    double cell = 1, inv_cell = 1;
    int sub = 1;             // linear keys: the FAST axis (z in 3D, y in 2D) is cut `sub` times finer than `cell`, so the fine
                             // cells of a column stay one contiguous run and a particle's stencil along it is +-sub fine cells
                             // (option "zsub"); n[fast] counts FINE cells, nc_fast the coarse ones
    int nc_fast = 1;
    int morton = 0;
    int bits = 0;            // morton: bits per axis
    int key_bits = 1;        // radix-sort end bit
    uint32_t ncells = 1;     // size of the key space (table has ncells + 1 entries)
    int cx_lo = 0, cx_hi = 0;  // clamp range of owned particles along x
};

struct PstComm;  // halo.cu

struct pst_ctx {
    pst_config cfg{};
    bool f64 = true;
    bool coupled = false;    // physics = WCSPH | DEM: rigid spheres in fluid (tags 0 fluid, 1 boundary, 2 solid)
    int dim = 3;
    cudaStream_t stream = nullptr;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;   // pst_upload_async / pst_download_async
    static constexpr int kRing = 32;                           // staging slots of the async transfers (lazily allocated): [0,24) uploads, [24,32) downloads
    char* ring[kRing] = {};
    cudaEvent_t ring_free[kRing] = {};                         // buffer k may be overwritten once this event completed
    cudaEvent_t ring_ready[kRing] = {};                        // producer -> consumer hand-over
    int ring_next_up = 0, ring_next_down = 0;
    std::vector<PstArray> arrays;
    std::map<std::string, int> index;
    std::map<std::string, double> params;
    std::map<std::string, int> options;
    uint64_t n = 0;          // owned particles
    int64_t n_ghost_l = 0, n_ghost_r = 0;   // ghost index range [-n_ghost_l, n + n_ghost_r); bounds only when !ghost_exact
    bool ghost_exact = true;                // false after a windowed halo exchange: the true counts live on the device
    uint64_t capacity = 0, ghost_cap = 0;
    PstGrid grid;
    // neighbour-search scratch
    uint32_t *keys_in = nullptr, *keys_out = nullptr, *vals_in = nullptr, *vals_out = nullptr;
    int32_t* cell_start = nullptr;   // ncells + 1, signed: left ghosts live at negative indices
    void* sort_tmp = nullptr;
    int32_t* scan_sums = nullptr;    // tile sums of the count scan (counting sort)
    size_t scan_sums_cap = 0;        // entries
    int32_t *nf_pos = nullptr, *nf_idx = nullptr;   // coupled contexts: compacted non-fluid particles of the sorted order (dem.cu)
    void* nf_rec = nullptr;                         // their (x, y, z, rad) records
    bool nf_dirty = true;                           // force rows of non-spheres are not known to be zero
    int32_t *wp_pos = nullptr, *wp_idx = nullptr;   // wall_pressure: compacted non-fluid particles of the owned range (wcsph.cu)
    uint32_t* big_list() const { return keys_out; }   // crowded-cell list reuses keys_out (unused by the counting sort)
    size_t sort_tmp_bytes = 0;
    char* stage = nullptr;           // capacity * 8 bytes
    int32_t* d_flags = nullptr;      // [0] contact overflow, [1] pair counter lo, ... (8 ints)
    unsigned long long* d_counters = nullptr;  // [0] pair count, [1] max cell count, ...
    int32_t* h_flags = nullptr;      // pinned mirror
    unsigned long long* h_counters = nullptr;
    bool ordered = false;            // device order != id order (a sort has happened)
    bool nbrs_valid = false;
    bool eos_valid = false;
    // packed neighbour-state records of the tiled pair kernel (wcsph.cu: rec_*): valid while rec_epoch == state_epoch;
    // everything that changes particle state (uploads, re-sorts, halo, stages, wall pressure ...) bumps state_epoch
    void* rec = nullptr;
    float* posf = nullptr;           // f32 copy of the positions relative to the grid origin (3 rows; written with the records): what the
                                     // pair kernel's tiles stage by TMA bulk copies
    size_t posf_stride = 0, posf_g4 = 0;   // row length / element 0 offset (ghost_cap rounded up to a multiple of 4: 16-byte aligned rows)
    uint64_t state_epoch = 1, rec_epoch = 0;
    bool ghost_eos_pending = false;  // the owned rows' EOS and records are current (fused permute), the ghost rows just arrived: tait_eos runs over them alone
    // tile list of the variant-3 pair kernel (wcsph_zrun.cuh): rebuilt when the cell table changed (build_epoch)
    void* ztiles = nullptr;
    size_t ztiles_cap = 0, ztile_off_cap = 0;
    int* d_ztile_count = nullptr;
    uint64_t ztiles_key = 0, build_epoch = 1;
    // every owned particle has the same mass / smoothing length (decided on the device, pst_uniform_refresh): the tiled
    // pair kernels then take them as constants instead of gathering m[j] and carrying the h-derived terms in registers
    bool m_uniform = false, h_uniform = false;
    double h_max = 0.0, rad_max = 0.0;   // largest smoothing length / radius of the owned particles (pst_uniform_refresh)
    double m_value = 0.0, h_value = 0.0;
    bool uni_dirty = true;           // m or h may have changed since the last check (upload, pst_array pointer, new particle set)
    unsigned long long *d_uni = nullptr, *h_uni = nullptr;   // min / max bit patterns of m and h (device, pinned mirror)
    bool hist_lag = false;           // contact-history rows still sit at their PRE-sort index (vals_out maps new -> old)
    uint64_t launches = 0;
    const void* last_kernel_fn = nullptr;        // every PST_LAUNCH records its kernel; the stage entry points keep theirs
    const void* pair_kernel_fn = nullptr;
    const void* contact_kernel_fn = nullptr;
    cudaEvent_t ev_stats = nullptr;  // completion of the async read-back of d_counters (occupied cells)
    bool stats_pending = false;
    // pst_step as a CUDA graph (option graph = 1, single-GPU contexts): TWO consecutive steps captured once -- the double-buffered
    // arrays are back on their original buffers after two flips, so the captured pointers stay valid -- and replayed
    cudaGraphExec_t step_graph = nullptr;
    uint64_t step_graph_key = 0, step_graph_parity = 0;
    uint64_t step_graph_launches = 0;      // kernel launches one replay stands for
    bool capturing = false;                // no host round trips while the stream is being captured
    PstComm* comm = nullptr;
    // multi-particle rigid bodies (rigid.cu): per-body records, field-major (rigid_core.h RbField), always double
    uint32_t n_bodies = 0;
    double* d_bodies = nullptr;
    int32_t* d_body_start = nullptr; // n_bodies + 1: member-list range of every body (members grouped by body, fixed order)
    double* d_body_vals = nullptr;   // per-member scratch of the body sums, field-major with stride body_members
    size_t body_members = 0;         // total member particles
    bool bodies_ready = false;       // pst_bodies_setup has run for the current particle set
    mutable std::string err;
};

pst_status pst_fail(const pst_ctx* ctx, pst_status s, const char* fmt, ...);
// NVTX range around a stage of the path (tracing): `nsys` timelines and `ncu --nvtx --nvtx-include "pst_apply/"`
// can then select a stage by name.  Costs one predictable branch when no tool is attached.
struct PstRange {
    explicit PstRange(const char* name) { nvtxRangePushA(name); }
    ~PstRange() { nvtxRangePop(); }
    PstRange(const PstRange&) = delete;
    PstRange& operator=(const PstRange&) = delete;
};
#define PST_CUDA(ctx, expr)                                                                       \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return pst_fail(ctx, PST_ECUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define PST_TRY(expr)                          \
    do {                                       \
        pst_status s__ = (expr);               \
        if (s__ != PST_OK) return s__;         \
    } while (0)
// every kernel launch goes through here so `launches` is an honest count
#define PST_LAUNCH(ctx, kern, grid, block, smem, ...)                                             \
    do {                                                                                          \
        auto k__ = kern;                                                                          \
        k__<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                             \
        (ctx)->launches++;                                                                        \
        (ctx)->last_kernel_fn = (const void*)k__;                                                 \
        PST_CUDA(ctx, cudaGetLastError());                                                        \
    } while (0)

PstArray* pst_find(pst_ctx* ctx, const char* name);
// owned-view pointer of the current / alternate buffer, row r
template <class T>
inline T* pst_ptr(pst_ctx* ctx, PstArray* a, int row = 0, int which = -1) {
    if (!a) return nullptr;
    const int w = which < 0 ? a->cur : which;
    const size_t stride = ctx->capacity + 2 * ctx->ghost_cap;
    return reinterpret_cast<T*>(a->buf[w] + (row * stride + ctx->ghost_cap) * a->esize);
}
template <class T>
inline T* pst_ptr(pst_ctx* ctx, const char* name, int row = 0, int which = -1) {
    return pst_ptr<T>(ctx, pst_find(ctx, name), row, which);
}
double pst_param(pst_ctx* ctx, const char* name, double dflt = 0.0);
int pst_option(pst_ctx* ctx, const char* name, int dflt = 0);

// stage implementations (one per .cu)
pst_status pst_nnps_build(pst_ctx* ctx);                                 // nnps.cu
pst_status pst_nnps_alloc(pst_ctx* ctx);
pst_status pst_nnps_alloc_table(pst_ctx* ctx);                           // (re)allocate the cell table for the current grid
pst_status pst_grid_finalize(pst_ctx* ctx);                              // api.cu: box + slab + zsub -> grid, then the table
pst_status pst_scan_exclusive(pst_ctx* ctx, int32_t* a, int m);        // in place, m entries (nnps.cu)
pst_status pst_sort_pairs_u32(pst_ctx* ctx, int n);                       // keys_in/vals_in -> keys_out/vals_out, stable (nnps.cu)
pst_status pst_resolve_history(pst_ctx* ctx);   // run the deferred k_remap_history, if any
pst_status pst_nnps_dump_pairs(pst_ctx* ctx, int mode, uint32_t* i, uint32_t* j, size_t cap, size_t* n_pairs);
pst_status pst_reorder_upload(pst_ctx* ctx, PstArray* a, int row, size_t n);   // stage -> array (by id)
pst_status pst_reorder_download(pst_ctx* ctx, PstArray* a, int row, size_t n); // array -> stage (by id)
pst_status pst_iota_ids(pst_ctx* ctx);
pst_status pst_eq1_apply(pst_ctx* ctx);                                   // eq1.cu
pst_status pst_wcsph_eos(pst_ctx* ctx);                                   // wcsph.cu
pst_status pst_uniform_refresh(pst_ctx* ctx);                             // wcsph.cu: are m and h uniform? (device-side reduction, once per change)
pst_status pst_wcsph_wall_pressure(pst_ctx* ctx);                         // dummy-particle pressure extrapolation (after the EOS)
pst_status pst_wcsph_forces(pst_ctx* ctx, bool continuity, bool momentum);
pst_status pst_wcsph_integrate(pst_ctx* ctx, double dt);
pst_status pst_dem_forces(pst_ctx* ctx);                                  // dem.cu
void pst_note_ghosts_changed(pst_ctx* ctx);                                // wcsph.cu: a halo exchange wrote the ghost rows
bool pst_wcsph_fused_permute(pst_ctx* ctx);                                // wcsph.cu: the WCSPH state goes through the fused permute + EOS + records kernel
pst_status pst_wcsph_permute_eos(pst_ctx* ctx, const uint32_t* perm, int n);
pst_status pst_check_cell_size(pst_ctx* ctx);                             // wcsph.cu: cell_size >= kfac max(h), >= 2 max(rad)
pst_status pst_dem_integrate(pst_ctx* ctx, double dt);
pst_status pst_coupled_integrate(pst_ctx* ctx, double dt);                // wcsph.cu
pst_status pst_rb_reduce(pst_ctx* ctx);                                   // rigid.cu: per-particle forces -> body force / torque
pst_status pst_rb_integrate(pst_ctx* ctx, double dt);                     // body stage + member scatter
pst_status pst_comm_destroy(pst_ctx* ctx);                                // halo.cu
void pst_comm_neighbours(pst_ctx* ctx, int* has_left, int* has_right);
pst_status pst_ghost_counts(pst_ctx* ctx, int64_t* nl, int64_t* nr);
pst_status pst_migrate(pst_ctx* ctx, int* arrivals);   // after a build pass with sentinel keys: ship leavers, append arrivals

// ---------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Rounded-to-nearest, never-contracted arithmetic for the cutoff test (bit-exact sets).
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

// r2 = dx*dx + dy*dy (+ dz*dz), left to right, no FMA.  SURVEY.md Appendix A.1.
template <int DIM, class R>
__device__ __forceinline__ R dist2(R dx, R dy, R dz) {
    R r2 = add_rn(mul_rn(dx, dx), mul_rn(dy, dy));
    if (DIM == 3) r2 = add_rn(r2, mul_rn(dz, dz));
    return r2;
}

template <class R>
struct GridDev {
    R lo[3];
    R inv[3];                // 1 / cell edge per axis (the fast axis of a subdivided grid is finer)
    R cell;
    int n[3];
    int sub;                 // stencil half-width along the fast axis, in (fine) cells
    int cx_lo, cx_hi;
    int bits;
};

template <class R>
__device__ __forceinline__ int cell_coord(R x, R lo, R inv, int cmin, int cmax) {
    int c = (int)floor((x - lo) * inv);
    return min(max(c, cmin), cmax);
}

__device__ __forceinline__ uint32_t spread3(uint32_t v) {  // 10 bits -> every 3rd bit
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ uint32_t spread2(uint32_t v) {  // 16 bits -> every 2nd bit
    v &= 0xFFFFu;
    v = (v | (v << 8)) & 0x00FF00FFu;
    v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

template <int DIM, bool MORTON, class R>
__device__ __forceinline__ uint32_t cell_key(const GridDev<R>& g, int cx, int cy, int cz) {
    if (MORTON) {
        if (DIM == 3) return (spread3(cx) << 2) | (spread3(cy) << 1) | spread3(cz);
        return (spread2(cx) << 1) | spread2(cy);
    }
    if (DIM == 3) return ((uint32_t)cx * g.n[1] + cy) * g.n[2] + cz;
    return (uint32_t)cx * g.n[1] + cy;
}

// Visit the contiguous index runs [b, e) holding every candidate j of a particle in cell (cx,cy,cz).
// Linear keys put the last axis fastest, so the 3 cells of a stencil column are ONE run
// (9 runs in 3D, 3 in 2D).  Morton keys have no such contiguity: 27 / 9 single-cell runs.
template <int DIM, bool MORTON, class R, class F>
__device__ __forceinline__ void for_each_run(const GridDev<R>& g, const int32_t* __restrict__ cell_start, int cx,
                                             int cy, int cz, F&& f) {
    if (DIM == 3) {
        for (int ax = max(cx - 1, 0); ax <= min(cx + 1, g.n[0] - 1); ++ax)
            for (int ay = max(cy - 1, 0); ay <= min(cy + 1, g.n[1] - 1); ++ay) {
                const int zl = max(cz - (MORTON ? 1 : g.sub), 0), zh = min(cz + (MORTON ? 1 : g.sub), g.n[2] - 1);
                if (MORTON) {
                    for (int az = zl; az <= zh; ++az) {
                        const uint32_t k = cell_key<3, true>(g, ax, ay, az);
                        f(cell_start[k], cell_start[k + 1]);
                    }
                } else {
                    f(cell_start[cell_key<3, false>(g, ax, ay, zl)], cell_start[cell_key<3, false>(g, ax, ay, zh) + 1]);
                }
            }
    } else {
        for (int ax = max(cx - 1, 0); ax <= min(cx + 1, g.n[0] - 1); ++ax) {
            const int yl = max(cy - (MORTON ? 1 : g.sub), 0), yh = min(cy + (MORTON ? 1 : g.sub), g.n[1] - 1);
            if (MORTON) {
                for (int ay = yl; ay <= yh; ++ay) {
                    const uint32_t k = cell_key<2, true>(g, ax, ay, 0);
                    f(cell_start[k], cell_start[k + 1]);
                }
            } else {
                f(cell_start[cell_key<2, false>(g, ax, yl, 0)], cell_start[cell_key<2, false>(g, ax, yh, 0) + 1]);
            }
        }
    }
}

template <class R>
inline GridDev<R> make_grid_dev(const PstGrid& g) {
    GridDev<R> d;
    for (int a = 0; a < 3; ++a) { d.lo[a] = (R)g.lo[a]; d.n[a] = g.n[a]; }
    d.cell = (R)g.cell;
    d.sub = g.sub;
    for (int a = 0; a < 3; ++a) d.inv[a] = (R)(a == g.dim - 1 ? g.inv_cell * g.sub : g.inv_cell);
    d.cx_lo = g.cx_lo; d.cx_hi = g.cx_hi;
    d.bits = g.bits;
    return d;
}

// dispatch a <Real, DIM, MORTON> template on the context configuration
#define PST_DISPATCH(ctx, FN, ...)                                                                  \
    ((ctx)->f64 ? ((ctx)->dim == 3 ? ((ctx)->grid.morton ? FN<double, 3, true>(__VA_ARGS__) : FN<double, 3, false>(__VA_ARGS__)) \
                                   : ((ctx)->grid.morton ? FN<double, 2, true>(__VA_ARGS__) : FN<double, 2, false>(__VA_ARGS__))) \
                : ((ctx)->dim == 3 ? ((ctx)->grid.morton ? FN<float, 3, true>(__VA_ARGS__) : FN<float, 3, false>(__VA_ARGS__))   \
                                   : ((ctx)->grid.morton ? FN<float, 2, true>(__VA_ARGS__) : FN<float, 2, false>(__VA_ARGS__))))

#endif  // __CUDACC__
