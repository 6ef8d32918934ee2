// api.cu -- the extern "C" surface declared in include/prestige_b200.h.
// Context, named arrays, parameters, transfers, equation-set dispatch (the reference's fuse() made
// real: prestige/src/equations/fuse.rs:14-40 -> one fused kernel), and the step loop.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <set>
#include <vector>

#include "pst_internal.h"

static thread_local std::string g_create_err;

pst_status pst_fail(const pst_ctx* ctx, pst_status s, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_err = buf;
    return s;
}

PstArray* pst_find(pst_ctx* ctx, const char* name) {
    auto it = ctx->index.find(name);
    return it == ctx->index.end() ? nullptr : &ctx->arrays[it->second];
}
double pst_param(pst_ctx* ctx, const char* name, double dflt) {
    auto it = ctx->params.find(name);
    return it == ctx->params.end() ? dflt : it->second;
}
int pst_option(pst_ctx* ctx, const char* name, int dflt) {
    auto it = ctx->options.find(name);
    return it == ctx->options.end() ? dflt : it->second;
}

static size_t dtype_size(int dt) { return (dt == PST_F64) ? 8 : 4; }

// Derive the cell grid from the configured box, the slab (when a communicator is attached: one ghost cell layer on each x
// face) and the fast-axis subdivision `grid.sub`; (re)allocate the cell table for it.
pst_status pst_grid_finalize(pst_ctx* ctx) {
    PstGrid& g = ctx->grid;
    const pst_config& cfg = ctx->cfg;
    g.dim = cfg.dim;
    g.cell = cfg.cell_size;
    g.inv_cell = 1.0 / cfg.cell_size;
    g.morton = cfg.key == PST_KEY_MORTON;
    if (g.morton && g.sub != 1) return pst_fail(ctx, PST_EINVAL, "zsub > 1 needs linear keys");
    double ncell_total = 1;
    int nmax = 1;
    for (int a = 0; a < 3; ++a) {
        g.lo[a] = cfg.lo[a];
        g.n[a] = 1;
        if (a < cfg.dim) {
            const double ext = cfg.hi[a] - cfg.lo[a];
            if (!(ext >= 0)) return pst_fail(ctx, PST_EINVAL, "hi[%d] < lo[%d]", a, a);
            g.n[a] = std::max(1, (int)std::ceil(ext / cfg.cell_size));
        }
    }
    g.cx_lo = 0; g.cx_hi = g.n[0] - 1;
    if (ctx->comm) {   // slab: whole number of cells along x (checked by pst_comm_init) + the two ghost layers
        g.n[0] = (int)std::llround((cfg.hi[0] - cfg.lo[0]) / g.cell) + 2;
        g.lo[0] = cfg.lo[0] - g.cell;
        g.cx_lo = 1;
        g.cx_hi = g.n[0] - 2;
    }
    g.nc_fast = g.n[cfg.dim - 1];
    g.n[cfg.dim - 1] *= g.sub;
    for (int a = 0; a < 3; ++a) { ncell_total *= g.n[a]; nmax = std::max(nmax, g.n[a]); }
    if (g.morton) {
        g.bits = 1;
        while ((1 << g.bits) < nmax) ++g.bits;
        if (g.bits > (cfg.dim == 3 ? 10 : 15)) return pst_fail(ctx, PST_EINVAL, "grid too large for Morton keys");
        g.key_bits = g.bits * cfg.dim;
        g.ncells = 1u << g.key_bits;
    } else {
        if (ncell_total >= 2147483647.0) return pst_fail(ctx, PST_EINVAL, "grid has too many cells (%g)", ncell_total);
        g.ncells = (uint32_t)ncell_total;
        g.key_bits = 1;
        while ((1ull << g.key_bits) < g.ncells) ++g.key_bits;
    }
    ctx->nbrs_valid = false;
    ctx->build_epoch++;
    ctx->params.erase("_ppc");
    return pst_nnps_alloc_table(ctx);
}

// m and h decide which pair kernel runs (uniform values become constants); whoever may have changed them marks the check stale
static void note_uniform_dirty(pst_ctx* ctx, const PstArray* a) {
    if (a->name == "m" || a->name == "h" || a->name == "rad") ctx->uni_dirty = true;
}

static pst_status array_create(pst_ctx* ctx, const char* name, int dtype, uint32_t flags, int rows) {
    if (!name || !*name) return pst_fail(ctx, PST_EINVAL, "array name is empty");
    if (pst_find(ctx, name)) return pst_fail(ctx, PST_EINVAL, "array '%s' already exists", name);
    if (dtype == PST_REAL) dtype = ctx->f64 ? PST_F64 : PST_F32;
    if (dtype < PST_F32 || dtype > PST_I32) return pst_fail(ctx, PST_EINVAL, "array '%s': bad dtype %d", name, dtype);
    PstArray a;
    a.name = name; a.dtype = dtype; a.flags = flags; a.rows = rows; a.esize = dtype_size(dtype);
    const size_t bytes = (size_t)rows * (ctx->capacity + 2 * ctx->ghost_cap) * a.esize;
    const int nbuf = (flags & PST_ARRAY_PERSISTENT) ? 2 : 1;
    for (int b = 0; b < nbuf; ++b) {
        cudaError_t e = cudaMalloc((void**)&a.buf[b], bytes);
        if (e != cudaSuccess) return pst_fail(ctx, PST_ENOMEM, "cudaMalloc(%zu) for '%s': %s", bytes, name, cudaGetErrorString(e));
        PST_CUDA(ctx, cudaMemsetAsync(a.buf[b], 0, bytes, ctx->stream));
    }
    if (nbuf == 1) a.buf[1] = nullptr;
    ctx->index[name] = (int)ctx->arrays.size();
    ctx->arrays.push_back(a);
    return PST_OK;
}

extern "C" {

const char* pst_version(void) { return "prestige_b200 0.1 (sm_100a)"; }

const char* pst_last_error(const pst_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

pst_status pst_create(const pst_config* cfg, pst_ctx** out) {
    if (!cfg || !out) return pst_fail(nullptr, PST_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->struct_size != sizeof(pst_config))
        return pst_fail(nullptr, PST_EINVAL, "pst_config.struct_size %u != %zu (ABI mismatch)", cfg->struct_size, sizeof(pst_config));
    if (cfg->dim != 2 && cfg->dim != 3) return pst_fail(nullptr, PST_EINVAL, "dim must be 2 or 3");
    if (cfg->real != PST_F32 && cfg->real != PST_F64) return pst_fail(nullptr, PST_EINVAL, "real must be PST_F32 or PST_F64");
    if (cfg->key != PST_KEY_LINEAR && cfg->key != PST_KEY_MORTON) return pst_fail(nullptr, PST_EINVAL, "bad key mode");
    if (!(cfg->cell_size > 0)) return pst_fail(nullptr, PST_EINVAL, "cell_size must be > 0");
    if (cfg->capacity == 0 || cfg->capacity + 2 * cfg->ghost_capacity >= (1ull << 31))
        return pst_fail(nullptr, PST_EINVAL, "capacity must be in [1, 2^31)");
    if ((cfg->physics & PST_PHYS_DEM) && cfg->dim != 3) return pst_fail(nullptr, PST_EINVAL, "DEM needs dim = 3");
    if (cfg->max_contacts < 0 || cfg->max_contacts > 32) return pst_fail(nullptr, PST_EINVAL, "max_contacts must be in [0, 32]");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return pst_fail(nullptr, PST_ECUDA, "no CUDA device: %s (this library has no CPU path)", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return pst_fail(nullptr, PST_EINVAL, "device %d out of range", cfg->device);
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) return pst_fail(nullptr, PST_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));

    pst_ctx* ctx = new pst_ctx();
    ctx->cfg = *cfg;
    ctx->f64 = cfg->real == PST_F64;
    ctx->dim = cfg->dim;
    ctx->coupled = (cfg->physics & PST_PHYS_WCSPH) && (cfg->physics & PST_PHYS_DEM);
    ctx->capacity = cfg->capacity;
    ctx->ghost_cap = cfg->ghost_capacity;
    auto bail = [&](pst_status s) { g_create_err = ctx->err; pst_destroy(ctx); return s; };
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking) != cudaSuccess) {
        pst_fail(ctx, PST_ECUDA, "cudaStreamCreate failed");
        return bail(PST_ECUDA);
    }
    // default parameters
    ctx->params = {{"rho0", 1000.0}, {"c0", 10.0}, {"gamma", 7.0}, {"alpha", 0.1}, {"beta", 0.0}, {"kfac", 2.0},
                   {"gx", 0.0}, {"gy", 0.0}, {"gz", 0.0}, {"dem_model", 0.0}, {"kn", 1e5}, {"gn", 0.0}, {"kt", 2e4},
                   {"gt", 0.0}, {"mu", 0.5}, {"dt", 1e-6}, {"Estar", 1e7}, {"Gstar", 4e6}, {"erest", 0.8}, {"rho_solid", 2500.0},
                   {"boundary_model", 0.0}};   // 1: dummy-particle wall pressure in pst_step, wall density slaved
    // fast-axis subdivision of the cell grid (option "zsub"): the tiled WCSPH pair kernel (force_kernel 3, the default) scans
    // z-trimmed runs and wants fine cells; DEM-only contexts keep whole cells (a cell holds about one sphere)
    ctx->grid.sub = ((cfg->physics & PST_PHYS_WCSPH) && cfg->key == PST_KEY_LINEAR) ? 4 : 1;
    pst_status s = pst_grid_finalize(ctx);   // grid + cell table
    if (s != PST_OK) return bail(s);
    auto mk = [&](const char* name, int dt, uint32_t fl, int rows = 1) { if (s == PST_OK) s = array_create(ctx, name, dt, fl, rows); };
    const uint32_t P = PST_ARRAY_PERSISTENT, O = PST_ARRAY_OUTPUT;
    mk("id", PST_U32, P);
    if (cfg->physics & (PST_PHYS_WCSPH | PST_PHYS_DEM)) {
        mk("x", PST_REAL, P); mk("y", PST_REAL, P); if (cfg->dim == 3) mk("z", PST_REAL, P);
        mk("u", PST_REAL, P); mk("v", PST_REAL, P); if (cfg->dim == 3) mk("w", PST_REAL, P);
        mk("m", PST_REAL, P); mk("tag", PST_I32, P);
    }
    if (cfg->physics & PST_PHYS_WCSPH) {
        mk("rho", PST_REAL, P); mk("h", PST_REAL, P);
        mk("p", PST_REAL, O); mk("por2", PST_REAL, O);
        mk("au", PST_REAL, O); mk("av", PST_REAL, O); if (cfg->dim == 3) mk("aw", PST_REAL, O);
        mk("arho", PST_REAL, O);
        if (ctx->coupled) mk("msph", PST_REAL, O);   // signed SPH mass, written by the EOS pass
    }
    if (cfg->physics & PST_PHYS_DEM) {
        mk("wx", PST_REAL, P); mk("wy", PST_REAL, P); mk("wz", PST_REAL, P);
        mk("rad", PST_REAL, P); mk("inertia", PST_REAL, P);
        mk("fx", PST_REAL, O); mk("fy", PST_REAL, O); mk("fz", PST_REAL, O);
        mk("tx", PST_REAL, O); mk("ty", PST_REAL, O); mk("tz", PST_REAL, O);
        if (cfg->max_contacts > 0) {
            const int K = cfg->max_contacts;
            mk("hist_n", PST_I32, P);
            mk("hist_id", PST_U32, P, K);
            mk("hist_x", PST_REAL, P, K); mk("hist_y", PST_REAL, P, K); mk("hist_z", PST_REAL, P, K);
        }
    }
    if (s == PST_OK) s = pst_nnps_alloc(ctx);
    if (s != PST_OK) return bail(s);
    *out = ctx;
    return PST_OK;
}

void pst_destroy(pst_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    pst_comm_destroy(ctx);
    for (auto& a : ctx->arrays) { cudaFree(a.buf[0]); if (a.buf[1]) cudaFree(a.buf[1]); }
    cudaFree(ctx->keys_in); cudaFree(ctx->keys_out); cudaFree(ctx->vals_in); cudaFree(ctx->vals_out);
    cudaFree(ctx->cell_start); cudaFree(ctx->sort_tmp); cudaFree(ctx->scan_sums); cudaFree(ctx->nf_pos); cudaFree(ctx->nf_idx); cudaFree(ctx->nf_rec); cudaFree(ctx->wp_pos); cudaFree(ctx->wp_idx); cudaFree(ctx->rec); cudaFree(ctx->posf); cudaFree(ctx->ztiles); cudaFree(ctx->d_ztile_count); cudaFree(ctx->d_uni); cudaFree(ctx->stage); cudaFree(ctx->d_flags); cudaFree(ctx->d_bodies); cudaFree(ctx->d_body_start); cudaFree(ctx->d_body_vals);
    cudaFree(ctx->d_counters);
    if (ctx->ev_stats) cudaEventDestroy(ctx->ev_stats);
    if (ctx->step_graph) cudaGraphExecDestroy(ctx->step_graph);
    if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->h_uni) cudaFreeHost(ctx->h_uni);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->h2d_stream) { cudaStreamSynchronize(ctx->h2d_stream); cudaStreamDestroy(ctx->h2d_stream); }
    if (ctx->d2h_stream) { cudaStreamSynchronize(ctx->d2h_stream); cudaStreamDestroy(ctx->d2h_stream); }
    for (int k = 0; k < pst_ctx::kRing; ++k) {
        cudaFree(ctx->ring[k]);
        if (ctx->ring_free[k]) cudaEventDestroy(ctx->ring_free[k]);
        if (ctx->ring_ready[k]) cudaEventDestroy(ctx->ring_ready[k]);
    }
    delete ctx;
}

void* pst_stream(pst_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// Fetch device flags; turn a recorded contact overflow into PST_EOVERFLOW (never silent truncation).
static pst_status check_flags(pst_ctx* ctx) {
    PST_CUDA(ctx, cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_flags[0]) {
        const int worst = ctx->h_flags[1];
        PST_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, 8 * sizeof(int32_t), ctx->stream));
        return pst_fail(ctx, PST_EOVERFLOW, "a particle has %d contacts > max_contacts = %d", worst, ctx->cfg.max_contacts);
    }
    if (ctx->h_flags[4]) {
        PST_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, 8 * sizeof(int32_t), ctx->stream));
        ctx->nbrs_valid = false;        // the ghost rows and the ghost cell table are stale: nothing may be evaluated on them
        ctx->eos_valid = false;
        return pst_fail(ctx, PST_ENCCL, "peer-memory halo: a slab neighbour did not publish its edge layers in time");
    }
    if (ctx->h_flags[2]) {
        const int need = ctx->h_flags[3];
        PST_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, 8 * sizeof(int32_t), ctx->stream));
        return pst_fail(ctx, PST_ENOMEM, "a ghost layer of %d particles exceeds ghost_capacity %llu", need, (unsigned long long)ctx->ghost_cap);
    }
    return PST_OK;
}

pst_status pst_sync(pst_ctx* ctx) {
    if (!ctx) return PST_EINVAL;
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    PST_CUDA(ctx, cudaStreamSynchronize(ctx->h2d_stream));
    PST_TRY(check_flags(ctx));           // synchronises the compute stream
    PST_CUDA(ctx, cudaStreamSynchronize(ctx->d2h_stream));
    return PST_OK;
}

pst_status pst_wait_transfers(pst_ctx* ctx) {
    if (!ctx) return PST_EINVAL;
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    PST_CUDA(ctx, cudaStreamSynchronize(ctx->d2h_stream));
    return PST_OK;
}

// next staging buffer of the ring; the stream `user` is made to wait until the buffer's previous use is over
static pst_status ring_acquire(pst_ctx* ctx, cudaStream_t user, bool upload, int* idx) {
    // separate sub-rings, so an upload never waits for a download's D2H (and vice versa)
    constexpr int kUp = 24, kDown = pst_ctx::kRing - kUp;
    int k;
    if (upload) { k = ctx->ring_next_up; ctx->ring_next_up = (k + 1) % kUp; }
    else { k = kUp + ctx->ring_next_down; ctx->ring_next_down = (ctx->ring_next_down + 1) % kDown; }
    if (!ctx->ring[0]) {   // first asynchronous transfer: allocate the whole ring once (no cudaMalloc on the hot path later)
        const size_t bytes = (ctx->capacity + 2 * ctx->ghost_cap) * 8;
        for (int r = 0; r < pst_ctx::kRing; ++r) {
            if (cudaMalloc((void**)&ctx->ring[r], bytes) != cudaSuccess) return pst_fail(ctx, PST_ENOMEM, "staging ring (%d x %zu bytes)", pst_ctx::kRing, bytes);
            PST_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ring_free[r], cudaEventDisableTiming));
            PST_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ring_ready[r], cudaEventDisableTiming));
            PST_CUDA(ctx, cudaEventRecord(ctx->ring_free[r], ctx->stream));
        }
    }
    PST_CUDA(ctx, cudaStreamWaitEvent(user, ctx->ring_free[k], 0));
    *idx = k;
    return PST_OK;
}

// Asynchronous transfers: the copy engines run on their own streams (H2D and D2H are full duplex) and overlap the
// kernels of the compute stream.  `host` must be pinned (pst_host_alloc) and must not be touched until pst_sync
// (uploads) / pst_wait_transfers or pst_sync (downloads).
pst_status pst_upload_async(pst_ctx* ctx, const char* name, const void* host, size_t n) {
    if (!ctx || !name || !host) return PST_EINVAL;
    PstArray* a = pst_find(ctx, name);
    if (!a) return pst_fail(ctx, PST_EINVAL, "unknown array '%s'", name);
    if (n != ctx->n) return pst_fail(ctx, PST_EINVAL, "upload '%s': n = %zu but the context holds %llu particles", name, n, (unsigned long long)ctx->n);
    if (a->name == "id") return pst_fail(ctx, PST_EINVAL, "'id' is maintained by the library");
    if ((!ctx->ordered && !ctx->comm) || a->rows != 1) return pst_upload(ctx, name, host, n);   // identity order / history rows: plain path
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    note_uniform_dirty(ctx, a);
    int k = 0;
    PST_TRY(ring_acquire(ctx, ctx->h2d_stream, true, &k));
    PST_CUDA(ctx, cudaMemcpyAsync(ctx->ring[k], host, n * a->esize, cudaMemcpyHostToDevice, ctx->h2d_stream));
    PST_CUDA(ctx, cudaEventRecord(ctx->ring_ready[k], ctx->h2d_stream));
    PST_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ring_ready[k], 0));
    if (ctx->comm) {
        // distributed mode: host arrays are in DEVICE order, so the staged copy goes straight into the array -- on the compute
        // stream, after whatever still reads the array there (a copy engine writing into it directly would race with the step)
        PST_CUDA(ctx, cudaMemcpyAsync(pst_ptr<char>(ctx, a), ctx->ring[k], n * a->esize, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        char* saved = ctx->stage;
        ctx->stage = ctx->ring[k];
        const pst_status st = pst_reorder_upload(ctx, a, 0, n);
        ctx->stage = saved;
        PST_TRY(st);
    }
    PST_CUDA(ctx, cudaEventRecord(ctx->ring_free[k], ctx->stream));
    if (a->name == "x" || a->name == "y" || a->name == "z" || a->name == "rad" || a->name == "h") ctx->nbrs_valid = false;
    if (a->name == "rho" || a->name == "m" || a->name == "tag") ctx->eos_valid = false;
    ctx->state_epoch++;
    return PST_OK;
}

pst_status pst_download_async(pst_ctx* ctx, const char* name, void* host, size_t n) {
    if (!ctx || !name || !host) return PST_EINVAL;
    PstArray* a = pst_find(ctx, name);
    if (!a) return pst_fail(ctx, PST_EINVAL, "unknown array '%s'", name);
    if (n != ctx->n) return pst_fail(ctx, PST_EINVAL, "download '%s': n = %zu but the context holds %llu particles", name, n, (unsigned long long)ctx->n);
    if ((!ctx->ordered && !ctx->comm) || a->rows != 1 || (a->name == "id" && !ctx->comm)) return pst_download(ctx, name, host, n);
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    int k = 0;
    PST_TRY(ring_acquire(ctx, ctx->stream, false, &k));
    if (ctx->comm) {       // distributed mode: device order, a snapshot of the array taken on the compute stream
        PST_CUDA(ctx, cudaMemcpyAsync(ctx->ring[k], pst_ptr<char>(ctx, a), n * a->esize, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        char* saved = ctx->stage;
        ctx->stage = ctx->ring[k];
        const pst_status st = pst_reorder_download(ctx, a, 0, n);
        ctx->stage = saved;
        PST_TRY(st);
    }
    PST_CUDA(ctx, cudaEventRecord(ctx->ring_ready[k], ctx->stream));
    PST_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h_stream, ctx->ring_ready[k], 0));
    PST_CUDA(ctx, cudaMemcpyAsync(host, ctx->ring[k], n * a->esize, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    PST_CUDA(ctx, cudaEventRecord(ctx->ring_free[k], ctx->d2h_stream));
    return PST_OK;
}

pst_status pst_set_param(pst_ctx* ctx, const char* name, double value) {
    if (!ctx || !name) return PST_EINVAL;
    if (!ctx->params.count(name)) return pst_fail(ctx, PST_EINVAL, "unknown parameter '%s'", name);
    ctx->params[name] = value;
    ctx->eos_valid = false;
    ctx->state_epoch++;
    return PST_OK;
}
pst_status pst_get_param(pst_ctx* ctx, const char* name, double* value) {
    if (!ctx || !name || !value) return PST_EINVAL;
    auto it = ctx->params.find(name);
    if (it == ctx->params.end()) return pst_fail(ctx, PST_EINVAL, "unknown parameter '%s'", name);
    *value = it->second;
    return PST_OK;
}
pst_status pst_set_option(pst_ctx* ctx, const char* name, int value) {
    if (!ctx || !name) return PST_EINVAL;
    // every option the kernels read, with its valid range: a typo or an out-of-range value is an error, never silently ignored
    static const struct { const char* name; int lo, hi; } known[] = {
        {"force_kernel", 0, 3},          // WCSPH pair kernel: 0 gather | 1 warp per cell | 2 tiled lists | 3 tiled z-runs + bit masks
        {"dem_kernel", 0, 2},
        {"sort_impl", 0, 1},             // 0 library radix sort + bounds kernel (A/B) | 1 counting sort (default)
        {"halo_impl", 0, 2},             // 0 per-array NCCL | 1 packed NCCL | 2 peer memory (default; falls back to 1 where the GPUs cannot map each other)
        {"halo_timeout_ms", 0, 86400000},// peer-memory halo: how long a rank waits for its neighbour's step before PST_ENCCL (0 = for ever; default 10 min)
        {"uniform_mass", 0, 1},
        {"uniform_mass_global", 0, 1},
        {"tile_g", 0, 64},               // 0 = from the measured cell occupancy
        {"tile_lcap", 4, 256},
        {"tile_jcap", 0, 1 << 20},
        {"tile_ta", 2, 3},
        {"zsub", 1, 8},                  // fast-axis subdivision of the cell grid: 1, 2, 4 or 8
        {"tile_words", 4, 64},
        {"rec_impl", 0, 1},
        {"fuse_eos", 0, 1},              // 1 (default): single-GPU WCSPH contexts evaluate the EOS and write the packed records inside the state permute
        {"tile_dbg", 0, 9},
        {"tile_gf", 0, 256},
        {"tile_jc", 0, 1},              // timing ablations of the variant-3 kernel (WRONG results): 1 no gathers, 2 no pair body, 3 scan only
        {"graph", 0, 1},
    };
    const std::string nm = name;
    for (const auto& k : known) {
        if (nm != k.name) continue;
        if (value < k.lo || value > k.hi) return pst_fail(ctx, PST_EINVAL, "option '%s' = %d is outside [%d, %d]", name, value, k.lo, k.hi);
        if (nm == "zsub") {
            if (value & (value - 1)) return pst_fail(ctx, PST_EINVAL, "option 'zsub' must be 1, 2, 4 or 8");
            if (ctx->comm) return pst_fail(ctx, PST_ESTATE, "set 'zsub' before pst_comm_init (the halo windows are sized by the cell layer)");
            if (value != ctx->grid.sub) {
                PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
                const int old = ctx->grid.sub;
                ctx->grid.sub = value;
                const pst_status st = pst_grid_finalize(ctx);
                if (st != PST_OK) { ctx->grid.sub = old; pst_grid_finalize(ctx); return st; }
            }
        }
        ctx->options[nm] = value;
        return PST_OK;
    }
    return pst_fail(ctx, PST_EINVAL, "unknown option '%s'", name);
}

pst_status pst_set_count(pst_ctx* ctx, uint64_t n) {
    if (!ctx) return PST_EINVAL;
    if (n > ctx->capacity) return pst_fail(ctx, PST_ENOMEM, "count %llu > capacity %llu", (unsigned long long)n, (unsigned long long)ctx->capacity);
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    ctx->n = n;
    ctx->n_ghost_l = ctx->n_ghost_r = 0;
    ctx->ghost_exact = true;
    ctx->ordered = false;
    ctx->nbrs_valid = false;
    ctx->eos_valid = false;
    ctx->hist_lag = false;
    ctx->state_epoch++;
    ctx->bodies_ready = false;       // a new particle set: pst_bodies_setup again
    ctx->m_uniform = ctx->h_uniform = false;   // ... and its masses and smoothing lengths are not known yet
    ctx->uni_dirty = true;
    PST_TRY(pst_iota_ids(ctx));
    if (PstArray* hn = pst_find(ctx, "hist_n")) {
        const size_t stride = ctx->capacity + 2 * ctx->ghost_cap;
        for (int b = 0; b < 2; ++b) PST_CUDA(ctx, cudaMemsetAsync(hn->buf[b], 0, stride * hn->esize, ctx->stream));
    }
    return PST_OK;
}
pst_status pst_get_count(pst_ctx* ctx, uint64_t* n_owned, uint64_t* n_ghost) {
    if (!ctx) return PST_EINVAL;
    if (n_owned) *n_owned = ctx->n;
    if (n_ghost) {
        int64_t nl = 0, nr = 0;
        PST_TRY(pst_ghost_counts(ctx, &nl, &nr));
        *n_ghost = (uint64_t)(nl + nr);
    }
    return PST_OK;
}

pst_status pst_array_create(pst_ctx* ctx, const char* name, int dtype, uint32_t flags) {
    if (!ctx) return PST_EINVAL;
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    return array_create(ctx, name, dtype, flags, 1);
}

pst_status pst_array(pst_ctx* ctx, const char* name, void** dev_ptr, size_t* n, int* dtype, int* rows) {
    if (!ctx || !name) return PST_EINVAL;
    PstArray* a = pst_find(ctx, name);
    if (!a) return pst_fail(ctx, PST_EINVAL, "unknown array '%s'", name);
    if (dev_ptr) { *dev_ptr = pst_ptr<char>(ctx, a); ctx->state_epoch++; note_uniform_dirty(ctx, a); }   // the caller may write through it
    if (n) *n = ctx->n;
    if (dtype) *dtype = a->dtype;
    if (rows) *rows = a->rows;
    return PST_OK;
}

pst_status pst_upload(pst_ctx* ctx, const char* name, const void* host, size_t n) {
    if (!ctx || !name || !host) return PST_EINVAL;
    PstArray* a = pst_find(ctx, name);
    if (!a) return pst_fail(ctx, PST_EINVAL, "unknown array '%s'", name);
    if (n != ctx->n) return pst_fail(ctx, PST_EINVAL, "upload '%s': n = %zu but the context holds %llu particles (pst_set_count first)", name, n, (unsigned long long)ctx->n);
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    note_uniform_dirty(ctx, a);
    if (a->name.rfind("hist_", 0) == 0) PST_TRY(pst_resolve_history(ctx));
    if (a->name == "id") {
        // Restoring a checkpoint in DEVICE order: right after pst_set_count (identity order) the caller may declare
        // which stable id sits in which slot.  Must be a permutation of 0..n-1; from here on host arrays are id-ordered.
        if (ctx->comm) {   // distributed mode: ids are global labels, transfers are in device order
            PST_CUDA(ctx, cudaMemcpyAsync(pst_ptr<char>(ctx, a), host, n * 4, cudaMemcpyHostToDevice, ctx->stream));
            PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            return PST_OK;
        }
        if (ctx->ordered) return pst_fail(ctx, PST_ESTATE, "'id' can only be set right after pst_set_count");
        const uint32_t* ids = (const uint32_t*)host;
        std::vector<bool> seen(n, false);
        for (size_t k = 0; k < n; ++k) {
            if (ids[k] >= n || seen[ids[k]]) return pst_fail(ctx, PST_EINVAL, "'id' must be a permutation of 0..n-1");
            seen[ids[k]] = true;
        }
        PST_CUDA(ctx, cudaMemcpyAsync(pst_ptr<char>(ctx, a), host, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->ordered = true;
        ctx->nbrs_valid = false;
        return PST_OK;
    }
    for (int r = 0; r < a->rows; ++r) {
        const char* src = (const char*)host + (size_t)r * n * a->esize;
        if (!ctx->ordered || ctx->comm) {   // identity order, or distributed mode (host arrays in device order)
            PST_CUDA(ctx, cudaMemcpyAsync(pst_ptr<char>(ctx, a, r), src, n * a->esize, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            PST_CUDA(ctx, cudaMemcpyAsync(ctx->stage, src, n * a->esize, cudaMemcpyHostToDevice, ctx->stream));
            PST_TRY(pst_reorder_upload(ctx, a, r, n));
        }
    }
    if (a->name == "x" || a->name == "y" || a->name == "z" || a->name == "rad" || a->name == "h") ctx->nbrs_valid = false;
    if (a->name == "rho" || a->name == "m" || a->name == "tag") ctx->eos_valid = false;
    ctx->state_epoch++;
    return PST_OK;
}

pst_status pst_download(pst_ctx* ctx, const char* name, void* host, size_t n) {
    if (!ctx || !name || !host) return PST_EINVAL;
    PstArray* a = pst_find(ctx, name);
    if (!a) return pst_fail(ctx, PST_EINVAL, "unknown array '%s'", name);
    if (n != ctx->n) return pst_fail(ctx, PST_EINVAL, "download '%s': n = %zu but the context holds %llu particles", name, n, (unsigned long long)ctx->n);
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    if (a->name.rfind("hist_", 0) == 0) PST_TRY(pst_resolve_history(ctx));
    for (int r = 0; r < a->rows; ++r) {
        char* dst = (char*)host + (size_t)r * n * a->esize;
        if (!ctx->ordered || ctx->comm || a->name == "id") {
            PST_CUDA(ctx, cudaMemcpyAsync(dst, pst_ptr<char>(ctx, a, r), n * a->esize, cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            PST_TRY(pst_reorder_download(ctx, a, r, n));
            PST_CUDA(ctx, cudaMemcpyAsync(dst, ctx->stage, n * a->esize, cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    return check_flags(ctx);
}

void* pst_host_alloc(size_t bytes) {
    void* p = nullptr;
    return cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess ? p : nullptr;
}
void pst_host_free(void* p) { if (p) cudaFreeHost(p); }

pst_status pst_build_neighbours(pst_ctx* ctx) {
    PstRange range("pst_build_neighbours");
    if (!ctx) return PST_EINVAL;
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    return pst_nnps_build(ctx);
}

pst_status pst_apply(pst_ctx* ctx, const char* const* eq_names, int n_eq) {
    PstRange range("pst_apply");
    if (!ctx || !eq_names || n_eq <= 0) return PST_EINVAL;
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    std::set<std::string> eqs;
    for (int k = 0; k < n_eq; ++k) {
        if (!eq_names[k]) return PST_EINVAL;
        eqs.insert(eq_names[k]);
    }
    static const std::set<std::string> known = {"eq1", "tait_eos", "wall_pressure", "continuity", "momentum", "dem_contact", "body_reduce"};
    for (auto& e : eqs)
        if (!known.count(e)) return pst_fail(ctx, PST_EINVAL, "no hand-written kernel for equation '%s'", e.c_str());
    // fuse(): group the set into the fused kernels that exist.  Bodies that share an i,j loop
    // (continuity + momentum) run in ONE pair kernel; tait_eos is per-particle and runs first
    // because momentum reads p[j].
    if (eqs.count("eq1")) {
        if (eqs.size() != 1) return pst_fail(ctx, PST_EINVAL, "eq1 cannot be fused with cutoff equations (it is an all-pairs loop)");
        return pst_eq1_apply(ctx);
    }
    const bool wc = eqs.count("tait_eos") || eqs.count("wall_pressure") || eqs.count("continuity") || eqs.count("momentum");
    if (wc && !(ctx->cfg.physics & PST_PHYS_WCSPH)) return pst_fail(ctx, PST_ESTATE, "context was created without PST_PHYS_WCSPH");
    if (eqs.count("dem_contact") && !(ctx->cfg.physics & PST_PHYS_DEM)) return pst_fail(ctx, PST_ESTATE, "context was created without PST_PHYS_DEM");
    if (eqs.count("tait_eos")) PST_TRY(pst_wcsph_eos(ctx));
    if (eqs.count("wall_pressure")) {   // per dummy particle, a gather over its fluid neighbours: after the EOS, before the pair kernel reads p[j]
        if (!ctx->nbrs_valid) return pst_fail(ctx, PST_ESTATE, "pst_build_neighbours must run before wall_pressure");
        if (!ctx->eos_valid || ctx->ghost_eos_pending) return pst_fail(ctx, PST_ESTATE, "wall_pressure reads p of the fluid: apply tait_eos first (or in the same set)");
        PST_TRY(pst_wcsph_wall_pressure(ctx));
    }
    if (eqs.count("continuity") || eqs.count("momentum")) {
        if (!ctx->nbrs_valid) return pst_fail(ctx, PST_ESTATE, "pst_build_neighbours must run before pair equations");
        if ((eqs.count("momentum") || ctx->coupled) && (!ctx->eos_valid || ctx->ghost_eos_pending)) return pst_fail(ctx, PST_ESTATE, "momentum reads p: apply tait_eos first (or in the same set)");
        if (ctx->coupled && !(eqs.count("continuity") && eqs.count("momentum")))
            return pst_fail(ctx, PST_EINVAL, "coupled contexts fuse continuity and momentum: apply both in one set");
        PST_TRY(pst_wcsph_forces(ctx, eqs.count("continuity") > 0, eqs.count("momentum") > 0));
    }
    if (eqs.count("dem_contact")) {
        if (!ctx->nbrs_valid) return pst_fail(ctx, PST_ESTATE, "pst_build_neighbours must run before pair equations");
        PST_TRY(pst_dem_forces(ctx));
    }
    if (eqs.count("body_reduce")) {   // per-particle results of the force loop -> force and torque of every rigid body (runs last)
        if (!ctx->d_bodies) return pst_fail(ctx, PST_ESTATE, "body_reduce: the context has no rigid bodies (pst_bodies_create)");
        PST_TRY(pst_rb_reduce(ctx));
    }
    return PST_OK;
}

pst_status pst_dump_pairs(pst_ctx* ctx, int mode, uint32_t* i, uint32_t* j, size_t cap, size_t* n_pairs) {
    if (!ctx || !n_pairs || (cap && (!i || !j))) return PST_EINVAL;
    if (!ctx->nbrs_valid) return pst_fail(ctx, PST_ESTATE, "pst_build_neighbours must run before pst_dump_pairs");
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    return pst_nnps_dump_pairs(ctx, mode, i, j, cap, n_pairs);
}

pst_status pst_integrate(pst_ctx* ctx, double dt) {
    PstRange range("pst_integrate");
    if (!ctx) return PST_EINVAL;
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    if (ctx->coupled) {
        // rigid bodies: reduce the current per-particle forces, advance free particles (members are moved too, as if
        // free, and then overwritten), advance the bodies and scatter the rigid motion onto their members
        if (ctx->bodies_ready) PST_TRY(pst_rb_reduce(ctx));
        PST_TRY(pst_coupled_integrate(ctx, dt));
        if (ctx->bodies_ready) PST_TRY(pst_rb_integrate(ctx, dt));
    } else if (ctx->cfg.physics & PST_PHYS_WCSPH) PST_TRY(pst_wcsph_integrate(ctx, dt));
    else if (ctx->cfg.physics & PST_PHYS_DEM) PST_TRY(pst_dem_integrate(ctx, dt));
    ctx->nbrs_valid = false;
    ctx->eos_valid = false;
    ctx->state_epoch++;
    return PST_OK;
}

static pst_status step_once(pst_ctx* ctx, double dt) {
    PST_TRY(pst_nnps_build(ctx));
    if (ctx->comm) PST_TRY(pst_halo_exchange(ctx));
    if (ctx->cfg.physics & PST_PHYS_WCSPH) {
        PST_TRY(pst_wcsph_eos(ctx));
        if (pst_param(ctx, "boundary_model", 0.0) == 1.0) PST_TRY(pst_wcsph_wall_pressure(ctx));
        PST_TRY(pst_wcsph_forces(ctx, true, true));
    }
    if (ctx->cfg.physics & PST_PHYS_DEM) {
        ctx->params["dt"] = dt;
        PST_TRY(pst_dem_forces(ctx));
    }
    return pst_integrate(ctx, dt);
}

// everything the captured launch sequence depends on besides the device data: particle count, time step, every parameter and
// option, the set of arrays.  A change of any of them re-captures.
static uint64_t step_graph_key(pst_ctx* ctx, double dt) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; } };
    mix(&ctx->n, sizeof ctx->n); mix(&dt, sizeof dt);
    for (const auto& kv : ctx->params) { if (kv.first == "_ppc" || kv.first == "dt") continue; mix(kv.first.data(), kv.first.size()); mix(&kv.second, sizeof kv.second); }
    for (const auto& kv : ctx->options) { mix(kv.first.data(), kv.first.size()); mix(&kv.second, sizeof kv.second); }
    const size_t na = ctx->arrays.size();
    mix(&na, sizeof na);
    mix(&ctx->grid.sub, sizeof ctx->grid.sub);
    return h ? h : 1;
}

// which buffer of every double-buffered array is current: the captured pointers are only right for the parity they were
// captured at (every step flips them once; two steps restore them)
static uint64_t buffer_parity(pst_ctx* ctx) {
    uint64_t h = 1469598103934665603ull;
    for (const auto& a : ctx->arrays) { h ^= (uint64_t)(a.cur + 1); h *= 1099511628211ull; }
    return h;
}

pst_status pst_step(pst_ctx* ctx, double dt, int n_steps) {
    PstRange range("pst_step");
    if (!ctx || n_steps < 0) return PST_EINVAL;
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    int k = 0;
    // Launch-bound contexts (the 2D dam break: ~20 launches of a few microseconds each per step) run pst_step as a CUDA graph:
    // option graph = 1, single GPU, no rigid bodies.  Steps 1-2 of a sequence run eagerly (they allocate and tune), the next
    // two are captured, every further pair is one cudaGraphLaunch.
    const bool want_graph = pst_option(ctx, "graph", 0) == 1 && !ctx->comm && !ctx->d_bodies && ctx->n > 0 && n_steps >= 2;
    if (want_graph) {
        const uint64_t key = step_graph_key(ctx, dt);
        if (ctx->step_graph && (ctx->step_graph_key != key || ctx->uni_dirty)) { cudaGraphExecDestroy(ctx->step_graph); ctx->step_graph = nullptr; }
        if (ctx->step_graph && buffer_parity(ctx) != ctx->step_graph_parity) {     // an odd number of eager steps since the capture: one more realigns
            PST_TRY(step_once(ctx, dt));
            ++k;
            if (buffer_parity(ctx) != ctx->step_graph_parity) { cudaGraphExecDestroy(ctx->step_graph); ctx->step_graph = nullptr; }
        }
        if (!ctx->step_graph && n_steps - k >= 4) {
            for (int w = 0; w < 2; ++w, ++k) PST_TRY(step_once(ctx, dt));
            ctx->step_graph_parity = buffer_parity(ctx);
            const uint64_t l0 = ctx->launches;
            cudaGraph_t g = nullptr;
            PST_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
            ctx->capturing = true;
            pst_status st = step_once(ctx, dt);
            if (st == PST_OK) st = step_once(ctx, dt);
            ctx->capturing = false;
            const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
            if (st != PST_OK) { if (g) cudaGraphDestroy(g); return st; }
            PST_CUDA(ctx, ce);
            PST_CUDA(ctx, cudaGraphInstantiate(&ctx->step_graph, g, 0));
            cudaGraphDestroy(g);
            ctx->step_graph_key = key;
            ctx->step_graph_launches = ctx->launches - l0;
            ctx->launches = l0;                                   // nothing ran during the capture
            // (the host-side state now describes "two steps later" although the device has not run them: the first replay
            // below makes that true)
            PST_CUDA(ctx, cudaGraphLaunch(ctx->step_graph, ctx->stream));
            ctx->launches += ctx->step_graph_launches;
            k += 2;
        }
        if (ctx->step_graph)
            for (; k + 2 <= n_steps; k += 2) {
                PST_CUDA(ctx, cudaGraphLaunch(ctx->step_graph, ctx->stream));
                ctx->launches += ctx->step_graph_launches;
            }
    }
    for (; k < n_steps; ++k) PST_TRY(step_once(ctx, dt));
    return PST_OK;
}

pst_status pst_kernel_name(pst_ctx* ctx, const char* stage, char* buf, size_t cap) {
    if (!ctx || !stage || !buf || cap == 0) return PST_EINVAL;
    const std::string s = stage;
    const void* fn = s == "pair" ? ctx->pair_kernel_fn : s == "contact" ? ctx->contact_kernel_fn : nullptr;
    if (s != "pair" && s != "contact") return pst_fail(ctx, PST_EINVAL, "unknown stage '%s' (pair | contact)", stage);
    if (!fn) return pst_fail(ctx, PST_ESTATE, "no '%s' kernel has been launched yet", stage);
    const char* nm = nullptr;
    PST_CUDA(ctx, cudaFuncGetName(&nm, fn));
    std::snprintf(buf, cap, "%s", nm ? nm : "");
    return PST_OK;
}

pst_status pst_get_stat(pst_ctx* ctx, const char* name, double* value) {
    if (!ctx || !name || !value) return PST_EINVAL;
    const std::string s = name;
    if (s == "launches") { *value = (double)ctx->launches; return PST_OK; }
    if (s == "n_cells") { *value = (double)ctx->grid.ncells; return PST_OK; }
    if (s == "key_bits") { *value = (double)ctx->grid.key_bits; return PST_OK; }
    if (s == "nx") { *value = ctx->grid.n[0]; return PST_OK; }
    if (s == "ny") { *value = ctx->grid.n[1]; return PST_OK; }
    if (s == "nz") { *value = ctx->grid.n[2]; return PST_OK; }
    if (s == "n_ghost_l" || s == "n_ghost_r") {
        int64_t nl = 0, nr = 0;
        PST_TRY(pst_ghost_counts(ctx, &nl, &nr));
        *value = (double)(s == "n_ghost_l" ? nl : nr);
        return PST_OK;
    }
    if (s == "ordered") { *value = ctx->ordered; return PST_OK; }
    if (s == "particles_per_occupied_cell") { *value = pst_param(ctx, "_ppc", 0.0); return PST_OK; }
    if (s == "contacts_total") {  // sum of hist_n over owned particles
        PstArray* hn = pst_find(ctx, "hist_n");
        if (!hn) return pst_fail(ctx, PST_EINVAL, "no contact history in this context");
        std::vector<int32_t> h(ctx->n);
        PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
        PST_TRY(pst_resolve_history(ctx));
        PST_CUDA(ctx, cudaMemcpyAsync(h.data(), pst_ptr<int32_t>(ctx, hn), ctx->n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        double t = 0;
        for (auto v : h) t += v;
        *value = t;
        return PST_OK;
    }
    return pst_fail(ctx, PST_EINVAL, "unknown stat '%s'", name);
}

}  // extern "C"
