// halo.cu -- multi-GPU slab decomposition along x: ghost-particle exchange with the two slab neighbours
// by NCCL send/recv over NVLink.  One context = one rank = one GPU.
//
// Layout trick: with x as the SLOWEST key digit, a rank's outermost owned cell layer is one contiguous
// range of every sorted array, so there is no pack kernel: ncclSend reads straight out of the sorted
// arrays, ncclRecv writes straight into the ghost region.  Left ghosts land at indices [-nL, 0),
// right ghosts at [n, n + nR) of the same allocation, so the whole thing stays ONE monotone index
// space and the cell table keeps its prefix semantics (cell_start is signed for this reason).
// The neighbour's cell-table slice for the layer travels with the particles and is rebased on arrival.
//
// NCCL is loaded with dlopen at pst_comm_init, so single-GPU users need no NCCL at all and a host that
// already loaded a libnccl (e.g. torch's bundled one) shares it.
#include <dlfcn.h>

#include <cstring>

#include "pst_internal.h"

namespace {

// the slice of the NCCL ABI that is used (stable since NCCL 2.7)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0 };
typedef ncclResult_t (*fn_GetUniqueId)(ncclUniqueId*);
typedef ncclResult_t (*fn_CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
typedef ncclResult_t (*fn_CommDestroy)(ncclComm_t);
typedef ncclResult_t (*fn_Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef ncclResult_t (*fn_Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef ncclResult_t (*fn_Group)(void);
typedef const char* (*fn_GetErrorString)(ncclResult_t);

struct NcclApi {
    void* lib = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_Send Send = nullptr;
    fn_Recv Recv = nullptr;
    fn_Group GroupStart = nullptr, GroupEnd = nullptr;
    fn_GetErrorString GetErrorString = nullptr;
    std::string err;
};

NcclApi* nccl_api() {
    static NcclApi api;
    if (api.lib || !api.err.empty()) return &api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) { api.err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return &api; }
#define PST_SYM(field, sym)                                                   \
    api.field = (decltype(api.field))dlsym(api.lib, sym);                     \
    if (!api.field) { api.err = std::string("dlsym failed: ") + sym; api.lib = nullptr; return &api; }
    PST_SYM(GetUniqueId, "ncclGetUniqueId")
    PST_SYM(CommInitRank, "ncclCommInitRank")
    PST_SYM(CommDestroy, "ncclCommDestroy")
    PST_SYM(Send, "ncclSend")
    PST_SYM(Recv, "ncclRecv")
    PST_SYM(GroupStart, "ncclGroupStart")
    PST_SYM(GroupEnd, "ncclGroupEnd")
    PST_SYM(GetErrorString, "ncclGetErrorString")
#undef PST_SYM
    return &api;
}

}  // namespace

struct PstComm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    int32_t* d_counts = nullptr;   // [0..1] my send counts (left, right), [2..3] received counts
    int32_t* h_counts = nullptr;   // pinned: [0..3] as above, [4..7] boundary cell_start values
    int32_t* d_tab_l = nullptr;    // received cell-table slice for the left / right ghost layer
    int32_t* d_tab_r = nullptr;
};

#define PST_NCCL(ctx, expr)                                                                                  \
    do {                                                                                                     \
        ncclResult_t r__ = (expr);                                                                           \
        if (r__ != 0) return pst_fail(ctx, PST_ENCCL, "%s: %s", #expr, nccl_api()->GetErrorString(r__));     \
    } while (0)

namespace {

// ghost-layer cell table: received slice holds the SENDER's indices; rebase so that the layer's first
// particle lands on `base` in this rank's index space.
__global__ void k_rebase_table(int count, const int32_t* __restrict__ recv, int base, int32_t* __restrict__ cell_start) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < count) cell_start[t] = recv[t] - recv[0] + base;
}
std::vector<PstArray*> ghost_arrays(pst_ctx* ctx) {
    std::vector<PstArray*> v;
    auto add = [&](const char* nm) { if (PstArray* a = pst_find(ctx, nm)) v.push_back(a); };
    for (const char* nm : {"x", "y", "z", "u", "v", "w", "m"}) add(nm);
    if (ctx->cfg.physics & PST_PHYS_WCSPH) add("rho");
    if (ctx->cfg.physics & PST_PHYS_DEM) for (const char* nm : {"wx", "wy", "wz", "rad", "id"}) add(nm);
    if (ctx->coupled) add("tag");   // the signed SPH mass (k_eos) and the contact mask need it on ghosts too
    return v;
}

}  // namespace

extern "C" pst_status pst_comm_unique_id(void* id_bytes) {
    if (!id_bytes) return PST_EINVAL;
    NcclApi* api = nccl_api();
    if (!api->lib) return pst_fail(nullptr, PST_ENCCL, "%s", api->err.c_str());
    ncclUniqueId id;
    ncclResult_t r = api->GetUniqueId(&id);
    if (r != 0) return pst_fail(nullptr, PST_ENCCL, "ncclGetUniqueId: %s", api->GetErrorString(r));
    static_assert(sizeof(id) == PST_COMM_ID_BYTES, "ncclUniqueId size");
    std::memcpy(id_bytes, &id, sizeof id);
    return PST_OK;
}

// Attach a communicator.  The context's box [lo, hi) is this rank's slab; its x extent must be a whole
// number of cells so neighbouring slabs share cell boundaries.  The grid grows by one ghost layer on
// each x face (always, also at the outer faces, so every rank runs identical code).
extern "C" pst_status pst_comm_init(pst_ctx* ctx, const void* id_bytes, int rank, int n_ranks) {
    if (!ctx || !id_bytes || rank < 0 || rank >= n_ranks) return PST_EINVAL;
    if (ctx->comm) return pst_fail(ctx, PST_ESTATE, "communicator already attached");
    if (ctx->grid.morton) return pst_fail(ctx, PST_EINVAL, "slab decomposition needs linear keys (x slowest)");
    if (ctx->ghost_cap == 0) return pst_fail(ctx, PST_EINVAL, "pst_config.ghost_capacity is 0");
    NcclApi* api = nccl_api();
    if (!api->lib) return pst_fail(ctx, PST_ENCCL, "%s", api->err.c_str());
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    PstGrid& g = ctx->grid;
    const double ext = ctx->cfg.hi[0] - ctx->cfg.lo[0];
    const double cells = ext / g.cell;
    if (std::fabs(cells - std::round(cells)) > 1e-6 * std::max(1.0, cells))
        return pst_fail(ctx, PST_EINVAL, "slab x extent %.17g is not a whole number of cells (%.17g)", ext, g.cell);
    g.n[0] = (int)std::llround(cells) + 2;
    g.lo[0] = ctx->cfg.lo[0] - g.cell;
    g.cx_lo = 1;
    g.cx_hi = g.n[0] - 2;
    const double total = (double)g.n[0] * g.n[1] * g.n[2];
    if (total >= 2147483647.0) return pst_fail(ctx, PST_EINVAL, "grid has too many cells");
    g.ncells = (uint32_t)total;
    g.key_bits = 1;
    while ((1ull << g.key_bits) < g.ncells) ++g.key_bits;
    PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PST_CUDA(ctx, cudaFree(ctx->cell_start));
    ctx->cell_start = nullptr;
    PST_CUDA(ctx, cudaMalloc((void**)&ctx->cell_start, ((size_t)g.ncells + 4) * 4));
    PST_CUDA(ctx, cudaMemsetAsync(ctx->cell_start, 0, ((size_t)g.ncells + 4) * 4, ctx->stream));
    PST_CUDA(ctx, cudaFree(ctx->scan_sums));
    ctx->scan_sums = nullptr;
    ctx->scan_sums_cap = ((size_t)g.ncells + 4) / 4096 + 2;
    PST_CUDA(ctx, cudaMalloc((void**)&ctx->scan_sums, ctx->scan_sums_cap * 4));
    PstComm* c = new PstComm();
    c->rank = rank; c->nranks = n_ranks;
    const size_t layer = (size_t)g.n[1] * g.n[2] + 1;
    if (cudaMalloc((void**)&c->d_counts, 8 * 4) != cudaSuccess || cudaMalloc((void**)&c->d_tab_l, layer * 4) != cudaSuccess ||
        cudaMalloc((void**)&c->d_tab_r, layer * 4) != cudaSuccess || cudaHostAlloc((void**)&c->h_counts, 8 * 4, cudaHostAllocDefault) != cudaSuccess) {
        delete c;
        return pst_fail(ctx, PST_ENOMEM, "halo buffers");
    }
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, sizeof id);
    ncclResult_t r = api->CommInitRank(&c->comm, n_ranks, id, rank);
    if (r != 0) { delete c; return pst_fail(ctx, PST_ENCCL, "ncclCommInitRank: %s", api->GetErrorString(r)); }
    ctx->comm = c;
    ctx->nbrs_valid = false;
    return PST_OK;
}

pst_status pst_comm_destroy(pst_ctx* ctx) {
    if (!ctx->comm) return PST_OK;
    PstComm* c = ctx->comm;
    if (c->comm) nccl_api()->CommDestroy(c->comm);
    cudaFree(c->d_counts); cudaFree(c->d_tab_l); cudaFree(c->d_tab_r);
    if (c->h_counts) cudaFreeHost(c->h_counts);
    delete c;
    ctx->comm = nullptr;
    return PST_OK;
}

void pst_comm_neighbours(pst_ctx* ctx, int* has_left, int* has_right) {
    *has_left = ctx->comm && ctx->comm->rank > 0;
    *has_right = ctx->comm && ctx->comm->rank + 1 < ctx->comm->nranks;
}

// Particle migration.  build_pass(mig) has sorted the leavers behind the stayers: [n_stay, n_stay + nL) go left,
// [n_stay + nL, n) go right, as contiguous ranges of EVERY persistent array (state, id, tag, contact-history rows), so
// they are sent as they are.  Arrivals are received into the free halves of the double buffers and appended after the
// stayers; the caller re-sorts if anything arrived.  History rows travel with their particle (partner ids must then
// be globally unique: upload a global "id" array in distributed runs).
pst_status pst_migrate(pst_ctx* ctx, int* arrivals) {
    *arrivals = 0;
    PstComm* c = ctx->comm;
    NcclApi* api = nccl_api();
    const int left = c->rank > 0 ? c->rank - 1 : -1, right = c->rank + 1 < c->nranks ? c->rank + 1 : -1;
    const size_t nc = ctx->grid.ncells;
    PST_CUDA(ctx, cudaMemcpyAsync(c->h_counts + 4, ctx->cell_start + nc, 3 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int n_stay = c->h_counts[4], nL = c->h_counts[5] - c->h_counts[4], nR = c->h_counts[6] - c->h_counts[5];
    c->h_counts[0] = nL; c->h_counts[1] = nR; c->h_counts[2] = c->h_counts[3] = 0;
    PST_CUDA(ctx, cudaMemcpyAsync(c->d_counts, c->h_counts, 4 * 4, cudaMemcpyHostToDevice, ctx->stream));
    PST_NCCL(ctx, api->GroupStart());
    if (left >= 0) { PST_NCCL(ctx, api->Send(c->d_counts + 0, 4, ncclInt8, left, c->comm, ctx->stream)); PST_NCCL(ctx, api->Recv(c->d_counts + 2, 4, ncclInt8, left, c->comm, ctx->stream)); }
    if (right >= 0) { PST_NCCL(ctx, api->Send(c->d_counts + 1, 4, ncclInt8, right, c->comm, ctx->stream)); PST_NCCL(ctx, api->Recv(c->d_counts + 3, 4, ncclInt8, right, c->comm, ctx->stream)); }
    PST_NCCL(ctx, api->GroupEnd());
    PST_CUDA(ctx, cudaMemcpyAsync(c->h_counts + 2, c->d_counts + 2, 2 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int aL = left >= 0 ? c->h_counts[2] : 0, aR = right >= 0 ? c->h_counts[3] : 0;
    if ((uint64_t)n_stay + aL + aR > ctx->capacity)
        return pst_fail(ctx, PST_ENOMEM, "migration: %d stayers + %d arrivals exceed capacity %llu", n_stay, aL + aR, (unsigned long long)ctx->capacity);
    if (nL + nR + aL + aR > 0) {
        PST_NCCL(ctx, api->GroupStart());
        for (auto& a : ctx->arrays) {
            if (!(a.flags & PST_ARRAY_PERSISTENT)) continue;
            const size_t es = a.esize;
            for (int r = 0; r < a.rows; ++r) {
                char* cur = pst_ptr<char>(ctx, &a, r, a.cur);
                char* alt = pst_ptr<char>(ctx, &a, r, 1 - a.cur);
                if (left >= 0) {
                    if (nL > 0) PST_NCCL(ctx, api->Send(cur + (size_t)n_stay * es, (size_t)nL * es, ncclInt8, left, c->comm, ctx->stream));
                    if (aL > 0) PST_NCCL(ctx, api->Recv(alt, (size_t)aL * es, ncclInt8, left, c->comm, ctx->stream));
                }
                if (right >= 0) {
                    if (nR > 0) PST_NCCL(ctx, api->Send(cur + (size_t)(n_stay + nL) * es, (size_t)nR * es, ncclInt8, right, c->comm, ctx->stream));
                    if (aR > 0) PST_NCCL(ctx, api->Recv(alt + (size_t)aL * es, (size_t)aR * es, ncclInt8, right, c->comm, ctx->stream));
                }
            }
        }
        PST_NCCL(ctx, api->GroupEnd());
        if (aL + aR > 0)
            for (auto& a : ctx->arrays) {
                if (!(a.flags & PST_ARRAY_PERSISTENT)) continue;
                for (int r = 0; r < a.rows; ++r)
                    PST_CUDA(ctx, cudaMemcpyAsync(pst_ptr<char>(ctx, &a, r, a.cur) + (size_t)n_stay * a.esize, pst_ptr<char>(ctx, &a, r, 1 - a.cur),
                                                  (size_t)(aL + aR) * a.esize, cudaMemcpyDeviceToDevice, ctx->stream));
            }
    }
    ctx->n = (uint64_t)(n_stay + aL + aR);
    *arrivals = aL + aR;
    return PST_OK;
}

extern "C" pst_status pst_halo_exchange(pst_ctx* ctx) {
    if (!ctx) return PST_EINVAL;
    if (!ctx->comm) return pst_fail(ctx, PST_ESTATE, "no communicator attached (pst_comm_init)");
    if (!ctx->nbrs_valid) return pst_fail(ctx, PST_ESTATE, "pst_build_neighbours must run before pst_halo_exchange");
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    PstComm* c = ctx->comm;
    NcclApi* api = nccl_api();
    const PstGrid& g = ctx->grid;
    const int layer = g.n[1] * g.n[2];
    const int left = c->rank > 0 ? c->rank - 1 : -1, right = c->rank + 1 < c->nranks ? c->rank + 1 : -1;
    const size_t kL0 = (size_t)g.cx_lo * layer, kL1 = kL0 + layer;           // my first owned layer
    const size_t kR0 = (size_t)g.cx_hi * layer, kR1 = kR0 + layer;           // my last owned layer
    // 1. where do my edge layers start and end?  (4 table entries -> host)
    PST_CUDA(ctx, cudaMemcpyAsync(c->h_counts + 4, ctx->cell_start + kL0, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaMemcpyAsync(c->h_counts + 5, ctx->cell_start + kL1, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaMemcpyAsync(c->h_counts + 6, ctx->cell_start + kR0, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaMemcpyAsync(c->h_counts + 7, ctx->cell_start + kR1, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int sL0 = c->h_counts[4], sL1 = c->h_counts[5], sR0 = c->h_counts[6], sR1 = c->h_counts[7];
    c->h_counts[0] = sL1 - sL0;   // particles I send to the left
    c->h_counts[1] = sR1 - sR0;   // ... to the right
    c->h_counts[2] = c->h_counts[3] = 0;
    PST_CUDA(ctx, cudaMemcpyAsync(c->d_counts, c->h_counts, 4 * 4, cudaMemcpyHostToDevice, ctx->stream));
    // 2. counts
    PST_NCCL(ctx, api->GroupStart());
    if (left >= 0) { PST_NCCL(ctx, api->Send(c->d_counts + 0, 4, ncclInt8, left, c->comm, ctx->stream)); PST_NCCL(ctx, api->Recv(c->d_counts + 2, 4, ncclInt8, left, c->comm, ctx->stream)); }
    if (right >= 0) { PST_NCCL(ctx, api->Send(c->d_counts + 1, 4, ncclInt8, right, c->comm, ctx->stream)); PST_NCCL(ctx, api->Recv(c->d_counts + 3, 4, ncclInt8, right, c->comm, ctx->stream)); }
    PST_NCCL(ctx, api->GroupEnd());
    PST_CUDA(ctx, cudaMemcpyAsync(c->h_counts + 2, c->d_counts + 2, 2 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int nL = left >= 0 ? c->h_counts[2] : 0, nR = right >= 0 ? c->h_counts[3] : 0;
    if ((uint64_t)nL > ctx->ghost_cap || (uint64_t)nR > ctx->ghost_cap)
        return pst_fail(ctx, PST_ENOMEM, "ghost layer of %d / %d particles exceeds ghost_capacity %llu", nL, nR, (unsigned long long)ctx->ghost_cap);
    const int n = (int)ctx->n;
    // 3. particles + cell-table slices, one NCCL group
    PST_NCCL(ctx, api->GroupStart());
    for (PstArray* a : ghost_arrays(ctx)) {
        char* base = pst_ptr<char>(ctx, a);
        const size_t es = a->esize;
        if (left >= 0) {
            if (sL1 > sL0) PST_NCCL(ctx, api->Send(base + (ptrdiff_t)sL0 * es, (size_t)(sL1 - sL0) * es, ncclInt8, left, c->comm, ctx->stream));
            if (nL > 0) PST_NCCL(ctx, api->Recv(base - (ptrdiff_t)nL * es, (size_t)nL * es, ncclInt8, left, c->comm, ctx->stream));
        }
        if (right >= 0) {
            if (sR1 > sR0) PST_NCCL(ctx, api->Send(base + (ptrdiff_t)sR0 * es, (size_t)(sR1 - sR0) * es, ncclInt8, right, c->comm, ctx->stream));
            if (nR > 0) PST_NCCL(ctx, api->Recv(base + (ptrdiff_t)n * es, (size_t)nR * es, ncclInt8, right, c->comm, ctx->stream));
        }
    }
    const size_t tab_bytes = ((size_t)layer + 1) * 4;
    if (left >= 0) {
        PST_NCCL(ctx, api->Send(ctx->cell_start + kL0, tab_bytes, ncclInt8, left, c->comm, ctx->stream));
        PST_NCCL(ctx, api->Recv(c->d_tab_l, tab_bytes, ncclInt8, left, c->comm, ctx->stream));
    }
    if (right >= 0) {
        PST_NCCL(ctx, api->Send(ctx->cell_start + kR0, tab_bytes, ncclInt8, right, c->comm, ctx->stream));
        PST_NCCL(ctx, api->Recv(c->d_tab_r, tab_bytes, ncclInt8, right, c->comm, ctx->stream));
    }
    PST_NCCL(ctx, api->GroupEnd());
    // 4. splice the ghost layers into the cell table
    if (left >= 0) {   // ghost layer 0: keys [0, layer)
        PST_LAUNCH(ctx, k_rebase_table, (layer + 255) / 256, 256, 0, layer, c->d_tab_l, -nL, ctx->cell_start);
    }
    if (right >= 0) {  // ghost layer n[0]-1: keys [kR1, kR1 + layer], the last entry is the table end
        PST_LAUNCH(ctx, k_rebase_table, (layer + 1 + 255) / 256, 256, 0, layer + 1, c->d_tab_r, n, ctx->cell_start + kR1);
    }
    ctx->n_ghost_l = nL;
    ctx->n_ghost_r = nR;
    ctx->eos_valid = false;
    return PST_OK;
}
