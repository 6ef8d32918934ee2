// halo.cu -- multi-GPU slab decomposition along x: ghost-particle exchange with the two slab neighbours and
// particle migration.  One context = one rank = one GPU.
//
// Layout trick: with x as the SLOWEST key digit, a rank's outermost owned cell layer is one contiguous
// range of every sorted array.  Left ghosts land at indices [-nL, 0), right ghosts at [n, n + nR) of the same
// allocation, so the whole thing stays ONE monotone index space and the cell table keeps its prefix semantics
// (cell_start is signed for this reason).  The neighbour's cell-table slice for the layer travels with the
// particles and is rebased on arrival.
//
// Three exchanges (option halo_impl / PST_HALO_IMPL):
//   2 (default)  peer memory: k_halo_pack stores the edge layers into the neighbour's cudaIpc receive buffer over
//                NVLink, k_halo_publish raises an epoch word, k_halo_pull waits for it and unpacks.  No NCCL call and
//                no host round trip per step.
//   1            the same fixed-size windows as ONE packed ncclSend/ncclRecv message per neighbour.
//   0            exact per-array ncclSend/ncclRecv straight out of / into the sorted arrays after a count hand-shake.
//
// NCCL is loaded with dlopen at pst_comm_init, so single-GPU users need no NCCL at all and a host that
// already loaded a libnccl (e.g. torch's bundled one) shares it.
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "pst_internal.h"

namespace {

// the slice of the NCCL ABI that is used (stable since NCCL 2.7)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclInt32 = 2 };
enum { ncclMin = 3 };
typedef ncclResult_t (*fn_GetUniqueId)(ncclUniqueId*);
typedef ncclResult_t (*fn_CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
typedef ncclResult_t (*fn_CommDestroy)(ncclComm_t);
typedef ncclResult_t (*fn_Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef ncclResult_t (*fn_Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef ncclResult_t (*fn_AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef ncclResult_t (*fn_Group)(void);
typedef const char* (*fn_GetErrorString)(ncclResult_t);

struct NcclApi {
    void* lib = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_Send Send = nullptr;
    fn_Recv Recv = nullptr;
    fn_AllReduce AllReduce = nullptr;
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
    PST_SYM(AllReduce, "ncclAllReduce")
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
    int32_t* d_counts = nullptr;   // 16 ints: old path [0..1] my send counts (left, right), [2..3] received counts;
                                   // migration [4..6] left neighbour's (n_stay, n_stay+nL, n), [8..10] right neighbour's;
                                   // windowed halo [12] nL, [13] nR (ghost counts, device-side only)
    int32_t* h_counts = nullptr;   // pinned, 16 ints: mirrors of the above as needed
    int32_t* d_tab_l = nullptr;    // received cell-table slice for the left / right ghost layer
    int32_t* d_tab_r = nullptr;
    // windowed halo: one packed message per neighbour and direction = W elements of every ghost array + the table slice
    char* send_buf[2] = {nullptr, nullptr};   // [0] to the left, [1] to the right
    char* recv_buf[2] = {nullptr, nullptr};   // [0] from the left, [1] from the right
    size_t msg_bytes = 0;
    // peer-memory halo (halo_impl = 2): my RECEIVE buffers are double-buffered by step parity and exported with
    // cudaIpc; each neighbour's pack kernel STORES its window straight into them over NVLink (stores are fire-and-forget,
    // remote loads would be latency-bound) and then raises the epoch word at the end of the buffer.
    char* p2p_send[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [side][parity]: MY receive buffers ([0] filled by the left neighbour)
    char* p2p_peer[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [0] = the left neighbour's buffer for what comes from ITS right (= me), [1] likewise
    size_t p2p_bytes = 0;
    bool p2p_failed = false;       // a rank could not map its neighbour's buffer (no P2P / IPC): every rank uses the packed NCCL halo instead
    int epoch = 0;
};

#define PST_NCCL(ctx, expr)                                                                                  \
    do {                                                                                                     \
        ncclResult_t r__ = (expr);                                                                           \
        if (r__ != 0) return pst_fail(ctx, PST_ENCCL, "%s: %s", #expr, nccl_api()->GetErrorString(r__));     \
    } while (0)
// ... between ncclGroupStart and ncclGroupEnd: a failing call still closes the group before returning
#define PST_NCCL_G(ctx, expr)                                                                                \
    do {                                                                                                     \
        ncclResult_t r__ = (expr);                                                                           \
        if (r__ != 0) {                                                                                      \
            nccl_api()->GroupEnd();                                                                          \
            return pst_fail(ctx, PST_ENCCL, "%s: %s", #expr, nccl_api()->GetErrorString(r__));               \
        }                                                                                                    \
    } while (0)

namespace {

// ghost-layer cell table: received slice holds the SENDER's indices; rebase so that the layer's first
// particle lands on `base` in this rank's index space.
__global__ void k_rebase_table(int count, const int32_t* __restrict__ recv, int base, int32_t* __restrict__ cell_start) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < count) cell_start[t] = recv[t] - recv[0] + base;
}
constexpr int kMaxHalo = 20;
struct HaloList {
    char* arr[kMaxHalo];       // owned-view base of each ghost array
    size_t off[kMaxHalo + 1];  // byte offset of each array's window inside a packed message; [n_arr] = the table slice
    int esize[kMaxHalo];
    int n_arr;
};

// Pack both edge windows: blockIdx.y = array (n_arr = the cell-table slice of the edge layer), blockIdx.z = side.
// Left window = elements [0, W), right window = [n - W, n) of every array.
// `exact`: only the elements of the edge layer itself are written (head of the left window, tail of the right one).
__global__ void __launch_bounds__(256) k_halo_pack(HaloList L, int W, int n, int layer, const int32_t* __restrict__ cell_start, size_t kL0,
                                                   size_t kR0, char* __restrict__ buf_l, char* __restrict__ buf_r, int has_l, int has_r, int exact) {
    const int side = blockIdx.z, a = blockIdx.y;
    if (side == 0 ? !has_l : !has_r) return;
    char* buf = side == 0 ? buf_l : buf_r;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (a == L.n_arr) {
        if (t <= layer) reinterpret_cast<int32_t*>(buf + L.off[a])[t] = cell_start[(side == 0 ? kL0 : kR0) + t];
        return;
    }
    if (t >= W) return;
    if (exact) {
        const size_t k0 = side == 0 ? kL0 : kR0;
        const int ext = cell_start[k0 + layer] - cell_start[k0];
        if (side == 0 ? t >= ext : t < W - ext) return;
    }
    const ptrdiff_t s = side == 0 ? t : (ptrdiff_t)n - W + t;
    if (L.esize[a] == 8) reinterpret_cast<unsigned long long*>(buf + L.off[a])[t] = reinterpret_cast<const unsigned long long*>(L.arr[a])[s];
    else reinterpret_cast<uint32_t*>(buf + L.off[a])[t] = reinterpret_cast<const uint32_t*>(L.arr[a])[s];
}

// Unpack the received windows into the ghost regions -- the left neighbour's END-aligned into [-W, 0), the right
// neighbour's START-aligned into [n, n + W) -- and splice the (rebased) table slices into the cell table.
__global__ void __launch_bounds__(256) k_halo_unpack(HaloList L, int W, int n, int layer, int32_t* __restrict__ cell_start, size_t kR1,
                                                     const char* __restrict__ buf_l, const char* __restrict__ buf_r, int has_l, int has_r,
                                                     int32_t* __restrict__ counts, int32_t* __restrict__ flags) {
    const int side = blockIdx.z, a = blockIdx.y;
    if (side == 0 ? !has_l : !has_r) return;
    const char* buf = side == 0 ? buf_l : buf_r;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (a == L.n_arr) {
        const int32_t* recv = reinterpret_cast<const int32_t*>(buf + L.off[a]);
        const int ext = recv[layer] - recv[0];             // the layer's true particle count
        if (side == 0) { if (t < layer) cell_start[t] = recv[t] - recv[0] - ext; }
        else if (t <= layer) cell_start[kR1 + t] = recv[t] - recv[0] + n;
        if (t == 0) { counts[12 + side] = ext; if (ext > W) { atomicExch(&flags[2], 1); atomicMax(&flags[3], ext); } }
        return;
    }
    if (t >= W) return;
    const ptrdiff_t d = side == 0 ? (ptrdiff_t)t - W : (ptrdiff_t)n + t;
    if (L.esize[a] == 8) reinterpret_cast<unsigned long long*>(L.arr[a])[d] = reinterpret_cast<const unsigned long long*>(buf + L.off[a])[t];
    else reinterpret_cast<uint32_t*>(L.arr[a])[d] = reinterpret_cast<const uint32_t*>(buf + L.off[a])[t];
}

// ---- peer-memory variant -------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// after the pack kernel (stream order): make the windows visible system-wide, then raise the epoch words
__global__ void k_halo_publish(int* flag_l, int* flag_r, int epoch) {
    __threadfence_system();
    if (flag_l) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag_l), "r"(epoch) : "memory");
    if (flag_r) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag_r), "r"(epoch) : "memory");
}

// Wait + unpack in one kernel: every CTA waits until the neighbour has raised this step's epoch word at the end of my
// receive buffer, reads the table slice to learn the layer's true particle count, and copies only that many elements of
// every array into my ghost region -- the left neighbour's layer is the TAIL of its right window and lands on
// [-ext, 0), the right neighbour's the HEAD of its left window and lands on [n, n + ext).  No NCCL, no host round trip.
__global__ void __launch_bounds__(256) k_halo_pull(HaloList L, int W, int n, int layer, int32_t* __restrict__ cell_start, size_t kR1,
                                                   const char* __restrict__ peer_l, const char* __restrict__ peer_r, size_t flag_off,
                                                   int epoch, long long timeout_clk, int32_t* __restrict__ counts, int32_t* __restrict__ flags) {
    const int side = blockIdx.z, a = blockIdx.y;
    const char* buf = side == 0 ? peer_l : peer_r;
    if (!buf) return;
    __shared__ int ok;
    if (threadIdx.x == 0) {
        const int* flag = reinterpret_cast<const int*>(buf + flag_off);
        const long long t0 = clock64();
        int seen = ld_acquire_sys(flag);
        while (seen < epoch && (timeout_clk <= 0 || clock64() - t0 < timeout_clk) && *(volatile int*)&flags[4] == 0) { __nanosleep(200); seen = ld_acquire_sys(flag); }
        ok = seen >= epoch;
        if (!ok) atomicExch(&flags[4], 1);        // neighbour never published: reported as PST_ENCCL at the next sync
    }
    __syncthreads();
    if (!ok) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t* recv = reinterpret_cast<const int32_t*>(buf + L.off[L.n_arr]);
    const int r0 = __ldcg(recv), ext = __ldcg(recv + layer) - r0;          // the layer's true particle count
    if (a == L.n_arr) {
        if (side == 0) { if (t < layer) cell_start[t] = __ldcg(recv + t) - r0 - ext; }
        else if (t <= layer) cell_start[kR1 + t] = __ldcg(recv + t) - r0 + n;
        if (t == 0) { counts[12 + side] = ext; if (ext > W) { atomicExch(&flags[2], 1); atomicMax(&flags[3], ext); } }
        return;
    }
    const int m = min(ext, W);
    if (t >= m) return;
    const ptrdiff_t src = side == 0 ? (ptrdiff_t)W - m + t : t;
    const ptrdiff_t dst = side == 0 ? (ptrdiff_t)t - m : (ptrdiff_t)n + t;
    // loads bypass L1 (__ldcg): the neighbour rewrites this buffer over NVLink every second step
    if (L.esize[a] == 8) reinterpret_cast<unsigned long long*>(L.arr[a])[dst] = __ldcg(reinterpret_cast<const unsigned long long*>(buf + L.off[a]) + src);
    else reinterpret_cast<uint32_t*>(L.arr[a])[dst] = __ldcg(reinterpret_cast<const uint32_t*>(buf + L.off[a]) + src);
}

__global__ void k_set_flag(int32_t* p, int32_t v) { *p = v; }

// migration: does this rank have room for what its neighbours send?  mine = (n_stay, n_stay + nL, n); a neighbour's triple
// likewise -- the left one's right-leavers and the right one's left-leavers arrive here
__global__ void k_mig_room(const int32_t* mine, const int32_t* from_l, const int32_t* from_r, int has_l, int has_r, long long capacity, int32_t* ok) {
    const long long aL = has_l ? from_l[2] - from_l[1] : 0, aR = has_r ? from_r[1] - from_r[0] : 0;
    *ok = (long long)mine[0] + aL + aR <= capacity;
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
    if (ctx->d_bodies) return pst_fail(ctx, PST_ESTATE, "rigid bodies are not supported with a communicator attached");
    NcclApi* api = nccl_api();
    if (!api->lib) return pst_fail(ctx, PST_ENCCL, "%s", api->err.c_str());
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    PstGrid& g = ctx->grid;
    const double ext = ctx->cfg.hi[0] - ctx->cfg.lo[0];
    const double cells = ext / g.cell;
    if (std::fabs(cells - std::round(cells)) > 1e-6 * std::max(1.0, cells))
        return pst_fail(ctx, PST_EINVAL, "slab x extent %.17g is not a whole number of cells (%.17g)", ext, g.cell);
    PstComm* c = new PstComm();
    c->rank = rank; c->nranks = n_ranks;
    ctx->comm = c;                                   // the grid of a slab has the two ghost layers
    const pst_status gs = pst_grid_finalize(ctx);
    ctx->comm = nullptr;
    if (gs != PST_OK) { delete c; pst_grid_finalize(ctx); return gs; }
    const size_t layer = (size_t)g.n[1] * g.n[2] + 1;
    if (cudaMalloc((void**)&c->d_counts, 16 * 4) != cudaSuccess || cudaMalloc((void**)&c->d_tab_l, layer * 4) != cudaSuccess ||
        cudaMalloc((void**)&c->d_tab_r, layer * 4) != cudaSuccess || cudaHostAlloc((void**)&c->h_counts, 16 * 4, cudaHostAllocDefault) != cudaSuccess) {
        cudaFree(c->d_counts); cudaFree(c->d_tab_l); cudaFree(c->d_tab_r);
        delete c;
        pst_grid_finalize(ctx);
        return pst_fail(ctx, PST_ENOMEM, "halo buffers");
    }
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, sizeof id);
    ncclResult_t r = api->CommInitRank(&c->comm, n_ranks, id, rank);
    if (r != 0) { delete c; pst_grid_finalize(ctx); return pst_fail(ctx, PST_ENCCL, "ncclCommInitRank: %s", api->GetErrorString(r)); }
    ctx->comm = c;
    ctx->nbrs_valid = false;
    return PST_OK;
}

pst_status pst_comm_destroy(pst_ctx* ctx) {
    if (!ctx->comm) return PST_OK;
    PstComm* c = ctx->comm;
    if (c->comm) nccl_api()->CommDestroy(c->comm);
    cudaFree(c->d_counts); cudaFree(c->d_tab_l); cudaFree(c->d_tab_r);
    for (int k = 0; k < 2; ++k) { cudaFree(c->send_buf[k]); cudaFree(c->recv_buf[k]); }
    for (int sd = 0; sd < 2; ++sd)
        for (int par = 0; par < 2; ++par) {
            if (c->p2p_peer[sd][par]) cudaIpcCloseMemHandle(c->p2p_peer[sd][par]);
            cudaFree(c->p2p_send[sd][par]);
        }
    if (c->h_counts) cudaFreeHost(c->h_counts);
    delete c;
    ctx->comm = nullptr;
    return PST_OK;
}

// true ghost counts (the windowed halo keeps them on the device; this synchronises)
pst_status pst_ghost_counts(pst_ctx* ctx, int64_t* nl, int64_t* nr) {
    *nl = ctx->n_ghost_l; *nr = ctx->n_ghost_r;
    if (!ctx->comm || ctx->ghost_exact || (ctx->n_ghost_l == 0 && ctx->n_ghost_r == 0)) return PST_OK;
    PstComm* c = ctx->comm;
    PST_CUDA(ctx, cudaMemcpyAsync(c->h_counts + 12, c->d_counts + 12, 2 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *nl = c->h_counts[12]; *nr = c->h_counts[13];
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
    // The three table entries (n_stay, n_stay + nL, n) go to both neighbours straight from device memory, then ONE
    // read-back + stream sync tells the host its own split and what arrives (was: sync, count hand-shake, sync).
    PST_NCCL(ctx, api->GroupStart());
    if (left >= 0) { PST_NCCL_G(ctx, api->Send(ctx->cell_start + nc, 12, ncclInt8, left, c->comm, ctx->stream)); PST_NCCL_G(ctx, api->Recv(c->d_counts + 4, 12, ncclInt8, left, c->comm, ctx->stream)); }
    if (right >= 0) { PST_NCCL_G(ctx, api->Send(ctx->cell_start + nc, 12, ncclInt8, right, c->comm, ctx->stream)); PST_NCCL_G(ctx, api->Recv(c->d_counts + 8, 12, ncclInt8, right, c->comm, ctx->stream)); }
    PST_NCCL(ctx, api->GroupEnd());
    // Does everybody have room for its arrivals?  Decided on the device and combined over ALL ranks (one 4-byte all-reduce in
    // the same stream), so either every rank posts the payload exchange below or none does: a rank that is full can never
    // leave its neighbours waiting in a send that is not matched.
    k_mig_room<<<1, 1, 0, ctx->stream>>>(ctx->cell_start + nc, c->d_counts + 4, c->d_counts + 8, left >= 0, right >= 0, (long long)ctx->capacity, c->d_counts + 14);
    PST_NCCL(ctx, api->AllReduce(c->d_counts + 14, c->d_counts + 14, 1, ncclInt32, ncclMin, c->comm, ctx->stream));
    PST_CUDA(ctx, cudaMemcpyAsync(c->h_counts, ctx->cell_start + nc, 3 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaMemcpyAsync(c->h_counts + 4, c->d_counts + 4, 11 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int n_stay = c->h_counts[0], nL = c->h_counts[1] - c->h_counts[0], nR = c->h_counts[2] - c->h_counts[1];
    // what the left neighbour sends me is ITS right-leavers, and vice versa
    const int aL = left >= 0 ? c->h_counts[6] - c->h_counts[5] : 0, aR = right >= 0 ? c->h_counts[9] - c->h_counts[8] : 0;
    if (!c->h_counts[14]) {
        if ((uint64_t)n_stay + aL + aR > ctx->capacity)
            return pst_fail(ctx, PST_ENOMEM, "migration: %d stayers + %d arrivals exceed capacity %llu", n_stay, aL + aR, (unsigned long long)ctx->capacity);
        return pst_fail(ctx, PST_ENOMEM, "migration: another rank has no room for its arrivals (its capacity is exceeded); nothing was exchanged");
    }
    if (nL + nR + aL + aR > 0) {
        PST_TRY(pst_resolve_history(ctx));   // contact-history rows travel with their particles: put them in the new order first
        PST_NCCL(ctx, api->GroupStart());
        for (auto& a : ctx->arrays) {
            if (!(a.flags & PST_ARRAY_PERSISTENT)) continue;
            const size_t es = a.esize;
            for (int r = 0; r < a.rows; ++r) {
                char* cur = pst_ptr<char>(ctx, &a, r, a.cur);
                char* alt = pst_ptr<char>(ctx, &a, r, 1 - a.cur);
                if (left >= 0) {
                    if (nL > 0) PST_NCCL_G(ctx, api->Send(cur + (size_t)n_stay * es, (size_t)nL * es, ncclInt8, left, c->comm, ctx->stream));
                    if (aL > 0) PST_NCCL_G(ctx, api->Recv(alt, (size_t)aL * es, ncclInt8, left, c->comm, ctx->stream));
                }
                if (right >= 0) {
                    if (nR > 0) PST_NCCL_G(ctx, api->Send(cur + (size_t)(n_stay + nL) * es, (size_t)nR * es, ncclInt8, right, c->comm, ctx->stream));
                    if (aR > 0) PST_NCCL_G(ctx, api->Recv(alt + (size_t)aL * es, (size_t)aR * es, ncclInt8, right, c->comm, ctx->stream));
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
    if (aL + aR > 0) ctx->state_epoch++;      // (leavers alone only shorten the owned range: the rows that stay, and their records, are untouched)
    if (aL + aR > 0) ctx->uni_dirty = true;   // arrivals carry their own m and h
    *arrivals = aL + aR;
    return PST_OK;
}

// one-time setup of the peer-memory halo: allocate my double-buffered receive windows, swap cudaIpc handles with the two
// slab neighbours (the 64-byte handles travel by ncclSend/ncclRecv, once), map theirs.  Whether that works is a property of
// the machine (no P2P between the two GPUs, ranks in different IPC namespaces or on different nodes, threads of one process):
// every rank tries, the outcomes are combined with ONE ncclAllReduce(min), and if any rank failed ALL ranks release what they
// mapped and use the packed NCCL halo (halo_impl = 1) from then on -- a collective decision, so no rank waits on an epoch
// word nobody will raise.  *use_p2p reports the outcome.
static pst_status p2p_setup(pst_ctx* ctx, size_t bytes, int left, int right, bool* use_p2p) {
    PstComm* c = ctx->comm;
    NcclApi* api = nccl_api();
    *use_p2p = !c->p2p_failed;
    if (c->p2p_failed || c->p2p_bytes >= bytes) return PST_OK;
    if (c->p2p_bytes != 0) return pst_fail(ctx, PST_ESTATE, "peer-memory halo: the message size changed after setup");
    cudaIpcMemHandle_t mine[2][2], theirs[2][2];
    std::memset(mine, 0, sizeof mine);
    std::memset(theirs, 0, sizeof theirs);
    int ok = 1;
    for (int sd = 0; sd < 2 && ok; ++sd)
        for (int par = 0; par < 2 && ok; ++par) {
            if (cudaMalloc((void**)&c->p2p_send[sd][par], bytes) != cudaSuccess) { c->p2p_send[sd][par] = nullptr; ok = 0; break; }
            if (cudaMemsetAsync(c->p2p_send[sd][par], 0, bytes, ctx->stream) != cudaSuccess) ok = 0;
            if (cudaIpcGetMemHandle(&mine[sd][par], c->p2p_send[sd][par]) != cudaSuccess) ok = 0;
        }
    cudaGetLastError();                                // a failed probe must not poison the context
    constexpr size_t HB = sizeof(cudaIpcMemHandle_t);
    char* d_h = nullptr;                               // [0..1] mine for the left, [2..3] mine for the right, [4..5] from left, [6..7] from right
    PST_CUDA(ctx, cudaMalloc((void**)&d_h, 8 * HB + 16));
    int32_t* d_ok = reinterpret_cast<int32_t*>(d_h + 8 * HB);
    PST_CUDA(ctx, cudaMemcpyAsync(d_h, mine, 4 * HB, cudaMemcpyHostToDevice, ctx->stream));
    // the swap always runs (a rank whose allocation failed sends zeroed handles): the neighbours' receives must be matched
    PST_NCCL(ctx, api->GroupStart());
    if (left >= 0) { PST_NCCL_G(ctx, api->Send(d_h, 2 * HB, ncclInt8, left, c->comm, ctx->stream)); PST_NCCL_G(ctx, api->Recv(d_h + 4 * HB, 2 * HB, ncclInt8, left, c->comm, ctx->stream)); }
    if (right >= 0) { PST_NCCL_G(ctx, api->Send(d_h + 2 * HB, 2 * HB, ncclInt8, right, c->comm, ctx->stream)); PST_NCCL_G(ctx, api->Recv(d_h + 6 * HB, 2 * HB, ncclInt8, right, c->comm, ctx->stream)); }
    PST_NCCL(ctx, api->GroupEnd());
    PST_CUDA(ctx, cudaMemcpyAsync(theirs, d_h + 4 * HB, 4 * HB, cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // what the left neighbour sent me are the handles of ITS windows for its RIGHT neighbour (= me), and vice versa
    const bool forced_off = std::getenv("PST_P2P_DISABLE") != nullptr;     // (tests: exercise the fallback on a machine that has P2P)
    for (int par = 0; par < 2 && ok; ++par) {
        if (left >= 0 && (forced_off || cudaIpcOpenMemHandle((void**)&c->p2p_peer[0][par], theirs[0][par], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)) { c->p2p_peer[0][par] = nullptr; ok = 0; }
        if (right >= 0 && ok && (forced_off || cudaIpcOpenMemHandle((void**)&c->p2p_peer[1][par], theirs[1][par], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)) { c->p2p_peer[1][par] = nullptr; ok = 0; }
    }
    cudaGetLastError();
    // the collective decision
    k_set_flag<<<1, 1, 0, ctx->stream>>>(d_ok, ok);
    PST_NCCL(ctx, api->AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, c->comm, ctx->stream));
    int all_ok = 0;
    PST_CUDA(ctx, cudaMemcpyAsync(&all_ok, d_ok, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_h);
    if (!all_ok) {
        for (int sd = 0; sd < 2; ++sd)
            for (int par = 0; par < 2; ++par) {
                if (c->p2p_peer[sd][par]) cudaIpcCloseMemHandle(c->p2p_peer[sd][par]);
                cudaFree(c->p2p_send[sd][par]);
                c->p2p_peer[sd][par] = c->p2p_send[sd][par] = nullptr;
            }
        cudaGetLastError();
        c->p2p_failed = true;
        *use_p2p = false;
        return PST_OK;
    }
    c->p2p_bytes = bytes;
    return PST_OK;
}

extern "C" pst_status pst_halo_exchange(pst_ctx* ctx) {
    PstRange range("pst_halo_exchange");
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
    // 2 = peer memory (default), 1 = one packed NCCL message per neighbour, 0 = exact per-array NCCL exchange with count hand-shake
    static const int impl_default = std::getenv("PST_HALO_IMPL") ? std::atoi(std::getenv("PST_HALO_IMPL")) : 2;
    const int impl = pst_option(ctx, "halo_impl", impl_default);
    if (impl == 1 || impl == 2) {
        // ---- windowed exchange: no count hand-shake, no host round trip, ONE NCCL group.
        // My first owned layer starts at index 0 and my last one ends at n, so fixed windows of W = ghost_capacity
        // elements -- [0, W) to the left, [n - W, n) to the right -- always contain the edge layers.  The left
        // neighbour's window is received END-aligned into [-W, 0), the right neighbour's START-aligned into [n, n + W):
        // the layers land exactly where the cell table expects them, whatever their true size; the surplus is never
        // referenced (the table slices that travel alongside carry the true extents, rebased on the device).
        // All windows of one direction travel as ONE packed message (a pack kernel, one ncclSend/ncclRecv per
        // neighbour, an unpack kernel): 14 + 1 separate sends per neighbour cost ~0.5 ms of NCCL per-operation overhead.
        const int W = (int)ctx->ghost_cap, n = (int)ctx->n;
        HaloList L;
        L.n_arr = 0;
        size_t off = 0;
        std::vector<PstArray*> ga = ghost_arrays(ctx);
        std::stable_sort(ga.begin(), ga.end(), [](PstArray* x, PstArray* y) { return x->esize > y->esize; });   // 8-byte arrays first: alignment
        if ((int)ga.size() > kMaxHalo) return pst_fail(ctx, PST_EINVAL, "too many ghost arrays");
        for (PstArray* a : ga) {
            L.arr[L.n_arr] = pst_ptr<char>(ctx, a);
            L.esize[L.n_arr] = (int)a->esize;
            L.off[L.n_arr] = off;
            off += (size_t)W * a->esize;
            ++L.n_arr;
        }
        L.off[L.n_arr] = off;
        const size_t msg = off + ((size_t)layer + 1) * 4;
        const int span = std::max(W, layer + 1);
        const dim3 grid((unsigned)((span + 255) / 256), (unsigned)(L.n_arr + 1), 2);
        if (impl == 2) {
            // ---- peer memory: the pack kernel stores this step's edge layers into the neighbours' parity buffers over
            // NVLink, a one-thread kernel raises the epoch words there, the unpack kernel waits for MY buffers' epoch
            const size_t flag_off = (msg + 15) & ~(size_t)15;
            bool use_p2p = true;
            PST_TRY(p2p_setup(ctx, flag_off + 16, left, right, &use_p2p));
            if (use_p2p) {
            // how long the pull kernel waits for a neighbour's epoch word before it gives up (option halo_timeout_ms, default
            // 10 minutes, 0 = for ever): rank skew of seconds is normal (checkpoint or VTK output on one rank, first-call allocation)
            static int clk_khz = 0;
            if (!clk_khz) cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, ctx->cfg.device);
            const long long timeout_clk = (long long)pst_option(ctx, "halo_timeout_ms", 600000) * (long long)std::max(clk_khz, 1000000);
            const int e = ++c->epoch, par = e & 1;
            PST_LAUNCH(ctx, k_halo_pack, grid, 256, 0, L, W, n, layer, ctx->cell_start, kL0, kR0, c->p2p_peer[0][par], c->p2p_peer[1][par], left >= 0, right >= 0, 1);
            PST_LAUNCH(ctx, k_halo_publish, 1, 1, 0, left >= 0 ? (int*)(c->p2p_peer[0][par] + flag_off) : nullptr,
                       right >= 0 ? (int*)(c->p2p_peer[1][par] + flag_off) : nullptr, e);
            PST_CUDA(ctx, cudaMemsetAsync(c->d_counts + 12, 0, 2 * 4, ctx->stream));
            PST_LAUNCH(ctx, k_halo_pull, grid, 256, 0, L, W, n, layer, ctx->cell_start, kR1, left >= 0 ? (const char*)c->p2p_send[0][par] : nullptr,
                       right >= 0 ? (const char*)c->p2p_send[1][par] : nullptr, flag_off, e, timeout_clk, c->d_counts, ctx->d_flags);
            ctx->n_ghost_l = left >= 0 ? W : 0;
            ctx->n_ghost_r = right >= 0 ? W : 0;
            ctx->ghost_exact = false;
            pst_note_ghosts_changed(ctx);
            return PST_OK;
            }      // (no peer access on this machine: fall through to the packed NCCL message)
        }
        if (c->msg_bytes < msg) {
            PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            for (int k = 0; k < 2; ++k) {
                cudaFree(c->send_buf[k]); cudaFree(c->recv_buf[k]);
                c->send_buf[k] = c->recv_buf[k] = nullptr;
                if (cudaMalloc((void**)&c->send_buf[k], msg) != cudaSuccess || cudaMalloc((void**)&c->recv_buf[k], msg) != cudaSuccess)
                    return pst_fail(ctx, PST_ENOMEM, "halo message buffers (%zu bytes)", msg);
            }
            c->msg_bytes = msg;
        }
        PST_LAUNCH(ctx, k_halo_pack, grid, 256, 0, L, W, n, layer, ctx->cell_start, kL0, kR0, c->send_buf[0], c->send_buf[1], left >= 0, right >= 0, 0);
        PST_NCCL(ctx, api->GroupStart());
        if (left >= 0) {
            PST_NCCL_G(ctx, api->Send(c->send_buf[0], msg, ncclInt8, left, c->comm, ctx->stream));
            PST_NCCL_G(ctx, api->Recv(c->recv_buf[0], msg, ncclInt8, left, c->comm, ctx->stream));
        }
        if (right >= 0) {
            PST_NCCL_G(ctx, api->Send(c->send_buf[1], msg, ncclInt8, right, c->comm, ctx->stream));
            PST_NCCL_G(ctx, api->Recv(c->recv_buf[1], msg, ncclInt8, right, c->comm, ctx->stream));
        }
        PST_NCCL(ctx, api->GroupEnd());
        PST_CUDA(ctx, cudaMemsetAsync(c->d_counts + 12, 0, 2 * 4, ctx->stream));
        PST_LAUNCH(ctx, k_halo_unpack, grid, 256, 0, L, W, n, layer, ctx->cell_start, kR1, c->recv_buf[0], c->recv_buf[1], left >= 0, right >= 0,
                   c->d_counts, ctx->d_flags);
        // the host only knows bounds: per-particle passes over "owned + ghosts" cover the whole windows
        ctx->n_ghost_l = left >= 0 ? W : 0;
        ctx->n_ghost_r = right >= 0 ? W : 0;
        ctx->ghost_exact = false;
        pst_note_ghosts_changed(ctx);
        return PST_OK;
    }
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
    if (left >= 0) { PST_NCCL_G(ctx, api->Send(c->d_counts + 0, 4, ncclInt8, left, c->comm, ctx->stream)); PST_NCCL_G(ctx, api->Recv(c->d_counts + 2, 4, ncclInt8, left, c->comm, ctx->stream)); }
    if (right >= 0) { PST_NCCL_G(ctx, api->Send(c->d_counts + 1, 4, ncclInt8, right, c->comm, ctx->stream)); PST_NCCL_G(ctx, api->Recv(c->d_counts + 3, 4, ncclInt8, right, c->comm, ctx->stream)); }
    PST_NCCL(ctx, api->GroupEnd());
    PST_CUDA(ctx, cudaMemcpyAsync(c->h_counts + 2, c->d_counts + 2, 2 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int nL = left >= 0 ? c->h_counts[2] : 0, nR = right >= 0 ? c->h_counts[3] : 0;
    {   // every rank posts the payload group below or none does (a rank that is over capacity must not leave unmatched sends behind)
        const int room = (uint64_t)nL <= ctx->ghost_cap && (uint64_t)nR <= ctx->ghost_cap;
        k_set_flag<<<1, 1, 0, ctx->stream>>>(c->d_counts + 14, room);
        PST_NCCL(ctx, api->AllReduce(c->d_counts + 14, c->d_counts + 14, 1, ncclInt32, ncclMin, c->comm, ctx->stream));
        PST_CUDA(ctx, cudaMemcpyAsync(c->h_counts + 14, c->d_counts + 14, 4, cudaMemcpyDeviceToHost, ctx->stream));
        PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (!room) return pst_fail(ctx, PST_ENOMEM, "ghost layer of %d / %d particles exceeds ghost_capacity %llu", nL, nR, (unsigned long long)ctx->ghost_cap);
        if (!c->h_counts[14]) return pst_fail(ctx, PST_ENOMEM, "halo: another rank's ghost layer exceeds its ghost_capacity; nothing was exchanged");
    }
    const int n = (int)ctx->n;
    // 3. particles + cell-table slices, one NCCL group
    PST_NCCL(ctx, api->GroupStart());
    for (PstArray* a : ghost_arrays(ctx)) {
        char* base = pst_ptr<char>(ctx, a);
        const size_t es = a->esize;
        if (left >= 0) {
            if (sL1 > sL0) PST_NCCL_G(ctx, api->Send(base + (ptrdiff_t)sL0 * es, (size_t)(sL1 - sL0) * es, ncclInt8, left, c->comm, ctx->stream));
            if (nL > 0) PST_NCCL_G(ctx, api->Recv(base - (ptrdiff_t)nL * es, (size_t)nL * es, ncclInt8, left, c->comm, ctx->stream));
        }
        if (right >= 0) {
            if (sR1 > sR0) PST_NCCL_G(ctx, api->Send(base + (ptrdiff_t)sR0 * es, (size_t)(sR1 - sR0) * es, ncclInt8, right, c->comm, ctx->stream));
            if (nR > 0) PST_NCCL_G(ctx, api->Recv(base + (ptrdiff_t)n * es, (size_t)nR * es, ncclInt8, right, c->comm, ctx->stream));
        }
    }
    const size_t tab_bytes = ((size_t)layer + 1) * 4;
    if (left >= 0) {
        PST_NCCL_G(ctx, api->Send(ctx->cell_start + kL0, tab_bytes, ncclInt8, left, c->comm, ctx->stream));
        PST_NCCL_G(ctx, api->Recv(c->d_tab_l, tab_bytes, ncclInt8, left, c->comm, ctx->stream));
    }
    if (right >= 0) {
        PST_NCCL_G(ctx, api->Send(ctx->cell_start + kR0, tab_bytes, ncclInt8, right, c->comm, ctx->stream));
        PST_NCCL_G(ctx, api->Recv(c->d_tab_r, tab_bytes, ncclInt8, right, c->comm, ctx->stream));
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
    ctx->ghost_exact = true;
    pst_note_ghosts_changed(ctx);
    return PST_OK;
}
