// nnps.cu -- cell-linked-list neighbour search on device-resident SoA arrays:
//   k_keys      cell key (linear, last axis fastest | Morton) per particle          
//   radix sort  (key, previous index) pairs, only the key bits that are in use      (a8)
//   k_bounds    cell start table with prefix semantics (start[c+1] = end of cell c) (a8)
//   k_permute   gather every persistent array into cell order in one pass           (a9)
//   k_remap_history  move each particle's contact-history row with it               (a9)
// plus the id-order <-> cell-order reorder kernels behind pst_upload / pst_download and the
// neighbour/contact-set dump used as the parity hook.
// No reference code exists for any of this; the loop it feeds is
// prestige/src/codegen/simple_cpu.rs:7-16.
#include <cub/device/device_radix_sort.cuh>

#include "pst_internal.h"

namespace {

constexpr int kThreads = 256;
inline unsigned blocks_for(size_t n, int t = kThreads) { return (unsigned)((n + t - 1) / t); }

__global__ void k_iota(uint32_t* __restrict__ id, int n) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) id[s] = (uint32_t)s;
}

// mig_l / mig_r (slab decomposition only): an owned particle that has left the slab through the left / right face gets
// the sentinel key ncells / ncells + 1, so the sort parks the leavers behind the stayers as two contiguous ranges that
// can be sent to the neighbour rank as they are.
template <class R, int DIM, bool MORTON>
__global__ void __launch_bounds__(kThreads) k_keys(GridDev<R> g, int n, const R* __restrict__ x, const R* __restrict__ y,
                                                   const R* __restrict__ z, uint32_t* __restrict__ keys,
                                                   uint32_t* __restrict__ vals, uint32_t ncells, int mig_l, int mig_r) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int cxu = (int)floor((x[s] - g.lo[0]) * g.inv[0]);
    const int cx = min(max(cxu, g.cx_lo), g.cx_hi);
    const int cy = cell_coord<R>(y[s], g.lo[1], g.inv[1], 0, g.n[1] - 1);
    const int cz = DIM == 3 ? cell_coord<R>(z[s], g.lo[2], g.inv[2], 0, g.n[2] - 1) : 0;
    uint32_t key = cell_key<DIM, MORTON>(g, cx, cy, cz);
    if (mig_l && cxu < g.cx_lo) key = ncells;
    else if (mig_r && cxu > g.cx_hi) key = ncells + 1u;
    keys[s] = key;
    vals[s] = (uint32_t)s;
}

// start[k] = first sorted index whose key is >= k.  Thread s owns the keys in (key[s-1], key[s]].
// Also counts the occupied cells (counters[2]) so the host can size pair-kernel tiles from the real
// mean occupancy instead of a guess.
__global__ void __launch_bounds__(kThreads) k_bounds(int n, uint32_t ncells, const uint32_t* __restrict__ keys,
                                                     int32_t* __restrict__ cell_start, unsigned long long* __restrict__ counters) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = s <= n;
    long long prev = 0, cur = -1;
    if (active) {
        prev = s == 0 ? -1ll : (long long)keys[s - 1];
        cur = s == n ? (long long)ncells : (long long)keys[s];
    }
    const int occupied = __syncthreads_count(active && s < n && cur != prev);   // one atomic per block, not per warp
    if (threadIdx.x == 0 && occupied) atomicAdd(&counters[2], (unsigned long long)occupied);
    for (long long k = prev + 1; k <= cur; ++k) cell_start[k] = s;
}

// ---------------------------------------------------------------------------------------------
// Hand-written cell sort (default; option sort_impl = 1): a counting sort over the cell keys that also PRODUCES the
// cell table, so it replaces both the radix sort and k_bounds.
//   k_keys_count   key per particle + provisional rank r = atomicAdd(count[key], 1)
//   k_scan_*       exclusive scan of the counts, in place -> cell_start (prefix semantics), 3 small kernels
//   k_place        slot[cell_start[key] + r] = previous index
//   k_cell_order   per cell: sort its <= 32 slots ascending (cells with more go through k_cell_order_big)
// The atomics make r nondeterministic; sorting each cell's slots by previous index restores exactly the order a
// STABLE sort gives, so results are run-to-run reproducible and bit-identical to the CUB path (tested).
// ---------------------------------------------------------------------------------------------
template <class R, int DIM, bool MORTON>
__global__ void __launch_bounds__(kThreads) k_keys_count(GridDev<R> g, int n, const R* __restrict__ x, const R* __restrict__ y,
                                                         const R* __restrict__ z, uint32_t* __restrict__ keys, uint32_t* __restrict__ prov,
                                                         int32_t* __restrict__ count, uint32_t ncells, int mig_l, int mig_r) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int cxu = (int)floor((x[s] - g.lo[0]) * g.inv[0]);
    const int cx = min(max(cxu, g.cx_lo), g.cx_hi);
    const int cy = cell_coord<R>(y[s], g.lo[1], g.inv[1], 0, g.n[1] - 1);
    const int cz = DIM == 3 ? cell_coord<R>(z[s], g.lo[2], g.inv[2], 0, g.n[2] - 1) : 0;
    uint32_t key = cell_key<DIM, MORTON>(g, cx, cy, cz);
    if (mig_l && cxu < g.cx_lo) key = ncells;
    else if (mig_r && cxu > g.cx_hi) key = ncells + 1u;
    keys[s] = key;
    prov[s] = (uint32_t)atomicAdd(&count[key], 1);
}

constexpr int kScanThreads = 512, kScanItems = 8, kScanTile = kScanThreads * kScanItems;

// block-level exclusive scan of one tile, in place; tile total -> sums[blockIdx.x]; also counts non-empty cells
// (`sub`: the fast axis of the grid is subdivided; `sub` consecutive entries -- aligned, since every column holds a multiple
// of `sub` fine cells -- form one coarse cell, and it is the coarse cells that are counted)
__global__ void __launch_bounds__(kScanThreads) k_scan_tiles(int m, int32_t* __restrict__ a, int32_t* __restrict__ sums,
                                                             unsigned long long* __restrict__ counters, int sub) {
    __shared__ int warp_tot[kScanThreads / 32];
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems], t = 0, occ = 0, grp = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = base + k < m ? a[base + k] : 0;
        grp += v[k];
        if (((k + 1) & (sub - 1)) == 0) { occ += grp != 0; grp = 0; }
        const int x = v[k]; v[k] = t; t += x;          // exclusive within the thread
    }
    int incl = t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += warp_tot[w];
    const int toff = woff + incl - t;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < m) a[base + k] = v[k] + toff;
    if (threadIdx.x == kScanThreads - 1) sums[blockIdx.x] = woff + incl;
    // occupied cells, for the tile-depth heuristic of the pair kernels
    for (int d = 16; d > 0; d >>= 1) occ += __shfl_down_sync(0xffffffffu, occ, d);
    if (counters && lane == 0 && occ) atomicAdd(&counters[2], (unsigned long long)occ);
}

// single block: exclusive scan of the tile sums
__global__ void __launch_bounds__(1024) k_scan_sums(int nb, int32_t* __restrict__ sums) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const int x = i < nb ? sums[i] : 0;
        int incl = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += warp_tot[w];
        const int carry = carry_s;
        if (i < nb) sums[i] = carry + woff + incl - x;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + woff + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kScanThreads) k_scan_add(int m, int32_t* __restrict__ a, const int32_t* __restrict__ sums) {
    const int off = sums[blockIdx.x];
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < m) a[base + k] += off;
}

__global__ void __launch_bounds__(kThreads) k_place(int n, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ prov,
                                                    const int32_t* __restrict__ cell_start, uint32_t* __restrict__ slot) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) slot[cell_start[keys[s]] + prov[s]] = (uint32_t)s;
}

constexpr int kSmallCell = 32;
// one thread per cell: insertion sort of its slots (nearly sorted already: the atomics arrive roughly in index order)
__global__ void __launch_bounds__(kThreads) k_cell_order(uint32_t nkeys, const int32_t* __restrict__ cell_start, uint32_t* __restrict__ slot,
                                                         uint32_t* __restrict__ big_list, unsigned long long* __restrict__ counters) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nkeys) return;
    const int b = cell_start[c], e = cell_start[c + 1], cnt = e - b;
    if (cnt <= 1) return;
    if (cnt > kSmallCell) {
        const unsigned long long k = atomicAdd(&counters[4], 1ull);
        big_list[k] = c;                       // capacity nkeys: cannot overflow
        return;
    }
    uint32_t v[kSmallCell];
    for (int i = 0; i < cnt; ++i) v[i] = slot[b + i];
    for (int i = 1; i < cnt; ++i) {
        const uint32_t x = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; }
        v[j + 1] = x;
    }
    for (int i = 0; i < cnt; ++i) slot[b + i] = v[i];
}

// crowded cells (e.g. out-of-box particles clamped into an edge cell): one block per cell, rank by counting.
// Values are distinct, so rank = #smaller is a permutation.  `tmp` is scratch of the same size as `slot`.
__global__ void __launch_bounds__(256) k_cell_order_big(const int32_t* __restrict__ cell_start, uint32_t* __restrict__ slot,
                                                        uint32_t* __restrict__ tmp, const uint32_t* __restrict__ big_list,
                                                        const unsigned long long* __restrict__ counters) {
    const unsigned long long nbig = counters[4];
    for (unsigned long long q = blockIdx.x; q < nbig; q += gridDim.x) {
        const uint32_t c = big_list[q];
        const int b = cell_start[c], e = cell_start[c + 1];
        for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
            const uint32_t x = slot[i];
            int r = 0;
            for (int j = b; j < e; ++j) r += slot[j] < x;
            tmp[b + r] = x;
        }
        __syncthreads();
        for (int i = b + threadIdx.x; i < e; i += blockDim.x) slot[i] = tmp[i];
        __syncthreads();
    }
}

constexpr int kMaxPermute = 40;
struct PermuteList {
    const void* src[kMaxPermute];
    void* dst[kMaxPermute];
    int n8, n4;  // entries [0, n8) are 8-byte, [n8, n8 + n4) are 4-byte
};

// One thread per destination slot: read the source index once, then issue every gather back to back
// (independent loads -> deep memory-level parallelism), stores are fully coalesced.
__global__ void __launch_bounds__(kThreads) k_permute(PermuteList L, int n, const uint32_t* __restrict__ perm) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t p = perm[s];
#pragma unroll 4
    for (int a = 0; a < L.n8; ++a)
        reinterpret_cast<unsigned long long*>(L.dst[a])[s] = __ldg(reinterpret_cast<const unsigned long long*>(L.src[a]) + p);
#pragma unroll 4
    for (int a = L.n8; a < L.n8 + L.n4; ++a)
        reinterpret_cast<uint32_t*>(L.dst[a])[s] = __ldg(reinterpret_cast<const uint32_t*>(L.src[a]) + p);
}

// Contact history follows its particle.  Only the slots in use are moved (slot-major layout, so
// slot k of consecutive particles is contiguous).
template <class R>
__global__ void __launch_bounds__(kThreads) k_remap_history(int n, size_t stride, const uint32_t* __restrict__ perm,
                                                            const int32_t* __restrict__ hn_s, const uint32_t* __restrict__ hid_s,
                                                            const R* __restrict__ hx_s, const R* __restrict__ hy_s,
                                                            const R* __restrict__ hz_s, int32_t* __restrict__ hn_d,
                                                            uint32_t* __restrict__ hid_d, R* __restrict__ hx_d,
                                                            R* __restrict__ hy_d, R* __restrict__ hz_d) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t p = perm[s];
    const int cnt = hn_s[p];
    hn_d[s] = cnt;
    for (int k = 0; k < cnt; ++k) {
        const size_t o = (size_t)k * stride;
        hid_d[o + s] = hid_s[o + p];
        hx_d[o + s] = hx_s[o + p];
        hy_d[o + s] = hy_s[o + p];
        hz_d[o + s] = hz_s[o + p];
    }
}

template <class T>
__global__ void __launch_bounds__(kThreads) k_gather_by_id(int n, const uint32_t* __restrict__ id, const T* __restrict__ stage,
                                                           T* __restrict__ dst) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) dst[s] = stage[id[s]];
}
template <class T>
__global__ void __launch_bounds__(kThreads) k_scatter_by_id(int n, const uint32_t* __restrict__ id, const T* __restrict__ src,
                                                            T* __restrict__ stage) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) stage[id[s]] = src[s];
}

// Neighbour / contact set dump (parity hook).  mode 0: r2 < (kfac*s_i)^2, s = h.  mode 1: r2 < (s_i+s_j)^2, s = rad.
// Coupled contexts pass `tag`: mode 0 keeps the pairs with a fluid member, mode 1 those with a solid i and a non-fluid j.
template <class R, int DIM, bool MORTON>
__global__ void __launch_bounds__(kThreads) k_dump_pairs(GridDev<R> g, int n, int mode, R kfac, const int32_t* __restrict__ cell_start,
                                                         const R* __restrict__ x, const R* __restrict__ y, const R* __restrict__ z,
                                                         const R* __restrict__ sz, const uint32_t* __restrict__ id,
                                                         const int32_t* __restrict__ tag, const int32_t* __restrict__ body,
                                                         uint32_t* __restrict__ oi, uint32_t* __restrict__ oj,
                                                         unsigned long long cap, unsigned long long* __restrict__ counter) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const R xi = x[s], yi = y[s], zi = DIM == 3 ? z[s] : (R)0, si = sz[s];
    const int cx = cell_coord<R>(xi, g.lo[0], g.inv[0], g.cx_lo, g.cx_hi);
    const int cy = cell_coord<R>(yi, g.lo[1], g.inv[1], 0, g.n[1] - 1);
    const int cz = DIM == 3 ? cell_coord<R>(zi, g.lo[2], g.inv[2], 0, g.n[2] - 1) : 0;
    const uint32_t idi = id[s];
    const int ti = tag ? tag[s] : 0;
    if (tag && mode == 1 && ti != 2) return;
    const int bi = (body && mode == 1) ? body[s] : -1;       // rigid bodies: members of one body are not contact partners
    for_each_run<DIM, MORTON>(g, cell_start, cx, cy, cz, [&](int b, int e) {
        for (int j = b; j < e; ++j) {
            if (j == s) continue;
            if (tag && (mode == 0 ? (ti != 0 && tag[j] != 0) : tag[j] == 0)) continue;
            if (bi >= 0 && body[j] == bi) continue;
            const R r2 = dist2<DIM, R>(xi - x[j], yi - y[j], DIM == 3 ? zi - z[j] : (R)0);
            const R rc = mode == 0 ? mul_rn(kfac, si) : add_rn(si, sz[j]);
            if (r2 < mul_rn(rc, rc)) {
                const unsigned long long k = atomicAdd(counter, 1ull);
                if (k < cap) { oi[k] = idi; oj[k] = id[j]; }
            }
        }
    });
}

template <class R, int DIM, bool MORTON>
pst_status launch_keys(pst_ctx* ctx, int mig_l, int mig_r) {
    const int n = (int)ctx->n;
    PST_LAUNCH(ctx, (k_keys<R, DIM, MORTON>), blocks_for(n), kThreads, 0, make_grid_dev<R>(ctx->grid), n,
               pst_ptr<R>(ctx, "x"), pst_ptr<R>(ctx, "y"), DIM == 3 ? pst_ptr<R>(ctx, "z") : nullptr, ctx->keys_in, ctx->vals_in,
               ctx->grid.ncells, mig_l, mig_r);
    return PST_OK;
}

template <class R, int DIM, bool MORTON>
pst_status launch_keys_count(pst_ctx* ctx, int mig_l, int mig_r) {
    const int n = (int)ctx->n;
    PST_LAUNCH(ctx, (k_keys_count<R, DIM, MORTON>), blocks_for(n), kThreads, 0, make_grid_dev<R>(ctx->grid), n,
               pst_ptr<R>(ctx, "x"), pst_ptr<R>(ctx, "y"), DIM == 3 ? pst_ptr<R>(ctx, "z") : nullptr, ctx->keys_in, ctx->vals_in,
               ctx->cell_start, ctx->grid.ncells, mig_l, mig_r);
    return PST_OK;
}

template <class R, int DIM, bool MORTON>
pst_status launch_dump(pst_ctx* ctx, int mode, const void* sz, uint32_t* oi, uint32_t* oj, size_t cap) {
    const int n = (int)ctx->n;
    PST_LAUNCH(ctx, (k_dump_pairs<R, DIM, MORTON>), blocks_for(n), kThreads, 0, make_grid_dev<R>(ctx->grid), n, mode,
               (R)pst_param(ctx, "kfac", 2.0), ctx->cell_start, pst_ptr<R>(ctx, "x"), pst_ptr<R>(ctx, "y"),
               DIM == 3 ? pst_ptr<R>(ctx, "z") : nullptr, (const R*)sz, pst_ptr<uint32_t>(ctx, "id"),
               ctx->coupled ? pst_ptr<int32_t>(ctx, "tag") : nullptr, ctx->bodies_ready ? pst_ptr<int32_t>(ctx, "body") : nullptr, oi, oj,
               (unsigned long long)cap, ctx->d_counters);
    return PST_OK;
}

template <class R>
pst_status launch_remap(pst_ctx* ctx) {
    PstArray* hn = pst_find(ctx, "hist_n");
    if (!hn) return PST_OK;
    PstArray *hid = pst_find(ctx, "hist_id"), *hx = pst_find(ctx, "hist_x"), *hy = pst_find(ctx, "hist_y"), *hz = pst_find(ctx, "hist_z");
    const int n = (int)ctx->n;
    const size_t stride = ctx->capacity + 2 * ctx->ghost_cap;
    const int c = hn->cur, d = 1 - hn->cur;
    PST_LAUNCH(ctx, (k_remap_history<R>), blocks_for(n), kThreads, 0, n, stride, ctx->vals_out,
               pst_ptr<int32_t>(ctx, hn, 0, c), pst_ptr<uint32_t>(ctx, hid, 0, c), pst_ptr<R>(ctx, hx, 0, c),
               pst_ptr<R>(ctx, hy, 0, c), pst_ptr<R>(ctx, hz, 0, c), pst_ptr<int32_t>(ctx, hn, 0, d),
               pst_ptr<uint32_t>(ctx, hid, 0, d), pst_ptr<R>(ctx, hx, 0, d), pst_ptr<R>(ctx, hy, 0, d), pst_ptr<R>(ctx, hz, 0, d));
    for (PstArray* a : {hn, hid, hx, hy, hz}) a->cur = d;
    return PST_OK;
}

bool is_history(const PstArray& a) { return a.name.rfind("hist_", 0) == 0; }

}  // namespace

pst_status pst_nnps_alloc(pst_ctx* ctx) {
    const size_t cap = ctx->capacity + 2 * ctx->ghost_cap;
    PST_CUDA(ctx, cudaMalloc((void**)&ctx->keys_in, cap * 4));
    PST_CUDA(ctx, cudaMalloc((void**)&ctx->keys_out, cap * 4));
    PST_CUDA(ctx, cudaMalloc((void**)&ctx->vals_in, cap * 4));
    PST_CUDA(ctx, cudaMalloc((void**)&ctx->vals_out, cap * 4));
    PST_CUDA(ctx, cudaMalloc((void**)&ctx->stage, cap * 8));
    PST_CUDA(ctx, cudaMalloc((void**)&ctx->d_flags, 8 * sizeof(int32_t)));
    PST_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, 8 * sizeof(int32_t), ctx->stream));
    PST_CUDA(ctx, cudaMalloc((void**)&ctx->d_counters, 8 * sizeof(unsigned long long)));
    PST_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
    PST_CUDA(ctx, cudaHostAlloc((void**)&ctx->h_flags, 8 * sizeof(int32_t), cudaHostAllocDefault));
    PST_CUDA(ctx, cudaHostAlloc((void**)&ctx->h_counters, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
    PST_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_stats, cudaEventDisableTiming));
    ctx->sort_tmp_bytes = 0;
    PST_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, ctx->sort_tmp_bytes, ctx->keys_in, ctx->keys_out, ctx->vals_in,
                                                  ctx->vals_out, (int)cap, 0, 32, ctx->stream));
    PST_CUDA(ctx, cudaMalloc(&ctx->sort_tmp, ctx->sort_tmp_bytes));
    return PST_OK;
}

// cell table (ncells + 1 entries, + the two migration sentinels) and the tile sums of its scan, for the current grid
pst_status pst_nnps_alloc_table(pst_ctx* ctx) {
    if (ctx->stream) PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->cell_start); ctx->cell_start = nullptr;
    cudaFree(ctx->scan_sums); ctx->scan_sums = nullptr;
    const size_t entries = (size_t)ctx->grid.ncells + 4;
    PST_CUDA(ctx, cudaMalloc((void**)&ctx->cell_start, entries * 4));
    PST_CUDA(ctx, cudaMemsetAsync(ctx->cell_start, 0, entries * 4, ctx->stream));
    ctx->scan_sums_cap = entries / kScanTile + 2;
    PST_CUDA(ctx, cudaMalloc((void**)&ctx->scan_sums, ctx->scan_sums_cap * 4));
    return PST_OK;
}

pst_status pst_iota_ids(pst_ctx* ctx) {
    if (ctx->n == 0) return PST_OK;
    PST_LAUNCH(ctx, k_iota, blocks_for(ctx->n), kThreads, 0, pst_ptr<uint32_t>(ctx, "id"), (int)ctx->n);
    return PST_OK;
}

// one pass of keys -> sort -> cell table -> permute (+ history remap).  With mig_l / mig_r the leavers end up behind the
// stayers: cell_start[ncells] = n_stay, cell_start[ncells + 1] = n_stay + n_left, cell_start[ncells + 2] = n.
static pst_status build_pass(pst_ctx* ctx, int mig_l, int mig_r) {
    PST_TRY(pst_resolve_history(ctx));   // vals_out of the previous sort is about to be overwritten
    const int n = (int)ctx->n;
    const bool mig = mig_l || mig_r;
    const uint32_t nkeys = ctx->grid.ncells + (mig ? 2u : 0u);
    int key_bits = ctx->grid.key_bits;
    while (mig && (1ull << key_bits) < (unsigned long long)nkeys) ++key_bits;
    ctx->n_ghost_l = ctx->n_ghost_r = 0;
    ctx->ghost_exact = true;
    // occupied-cell count of the PREVIOUS build (read back asynchronously; the very first build waits once)
    if (ctx->stats_pending && !ctx->capturing) {
        PST_CUDA(ctx, cudaEventSynchronize(ctx->ev_stats));
        ctx->stats_pending = false;
        if (ctx->h_counters[2] > 0) ctx->params["_ppc"] = (double)ctx->h_counters[3] / (double)ctx->h_counters[2];
    }
    PST_CUDA(ctx, cudaMemsetAsync(ctx->d_counters + 2, 0, 3 * sizeof(unsigned long long), ctx->stream));
    if (pst_option(ctx, "sort_impl", 1) == 1) {
        // ---- hand-written counting sort by cell key; the scanned counts ARE the cell table
        const int m = (int)nkeys + 1;                       // entries [0, nkeys]: the last one becomes n
        PST_CUDA(ctx, cudaMemsetAsync(ctx->cell_start, 0, ((size_t)nkeys + 2) * 4, ctx->stream));
        if (n > 0) PST_TRY(PST_DISPATCH(ctx, launch_keys_count, ctx, mig_l, mig_r));
        const int nb = (m + kScanTile - 1) / kScanTile;
        PST_LAUNCH(ctx, k_scan_tiles, nb, kScanThreads, 0, m, ctx->cell_start, ctx->scan_sums, ctx->d_counters, ctx->grid.sub);
        if (nb > 1) {
            PST_LAUNCH(ctx, k_scan_sums, 1, 1024, 0, nb, ctx->scan_sums);
            PST_LAUNCH(ctx, k_scan_add, nb, kScanThreads, 0, m, ctx->cell_start, ctx->scan_sums);
        }
        if (n > 0) {
            PST_LAUNCH(ctx, k_place, blocks_for(n), kThreads, 0, n, ctx->keys_in, ctx->vals_in, ctx->cell_start, ctx->vals_out);
            PST_LAUNCH(ctx, k_cell_order, blocks_for(nkeys), kThreads, 0, nkeys, ctx->cell_start, ctx->vals_out, ctx->keys_out, ctx->d_counters);
            PST_LAUNCH(ctx, k_cell_order_big, 296, 256, 0, ctx->cell_start, ctx->vals_out, ctx->vals_in, ctx->big_list(), ctx->d_counters);
        }
    } else {
        // ---- library path (A/B): CUB onesweep radix sort on the key bits in use + k_bounds
        if (n > 0) {
            PST_TRY(PST_DISPATCH(ctx, launch_keys, ctx, mig_l, mig_r));
            size_t tmp = ctx->sort_tmp_bytes;
            PST_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->sort_tmp, tmp, ctx->keys_in, ctx->keys_out, ctx->vals_in, ctx->vals_out,
                                                          n, 0, key_bits, ctx->stream));
            ctx->launches += (key_bits + 7) / 8 + 2;  // CUB onesweep: histogram + scan + one pass per 8 bits
        }
        PST_LAUNCH(ctx, k_bounds, blocks_for((size_t)n + 1), kThreads, 0, n, nkeys, ctx->keys_out, ctx->cell_start, ctx->d_counters);
    }
    if (!ctx->capturing) {     // (a captured step keeps the occupancy figure of the steps before it)
        PST_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters + 2, ctx->d_counters + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        PST_CUDA(ctx, cudaEventRecord(ctx->ev_stats, ctx->stream));
        ctx->h_counters[3] = (unsigned long long)n;
        ctx->stats_pending = true;
    }
    if (!ctx->capturing && ctx->params.find("_ppc") == ctx->params.end()) {   // first build: one-time wait so the first force pass is tuned too
        PST_CUDA(ctx, cudaEventSynchronize(ctx->ev_stats));
        ctx->stats_pending = false;
        if (ctx->h_counters[2] > 0) ctx->params["_ppc"] = (double)n / (double)ctx->h_counters[2];
    }
    if (n > 0) {
        PermuteList L;
        L.n8 = L.n4 = 0;
        std::vector<PstArray*> moved;
        const bool fused = pst_wcsph_fused_permute(ctx);      // x y z u v w rho m h travel through k_permute_eos (wcsph.cu) instead
        auto in_fused = [&](const PstArray& a) {
            if (!fused || a.rows != 1) return false;
            for (const char* nm : {"x", "y", "z", "u", "v", "w", "rho", "m", "h"}) if (a.name == nm) return true;
            if ((a.name == "tag" || a.name == "id") && a.esize == 4) return true;      // ride along in k_permute_eos
            return false;
        };
        for (int pass = 0; pass < 2; ++pass)
            for (auto& a : ctx->arrays) {
                if (!(a.flags & PST_ARRAY_PERSISTENT) || is_history(a)) continue;
                if ((pass == 0) != (a.esize == 8)) continue;
                if (in_fused(a)) { if (pass == 0 || a.esize != 8) moved.push_back(&a); continue; }
                for (int r = 0; r < a.rows; ++r) {
                    const int k = L.n8 + L.n4;
                    if (k >= kMaxPermute) return pst_fail(ctx, PST_EINVAL, "too many persistent arrays (max %d)", kMaxPermute);
                    L.src[k] = pst_ptr<char>(ctx, &a, r, a.cur);
                    L.dst[k] = pst_ptr<char>(ctx, &a, r, 1 - a.cur);
                    (pass == 0 ? L.n8 : L.n4)++;
                }
                moved.push_back(&a);
            }
        if (L.n8 + L.n4 > 0) PST_LAUNCH(ctx, k_permute, blocks_for(n), kThreads, 0, L, n, ctx->vals_out);
        if (fused) PST_TRY(pst_wcsph_permute_eos(ctx, ctx->vals_out, n));
        for (PstArray* a : moved) a->cur = 1 - a->cur;
        // Contact history: the remap is DEFERRED -- the next contact pass reads each row through vals_out (new -> old
        // index) and writes it back in place of a separate 2 x (28 Z + 4) B/particle copy.  Anything else that needs the
        // rows in the new order (a second re-sort, a download, a migration that moves particles) resolves the lag first.
        if (pst_find(ctx, "hist_n")) ctx->hist_lag = true;   // (migration resolves it first, and only when particles actually move)
    }
    ctx->ordered = true;
    ctx->nbrs_valid = true;
    ctx->state_epoch++;
    ctx->build_epoch++;
    ctx->ghost_eos_pending = false;                              // (no ghost rows until the next halo exchange)
    ctx->eos_valid = n > 0 && pst_wcsph_fused_permute(ctx);      // the fused permute has evaluated the EOS and written the records
    if (ctx->eos_valid) ctx->rec_epoch = ctx->state_epoch;
    return PST_OK;
}

// stable sort of n (key, value) u32 pairs: keys_in/vals_in -> keys_out/vals_out (library sort; used by one-time set-up
// work such as the rigid-body member lists, never on the per-step path, whose sort is the counting sort above)
pst_status pst_sort_pairs_u32(pst_ctx* ctx, int n) {
    if (n <= 0) return PST_OK;
    size_t tmp = ctx->sort_tmp_bytes;
    PST_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->sort_tmp, tmp, ctx->keys_in, ctx->keys_out, ctx->vals_in, ctx->vals_out, n, 0, 32, ctx->stream));
    ctx->launches += 6;
    return PST_OK;
}

// exclusive scan of m ints, in place (the three scan kernels of the counting sort; scan_sums must hold m / kScanTile + 1 entries)
pst_status pst_scan_exclusive(pst_ctx* ctx, int32_t* a, int m) {
    const int nb = (m + kScanTile - 1) / kScanTile;
    if ((size_t)nb + 1 > ctx->scan_sums_cap) {
        PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        PST_CUDA(ctx, cudaFree(ctx->scan_sums));
        ctx->scan_sums = nullptr;
        ctx->scan_sums_cap = (size_t)nb + 2;
        PST_CUDA(ctx, cudaMalloc((void**)&ctx->scan_sums, ctx->scan_sums_cap * 4));
    }
    PST_LAUNCH(ctx, k_scan_tiles, nb, kScanThreads, 0, m, a, ctx->scan_sums, (unsigned long long*)nullptr, 1);
    if (nb > 1) {       // (one tile: its local scan is the scan)
        PST_LAUNCH(ctx, k_scan_sums, 1, 1024, 0, nb, ctx->scan_sums);
        PST_LAUNCH(ctx, k_scan_add, nb, kScanThreads, 0, m, a, ctx->scan_sums);
    }
    return PST_OK;
}

pst_status pst_resolve_history(pst_ctx* ctx) {
    if (!ctx->hist_lag) return PST_OK;
    ctx->hist_lag = false;
    if (ctx->n == 0) return PST_OK;
    return ctx->f64 ? launch_remap<double>(ctx) : launch_remap<float>(ctx);
}

pst_status pst_nnps_build(pst_ctx* ctx) {
    if (!pst_find(ctx, "x")) return pst_fail(ctx, PST_ESTATE, "context has no position arrays (physics = PST_PHYS_NONE)");
    if (!ctx->comm) return build_pass(ctx, 0, 0);
    // slab decomposition: park the particles that left the slab behind the stayers, hand them to the neighbour ranks,
    // take theirs in, and re-sort only if somebody arrived
    int mig_l = 0, mig_r = 0;
    pst_comm_neighbours(ctx, &mig_l, &mig_r);
    PST_TRY(build_pass(ctx, mig_l, mig_r));
    int arrivals = 0;
    PST_TRY(pst_migrate(ctx, &arrivals));
    if (arrivals > 0) PST_TRY(build_pass(ctx, 0, 0));
    return PST_OK;
}

pst_status pst_reorder_upload(pst_ctx* ctx, PstArray* a, int row, size_t n) {
    const uint32_t* id = pst_ptr<uint32_t>(ctx, "id");
    if (a->esize == 8)
        PST_LAUNCH(ctx, k_gather_by_id<unsigned long long>, blocks_for(n), kThreads, 0, (int)n, id,
                   (const unsigned long long*)ctx->stage, pst_ptr<unsigned long long>(ctx, a, row));
    else
        PST_LAUNCH(ctx, k_gather_by_id<uint32_t>, blocks_for(n), kThreads, 0, (int)n, id, (const uint32_t*)ctx->stage,
                   pst_ptr<uint32_t>(ctx, a, row));
    return PST_OK;
}

pst_status pst_reorder_download(pst_ctx* ctx, PstArray* a, int row, size_t n) {
    const uint32_t* id = pst_ptr<uint32_t>(ctx, "id");
    if (a->esize == 8)
        PST_LAUNCH(ctx, k_scatter_by_id<unsigned long long>, blocks_for(n), kThreads, 0, (int)n, id,
                   pst_ptr<unsigned long long>(ctx, a, row), (unsigned long long*)ctx->stage);
    else
        PST_LAUNCH(ctx, k_scatter_by_id<uint32_t>, blocks_for(n), kThreads, 0, (int)n, id, pst_ptr<uint32_t>(ctx, a, row),
                   (uint32_t*)ctx->stage);
    return PST_OK;
}

pst_status pst_nnps_dump_pairs(pst_ctx* ctx, int mode, uint32_t* hi, uint32_t* hj, size_t cap, size_t* n_pairs) {
    if (mode != 0 && mode != 1) return pst_fail(ctx, PST_EINVAL, "dump_pairs mode must be 0 (neighbours) or 1 (contacts)");
    PstArray* sa = pst_find(ctx, mode == 0 ? "h" : "rad");
    if (!sa) return pst_fail(ctx, PST_ESTATE, "dump_pairs mode %d needs array '%s'", mode, mode == 0 ? "h" : "rad");
    uint32_t *di = nullptr, *dj = nullptr;
    if (cap) {
        if (cudaMalloc((void**)&di, cap * 4) != cudaSuccess || cudaMalloc((void**)&dj, cap * 4) != cudaSuccess) {
            cudaFree(di);
            return pst_fail(ctx, PST_ENOMEM, "pair buffer of %zu entries does not fit on the device", cap);
        }
    }
    pst_status s = PST_OK;
    cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned long long), ctx->stream);
    if (ctx->n > 0) {
        s = PST_DISPATCH(ctx, launch_dump, ctx, mode, pst_ptr<char>(ctx, sa), di, dj, cap);
    }
    if (s == PST_OK) {
        cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) s = pst_fail(ctx, PST_ECUDA, "dump_pairs: %s", cudaGetErrorString(e));
    }
    if (s == PST_OK) {
        const size_t cnt = (size_t)ctx->h_counters[0];
        *n_pairs = cnt;
        const size_t m = cnt < cap ? cnt : cap;
        if (m) {
            cudaMemcpy(hi, di, m * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hj, dj, m * 4, cudaMemcpyDeviceToHost);
        }
        if (cnt > cap) s = pst_fail(ctx, PST_EOVERFLOW, "%zu pairs > buffer capacity %zu", cnt, cap);
    }
    cudaFree(di);
    cudaFree(dj);
    return s;
}
