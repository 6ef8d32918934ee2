// wcsph_zrun.cuh -- variant 3 of the fused continuity + momentum pair kernel (option force_kernel = 3).  Included by
// wcsph.cu inside its anonymous namespace (it shares ForceArgs, gather_one, TileDims, the packed records with the rest).
// No reference code exists for the physics; the loop shape it honours is the reference's
// gather -- write only [i], bodies of a fused set in ONE i,j loop (prestige/src/codegen/simple_cpu.rs:7-16, fuse.rs:14-40).
//
// The design in one place (DESIGN.md section 4 has the measurements; profiles/r2_exp_log.txt every step on the way):
//   * the cell grid is `sub` times finer along the FAST axis (option zsub; nnps.cu sorts by the fine key), so inside a
//     stencil column the particles are ordered by fine z.  A particle scans, per column, only the fine cells within
//     +-sqrt(rc^2 - d_xy^2) of its own z, d_xy = its distance to that column's footprint: ~190 candidates instead of the
//     373 of the 27-cell stencil;
//   * the scan produces BIT MASKS, not lists: d = dx^2 + dy^2 + dz^2 - rc^2 comes out of three packed FFMA2, its sign bit
//     is funnel-shifted into a 32-bit word (one SHF per candidate, no compare, no predicated store, no serial list
//     pointer); non-empty words are kept as 8 bytes (mask, global index of the word's first candidate + 31);
//   * all lanes of a warp scan a run in lock-step (trip count = the warp's longest range, the surplus bits are cut off),
//     every LDS.128 is 16-byte aligned by construction (runs are staged at multiples of 4, ranges start aligned down);
//   * staging holds 12 B per candidate: f32 coordinates relative to the GRID origin, copied by TMA (cp.async.bulk +
//     mbarrier) out of the f32 position rows that are written with the packed records; without records (rec_impl = 0) the
//     warps convert tile-local coordinates on the fly;
//   * phase 2 is software-pipelined and branch-free: a count-leading-zeros iterator over the words (one predicated LDS.64
//     per refill) runs one hit ahead of the gathers, the gathers -- 4 x LDG.128 off one address into the packed AoSoA
//     records (wcsph.cu rec_*) -- one body ahead of their use; the exact FMA-free f64 test on the gathered record decides
//     membership, so the neighbour set stays bit-exact;
//   * tiles are cut on the device from the cell table (k_ztile_list: 2 x 2 columns x as many fine cells as hold <= 2 NT own
//     particles and <= 2560 staged candidates; list in (tx, ty, f0) order so neighbouring tiles meet in L2);
//   * CTAs are PERSISTENT (2 per SM), fetch tiles from an atomic counter, and every warp evaluates two rows of 32 particles
//     per tile; ONE staging buffer (TMA makes the refill cheap; two buffers of half the size measured slower).
// The unit this kernel sits on is the L1 data pipe (90 % busy: the 64-byte gathers cost ~36 wavefronts per warp whatever the
// layout), not HBM and not FP64.

struct ZTile {
    int gfcap;       // deepest tile, in fine cells
    int wmax;        // fine-cell boundaries per staged run at that depth: gfcap + 2 sub + 1
    int jcap;        // staged-candidate capacity (<= kZJcap; tests shrink it to force the fallback)
    int maxw;        // mask words per thread
    const int4* tiles;   // (tx, ty, f0, gf) per tile
    const int* ntiles;   // device-side count
    int* next;           // dynamic tile counter (starts at 2 x grid: the first two tiles of every CTA are static)
    float marg;          // TMA staging (global f32 coordinates): relative margin of the pre-filter on rc^2, from the box size
    float zslack;        // ... and the slack, in fine cells, of the z-range look-up
};

constexpr int kZPad = 64;      // readable slack behind the staged rows: lanes with a short range over-scan with the warp
// Tiles of up to kZRounds x NT particles (every warp evaluates kZRounds rows of 32 per tile) with ONE staging buffer: fewer tiles,
// fewer barriers and table set-ups, and a better ratio of staged candidates to own particles (4.8 instead of 5.7) than
// NT-particle tiles with two buffers in the same shared memory: 5.04 -> 4.92 ms at 10 M, L1 hit rate unchanged at 90 %
// (profiles/r2_exp_log.txt).
constexpr int kZRounds = 2, kZNbuf = 1;
template <int NT> constexpr int zjcap() { return NT >= 256 ? 2560 : 1024; }   // staged candidates per buffer (12 B each)

__global__ void k_set_int(int* p, int v) { *p = v; }

// ---- tile list: one warp per tile column pair (tx, ty); greedy cut along the fast axis so that a tile holds <= target particles
// of its own and its staged runs (own columns + one column around, `sub` fine cells above and below) hold <= jcap candidates
template <int DIM, int TA, int TB>
__global__ void __launch_bounds__(128) k_ztile_list(int ncx, int ncy, int nf, int cx_lo, int cx_hi, int tiles_x, int tiles_y, int target, int gfcap,
                                                    int sub, int jcap, const int32_t* __restrict__ cs, int4* __restrict__ slots, int* __restrict__ off) {
    constexpr int BB = DIM == 3 ? TB : 1;
    constexpr int NRX = TA + 2, NRY = DIM == 3 ? BB + 2 : 1;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= tiles_x * tiles_y) return;
    const int tx = wid / tiles_y, ty = wid - tx * tiles_y;
    // the owned columns of this tile column pair
    size_t colbase[TA * BB];
    bool colok[TA * BB];
#pragma unroll
    for (int c = 0; c < TA * BB; ++c) {
        const int lx = c / BB, ly = c - lx * BB;
        const int cx = tx * TA + lx, cy = ty * BB + ly;
        colok[c] = cx >= cx_lo && cx <= cx_hi && cx < ncx && cy < ncy;
        colbase[c] = (size_t)(DIM == 3 ? cx * ncy + cy : cx) * nf;
    }
    auto C = [&](int f) {     // particles of the owned columns below fine boundary f
        int t = 0;
#pragma unroll
        for (int c = 0; c < TA * BB; ++c)
            if (colok[c]) t += cs[colbase[c] + f];
        return t;
    };
    auto Sg = [&](int f) {    // particles of the staged columns below fine boundary f
        int t = 0;
        for (int rx = 0; rx < NRX; ++rx)
            for (int ry = 0; ry < NRY; ++ry) {
                const int cx = tx * TA - 1 + rx, cy = DIM == 3 ? ty * BB - 1 + ry : 0;
                if (cx >= 0 && cx < ncx && cy >= 0 && cy < ncy) t += cs[(size_t)(DIM == 3 ? cx * ncy + cy : cx) * nf + f];
            }
        return t;
    };
    int f0 = 0;
    int c0 = C(0);
    const int cend = C(nf);
    const int pad = 3 * NRX * NRY;          // alignment slack of the staged runs
    // ONE pass of the cut: the tiles of this column pair go to its own slot range (stride nf: a tile holds at least one fine cell)
    // and their number to off[wid]; after an exclusive scan of the counts k_ztile_compact moves them to their final places, so the
    // list is in (tx, ty, f0) order: the CTAs then walk the domain like a launch in that order would, and the candidates two
    // neighbouring tile columns share are still in L2 when the second one needs them (an unordered list reads every record
    // ~3x from DRAM: 2.8 GB instead of 0.9 GB at 10 M particles, profiles/r2_exp_log.txt)
    int nt = 0;
    int4* const mine = slots + (size_t)wid * nf;
    while (f0 < nf && c0 < cend) {          // (c0 == cend: nothing above f0)
        int best = f0 + 1;                  // at least one fine cell per tile (a cell denser than the target runs in several rounds)
        const int s0 = Sg(max(f0 - sub, 0));
        for (int off = 0; off < gfcap; off += 32) {
            const int fe = f0 + 1 + off + lane;
            const bool valid = fe <= nf && fe - f0 <= gfcap;
            const int cnt = valid ? C(fe) - c0 : 0x7fffffff;
            const int staged = valid ? Sg(min(fe + sub, nf)) - s0 + pad : 0x7fffffff;
            const unsigned m = __ballot_sync(0xffffffffu, valid && cnt <= target && staged <= jcap);
            if (m) best = max(best, f0 + off + 32 - __clz(m));     // counts are monotone: the set lanes form a prefix
            if (m != 0xffffffffu) break;
        }
        const int cb = C(best);
        if (cb > c0) {
            if (lane == 0) mine[nt] = make_int4(tx, ty, f0, best - f0);
            ++nt;
        }
        f0 = best; c0 = cb;
    }
    if (lane == 0) off[wid] = nt;
}

// one warp per column pair: its tiles from the slot range to their place in the ordered list (off = the scanned counts)
__global__ void __launch_bounds__(128) k_ztile_compact(int npairs, int nf, const int4* __restrict__ slots, const int* __restrict__ off, int4* __restrict__ tiles) {
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= npairs) return;
    const int b = off[wid], e = off[wid + 1];
    for (int i = lane; i < e - b; i += 32) tiles[b + i] = slots[(size_t)wid * nf + i];
}

// ---- TMA (1-D bulk copy) + mbarrier, sm_90+ PTX: the staged rows of a tile are contiguous ranges of the f32 position arrays
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "W_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// predicated 8-byte shared-memory load (no branch, no wavefront for lanes whose predicate is off): a, b keep their values when p is false
__device__ __forceinline__ void lds_u64_if(unsigned saddr, bool p, unsigned& a, unsigned& b) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q ld.shared.v2.u32 {%0, %1}, [%2];\n\t}" : "+r"(a), "+r"(b) : "r"(saddr), "r"((unsigned)p) : "memory");
}

template <class R, int DIM, int TA, int TB, int NT, bool CONT, bool MOM, bool COUPLED = false, bool UNI = false, bool REC = false, int DBG = 0, int NBUF = 2, bool BETA0 = false>
__global__ void __launch_bounds__(NT, 512 / NT) k_wcsph_zrun(GridDev<R> g, WcsphConst<R> C, ForceArgs<R> A, ZTile T) {
    using D = TileDims<DIM, TA, TB>;
    using P2 = typename RecPair<R>::type;
    constexpr int NR = D::NR, NI = D::NI, RY = D::RY, BB = D::BB;
    constexpr int NW = NT / 32;
    constexpr int NRUN = DIM == 3 ? 9 : 3;
    constexpr int FAST = DIM - 1;
    constexpr int kZJcap = zjcap<NT>(), kZRow = kZJcap + kZPad;
    constexpr int NTAB = NR + NR + 1 + NI + NI + 1 + 4 + NW;     // per buffer: gbeg, voff, ibeg, ipre, meta (tx ty f0 gf), far flag per warp
    static_assert(NR <= 32 && NI <= 32, "one warp scans the run and column tables");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = g.sub, WM = T.wmax;
    // ---- shared memory carve-up: two buffers of (tables, boundary table, staged rows), then the mask words
    const int tab_ints = (NTAB + NR * WM + 3) & ~3;
    int* const s_tab0 = reinterpret_cast<int*>(smem_raw);
    float* const s_row0 = reinterpret_cast<float*>(smem_raw + (size_t)NBUF * tab_ints * sizeof(int));
    constexpr int kRows = DIM == 3 ? 3 : 2;
    uint2* const s_word = reinterpret_cast<uint2*>(s_row0 + NBUF * kRows * kZRow);   // maxw * NT: (mask, global index of the word's first candidate + 31)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nf = g.n[FAST];                       // fine cells along the fast axis
    const int ncx = g.n[0], ncy = DIM == 3 ? g.n[1] : 1;
    const int ntiles = *T.ntiles;
    const R cellR = g.cell;
    const float cellf = (float)g.cell;
    const float inv_cf = (float)S / cellf;          // 1 / fine cell edge
    const int MAXW = T.maxw;
    uint2* const my_word = s_word + tid;
    // TMA staging needs the f32 position rows that are written with the packed records
    constexpr bool TMA = REC;
    __shared__ __align__(8) uint64_t s_bar;
    if (TMA) {
        if (tid == 0) mbar_init(&s_bar, 1);
        __syncthreads();          // the barrier object is initialised before anybody arrives on it or waits for it
    }
    // coordinates of the scan: TMA: relative to the grid origin (what the f32 rows hold); else: relative to the tile origin
    const float marg = TMA ? T.marg : 1.0f / 32768.0f;

    // ---- stage tile k into buffer b: every warp derives the run table itself (two look-ups per lane + a shuffle scan),
    // then copies its own runs (boundaries + candidates).  No block-wide barrier inside.
    auto stage_tile = [&](int k, int b) {
        int* const tab = s_tab0 + b * tab_ints;
        int* const s_gbeg = tab; int* const s_voff = s_gbeg + NR; int* const s_ibeg = s_voff + NR + 1; int* const s_ipre = s_ibeg + NI;
        int* const s_meta = s_ipre + NI + 1; int* const s_far = s_meta + 4; int* const s_cs = s_far + NW;
        float* const s_x = s_row0 + b * kRows * kZRow;
        const int4 t = T.tiles[k];
        const int cx0 = t.x * TA, cy0 = t.y * BB, f0 = t.z, gf = t.w, W = gf + 2 * S + 1;
        // run extents (lane q < NR)
        int gb = 0, len = 0;
        long long colq = 0;
        bool okq = false;
        if (lane < NR) {
            const int rx = lane / RY, ry = lane - rx * RY;
            const int cx = cx0 - 1 + rx, cy = DIM == 3 ? cy0 - 1 + ry : 0;
            okq = cx >= 0 && cx < ncx && cy >= 0 && cy < ncy;
            colq = (long long)(DIM == 3 ? cx * ncy + cy : cx) * nf;
            if (okq) { gb = A.cell_start[colq + max(f0 - S, 0)]; len = A.cell_start[colq + min(f0 + gf + S, nf)] - gb; }
        }
        // every run starts 16-byte aligned in the staged rows.  TMA: the SOURCE must be aligned too, so the run is copied from the
        // multiple of 4 below its first particle (the head and tail slack hold neighbours of the range, never scanned)
        const int gb4 = TMA ? gb & ~3 : gb;
        const int alen = len > 0 ? (gb - gb4 + len + 3) & ~3 : 0;
        int incl = alen;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        const int voff = incl - alen;
        const int M = __shfl_sync(0xffffffffu, incl, NR - 1);
        const bool dense = M > min(kZJcap, T.jcap);
        if (warp == 0) {
            if (lane < NR) { s_gbeg[lane] = gb4 - voff; s_voff[lane] = voff; }      // s_gbeg: global index minus staged index of the run
            if (lane == NR - 1) s_voff[NR] = M;
            // i segments: the tile's own columns, fine cells [f0, f0 + gf)
            int cnt = 0, beg = 0;
            if (lane < NI) {
                const int lx = lane / BB, ly = lane - lx * BB;
                const int cx = cx0 + lx, cy = cy0 + ly;
                if (cx >= g.cx_lo && cx <= g.cx_hi && cx < ncx && cy < ncy) {       // ghost layers are never i
                    const long long col = (long long)(DIM == 3 ? cx * ncy + cy : cx) * nf;
                    beg = A.cell_start[col + f0]; cnt = A.cell_start[col + f0 + gf] - beg;
                }
            }
            int ipre = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, ipre, d);
                if (lane >= d) ipre += y;
            }
            if (lane < NI) { s_ibeg[lane] = beg; s_ipre[lane] = ipre - cnt; }
            if (lane == NI - 1) s_ipre[NI] = ipre;
            if (lane == 0) { s_meta[0] = t.x; s_meta[1] = t.y; s_meta[2] = f0; s_meta[3] = dense ? -gf : gf; }
            if (TMA) {
                // one arrival + the byte count, then every run's rows as bulk copies straight into the staged rows
                fence_proxy_async();                 // the rows were read through the generic proxy by the previous tile
                if (lane == 0) mbar_expect_tx(&s_bar, dense ? 0u : (unsigned)(kRows * M * (int)sizeof(float)));
                __syncwarp();
                if (!dense && lane < NR && alen > 0) {
                    tma_load_1d(s_x + voff, A.F.x + gb4, (unsigned)alen * 4u, &s_bar);
                    tma_load_1d(s_x + kZRow + voff, A.F.y + gb4, (unsigned)alen * 4u, &s_bar);
                    if (DIM == 3) tma_load_1d(s_x + 2 * kZRow + voff, A.F.z + gb4, (unsigned)alen * 4u, &s_bar);
                }
            }
        }
        bool far = false;
        if (TMA) {
            if (!dense)
                for (int q = warp; q < NR; q += NW) {      // staged offset of every fine-cell boundary of the run
                    const int gbq = __shfl_sync(0xffffffffu, gb4, q), vo = __shfl_sync(0xffffffffu, voff, q);
                    const bool ok = __shfl_sync(0xffffffffu, (int)okq, q) != 0;
                    const long long col = __shfl_sync(0xffffffffu, colq, q);
                    for (int tt = lane; tt < W; tt += 32)
                        s_cs[q * WM + tt] = vo + (ok ? A.cell_start[col + min(max(f0 - S + tt, 0), nf)] - gbq : 0);
                }
        } else if (!dense) {
            // tile-local coordinates: origin = low corner of the tile's own cells (no global load needed)
            const R ox = g.lo[0] + (R)cx0 * cellR;
            const R oy = DIM == 3 ? g.lo[1] + (R)cy0 * cellR : g.lo[1] + (R)f0 * (cellR / (R)S);
            const R oz = DIM == 3 ? g.lo[2] + (R)f0 * (cellR / (R)S) : (R)0;
            const float far_lim = (float)(gf / S + 6) * cellf;
            for (int q = warp; q < NR; q += NW) {
                const int gbq = __shfl_sync(0xffffffffu, gb, q), vo = __shfl_sync(0xffffffffu, voff, q);
                const int lenq = __shfl_sync(0xffffffffu, len, q), plen = __shfl_sync(0xffffffffu, alen, q);
                const bool ok = __shfl_sync(0xffffffffu, (int)okq, q) != 0;
                const long long col = __shfl_sync(0xffffffffu, colq, q);
                // staged offset of every fine-cell boundary of the run
                for (int tt = lane; tt < W; tt += 32)
                    s_cs[q * WM + tt] = vo + (ok ? A.cell_start[col + min(max(f0 - S + tt, 0), nf)] - gbq : 0);
                // candidates, coalesced; the slack up to the next multiple of 4 holds a far-away dummy
                for (int v = lane; v < plen; v += 32) {
                    float px = 1e30f, py = 0.0f, pz = 0.0f;    // dummy: never within any cutoff
                    if (v < lenq) {
                        const int gj = gbq + v;
                        px = (float)(A.x[gj] - ox); py = (float)(A.y[gj] - oy); pz = DIM == 3 ? (float)(A.z[gj] - oz) : 0.0f;
                        far |= !(fabsf(px) <= far_lim && fabsf(py) <= far_lim && fabsf(pz) <= far_lim);
                    }
                    s_x[vo + v] = px; s_x[kZRow + vo + v] = py; if (DIM == 3) s_x[2 * kZRow + vo + v] = pz;
                }
            }
            if (warp == NW - 1)
                for (int v = lane; v < kZPad; v += 32) { s_x[M + v] = 1e30f; s_x[kZRow + M + v] = 0.0f; if (DIM == 3) s_x[2 * kZRow + M + v] = 0.0f; }
        }
        far = __any_sync(0xffffffffu, far);
        if (lane == 0) s_far[warp] = far;
    };

    // ---- evaluate the staged tile of buffer b
    auto compute_tile = [&](int b) {
        const int* const tab = s_tab0 + b * tab_ints;
        const int* const s_gbeg = tab; const int* const s_voff = s_gbeg + NR; const int* const s_ibeg = s_voff + NR + 1; const int* const s_ipre = s_ibeg + NI;
        const int* const s_meta = s_ipre + NI + 1; const int* const s_far = s_meta + 4; const int* const s_cs = s_far + NW;
        const float* const s_x = s_row0 + b * kRows * kZRow;
        const int ni = s_ipre[NI];
        if (ni == 0) return;
        const int cx0 = s_meta[0] * TA, cy0 = s_meta[1] * BB, f0 = s_meta[2];
        bool fallback = s_meta[3] < 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) fallback |= s_far[w] != 0;
        if (fallback) {
            // tile denser than the staging buffer, or holding particles far outside the box (clamped into its cells): the exact
            // per-particle gather for its particles
            for (int ii = tid; ii < ni; ii += NT) {
                int c = 0;
                while (c + 1 < NI && ii >= s_ipre[c + 1]) ++c;
                gather_one<R, DIM, false, CONT, MOM, COUPLED>(g, C, A, s_ibeg[c] + (ii - s_ipre[c]));
            }
            return;
        }
        const int gf = s_meta[3], W = gf + 2 * S + 1;
        const R ox = g.lo[0] + (R)cx0 * cellR;
        const R oy = DIM == 3 ? g.lo[1] + (R)cy0 * cellR : g.lo[1] + (R)f0 * (cellR / (R)S);
        const R oz = DIM == 3 ? g.lo[2] + (R)f0 * (cellR / (R)S) : (R)0;
        // origin of the scan coordinates in cells: the grid's (TMA) or the tile's
        const int cxo = TMA ? cx0 : 0, cyo = TMA ? (DIM == 3 ? cy0 : 0) : 0, fzo = TMA ? 0 : f0;
        const float zslack = TMA ? T.zslack : 1e-3f;
        for (int ii0 = warp * 32; ii0 < ni; ii0 += NT) {   // a warp takes 32 consecutive particles per round (warp-uniform trip count)
            const int ii = ii0 + lane;
            const bool active = ii < ni;
            int c = 0, gi = 0, lx = 0, ly = 0;
            IState<R, DIM> I;
            Acc<R> a{0, 0, 0, 0};
            float xf = 0, yf = 0, zf = 0, rc2f = 0, rc2m = -1.0f;
            bool fluid_i = true;
            if (active) {
                while (c + 1 < NI && ii >= s_ipre[c + 1]) ++c;
                gi = s_ibeg[c] + (ii - s_ipre[c]);
                lx = c / BB; ly = c - lx * BB;
                if (REC) {
                    const P2* q = reinterpret_cast<const P2*>(A.rec) + rec_index(gi);
                    const P2 r0 = q[0], r1 = q[8], r2 = q[16], r3 = q[24], r4 = q[32];
                    if (COUPLED) fluid_i = r4.x > (R)0;
                    load_i<R, DIM>(I, C, r0.x, r0.y, r1.x, r1.y, r2.x, r2.y, r3.x, r3.y, r4.y);
                } else {
                    if (COUPLED) fluid_i = A.m[gi] > (R)0;
                    load_i<R, DIM>(I, C, A.x[gi], A.y[gi], DIM == 3 ? A.z[gi] : (R)0, A.u[gi], A.v[gi], DIM == 3 ? A.w[gi] : (R)0, A.rho[gi],
                                   A.por2[gi], A.h[gi]);
                }
                if (TMA) {
                    xf = pos_f32<R>(I.x, A.F.lo[0], A.F.cmin, A.F.cmax[0]); yf = pos_f32<R>(I.y, A.F.lo[1], A.F.cmin, A.F.cmax[1]);
                    zf = DIM == 3 ? pos_f32<R>(I.z, A.F.lo[2], A.F.cmin, A.F.cmax[2]) : 0.0f;
                } else {
                    xf = (float)(I.x - ox); yf = (float)(I.y - oy); zf = DIM == 3 ? (float)(I.z - oz) : 0.0f;
                }
                // conservative f32 pre-filter: relative margin on rc^2 (2^-15 for tile-local coordinates, from the box size for
                // grid-relative ones: see launch_zrun); the range cull below uses twice that on top
                rc2f = __double2float_ru((double)(UNI ? C.u_rc2 : I.rc2) * (1.0 + (double)marg));
                rc2m = rc2f * (1.0f + 2.0f * marg);
            }
            const float fz = DIM == 3 ? zf : yf;            // coordinate along the fast axis, relative to fine cell f0
            const float2 xf2 = make_float2(xf, xf), yf2 = make_float2(yf, yf), zf2 = make_float2(zf, zf);
            const float2 nrc2 = make_float2(-rc2f, -rc2f);

            // ---- per-lane scan range of stencil run k: staged [vs, ve), aligned start a4 = vs & ~3, ng 4-groups
            int q_run = 0, vs = 0, ve = 0;
            auto open_run = [&](int k) {
                const int ax = DIM == 3 ? k / 3 : k, ay = DIM == 3 ? k - ax * 3 : 0;
                q_run = (lx + ax) * RY + (DIM == 3 ? ly + ay : 0);
                // distance from the particle to the footprint of that column (0 for its own column)
                float dxc = 0.0f, dyc = 0.0f;
                if (ax == 0) dxc = fmaxf(xf - (float)(cxo + lx) * cellf, 0.0f);
                if (ax == 2) dxc = fmaxf((float)(cxo + lx + 1) * cellf - xf, 0.0f);
                if (DIM == 3) {
                    if (ay == 0) dyc = fmaxf(yf - (float)(cyo + ly) * cellf, 0.0f);
                    if (ay == 2) dyc = fmaxf((float)(cyo + ly + 1) * cellf - yf, 0.0f);
                }
                const float rem = rc2m - (dxc * dxc + dyc * dyc);
                vs = ve = 0;
                if (active && rem > 0.0f) {
                    const float zext = sqrtf(rem) * 1.00001f;
                    // fine cells [tlo, thi] relative to boundary 0 of the staged run (= fine cell f0 - S); +-1e-3 cell of slack,
                    // then clamped like the keys themselves (particles outside the box sit in the edge cells)
                    int flo = (int)floorf((fz - zext) * inv_cf - zslack) + fzo, fhi = (int)floorf((fz + zext) * inv_cf + zslack) + fzo;
                    flo = min(max(flo, 0), nf - 1); fhi = min(max(fhi, 0), nf - 1);
                    const int tlo = min(max(flo - (f0 - S), 0), W - 2), thi = min(max(fhi - (f0 - S), 0), W - 2);
                    vs = s_cs[q_run * WM + tlo];
                    ve = s_cs[q_run * WM + thi + 1];
                }
            };

            int k = 0, wc = 0;          // warp-uniform scan cursor: stencil run, 32-candidate chunk inside it
            while (DBG != 5) {           // (DBG 5, timing ablation: staging and barriers only)
                // ---- phase 1: pre-filter on staged f32 coordinates -> bit masks (bit 31 = first candidate of the word)
                int nw = 0;
                while (k < NRUN) {
                    open_run(k);
                    const int gb31 = s_gbeg[q_run] + 31;      // staged index -> global index (+ 31: the iterator counts leading zeros)
                    const int a4 = vs & ~3;
                    const int ng = (ve - a4 + 3) >> 2;
                    const int Tg = __reduce_max_sync(0xffffffffu, ve > vs ? ng : 0);   // the warp scans its longest range
                    bool full = false;
                    if (DBG == 4) { a.au += (R)(vs + ve + Tg); ++k; continue; }      // timing ablation: range set-up only
                    while (wc * 8 < Tg) {
                        if (__any_sync(0xffffffffu, nw == MAXW)) { full = true; break; }
                        const int iters = min(8, Tg - wc * 8);
                        const int j0 = a4 + wc * 32;
                        const float* px = s_x + min(j0, kZRow - 32);           // (only a lane past its own range is ever clamped)
                        unsigned m = 0;
#pragma unroll 4
                        for (int it = 0; it < iters; ++it, px += 4) {
                            const float4 X = *reinterpret_cast<const float4*>(px), Y = *reinterpret_cast<const float4*>(px + kZRow);
                            const float4 Z = DIM == 3 ? *reinterpret_cast<const float4*>(px + 2 * kZRow) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                            const float2 dxa = __fadd2_rn(xf2, make_float2(-X.x, -X.y)), dxb = __fadd2_rn(xf2, make_float2(-X.z, -X.w));
                            const float2 dya = __fadd2_rn(yf2, make_float2(-Y.x, -Y.y)), dyb = __fadd2_rn(yf2, make_float2(-Y.z, -Y.w));
                            float2 da = __ffma2_rn(dya, dya, __ffma2_rn(dxa, dxa, nrc2));
                            float2 db = __ffma2_rn(dyb, dyb, __ffma2_rn(dxb, dxb, nrc2));
                            if (DIM == 3) {
                                const float2 dza = __fadd2_rn(zf2, make_float2(-Z.x, -Z.y)), dzb = __fadd2_rn(zf2, make_float2(-Z.z, -Z.w));
                                da = __ffma2_rn(dza, dza, da);
                                db = __ffma2_rn(dzb, dzb, db);
                            }
                            // sign bit of d = r^2 - rc^2 -> the word (d < 0: inside the margin-inflated cutoff)
                            m = __funnelshift_l(__float_as_uint(da.x), m, 1);
                            m = __funnelshift_l(__float_as_uint(da.y), m, 1);
                            m = __funnelshift_l(__float_as_uint(db.x), m, 1);
                            m = __funnelshift_l(__float_as_uint(db.y), m, 1);
                        }
                        m <<= 32 - 4 * iters;                                  // left-align: bit 31 = candidate j0
                        const int lim = min(max(ve - j0, 0), 32);              // this lane's own range ends here; beyond it: over-scan with the warp
                        m &= (unsigned)(0xFFFFFFFF00000000ull >> lim);
                        if (wc == 0) m &= 0xFFFFFFFFu >> (vs - a4);            // aligned-down head
                        if (m) {
                            my_word[nw * NT] = make_uint2(m, (unsigned)(j0 + gb31));
                            ++nw;
                        }
                        ++wc;
                    }
                    if (full) break;
                    ++k; wc = 0;
                }
                // ---- phase 2: one hit per body, software-pipelined -- the record of the NEXT hit is in flight while the body of
                // the current one runs (two register sets, loop unrolled by two), so the L1 latency of the gathers hides behind
                // ~40 FP64 instructions instead of stalling the warp at the head of every trip.  The loop is warp-uniform (runs
                // until no lane has a hit left); a lane without a hit evaluates a masked dummy pair against a warp-common record.
                // Exact FMA-free test on the f64 record: the neighbour set is decided here.
                if (DBG >= 3) a.au += (R)nw;      // timing ablation: phase 1 only
                else {
                    // hit iterator over the mask words (none empty) with one word of lookahead: m = bits left in the current
                    // word, base31 = global index of its first candidate + 31; (mN, bN) = the next word, raw; w = the one after
                    // All of it branch-free (one predicated 8-byte shared-memory load per word), so the pipelined loop below is ONE basic
                    // block and ptxas can hoist a pop and its gathers above the body that precedes them.
                    const unsigned a_word = (unsigned)__cvta_generic_to_shared(my_word);
                    const unsigned pend = a_word + (unsigned)nw * (NT * 8);
                    unsigned m = 0, mN = 0, base31 = 0, baseN = 0, pw = a_word + 2 * (NT * 8);
                    if (nw > 0) { const uint2 w0 = my_word[0]; m = w0.x; base31 = w0.y; }
                    if (nw > 1) { const uint2 w1 = my_word[NT]; mN = w1.x; baseN = w1.y; }
                    const int jdummy = s_ibeg[0];
                    auto pop = [&](bool& valid) -> int {
                        valid = m != 0;
                        const int c = __clz(m);
                        int j = (int)base31 - 31 + c;
                        m &= __funnelshift_rc(0x7fffffffu, 0u, c);      // clear the bit just taken (c = 32: nothing left anyway)
                        const bool adv = valid && m == 0;      // word exhausted: step to the preloaded one, preload the one after
                        const bool ld = adv && pw < pend;
                        m = adv ? mN : m;
                        base31 = adv ? baseN : base31;
                        mN = adv ? 0u : mN;
                        lds_u64_if(pw, ld, mN, baseN);
                        pw += adv ? (unsigned)(NT * 8) : 0u;
                        j = valid ? j : jdummy;
                        asm("" : "+r"(j));        // keep j a plain 32-bit value: the record index below is 32-bit arithmetic + one IMAD.WIDE
                        return j;
                    };
                    struct JRec { R x, y, z, u, v, w, rho, por2, m; };
                    auto load_j = [&](int j) -> JRec {
                        JRec r;
                        r.m = A.m_uni;
                        if (DBG == 1) {     // timing ablation (wrong results): no gathers, the j state is made up from the index
                            r.x = I.x + (R)(j & 15) * (R)1e-3; r.y = I.y + (R)(j & 7) * (R)1e-3; r.z = I.z; r.u = I.u; r.v = I.v; r.w = (R)j; r.rho = I.rho; r.por2 = I.por2;
                        } else if (REC) {
                            const P2* q = reinterpret_cast<const P2*>(A.rec) + rec_index(j);
                            const P2 a0 = q[0], b0 = q[8], c0 = q[16], d0 = q[24];
                            r.x = a0.x; r.y = a0.y; r.z = b0.x; r.u = b0.y; r.v = c0.x; r.w = c0.y; r.rho = d0.x; r.por2 = d0.y;
                            if (!UNI) r.m = q[32].x;
                        } else {
                            r.x = A.x[j]; r.y = A.y[j]; r.z = DIM == 3 ? A.z[j] : (R)0; r.u = A.u[j]; r.v = A.v[j]; r.w = DIM == 3 ? A.w[j] : (R)0;
                            r.rho = A.rho[j]; r.por2 = A.por2[j];
                            if (!UNI) r.m = A.m[j];
                        }
                        return r;
                    };
                    const R rc2 = UNI ? C.u_rc2 : I.rc2;
                    auto body = [&](const JRec& J, bool valid) {
                        const R dx = I.x - J.x, dy = I.y - J.y, dz = DIM == 3 ? I.z - J.z : (R)0;
                        R r2 = dist2<DIM, R>(dx, dy, dz);
                        const bool in = valid && r2 < rc2 && r2 > (R)0;              // the exact test (the set is defined here)
                        r2 = in ? r2 : (R)1;
                        R mj = in ? J.m : (R)0;
                        if (COUPLED) mj = (fluid_i || mj > (R)0) ? fabs(mj) : (R)0;     // signed SPH mass: the pair counts iff i or j is fluid
                        pair_body<R, DIM, CONT, MOM, UNI, BETA0, UNI && !COUPLED>(C, I, dx, dy, dz, r2, J.u, J.v, J.w, J.rho, mj, J.por2, a, in);
                    };
                    // pops run one hit ahead of the gathers, gathers one body ahead of their use
                    bool vA, vB;
                    const int jA0 = pop(vA);
                    JRec rA = load_j(jA0);
                    int jB = pop(vB);
                    while (__any_sync(0xffffffffu, vA)) {
                        const JRec rB = load_j(jB);
                        const bool vb = vB;
                        bool vA2;
                        const int jA2 = pop(vA2);
                        body(rA, vA);
                        rA = load_j(jA2);
                        vA = vA2;
                        jB = pop(vB);
                        body(rB, vb);
                    }
                }
                if (k == NRUN) break;      // warp-uniform
            }
            if (active) {
                store_acc<R, DIM, CONT, MOM>(A, C, gi, a);
            }
        }
    };


    // ---- persistent loop over this CTA's tiles, staging one tile ahead
    // tiles are handed out dynamically (one atomic per tile, fetched two iterations ahead by thread 0 and published through
    // shared memory at the barrier in between): neighbours in the list are evaluated at about the same time by different SMs
    // (their common candidates meet in L2) and no CTA waits for a slow one at the end
    __shared__ int s_next[2];
    int k = blockIdx.x, k1 = blockIdx.x + gridDim.x;        // the first two tiles are static
    if (k >= ntiles) return;
    stage_tile(k, 0);
    __syncthreads();
    unsigned phase = 0;
    for (int n = 0; k < ntiles; ++n) {
        if (TMA) { mbar_wait(&s_bar, phase); phase ^= 1u; }  // the rows of this tile have landed
        if (tid == 0) s_next[n & 1] = atomicAdd(T.next, 1);  // tile of iteration n + 2
        compute_tile(NBUF == 2 ? n & 1 : 0);
        if (NBUF == 1) __syncthreads();                      // single buffer: everybody is done with it before it is refilled
        if (k1 < ntiles) stage_tile(k1, NBUF == 2 ? (n + 1) & 1 : 0);
        __syncthreads();
        k = k1;
        k1 = s_next[n & 1];
    }
}

template <class R, int DIM, int TA, int TB, int NT, int NBUF, bool CONT, bool MOM, bool COUPLED = false, bool UNI = false>
pst_status launch_zrun_k(pst_ctx* ctx, const ZTile& T, size_t smem, unsigned grid) {
    const bool rec = rec_wanted(ctx);
    if (rec) PST_TRY(rec_refresh<R>(ctx));
    auto kern = rec ? k_wcsph_zrun<R, DIM, TA, TB, NT, CONT, MOM, COUPLED, UNI, true, 0, NBUF> : k_wcsph_zrun<R, DIM, TA, TB, NT, CONT, MOM, COUPLED, UNI, false, 0, NBUF>;
    if (MOM && pst_param(ctx, "beta") == 0.0)      // no quadratic viscosity term (every BASELINE config): the cheaper form of Pi
        kern = rec ? k_wcsph_zrun<R, DIM, TA, TB, NT, CONT, MOM, COUPLED, UNI, true, 0, NBUF, MOM> : k_wcsph_zrun<R, DIM, TA, TB, NT, CONT, MOM, COUPLED, UNI, false, 0, NBUF, MOM>;
    if (UNI && rec && CONT && MOM && !COUPLED && DIM == 3 && sizeof(R) == 8) {     // timing ablations of the headline kernel (wrong results)
        const int dbg = pst_option(ctx, "tile_dbg", 0);
        if (dbg == 1) kern = k_wcsph_zrun<R, DIM, TA, TB, NT, CONT, MOM, COUPLED, UNI, true, 1, NBUF, MOM>;
        if (dbg == 3) kern = k_wcsph_zrun<R, DIM, TA, TB, NT, CONT, MOM, COUPLED, UNI, true, 3, NBUF, MOM>;
        if (dbg == 4) kern = k_wcsph_zrun<R, DIM, TA, TB, NT, CONT, MOM, COUPLED, UNI, true, 4, NBUF, MOM>;
        if (dbg == 5) kern = k_wcsph_zrun<R, DIM, TA, TB, NT, CONT, MOM, COUPLED, UNI, true, 5, NBUF, MOM>;
    }
    PST_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WcsphConst<R> C = make_const<R>(ctx);
    if (UNI) fill_uniform<R, DIM>(ctx, C);
    PST_LAUNCH(ctx, kern, grid, NT, smem, make_grid_dev<R>(ctx->grid), C, make_args<R>(ctx), T);
    return PST_OK;
}

template <class R, int DIM, int NT, int NBUF>
pst_status launch_zrun_shape(pst_ctx* ctx, bool cont, bool mom) {
    // 2 x 2 columns, 256 threads, two CTAs per SM.  Measured and dropped (profiles/r2_exp_log.txt): 4 x 2 columns with 512 threads
    // and one CTA per SM (+6 %), three CTAs per SM at 80 registers (+12..21 %: spills, and the third CTA's shared memory leaves too
    // little L1 for the gathers), prefetch.global.L1 of the next trip's records (+25 %)
    constexpr int TA = 2, TB = 2;
    constexpr int kZJcap = zjcap<NT>();
    using D = TileDims<DIM, TA, TB>;
    const PstGrid& g = ctx->grid;
    const int S = g.sub;
    const int nf = g.n[DIM - 1];
    double ppc = pst_param(ctx, "_ppc", 0.0);      // mean occupancy of the occupied COARSE cells (k_scan_tiles)
    if (!(ppc > 0)) ppc = DIM == 3 ? 14.0 : 6.0;
    ZTile T;
    T.maxw = pst_option(ctx, "tile_words", 10);
    // deepest tile (fine cells): twice the depth that holds one thread per particle at the mean occupancy, bounded by the f32
    // pre-filter (tile-local coordinates reach gf / S + 6 cells: <= 32 cells keeps the f32 error of r^2 below half of the 2^-15
    // margin, tests/test_prefilter_margin.py), by the boundary table and by the staging rows
    const int user_G = pst_option(ctx, "tile_g", 0) * S + pst_option(ctx, "tile_gf", 0);   // tile_g in COARSE cells, tile_gf in fine cells
    int gfcap = user_G > 0 ? user_G : (int)std::ceil(2.0 * NT * kZRounds * S / (D::NI * ppc));
    gfcap = std::min(std::max(gfcap, 1), std::max(1, nf));
    gfcap = std::min(gfcap, kMaxTileG * S);
    while (gfcap > 1 && D::NR * (gfcap + 2 * S + 1) > 2048) --gfcap;
    T.gfcap = gfcap;
    T.wmax = gfcap + 2 * S + 1;
    T.jcap = std::min(kZJcap, std::max(0, pst_option(ctx, "tile_jcap", kZJcap)));
    // ---- the tile list (device side; rebuilt when the cell table or the cut parameters changed)
    const int tiles_x = (g.n[0] + TA - 1) / TA, tiles_y = DIM == 3 ? (g.n[1] + D::BB - 1) / D::BB : 1;
    const int target = user_G > 0 ? 0x3fffffff : NT * kZRounds;       // a forced depth (tests) cuts by depth alone
    const size_t cap = (size_t)tiles_x * tiles_y * nf;     // a tile holds at least one fine cell: the list cannot overflow
    if (cap > ctx->ztiles_cap) {
        PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->ztiles); ctx->ztiles = nullptr;
        PST_CUDA(ctx, cudaMalloc(&ctx->ztiles, 2 * cap * sizeof(int4)));      // the ordered list, then the per-pair slot ranges of the cut
        ctx->ztiles_cap = cap;
        ctx->ztiles_key = 0;
    }
    const int nwarps = tiles_x * tiles_y;
    if ((size_t)nwarps + 2 > ctx->ztile_off_cap) {
        PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_ztile_count); ctx->d_ztile_count = nullptr;
        PST_CUDA(ctx, cudaMalloc((void**)&ctx->d_ztile_count, ((size_t)nwarps + 2) * sizeof(int)));   // [0] dynamic tile counter, [1 ..] tiles per column pair -> offsets
        ctx->ztile_off_cap = (size_t)nwarps + 2;
        ctx->ztiles_key = 0;
    }
    int* const d_off = ctx->d_ztile_count + 1;
    const uint64_t key = ctx->build_epoch * 1000003ull + (uint64_t)gfcap * 4099 + (uint64_t)target + (uint64_t)T.jcap * 7919 + (uint64_t)NT * 104729;
    if (key != ctx->ztiles_key) {
        int4* const slots = (int4*)ctx->ztiles + cap;
        PST_CUDA(ctx, cudaMemsetAsync(d_off + nwarps, 0, sizeof(int), ctx->stream));
        PST_LAUNCH(ctx, (k_ztile_list<DIM, TA, TB>), blocks_for((size_t)nwarps * 32, 128), 128, 0, g.n[0], DIM == 3 ? g.n[1] : 1, nf, g.cx_lo, g.cx_hi,
                   tiles_x, tiles_y, target, gfcap, S, user_G > 0 ? 0x3fffffff : T.jcap, ctx->cell_start, slots, d_off);
        PST_TRY(pst_scan_exclusive(ctx, d_off, nwarps + 1));     // off[nwarps] = number of tiles
        PST_LAUNCH(ctx, k_ztile_compact, blocks_for((size_t)nwarps * 32, 128), 128, 0, nwarps, nf, slots, d_off, (int4*)ctx->ztiles);
        ctx->ztiles_key = key;
    }
    // TMA staging works on f32 coordinates relative to the GRID origin: their error grows with the box, and so must the margin of
    // the pre-filter.  Two stored coordinates are each within half an ulp of E = the largest clamped coordinate, so a difference is
    // off by <= ulp(E) = 2^-23 E per axis and r^2 by <= 2 rc sqrt(3) 2^-23 E (+ 3 roundings of the f32 evaluation): relative to
    // rc^2 that is 3.5 * 2^-23 * E / rc; the margin is 8 * 2^-23 * E / rc, never below the 2^-15 of the tile-local form
    // (tests/test_prefilter_margin.py::test_grid_relative_coordinates).  At E / rc = 333 (the 80 M tank) it is 3.2e-4: 0.1 % more
    // candidates reach the exact test.
    {
        double E = 0;
        for (int a = 0; a < DIM; ++a) E = std::max(E, (double)(g.n[a] / (a == DIM - 1 ? S : 1) + 2) * g.cell);
        // (the SMALLEST cutoff sets the relative error: h_value is the minimum of h over the owned particles)
        const double rc = std::max(1e-300, pst_param(ctx, "kfac", 2.0) * (ctx->h_value > 0 ? ctx->h_value : g.cell / pst_param(ctx, "kfac", 2.0)));
        T.marg = (float)std::max(1.0 / 32768.0, 8.0 * std::ldexp(1.0, -23) * E / rc);
        T.zslack = (float)std::max(1e-3, (double)nf * std::ldexp(1.0, -20));
    }
    T.tiles = (const int4*)ctx->ztiles;
    T.ntiles = d_off + nwarps;
    T.next = ctx->d_ztile_count;
    constexpr int kZRow = kZJcap + kZPad;
    const size_t tab_ints = (size_t)((D::NR + D::NR + 1 + D::NI + D::NI + 1 + 4 + NT / 32 + D::NR * T.wmax + 3) & ~3);
    constexpr int nbuf = NBUF;
    const size_t smem = nbuf * tab_ints * sizeof(int) + (size_t)nbuf * (DIM == 3 ? 3 : 2) * kZRow * sizeof(float) + (size_t)T.maxw * NT * 8;
    if (smem > 113 * 1024) return pst_fail(ctx, PST_EINVAL, "tile_words too large (%zu bytes of shared memory per CTA)", smem);
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)((512 / NT) * nsm);        // persistent: 512 threads per SM
    k_set_int<<<1, 1, 0, ctx->stream>>>(ctx->d_ztile_count, (int)(2 * grid));
#ifdef PST_DEV_HEADLINE
    PST_TRY(pst_uniform_refresh(ctx));
    if (!(cont && mom && ctx->m_uniform && ctx->h_uniform && !ctx->coupled)) return pst_fail(ctx, PST_EINVAL, "development build: uniform m and h, continuity + momentum only");
    return launch_zrun_k<R, DIM, TA, TB, NT, NBUF, true, true, false, true>(ctx, T, smem, grid);
#else
    if (ctx->coupled) {
        if (DIM == 3) return launch_zrun_k<R, 3, TA, TB, NT, NBUF, true, true, true>(ctx, T, smem, grid);
        return pst_fail(ctx, PST_EINVAL, "coupled contexts need dim = 3");
    }
    // every particle this rank can see has the same mass AND smoothing length (owned ones: checked on the device; ghosts and
    // migrants of other ranks: the caller vouches for them with "uniform_mass_global"): both become kernel constants
    PST_TRY(pst_uniform_refresh(ctx));
    const bool uni = ctx->m_uniform && ctx->h_uniform && (!ctx->comm || pst_option(ctx, "uniform_mass_global", 0) != 0) && pst_option(ctx, "uniform_mass", 1) != 0;
    if (cont && mom && uni) return launch_zrun_k<R, DIM, TA, TB, NT, NBUF, true, true, false, true>(ctx, T, smem, grid);
    if (cont && mom) return launch_zrun_k<R, DIM, TA, TB, NT, NBUF, true, true>(ctx, T, smem, grid);
    if (cont) return launch_zrun_k<R, DIM, TA, TB, NT, NBUF, true, false>(ctx, T, smem, grid);
    return launch_zrun_k<R, DIM, TA, TB, NT, NBUF, false, true>(ctx, T, smem, grid);
#endif
}

template <class R, int DIM>
pst_status launch_zrun(pst_ctx* ctx, bool cont, bool mom) {
    // two CTAs of 256 threads per SM, two staging buffers.  Measured and dropped (profiles/r2_exp_log.txt): one staging buffer
    // (+2 %), four CTAs of 128 threads with one or two buffers (+8 %)
    return launch_zrun_shape<R, DIM, 256, kZNbuf>(ctx, cont, mom);
}
