// wcsph.cu -- weakly-compressible SPH: Tait EOS, continuity, momentum (+ artificial viscosity,
// gravity), semi-implicit Euler stage.  Formulation: SURVEY.md Appendix A.2 / DESIGN.md.
// No reference code exists for the physics; what the reference fixes is
// the loop shape -- gather into [i], bodies of a fused set share one i,j loop
// (prestige/src/codegen/simple_cpu.rs:7-16, prestige/src/equations/fuse.rs:14-40).
//
// Three pair kernels (option force_kernel), all computing continuity and/or momentum in ONE j loop:
//   0 k_wcsph_gather    one thread per particle, walks its 9 (3D) / 3 (2D) contiguous candidate runs
//                       straight from global memory.  Any key mode, any occupancy.
//   1 k_wcsph_cellwarp  one warp per cell, candidates cached in registers, hits ballot-compacted.  Linear keys.
//   2 k_wcsph_tiled     (default) one CTA per tile of A x B cell columns x G cells along the fast axis.  The
//                       (A+2)(B+2) candidate runs are staged ONCE in shared memory (coalesced loads of the
//                       sorted SoA arrays) as f32 coordinates + index, then every thread (a) scans its candidates
//                       (packed f32x2 arithmetic; a conservative pre-filter in f64 contexts) and compacts the hits
//                       into a private list, (b) evaluates the expensive pair body only for the hits -- the exact,
//                       FMA-free cutoff test decides there -- with all lanes busy.  Linear keys only.
// Also here: the EOS, the dummy-particle wall pressure (k_wall_pressure) and the semi-implicit Euler stages.
#include <cstring>

#include "pst_internal.h"
#include "wcsph_core.h"   // WcsphConst, IState, Acc, load_i, pair_body, wall_accumulate / wall_finish (shared with the host test harness)

namespace {

constexpr int kThreads = 128;
inline unsigned blocks_for(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

// Are the masses / smoothing lengths of the owned particles all equal?  min and max of the BIT PATTERNS (any total order
// does: all equal <=> min == max), one atomic pair per block.
template <class R>
__global__ void __launch_bounds__(256) k_uniform_check(int n, const R* __restrict__ m, const R* __restrict__ h, const R* __restrict__ rad,
                                                       unsigned long long* __restrict__ out) {
    unsigned long long lo_m = ~0ull, hi_m = 0, lo_h = ~0ull, hi_h = 0, hi_r = 0;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const unsigned long long km = (unsigned long long)__double_as_longlong((double)m[s]);
        const unsigned long long kh = h ? (unsigned long long)__double_as_longlong((double)h[s]) : 0ull;
        lo_m = min(lo_m, km); hi_m = max(hi_m, km); lo_h = min(lo_h, kh); hi_h = max(hi_h, kh);
        if (rad) hi_r = max(hi_r, (unsigned long long)__double_as_longlong(fabs((double)rad[s])));     // (largest radius: the cell-size precondition)
    }
    for (int d = 16; d > 0; d >>= 1) {
        lo_m = min(lo_m, __shfl_xor_sync(0xffffffffu, lo_m, d)); hi_m = max(hi_m, __shfl_xor_sync(0xffffffffu, hi_m, d));
        lo_h = min(lo_h, __shfl_xor_sync(0xffffffffu, lo_h, d)); hi_h = max(hi_h, __shfl_xor_sync(0xffffffffu, hi_h, d));
        hi_r = max(hi_r, __shfl_xor_sync(0xffffffffu, hi_r, d));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(out, lo_m); atomicMax(out + 1, hi_m); atomicMin(out + 2, lo_h); atomicMax(out + 3, hi_h); atomicMax(out + 4, hi_r); }
}

template <class R>
WcsphConst<R> make_const(pst_ctx* ctx) {
    WcsphConst<R> C;
    const double rho0 = pst_param(ctx, "rho0"), c0 = pst_param(ctx, "c0"), gamma = pst_param(ctx, "gamma");
    C.kfac = (R)pst_param(ctx, "kfac", 2.0);
    C.rho0 = (R)rho0; C.c0 = (R)c0; C.gamma = (R)gamma;
    C.B = (R)(rho0 * c0 * c0 / gamma);
    C.alpha_c0 = (R)(pst_param(ctx, "alpha") * c0);
    C.beta = (R)pst_param(ctx, "beta");
    C.g[0] = (R)pst_param(ctx, "gx"); C.g[1] = (R)pst_param(ctx, "gy"); C.g[2] = (R)pst_param(ctx, "gz");
    C.gamma_is_7 = gamma == 7.0;
    C.u_h = C.u_half_inv_h = C.u_gfc = C.u_eta2 = C.u_rc2 = C.u_mgfc = C.u_visc = (R)0;
    return C;
}

// the h-derived constants of a context whose particles all share one smoothing length: the same load_i the kernels run
template <class R, int DIM>
void fill_uniform(pst_ctx* ctx, WcsphConst<R>& C) {
    IState<R, DIM> I;
    load_i<R, DIM>(I, C, (R)0, (R)0, (R)0, (R)0, (R)0, (R)0, (R)0, (R)0, (R)ctx->h_value);
    C.u_h = I.h; C.u_half_inv_h = I.half_inv_h; C.u_gfc = I.gfc; C.u_eta2 = I.eta2; C.u_rc2 = I.rc2;
    C.u_mgfc = (R)ctx->m_value * I.gfc; C.u_visc = (R)-2 * C.alpha_c0 * I.h;
}

// ---------------------------------------------------------------------------------------------
// EOS: p = B((rho/rho0)^gamma - 1) = B expm1(gamma log1p(rho/rho0 - 1)), and p/rho^2 which is what the momentum body consumes.
// Runs over owned + ghost particles.
// ---------------------------------------------------------------------------------------------
// Coupled SPH-DEM contexts (DESIGN.md "Coupled formulation") also get `msph`, the SPH mass the pair kernels read in
// place of m: +m for fluid, -m for boundaries, -m rho0/rho_solid (the displaced fluid mass) for solids.  The sign
// carries "is fluid" into the pair loop without another gather: a pair is active iff one of its two particles is fluid.
// Packed neighbour-state records for the tiled pair kernel (variant 3, option rec_impl = 1): blocks of 8 particles x 5
// rows of 8 pairs -- (x,y) (z,u) (v,w) (rho, p/rho^2) (m or signed SPH mass, h) -- so one pair of the loop costs 4 (5) 16-byte
// loads off ONE address instead of 8 (9) 8-byte loads off 8 (9) array bases; lanes of a warp that gather different j hit
// bank group j mod 8 of the L1 data array, whatever the row.  Element (pair) index of particle j, row r: j + (j & ~7) * 4 + 8 r
// (valid for the negative ghost indices too).
template <class R> struct RecPair;
template <> struct RecPair<double> { using type = double2; };
template <> struct RecPair<float> { using type = float2; };
constexpr int kRecRows = 5;
// 32-bit arithmetic: the context refuses records beyond 2^31 / 5 particles (rec_alloc)
__host__ __device__ inline int rec_index(int j) { return j + (j & ~7) * (kRecRows - 1); }

// f32 position relative to the grid origin, clamped to the box + 2 cells.  The clamp keeps the f32 error bounded by the box
// size whatever the coordinates of particles far outside (they are binned into the edge cells); it is a contraction, so a
// distance between clamped coordinates never exceeds the true one: the pre-filter stays conservative.
template <class R>
__host__ __device__ __forceinline__ float pos_f32(R x, R lo, float cmin, float cmax) {
    const float v = (float)(x - lo);
    return v < cmin ? cmin : (v > cmax ? cmax : v);
}
template <class R>
struct PosF { float *x, *y, *z; R lo[3]; float cmin, cmax[3]; };

template <class R>
struct RecSrc { const R *x, *y, *z, *u, *v, *w, *rho, *por2, *m, *h; PosF<R> F; };

template <class R>
__device__ __forceinline__ void rec_store_vals(R* __restrict__ rec, const PosF<R>& F, int s, R x, R y, R z, R u, R v, R w, R rho_s, R por2_s, R m, R h) {
    using P = typename RecPair<R>::type;
    P* q = reinterpret_cast<P*>(rec) + rec_index(s);
    P a; a.x = x; a.y = y; q[0] = a;
    a.x = z; a.y = u; q[8] = a;
    a.x = v; a.y = w; q[16] = a;
    a.x = rho_s; a.y = por2_s; q[24] = a;
    a.x = m; a.y = h; q[32] = a;
    if (F.x) {
        F.x[s] = pos_f32<R>(x, F.lo[0], F.cmin, F.cmax[0]);
        F.y[s] = pos_f32<R>(y, F.lo[1], F.cmin, F.cmax[1]);
        if (F.z) F.z[s] = pos_f32<R>(z, F.lo[2], F.cmin, F.cmax[2]);
    }
}
template <class R>
__device__ __forceinline__ void rec_store(R* __restrict__ rec, const RecSrc<R>& S, int s, R rho_s, R por2_s) {
    rec_store_vals<R>(rec, S.F, s, S.x[s], S.y[s], S.z ? S.z[s] : (R)0, S.u[s], S.v[s], S.w ? S.w[s] : (R)0, rho_s, por2_s, S.m[s], S.h[s]);
}

// Tait EOS in the cancellation-free form (same as the oracle): p = B expm1(gamma log1p(rho/rho0 - 1))
template <class R>
__device__ __forceinline__ R tait_pressure(const WcsphConst<R>& C, R rho) {
    const R e = (rho - C.rho0) / C.rho0;
    return C.B * expm1(C.gamma * log1p(e));
}

// Re-sort + EOS + records in ONE pass (single-GPU WCSPH contexts with packed records): the state permute already holds
// x y z u v w rho m h of a particle in registers, so it also evaluates the EOS and writes p, p/rho^2, the packed record and the
// f32 position -- instead of k_eos reading the nine arrays back (76 B per particle and step less HBM traffic, one launch less).
template <class R>
struct PermEosArgs {
    const R* src[9];      // x y z u v w rho m h (z, w null in 2D), the buffers of the OLD order
    R* dst[9];
    R *p, *por2, *rec;
    PosF<R> F;
    const uint32_t* src4[2];   // tag and id (4-byte rows; null if the context has none): they ride along, so a WCSPH context needs no
    uint32_t* dst4[2];         // second, generic permute launch for two small arrays
};
template <class R>
__global__ void __launch_bounds__(256) k_permute_eos(WcsphConst<R> C, int n, const uint32_t* __restrict__ perm, PermEosArgs<R> P) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t o = perm[s];
    R v[9];
#pragma unroll
    for (int a = 0; a < 9; ++a) v[a] = P.src[a] ? __ldg(P.src[a] + o) : (R)0;
#pragma unroll
    for (int a = 0; a < 9; ++a)
        if (P.dst[a]) P.dst[a][s] = v[a];
#pragma unroll
    for (int a = 0; a < 2; ++a)
        if (P.src4[a]) P.dst4[a][s] = __ldg(P.src4[a] + o);
    const R pr = tait_pressure<R>(C, v[6]);
    const R q = pr / (v[6] * v[6]);
    P.p[s] = pr;
    P.por2[s] = q;
    rec_store_vals<R>(P.rec, P.F, s, v[0], v[1], v[2], v[3], v[4], v[5], v[6], q, v[7], v[8]);
}

template <class R>
__global__ void __launch_bounds__(256) k_eos(WcsphConst<R> C, int lo, int hi, const R* __restrict__ rho, R* __restrict__ p,
                                             R* __restrict__ por2, const int32_t* __restrict__ tag, const R* __restrict__ m,
                                             R* __restrict__ msph, R solid_ratio, R* __restrict__ rec, RecSrc<R> S) {
    const int s = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= hi) return;
    if (msph) {
        const int t = tag[s];
        msph[s] = t == 0 ? m[s] : (t == 2 ? -(m[s] * solid_ratio) : -m[s]);
    }
    const R r = rho[s];
    const R pr = tait_pressure<R>(C, r);        // (rho/rho0)^gamma - 1 without cancellation near rho0 (same form as the oracle)
    const R q = pr / (r * r);
    p[s] = pr;
    por2[s] = q;
    if (rec) rec_store<R>(rec, S, s, r, q);      // S.m is msph in coupled contexts: written above by this same thread
}

// records alone (state changed after the EOS pass, e.g. by the wall-pressure equation, or continuity without the EOS)
template <class R>
__global__ void __launch_bounds__(256) k_rec_pack(int lo, int hi, R* __restrict__ rec, RecSrc<R> S) {
    const int s = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (s < hi) rec_store<R>(rec, S, s, S.rho[s], S.por2[s]);
}

template <class R>
struct ForceArgs {
    const R* rec;   // packed records (variant 3 with rec_impl = 1), element 0 = particle 0
    PosF<R> F;      // ... and the f32 positions written with them (const in the pair kernel)
    R m_uni;   // UMASS kernels: the mass every particle has
    const R *x, *y, *z, *u, *v, *w, *rho, *m, *h, *por2;
    R *au, *av, *aw, *arho;
    const int32_t* cell_start;
    int n;
};

template <class R, int DIM, bool CONT, bool MOM>
__device__ __forceinline__ void store_acc(const ForceArgs<R>& A, const WcsphConst<R>& C, int s, const Acc<R>& a) {
    if (CONT) A.arho[s] = a.arho;
    if (MOM) {
        A.au[s] = a.au + C.g[0];
        A.av[s] = a.av + C.g[1];
        if (DIM == 3) A.aw[s] = a.aw + C.g[2];
    }
}

// ---------------------------------------------------------------------------------------------
// variant 0: per-particle gather
// ---------------------------------------------------------------------------------------------
// COUPLED: A.m is the signed SPH mass (k_eos); the pair (i, j) counts iff i or j is fluid (mass > 0).
template <class R, int DIM, bool MORTON, bool CONT, bool MOM, bool COUPLED = false>
__device__ __forceinline__ void gather_one(const GridDev<R>& g, const WcsphConst<R>& C, const ForceArgs<R>& A, int s) {
    IState<R, DIM> I;
    load_i<R, DIM>(I, C, A.x[s], A.y[s], DIM == 3 ? A.z[s] : (R)0, A.u[s], A.v[s], DIM == 3 ? A.w[s] : (R)0, A.rho[s], A.por2[s], A.h[s]);
    const int cx = cell_coord<R>(I.x, g.lo[0], g.inv[0], g.cx_lo, g.cx_hi);
    const int cy = cell_coord<R>(I.y, g.lo[1], g.inv[1], 0, g.n[1] - 1);
    const int cz = DIM == 3 ? cell_coord<R>(I.z, g.lo[2], g.inv[2], 0, g.n[2] - 1) : 0;
    Acc<R> a{0, 0, 0, 0};
    const bool fluid_i = !COUPLED || A.m[s] > (R)0;
    for_each_run<DIM, MORTON>(g, A.cell_start, cx, cy, cz, [&](int b, int e) {
        for (int j = b; j < e; ++j) {
            const R dx = I.x - A.x[j], dy = I.y - A.y[j], dz = DIM == 3 ? I.z - A.z[j] : (R)0;
            const R r2 = dist2<DIM, R>(dx, dy, dz);
            if (r2 < I.rc2 && r2 > (R)0 && j != s) {
                const R mj = A.m[j];
                if (COUPLED && !(fluid_i || mj > (R)0)) continue;
                pair_body<R, DIM, CONT, MOM>(C, I, dx, dy, dz, r2, A.u[j], A.v[j], DIM == 3 ? A.w[j] : (R)0, A.rho[j], COUPLED ? fabs(mj) : mj, A.por2[j], a);
            }
        }
    });
    store_acc<R, DIM, CONT, MOM>(A, C, s, a);
}

template <class R, int DIM, bool MORTON, bool CONT, bool MOM, bool COUPLED = false>
__global__ void __launch_bounds__(kThreads) k_wcsph_gather(GridDev<R> g, WcsphConst<R> C, ForceArgs<R> A) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.n) return;
    gather_one<R, DIM, MORTON, CONT, MOM, COUPLED>(g, C, A, s);
}

// ---------------------------------------------------------------------------------------------
// variant 2: shared-memory tiles, ONE THREAD PER PARTICLE, private hit lists
//
//   * staged per tile: 16 B per candidate -- f32 coordinates (tile-local for f64 contexts) and the candidate's
//     global index, as four SoA rows x[] y[] z[] index[] (-DPST_P1_AOS: one float4 per candidate, the layout the
//     packed form was measured against).  16 B instead of 72 B per candidate leaves room for two CTAs per SM;
//   * phase 1: each thread scans its 9 (3) runs four candidates at a time: 3 LDS.128 + 6 FADD2 + 2 FMUL2 + 4 FFMA2
//     (sm_100a f32x2: two candidates per instruction) + 4 FSETP, appending hits to its private 16-bit list.  For
//     f64 contexts this is a CONSERVATIVE pre-filter (margin 2^-15, switched off for tiles holding far-away particles);
//   * phase 2: two hits per trip; the f64 state of j is gathered straight from global memory -- L1/L2 hits, since
//     the CTA's threads share the same ~1300 candidates and staging has just touched them -- the exact
//     FMA-free cutoff test decides membership, then the branch-free pair body runs.
// ---------------------------------------------------------------------------------------------
struct TileShape {
    int G;          // cells per tile along the fast axis
    int tiles[3];   // tile grid: [0] columns-x, [1] columns-y (1 in 2D), [2] fast axis
    int jcap;       // staged-candidate capacity (elements)
    int lcap;       // hit-list capacity per thread
};

template <int DIM, int A, int B>
struct TileDims {
    static constexpr int BB = DIM == 3 ? B : 1;
    static constexpr int NI = A * BB;                              // i columns
    static constexpr int NR = (A + 2) * (DIM == 3 ? BB + 2 : 1);   // staged runs
    static constexpr int RY = DIM == 3 ? BB + 2 : 1;               // runs along y
};

constexpr int kMaxTileG = 26;      // deepest tile: bounds the f32 error of the tile-local pre-filter (see launch_tiled)
constexpr int kListsJcap = 1536;   // float4 slots staged per tile (24 KB): every KB not requested stays L1 for the phase-2 gathers (2048 -> 1536: -3 %)

// UMASS: every particle has the mass A.m_uni (known from the host upload), so m[j] is not gathered.
template <class R, int DIM, int TA, int TB, int NT, bool CONT, bool MOM, bool COUPLED = false, bool UMASS = false>
__global__ void __launch_bounds__(NT, 2) k_wcsph_tiled(GridDev<R> g, WcsphConst<R> C, ForceArgs<R> A, TileShape T) {
    using D = TileDims<DIM, TA, TB>;
    constexpr int NR = D::NR, NI = D::NI, RY = D::RY, BB = D::BB;
    constexpr bool LOCAL = sizeof(R) == 8;
    constexpr int JC = kListsJcap;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int G = T.G, W = G + 3;  // W boundaries per run
    // ---- shared memory carve-up
    int* s_cs = reinterpret_cast<int*>(smem_raw);            // NR * W   virtual cell boundaries
    int* s_gbeg = s_cs + NR * W;                              // NR       global begin of each run
    int* s_voff = s_gbeg + NR;                                // NR + 1   virtual offset of each run
    int* s_ibeg = s_voff + NR + 1;                            // NI       global begin of each i segment
    int* s_ipre = s_ibeg + NI;                                // NI + 1   prefix of i counts
    size_t off = ((size_t)(NR * W + NR + NR + 1 + NI + NI + 1) * sizeof(int) + 15) & ~(size_t)15;
    double* s_org = reinterpret_cast<double*>(smem_raw + off);   // 3 (+1 pad): tile origin
    off += 4 * sizeof(double);
#if !defined(PST_P1_AOS)   // default: candidates staged as SoA rows so phase 1 runs packed f32x2 arithmetic (-DPST_P1_AOS: the float4 layout, A/B)
    float* s_x = reinterpret_cast<float*>(smem_raw + off);
    float* s_y = s_x + JC;
    float* s_z = s_y + JC;
    int* s_gj = reinterpret_cast<int*>(s_z + JC);
    unsigned short* s_list = reinterpret_cast<unsigned short*>(s_gj + JC);   // lcap * NT
#else
    float4* s_p4 = reinterpret_cast<float4*>(smem_raw + off);
    unsigned short* s_list = reinterpret_cast<unsigned short*>(s_p4 + JC);   // lcap * NT
#endif

    const int tid = threadIdx.x;
    // ---- which tile
    int b = blockIdx.x;
    const int tf = b % T.tiles[2]; b /= T.tiles[2];
    const int ty = DIM == 3 ? b % T.tiles[1] : 0; if (DIM == 3) b /= T.tiles[1];
    const int tx = b;
    const int cx0 = tx * TA, cy0 = ty * BB, f0 = tf * G;
    const int nf = DIM == 3 ? g.n[2] : g.n[1];   // cells along the fast axis
    const int ncx = g.n[0], ncy = DIM == 3 ? g.n[1] : 1;

    // ---- cell boundaries of every staged run (global indices first)
    for (int t = tid; t < NR * W; t += NT) {
        const int q = t / W, tt = t - q * W;
        const int rx = q / RY, ry = q - rx * RY;
        const int cx = cx0 - 1 + rx, cy = DIM == 3 ? cy0 - 1 + ry : 0;
        int gi = 0;   // column outside the grid: all boundaries equal -> empty run
        if (cx >= 0 && cx < ncx && cy >= 0 && cy < ncy) {
            const int col = DIM == 3 ? cx * ncy + cy : cx;
            gi = A.cell_start[(size_t)col * nf + min(max(f0 - 1 + tt, 0), nf)];
        }
        s_cs[t] = gi;
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int q = 0; q < NR; ++q) {
            s_gbeg[q] = s_cs[q * W];
            s_voff[q] = acc;
            acc += s_cs[q * W + W - 1] - s_cs[q * W];
        }
        s_voff[NR] = acc;
        // i segments: tile columns, fast cells [f0, f0 + G)  == boundaries tt = 1 .. G+1 of the centre runs
        int ia = 0, first = -1;
        for (int c = 0; c < NI; ++c) {
            const int lx = c / BB, ly = c - lx * BB;
            const int q = (lx + 1) * RY + (DIM == 3 ? ly + 1 : 0);
            const int cx = cx0 + lx, cy = cy0 + ly;
            int cnt = 0, beg = 0;
            if (cx >= g.cx_lo && cx <= g.cx_hi && cy < ncy) { beg = s_cs[q * W + 1]; cnt = s_cs[q * W + G + 1] - beg; }  // ghost layers are never i
            s_ibeg[c] = beg;
            s_ipre[c] = ia;
            ia += cnt;
            if (cnt > 0 && first < 0) first = beg;
        }
        s_ipre[NI] = ia;
        if (LOCAL && first >= 0) { s_org[0] = (double)A.x[first]; s_org[1] = (double)A.y[first]; s_org[2] = DIM == 3 ? (double)A.z[first] : 0.0; }
        else { s_org[0] = s_org[1] = s_org[2] = 0.0; }
    }
    __syncthreads();
    const int ni = s_ipre[NI];
    if (ni == 0) return;                      // empty tile (uniform exit)
    const int M = s_voff[NR];
    if (M > min(JC, T.jcap)) {
        // tile denser than the staging buffer: exact per-particle gather for its particles
        for (int ii = tid; ii < ni; ii += NT) {
            int c = 0;
            while (c + 1 < NI && ii >= s_ipre[c + 1]) ++c;
            gather_one<R, DIM, false, CONT, MOM, COUPLED>(g, C, A, s_ibeg[c] + (ii - s_ipre[c]));
        }
        return;
    }
    // ---- rebase boundaries to virtual offsets
    for (int t = tid; t < NR * W; t += NT) {
        const int q = t / W;
        s_cs[t] = s_voff[q] + (s_cs[t] - s_gbeg[q]);
    }
    // ---- stage candidates: flattened over the concatenated runs, coalesced inside each run
    const R ox = (R)s_org[0], oy = (R)s_org[1], oz = (R)s_org[2];
    const float far_lim = (float)(G + 6) * (float)g.cell;
    bool far = false;
    for (int v = tid; v < M; v += NT) {
        int q = 0;
        while (q + 1 < NR && v >= s_voff[q + 1]) ++q;
        const int gj = s_gbeg[q] + (v - s_voff[q]);
        float4 p;
        p.x = (float)(A.x[gj] - ox); p.y = (float)(A.y[gj] - oy); p.z = DIM == 3 ? (float)(A.z[gj] - oz) : 0.0f;
        p.w = __int_as_float(gj);
        if (LOCAL) far |= fabsf(p.x) > far_lim || fabsf(p.y) > far_lim || fabsf(p.z) > far_lim;
#if !defined(PST_P1_AOS)
        s_x[v] = p.x; s_y[v] = p.y; if (DIM == 3) s_z[v] = p.z;
        s_gj[v] = gj;
#else
        s_p4[v] = p;
#endif
    }
    far = __syncthreads_or(far);

    const int LC = T.lcap;
    for (int ii0 = 0; ii0 < ni; ii0 += NT) {   // uniform trip count: rounds of NT particles
        const int ii = ii0 + tid;
        const bool active = ii < ni;
        int c = 0, gi = 0, lx = 0, ly = 0, lf = 0;
        IState<R, DIM> I;
        Acc<R> a{0, 0, 0, 0}, a2{0, 0, 0, 0};
        float xf = 0, yf = 0, zf = 0, rc2f = 0;
        bool fluid_i = true;
        if (active) {
            while (c + 1 < NI && ii >= s_ipre[c + 1]) ++c;
            gi = s_ibeg[c] + (ii - s_ipre[c]);
            if (COUPLED) fluid_i = A.m[gi] > (R)0;
            lx = c / BB; ly = c - lx * BB;
            load_i<R, DIM>(I, C, A.x[gi], A.y[gi], DIM == 3 ? A.z[gi] : (R)0, A.u[gi], A.v[gi], DIM == 3 ? A.w[gi] : (R)0, A.rho[gi],
                           A.por2[gi], A.h[gi]);
            const int cf = DIM == 3 ? cell_coord<R>(I.z, g.lo[2], g.inv[2], 0, g.n[2] - 1) : cell_coord<R>(I.y, g.lo[1], g.inv[1], 0, g.n[1] - 1);
            lf = cf - f0;   // 0 .. G-1
            xf = (float)(I.x - ox); yf = (float)(I.y - oy); zf = DIM == 3 ? (float)(I.z - oz) : 0.0f;
            if (LOCAL) rc2f = far ? __int_as_float(0x7f800000) : __double2float_ru((double)I.rc2 * (1.0 + 1.0 / 32768.0));
            else rc2f = (float)I.rc2;
        }
        auto test = [&](const float4& p) -> bool {
            const float dxf = xf - p.x, dyf = yf - p.y, dzf = zf - p.z;
            float r2f;
            if (LOCAL) { r2f = dxf * dxf + dyf * dyf; if (DIM == 3) r2f += dzf * dzf; return r2f <= rc2f; }
            r2f = dist2<DIM, float>(dxf, dyf, dzf);
            return r2f < rc2f && r2f > 0.0f;
        };
#if !defined(PST_P1_AOS)
        // two candidates per instruction: FADD2 / FMUL2 / FFMA2 on register pairs (x_k, x_k+1) as LDS.128 delivers them
        const float2 xf2 = make_float2(xf, xf), yf2 = make_float2(yf, yf), zf2 = make_float2(zf, zf);
        auto r2_pair = [&](float xa, float xb, float ya, float yb, float za, float zb) -> float2 {
            const float2 dx = __fadd2_rn(xf2, make_float2(-xa, -xb)), dy = __fadd2_rn(yf2, make_float2(-ya, -yb));
            const float2 dz = DIM == 3 ? __fadd2_rn(zf2, make_float2(-za, -zb)) : make_float2(0.0f, 0.0f);
            if (LOCAL) {
                float2 r = __ffma2_rn(dy, dy, __fmul2_rn(dx, dx));
                if (DIM == 3) r = __ffma2_rn(dz, dz, r);
                return r;
            }
            float2 r = __fadd2_rn(__fmul2_rn(dx, dx), __fmul2_rn(dy, dy));     // exact test of f32 contexts: left to right, no FMA
            if (DIM == 3) r = __fadd2_rn(r, __fmul2_rn(dz, dz));
            return r;
        };
        auto hit = [&](float r2f) -> bool { return LOCAL ? r2f <= rc2f : (r2f < rc2f && r2f > 0.0f); };
        auto test_at = [&](int j) -> bool { return test(make_float4(s_x[j], s_y[j], DIM == 3 ? s_z[j] : 0.0f, 0.0f)); };
#endif
        // scan cursor over this particle's 9 (3) runs
        int run = 0, jv = 0, jend = 0;
        constexpr int NRUN = DIM == 3 ? 9 : 3;
        auto open_run = [&](int k) {
            const int ax = DIM == 3 ? k / 3 : k, ay = DIM == 3 ? k - ax * 3 : 0;
            const int q = (lx + ax) * RY + (DIM == 3 ? ly + ay : 0);
            jv = s_cs[q * W + lf];
            jend = s_cs[q * W + lf + 3];
        };
        bool done = !active;
        if (active) open_run(0);
        unsigned short* const my_list = s_list + tid;
        while (true) {
            // ---- phase 1: pre-filter on staged f32 coordinates, compact hits into the private list
            int cnt = 0;
            if (!done) {
                // private list = column `tid` of s_list (stride NT): one running shared-memory byte address
                const unsigned la0 = (unsigned)__cvta_generic_to_shared(my_list), lend = la0 + (unsigned)(LC * NT * 2);
                unsigned la = la0;
                // (no per-store "memory" clobber: it would pin the next LDS.128 batch behind every append; one compiler
                // barrier after the scan orders the list stores before phase 2 reads them)
                auto push = [&](int v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(la), "h"((unsigned short)v)); la += 2 * NT; };
                while (true) {
#if !defined(PST_P1_AOS)
                    while ((jv & 3) && jv < jend && la < lend) {             // head: up to the next 16-byte boundary of the SoA rows
                        if (test_at(jv)) push(jv);
                        ++jv;
                    }
                    while (jv + 4 <= jend && la + 8 * NT <= lend) {          // 4 candidates = 3 LDS.128 + 2 x (3 FADD2, FMUL2, 2 FFMA2)
                        const float4 X = *reinterpret_cast<const float4*>(s_x + jv), Y = *reinterpret_cast<const float4*>(s_y + jv);
                        const float4 Z = DIM == 3 ? *reinterpret_cast<const float4*>(s_z + jv) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        const float2 ra = r2_pair(X.x, X.y, Y.x, Y.y, Z.x, Z.y), rb = r2_pair(X.z, X.w, Y.z, Y.w, Z.z, Z.w);
                        const bool h0 = hit(ra.x), h1 = hit(ra.y), h2 = hit(rb.x), h3 = hit(rb.y);
                        if (h0) push(jv);
                        if (h1) push(jv + 1);
                        if (h2) push(jv + 2);
                        if (h3) push(jv + 3);
                        jv += 4;
                    }
                    while (jv < jend && la < lend) {
                        if (test_at(jv)) push(jv);
                        ++jv;
                    }
#else
                    while (jv + 4 <= jend && la + 8 * NT <= lend) {          // 4 candidates in flight
                        const float4 p0 = s_p4[jv], p1 = s_p4[jv + 1], p2 = s_p4[jv + 2], p3 = s_p4[jv + 3];
                        const bool h0 = test(p0), h1 = test(p1), h2 = test(p2), h3 = test(p3);
                        if (h0) push(jv);
                        if (h1) push(jv + 1);
                        if (h2) push(jv + 2);
                        if (h3) push(jv + 3);
                        jv += 4;
                    }
                    while (jv < jend && la < lend) {
                        if (test(s_p4[jv])) push(jv);
                        ++jv;
                    }
#endif
                    if (jv < jend) break;          // list full: drain, then resume here
                    if (++run == NRUN) { done = true; break; }
                    open_run(run);
                }
                cnt = (int)((la - la0) / (2 * NT));
                asm volatile("" ::: "memory");
            }
            // ---- phase 2: two hits per trip, j state gathered from global memory (L1/L2 hits), exact test, branch-free body
            for (int k = 0; k < cnt; k += 2) {
                const bool v1 = k + 1 < cnt;
#if !defined(PST_P1_AOS)
                const int j0 = s_gj[my_list[k * NT]];
                const int j1 = v1 ? s_gj[my_list[(k + 1) * NT]] : j0;
#else
                const int j0 = __float_as_int(s_p4[my_list[k * NT]].w);
                const int j1 = v1 ? __float_as_int(s_p4[my_list[(k + 1) * NT]].w) : j0;
#endif
#if defined(PST_EXP_SMEM) && !defined(PST_P1_AOS)
#error "PST_EXP_SMEM (timing experiment) reads the float4 staging buffer: build it together with -DPST_P1_AOS"
#endif
#if defined(PST_EXP_SMEM)    // timing experiment only (wrong results): the 9 gathers come from shared memory
                auto LD = [&](const R*, int j, int f) -> R { return (R)reinterpret_cast<const double*>(s_p4)[(j * 9 + f) & 2047]; };
#else
                auto LD = [&](const R* arr, int j, int) -> R { return arr[j]; };
#endif
                const R dx0 = I.x - LD(A.x, j0, 0), dy0 = I.y - LD(A.y, j0, 1), dz0 = DIM == 3 ? I.z - LD(A.z, j0, 2) : (R)0;
                const R dx1 = I.x - LD(A.x, j1, 0), dy1 = I.y - LD(A.y, j1, 1), dz1 = DIM == 3 ? I.z - LD(A.z, j1, 2) : (R)0;
                R r20 = dist2<DIM, R>(dx0, dy0, dz0), r21 = dist2<DIM, R>(dx1, dy1, dz1);
                const bool in0 = r20 < I.rc2 && r20 > (R)0;            // the exact test (the set is defined here)
                const bool in1 = v1 && r21 < I.rc2 && r21 > (R)0;
                r20 = in0 ? r20 : (R)1; r21 = in1 ? r21 : (R)1;
                R m0 = in0 ? (UMASS ? A.m_uni : LD(A.m, j0, 3)) : (R)0, m1 = in1 ? (UMASS ? A.m_uni : LD(A.m, j1, 3)) : (R)0;
                if (COUPLED) {   // signed SPH mass: the pair counts iff i or j is fluid
                    m0 = (fluid_i || m0 > (R)0) ? fabs(m0) : (R)0;
                    m1 = (fluid_i || m1 > (R)0) ? fabs(m1) : (R)0;
                }
                pair_body<R, DIM, CONT, MOM>(C, I, dx0, dy0, dz0, r20, LD(A.u, j0, 4), LD(A.v, j0, 5), DIM == 3 ? LD(A.w, j0, 6) : (R)0, LD(A.rho, j0, 7), m0, LD(A.por2, j0, 8), a);
                pair_body<R, DIM, CONT, MOM>(C, I, dx1, dy1, dz1, r21, LD(A.u, j1, 4), LD(A.v, j1, 5), DIM == 3 ? LD(A.w, j1, 6) : (R)0, LD(A.rho, j1, 7), m1, LD(A.por2, j1, 8), a2);
            }
            if (__all_sync(0xffffffffu, done)) break;
        }
        if (active) {
            a.au += a2.au; a.av += a2.av; a.aw += a2.aw; a.arho += a2.arho;
            store_acc<R, DIM, CONT, MOM>(A, C, gi, a);
        }
    }
}


// ---------------------------------------------------------------------------------------------
// variant 1: shared-memory tiles, ONE WARP PER CELL, candidate-parallel scan
//
//   * tile staging of the full f64 neighbour state (the (A+2)(B+2) candidate runs, coalesced, once per CTA);
//   * a warp takes one cell at a time.  All particles of a cell share the same 9 (3) candidate runs,
//     so the warp flattens them ONCE into a per-lane register cache: lane l holds candidates
//     l, l+32, l+64, ... as f32 coordinates (tile-local for f64 contexts) plus their staged index;
//   * per particle i (warp-uniform, broadcast from shared memory) the 32 lanes test 32 candidates
//     per step -- no divergence, no shared-memory traffic -- and ballot-compact the hits into a
//     per-warp queue;  f64 contexts use the f32 distance only as a CONSERVATIVE pre-filter
//     (relative margin 2^-15 >> the 1.5e-6 worst-case f32 error for coordinates within 4 cells of
//     the origin; cells holding farther particles switch the filter off), the exact FMA-free f64
//     test runs on the survivors, so the neighbour set is still bit-exact;
//   * the queue is drained 32 hits at a time: every lane evaluates the pair body for a different j
//     of the SAME i, partial sums are combined with a 6-step transposing butterfly.
// ---------------------------------------------------------------------------------------------
// phase 1 of the warp-per-cell kernel: NS steps of 32 candidates each, straight-line.
template <int DIM, bool LOCAL, int NS>
__device__ __forceinline__ int scan_steps(const float (&jx)[16], const float (&jy)[16], const float (&jz)[16], float xf, float yf, float zf,
                                          float rc2f, unsigned lt_mask, int lane, unsigned short* __restrict__ q_w) {
    unsigned m[NS];
    bool hit[NS];
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        const float dxf = xf - jx[t], dyf = yf - jy[t], dzf = zf - jz[t];
        float r2f;
        if (LOCAL) { r2f = dxf * dxf + dyf * dyf; if (DIM == 3) r2f += dzf * dzf; }
        else r2f = dist2<DIM, float>(dxf, dyf, dzf);
        hit[t] = LOCAL ? (r2f <= rc2f) : (r2f < rc2f && r2f > 0.0f);
        m[t] = __ballot_sync(0xffffffffu, hit[t]);
    }
    int tail = 0;
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        if (hit[t]) q_w[tail + __popc(m[t] & lt_mask)] = (unsigned short)(t * 32 + lane);
        tail += __popc(m[t]);
    }
    return tail;
}

// staging capacity of the warp-per-cell kernel: 113 KB per CTA (two CTAs per SM) minus 1 KB of tables minus the
// per-warp queue and slot->index tables (2 x 512 x u16 each)
template <class R, int NT>
__host__ __device__ constexpr int kCellwarpJcap() { return (int)((113 * 1024 - 1024 - (NT / 32) * 512 * 4) / (9 * sizeof(R))) & ~1; }

template <class R, int DIM, int TA, int TB, int NT, bool CONT, bool MOM>
__global__ void __launch_bounds__(NT, 384 / NT) k_wcsph_cellwarp(GridDev<R> g, WcsphConst<R> C, ForceArgs<R> A, TileShape T) {
    using D = TileDims<DIM, TA, TB>;
    constexpr int NR = D::NR, NI = D::NI, RY = D::RY, BB = D::BB;
    constexpr int NW = NT / 32, MAXT = 16, QCAP = 32 * MAXT;
    constexpr int NRUN = DIM == 3 ? 9 : 3;
    constexpr bool LOCAL = sizeof(R) == 8;       // f64: f32 pre-filter in local coordinates
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int G = T.G, W = G + 3;
    int* s_cs = reinterpret_cast<int*>(smem_raw);            // NR * W   virtual cell boundaries
    int* s_gbeg = s_cs + NR * W;                              // NR
    int* s_voff = s_gbeg + NR;                                // NR + 1
    int* s_next = s_voff + NR + 1;                            // 1        next cell to hand out
    size_t off = ((size_t)(NR * W + NR + NR + 1 + 1) * sizeof(int) + 15) & ~(size_t)15;
    unsigned short* s_q = reinterpret_cast<unsigned short*>(smem_raw + off);   // NW * QCAP
    off += (size_t)NW * QCAP * sizeof(unsigned short);
    unsigned short* s_jid = reinterpret_cast<unsigned short*>(smem_raw + off);   // NW * QCAP: staged index of cached candidate (t, lane)
    off += (size_t)NW * QCAP * sizeof(unsigned short);
    R* s_x = reinterpret_cast<R*>(smem_raw + off);
    constexpr int JC = kCellwarpJcap<R, NT>();   // compile-time, so the 9 staged arrays are one base + immediates
    R* s_y = s_x + JC; R* s_z = s_y + JC; R* s_u = s_z + JC; R* s_v = s_u + JC; R* s_w = s_v + JC;
    R* s_rho = s_w + JC; R* s_m = s_rho + JC; R* s_por2 = s_m + JC;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int b = blockIdx.x;
    const int tf = b % T.tiles[2]; b /= T.tiles[2];
    const int ty = DIM == 3 ? b % T.tiles[1] : 0; if (DIM == 3) b /= T.tiles[1];
    const int tx = b;
    const int cx0 = tx * TA, cy0 = ty * BB, f0 = tf * G;
    const int nf = DIM == 3 ? g.n[2] : g.n[1];
    const int ncx = g.n[0], ncy = DIM == 3 ? g.n[1] : 1;

    for (int t = tid; t < NR * W; t += NT) {
        const int q = t / W, tt = t - q * W;
        const int rx = q / RY, ry = q - rx * RY;
        const int cx = cx0 - 1 + rx, cy = DIM == 3 ? cy0 - 1 + ry : 0;
        int gi = 0;
        if (cx >= 0 && cx < ncx && cy >= 0 && cy < ncy) {
            const int col = DIM == 3 ? cx * ncy + cy : cx;
            gi = A.cell_start[(size_t)col * nf + min(max(f0 - 1 + tt, 0), nf)];
        }
        s_cs[t] = gi;
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int q = 0; q < NR; ++q) {
            s_gbeg[q] = s_cs[q * W];
            s_voff[q] = acc;
            acc += s_cs[q * W + W - 1] - s_cs[q * W];
        }
        s_voff[NR] = acc;
        *s_next = NW;
    }
    __syncthreads();
    // number of i particles in the tile (sum over the centre runs, cells f0 .. f0+G-1)
    int ni = 0;
    for (int c = 0; c < NI; ++c) {
        const int lx = c / BB, ly = c - lx * BB;
        const int q = (lx + 1) * RY + (DIM == 3 ? ly + 1 : 0);
        if (cx0 + lx >= g.cx_lo && cx0 + lx <= g.cx_hi && cy0 + ly < ncy) ni += s_cs[q * W + G + 1] - s_cs[q * W + 1];
    }
    if (ni == 0) return;
    const int M = s_voff[NR];
    if (M > min(JC, T.jcap)) {   // tile denser than the staging buffer: exact per-particle gather for its particles
        for (int c = 0; c < NI; ++c) {
            const int lx = c / BB, ly = c - lx * BB;
            const int q = (lx + 1) * RY + (DIM == 3 ? ly + 1 : 0);
            if (!(cx0 + lx >= g.cx_lo && cx0 + lx <= g.cx_hi && cy0 + ly < ncy)) continue;
            for (int s = s_cs[q * W + 1] + tid; s < s_cs[q * W + G + 1]; s += NT) gather_one<R, DIM, false, CONT, MOM>(g, C, A, s);
        }
        return;
    }
    __syncthreads();   // everyone has read the global-index form of s_cs
    for (int t = tid; t < NR * W; t += NT) {
        const int q = t / W;
        s_cs[t] = s_voff[q] + (s_cs[t] - s_gbeg[q]);
    }
    for (int v = tid; v < M; v += NT) {
        int q = 0;
        while (q + 1 < NR && v >= s_voff[q + 1]) ++q;
        const int gj = s_gbeg[q] + (v - s_voff[q]);
        s_x[v] = A.x[gj]; s_y[v] = A.y[gj]; if (DIM == 3) s_z[v] = A.z[gj];
        s_u[v] = A.u[gj]; s_v[v] = A.v[gj]; if (DIM == 3) s_w[v] = A.w[gj];
        s_rho[v] = A.rho[gj]; s_m[v] = A.m[gj]; s_por2[v] = A.por2[gj];
    }
    __syncthreads();

    unsigned short* q_w = s_q + warp * QCAP;
    unsigned short* jid_w = s_jid + warp * QCAP;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float far_lim = 4.0f * (float)g.cell;
    // result routing after the butterfly: lane 0 au, 8 av, 16 aw, 24 arho
    R* const out_ptr = lane < 8 ? A.au : lane < 16 ? A.av : lane < 24 ? A.aw : A.arho;
    const R out_g = lane < 8 ? C.g[0] : lane < 16 ? C.g[1] : lane < 24 ? C.g[2] : (R)0;
    const bool out_lane = (lane & 7) == 0 && (lane < 24 ? (MOM && (DIM == 3 || lane < 16)) : CONT);
    IState<R, DIM> I;
    I.h = (R)-1;
    float rc2f_h = 0.0f;
    for (int c = warp;;) {     // cells of the tile, handed out dynamically (warp-uniform)
        if (c >= NI * G) break;
        const int col = c / G, lf = c - col * G;
        const int lx = col / BB, ly = col - lx * BB;
        const int q0 = (lx + 1) * RY + (DIM == 3 ? ly + 1 : 0);
        int ib = 0, ie = 0;
        if (cx0 + lx >= g.cx_lo && cx0 + lx <= g.cx_hi && cy0 + ly < ncy && f0 + lf < nf) { ib = s_cs[q0 * W + lf + 1]; ie = s_cs[q0 * W + lf + 2]; }
        if (ib < ie) {
            // the cell's candidate runs, flattened
            int rb[NRUN], rp[NRUN + 1];
            rp[0] = 0;
#pragma unroll
            for (int k = 0; k < NRUN; ++k) {
                const int ax = DIM == 3 ? k / 3 : k, ay = DIM == 3 ? k - ax * 3 : 0;
                const int q = (lx + ax) * RY + (DIM == 3 ? ly + ay : 0);
                rb[k] = s_cs[q * W + lf];
                rp[k + 1] = rp[k] + (s_cs[q * W + lf + 3] - rb[k]);
            }
            const int total = rp[NRUN];
            const int gi0 = s_gbeg[q0] - s_voff[q0];
            const R ox = LOCAL ? s_x[ib] : (R)0, oy = LOCAL ? s_y[ib] : (R)0, oz = (LOCAL && DIM == 3) ? s_z[ib] : (R)0;
            for (int cbase = 0; cbase < total; cbase += 32 * MAXT) {
                // ---- per-lane register cache of this chunk's candidates (f32; tile-local for f64 contexts)
                float jx[MAXT], jy[MAXT], jz[MAXT];
                bool far = false;
                __syncwarp();
#pragma unroll
                for (int t = 0; t < MAXT; ++t) {
                    const int v = cbase + t * 32 + lane;
                    int jv = -1;
                    if (v < total) {
#pragma unroll
                        for (int k = 0; k < NRUN; ++k)
                            if (v >= rp[k] && v < rp[k + 1]) jv = rb[k] + (v - rp[k]);
                    }
                    jid_w[t * 32 + lane] = (unsigned short)jv;
                    if (jv >= 0) {
                        jx[t] = (float)(s_x[jv] - ox); jy[t] = (float)(s_y[jv] - oy); jz[t] = DIM == 3 ? (float)(s_z[jv] - oz) : 0.0f;
                        if (LOCAL) far |= fabsf(jx[t]) > far_lim || fabsf(jy[t]) > far_lim || fabsf(jz[t]) > far_lim;
                    } else {
                        jx[t] = jy[t] = jz[t] = __int_as_float(0x7fc00000);   // NaN: never a hit, whatever the cutoff
                    }
                }
                if (LOCAL) far = __any_sync(0xffffffffu, far);
                __syncwarp();
                const int nt = min(MAXT, (total - cbase + 31) >> 5);
                const bool first = cbase == 0;
                R h_next = A.h[gi0 + ib];
                for (int iv = ib; iv < ie; ++iv) {    // warp-uniform: one particle i at a time
                    const int gi = gi0 + iv;
                    const R hi = h_next;
                    if (iv + 1 < ie) h_next = A.h[gi + 1];           // prefetch: keeps the global-load latency off the chain
                    if (hi != I.h) {    // h-derived constants (uniform h: computed once per warp)
                        load_i<R, DIM>(I, C, (R)0, (R)0, (R)0, (R)0, (R)0, (R)0, (R)0, (R)0, hi);
                        rc2f_h = LOCAL ? __double2float_ru((double)I.rc2 * (1.0 + 1.0 / 32768.0)) : (float)I.rc2;
                    }
                    const float rc2f = (LOCAL && far) ? __int_as_float(0x7f800000) : rc2f_h;
                    I.x = s_x[iv]; I.y = s_y[iv]; I.z = DIM == 3 ? s_z[iv] : (R)0;
                    const float xf = (float)(I.x - ox), yf = (float)(I.y - oy), zf = DIM == 3 ? (float)(I.z - oz) : 0.0f;
                    // ---- phase 1: 32 candidates per step, ballot-compacted (slot numbers) into the warp's queue.
                    // One straight-line block per group count, so all its steps overlap; padding slots hold NaN.
                    int tail;
                    switch ((nt + 3) >> 2) {
                        case 1: tail = scan_steps<DIM, LOCAL, 4>(jx, jy, jz, xf, yf, zf, rc2f, lt_mask, lane, q_w); break;
                        case 2: tail = scan_steps<DIM, LOCAL, 8>(jx, jy, jz, xf, yf, zf, rc2f, lt_mask, lane, q_w); break;
                        case 3: tail = scan_steps<DIM, LOCAL, 12>(jx, jy, jz, xf, yf, zf, rc2f, lt_mask, lane, q_w); break;
                        default: tail = scan_steps<DIM, LOCAL, 16>(jx, jy, jz, xf, yf, zf, rc2f, lt_mask, lane, q_w); break;
                    }
                    __syncwarp();
                    // ---- phase 2: the pair body, TWO hits per lane per pass (independent chains -> ILP), branch-free:
                    // an empty slot evaluates a harmless dummy pair (r2 = 1, m_j = 0).
                    I.u = s_u[iv]; I.v = s_v[iv]; I.w = DIM == 3 ? s_w[iv] : (R)0;
                    I.rho = s_rho[iv]; I.por2 = s_por2[iv];
                    Acc<R> a{0, 0, 0, 0}, a2{0, 0, 0, 0};
                    for (int head = 0; head < tail; head += 64) {
                        const bool v0 = head + lane < tail, v1 = head + 32 + lane < tail;
                        const int j0 = v0 ? jid_w[q_w[head + lane]] : iv;
                        const int j1 = v1 ? jid_w[q_w[head + 32 + lane]] : iv;
                        const R dx0 = I.x - s_x[j0], dy0 = I.y - s_y[j0], dz0 = DIM == 3 ? I.z - s_z[j0] : (R)0;
                        const R dx1 = I.x - s_x[j1], dy1 = I.y - s_y[j1], dz1 = DIM == 3 ? I.z - s_z[j1] : (R)0;
                        R r20 = dist2<DIM, R>(dx0, dy0, dz0), r21 = dist2<DIM, R>(dx1, dy1, dz1);
                        const bool in0 = v0 && r20 < I.rc2 && r20 > (R)0;     // the exact test (the set is defined here)
                        const bool in1 = v1 && r21 < I.rc2 && r21 > (R)0;
                        r20 = in0 ? r20 : (R)1; r21 = in1 ? r21 : (R)1;
                        const R m0 = in0 ? s_m[j0] : (R)0, m1 = in1 ? s_m[j1] : (R)0;
                        pair_body<R, DIM, CONT, MOM>(C, I, dx0, dy0, dz0, r20, s_u[j0], s_v[j0], DIM == 3 ? s_w[j0] : (R)0, s_rho[j0], m0, s_por2[j0], a);
                        pair_body<R, DIM, CONT, MOM>(C, I, dx1, dy1, dz1, r21, s_u[j1], s_v[j1], DIM == 3 ? s_w[j1] : (R)0, s_rho[j1], m1, s_por2[j1], a2);
                    }
                    a.au += a2.au; a.av += a2.av; a.aw += a2.aw; a.arho += a2.arho;
                    __syncwarp();
                    // ---- combine the 32 partial sums: transposing butterfly, 6 shuffles for 4 values
                    const bool b4 = lane & 16, b3 = lane & 8;
                    const R s0 = b4 ? a.au : a.aw, s1 = b4 ? a.av : a.arho;
                    R k0 = b4 ? a.aw : a.au, k1 = b4 ? a.arho : a.av;
                    k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
                    k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                    const R sx = b3 ? k0 : k1;
                    R k = b3 ? k1 : k0;
                    k += __shfl_xor_sync(0xffffffffu, sx, 8);
                    k += __shfl_xor_sync(0xffffffffu, k, 4);
                    k += __shfl_xor_sync(0xffffffffu, k, 2);
                    k += __shfl_xor_sync(0xffffffffu, k, 1);
                    if (out_lane) out_ptr[gi] = first ? k + out_g : out_ptr[gi] + k;
                }
            }
        }
        int nc = 0;
        if (lane == 0) nc = atomicAdd(s_next, 1);
        c = __shfl_sync(0xffffffffu, nc, 0);
    }
}

// element 0 of the record buffer (the allocation starts ghost_cap rounded up to a block of 8 earlier)
template <class R>
R* rec_ptr(pst_ctx* ctx) {
    if (!ctx->rec) return nullptr;
    const size_t g8 = (ctx->ghost_cap + 7) & ~(size_t)7;
    return reinterpret_cast<R*>(ctx->rec) + g8 * kRecRows * 2;
}
inline bool rec_wanted(pst_ctx* ctx) {
    return !ctx->grid.morton && pst_option(ctx, "force_kernel", 3) == 3 && pst_option(ctx, "rec_impl", 1) == 1;
}
pst_status rec_alloc(pst_ctx* ctx) {
    if (ctx->rec) return PST_OK;
    if ((ctx->capacity + 2 * ctx->ghost_cap + 16) * kRecRows >= (1ull << 31)) return pst_fail(ctx, PST_EINVAL, "rec_impl = 1 supports up to 2^31 / 5 particles per context: set option rec_impl = 0");
    const size_t g8 = (ctx->ghost_cap + 7) & ~(size_t)7;
    const size_t elems = (g8 + ((ctx->capacity + ctx->ghost_cap + 7) & ~(size_t)7) + 8) * kRecRows * 2;
    if (cudaMalloc(&ctx->rec, elems * (ctx->f64 ? 8 : 4)) != cudaSuccess) return pst_fail(ctx, PST_ENOMEM, "neighbour-state records (%zu bytes)", elems * (ctx->f64 ? 8 : 4));
    // f32 positions: 3 rows, element 0 at a multiple of 4 (the tiles' TMA bulk copies need 16-byte aligned sources), 8 floats of slack
    ctx->posf_g4 = (ctx->ghost_cap + 3) & ~(size_t)3;
    ctx->posf_stride = (ctx->posf_g4 + ctx->capacity + ctx->ghost_cap + 8 + 3) & ~(size_t)3;
    if (cudaMalloc((void**)&ctx->posf, 3 * ctx->posf_stride * sizeof(float)) != cudaSuccess) return pst_fail(ctx, PST_ENOMEM, "f32 position rows");
    PST_CUDA(ctx, cudaMemsetAsync(ctx->posf, 0, 3 * ctx->posf_stride * sizeof(float), ctx->stream));
    return PST_OK;
}
template <class R>
PosF<R> pos_f(pst_ctx* ctx) {
    PosF<R> F;
    F.x = F.y = F.z = nullptr;
    const PstGrid& g = ctx->grid;
    for (int a = 0; a < 3; ++a) { F.lo[a] = (R)g.lo[a]; F.cmax[a] = (float)(g.n[a] / (a == g.dim - 1 ? g.sub : 1) + 2) * (float)g.cell; }
    F.cmin = -2.0f * (float)g.cell;
    if (ctx->posf) {
        F.x = ctx->posf + ctx->posf_g4; F.y = F.x + ctx->posf_stride;
        F.z = g.dim == 3 ? F.y + ctx->posf_stride : nullptr;
    }
    return F;
}
template <class R>
RecSrc<R> rec_src(pst_ctx* ctx) {
    RecSrc<R> S;
    S.x = pst_ptr<R>(ctx, "x"); S.y = pst_ptr<R>(ctx, "y"); S.z = pst_ptr<R>(ctx, "z");
    S.u = pst_ptr<R>(ctx, "u"); S.v = pst_ptr<R>(ctx, "v"); S.w = pst_ptr<R>(ctx, "w");
    S.rho = pst_ptr<R>(ctx, "rho"); S.por2 = pst_ptr<R>(ctx, "por2"); S.m = pst_ptr<R>(ctx, ctx->coupled ? "msph" : "m"); S.h = pst_ptr<R>(ctx, "h");
    S.F = pos_f<R>(ctx);
    return S;
}

template <class R>
ForceArgs<R> make_args(pst_ctx* ctx) {
    ForceArgs<R> A;
    A.x = pst_ptr<R>(ctx, "x"); A.y = pst_ptr<R>(ctx, "y"); A.z = pst_ptr<R>(ctx, "z");
    A.u = pst_ptr<R>(ctx, "u"); A.v = pst_ptr<R>(ctx, "v"); A.w = pst_ptr<R>(ctx, "w");
    A.rho = pst_ptr<R>(ctx, "rho"); A.m = pst_ptr<R>(ctx, ctx->coupled ? "msph" : "m"); A.h = pst_ptr<R>(ctx, "h"); A.por2 = pst_ptr<R>(ctx, "por2");
    A.au = pst_ptr<R>(ctx, "au"); A.av = pst_ptr<R>(ctx, "av"); A.aw = pst_ptr<R>(ctx, "aw"); A.arho = pst_ptr<R>(ctx, "arho");
    A.cell_start = ctx->cell_start;
    A.n = (int)ctx->n;
    A.m_uni = (R)ctx->m_value;
    A.rec = rec_ptr<R>(ctx);
    A.F = pos_f<R>(ctx);
    return A;
}

template <class R, int DIM, bool MORTON>
pst_status launch_gather(pst_ctx* ctx, bool cont, bool mom) {
    const GridDev<R> g = make_grid_dev<R>(ctx->grid);
    const WcsphConst<R> C = make_const<R>(ctx);
    const ForceArgs<R> A = make_args<R>(ctx);
    const unsigned grid = blocks_for(ctx->n, kThreads);
    if (ctx->coupled) {
        if (DIM == 3) PST_LAUNCH(ctx, (k_wcsph_gather<R, 3, MORTON, true, true, true>), grid, kThreads, 0, g, C, A);
        return PST_OK;
    }
    if (cont && mom) PST_LAUNCH(ctx, (k_wcsph_gather<R, DIM, MORTON, true, true>), grid, kThreads, 0, g, C, A);
    else if (cont) PST_LAUNCH(ctx, (k_wcsph_gather<R, DIM, MORTON, true, false>), grid, kThreads, 0, g, C, A);
    else PST_LAUNCH(ctx, (k_wcsph_gather<R, DIM, MORTON, false, true>), grid, kThreads, 0, g, C, A);
    return PST_OK;
}

template <class R, int DIM, int TA, int TB, int NT, int VARIANT, bool CONT, bool MOM, bool COUPLED = false, bool UMASS = false>
pst_status launch_tiled_k(pst_ctx* ctx, const TileShape& T, size_t smem) {
    void (*kern)(GridDev<R>, WcsphConst<R>, ForceArgs<R>, TileShape);
    if (VARIANT == 1) kern = k_wcsph_cellwarp<R, DIM, TA, TB, NT, CONT, MOM>;
    else kern = k_wcsph_tiled<R, DIM, TA, TB, NT, CONT, MOM, COUPLED, UMASS>;
    PST_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const unsigned grid = (unsigned)T.tiles[0] * T.tiles[1] * T.tiles[2];
    PST_LAUNCH(ctx, kern, grid, NT, smem, make_grid_dev<R>(ctx->grid), make_const<R>(ctx), make_args<R>(ctx), T);
    return PST_OK;
}

// VARIANT 1: warp-per-cell (no per-thread lists); VARIANT 2: thread-per-particle with private hit lists.  Both 2 CTAs/SM.
template <class R, int DIM, int VARIANT, int TA>
pst_status launch_tiled(pst_ctx* ctx, bool cont, bool mom) {
    constexpr int TB = 2, NT = VARIANT == 1 ? 192 : 256;
    using D = TileDims<DIM, TA, TB>;
    const PstGrid& g = ctx->grid;
    const int nf = DIM == 3 ? g.n[2] : g.n[1];
    double ppc = pst_param(ctx, "_ppc", 0.0);      // mean occupancy of the occupied cells (measured by k_bounds)
    if (!(ppc > 0)) ppc = DIM == 3 ? 14.0 : 6.0;
    TileShape T;
    T.lcap = pst_option(ctx, "tile_lcap", DIM == 3 ? 80 : 32);
    const int jc_max = VARIANT == 1 ? kCellwarpJcap<R, NT>() : kListsJcap;
    const int user_G = pst_option(ctx, "tile_g", 0);
    int G = user_G;
    if (G <= 0) {
        if (VARIANT == 1) G = (int)std::floor(jc_max / (1.12 * D::NR * ppc)) - 2;   // deepest tile that fits with 12 % headroom
        else G = (int)std::floor(0.95 * NT / (D::NI * ppc));                        // one thread per particle
    }
    G = std::min(std::max(G, 1), std::max(1, nf));
    // f32 pre-filter of f64 contexts: tile-local coordinates reach (G + 6) cells, where one f32 ulp is (G + 6) 2^-23 cells; the
    // worst-case relative error of r2f is 2 sqrt(3) ulp / cutoff.  (G + 6) <= 32 keeps it below half of the 2^-15 margin
    // (1.3e-5 < 1.5e-5; tests/test_prefilter_margin.py), so a true neighbour can never be filtered out.
    G = std::min(G, kMaxTileG);
    G = std::min(G, (256 - 2 * D::NR - 2 * D::NI - 3) / D::NR - 3);   // boundary tables stay within 1 KB
    while (user_G <= 0 && G > 1 && 1.08 * D::NR * (G + 2) * ppc > jc_max) --G;   // the staged runs must fit (else: slow exact fallback)
    T.G = G;
    T.tiles[0] = (g.n[0] + TA - 1) / TA;
    T.tiles[1] = DIM == 3 ? (g.n[1] + D::BB - 1) / D::BB : 1;
    T.tiles[2] = (nf + G - 1) / G;
    // "tile_jcap" (tests): pretend the staging buffer is smaller, to force the in-kernel exact fallback
    T.jcap = std::min(jc_max, std::max(0, pst_option(ctx, "tile_jcap", jc_max)));
    const size_t ints = ((size_t)(D::NR * (G + 3) + D::NR + D::NR + 1 + D::NI + D::NI + 1) * sizeof(int) + 15) & ~(size_t)15;
    const size_t smem = VARIANT == 1 ? (size_t)113 * 1024
                                     : ints + 4 * sizeof(double) + (size_t)kListsJcap * sizeof(float4) + (size_t)T.lcap * NT * sizeof(unsigned short);
    if (smem > 227 * 1024) return pst_fail(ctx, PST_EINVAL, "tile_lcap too large");
    if (ctx->coupled) {
        if (DIM == 3 && VARIANT == 2) return launch_tiled_k<R, 3, TA, TB, NT, 2, true, true, true>(ctx, T, smem);
        return pst_fail(ctx, PST_EINVAL, "coupled contexts need dim = 3 and force_kernel 0 or 2");
    }
    // all masses equal (seen at upload): the fused kernel without the m[j] gather.  With a communicator the ghosts and the
    // migrants come from other ranks, whose uploads this rank has not seen: the caller vouches for them with the option
    // "uniform_mass_global" = 1 (every rank uploaded the same single mass value; bench.py checks it with an all-reduce).
    PST_TRY(pst_uniform_refresh(ctx));
    const bool umass = VARIANT == 2 && ctx->m_uniform && (!ctx->comm || pst_option(ctx, "uniform_mass_global", 0) != 0) &&
                       pst_option(ctx, "uniform_mass", 1) != 0;
    if constexpr (VARIANT == 2) {
        if (cont && mom && umass) return launch_tiled_k<R, DIM, TA, TB, NT, 2, true, true, false, true>(ctx, T, smem);
    }
    if (cont && mom) return launch_tiled_k<R, DIM, TA, TB, NT, VARIANT, true, true>(ctx, T, smem);
    if (cont) return launch_tiled_k<R, DIM, TA, TB, NT, VARIANT, true, false>(ctx, T, smem);
    return launch_tiled_k<R, DIM, TA, TB, NT, VARIANT, false, true>(ctx, T, smem);
}

// bring the packed records up to date (no-op when the EOS pass has just written them)
template <class R>
pst_status rec_refresh(pst_ctx* ctx) {
    if (ctx->rec && ctx->rec_epoch == ctx->state_epoch) {
        if (ctx->ghost_eos_pending) {       // ghost rows arrived after the fused permute and no EOS pass has run over them yet (continuity alone)
            const int n = (int)ctx->n, nl = (int)ctx->n_ghost_l, nr = (int)ctx->n_ghost_r;
            if (nl > 0) PST_LAUNCH(ctx, k_rec_pack<R>, blocks_for(nl, 256), 256, 0, -nl, 0, rec_ptr<R>(ctx), rec_src<R>(ctx));
            if (nr > 0) PST_LAUNCH(ctx, k_rec_pack<R>, blocks_for(nr, 256), 256, 0, n, n + nr, rec_ptr<R>(ctx), rec_src<R>(ctx));
        }
        return PST_OK;
    }
    PST_TRY(rec_alloc(ctx));
    const int lo = -(int)ctx->n_ghost_l, hi = (int)ctx->n + (int)ctx->n_ghost_r;
    if (hi > lo) PST_LAUNCH(ctx, k_rec_pack<R>, blocks_for(hi - lo, 256), 256, 0, lo, hi, rec_ptr<R>(ctx), rec_src<R>(ctx));
    ctx->rec_epoch = ctx->state_epoch;
    return PST_OK;
}

#include "wcsph_zrun.cuh"   // variant 3: tiles + fine z-runs + bit masks

template <class R, int DIM, bool MORTON>
pst_status launch_forces(pst_ctx* ctx, bool cont, bool mom) {
    int variant = pst_option(ctx, "force_kernel", 3);   // 3 = tiled z-runs + bit masks (default), 2 = thread-per-particle lists, 1 = warp-per-cell, 0 = gather
    if (ctx->coupled && variant == 1) variant = 2;      // the warp-per-cell kernel has no coupled form
#ifdef PST_DEV_HEADLINE     // kernel-development build (make dev): only the headline instantiation, compiles in seconds
    if constexpr (sizeof(R) == 8 && DIM == 3 && !MORTON) { if (variant == 3) return launch_zrun<R, DIM>(ctx, cont, mom); }
    return pst_fail(ctx, PST_EINVAL, "development build: only force_kernel 3, f64, 3D, linear keys");
#else
    if (variant == 3 && !MORTON) return launch_zrun<R, DIM>(ctx, cont, mom);
    if ((variant == 1 || variant == 2) && !MORTON && ctx->grid.sub != 1)
        return pst_fail(ctx, PST_EINVAL, "force_kernel %d needs zsub = 1 (its tiles are cut in whole cells)", variant);
    if (variant == 1 && !MORTON) return launch_tiled<R, DIM, 1, 2>(ctx, cont, mom);
    if (variant == 2 && !MORTON) return pst_option(ctx, "tile_ta", 2) == 3 ? launch_tiled<R, DIM, 2, 3>(ctx, cont, mom) : launch_tiled<R, DIM, 2, 2>(ctx, cont, mom);
    return launch_gather<R, DIM, MORTON>(ctx, cont, mom);
#endif
}

// semi-implicit Euler stage: fluid (tag 0): v += a dt, x += v dt;
// every particle: rho += arho dt.  Boundary particles (tag 1) keep position and velocity.
// `slaved` (boundary_model = 1): the density of non-fluid particles is set by wall_pressure, not integrated.
template <class R, int DIM>
__global__ void __launch_bounds__(256) k_integrate(int n, R dt, bool slaved, const int32_t* __restrict__ tag, R* __restrict__ x, R* __restrict__ y,
                                                   R* __restrict__ z, R* __restrict__ u, R* __restrict__ v, R* __restrict__ w,
                                                   R* __restrict__ rho, const R* __restrict__ au, const R* __restrict__ av,
                                                   const R* __restrict__ aw, const R* __restrict__ arho) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const bool fluid = tag[s] == 0;
    if (fluid || !slaved) rho[s] += arho[s] * dt;
    if (!fluid) return;
    const R un = u[s] + au[s] * dt, vn = v[s] + av[s] * dt;
    u[s] = un; v[s] = vn;
    x[s] += un * dt; y[s] += vn * dt;
    if (DIM == 3) {
        const R wn = w[s] + aw[s] * dt;
        w[s] = wn;
        z[s] += wn * dt;
    }
}

template <class R, int DIM, bool MORTON>
pst_status launch_integrate(pst_ctx* ctx, double dt) {
    const int n = (int)ctx->n;
    PST_LAUNCH(ctx, (k_integrate<R, DIM>), blocks_for(n, 256), 256, 0, n, (R)dt, pst_param(ctx, "boundary_model", 0.0) == 1.0, pst_ptr<int32_t>(ctx, "tag"), pst_ptr<R>(ctx, "x"),
               pst_ptr<R>(ctx, "y"), pst_ptr<R>(ctx, "z"), pst_ptr<R>(ctx, "u"), pst_ptr<R>(ctx, "v"), pst_ptr<R>(ctx, "w"),
               pst_ptr<R>(ctx, "rho"), pst_ptr<R>(ctx, "au"), pst_ptr<R>(ctx, "av"), pst_ptr<R>(ctx, "aw"), pst_ptr<R>(ctx, "arho"));
    return PST_OK;
}

// Coupled SPH-DEM stage (DESIGN.md "Coupled formulation"): every particle rho += arho dt; fluid (tag 0) v += a dt,
// x += v dt; solids (tag 2) v += (F_contact/m + (rho0/rho_solid)(a - g) + g) dt, x += v dt, omega += T/I dt -- the SPH
// sum `a` of a solid runs over its fluid neighbours only, so m_sph (a - g) is the hydrodynamic force on the sphere;
// boundaries (tag 1) keep position and velocity.
template <class R>
struct CoupledIntArgs {
    const int32_t* tag;
    R *x, *y, *z, *u, *v, *w, *rho, *wx, *wy, *wz;
    const R *m, *inertia, *au, *av, *aw, *arho, *fx, *fy, *fz, *tx, *ty, *tz;
    R dt, g[3], ratio;
    int n;
    bool slaved;   // boundary_model = 1: the density of non-fluid particles is set by wall_pressure, not integrated
};
template <class R>
__global__ void __launch_bounds__(256) k_coupled_integrate(CoupledIntArgs<R> A) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.n) return;
    const R dt = A.dt;
    const int t = A.tag[s];
    if (t == 0 || !A.slaved) A.rho[s] += A.arho[s] * dt;
    if (t == 1) return;
    R ax = A.au[s], ay = A.av[s], az = A.aw[s];
    if (t == 2) {
        const R im = (R)1 / A.m[s], ii = (R)1 / A.inertia[s];
        ax = (A.fx[s] * im + A.ratio * (ax - A.g[0])) + A.g[0];
        ay = (A.fy[s] * im + A.ratio * (ay - A.g[1])) + A.g[1];
        az = (A.fz[s] * im + A.ratio * (az - A.g[2])) + A.g[2];
        A.wx[s] += A.tx[s] * ii * dt; A.wy[s] += A.ty[s] * ii * dt; A.wz[s] += A.tz[s] * ii * dt;
    }
    const R un = A.u[s] + ax * dt, vn = A.v[s] + ay * dt, wn = A.w[s] + az * dt;
    A.u[s] = un; A.v[s] = vn; A.w[s] = wn;
    A.x[s] += un * dt; A.y[s] += vn * dt; A.z[s] += wn * dt;
}

template <class R>
pst_status launch_coupled_integrate(pst_ctx* ctx, double dt) {
    CoupledIntArgs<R> A;
    A.tag = pst_ptr<int32_t>(ctx, "tag");
    A.x = pst_ptr<R>(ctx, "x"); A.y = pst_ptr<R>(ctx, "y"); A.z = pst_ptr<R>(ctx, "z");
    A.u = pst_ptr<R>(ctx, "u"); A.v = pst_ptr<R>(ctx, "v"); A.w = pst_ptr<R>(ctx, "w"); A.rho = pst_ptr<R>(ctx, "rho");
    A.wx = pst_ptr<R>(ctx, "wx"); A.wy = pst_ptr<R>(ctx, "wy"); A.wz = pst_ptr<R>(ctx, "wz");
    A.m = pst_ptr<R>(ctx, "m"); A.inertia = pst_ptr<R>(ctx, "inertia");
    A.au = pst_ptr<R>(ctx, "au"); A.av = pst_ptr<R>(ctx, "av"); A.aw = pst_ptr<R>(ctx, "aw"); A.arho = pst_ptr<R>(ctx, "arho");
    A.fx = pst_ptr<R>(ctx, "fx"); A.fy = pst_ptr<R>(ctx, "fy"); A.fz = pst_ptr<R>(ctx, "fz");
    A.tx = pst_ptr<R>(ctx, "tx"); A.ty = pst_ptr<R>(ctx, "ty"); A.tz = pst_ptr<R>(ctx, "tz");
    A.dt = (R)dt;
    A.g[0] = (R)pst_param(ctx, "gx"); A.g[1] = (R)pst_param(ctx, "gy"); A.g[2] = (R)pst_param(ctx, "gz");
    A.ratio = (R)pst_param(ctx, "rho0") / (R)pst_param(ctx, "rho_solid", 1.0);
    A.n = (int)ctx->n;
    A.slaved = pst_param(ctx, "boundary_model", 0.0) == 1.0;
    PST_LAUNCH(ctx, k_coupled_integrate<R>, blocks_for(A.n, 256), 256, 0, A);
    return PST_OK;
}

template <class R>
pst_status launch_eos(pst_ctx* ctx, int lo, int hi) {
    if (hi <= lo) return PST_OK;
    const double rs = pst_param(ctx, "rho_solid", 0.0);
    if (ctx->coupled && !(rs > 0)) return pst_fail(ctx, PST_EINVAL, "coupled context: parameter rho_solid must be > 0");
    const bool pack = rec_wanted(ctx);
    if (pack) PST_TRY(rec_alloc(ctx));
    PST_LAUNCH(ctx, k_eos<R>, blocks_for(hi - lo, 256), 256, 0, make_const<R>(ctx), lo, hi, pst_ptr<R>(ctx, "rho"), pst_ptr<R>(ctx, "p"),
               pst_ptr<R>(ctx, "por2"), ctx->coupled ? pst_ptr<int32_t>(ctx, "tag") : nullptr, pst_ptr<R>(ctx, "m"),
               ctx->coupled ? pst_ptr<R>(ctx, "msph") : nullptr, ctx->coupled ? (R)pst_param(ctx, "rho0") / (R)rs : (R)0,
               pack ? rec_ptr<R>(ctx) : nullptr, rec_src<R>(ctx));
    if (pack) ctx->rec_epoch = ctx->state_epoch;
    return PST_OK;
}

// ---------------------------------------------------------------------------------------------
// Dummy-particle wall pressure (after Adami, Hu & Adams 2012).  No reference code exists
// for it; the formulation is restated in oracle/oracle.cpp:wall_pressure.  For every NON-fluid particle w (tag != 0),
// over its FLUID neighbours f (A.1 rule with the support of w):
//     S0 = sum W_wf,  Sp = sum p_f W_wf,  S = sum rho_f x_wf W_wf,   p_w = (Sp + g . S) / S0  (0 without fluid neighbours)
//     rho_w = rho0 (max(p_w / B, -1/2) + 1)^(1/gamma)
// and p, p/rho^2 and the STATE density of w are overwritten, so the fused pair kernel runs unchanged: it simply gathers
// the extrapolated values for dummy neighbours.  Runs after k_eos (reads p of fluid rows, writes non-fluid rows: no race).
//   k_wp_flags -> in-place exclusive scan (the counting sort's scan kernels) -> k_wp_fill: the non-fluid particles of the
//   owned range, compacted in sorted order, so k_wall_pressure runs full warps of neighbouring dummy particles (their
//   candidate runs overlap: L1 hits) instead of one warp in ten lanes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wp_flags(int n, const int32_t* __restrict__ tag, int32_t* __restrict__ pos) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n) return;
    pos[s] = s < n && tag[s] != 0;
}

__global__ void __launch_bounds__(256) k_wp_fill(int n, const int32_t* __restrict__ tag, const int32_t* __restrict__ pos, int32_t* __restrict__ idx) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n && tag[s] != 0) idx[pos[s]] = s;
}

template <class R, int DIM, bool MORTON>
__global__ void __launch_bounds__(kThreads) k_wall_pressure(GridDev<R> g, WcsphConst<R> C, int n, const int32_t* __restrict__ pos,
                                                            const int32_t* __restrict__ idx, const int32_t* __restrict__ tag,
                                                            const R* __restrict__ x, const R* __restrict__ y, const R* __restrict__ z,
                                                            const R* __restrict__ h, R* rho, R* p, R* __restrict__ por2,
                                                            const int32_t* __restrict__ cell_start, PosF<R> F, float marg) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= pos[n]) return;                 // pos[n] = number of non-fluid particles (device-side count, no host sync)
    const int s = idx[t];
    const R xi = x[s], yi = y[s], zi = DIM == 3 ? z[s] : (R)0, hi = h[s];
    const R ad = wendland_alpha<R, DIM>(hi);
    const R inv_h = (R)1 / hi;
    const R rc = mul_rn(C.kfac, hi);
    const R rc2 = mul_rn(rc, rc);
    const int cx = cell_coord<R>(xi, g.lo[0], g.inv[0], g.cx_lo, g.cx_hi);
    const int cy = cell_coord<R>(yi, g.lo[1], g.inv[1], 0, g.n[1] - 1);
    const int cz = DIM == 3 ? cell_coord<R>(zi, g.lo[2], g.inv[2], 0, g.n[2] - 1) : 0;
    WallSums<R> S{0, 0, 0, 0, 0};
    // With the f32 position rows of the packed records at hand (F.x != null) the ~370 candidates are pre-filtered in f32 -- the
    // pair kernel's conservative test, same margin -- and only the survivors (~58) reach the f64 loads and the exact test.
    const bool pre = F.x != nullptr;
    const float xf = pre ? pos_f32<R>(xi, F.lo[0], F.cmin, F.cmax[0]) : 0.0f, yf = pre ? pos_f32<R>(yi, F.lo[1], F.cmin, F.cmax[1]) : 0.0f;
    const float zf = pre && DIM == 3 ? pos_f32<R>(zi, F.lo[2], F.cmin, F.cmax[2]) : 0.0f;
    const float rc2f = pre ? __double2float_ru((double)rc2 * (1.0 + (double)marg)) : 0.0f;
    auto exact = [&](int j) {
        const R dx = xi - x[j], dy = yi - y[j], dz = DIM == 3 ? zi - z[j] : (R)0;
        const R r2 = dist2<DIM, R>(dx, dy, dz);
        if (r2 < rc2 && r2 > (R)0 && tag[j] == 0) wall_accumulate<R, DIM>(S, ad, inv_h, dx, dy, dz, r2, p[j], rho[j]);
    };
    for_each_run<DIM, MORTON>(g, cell_start, cx, cy, cz, [&](int b, int e) {
        if (!pre) {
            for (int j = b; j < e; ++j) exact(j);
            return;
        }
        for (int j0 = b; j0 < e; j0 += 8) {       // eight candidates' loads in flight per trip (the loop is latency-bound)
            float d[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int j = min(j0 + t, e - 1);
                const float ex = xf - F.x[j], ey = yf - F.y[j], ez = DIM == 3 ? zf - F.z[j] : 0.0f;
                d[t] = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
            }
#pragma unroll
            for (int t = 0; t < 8; ++t)
                if (j0 + t < e && d[t] <= rc2f) exact(j0 + t);
        }
    });
    R pv, rw;
    wall_finish<R>(C, S, pv, rw);
    p[s] = pv;
    rho[s] = rw;
    por2[s] = pv / (rw * rw);
}

template <class R, int DIM, bool MORTON>
pst_status launch_wall_pressure(pst_ctx* ctx) {
    const int n = (int)ctx->n;
    const int32_t* tag = pst_ptr<int32_t>(ctx, "tag");
    if (!tag) return pst_fail(ctx, PST_ESTATE, "wall_pressure needs the tag array");
    if (!ctx->wp_pos) {
        const size_t cap = ctx->capacity + 2;
        if (cudaMalloc((void**)&ctx->wp_pos, cap * 4) != cudaSuccess || cudaMalloc((void**)&ctx->wp_idx, cap * 4) != cudaSuccess)
            return pst_fail(ctx, PST_ENOMEM, "wall_pressure compaction buffers");
    }
    PST_LAUNCH(ctx, k_wp_flags, blocks_for(n + 1, 256), 256, 0, n, tag, ctx->wp_pos);
    PST_TRY(pst_scan_exclusive(ctx, ctx->wp_pos, n + 1));
    PST_LAUNCH(ctx, k_wp_fill, blocks_for(n, 256), 256, 0, n, tag, ctx->wp_pos, ctx->wp_idx);
    // f32 pre-filter: only when the f32 position rows are current (written with the packed records by the EOS pass that must
    // precede this equation) and the keys are linear; the margin is the pair kernel's (wcsph_zrun.cuh: 8 * 2^-23 * E / rc)
    PosF<R> F = pos_f<R>(ctx);
    PST_TRY(pst_uniform_refresh(ctx));
    if (MORTON || !ctx->rec || ctx->rec_epoch != ctx->state_epoch || ctx->ghost_eos_pending) F.x = F.y = F.z = nullptr;
    float marg = 0.0f;
    {
        const PstGrid& gg = ctx->grid;
        double E = 0;
        for (int a = 0; a < DIM; ++a) E = std::max(E, (double)(gg.n[a] / (a == DIM - 1 ? gg.sub : 1) + 2) * gg.cell);
        const double kf = pst_param(ctx, "kfac", 2.0);
        const double rc = std::max(1e-300, kf * (ctx->h_value > 0 ? ctx->h_value : gg.cell / kf));
        marg = (float)std::max(1.0 / 32768.0, 8.0 * std::ldexp(1.0, -23) * E / rc);
    }
    // the grid covers the worst case (every particle a dummy); threads beyond the device-side count exit at once
    PST_LAUNCH(ctx, (k_wall_pressure<R, DIM, MORTON>), blocks_for(n, kThreads), kThreads, 0, make_grid_dev<R>(ctx->grid), make_const<R>(ctx), n,
               ctx->wp_pos, ctx->wp_idx, tag, pst_ptr<R>(ctx, "x"), pst_ptr<R>(ctx, "y"), pst_ptr<R>(ctx, "z"), pst_ptr<R>(ctx, "h"),
               pst_ptr<R>(ctx, "rho"), pst_ptr<R>(ctx, "p"), pst_ptr<R>(ctx, "por2"), ctx->cell_start, F, marg);
    return PST_OK;
}

}  // namespace

pst_status pst_wcsph_wall_pressure(pst_ctx* ctx) {
    if (ctx->n == 0) return PST_OK;
    if (ctx->comm) return pst_fail(ctx, PST_EINVAL, "wall_pressure is single-GPU for now: ghost dummy particles would need a second density exchange");
    const pst_status st = PST_DISPATCH(ctx, launch_wall_pressure, ctx);   // (reads the f32 position rows while they are still current)
    ctx->state_epoch++;     // rho, p, p/rho^2 of the non-fluid rows change: packed records are stale
    return st;
}

pst_status pst_uniform_refresh(pst_ctx* ctx) {
    if (!ctx->uni_dirty) return PST_OK;
    ctx->m_uniform = ctx->h_uniform = false;
    const int n = (int)ctx->n;
    PstArray* m = pst_find(ctx, "m");
    if (n == 0 || !m) { ctx->uni_dirty = false; return PST_OK; }
    if (!ctx->d_uni) {
        PST_CUDA(ctx, cudaMalloc((void**)&ctx->d_uni, 8 * sizeof(unsigned long long)));
        PST_CUDA(ctx, cudaHostAlloc((void**)&ctx->h_uni, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
    }
    ctx->h_uni[0] = ctx->h_uni[2] = ~0ull; ctx->h_uni[1] = ctx->h_uni[3] = ctx->h_uni[4] = 0ull;
    PST_CUDA(ctx, cudaMemcpyAsync(ctx->d_uni, ctx->h_uni, 5 * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
    const unsigned grid = std::min(blocks_for(n, 256), 1184u);
    if (ctx->f64) PST_LAUNCH(ctx, k_uniform_check<double>, grid, 256, 0, n, pst_ptr<double>(ctx, "m"), pst_ptr<double>(ctx, "h"), pst_ptr<double>(ctx, "rad"), ctx->d_uni);
    else PST_LAUNCH(ctx, k_uniform_check<float>, grid, 256, 0, n, pst_ptr<float>(ctx, "m"), pst_ptr<float>(ctx, "h"), pst_ptr<float>(ctx, "rad"), ctx->d_uni);
    PST_CUDA(ctx, cudaMemcpyAsync(ctx->h_uni, ctx->d_uni, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // once per change of m / h, not per step
    auto as_double = [](unsigned long long k) { double d; std::memcpy(&d, &k, sizeof d); return d; };
    ctx->m_uniform = ctx->h_uni[0] == ctx->h_uni[1];
    ctx->m_value = as_double(ctx->h_uni[0]);
    ctx->h_uniform = pst_find(ctx, "h") && ctx->h_uni[2] == ctx->h_uni[3];
    ctx->h_value = as_double(ctx->h_uni[2]);
    ctx->h_max = pst_find(ctx, "h") ? as_double(ctx->h_uni[3]) : 0.0;
    ctx->rad_max = pst_find(ctx, "rad") ? as_double(ctx->h_uni[4]) : 0.0;
    ctx->uni_dirty = false;
    return PST_OK;
}

// The stencil walkers visit one cell around a particle's own: a cutoff (kfac h, or R_i + R_j <= 2 max R for contacts) larger than
// the cell edge would silently lose neighbours.  The largest h and radius of the owned particles are part of the device-side
// check of uploaded data (pst_uniform_refresh: once per change of m / h / rad, no host scan), so the precondition is an error here.
pst_status pst_check_cell_size(pst_ctx* ctx) {
    PST_TRY(pst_uniform_refresh(ctx));
    const double cell = ctx->grid.cell * (1.0 + 1e-12);
    if ((ctx->cfg.physics & PST_PHYS_WCSPH) && pst_param(ctx, "kfac", 2.0) * ctx->h_max > cell)
        return pst_fail(ctx, PST_EINVAL, "cell_size %.17g is smaller than the largest cutoff kfac * max(h) = %.17g: neighbours would be lost",
                        ctx->grid.cell, pst_param(ctx, "kfac", 2.0) * ctx->h_max);
    if ((ctx->cfg.physics & PST_PHYS_DEM) && 2.0 * ctx->rad_max > cell)
        return pst_fail(ctx, PST_EINVAL, "cell_size %.17g is smaller than the largest contact distance 2 * max(rad) = %.17g: contacts would be lost",
                        ctx->grid.cell, 2.0 * ctx->rad_max);
    return PST_OK;
}

// nnps.cu asks before it builds its permute list: does the WCSPH state travel through the fused permute + EOS kernel?
bool pst_wcsph_fused_permute(pst_ctx* ctx) {
    return (ctx->cfg.physics & PST_PHYS_WCSPH) && !ctx->coupled && rec_wanted(ctx) && pst_option(ctx, "fuse_eos", 1) == 1 &&
           pst_find(ctx, "rho") && pst_find(ctx, "p") && pst_find(ctx, "por2");
}

template <class R>
static pst_status launch_permute_eos(pst_ctx* ctx, const uint32_t* perm, int n) {
    PST_TRY(rec_alloc(ctx));
    PermEosArgs<R> P;
    const char* names[9] = {"x", "y", "z", "u", "v", "w", "rho", "m", "h"};
    for (int a = 0; a < 9; ++a) {
        PstArray* arr = pst_find(ctx, names[a]);
        P.src[a] = arr ? pst_ptr<R>(ctx, arr, 0, arr->cur) : nullptr;
        P.dst[a] = arr ? pst_ptr<R>(ctx, arr, 0, 1 - arr->cur) : nullptr;
    }
    P.p = pst_ptr<R>(ctx, "p"); P.por2 = pst_ptr<R>(ctx, "por2");
    P.rec = rec_ptr<R>(ctx);
    P.F = pos_f<R>(ctx);
    const char* names4[2] = {"tag", "id"};
    for (int a = 0; a < 2; ++a) {
        PstArray* arr = pst_find(ctx, names4[a]);
        const bool ok = arr && arr->esize == 4 && arr->rows == 1 && (arr->flags & PST_ARRAY_PERSISTENT);
        P.src4[a] = ok ? pst_ptr<uint32_t>(ctx, arr, 0, arr->cur) : nullptr;
        P.dst4[a] = ok ? pst_ptr<uint32_t>(ctx, arr, 0, 1 - arr->cur) : nullptr;
    }
    PST_LAUNCH(ctx, k_permute_eos<R>, blocks_for(n, 256), 256, 0, make_const<R>(ctx), n, perm, P);
    return PST_OK;
}
// the caller (build_pass) flips the nine arrays afterwards and marks EOS and records current
pst_status pst_wcsph_permute_eos(pst_ctx* ctx, const uint32_t* perm, int n) {
    return ctx->f64 ? launch_permute_eos<double>(ctx, perm, n) : launch_permute_eos<float>(ctx, perm, n);
}

// A halo exchange wrote the ghost rows.  If the owned rows' EOS and records are current (the fused permute of this step), they stay
// so: only the ghost rows are pending, and tait_eos will run over those alone (two short ranges instead of the whole slab).
void pst_note_ghosts_changed(pst_ctx* ctx) {
    const bool keep = ctx->eos_valid && ctx->rec && ctx->rec_epoch == ctx->state_epoch && pst_wcsph_fused_permute(ctx);
    ctx->state_epoch++;
    if (keep) { ctx->rec_epoch = ctx->state_epoch; ctx->ghost_eos_pending = true; }
    else ctx->eos_valid = false;
}

pst_status pst_wcsph_eos(pst_ctx* ctx) {
    const int n = (int)ctx->n, nl = (int)ctx->n_ghost_l, nr = (int)ctx->n_ghost_r;
    // nothing to do for the owned rows when the re-sort has just evaluated them (fused permute) and nothing changed since
    if (ctx->eos_valid && ctx->rec && ctx->rec_epoch == ctx->state_epoch && pst_wcsph_fused_permute(ctx)) {
        if (ctx->ghost_eos_pending) {
            PST_TRY(ctx->f64 ? launch_eos<double>(ctx, -nl, 0) : launch_eos<float>(ctx, -nl, 0));
            PST_TRY(ctx->f64 ? launch_eos<double>(ctx, n, n + nr) : launch_eos<float>(ctx, n, n + nr));
            ctx->ghost_eos_pending = false;
        }
        return PST_OK;
    }
    PST_TRY(ctx->f64 ? launch_eos<double>(ctx, -nl, n + nr) : launch_eos<float>(ctx, -nl, n + nr));
    ctx->eos_valid = true;
    ctx->ghost_eos_pending = false;
    return PST_OK;
}

pst_status pst_wcsph_forces(pst_ctx* ctx, bool continuity, bool momentum) {
    if (ctx->n == 0) return PST_OK;
    PST_TRY(pst_check_cell_size(ctx));
    PST_TRY(PST_DISPATCH(ctx, launch_forces, ctx, continuity, momentum));
    ctx->pair_kernel_fn = ctx->last_kernel_fn;
    return PST_OK;
}

pst_status pst_coupled_integrate(pst_ctx* ctx, double dt) {
    if (ctx->n == 0) return PST_OK;
    return ctx->f64 ? launch_coupled_integrate<double>(ctx, dt) : launch_coupled_integrate<float>(ctx, dt);
}

pst_status pst_wcsph_integrate(pst_ctx* ctx, double dt) {
    if (ctx->n == 0) return PST_OK;
    return PST_DISPATCH(ctx, launch_integrate, ctx, dt);
}
