// wcsph_zrun.cuh -- variant 3 of the fused continuity + momentum pair kernel (option force_kernel = 3).  Included by
// wcsph.cu inside its anonymous namespace (it shares ForceArgs, gather_one, TileDims with the other variants).
// No reference code exists for the physics; the loop shape it honours is the reference's
// gather -- write only [i], bodies of a fused set in ONE i,j loop (prestige/src/codegen/simple_cpu.rs:7-16, fuse.rs:14-40).
//
// What changed against variant 2 (k_wcsph_tiled), and why (profiles/r2_ncu_k_wcsph_tiled_10m_base.txt: 43 % of the warp
// samples sat in the candidate scan, 22 % in the tile preamble, 33 % in the pair bodies; L1 data pipe 80 % busy):
//   * the cell grid is `sub` times finer along the FAST axis (option zsub; nnps.cu sorts by the fine key), so inside a
//     stencil column the particles are ordered by fine z.  A particle scans, per column, only the fine cells within
//     +-sqrt(rc^2 - d_xy^2) of its own z, d_xy = its distance to that column's footprint: ~190 candidates instead of the
//     373 of the 27-cell stencil;
//   * the scan produces BIT MASKS, not lists: d = dx^2 + dy^2 + dz^2 - rc^2 comes out of three packed FFMA2, its sign bit
//     is funnel-shifted into a 32-bit word (one SHF per candidate, no compare, no predicated store, no serial list
//     pointer), one word per 32 scanned candidates, empty words dropped;
//   * all lanes of a warp scan a run in lock-step (trip count = the warp's longest range, the surplus bits are cut off),
//     every LDS.128 is 16-byte aligned by construction (runs are staged at multiples of 4, ranges start aligned down);
//   * staging holds 12 B per candidate (f32 tile-local x, y, z; the global index follows from the word's base);
//   * the tile preamble is parallel (warp-shuffle scans, one warp per staged run, geometric tile origin: no global load on
//     the critical path).
// The pair bodies (phase 2) are unchanged: the exact FMA-free test decides membership, so the neighbour set stays bit-exact.

struct ZTile {
    int GF;          // fine cells per tile along the fast axis
    int tiles[3];    // tile grid: [0] columns-x, [1] columns-y (1 in 2D), [2] fast axis
    int jcap;        // staged-candidate capacity (<= kZJcap; tests shrink it to force the fallback)
    int maxw;        // mask words per thread
};

constexpr int kZPad = 64;      // readable slack behind the staged rows: lanes with a short range over-scan with the warp

template <class R, int DIM, int TA, int TB, int NT, bool CONT, bool MOM, bool COUPLED = false, bool UNI = false, bool REC = false, int MINB = 2, int JC = 2304, int DBG = 0>
__global__ void __launch_bounds__(NT, MINB) k_wcsph_zrun(GridDev<R> g, WcsphConst<R> C, ForceArgs<R> A, ZTile T) {
    constexpr int kZJcap = JC, kZRow = JC + kZPad;   // staged candidates per tile (12 B each)
    using D = TileDims<DIM, TA, TB>;
    constexpr int NR = D::NR, NI = D::NI, RY = D::RY, BB = D::BB;
    constexpr int NW = NT / 32;
    constexpr int NRUN = DIM == 3 ? 9 : 3;
    constexpr int FAST = DIM - 1;
    static_assert(NR <= 32 && NI <= 32, "one warp scans the run and column tables");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = g.sub, GF = T.GF, W = GF + 2 * S + 1;   // W fine-cell boundaries per staged run
    // ---- shared memory carve-up
    int* s_cs = reinterpret_cast<int*>(smem_raw);             // NR * W   staged offset of every fine-cell boundary
    int* s_gbeg = s_cs + NR * W;                               // NR       global begin of each run
    int* s_voff = s_gbeg + NR;                                 // NR + 1   staged offset of each run (multiples of 4)
    int* s_ibeg = s_voff + NR + 1;                             // NI       global begin of each i segment
    int* s_ipre = s_ibeg + NI;                                 // NI + 1   prefix of i counts
    size_t off = ((size_t)(NR * W + NR + NR + 1 + NI + NI + 1) * sizeof(int) + 15) & ~(size_t)15;
    float* s_x = reinterpret_cast<float*>(smem_raw + off);
    float* s_y = s_x + kZRow;
    float* s_z = s_y + kZRow;
    unsigned* s_mask = reinterpret_cast<unsigned*>(s_z + (DIM == 3 ? kZRow : 0));   // maxw * NT
    int* s_base = reinterpret_cast<int*>(s_mask + T.maxw * NT);                      // maxw * NT: global index of a word's first candidate

    using P2 = typename RecPair<R>::type;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- which tile
    int b = blockIdx.x;
    const int tf = b % T.tiles[2]; b /= T.tiles[2];
    const int ty = DIM == 3 ? b % T.tiles[1] : 0; if (DIM == 3) b /= T.tiles[1];
    const int tx = b;
    const int cx0 = tx * TA, cy0 = ty * BB, f0 = tf * GF;
    const int nf = g.n[FAST];                       // fine cells along the fast axis
    const int ncx = g.n[0], ncy = DIM == 3 ? g.n[1] : 1;

    // ---- fine-cell boundaries of every staged run (global indices first)
    for (int t = tid; t < NR * W; t += NT) {
        const int q = t / W, tt = t - q * W;
        const int rx = q / RY, ry = q - rx * RY;
        const int cx = cx0 - 1 + rx, cy = DIM == 3 ? cy0 - 1 + ry : 0;
        int gi = 0;   // column outside the grid: all boundaries equal -> empty run
        if (cx >= 0 && cx < ncx && cy >= 0 && cy < ncy) {
            const int col = DIM == 3 ? cx * ncy + cy : cx;
            gi = A.cell_start[(size_t)col * nf + min(max(f0 - S + tt, 0), nf)];
        }
        s_cs[t] = gi;
    }
    __syncthreads();
    if (warp == 0) {
        // staged offsets: exclusive scan of the run lengths rounded up to 4 (so every run starts 16-byte aligned)
        int len = 0, gb = 0;
        if (lane < NR) { gb = s_cs[lane * W]; len = s_cs[lane * W + W - 1] - gb; }
        int incl = (len + 3) & ~3;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        if (lane < NR) { s_gbeg[lane] = gb; s_voff[lane] = incl - ((len + 3) & ~3); }
        if (lane == NR - 1) s_voff[NR] = incl;
        // i segments: tile columns, fine cells [f0, f0 + GF) == boundaries tt = S .. S + GF of the centre runs
        int cnt = 0, beg = 0;
        if (lane < NI) {
            const int lx = lane / BB, ly = lane - lx * BB;
            const int q = (lx + 1) * RY + (DIM == 3 ? ly + 1 : 0);
            const int cx = cx0 + lx, cy = cy0 + ly;
            if (cx >= g.cx_lo && cx <= g.cx_hi && cy < ncy) { beg = s_cs[q * W + S]; cnt = s_cs[q * W + S + GF] - beg; }   // ghost layers are never i
        }
        int ipre = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, ipre, d);
            if (lane >= d) ipre += y;
        }
        if (lane < NI) { s_ibeg[lane] = beg; s_ipre[lane] = ipre - cnt; }
        if (lane == NI - 1) s_ipre[NI] = ipre;
    }
    __syncthreads();
    const int ni = s_ipre[NI];
    if (ni == 0) return;                      // empty tile (uniform exit)
    const int M = s_voff[NR];
    // tile-local coordinates: origin = low corner of the tile's own cells (no global load needed)
    const R cellR = g.cell;
    const R ox = g.lo[0] + (R)cx0 * cellR;
    const R oy = DIM == 3 ? g.lo[1] + (R)cy0 * cellR : g.lo[1] + (R)f0 * (cellR / (R)S);
    const R oz = DIM == 3 ? g.lo[2] + (R)f0 * (cellR / (R)S) : (R)0;
    bool far = false;
    const bool dense = M > min(kZJcap, T.jcap);
    if (!dense) {
        // ---- rebase boundaries to staged offsets (every entry by the thread that owns it; the run tables are read-only by now)
        for (int t = tid; t < NR * W; t += NT) {
            const int q = t / W;
            s_cs[t] = s_voff[q] + (s_cs[t] - s_gbeg[q]);
        }
    }
    __syncthreads();
    if (!dense) {
        // ---- stage candidates: one warp per run, coalesced; the slack up to the next multiple of 4 holds a far-away dummy
        const float far_lim = (float)(GF / S + 6) * (float)g.cell;
        for (int q = warp; q < NR; q += NW) {
            const int gb = s_gbeg[q], vo = s_voff[q], plen = s_voff[q + 1] - vo;
            const int len = s_cs[q * W + W - 1] - vo;      // true length (rebased boundaries)
            for (int v = lane; v < plen; v += 32) {
                float px = 1e30f, py = 0.0f, pz = 0.0f;    // dummy: never within any cutoff
                if (v < len) {
                    const int gj = gb + v;
                    px = (float)(A.x[gj] - ox); py = (float)(A.y[gj] - oy); pz = DIM == 3 ? (float)(A.z[gj] - oz) : 0.0f;
                    far |= !(fabsf(px) <= far_lim && fabsf(py) <= far_lim && fabsf(pz) <= far_lim);
                }
                s_x[vo + v] = px; s_y[vo + v] = py; if (DIM == 3) s_z[vo + v] = pz;
            }
        }
        for (int v = tid; v < kZPad; v += NT) { s_x[M + v] = 1e30f; s_y[M + v] = 0.0f; if (DIM == 3) s_z[M + v] = 0.0f; }
    }
    far = __syncthreads_or(far);
    if (dense || far) {
        // tile denser than the staging buffer, or holding particles far outside the box (clamped into its cells): the exact
        // per-particle gather for its particles
        for (int ii = tid; ii < ni; ii += NT) {
            int c = 0;
            while (c + 1 < NI && ii >= s_ipre[c + 1]) ++c;
            gather_one<R, DIM, false, CONT, MOM, COUPLED>(g, C, A, s_ibeg[c] + (ii - s_ipre[c]));
        }
        return;
    }

    const int MAXW = T.maxw;
    const float cellf = (float)g.cell;
    const float inv_cf = (float)S / cellf;          // 1 / fine cell edge
    unsigned* const my_mask = s_mask + tid;
    int* const my_base = s_base + tid;
    for (int ii0 = warp * 32; ii0 < ni; ii0 += NT) {   // a warp takes 32 consecutive particles per round (warp-uniform trip count)
        const int ii = ii0 + lane;
        const bool active = ii < ni;
        int c = 0, gi = 0, lx = 0, ly = 0;
        IState<R, DIM> I;
        Acc<R> a{0, 0, 0, 0}, a2{0, 0, 0, 0};
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
            xf = (float)(I.x - ox); yf = (float)(I.y - oy); zf = DIM == 3 ? (float)(I.z - oz) : 0.0f;
            // conservative f32 pre-filter: relative margin 2^-15 on rc^2 (>> the f32 error of tile-local coordinates, see
            // launch_zrun); the range cull below uses twice that
            rc2f = __double2float_ru((double)(UNI ? C.u_rc2 : I.rc2) * (1.0 + 1.0 / 32768.0));
            rc2m = rc2f * (1.0f + 1.0f / 16384.0f);
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
            if (ax == 0) dxc = fmaxf(xf - (float)lx * cellf, 0.0f);
            if (ax == 2) dxc = fmaxf((float)(lx + 1) * cellf - xf, 0.0f);
            if (DIM == 3) {
                if (ay == 0) dyc = fmaxf(yf - (float)ly * cellf, 0.0f);
                if (ay == 2) dyc = fmaxf((float)(ly + 1) * cellf - yf, 0.0f);
            }
            const float rem = rc2m - (dxc * dxc + dyc * dyc);
            vs = ve = 0;
            if (active && rem > 0.0f) {
                const float zext = sqrtf(rem) * 1.00001f;
                // fine cells [tlo, thi] relative to boundary 0 of the staged run (= fine cell f0 - S); +-1e-3 cell of slack,
                // then clamped like the keys themselves (particles outside the box sit in the edge cells)
                int flo = (int)floorf((fz - zext) * inv_cf - 1e-3f) + f0, fhi = (int)floorf((fz + zext) * inv_cf + 1e-3f) + f0;
                flo = min(max(flo, 0), nf - 1); fhi = min(max(fhi, 0), nf - 1);
                const int tlo = min(max(flo - (f0 - S), 0), W - 2), thi = min(max(fhi - (f0 - S), 0), W - 2);
                vs = s_cs[q_run * W + tlo];
                ve = s_cs[q_run * W + thi + 1];
            }
        };

        int k = 0, wc = 0;          // warp-uniform scan cursor: stencil run, 32-candidate chunk inside it
        while (true) {
            // ---- phase 1: pre-filter on staged f32 coordinates -> bit masks (bit 31 = first candidate of the word)
            int nw = 0, nch = 0;
            while (k < NRUN) {
                open_run(k);
                int a4 = vs & ~3;
                const int ng = (ve - a4 + 3) >> 2;
                const int Tg = __reduce_max_sync(0xffffffffu, ve > vs ? ng : 0);   // the warp scans its longest range
                bool full = false;
                while (wc * 8 < Tg) {
                    if (nch == MAXW) { full = true; break; }
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
                        my_mask[nw * NT] = m;
                        my_base[nw * NT] = j0 + (s_gbeg[q_run] - s_voff[q_run]);
                        ++nw;
                    }
                    ++nch;
                    ++wc;
                }
                if (full) break;
                ++k; wc = 0;
            }
            // ---- phase 2: two hits per trip, j state gathered from global memory (L1/L2 hits), exact test, branch-free body
            {
                // hit iterator over the mask words (none of them empty): m = bits left in the current word, base31 = global index
                // of its first candidate + 31, w = next word
                int w = 0, base31 = 0;
                unsigned m = 0;
                if (nw > 0 && DBG != 3) { m = my_mask[0]; base31 = my_base[0] + 31; w = 1; }
                if (DBG == 3) a.au += (R)nw;      // timing ablation: phase 1 only
                auto pop = [&]() -> int {      // requires m != 0
                    const int msb = 31 - __clz(m);
                    int j = base31 - msb;
                    asm("" : "+r"(j));        // keep j a plain 32-bit value: the record index below is 32-bit arithmetic + one IMAD.WIDE
                    m ^= 1u << msb;
                    if (m == 0 && w < nw) { m = my_mask[w * NT]; base31 = my_base[w * NT] + 31; ++w; }
                    return j;
                };
                while (m != 0) {
                    const int j0 = pop();
                    const bool v1 = m != 0;
                    const int j1 = v1 ? pop() : j0;
                    R xj0, yj0, zj0, uj0, vj0, wj0, rj0, pj0, xj1, yj1, zj1, uj1, vj1, wj1, rj1, pj1, mj0 = A.m_uni, mj1 = A.m_uni;
                    const R rc2 = UNI ? C.u_rc2 : I.rc2;
                    if (DBG == 1) {     // timing ablation (wrong results): no gathers, the j state is made up from the index
                        xj0 = I.x + (R)(j0 & 15) * (R)1e-3; yj0 = I.y + (R)(j0 & 7) * (R)1e-3; zj0 = I.z; uj0 = I.u; vj0 = I.v; wj0 = (R)j0; rj0 = I.rho; pj0 = I.por2;
                        xj1 = I.x + (R)(j1 & 15) * (R)1e-3; yj1 = I.y + (R)(j1 & 7) * (R)1e-3; zj1 = I.z; uj1 = I.u; vj1 = I.v; wj1 = (R)j1; rj1 = I.rho; pj1 = I.por2;
                    } else if (REC) {
                        const P2* q0 = reinterpret_cast<const P2*>(A.rec) + rec_index(j0);
                        const P2* q1 = reinterpret_cast<const P2*>(A.rec) + rec_index(j1);
                        const P2 a0 = q0[0], b0 = q0[8], a1 = q1[0], b1 = q1[8];
                        P2 c0, d0, c1, d1;
                        if (DBG == 4) { c0.x = I.v; c0.y = I.w; d0.x = I.rho; d0.y = I.por2; c1 = c0; d1 = d0; }   // timing ablation: half the gather
                        else { c0 = q0[16]; d0 = q0[24]; c1 = q1[16]; d1 = q1[24]; }
                        xj0 = a0.x; yj0 = a0.y; zj0 = b0.x; uj0 = b0.y; vj0 = c0.x; wj0 = c0.y; rj0 = d0.x; pj0 = d0.y;
                        xj1 = a1.x; yj1 = a1.y; zj1 = b1.x; uj1 = b1.y; vj1 = c1.x; wj1 = c1.y; rj1 = d1.x; pj1 = d1.y;
                        if (!UNI) { mj0 = q0[32].x; mj1 = q1[32].x; }
                    } else {
                        xj0 = A.x[j0]; yj0 = A.y[j0]; zj0 = DIM == 3 ? A.z[j0] : (R)0; uj0 = A.u[j0]; vj0 = A.v[j0]; wj0 = DIM == 3 ? A.w[j0] : (R)0;
                        rj0 = A.rho[j0]; pj0 = A.por2[j0];
                        xj1 = A.x[j1]; yj1 = A.y[j1]; zj1 = DIM == 3 ? A.z[j1] : (R)0; uj1 = A.u[j1]; vj1 = A.v[j1]; wj1 = DIM == 3 ? A.w[j1] : (R)0;
                        rj1 = A.rho[j1]; pj1 = A.por2[j1];
                        if (!UNI) { mj0 = A.m[j0]; mj1 = A.m[j1]; }
                    }
                    if (DBG == 2) {     // timing ablation (wrong results): gathers only, no pair body
                        a.au += xj0 + yj0 + zj0 + uj0; a.av += vj0 + wj0 + rj0 + pj0; a2.au += xj1 + yj1 + zj1 + uj1; a2.av += vj1 + wj1 + rj1 + pj1;
                        continue;
                    }
                    const R dx0 = I.x - xj0, dy0 = I.y - yj0, dz0 = DIM == 3 ? I.z - zj0 : (R)0;
                    const R dx1 = I.x - xj1, dy1 = I.y - yj1, dz1 = DIM == 3 ? I.z - zj1 : (R)0;
                    R r20 = dist2<DIM, R>(dx0, dy0, dz0), r21 = dist2<DIM, R>(dx1, dy1, dz1);
                    const bool in0 = r20 < rc2 && r20 > (R)0;              // the exact test (the set is defined here)
                    const bool in1 = v1 && r21 < rc2 && r21 > (R)0;
                    r20 = in0 ? r20 : (R)1; r21 = in1 ? r21 : (R)1;
                    R m0 = in0 ? mj0 : (R)0, m1 = in1 ? mj1 : (R)0;
                    if (COUPLED) {   // signed SPH mass: the pair counts iff i or j is fluid
                        m0 = (fluid_i || m0 > (R)0) ? fabs(m0) : (R)0;
                        m1 = (fluid_i || m1 > (R)0) ? fabs(m1) : (R)0;
                    }
                    pair_body<R, DIM, CONT, MOM, UNI>(C, I, dx0, dy0, dz0, r20, uj0, vj0, wj0, rj0, m0, pj0, a);
                    pair_body<R, DIM, CONT, MOM, UNI>(C, I, dx1, dy1, dz1, r21, uj1, vj1, wj1, rj1, m1, pj1, a2);
                }
            }
            if (k == NRUN) break;      // warp-uniform
        }
        if (active) {
            a.au += a2.au; a.av += a2.av; a.aw += a2.aw; a.arho += a2.arho;
            store_acc<R, DIM, CONT, MOM>(A, C, gi, a);
        }
    }
}

template <class R, int DIM, int TA, int TB, int NT, int MINB, int JC, bool CONT, bool MOM, bool COUPLED = false, bool UNI = false>
pst_status launch_zrun_k(pst_ctx* ctx, const ZTile& T, size_t smem) {
    const bool rec = rec_wanted(ctx);
    if (rec) PST_TRY(rec_refresh<R>(ctx));
    auto kern = rec ? k_wcsph_zrun<R, DIM, TA, TB, NT, CONT, MOM, COUPLED, UNI, true, MINB, JC> : k_wcsph_zrun<R, DIM, TA, TB, NT, CONT, MOM, COUPLED, UNI, false, MINB, JC>;
    if (UNI && rec && CONT && MOM && !COUPLED && DIM == 3 && MINB == 2) {     // timing ablations of the headline kernel (wrong results)
        const int dbg = pst_option(ctx, "tile_dbg", 0);
        if (dbg == 1) kern = k_wcsph_zrun<R, DIM, TA, TB, NT, CONT, MOM, COUPLED, UNI, true, MINB, JC, 1>;
        if (dbg == 2) kern = k_wcsph_zrun<R, DIM, TA, TB, NT, CONT, MOM, COUPLED, UNI, true, MINB, JC, 2>;
        if (dbg == 3) kern = k_wcsph_zrun<R, DIM, TA, TB, NT, CONT, MOM, COUPLED, UNI, true, MINB, JC, 3>;
        if (dbg == 4) kern = k_wcsph_zrun<R, DIM, TA, TB, NT, CONT, MOM, COUPLED, UNI, true, MINB, JC, 4>;
    }
    PST_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const unsigned grid = (unsigned)T.tiles[0] * T.tiles[1] * T.tiles[2];
    WcsphConst<R> C = make_const<R>(ctx);
    if (UNI) fill_uniform<R, DIM>(ctx, C);
    PST_LAUNCH(ctx, kern, grid, NT, smem, make_grid_dev<R>(ctx->grid), C, make_args<R>(ctx), T);
    return PST_OK;
}

template <class R, int DIM, int TA, int TB, int NT, int MINB, int JC>
pst_status launch_zrun_shape(pst_ctx* ctx, bool cont, bool mom) {
    constexpr int kZJcap = JC, kZRow = JC + kZPad;
    using D = TileDims<DIM, TA, TB>;
    const PstGrid& g = ctx->grid;
    const int S = g.sub;
    const int nf = g.n[DIM - 1];
    double ppc = pst_param(ctx, "_ppc", 0.0);      // mean occupancy of the occupied COARSE cells (k_scan_tiles)
    if (!(ppc > 0)) ppc = DIM == 3 ? 14.0 : 6.0;
    ZTile T;
    T.maxw = pst_option(ctx, "tile_words", 16);
    const int user_G = pst_option(ctx, "tile_g", 0) * S + pst_option(ctx, "tile_gf", 0);   // tile_g in COARSE cells, tile_gf in fine cells
    // tile depth in fine cells: about one thread per particle
    int GF = user_G > 0 ? user_G : (int)std::floor(0.95 * NT * S / (D::NI * ppc));
    GF = std::min(std::max(GF, 1), std::max(1, nf));
    // f32 pre-filter: tile-local coordinates reach (GF / S + 2 + TA) cells (see test_prefilter_margin.py: <= 32 cells keeps the
    // f32 error of r^2 below half of the 2^-15 margin)
    GF = std::min(GF, kMaxTileG * S);
    // the boundary tables and the staged runs must fit
    while (GF > 1 && (D::NR * (GF + 2 * S + 1) > 3072 || (user_G <= 0 && 1.10 * D::NR * (GF + 2 * S) * ppc / S + 4 * D::NR > kZJcap))) --GF;
    T.GF = GF;
    T.tiles[0] = (g.n[0] + TA - 1) / TA;
    T.tiles[1] = DIM == 3 ? (g.n[1] + D::BB - 1) / D::BB : 1;
    T.tiles[2] = (nf + GF - 1) / GF;
    T.jcap = std::min(kZJcap, std::max(0, pst_option(ctx, "tile_jcap", kZJcap)));
    const size_t ints = ((size_t)(D::NR * (GF + 2 * S + 1) + D::NR + D::NR + 1 + D::NI + D::NI + 1) * sizeof(int) + 15) & ~(size_t)15;
    const size_t smem = ints + (size_t)(DIM == 3 ? 3 : 2) * kZRow * sizeof(float) + (size_t)T.maxw * NT * 8;
    if (smem > 227 * 1024) return pst_fail(ctx, PST_EINVAL, "tile_words too large");
    if (ctx->coupled) {
        if (DIM == 3) return launch_zrun_k<R, 3, TA, TB, NT, MINB, JC, true, true, true>(ctx, T, smem);
        return pst_fail(ctx, PST_EINVAL, "coupled contexts need dim = 3");
    }
    // every particle this rank can see has the same mass AND smoothing length (owned ones: checked on the device; ghosts and
    // migrants of other ranks: the caller vouches for them with "uniform_mass_global"): both become kernel constants
    PST_TRY(pst_uniform_refresh(ctx));
    const bool uni = ctx->m_uniform && ctx->h_uniform && (!ctx->comm || pst_option(ctx, "uniform_mass_global", 0) != 0) && pst_option(ctx, "uniform_mass", 1) != 0;
    if (cont && mom && uni) return launch_zrun_k<R, DIM, TA, TB, NT, MINB, JC, true, true, false, true>(ctx, T, smem);
    if (cont && mom) return launch_zrun_k<R, DIM, TA, TB, NT, MINB, JC, true, true>(ctx, T, smem);
    if (cont) return launch_zrun_k<R, DIM, TA, TB, NT, MINB, JC, true, false>(ctx, T, smem);
    return launch_zrun_k<R, DIM, TA, TB, NT, MINB, JC, false, true>(ctx, T, smem);
}

template <class R, int DIM>
pst_status launch_zrun(pst_ctx* ctx, bool cont, bool mom) {
    // 2 x 2 columns, 256 threads, two CTAs per SM.  Measured and dropped (profiles/r2_exp_log.txt): 4 x 2 columns with 512 threads
    // and one CTA per SM (+6 %), three CTAs per SM at 80 registers (+12..21 %: spills, and the third CTA's shared memory leaves too
    // little L1 for the gathers), prefetch.global.L1 of the next trip's records (+25 %)
    if (pst_option(ctx, "tile_jc", 0) == 1) return launch_zrun_shape<R, DIM, 2, 2, 256, 2, 1664>(ctx, cont, mom);
    return launch_zrun_shape<R, DIM, 2, 2, 256, 2, 2304>(ctx, cont, mom);
}
