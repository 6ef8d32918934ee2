// dem.cu -- discrete-element contact forces with per-pair tangential history.
// Linear spring-dashpot or Hertz-Mindlin normal/tangential law, Coulomb cap, torque; history rows
// keyed by the partner's STABLE id so they survive every re-sort.
// No reference code exists for this physics; the loop shape it honours is the reference's gather
// (prestige/src/codegen/simple_cpu.rs:7-16: write only [i]); each side of a contact evaluates its
// own copy, and the operand order below makes F_ji = -F_ij and xi_ji = -xi_ij bit-exact.
//
// One thread per particle.  Cells hold ~1 sphere, so the 27-cell stencil is 9 short contiguous runs
// (28 candidates, ~6 contacts): the kernel is a latency/bandwidth-bound gather, not FP64-bound.
// History is slot-major ([k][particle]) so slot k of neighbouring threads coalesces; the pass reads
// the current buffer (already permuted by k_remap_history) and writes the other one, then flips.
#include "pst_internal.h"

namespace {

constexpr int kThreads = 128;
inline unsigned blocks_for(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

template <class R>
struct DemConst {
    int model, K;
    R kn, gn, kt, gt, mu, dt, Estar, Gstar, damp_c;  // damp_c = -2 sqrt(5/6) beta_e  (> 0)
};

template <class R>
struct DemArgs {
    const R *x, *y, *z, *u, *v, *w, *wx, *wy, *wz, *rad, *m;
    const uint32_t* id;
    const int32_t* tag;      // coupled SPH-DEM contexts only (else nullptr): i must be a solid (tag 2), j must not be fluid (tag 0)
    const int32_t* body;     // rigid bodies (rigid.cu; else nullptr): members of the same body (index >= 0) are not contact partners
    const int32_t* hn_in; const uint32_t* hid_in; const R *hx_in, *hy_in, *hz_in;
    const uint32_t* hperm;   // deferred history remap: old row of particle s is hperm[s] (nullptr: rows already in place)
    int32_t* hn_out; uint32_t* hid_out; R *hx_out, *hy_out, *hz_out;
    R *fx, *fy, *fz, *tx, *ty, *tz;
    const int32_t* cell_start;
    int32_t* flags;
    size_t stride;
    int n;
};

// one contact: normal + tangential law, history lookup by stable partner id, torque.  Shared by both kernels.
template <class R>
__device__ __forceinline__ void dem_contact(const DemConst<R>& C, const DemArgs<R>& A, int s, int j, R xi, R yi, R zi, R ui, R vi, R wi, R ri,
                                            R mi, R owx, R owy, R owz, int nold, int hrow, R& fx, R& fy, R& fz, R& tx, R& ty, R& tz, int& cnt) {
    const R dx = xi - A.x[j], dy = yi - A.y[j], dz = zi - A.z[j];
    const R r2 = dist2<3, R>(dx, dy, dz);
    const R rj = A.rad[j];
    const R rs = add_rn(ri, rj);
    const R r = sqrt(r2);
    const R rinv = (R)1 / r;
    const R nx = dx * rinv, ny = dy * rinv, nz = dz * rinv;
    const R delta = rs - r;
    // R_i w_i + R_j w_j as a commutative sum of two rounded products (no FMA), so both sides of the
    // contact see the same bits and F_ji = -F_ij exactly
    const R ox = add_rn(owx, mul_rn(rj, A.wx[j])), oy = add_rn(owy, mul_rn(rj, A.wy[j])), oz = add_rn(owz, mul_rn(rj, A.wz[j]));
    const R vcx = (ui - A.u[j]) - (oy * nz - oz * ny);
    const R vcy = (vi - A.v[j]) - (oz * nx - ox * nz);
    const R vcz = (wi - A.w[j]) - (ox * ny - oy * nx);
    const R vn = vcx * nx + vcy * ny + vcz * nz;
    const R vtx = vcx - vn * nx, vty = vcy - vn * ny, vtz = vcz - vn * nz;
    R kn = C.kn, gn = C.gn, kt = C.kt, gt = C.gt;
    if (C.model == 1) {
        const R mj = A.m[j];
        const R Rs = ri * rj / rs;
        const R ms = mi * mj / (mi + mj);
        const R sq = sqrt(Rs * delta);
        const R Sn = (R)2 * C.Estar * sq, St = (R)8 * C.Gstar * sq;
        kn = (R)(4.0 / 3.0) * C.Estar * sq;
        kt = St;
        gn = C.damp_c * sqrt(Sn * ms);
        gt = C.damp_c * sqrt(St * ms);
    }
    const R fnm = kn * delta - gn * vn;
    // history lookup by stable partner id (new contact => xi = 0); branch-free scan so the slot loads overlap
    const uint32_t pid = A.id[j];
    int slot = -1;
    for (int k = 0; k < nold; ++k)
        if (A.hid_in[(size_t)k * A.stride + hrow] == pid) slot = k;
    R hx = 0, hy = 0, hz = 0;
    if (slot >= 0) {
        const size_t o = (size_t)slot * A.stride + hrow;
        hx = A.hx_in[o]; hy = A.hy_in[o]; hz = A.hz_in[o];
    }
    const R xn = hx * nx + hy * ny + hz * nz;
    hx = hx - xn * nx + vtx * C.dt;
    hy = hy - xn * ny + vty * C.dt;
    hz = hz - xn * nz + vtz * C.dt;
    R ftx = -kt * hx - gt * vtx, fty = -kt * hy - gt * vty, ftz = -kt * hz - gt * vtz;
    const R ftm = sqrt(ftx * ftx + fty * fty + ftz * ftz);
    const R fmax = C.mu * fabs(fnm);
    if (ftm > fmax) {
        const R sc = fmax / ftm;
        ftx *= sc; fty *= sc; ftz *= sc;
        const R ikt = (R)1 / kt;
        hx = -(ftx + gt * vtx) * ikt; hy = -(fty + gt * vty) * ikt; hz = -(ftz + gt * vtz) * ikt;
    }
    fx += fnm * nx + ftx; fy += fnm * ny + fty; fz += fnm * nz + ftz;
    const R lx = -ri * nx, ly = -ri * ny, lz = -ri * nz;
    tx += ly * ftz - lz * fty;
    ty += lz * ftx - lx * ftz;
    tz += lx * fty - ly * ftx;
    if (cnt < C.K) {
        const size_t o = (size_t)cnt * A.stride + s;
        A.hid_out[o] = pid; A.hx_out[o] = hx; A.hy_out[o] = hy; A.hz_out[o] = hz;
    }
    ++cnt;
}

template <class R>
__device__ __forceinline__ void dem_finish(const DemConst<R>& C, const DemArgs<R>& A, int s, int cnt, R fx, R fy, R fz, R tx, R ty, R tz) {
    if (cnt > C.K) {            // never silent truncation: PST_EOVERFLOW at the next sync
        atomicExch(&A.flags[0], 1);
        atomicMax(&A.flags[1], cnt);
        cnt = C.K;
    }
    A.hn_out[s] = cnt;
    A.fx[s] = fx; A.fy[s] = fy; A.fz[s] = fz;
    A.tx[s] = tx; A.ty[s] = ty; A.tz[s] = tz;
}

// generic kernel (any key mode): contact body evaluated inside the candidate loop
template <class R, bool MORTON>
__global__ void __launch_bounds__(kThreads) k_dem_forces_generic(GridDev<R> g, DemConst<R> C, DemArgs<R> A) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.n) return;
    if (A.tag && A.tag[s] != 2) { dem_finish<R>(C, A, s, 0, (R)0, (R)0, (R)0, (R)0, (R)0, (R)0); return; }
    const R xi = A.x[s], yi = A.y[s], zi = A.z[s];
    const R ui = A.u[s], vi = A.v[s], wi = A.w[s];
    const R ri = A.rad[s], mi = A.m[s];
    const R owx = mul_rn(ri, A.wx[s]), owy = mul_rn(ri, A.wy[s]), owz = mul_rn(ri, A.wz[s]);   // R_i w_i
    const int hrow = A.hperm ? (int)A.hperm[s] : s;
    const int nold = A.hn_in[hrow];
    const int cx = cell_coord<R>(xi, g.lo[0], g.inv[0], g.cx_lo, g.cx_hi);
    const int cy = cell_coord<R>(yi, g.lo[1], g.inv[1], 0, g.n[1] - 1);
    const int cz = cell_coord<R>(zi, g.lo[2], g.inv[2], 0, g.n[2] - 1);
    R fx = 0, fy = 0, fz = 0, tx = 0, ty = 0, tz = 0;
    int cnt = 0;
    for_each_run<3, MORTON>(g, A.cell_start, cx, cy, cz, [&](int b, int e) {
        for (int j = b; j < e; ++j) {
            const R dx = xi - A.x[j], dy = yi - A.y[j], dz = zi - A.z[j];
            const R r2 = dist2<3, R>(dx, dy, dz);
            const R rs = add_rn(ri, A.rad[j]);
            if (!(r2 < mul_rn(rs, rs)) || !(r2 > (R)0) || j == s) continue;
            if (A.tag && A.tag[j] == 0) continue;
            if (A.body && A.body[s] >= 0 && A.body[j] == A.body[s]) continue;
            dem_contact<R>(C, A, s, j, xi, yi, zi, ui, vi, wi, ri, mi, owx, owy, owz, nold, hrow, fx, fy, fz, tx, ty, tz, cnt);
        }
    });
    dem_finish<R>(C, A, s, cnt, fx, fy, fz, tx, ty, tz);
}

// Coupled SPH-DEM contexts: fluid particles share the (SPH-sized) cells, so a sphere's 27-cell stencil holds ~370
// candidates of which only the non-fluid ones can be contact partners, and only ~1 thread in 10 is a sphere.  Each
// contact pass therefore first compacts the NON-FLUID particles of the sorted order (owned + ghosts):
//   k_nf_clear  zero the force rows the PREVIOUS pass wrote (its compacted list is still there): the output arrays are
//               not permuted by a re-sort, so this keeps "rows of non-spheres are zero" without touching all N rows
//   k_nf_flags  pos[e] = (tag != 0), e = index - lo
//   scan        exclusive, in place (nnps.cu): pos[e] = number of non-fluid particles before e, pos[hi - lo] = their total
//   k_nf_fill   idx[pos[e]] = index, rec[pos[e]] = (x, y, z, rad): the contact TEST then streams 32-byte records that
//               are contiguous along a compacted run instead of four scattered 8-byte gathers per candidate
// Because the compaction keeps the sorted order, a cell range [b, e) of the cell table maps to the contiguous range
// [pos[b], pos[e]) of idx: the contact kernel runs one thread per compacted entry (full warps of spheres) and scans
// compacted runs (no fluid candidates at all).
struct NfArgs {
    const int32_t* pos;   // hi - lo + 1 entries
    const int32_t* idx;
    const void* rec;      // 4 reals per compacted entry: x y z rad
    int lo, hi;           // extended index range [-n_ghost_l, n + n_ghost_r)
};

template <class R>
__global__ void __launch_bounds__(256) k_nf_clear(const int32_t* __restrict__ old_total, const int32_t* __restrict__ idx, int n,
                                                  R* __restrict__ fx, R* __restrict__ fy, R* __restrict__ fz, R* __restrict__ tx,
                                                  R* __restrict__ ty, R* __restrict__ tz) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *old_total) return;
    const int s = idx[t];
    if (s >= 0 && s < n) fx[s] = fy[s] = fz[s] = tx[s] = ty[s] = tz[s] = (R)0;
}

__global__ void __launch_bounds__(256) k_nf_flags(int lo, int hi, const int32_t* __restrict__ tag, int32_t* __restrict__ pos) {
    const int s = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (s > hi) return;
    pos[s - lo] = s < hi && tag[s] != 0;
}

template <class R>
__global__ void __launch_bounds__(256) k_nf_fill(int lo, int hi, const int32_t* __restrict__ tag, const int32_t* __restrict__ pos,
                                                 const R* __restrict__ x, const R* __restrict__ y, const R* __restrict__ z,
                                                 const R* __restrict__ rad, int32_t* __restrict__ idx, R* __restrict__ rec,
                                                 int32_t* __restrict__ total_copy) {
    const int s = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (s == hi) *total_copy = pos[hi - lo];      // survives the next pass's scan: k_nf_clear reads it
    if (s < hi && tag[s] != 0) {
        const int q = pos[s - lo];
        idx[q] = s;
        R* r = rec + 4 * (size_t)q;
        r[0] = x[s]; r[1] = y[s]; r[2] = z[s]; r[3] = rad[s];
    }
}

// linear keys: (1) all 18 run bounds are fetched at once, (2) the contact TEST runs over the candidates four
// at a time with their 16 loads in flight together and only records the hits, (3) the heavy contact body then
// runs on the recorded hits.  The kernel is a latency-bound gather, so the win is memory-level parallelism.
// NF (coupled contexts): threads and candidates come from the compacted non-fluid list (see above).
// BODIES (coupled contexts with rigid bodies): members of the same body are skipped as contact partners.
template <class R, bool NF, bool BODIES = false>
__global__ void __launch_bounds__(kThreads, 6) k_dem_forces(GridDev<R> g, DemConst<R> C, DemArgs<R> A, NfArgs F) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int own = s;                                        // NF: this thread's compacted index
    if (NF) {
        if (s >= F.pos[F.hi - F.lo]) return;
        s = F.idx[s];
        if (s < 0 || s >= A.n) return;                        // ghosts are partners only
        if (A.tag[s] != 2) { dem_finish<R>(C, A, s, 0, (R)0, (R)0, (R)0, (R)0, (R)0, (R)0); return; }   // boundaries too
    } else {
        if (s >= A.n) return;
        if (A.tag && A.tag[s] != 2) { dem_finish<R>(C, A, s, 0, (R)0, (R)0, (R)0, (R)0, (R)0, (R)0); return; }
    }
    const R xi = A.x[s], yi = A.y[s], zi = A.z[s];
    const R ri = A.rad[s];
    const int cx = cell_coord<R>(xi, g.lo[0], g.inv[0], g.cx_lo, g.cx_hi);
    const int cy = cell_coord<R>(yi, g.lo[1], g.inv[1], 0, g.n[1] - 1);
    const int cz = cell_coord<R>(zi, g.lo[2], g.inv[2], 0, g.n[2] - 1);
    const int zl = max(cz - g.sub, 0), zh = min(cz + g.sub, g.n[2] - 1);
    int rb[9], re[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int ax = cx + k / 3 - 1, ay = cy + k % 3 - 1;
        const bool ok = ax >= 0 && ax < g.n[0] && ay >= 0 && ay < g.n[1];
        const uint32_t k0 = ok ? ((uint32_t)ax * g.n[1] + ay) * g.n[2] : 0u;
        rb[k] = A.cell_start[k0 + zl];
        re[k] = ok ? A.cell_start[k0 + zh + 1] : rb[k];
        if (NF) { rb[k] = F.pos[rb[k] - F.lo]; re[k] = F.pos[re[k] - F.lo]; }
    }
    constexpr int kHits = 32;          // >= max_contacts (<= 32)
    int hits[kHits];
    int nh = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        for (int j0 = rb[k]; j0 < re[k]; j0 += 4) {
            R r2[4], rs[4];
            int tg[4], jj[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int q = min(j0 + t, re[k] - 1);
                jj[t] = q;
                if (NF) {
                    const R* r = reinterpret_cast<const R*>(F.rec) + 4 * (size_t)q;
                    r2[t] = dist2<3, R>(xi - r[0], yi - r[1], zi - r[2]);
                    rs[t] = add_rn(ri, r[3]);
                    tg[t] = 1;
                } else {
                    r2[t] = dist2<3, R>(xi - A.x[q], yi - A.y[q], zi - A.z[q]);
                    rs[t] = add_rn(ri, A.rad[q]);
                    tg[t] = A.tag ? A.tag[q] : 1;
                }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int j = NF ? jj[t] : j0 + t;            // NF: compacted index, mapped to the particle index below
                if (j0 + t < re[k] && r2[t] < mul_rn(rs[t], rs[t]) && r2[t] > (R)0 && j != (NF ? own : s) && tg[t] != 0) {
                    if (nh < kHits) hits[nh] = j;
                    ++nh;
                }
            }
        }
    }
    const R ui = A.u[s], vi = A.v[s], wi = A.w[s], mi = A.m[s];
    const R owx = mul_rn(ri, A.wx[s]), owy = mul_rn(ri, A.wy[s]), owz = mul_rn(ri, A.wz[s]);   // R_i w_i
    const int hrow = A.hperm ? (int)A.hperm[s] : s;
    const int nold = A.hn_in[hrow];
    R fx = 0, fy = 0, fz = 0, tx = 0, ty = 0, tz = 0;
    int cnt = 0;
    const int bi = BODIES ? A.body[s] : -1;
    for (int h = 0; h < min(nh, kHits); ++h) {
        const int j = NF ? F.idx[hits[h]] : hits[h];
        if (BODIES && bi >= 0 && A.body[j] == bi) continue;          // same rigid body: no internal contacts
        dem_contact<R>(C, A, s, j, xi, yi, zi, ui, vi, wi, ri, mi, owx, owy, owz, nold, hrow, fx, fy, fz, tx, ty, tz, cnt);
    }
    if (nh > kHits) cnt = nh;          // more contacts than can ever be stored: reported as overflow below
    dem_finish<R>(C, A, s, cnt, fx, fy, fz, tx, ty, tz);
}

// semi-implicit Euler for spheres: tag 0 only
template <class R>
__global__ void __launch_bounds__(256) k_dem_integrate(int n, R dt, R gx, R gy, R gz, const int32_t* __restrict__ tag, R* __restrict__ x,
                                                       R* __restrict__ y, R* __restrict__ z, R* __restrict__ u, R* __restrict__ v,
                                                       R* __restrict__ w, R* __restrict__ wx, R* __restrict__ wy, R* __restrict__ wz,
                                                       const R* __restrict__ m, const R* __restrict__ inertia, const R* __restrict__ fx,
                                                       const R* __restrict__ fy, const R* __restrict__ fz, const R* __restrict__ tx,
                                                       const R* __restrict__ ty, const R* __restrict__ tz) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n || tag[s] != 0) return;
    const R im = (R)1 / m[s], ii = (R)1 / inertia[s];
    const R un = u[s] + (fx[s] * im + gx) * dt, vn = v[s] + (fy[s] * im + gy) * dt, wn = w[s] + (fz[s] * im + gz) * dt;
    u[s] = un; v[s] = vn; w[s] = wn;
    x[s] += un * dt; y[s] += vn * dt; z[s] += wn * dt;
    wx[s] += tx[s] * ii * dt; wy[s] += ty[s] * ii * dt; wz[s] += tz[s] * ii * dt;
}

template <class R, int DIM, bool MORTON>
pst_status launch_dem(pst_ctx* ctx) {
    if (DIM != 3) return pst_fail(ctx, PST_EINVAL, "DEM needs dim = 3");
    DemConst<R> C;
    C.model = (int)pst_param(ctx, "dem_model");
    C.K = ctx->cfg.max_contacts;
    C.kn = (R)pst_param(ctx, "kn"); C.gn = (R)pst_param(ctx, "gn"); C.kt = (R)pst_param(ctx, "kt"); C.gt = (R)pst_param(ctx, "gt");
    C.mu = (R)pst_param(ctx, "mu"); C.dt = (R)pst_param(ctx, "dt");
    C.Estar = (R)pst_param(ctx, "Estar"); C.Gstar = (R)pst_param(ctx, "Gstar");
    const double le = std::log(pst_param(ctx, "erest", 1.0));
    const double be = le / std::sqrt(le * le + 3.14159265358979323846 * 3.14159265358979323846);
    C.damp_c = (R)(-2.0 * std::sqrt(5.0 / 6.0) * be);
    PstArray *hn = pst_find(ctx, "hist_n"), *hid = pst_find(ctx, "hist_id"), *hx = pst_find(ctx, "hist_x"), *hy = pst_find(ctx, "hist_y"),
             *hz = pst_find(ctx, "hist_z");
    if (!hn) return pst_fail(ctx, PST_ESTATE, "dem_contact needs max_contacts > 0");
    DemArgs<R> A;
    A.x = pst_ptr<R>(ctx, "x"); A.y = pst_ptr<R>(ctx, "y"); A.z = pst_ptr<R>(ctx, "z");
    A.u = pst_ptr<R>(ctx, "u"); A.v = pst_ptr<R>(ctx, "v"); A.w = pst_ptr<R>(ctx, "w");
    A.wx = pst_ptr<R>(ctx, "wx"); A.wy = pst_ptr<R>(ctx, "wy"); A.wz = pst_ptr<R>(ctx, "wz");
    A.rad = pst_ptr<R>(ctx, "rad"); A.m = pst_ptr<R>(ctx, "m"); A.id = pst_ptr<uint32_t>(ctx, "id");
    A.tag = ctx->coupled ? pst_ptr<int32_t>(ctx, "tag") : nullptr;
    A.body = ctx->bodies_ready ? pst_ptr<int32_t>(ctx, "body") : nullptr;
    const int c = hn->cur, d = 1 - hn->cur;
    A.hn_in = pst_ptr<int32_t>(ctx, hn, 0, c); A.hid_in = pst_ptr<uint32_t>(ctx, hid, 0, c);
    A.hx_in = pst_ptr<R>(ctx, hx, 0, c); A.hy_in = pst_ptr<R>(ctx, hy, 0, c); A.hz_in = pst_ptr<R>(ctx, hz, 0, c);
    A.hn_out = pst_ptr<int32_t>(ctx, hn, 0, d); A.hid_out = pst_ptr<uint32_t>(ctx, hid, 0, d);
    A.hx_out = pst_ptr<R>(ctx, hx, 0, d); A.hy_out = pst_ptr<R>(ctx, hy, 0, d); A.hz_out = pst_ptr<R>(ctx, hz, 0, d);
    A.fx = pst_ptr<R>(ctx, "fx"); A.fy = pst_ptr<R>(ctx, "fy"); A.fz = pst_ptr<R>(ctx, "fz");
    A.tx = pst_ptr<R>(ctx, "tx"); A.ty = pst_ptr<R>(ctx, "ty"); A.tz = pst_ptr<R>(ctx, "tz");
    A.hperm = ctx->hist_lag ? ctx->vals_out : nullptr;
    A.cell_start = ctx->cell_start;
    A.flags = ctx->d_flags;
    A.stride = ctx->capacity + 2 * ctx->ghost_cap;
    A.n = (int)ctx->n;
    const int variant = MORTON ? 0 : pst_option(ctx, "dem_kernel", ctx->coupled ? 2 : 1);
    NfArgs F{nullptr, nullptr, nullptr, 0, 0};
    if (!(variant == 2 && ctx->coupled)) ctx->nf_dirty = true;
    if (variant == 0) {
        PST_LAUNCH(ctx, (k_dem_forces_generic<R, MORTON>), blocks_for(ctx->n, kThreads), kThreads, 0, make_grid_dev<R>(ctx->grid), C, A);
    } else if (variant == 2 && ctx->coupled) {
        // compact the non-fluid particles (owned + ghosts), then one thread per compacted entry
        const int lo = -(int)ctx->n_ghost_l, hi = (int)ctx->n + (int)ctx->n_ghost_r, m = hi - lo + 1;
        const size_t cap = ctx->capacity + 2 * ctx->ghost_cap + 2;
        const size_t row = (ctx->capacity + 2 * ctx->ghost_cap) * sizeof(R);
        if (!ctx->nf_pos) {
            if (cudaMalloc((void**)&ctx->nf_pos, (cap + 1) * 4) != cudaSuccess || cudaMalloc((void**)&ctx->nf_idx, cap * 4) != cudaSuccess ||
                cudaMalloc((void**)&ctx->nf_rec, cap * 4 * sizeof(R)) != cudaSuccess)
                return pst_fail(ctx, PST_ENOMEM, "non-fluid compaction buffers");
        }
        if (ctx->nf_dirty) {   // no usable previous list (first pass, or another kernel variant wrote the rows): zero them all once
            R* const rows[6] = {A.fx, A.fy, A.fz, A.tx, A.ty, A.tz};
            for (R* a : rows) PST_CUDA(ctx, cudaMemsetAsync(a - ctx->ghost_cap, 0, row, ctx->stream));
            PST_CUDA(ctx, cudaMemsetAsync(ctx->nf_pos + cap, 0, 4, ctx->stream));
            ctx->nf_dirty = false;
        }
        int32_t* total_copy = ctx->nf_pos + cap;
        PST_LAUNCH(ctx, k_nf_clear<R>, blocks_for(cap, 256), 256, 0, total_copy, ctx->nf_idx, (int)ctx->capacity, A.fx, A.fy, A.fz, A.tx, A.ty, A.tz);
        PST_CUDA(ctx, cudaMemsetAsync(A.hn_out - ctx->ghost_cap, 0, (ctx->capacity + 2 * ctx->ghost_cap) * 4, ctx->stream));
        PST_LAUNCH(ctx, k_nf_flags, blocks_for(m, 256), 256, 0, lo, hi, A.tag, ctx->nf_pos);
        PST_TRY(pst_scan_exclusive(ctx, ctx->nf_pos, m));
        PST_LAUNCH(ctx, k_nf_fill<R>, blocks_for(m, 256), 256, 0, lo, hi, A.tag, ctx->nf_pos, A.x, A.y, A.z, A.rad, ctx->nf_idx, (R*)ctx->nf_rec, total_copy);
        F = NfArgs{ctx->nf_pos, ctx->nf_idx, ctx->nf_rec, lo, hi};
        if (A.body) PST_LAUNCH(ctx, (k_dem_forces<R, true, true>), blocks_for(m, kThreads), kThreads, 0, make_grid_dev<R>(ctx->grid), C, A, F);
        else PST_LAUNCH(ctx, (k_dem_forces<R, true>), blocks_for(m, kThreads), kThreads, 0, make_grid_dev<R>(ctx->grid), C, A, F);
    } else if (A.body) {
        PST_LAUNCH(ctx, (k_dem_forces<R, false, true>), blocks_for(ctx->n, kThreads), kThreads, 0, make_grid_dev<R>(ctx->grid), C, A, F);
    } else {
        PST_LAUNCH(ctx, (k_dem_forces<R, false>), blocks_for(ctx->n, kThreads), kThreads, 0, make_grid_dev<R>(ctx->grid), C, A, F);
    }
    for (PstArray* a : {hn, hid, hx, hy, hz}) a->cur = d;
    ctx->hist_lag = false;   // the pass wrote every row at its new index
    return PST_OK;
}

template <class R>
pst_status launch_dem_integrate(pst_ctx* ctx, double dt) {
    const int n = (int)ctx->n;
    PST_LAUNCH(ctx, k_dem_integrate<R>, blocks_for(n, 256), 256, 0, n, (R)dt, (R)pst_param(ctx, "gx"), (R)pst_param(ctx, "gy"),
               (R)pst_param(ctx, "gz"), pst_ptr<int32_t>(ctx, "tag"), pst_ptr<R>(ctx, "x"), pst_ptr<R>(ctx, "y"), pst_ptr<R>(ctx, "z"),
               pst_ptr<R>(ctx, "u"), pst_ptr<R>(ctx, "v"), pst_ptr<R>(ctx, "w"), pst_ptr<R>(ctx, "wx"), pst_ptr<R>(ctx, "wy"),
               pst_ptr<R>(ctx, "wz"), pst_ptr<R>(ctx, "m"), pst_ptr<R>(ctx, "inertia"), pst_ptr<R>(ctx, "fx"), pst_ptr<R>(ctx, "fy"),
               pst_ptr<R>(ctx, "fz"), pst_ptr<R>(ctx, "tx"), pst_ptr<R>(ctx, "ty"), pst_ptr<R>(ctx, "tz"));
    return PST_OK;
}

}  // namespace

pst_status pst_dem_forces(pst_ctx* ctx) {
    if (ctx->n == 0) return PST_OK;
    PST_TRY(pst_check_cell_size(ctx));
    PST_TRY(PST_DISPATCH(ctx, launch_dem, ctx));
    ctx->contact_kernel_fn = ctx->last_kernel_fn;      // (the contact kernel is the last launch of the pass)
    return PST_OK;
}

pst_status pst_dem_integrate(pst_ctx* ctx, double dt) {
    if (ctx->n == 0) return PST_OK;
    return ctx->f64 ? launch_dem_integrate<double>(ctx, dt) : launch_dem_integrate<float>(ctx, dt);
}
