// oracle.cpp -- CPU restatement of the prestige particle hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in prestige_b200/ (the product) may link,
// import or call this file.  Users: tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py.
//
// PARITY UNPINNED.  The reference (dineshadepu/prestige) holds no neighbour
// search, SPH or DEM code and its two tests assert nothing
// (/root/reference/prestige/src/lib.rs:20-27, :32-52).  What the reference does
// fix, and what this file follows literally:
//   * loop shape: `for i in 0..n { for j in 0..n { bodies } }`, j == i included,
//     bodies in fuse() input order        (prestige/src/codegen/simple_cpu.rs:7-16)
//   * gather form, f64 slices, accumulate into [i] only  (prestige/src/lib.rs:8-10)
//   * the one derivable known answer: eq1 => force[i] = sum_j mass[j], self term
//     included                             (prestige/src/lib.rs:7-12)
// The physics is this repo's own written contract; the
// all-pairs functions below are its definition of truth for neighbour sets,
// contact sets and per-particle rates.  The cell-list functions must reproduce
// the all-pairs sets bit-exactly and are what gets timed as the CPU baseline.
//
// Build: see oracle/Makefile (-O3 -ffp-contract=off -fopenmp).
// -ffp-contract=off matters: the cutoff test r2 = dx*dx + dy*dy + dz*dz must be
// evaluated without FMA on both CPU and GPU for the sets to be bit-exact.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ----------------------------------------------------------------------------
// parameters (plain doubles so ctypes can fill them; cast to Real inside)
// ----------------------------------------------------------------------------
struct WcsphParams {
    int32_t dim;      // 2 | 3
    int32_t pad;
    double kfac;      // support radius = kfac * h   (2 for Wendland C2)
    double rho0, c0, gamma, alpha, beta;
    double g[3];
    int32_t p_given;  // 1: `p` is an INPUT of the pair loops (EOS + wall_pressure already applied by the caller)
    int32_t pad2;
};

struct DemParams {
    int32_t model;    // 0 = linear spring-dashpot, 1 = Hertz-Mindlin
    int32_t K;        // history slots per particle
    double kn, gn, kt, gt, mu, dt;
    double Estar, Gstar, erest;   // Hertz only
};

struct Grid {
    int32_t dim;
    int32_t n[3];
    double lo[3];
    double cell;
};

// Cell coordinate: floor((x - lo) / cell) clamped to the grid.  Monotone in x,
// so |x_i - x_j| < cutoff <= cell never puts i and j more than one cell apart.
template <class R>
inline int cell_coord(R x, double lo, double cell, int n) {
    R t = (x - (R)lo) * (R)(1.0 / cell);
    int c = (int)std::floor(t);
    if (c < 0) c = 0;
    if (c > n - 1) c = n - 1;
    return c;
}

// linear key, x slowest, last axis fastest (slabs along x are
// contiguous key ranges)
inline int64_t lin_key(const Grid& g, int cx, int cy, int cz) {
    if (g.dim == 2) return (int64_t)cx * g.n[1] + cy;
    return ((int64_t)cx * g.n[1] + cy) * g.n[2] + cz;
}

template <class R>
struct CellList {
    Grid g;
    int64_t ncells;
    std::vector<int64_t> start;     // ncells + 1
    std::vector<uint32_t> order;    // particle ids, ascending id inside a cell
    std::vector<int32_t> cx, cy, cz;

    void build(const Grid& grid, int64_t n, const R* x, const R* y, const R* z) {
        g = grid;
        ncells = (int64_t)g.n[0] * g.n[1] * (g.dim == 3 ? g.n[2] : 1);
        cx.resize(n); cy.resize(n); cz.resize(n);
        std::vector<int64_t> key(n);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            cx[i] = cell_coord<R>(x[i], g.lo[0], g.cell, g.n[0]);
            cy[i] = cell_coord<R>(y[i], g.lo[1], g.cell, g.n[1]);
            cz[i] = g.dim == 3 ? cell_coord<R>(z[i], g.lo[2], g.cell, g.n[2]) : 0;
            key[i] = lin_key(g, cx[i], cy[i], cz[i]);
        }
        start.assign(ncells + 1, 0);
        for (int64_t i = 0; i < n; ++i) start[key[i] + 1]++;
        for (int64_t c = 0; c < ncells; ++c) start[c + 1] += start[c];
        std::vector<int64_t> fill(start.begin(), start.end() - 1);
        order.resize(n);
        for (int64_t i = 0; i < n; ++i) order[fill[key[i]]++] = (uint32_t)i;  // stable
    }

    // visit candidates j of particle i in cell order (27 / 9 stencil)
    template <class F>
    inline void for_candidates(int64_t i, F&& f) const {
        const int z0 = g.dim == 3 ? cz[i] - 1 : 0, z1 = g.dim == 3 ? cz[i] + 1 : 0;
        for (int ax = cx[i] - 1; ax <= cx[i] + 1; ++ax) {
            if (ax < 0 || ax >= g.n[0]) continue;
            for (int ay = cy[i] - 1; ay <= cy[i] + 1; ++ay) {
                if (ay < 0 || ay >= g.n[1]) continue;
                for (int az = z0; az <= z1; ++az) {
                    if (g.dim == 3 && (az < 0 || az >= g.n[2])) continue;
                    const int64_t c = lin_key(g, ax, ay, az);
                    for (int64_t s = start[c]; s < start[c + 1]; ++s) f((int64_t)order[s]);
                }
            }
        }
    }
};

// The cutoff arithmetic, written once.  Left to right, no FMA (file is built
// with -ffp-contract=off).  SURVEY.md Appendix A.1.
template <class R>
inline R dist2(int dim, R dx, R dy, R dz) {
    R r2 = dx * dx + dy * dy;
    if (dim == 3) r2 = r2 + dz * dz;
    return r2;
}

// ----------------------------------------------------------------------------
// eq1: the reference's sample equation, executed exactly as generate_simple_cpu
// would emit it (simple_cpu.rs:7-16 with body lib.rs:10).
// ----------------------------------------------------------------------------
template <class R>
void eq1_allpairs(int64_t n, const R* mass, R* force) {
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < n; ++j)
            force[i] += mass[j];
}

// ----------------------------------------------------------------------------
// neighbour / contact sets
//   mode 0 (SPH): j != i and r2 < (kfac*h_i)^2
//   mode 1 (DEM): j != i and r2 < (R_i + R_j)^2     (s = radius array)
// ----------------------------------------------------------------------------
template <class R>
inline bool in_range(int mode, R r2, R si, R sj, R kfac) {
    if (mode == 0) { R rc = kfac * si; return r2 < rc * rc; }
    R rc = si + sj; return r2 < rc * rc;
}

template <class R>
int64_t pairs_allpairs(int dim, int mode, double kfac, int64_t n, const R* x, const R* y, const R* z,
                       const R* s, uint32_t* oi, uint32_t* oj, int64_t cap, double* min_margin) {
    int64_t cnt = 0;
    double mm = 1e300;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            R r2 = dist2<R>(dim, x[i] - x[j], y[i] - y[j], dim == 3 ? z[i] - z[j] : (R)0);
            R rc = mode == 0 ? (R)kfac * s[i] : s[i] + s[j];
            double m = std::fabs((double)r2 - (double)(rc * rc)) / (double)(rc * rc);
            if (m < mm) mm = m;
            if (in_range<R>(mode, r2, s[i], s[j], (R)kfac)) {
                if (cnt < cap) { oi[cnt] = (uint32_t)i; oj[cnt] = (uint32_t)j; }
                ++cnt;
            }
        }
    if (min_margin) *min_margin = mm;
    return cnt;
}

template <class R>
int64_t pairs_cells(const Grid& g, int mode, double kfac, int64_t n, const R* x, const R* y, const R* z,
                    const R* s, uint32_t* oi, uint32_t* oj, int64_t cap) {
    CellList<R> cl;
    cl.build(g, n, x, y, z);
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i)
        cl.for_candidates(i, [&](int64_t j) {
            if (j == i) return;
            R r2 = dist2<R>(g.dim, x[i] - x[j], y[i] - y[j], g.dim == 3 ? z[i] - z[j] : (R)0);
            if (in_range<R>(mode, r2, s[i], s[j], (R)kfac)) {
                if (cnt < cap) { oi[cnt] = (uint32_t)i; oj[cnt] = (uint32_t)j; }
                ++cnt;
            }
        });
    return cnt;
}

// ----------------------------------------------------------------------------
// WCSPH
// ----------------------------------------------------------------------------
template <class R>
inline R tait_eos(const WcsphParams& P, R rho) {
    // (rho/rho0)^gamma - 1 evaluated as expm1(gamma * log1p((rho - rho0)/rho0)): the same function, but
    // without the cancellation of pow(..) - 1 near rho0 (which costs ~2 digits per 1e-2 of compression and
    // would make an f32 comparison at 1e-5 meaningless).
    const R B = (R)(P.rho0 * P.c0 * P.c0 / P.gamma);
    const R e = (rho - (R)P.rho0) / (R)P.rho0;
    return B * std::expm1((R)P.gamma * std::log1p(e));
}

template <class R>
void wcsph_eos(const WcsphParams& P, int64_t n, const R* rho, R* p) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) p[i] = tait_eos<R>(P, rho[i]);
}

template <class R>
struct WcsphAcc { R au, av, aw, arho; };

// one (i, j) body: continuity then momentum, in fuse() input order
template <class R>
inline void wcsph_pair(const WcsphParams& P, int64_t i, int64_t j, const R* x, const R* y, const R* z,
                       const R* u, const R* v, const R* w, const R* rho, const R* m, const R* h,
                       const R* p, WcsphAcc<R>& a) {
    const int dim = P.dim;
    const R dx = x[i] - x[j], dy = y[i] - y[j], dz = dim == 3 ? z[i] - z[j] : (R)0;
    const R r2 = dist2<R>(dim, dx, dy, dz);
    const R hi = h[i];
    const R rc = (R)P.kfac * hi;
    if (!(r2 < rc * rc)) return;
    const R r = std::sqrt(r2);
    if (!(r > (R)0)) return;                       // grad W(0) = 0
    const double pi = 3.14159265358979323846;
    const R ad = dim == 3 ? (R)(21.0 / (16.0 * pi)) / (hi * hi * hi) : (R)(7.0 / (4.0 * pi)) / (hi * hi);
    const R q = r / hi;
    const R t = (R)1 - (R)0.5 * q;
    const R dwdq = (R)-5 * ad * q * t * t * t;
    const R gf = dwdq / (hi * r);                  // grad_i W = gf * x_ij
    const R du = u[i] - u[j], dv = v[i] - v[j], dw = dim == 3 ? w[i] - w[j] : (R)0;
    R vx = du * dx + dv * dy;
    if (dim == 3) vx = vx + dw * dz;
    // continuity: d rho_i/dt += m_j v_ij . grad W
    a.arho += m[j] * gf * vx;
    // momentum
    R Pi = (R)0;
    if (vx < (R)0) {
        const R mu = hi * vx / (r2 + (R)0.01 * hi * hi);
        const R rhob = (R)0.5 * (rho[i] + rho[j]);
        Pi = ((R)(-P.alpha * P.c0) * mu + (R)P.beta * mu * mu) / rhob;
    }
    const R c = -m[j] * (p[i] / (rho[i] * rho[i]) + p[j] / (rho[j] * rho[j]) + Pi) * gf;
    a.au += c * dx;
    a.av += c * dy;
    if (dim == 3) a.aw += c * dz;
}

template <class R>
void wcsph_allpairs(const WcsphParams& P, int64_t n, const R* x, const R* y, const R* z, const R* u,
                    const R* v, const R* w, const R* rho, const R* m, const R* h, R* p, R* au, R* av,
                    R* aw, R* arho) {
    if (!P.p_given) wcsph_eos<R>(P, n, rho, p);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
        WcsphAcc<R> a{0, 0, 0, 0};
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            wcsph_pair<R>(P, i, j, x, y, z, u, v, w, rho, m, h, p, a);
        }
        au[i] = a.au + (R)P.g[0];
        av[i] = a.av + (R)P.g[1];
        if (P.dim == 3) aw[i] = a.aw + (R)P.g[2];
        arho[i] = a.arho;
    }
}

template <class R>
void wcsph_cells(const WcsphParams& P, const Grid& g, int64_t n, const R* x, const R* y, const R* z,
                 const R* u, const R* v, const R* w, const R* rho, const R* m, const R* h, R* p, R* au,
                 R* av, R* aw, R* arho) {
    CellList<R> cl;
    cl.build(g, n, x, y, z);
    if (!P.p_given) wcsph_eos<R>(P, n, rho, p);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) {
        WcsphAcc<R> a{0, 0, 0, 0};
        cl.for_candidates(i, [&](int64_t j) {
            if (j == i) return;
            wcsph_pair<R>(P, i, j, x, y, z, u, v, w, rho, m, h, p, a);
        });
        au[i] = a.au + (R)P.g[0];
        av[i] = a.av + (R)P.g[1];
        if (P.dim == 3) aw[i] = a.aw + (R)P.g[2];
        arho[i] = a.arho;
    }
}

// The timed CPU step: the same stages the GPU path runs (key, sort, cell table,
// permute of the persistent state into cell order, EOS, fused pair loop), all
// threads.  State layout after the call is left untouched; outputs are written
// in the caller's (id) order.
template <class R>
void wcsph_step_sorted(const WcsphParams& P, const Grid& g, int64_t n, const R* x, const R* y, const R* z,
                       const R* u, const R* v, const R* w, const R* rho, const R* m, const R* h, R* p,
                       R* au, R* av, R* aw, R* arho) {
    CellList<R> cl;
    cl.build(g, n, x, y, z);
    // permute persistent state into cell order (what a9 does on the GPU)
    std::vector<R> sx(n), sy(n), sz(n), su(n), sv(n), sw(n), sr(n), sm(n), sh(n), sp(n);
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < n; ++s) {
        const uint32_t i = cl.order[s];
        sx[s] = x[i]; sy[s] = y[i]; sz[s] = g.dim == 3 ? z[i] : (R)0;
        su[s] = u[i]; sv[s] = v[i]; sw[s] = g.dim == 3 ? w[i] : (R)0;
        sr[s] = rho[i]; sm[s] = m[i]; sh[s] = h[i];
        sp[s] = P.p_given ? p[i] : tait_eos<R>(P, rho[i]);
    }
    // cell coordinates of sorted particle s are those of order[s]
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t s = 0; s < n; ++s) {
        const uint32_t i = cl.order[s];
        WcsphAcc<R> a{0, 0, 0, 0};
        const int z0 = g.dim == 3 ? cl.cz[i] - 1 : 0, z1 = g.dim == 3 ? cl.cz[i] + 1 : 0;
        for (int ax = cl.cx[i] - 1; ax <= cl.cx[i] + 1; ++ax) {
            if (ax < 0 || ax >= g.n[0]) continue;
            for (int ay = cl.cy[i] - 1; ay <= cl.cy[i] + 1; ++ay) {
                if (ay < 0 || ay >= g.n[1]) continue;
                // last axis is fastest in the key, so the 3 cells are one run
                const int lo = g.dim == 3 ? std::max(z0, 0) : 0;
                const int hi = g.dim == 3 ? std::min(z1, g.n[2] - 1) : 0;
                const int64_t b = cl.start[lin_key(g, ax, ay, lo)], e = cl.start[lin_key(g, ax, ay, hi) + 1];
                for (int64_t t = b; t < e; ++t) {
                    if (t == s) continue;
                    wcsph_pair<R>(P, s, t, sx.data(), sy.data(), sz.data(), su.data(), sv.data(), sw.data(),
                                  sr.data(), sm.data(), sh.data(), sp.data(), a);
                }
            }
        }
        p[i] = sp[s];
        au[i] = a.au + (R)P.g[0];
        av[i] = a.av + (R)P.g[1];
        if (P.dim == 3) aw[i] = a.aw + (R)P.g[2];
        arho[i] = a.arho;
    }
}

// ----------------------------------------------------------------------------
// Dummy-particle wall pressure (formulation: DESIGN.md 4d, after Adami, Hu & Adams 2012).
// No reference code exists for it; this is the repo's own written contract.
// For every NON-fluid particle w (tag != 0), over its FLUID neighbours f (tag == 0) under the neighbour rule of
// Appendix A.1 with the support of w (0 < r2 < (kfac h_w)^2, same FMA-free arithmetic):
//     S0 = sum W_wf,   Sp = sum p_f W_wf,   Sx = sum rho_f x_wf W_wf          (x_wf = x_w - x_f)
//     p_w = (Sp + g . Sx) / S0  if S0 > 0 else 0       (wall acceleration a_w taken as 0: quasi-static walls)
//     rho_w = rho0 (max(p_w / B, -1/2) + 1)^(1/gamma)  evaluated as rho0 exp(log1p(.) / gamma)
// and p[w] = p_w, rho[w] = rho_w replace the EOS pressure and the state density of w: the pair loops then see the
// extrapolated fluid pressure on dummy particles instead of one evolved by continuity.  Fluid rows are untouched.
// W = Wendland C2 (Appendix A.2) with h = h_w.
// ----------------------------------------------------------------------------
template <class R>
void wall_pressure(const WcsphParams& P, const Grid* g /* null => all pairs */, int64_t n, const R* x, const R* y,
                   const R* z, const R* h, const int32_t* tag, R* rho /* in/out */, R* p /* in/out */) {
    CellList<R> cl;
    if (g) cl.build(*g, n, x, y, z);
    const int dim = P.dim;
    const R B = (R)(P.rho0 * P.c0 * P.c0 / P.gamma);
    std::vector<R> pw(n), rw(n);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) {
        if (tag[i] == 0) continue;
        const R hi = h[i];
        const R rc = (R)P.kfac * hi;
        const double pi = 3.14159265358979323846;
        const R ad = dim == 3 ? (R)(21.0 / (16.0 * pi)) / (hi * hi * hi) : (R)(7.0 / (4.0 * pi)) / (hi * hi);
        R S0 = 0, Sp = 0, Sx = 0, Sy = 0, Sz = 0;
        auto body = [&](int64_t j) {
            if (j == i || tag[j] != 0) return;
            const R dx = x[i] - x[j], dy = y[i] - y[j], dz = dim == 3 ? z[i] - z[j] : (R)0;
            const R r2 = dist2<R>(dim, dx, dy, dz);
            if (!(r2 < rc * rc) || !(r2 > (R)0)) return;
            const R q = std::sqrt(r2) / hi;
            const R t = (R)1 - (R)0.5 * q;
            const R t2 = t * t;
            const R W = ad * (t2 * t2) * ((R)2 * q + (R)1);
            S0 += W;
            Sp += p[j] * W;
            const R rW = rho[j] * W;
            Sx += rW * dx; Sy += rW * dy; Sz += rW * dz;
        };
        if (g) cl.for_candidates(i, body);
        else for (int64_t j = 0; j < n; ++j) body(j);
        R pv = 0;
        if (S0 > (R)0) pv = (Sp + ((R)P.g[0] * Sx + (R)P.g[1] * Sy + (R)P.g[2] * Sz)) / S0;
        pw[i] = pv;
        rw[i] = (R)P.rho0 * std::exp(std::log1p(std::max(pv / B, (R)-0.5)) / (R)P.gamma);
    }
    for (int64_t i = 0; i < n; ++i)
        if (tag[i] != 0) { p[i] = pw[i]; rho[i] = rw[i]; }
}

// ----------------------------------------------------------------------------
// DEM.  History: K slots per particle, slot-major:
// hid[k*n + i] = stable id of partner, hx/hy/hz[k*n + i] = xi.  hn[i] = count.
// Returns 0, or 1 if some particle had more than K contacts (PST_EOVERFLOW).
// ----------------------------------------------------------------------------
template <class R>
struct DemAcc { R fx, fy, fz, tx, ty, tz; };

template <class R>
inline void dem_pair(const DemParams& P, int64_t n, int64_t i, int64_t j, const R* x, const R* y, const R* z,
                     const R* u, const R* v, const R* w, const R* wx, const R* wy, const R* wz,
                     const R* rad, const R* m, const uint32_t* id,
                     const int32_t* hn_in, const uint32_t* hid_in, const R* hx_in, const R* hy_in, const R* hz_in,
                     int64_t n_in_stride,
                     int32_t& cnt, uint32_t* hid_out, R* hx_out, R* hy_out, R* hz_out, int64_t out_row,
                     DemAcc<R>& a, int& overflow) {
    const R dx = x[i] - x[j], dy = y[i] - y[j], dz = z[i] - z[j];
    const R r2 = dist2<R>(3, dx, dy, dz);
    const R rs = rad[i] + rad[j];
    if (!(r2 < rs * rs)) return;
    const R r = std::sqrt(r2);
    if (!(r > (R)0)) return;
    const R nx = dx / r, ny = dy / r, nz = dz / r;       // points j -> i
    const R delta = rs - r;
    // contact-point relative velocity: v_ij - (R_i w_i + R_j w_j) x n
    const R ox = rad[i] * wx[i] + rad[j] * wx[j];
    const R oy = rad[i] * wy[i] + rad[j] * wy[j];
    const R oz = rad[i] * wz[i] + rad[j] * wz[j];
    const R vcx = (u[i] - u[j]) - (oy * nz - oz * ny);
    const R vcy = (v[i] - v[j]) - (oz * nx - ox * nz);
    const R vcz = (w[i] - w[j]) - (ox * ny - oy * nx);
    const R vn = vcx * nx + vcy * ny + vcz * nz;
    const R vtx = vcx - vn * nx, vty = vcy - vn * ny, vtz = vcz - vn * nz;
    R kn = (R)P.kn, gn = (R)P.gn, kt = (R)P.kt, gt = (R)P.gt;
    if (P.model == 1) {
        const R Rs = rad[i] * rad[j] / rs;
        const R ms = m[i] * m[j] / (m[i] + m[j]);
        const R sq = std::sqrt(Rs * delta);
        const double le = std::log(P.erest);
        const R be = (R)(le / std::sqrt(le * le + 3.14159265358979323846 * 3.14159265358979323846));  // < 0
        const R Sn = (R)2 * (R)P.Estar * sq, St = (R)8 * (R)P.Gstar * sq;
        kn = (R)(4.0 / 3.0) * (R)P.Estar * sq;
        kt = St;
        const R c = (R)-2 * (R)std::sqrt(5.0 / 6.0) * be;
        gn = c * std::sqrt(Sn * ms);
        gt = c * std::sqrt(St * ms);
    }
    const R fnm = kn * delta - gn * vn;                  // signed normal magnitude
    const R fnx = fnm * nx, fny = fnm * ny, fnz = fnm * nz;
    // history lookup by stable partner id; new contact => xi = 0
    R xx = 0, xy = 0, xz = 0;
    const uint32_t pid = id[j];
    for (int32_t k = 0; k < hn_in[i]; ++k)
        if (hid_in[(int64_t)k * n_in_stride + i] == pid) {
            xx = hx_in[(int64_t)k * n_in_stride + i];
            xy = hy_in[(int64_t)k * n_in_stride + i];
            xz = hz_in[(int64_t)k * n_in_stride + i];
            break;
        }
    // rotate into the current tangent plane, then integrate
    const R xn = xx * nx + xy * ny + xz * nz;
    xx = xx - xn * nx; xy = xy - xn * ny; xz = xz - xn * nz;
    xx = xx + vtx * (R)P.dt; xy = xy + vty * (R)P.dt; xz = xz + vtz * (R)P.dt;
    R ftx = -kt * xx - gt * vtx, fty = -kt * xy - gt * vty, ftz = -kt * xz - gt * vtz;
    const R ftm = std::sqrt(ftx * ftx + fty * fty + ftz * ftz);
    const R fmax = (R)P.mu * std::fabs(fnm);
    if (ftm > fmax) {
        const R sc = fmax / ftm;
        ftx = ftx * sc; fty = fty * sc; ftz = ftz * sc;
        xx = -(ftx + gt * vtx) / kt; xy = -(fty + gt * vty) / kt; xz = -(ftz + gt * vtz) / kt;
    }
    a.fx += fnx + ftx; a.fy += fny + fty; a.fz += fnz + ftz;
    // T_i += (-R_i n) x F_t
    const R lx = -rad[i] * nx, ly = -rad[i] * ny, lz = -rad[i] * nz;
    a.tx += ly * ftz - lz * fty;
    a.ty += lz * ftx - lx * ftz;
    a.tz += lx * fty - ly * ftx;
    if (cnt < P.K) {
        hid_out[(int64_t)cnt * n + out_row] = pid;
        hx_out[(int64_t)cnt * n + out_row] = xx;
        hy_out[(int64_t)cnt * n + out_row] = xy;
        hz_out[(int64_t)cnt * n + out_row] = xz;
        ++cnt;
    } else {
        overflow = 1;
    }
}

template <class R>
int dem_forces(const DemParams& P, const Grid* g /* null => all pairs */, int64_t n, const R* x, const R* y,
               const R* z, const R* u, const R* v, const R* w, const R* wx, const R* wy, const R* wz,
               const R* rad, const R* m, const uint32_t* id, const int32_t* hn_in, const uint32_t* hid_in,
               const R* hx_in, const R* hy_in, const R* hz_in, int32_t* hn_out, uint32_t* hid_out, R* hx_out,
               R* hy_out, R* hz_out, R* fx, R* fy, R* fz, R* tx, R* ty, R* tz) {
    CellList<R> cl;
    if (g) cl.build(*g, n, x, y, z);
    int overflow = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(| : overflow)
    for (int64_t i = 0; i < n; ++i) {
        DemAcc<R> a{0, 0, 0, 0, 0, 0};
        int32_t cnt = 0;
        int ov = 0;
        auto body = [&](int64_t j) {
            if (j == i) return;
            dem_pair<R>(P, n, i, j, x, y, z, u, v, w, wx, wy, wz, rad, m, id, hn_in, hid_in, hx_in, hy_in,
                        hz_in, n, cnt, hid_out, hx_out, hy_out, hz_out, i, a, ov);
        };
        if (g) cl.for_candidates(i, body);
        else for (int64_t j = 0; j < n; ++j) body(j);
        hn_out[i] = cnt;
        fx[i] = a.fx; fy[i] = a.fy; fz[i] = a.fz;
        tx[i] = a.tx; ty[i] = a.ty; tz[i] = a.tz;
        overflow |= ov;
    }
    return overflow;
}

// ----------------------------------------------------------------------------
// Coupled SPH-DEM (DESIGN.md "Coupled formulation", SURVEY.md 8f-4, BASELINE configs[4]).
// Tags: 0 fluid, 1 boundary (static), 2 solid (mobile rigid sphere, one particle per body).
//   * SPH mass of j: m_j for fluid and boundary, m_j * rho0 / rho_solid (the displaced fluid mass) for solids;
//   * an SPH pair (i, j) is active iff tag_i == 0 or tag_j == 0 (same neighbour rule and body as WCSPH);
//   * a DEM pair (i, j) is active iff tag_i == 2 and tag_j != 0 (same contact rule, body and history as DEM) and,
//     with rigid bodies, i and j are not members of the same body;
//   * outputs: au av aw arho (SPH sums, g added to a) for every particle, fx..tz + history for solids only.
// `ms` is the SPH-mass array (computed by the caller from m, tag, rho0 / rho_solid).
// ----------------------------------------------------------------------------
template <class R>
int coupled_forces(const WcsphParams& PW, const DemParams& PD, const Grid* g /* null => all pairs */, int64_t n,
                   const R* x, const R* y, const R* z, const R* u, const R* v, const R* w, const R* rho, const R* ms,
                   const R* h, const R* wx, const R* wy, const R* wz, const R* rad, const R* m, const int32_t* tag,
                   const int32_t* rbody /* null, or rigid-body index per particle (-1 = none) */,
                   const uint32_t* id, const int32_t* hn_in, const uint32_t* hid_in, const R* hx_in, const R* hy_in,
                   const R* hz_in, R* p, R* au, R* av, R* aw, R* arho, int32_t* hn_out, uint32_t* hid_out, R* hx_out,
                   R* hy_out, R* hz_out, R* fx, R* fy, R* fz, R* tx, R* ty, R* tz) {
    CellList<R> cl;
    if (g) cl.build(*g, n, x, y, z);
    if (!PW.p_given) wcsph_eos<R>(PW, n, rho, p);
    int overflow = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(| : overflow)
    for (int64_t i = 0; i < n; ++i) {
        WcsphAcc<R> a{0, 0, 0, 0};
        DemAcc<R> d{0, 0, 0, 0, 0, 0};
        int32_t cnt = 0;
        int ov = 0;
        auto body = [&](int64_t j) {
            if (j == i) return;
            if (tag[i] == 0 || tag[j] == 0)
                wcsph_pair<R>(PW, i, j, x, y, z, u, v, w, rho, ms, h, p, a);
            if (tag[i] == 2 && tag[j] != 0 && !(rbody && rbody[i] >= 0 && rbody[j] == rbody[i]))
                dem_pair<R>(PD, n, i, j, x, y, z, u, v, w, wx, wy, wz, rad, m, id, hn_in, hid_in, hx_in, hy_in, hz_in,
                            n, cnt, hid_out, hx_out, hy_out, hz_out, i, d, ov);
        };
        if (g) cl.for_candidates(i, body);
        else for (int64_t j = 0; j < n; ++j) body(j);
        au[i] = a.au + (R)PW.g[0];
        av[i] = a.av + (R)PW.g[1];
        aw[i] = a.aw + (R)PW.g[2];
        arho[i] = a.arho;
        hn_out[i] = cnt;
        fx[i] = d.fx; fy[i] = d.fy; fz[i] = d.fz;
        tx[i] = d.tx; ty[i] = d.ty; tz[i] = d.tz;
        overflow |= ov;
    }
    return overflow;
}

}  // namespace

// ----------------------------------------------------------------------------
// C entry points (ctypes).  _f64 / _f32 suffixes.
// ----------------------------------------------------------------------------
#define ORC_API extern "C" __attribute__((visibility("default")))

ORC_API int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

#define ORC_INSTANTIATE(SFX, R)                                                                               \
    ORC_API void orc_eq1_allpairs_##SFX(int64_t n, const R* mass, R* force) { eq1_allpairs<R>(n, mass, force); } \
    ORC_API int64_t orc_pairs_allpairs_##SFX(int dim, int mode, double kfac, int64_t n, const R* x, const R* y, \
                                             const R* z, const R* s, uint32_t* oi, uint32_t* oj, int64_t cap, \
                                             double* min_margin) {                                            \
        return pairs_allpairs<R>(dim, mode, kfac, n, x, y, z, s, oi, oj, cap, min_margin);                    \
    }                                                                                                         \
    ORC_API int64_t orc_pairs_cells_##SFX(const Grid* g, int mode, double kfac, int64_t n, const R* x,         \
                                          const R* y, const R* z, const R* s, uint32_t* oi, uint32_t* oj,     \
                                          int64_t cap) {                                                      \
        return pairs_cells<R>(*g, mode, kfac, n, x, y, z, s, oi, oj, cap);                                    \
    }                                                                                                         \
    ORC_API void orc_wcsph_eos_##SFX(const WcsphParams* P, int64_t n, const R* rho, R* p) {                    \
        wcsph_eos<R>(*P, n, rho, p);                                                                          \
    }                                                                                                         \
    ORC_API void orc_wall_pressure_##SFX(const WcsphParams* P, const Grid* g, int64_t n, const R* x, const R* y, \
                                         const R* z, const R* h, const int32_t* tag, R* rho, R* p) {           \
        wall_pressure<R>(*P, g, n, x, y, z, h, tag, rho, p);                                                  \
    }                                                                                                         \
    ORC_API void orc_wcsph_allpairs_##SFX(const WcsphParams* P, int64_t n, const R* x, const R* y, const R* z, \
                                          const R* u, const R* v, const R* w, const R* rho, const R* m,       \
                                          const R* h, R* p, R* au, R* av, R* aw, R* arho) {                   \
        wcsph_allpairs<R>(*P, n, x, y, z, u, v, w, rho, m, h, p, au, av, aw, arho);                           \
    }                                                                                                         \
    ORC_API void orc_wcsph_cells_##SFX(const WcsphParams* P, const Grid* g, int64_t n, const R* x, const R* y, \
                                       const R* z, const R* u, const R* v, const R* w, const R* rho,          \
                                       const R* m, const R* h, R* p, R* au, R* av, R* aw, R* arho) {          \
        wcsph_cells<R>(*P, *g, n, x, y, z, u, v, w, rho, m, h, p, au, av, aw, arho);                          \
    }                                                                                                         \
    ORC_API void orc_wcsph_step_sorted_##SFX(const WcsphParams* P, const Grid* g, int64_t n, const R* x,       \
                                             const R* y, const R* z, const R* u, const R* v, const R* w,      \
                                             const R* rho, const R* m, const R* h, R* p, R* au, R* av, R* aw, \
                                             R* arho) {                                                       \
        wcsph_step_sorted<R>(*P, *g, n, x, y, z, u, v, w, rho, m, h, p, au, av, aw, arho);                    \
    }                                                                                                         \
    ORC_API int orc_dem_forces_##SFX(const DemParams* P, const Grid* g, int64_t n, const R* x, const R* y,     \
                                     const R* z, const R* u, const R* v, const R* w, const R* wx,             \
                                     const R* wy, const R* wz, const R* rad, const R* m, const uint32_t* id,  \
                                     const int32_t* hn_in, const uint32_t* hid_in, const R* hx_in,            \
                                     const R* hy_in, const R* hz_in, int32_t* hn_out, uint32_t* hid_out,      \
                                     R* hx_out, R* hy_out, R* hz_out, R* fx, R* fy, R* fz, R* tx, R* ty,      \
                                     R* tz) {                                                                 \
        return dem_forces<R>(*P, g, n, x, y, z, u, v, w, wx, wy, wz, rad, m, id, hn_in, hid_in, hx_in, hy_in, \
                             hz_in, hn_out, hid_out, hx_out, hy_out, hz_out, fx, fy, fz, tx, ty, tz);         \
    }                                                                                                         \
    ORC_API int orc_coupled_forces_##SFX(const WcsphParams* PW, const DemParams* PD, const Grid* g, int64_t n, \
                                         const R* const* in /* x y z u v w rho ms h wx wy wz rad m */,         \
                                         const int32_t* tag, const int32_t* body, const uint32_t* id,          \
                                         const int32_t* hn_in,                                                 \
                                         const uint32_t* hid_in, const R* hx_in, const R* hy_in,               \
                                         const R* hz_in, R* const* out /* p au av aw arho fx fy fz tx ty tz */, \
                                         int32_t* hn_out, uint32_t* hid_out, R* hx_out, R* hy_out, R* hz_out) { \
        return coupled_forces<R>(*PW, *PD, g, n, in[0], in[1], in[2], in[3], in[4], in[5], in[6], in[7], in[8], \
                                 in[9], in[10], in[11], in[12], in[13], tag, body, id, hn_in, hid_in, hx_in, hy_in,  \
                                 hz_in, out[0], out[1], out[2], out[3], out[4], hn_out, hid_out, hx_out,       \
                                 hy_out, hz_out, out[5], out[6], out[7], out[8], out[9], out[10]);             \
    }

ORC_INSTANTIATE(f64, double)
ORC_INSTANTIATE(f32, float)
