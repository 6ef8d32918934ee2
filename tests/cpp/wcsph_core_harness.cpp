// Host-compiled view of prestige_b200/csrc/wcsph_core.h: the SAME pair-body and wall-pressure source the device kernels
// of wcsph.cu run (hardware rsqrt / rcp seeds replaced by 1/sqrt and 1/x), driven by the literal all-pairs loop of the
// reference's back-end (prestige/src/codegen/simple_cpu.rs:7-16) and exported with C linkage so tests/test_oracle.py can
// compare it against the oracle without a GPU.  Test infrastructure only.  Build: -O2 -ffp-contract=off.
#include <stdint.h>

#include "../../prestige_b200/csrc/wcsph_core.h"

namespace {
template <class R>
WcsphConst<R> make_const(const double* P /* kfac rho0 c0 gamma alpha beta gx gy gz */) {
    WcsphConst<R> C;
    C.kfac = (R)P[0]; C.rho0 = (R)P[1]; C.c0 = (R)P[2]; C.gamma = (R)P[3];
    C.B = (R)(P[1] * P[2] * P[2] / P[3]);
    C.alpha_c0 = (R)(P[4] * P[2]);
    C.beta = (R)P[5];
    C.g[0] = (R)P[6]; C.g[1] = (R)P[7]; C.g[2] = (R)P[8];
    C.gamma_is_7 = P[3] == 7.0;
    return C;
}

template <class R, int DIM>
R dist2(R dx, R dy, R dz) {
    R r2 = dx * dx + dy * dy;          // left to right, no FMA (-ffp-contract=off)
    if (DIM == 3) r2 = r2 + dz * dz;
    return r2;
}

// `ms` < 0 marks a non-fluid particle of a coupled context (k_eos' signed SPH mass); pass ms = m for plain WCSPH.
template <class R, int DIM>
void forces(const double* P, int64_t n, const R* x, const R* y, const R* z, const R* u, const R* v, const R* w, const R* rho,
            const R* ms, const R* h, const R* p, R* au, R* av, R* aw, R* arho) {
    const WcsphConst<R> C = make_const<R>(P);
    for (int64_t i = 0; i < n; ++i) {
        IState<R, DIM> I;
        load_i<R, DIM>(I, C, x[i], y[i], DIM == 3 ? z[i] : (R)0, u[i], v[i], DIM == 3 ? w[i] : (R)0, rho[i], p[i] / (rho[i] * rho[i]), h[i]);
        Acc<R> a{0, 0, 0, 0};
        const bool fluid_i = ms[i] > (R)0;
        for (int64_t j = 0; j < n; ++j) {
            const R dx = I.x - x[j], dy = I.y - y[j], dz = DIM == 3 ? I.z - z[j] : (R)0;
            R r2 = dist2<R, DIM>(dx, dy, dz);
            const bool in = r2 < I.rc2 && r2 > (R)0;
            r2 = in ? r2 : (R)1;                               // the kernels' branch-free dummy pair
            R mj = in ? ms[j] : (R)0;
            mj = (fluid_i || mj > (R)0) ? (mj < (R)0 ? -mj : mj) : (R)0;
            pair_body<R, DIM, true, true>(C, I, dx, dy, dz, r2, u[j], v[j], DIM == 3 ? w[j] : (R)0, rho[j], mj, p[j] / (rho[j] * rho[j]), a);
        }
        au[i] = a.au + C.g[0]; av[i] = a.av + C.g[1];
        if (DIM == 3) aw[i] = a.aw + C.g[2];
        arho[i] = a.arho;
    }
}

template <class R, int DIM>
void wall(const double* P, int64_t n, const R* x, const R* y, const R* z, const R* h, const int32_t* tag, R* rho, R* p) {
    const WcsphConst<R> C = make_const<R>(P);
    for (int64_t i = 0; i < n; ++i) {
        if (tag[i] == 0) continue;
        const R ad = wendland_alpha<R, DIM>(h[i]), inv_h = (R)1 / h[i];
        const R rc = wc_mul_rn(C.kfac, h[i]), rc2 = wc_mul_rn(rc, rc);
        WallSums<R> S{0, 0, 0, 0, 0};
        for (int64_t j = 0; j < n; ++j) {
            const R dx = x[i] - x[j], dy = y[i] - y[j], dz = DIM == 3 ? z[i] - z[j] : (R)0;
            const R r2 = dist2<R, DIM>(dx, dy, dz);
            if (r2 < rc2 && r2 > (R)0 && tag[j] == 0) wall_accumulate<R, DIM>(S, ad, inv_h, dx, dy, dz, r2, p[j], rho[j]);
        }
        R pw, rw;
        wall_finish<R>(C, S, pw, rw);
        p[i] = pw; rho[i] = rw;          // fluid rows are never written, so in-place is the device's semantics too
    }
}
}  // namespace

#define WCH(SFX, R)                                                                                                            \
    extern "C" void wch_forces_##SFX(int dim, const double* P, int64_t n, const R* x, const R* y, const R* z, const R* u,       \
                                     const R* v, const R* w, const R* rho, const R* ms, const R* h, const R* p, R* au, R* av,   \
                                     R* aw, R* arho) {                                                                         \
        if (dim == 3) forces<R, 3>(P, n, x, y, z, u, v, w, rho, ms, h, p, au, av, aw, arho);                                    \
        else forces<R, 2>(P, n, x, y, z, u, v, w, rho, ms, h, p, au, av, aw, arho);                                             \
    }                                                                                                                          \
    extern "C" void wch_wall_##SFX(int dim, const double* P, int64_t n, const R* x, const R* y, const R* z, const R* h,         \
                                   const int32_t* tag, R* rho, R* p) {                                                         \
        if (dim == 3) wall<R, 3>(P, n, x, y, z, h, tag, rho, p);                                                                \
        else wall<R, 2>(P, n, x, y, z, h, tag, rho, p);                                                                         \
    }
WCH(f64, double)
WCH(f32, float)
