// wcsph_core.h -- the per-pair arithmetic of the WCSPH kernels (continuity, momentum with
// artificial viscosity; DESIGN.md 4d: the dummy-particle wall-pressure sums).  No reference code exists for the physics
// (SURVEY.md 8a rows a10-a12); what the reference fixes is the loop shape -- gather into [i], bodies of a fused set in
// fuse() order (prestige/src/codegen/simple_cpu.rs:7-16, prestige/src/equations/fuse.rs:18,30).
//
// Plain templated C++, used by the device kernels of wcsph.cu and by a host-compiled test harness
// (tests/cpp/wcsph_core_harness.cpp), so the algebra of the bodies (the regrouped viscosity term, the branch-free
// masking, the wall-pressure sums) is checked against the oracle without a GPU.  On the host the two hardware
// approximations (MUFU.RSQ64H / RCP64H seeds + one correction) are replaced by 1/sqrt and 1/x; everything else is the
// same source.
#pragma once

#include <math.h>

#ifndef PST_HD
#if defined(__CUDACC__)
#define PST_HD __host__ __device__ __forceinline__
#else
#define PST_HD inline
#endif
#endif

// Rounded-to-nearest, never-contracted product for the cutoff radius (the host harness is built -ffp-contract=off).
PST_HD double wc_mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
PST_HD float wc_mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}

template <class R>
struct WcsphConst {
    R kfac, rho0, c0, gamma, B, alpha_c0, beta, g[3];
    int gamma_is_7;
    // UNI kernels (every particle has the same smoothing length and mass): the h-derived constants of IState, computed
    // once on the host by the same load_i, so they cost constant-bank operands instead of ten registers per thread
    R u_h, u_half_inv_h, u_gfc, u_eta2, u_rc2, u_mgfc, u_visc;     // u_mgfc = m gfc, u_visc = -2 alpha c0 h
};

template <class R, int DIM>
struct IState {         // everything about particle i the body needs, in registers
    R x, y, z, u, v, w, rho, por2, h, half_inv_h, gfc, eta2, rc2;
};
template <class R>
struct Acc { R au, av, aw, arho; };

// Wendland C2 normalisation alpha_d(h): 21/(16 pi h^3) in 3D, 7/(4 pi h^2) in 2D
template <class R, int DIM>
PST_HD R wendland_alpha(R h) {
    const R pi = (R)3.14159265358979323846;
    return DIM == 3 ? (R)(21.0 / 16.0) / (pi * h * h * h) : (R)(7.0 / 4.0) / (pi * h * h);
}

template <class R, int DIM>
PST_HD void load_i(IState<R, DIM>& I, const WcsphConst<R>& C, R x, R y, R z, R u, R v, R w, R rho, R por2, R h) {
    I.x = x; I.y = y; I.z = z; I.u = u; I.v = v; I.w = w; I.rho = rho; I.por2 = por2; I.h = h;
    I.half_inv_h = (R)0.5 / h;
    const R ad = wendland_alpha<R, DIM>(h);
    I.gfc = (R)-5 * ad / (h * h);       // (dW/dq)/(h r) = gfc * (1 - q/2)^3
    I.eta2 = (R)0.01 * h * h;
    const R rc = wc_mul_rn(C.kfac, h);
    I.rc2 = wc_mul_rn(rc, rc);
}

// 1/sqrt(x) and 1/x for normal positive x: hardware seed (MUFU.RSQ64H / RCP64H, ~2^-22) + ONE cubically
// convergent correction, no special-case branches.  Error ~2 ulp (e^3 ~ 2^-66 is below the rounding).
PST_HD double fast_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);                 // 1 - x y^2
    return fma(y, e * fma(0.375, e, 0.5), y);               // y (1 + e/2 + 3 e^2 / 8)
#else
    return 1.0 / sqrt(x);
#endif
}
PST_HD float fast_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}
PST_HD double fast_rcp(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);                       // 1 - x y
    return fma(y, fma(e, e, e), y);                         // y (1 + e + e^2)
#else
    return 1.0 / x;
#endif
}
PST_HD float fast_rcp(float x) {
#if defined(__CUDA_ARCH__)
    return __frcp_rn(x);
#else
    return 1.0f / x;
#endif
}
// The pair body's own forms.  sqrt(x) = r0 (1 + e/2 + 3 e^2/8) with r0 = x y, e = 1 - r0 y: the same cubically convergent
// correction as fast_rsqrt applied to r0 directly -- 5 FP64 instructions instead of 6 for x * fast_rsqrt(x), error ~2 ulp.
// 1/x by ONE Newton step (2 instructions instead of 3): quadratic from the ~2^-22 seed, relative error <= 2^-44 = 6e-14;
// it only enters the artificial-viscosity term Pi, a small part of the pair force, so the rates stay ~1e-14 from the oracle.
PST_HD double pair_sqrt(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double r0 = x * y;
    const double e = fma(-r0, y, 1.0);                      // 1 - x y^2
    return fma(r0, e * fma(0.375, e, 0.5), r0);
#else
    return sqrt(x);
#endif
}
PST_HD float pair_sqrt(float x) { return x * fast_rsqrt(x); }
PST_HD double pair_rcp(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return fma(y, fma(-x, y, 1.0), y);                      // y (1 + e),  e = 1 - x y
#else
    return 1.0 / x;
#endif
}
PST_HD float pair_rcp(float x) { return fast_rcp(x); }

// the pair body (continuity then momentum, the order fuse() keeps: fuse.rs:18,30).
// caller has already established 0 < r2 < rc2 with the exact test.  Branch-free: one rsqrt, one rcp.
//   BETA0: the quadratic viscosity coefficient is zero (every config of BASELINE.json): Pi = -2 alpha c0 h vx / ((r2 + eta2)(rho_i + rho_j))
//          costs 2 FP64 multiplies instead of 6 instructions.
//   FOLD (UNI kernels): the common mass is folded into the kernel-gradient constant (C.u_mgfc = m gfc); `in` masks the pair instead
//          of a zero mass.
template <class R, int DIM, bool CONT, bool MOM, bool UNI = false, bool BETA0 = false, bool FOLD = false>
PST_HD void pair_body(const WcsphConst<R>& C, const IState<R, DIM>& I, R dx, R dy, R dz, R r2, R uj, R vj, R wj,
                      R rhoj, R mj, R por2j, Acc<R>& a, bool in = true) {
    const R half_inv_h = UNI ? C.u_half_inv_h : I.half_inv_h, gfc = UNI ? C.u_gfc : I.gfc, eta2 = UNI ? C.u_eta2 : I.eta2, h = UNI ? C.u_h : I.h;
    const R r = pair_sqrt(r2);
    const R t = (R)1 - r * half_inv_h;
    const R du = I.u - uj, dv = I.v - vj, dw = DIM == 3 ? I.w - wj : (R)0;
    R vx = du * dx + dv * dy;
    if (DIM == 3) vx += dw * dz;
    R mgf;
    if (FOLD) { mgf = C.u_mgfc * (t * t * t); mgf = in ? mgf : (R)0; }
    else mgf = mj * (gfc * (t * t * t));
    if (CONT) a.arho += mgf * vx;
    if (MOM) {
        // Pi = (beta mu - alpha c0) mu / rho_bar,  mu = h vx / (r2 + eta2),  rho_bar = (rho_i + rho_j)/2
        const R rhos = I.rho + rhoj;
        const R inv = pair_rcp((r2 + eta2) * rhos);
        R Pi;
        if (BETA0) {
            const R k = UNI ? C.u_visc : (R)-2 * C.alpha_c0 * h;     // -2 alpha c0 h
            Pi = vx < (R)0 ? (k * vx) * inv : (R)0;
        } else {
            const R wv = h * vx * inv;                          // mu / (2 rho_bar)
            const R mu = wv * rhos;
            Pi = vx < (R)0 ? (C.beta * mu - C.alpha_c0) * (wv + wv) : (R)0;
        }
        const R c = -mgf * (I.por2 + por2j + Pi);
        a.au += c * dx;
        a.av += c * dy;
        if (DIM == 3) a.aw += c * dz;
    }
}

// Dummy-particle wall pressure: the five sums over the fluid neighbours of a non-fluid particle ...
template <class R>
struct WallSums { R S0, Sp, Sx, Sy, Sz; };

// ... one in-range fluid neighbour (0 < r2 < rc2 established by the caller): W = alpha_d (1 - q/2)^4 (2 q + 1), q = r / h_w
template <class R, int DIM>
PST_HD void wall_accumulate(WallSums<R>& S, R ad, R inv_h, R dx, R dy, R dz, R r2, R pj, R rhoj) {
    const R q = r2 * fast_rsqrt(r2) * inv_h;
    const R tt = (R)1 - (R)0.5 * q;
    const R t2 = tt * tt;
    const R W = ad * (t2 * t2) * ((R)2 * q + (R)1);
    S.S0 += W;
    S.Sp += pj * W;
    const R rW = rhoj * W;
    S.Sx += rW * dx; S.Sy += rW * dy;
    if (DIM == 3) S.Sz += rW * dz;
}

// ... and the result: p_w = (Sp + g . S) / S0 (0 without fluid neighbours), rho_w = rho0 (max(p_w / B, -1/2) + 1)^(1/gamma)
template <class R>
PST_HD void wall_finish(const WcsphConst<R>& C, const WallSums<R>& S, R& pw, R& rw) {
    pw = (R)0;
    if (S.S0 > (R)0) pw = (S.Sp + (C.g[0] * S.Sx + C.g[1] * S.Sy + C.g[2] * S.Sz)) / S.S0;
#if defined(__CUDA_ARCH__)
    const R e = max(pw / C.B, (R)-0.5);
#else
    const R e = pw / C.B > (R)-0.5 ? pw / C.B : (R)-0.5;
#endif
    rw = C.rho0 * exp(log1p(e) / C.gamma);
}
