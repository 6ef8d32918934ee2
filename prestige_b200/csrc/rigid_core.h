// rigid_core.h -- per-body arithmetic of the multi-particle rigid bodies ("rigid-body reduction of
// per-particle forces to body force/torque", the stage either side of the coupled force loop).  No reference code
// exists for this physics; the formulation is DESIGN.md section 4c.
//
// Plain C++ in double, usable from device code (rigid.cu) and from a host-compiled test harness
// (tests/cpp/rigid_core_harness.cpp), so the arithmetic can be checked against the numpy oracle without a GPU.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define PST_HD __host__ __device__ __forceinline__
#else
#define PST_HD inline
#endif

// per-body record, field-major on the device: field f of body b sits at base[f * nb + b]
enum RbField {
    RB_M = 0,         // mass
    RB_X = 1,         // centre of mass (3)
    RB_V = 4,         // velocity (3)
    RB_W = 7,         // angular velocity (3)
    RB_R = 10,        // rotation matrix, row-major (9); identity at setup
    RB_I0 = 19,       // body-frame inertia tensor about the centre of mass: xx yy zz xy xz yz (6)
    RB_F = 25,        // force (3)        -- written by body_reduce
    RB_T = 28,        // torque about the centre of mass (3)
    RB_SX = 31,       // setup scratch: sum m x (3)
    RB_SV = 34,       // setup scratch: sum m v (3)
    RB_NF = 37
};

struct RbState {
    double M, X[3], V[3], W[3], R[9], I0[6], F[3], T[3];
};

// total force on a member particle of a coupled context: contact force + hydrodynamic force m rho0/rho_s (a - g) +
// weight m g -- the same expression k_coupled_integrate uses for a single sphere, times m.
PST_HD void rb_particle_force(double m, double ratio, const double f[3], const double a[3], const double g[3], double out[3]) {
    for (int k = 0; k < 3; ++k) out[k] = m * ((f[k] / m + ratio * (a[k] - g[k])) + g[k]);
}

PST_HD void rb_cross(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// world inertia I = R I0 R^T (symmetric 3x3, returned as full row-major matrix)
PST_HD void rb_world_inertia(const double R[9], const double I0[6], double I[9]) {
    const double A[9] = {I0[0], I0[3], I0[4], I0[3], I0[1], I0[5], I0[4], I0[5], I0[2]};
    double RA[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) RA[3 * r + c] = R[3 * r] * A[c] + R[3 * r + 1] * A[3 + c] + R[3 * r + 2] * A[6 + c];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) I[3 * r + c] = RA[3 * r] * R[3 * c] + RA[3 * r + 1] * R[3 * c + 1] + RA[3 * r + 2] * R[3 * c + 2];
}

// y = A^-1 b for a 3x3 matrix by the adjugate (the inertia tensor is symmetric positive definite)
PST_HD void rb_solve3(const double A[9], const double b[3], double y[3]) {
    const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
    const double inv = 1.0 / det;
    const double c10 = A[2] * A[7] - A[1] * A[8], c11 = A[0] * A[8] - A[2] * A[6], c12 = A[1] * A[6] - A[0] * A[7];
    const double c20 = A[1] * A[5] - A[2] * A[4], c21 = A[2] * A[3] - A[0] * A[5], c22 = A[0] * A[4] - A[1] * A[3];
    y[0] = (c00 * b[0] + c10 * b[1] + c20 * b[2]) * inv;
    y[1] = (c01 * b[0] + c11 * b[1] + c21 * b[2]) * inv;
    y[2] = (c02 * b[0] + c12 * b[1] + c22 * b[2]) * inv;
}

// Semi-implicit Euler stage of one body (same scheme as the particle integrators):
//   V += F/M dt;  X += V dt;  w += I^-1 (T - w x (I w)) dt with I = R I0 R^T at the current orientation;
//   R <- exp([w dt]x) R  (Rodrigues; first-order form for a vanishing angle).
PST_HD void rb_integrate(RbState& b, double dt) {
    for (int k = 0; k < 3; ++k) {
        b.V[k] += b.F[k] / b.M * dt;
        b.X[k] += b.V[k] * dt;
    }
    double I[9], Iw[3], wxIw[3], rhs[3], dw[3];
    rb_world_inertia(b.R, b.I0, I);
    for (int r = 0; r < 3; ++r) Iw[r] = I[3 * r] * b.W[0] + I[3 * r + 1] * b.W[1] + I[3 * r + 2] * b.W[2];
    rb_cross(b.W, Iw, wxIw);
    for (int k = 0; k < 3; ++k) rhs[k] = b.T[k] - wxIw[k];
    rb_solve3(I, rhs, dw);
    for (int k = 0; k < 3; ++k) b.W[k] += dw[k] * dt;
    const double ax = b.W[0] * dt, ay = b.W[1] * dt, az = b.W[2] * dt;
    const double th2 = ax * ax + ay * ay + az * az;
    double s, c;   // E = 1 + s [a]x + c [a]x^2 with a = w dt:  s = sin(th)/th, c = (1 - cos(th))/th^2
    if (th2 < 1e-24) { s = 1.0; c = 0.5; }
    else { const double th = sqrt(th2); s = sin(th) / th; const double sh = sin(0.5 * th); c = 2.0 * sh * sh / th2; }
    const double K[9] = {0, -az, ay, az, 0, -ax, -ay, ax, 0};
    double K2[9], E[9], Rn[9];
    for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 3; ++q) K2[3 * r + q] = K[3 * r] * K[q] + K[3 * r + 1] * K[3 + q] + K[3 * r + 2] * K[6 + q];
    for (int k = 0; k < 9; ++k) E[k] = s * K[k] + c * K2[k];
    E[0] += 1.0; E[4] += 1.0; E[8] += 1.0;
    for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 3; ++q) Rn[3 * r + q] = E[3 * r] * b.R[q] + E[3 * r + 1] * b.R[3 + q] + E[3 * r + 2] * b.R[6 + q];
    for (int k = 0; k < 9; ++k) b.R[k] = Rn[k];
}

// member particle from the body state: x = X + R r0, v = V + w x (x - X); its spin is the body's
PST_HD void rb_member(const RbState& b, const double r0[3], double x[3], double v[3]) {
    double r[3], wxr[3];
    for (int k = 0; k < 3; ++k) r[k] = b.R[3 * k] * r0[0] + b.R[3 * k + 1] * r0[1] + b.R[3 * k + 2] * r0[2];
    rb_cross(b.W, r, wxr);
    for (int k = 0; k < 3; ++k) { x[k] = b.X[k] + r[k]; v[k] = b.V[k] + wxr[k]; }
}
