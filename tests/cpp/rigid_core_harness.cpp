// Host-compiled view of prestige_b200/csrc/rigid_core.h: the SAME per-body arithmetic the device kernels of rigid.cu
// run, exported with C linkage so tests/test_rigid.py can compare it against the numpy oracle without a GPU.
// Test infrastructure only.
#include "../../prestige_b200/csrc/rigid_core.h"

namespace {
void unpack(const double* s, RbState& b) {
    b.M = s[RB_M];
    for (int k = 0; k < 3; ++k) { b.X[k] = s[RB_X + k]; b.V[k] = s[RB_V + k]; b.W[k] = s[RB_W + k]; b.F[k] = s[RB_F + k]; b.T[k] = s[RB_T + k]; }
    for (int k = 0; k < 9; ++k) b.R[k] = s[RB_R + k];
    for (int k = 0; k < 6; ++k) b.I0[k] = s[RB_I0 + k];
}
void pack(const RbState& b, double* s) {
    s[RB_M] = b.M;
    for (int k = 0; k < 3; ++k) { s[RB_X + k] = b.X[k]; s[RB_V + k] = b.V[k]; s[RB_W + k] = b.W[k]; s[RB_F + k] = b.F[k]; s[RB_T + k] = b.T[k]; }
    for (int k = 0; k < 9; ++k) s[RB_R + k] = b.R[k];
    for (int k = 0; k < 6; ++k) s[RB_I0 + k] = b.I0[k];
}
}  // namespace

extern "C" {
int rbh_fields() { return RB_NF; }
// field offsets in the order M X V W R I0 F T (so the test does not hard-code the enum)
void rbh_layout(int* out) { const int f[8] = {RB_M, RB_X, RB_V, RB_W, RB_R, RB_I0, RB_F, RB_T}; for (int k = 0; k < 8; ++k) out[k] = f[k]; }
void rbh_particle_force(double m, double ratio, const double* f, const double* a, const double* g, double* out) { rb_particle_force(m, ratio, f, a, g, out); }
void rbh_integrate(double* state, double dt) { RbState b; unpack(state, b); rb_integrate(b, dt); pack(b, state); }
void rbh_member(const double* state, const double* r0, double* x, double* v) { RbState b; unpack(state, b); rb_member(b, r0, x, v); }
}
