// rigid.cu -- multi-particle rigid bodies in a coupled SPH-DEM context: reduction of the per-particle
// forces of the force loop to one force and torque per body, the body stage of the integrator, and the scatter of the
// rigid motion back onto the member particles.  No reference code exists for this physics; what the
// reference fixes is only that results are accumulated into named arrays by index (prestige/src/lib.rs:8-10).
// Formulation: DESIGN.md section 4c; per-body arithmetic: rigid_core.h (shared with the host-compiled test harness).
//
// Particles say which body they belong to through the persistent i32 array `body` (-1 = none: the particle is its own
// body, the behaviour of section 4b).  Members are tag-2 spheres; their contacts with members of the SAME body are
// skipped by the contact kernel (dem.cu), everything else in the force loop is unchanged.  Per-body records live in one
// device buffer, field-major and always in double (sums over many particles), whatever the context's `real`.
#include <algorithm>
#include <cstring>
#include <vector>

#include "pst_internal.h"
#include "rigid_core.h"

namespace {

inline unsigned blocks_for(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

// Body sums are DETERMINISTIC and atomic-free.  pst_bodies_setup groups the members by body once (stable sort of the
// device slots by body index) and gives every member its position `bpos` in that list -- a persistent particle array, so
// it follows the particle through every re-sort and into a checkpoint.  A sum over the members of each body is then
//   (1) one thread per particle writes its contribution to vals[k][bpos]   (scattered 8-byte stores, each slot once)
//   (2) one warp per body adds its contiguous range [start[b], start[b+1]) -- lanes stride the range, fixed butterfly --
//       and ASSIGNS the result, so the order of additions never depends on scheduling: two runs, or a run and its
//       resumed checkpoint, agree bit for bit, and large bodies do not serialise on one address.
template <int NV>
struct RbFields { int f[NV]; };

template <int NV>
__global__ void __launch_bounds__(256) k_rb_sum(int nb, const int32_t* __restrict__ start, const double* __restrict__ vals, size_t stride,
                                                double* __restrict__ B, RbFields<NV> F) {
    const int b = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (b >= nb) return;                                     // warp-uniform
    const int beg = start[b], end = start[b + 1];
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    for (int p = beg + lane; p < end; p += 32) {
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] += vals[(size_t)k * stride + p];
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if (lane == 0) B[(size_t)F.f[k] * nb + b] = acc[k];
    }
}

// ---- setup ---------------------------------------------------------------------------------------------------------
// members per body (integer atomics: the counts do not depend on the order) and the sort input (key = body, value = slot)
__global__ void __launch_bounds__(256) k_rb_count(int n, int nb, const int32_t* __restrict__ body, int32_t* __restrict__ count,
                                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, int32_t* __restrict__ flags) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int b = body[s];
    if (b >= nb) atomicExch(&flags[5], 1);                   // body index out of range: PST_EINVAL from pst_bodies_setup
    const bool member = b >= 0 && b < nb;
    if (member) atomicAdd(&count[b], 1);
    keys[s] = member ? (uint32_t)b : 0xffffffffu;            // non-members sort behind every body
    vals[s] = (uint32_t)s;
}

// sorted position p holds slot vals[p]: members come first, grouped by body, in slot order inside a body
__global__ void __launch_bounds__(256) k_rb_positions(int n, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                      int32_t* __restrict__ bpos) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) bpos[vals[p]] = keys[p] == 0xffffffffu ? -1 : p;
}

// contributions to M, sum m x, sum m v
template <class R>
__global__ void __launch_bounds__(256) k_rb_setup_vals(int n, const int32_t* __restrict__ bpos, const R* __restrict__ m, const R* __restrict__ x,
                                                       const R* __restrict__ y, const R* __restrict__ z, const R* __restrict__ u,
                                                       const R* __restrict__ v, const R* __restrict__ w, double* __restrict__ vals, size_t stride) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int p = bpos[s];
    if (p < 0) return;
    const double mi = (double)m[s];
    vals[p] = mi;
    vals[stride + p] = mi * (double)x[s]; vals[2 * stride + p] = mi * (double)y[s]; vals[3 * stride + p] = mi * (double)z[s];
    vals[4 * stride + p] = mi * (double)u[s]; vals[5 * stride + p] = mi * (double)v[s]; vals[6 * stride + p] = mi * (double)w[s];
}

// one thread per body: X = SX / M, V = SV / M, w = 0, R = identity, F = T = 0
__global__ void __launch_bounds__(256) k_rb_setup_cm(int nb, double* __restrict__ B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const double M = B[(size_t)RB_M * nb + b];
    const double inv = M > 0.0 ? 1.0 / M : 0.0;             // a body without members keeps zeros
    for (int k = 0; k < 3; ++k) {
        B[(size_t)(RB_X + k) * nb + b] = B[(size_t)(RB_SX + k) * nb + b] * inv;
        B[(size_t)(RB_V + k) * nb + b] = B[(size_t)(RB_SV + k) * nb + b] * inv;
        B[(size_t)(RB_W + k) * nb + b] = 0.0;
        B[(size_t)(RB_F + k) * nb + b] = 0.0;
        B[(size_t)(RB_T + k) * nb + b] = 0.0;
    }
    for (int k = 0; k < 9; ++k) B[(size_t)(RB_R + k) * nb + b] = (k == 0 || k == 4 || k == 8) ? 1.0 : 0.0;
}

// body-frame offsets r0 = x - X (stored in the context's real) and the contributions to
// I0 = sum m (|r0|^2 1 - r0 r0^T) + inertia_i 1
template <class R>
__global__ void __launch_bounds__(256) k_rb_setup_members(int n, int nb, const int32_t* __restrict__ body, const int32_t* __restrict__ bpos,
                                                          const R* __restrict__ m, const R* __restrict__ inertia, const R* __restrict__ x,
                                                          const R* __restrict__ y, const R* __restrict__ z, R* __restrict__ bx0,
                                                          R* __restrict__ by0, R* __restrict__ bz0, const double* __restrict__ B,
                                                          double* __restrict__ vals, size_t stride) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int p = bpos[s];
    if (p < 0) { bx0[s] = by0[s] = bz0[s] = (R)0; return; }
    const int b = body[s];
    const R rx = (R)((double)x[s] - B[(size_t)(RB_X + 0) * nb + b]);
    const R ry = (R)((double)y[s] - B[(size_t)(RB_X + 1) * nb + b]);
    const R rz = (R)((double)z[s] - B[(size_t)(RB_X + 2) * nb + b]);
    bx0[s] = rx; by0[s] = ry; bz0[s] = rz;
    const double ax = (double)rx, ay = (double)ry, az = (double)rz, mi = (double)m[s], is = (double)inertia[s];
    vals[p] = mi * (ay * ay + az * az) + is;
    vals[stride + p] = mi * (ax * ax + az * az) + is;
    vals[2 * stride + p] = mi * (ax * ax + ay * ay) + is;
    vals[3 * stride + p] = -mi * ax * ay;
    vals[4 * stride + p] = -mi * ax * az;
    vals[5 * stride + p] = -mi * ay * az;
}

// ---- reduction of the force loop's per-particle results --------------------------------------------------------------
template <class R>
struct RbReduceArgs {
    const int32_t *body, *bpos;
    const R *x, *y, *z, *m, *fx, *fy, *fz, *tx, *ty, *tz, *au, *av, *aw;
    double ratio, g[3];
    int n, nb;
};

// contributions to F_b = sum_i Ft_i and T_b = sum_i (x_i - X_b) x Ft_i + t_i, Ft_i = f_i + m_i rho0/rho_s (a_i - g) + m_i g
template <class R>
__global__ void __launch_bounds__(256) k_rb_reduce_vals(RbReduceArgs<R> A, const double* __restrict__ B, double* __restrict__ vals, size_t stride) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.n) return;
    const int p = A.bpos[s];
    if (p < 0) return;
    const int b = A.body[s];
    const double f[3] = {(double)A.fx[s], (double)A.fy[s], (double)A.fz[s]};
    const double a[3] = {(double)A.au[s], (double)A.av[s], (double)A.aw[s]};
    double F[3], r[3], rxF[3];
    rb_particle_force((double)A.m[s], A.ratio, f, a, A.g, F);
    r[0] = (double)A.x[s] - B[(size_t)(RB_X + 0) * A.nb + b];
    r[1] = (double)A.y[s] - B[(size_t)(RB_X + 1) * A.nb + b];
    r[2] = (double)A.z[s] - B[(size_t)(RB_X + 2) * A.nb + b];
    rb_cross(r, F, rxF);
    vals[p] = F[0]; vals[stride + p] = F[1]; vals[2 * stride + p] = F[2];
    vals[3 * stride + p] = rxF[0] + (double)A.tx[s];
    vals[4 * stride + p] = rxF[1] + (double)A.ty[s];
    vals[5 * stride + p] = rxF[2] + (double)A.tz[s];
}

// ---- integrator -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rb_load(const double* __restrict__ B, int nb, int b, RbState& st) {
    st.M = B[(size_t)RB_M * nb + b];
    for (int k = 0; k < 3; ++k) {
        st.X[k] = B[(size_t)(RB_X + k) * nb + b]; st.V[k] = B[(size_t)(RB_V + k) * nb + b]; st.W[k] = B[(size_t)(RB_W + k) * nb + b];
        st.F[k] = B[(size_t)(RB_F + k) * nb + b]; st.T[k] = B[(size_t)(RB_T + k) * nb + b];
    }
    for (int k = 0; k < 9; ++k) st.R[k] = B[(size_t)(RB_R + k) * nb + b];
    for (int k = 0; k < 6; ++k) st.I0[k] = B[(size_t)(RB_I0 + k) * nb + b];
}

__global__ void __launch_bounds__(128) k_rb_integrate(int nb, double dt, double* __restrict__ B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    RbState st;
    rb_load(B, nb, b, st);
    if (!(st.M > 0.0)) return;
    rb_integrate(st, dt);
    for (int k = 0; k < 3; ++k) {
        B[(size_t)(RB_X + k) * nb + b] = st.X[k]; B[(size_t)(RB_V + k) * nb + b] = st.V[k]; B[(size_t)(RB_W + k) * nb + b] = st.W[k];
    }
    for (int k = 0; k < 9; ++k) B[(size_t)(RB_R + k) * nb + b] = st.R[k];
}

// members take position, velocity and spin from their body: x = X + R r0, v = V + w x (x - X), omega = w
template <class R>
__global__ void __launch_bounds__(256) k_rb_scatter(int n, int nb, const int32_t* __restrict__ body, const R* __restrict__ bx0,
                                                    const R* __restrict__ by0, const R* __restrict__ bz0, R* __restrict__ x, R* __restrict__ y,
                                                    R* __restrict__ z, R* __restrict__ u, R* __restrict__ v, R* __restrict__ w,
                                                    R* __restrict__ wx, R* __restrict__ wy, R* __restrict__ wz, const double* __restrict__ B) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int b = body[s];
    if (b < 0 || b >= nb) return;
    RbState st;
    rb_load(B, nb, b, st);
    const double r0[3] = {(double)bx0[s], (double)by0[s], (double)bz0[s]};
    double xo[3], vo[3];
    rb_member(st, r0, xo, vo);
    x[s] = (R)xo[0]; y[s] = (R)xo[1]; z[s] = (R)xo[2];
    u[s] = (R)vo[0]; v[s] = (R)vo[1]; w[s] = (R)vo[2];
    wx[s] = (R)st.W[0]; wy[s] = (R)st.W[1]; wz[s] = (R)st.W[2];
}

template <class R>
pst_status launch_scatter(pst_ctx* ctx) {
    const int n = (int)ctx->n;
    if (n == 0) return PST_OK;
    PST_LAUNCH(ctx, k_rb_scatter<R>, blocks_for(n, 256), 256, 0, n, (int)ctx->n_bodies, pst_ptr<int32_t>(ctx, "body"), pst_ptr<R>(ctx, "bx0"),
               pst_ptr<R>(ctx, "by0"), pst_ptr<R>(ctx, "bz0"), pst_ptr<R>(ctx, "x"), pst_ptr<R>(ctx, "y"), pst_ptr<R>(ctx, "z"),
               pst_ptr<R>(ctx, "u"), pst_ptr<R>(ctx, "v"), pst_ptr<R>(ctx, "w"), pst_ptr<R>(ctx, "wx"), pst_ptr<R>(ctx, "wy"),
               pst_ptr<R>(ctx, "wz"), ctx->d_bodies);
    ctx->nbrs_valid = false;
    ctx->state_epoch++;
    return PST_OK;
}

template <int NV>
pst_status launch_sum(pst_ctx* ctx, const RbFields<NV>& F) {
    const int nb = (int)ctx->n_bodies;
    PST_LAUNCH(ctx, k_rb_sum<NV>, blocks_for((size_t)nb * 32, 256), 256, 0, nb, ctx->d_body_start, ctx->d_body_vals, ctx->body_members, ctx->d_bodies, F);
    return PST_OK;
}

// member lists: counts -> exclusive scan = start[]; with `assign`, a stable sort of the slots by body gives every member
// its position bpos (without: the caller restored bpos from a checkpoint; the ranges only depend on the counts)
pst_status build_member_lists(pst_ctx* ctx, bool assign) {
    const int n = (int)ctx->n, nb = (int)ctx->n_bodies;
    PST_TRY(pst_resolve_history(ctx));                       // the sort below reuses vals_out, which a deferred history remap still needs
    PST_CUDA(ctx, cudaMemsetAsync(ctx->d_body_start, 0, ((size_t)nb + 1) * 4, ctx->stream));
    if (n > 0) {
        PST_LAUNCH(ctx, k_rb_count, blocks_for(n, 256), 256, 0, n, nb, pst_ptr<int32_t>(ctx, "body"), ctx->d_body_start, ctx->keys_in,
                   ctx->vals_in, ctx->d_flags);
    }
    PST_TRY(pst_scan_exclusive(ctx, ctx->d_body_start, nb + 1));
    if (assign) PST_TRY(pst_sort_pairs_u32(ctx, n));
    if (assign && n > 0) PST_LAUNCH(ctx, k_rb_positions, blocks_for(n, 256), 256, 0, n, ctx->keys_out, ctx->vals_out, pst_ptr<int32_t>(ctx, "bpos"));
    PST_CUDA(ctx, cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaMemcpyAsync(ctx->h_flags + 6, ctx->d_body_start + nb, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_flags[5]) {
        PST_CUDA(ctx, cudaMemsetAsync(ctx->d_flags + 5, 0, sizeof(int32_t), ctx->stream));
        return pst_fail(ctx, PST_EINVAL, "array 'body' holds an index >= n_bodies = %u", ctx->n_bodies);
    }
    const size_t members = (size_t)ctx->h_flags[6];
    if (members > ctx->body_members || !ctx->d_body_vals) {
        PST_CUDA(ctx, cudaFree(ctx->d_body_vals));
        ctx->d_body_vals = nullptr;
        if (cudaMalloc((void**)&ctx->d_body_vals, 7 * std::max<size_t>(members, 1) * sizeof(double)) != cudaSuccess)
            return pst_fail(ctx, PST_ENOMEM, "rigid-body scratch (%zu members)", members);
    }
    ctx->body_members = std::max<size_t>(members, 1);        // also the stride of d_body_vals
    return PST_OK;
}

template <class R>
pst_status launch_setup(pst_ctx* ctx) {
    const int n = (int)ctx->n, nb = (int)ctx->n_bodies;
    PST_CUDA(ctx, cudaMemsetAsync(ctx->d_bodies, 0, (size_t)RB_NF * nb * sizeof(double), ctx->stream));
    const int32_t* bpos = pst_ptr<int32_t>(ctx, "bpos");
    if (n > 0) {
        PST_LAUNCH(ctx, k_rb_setup_vals<R>, blocks_for(n, 256), 256, 0, n, bpos, pst_ptr<R>(ctx, "m"), pst_ptr<R>(ctx, "x"), pst_ptr<R>(ctx, "y"),
                   pst_ptr<R>(ctx, "z"), pst_ptr<R>(ctx, "u"), pst_ptr<R>(ctx, "v"), pst_ptr<R>(ctx, "w"), ctx->d_body_vals, ctx->body_members);
    }
    PST_TRY(launch_sum<7>(ctx, RbFields<7>{{RB_M, RB_SX, RB_SX + 1, RB_SX + 2, RB_SV, RB_SV + 1, RB_SV + 2}}));
    PST_LAUNCH(ctx, k_rb_setup_cm, blocks_for(nb, 256), 256, 0, nb, ctx->d_bodies);
    if (n > 0) {
        PST_LAUNCH(ctx, k_rb_setup_members<R>, blocks_for(n, 256), 256, 0, n, nb, pst_ptr<int32_t>(ctx, "body"), bpos, pst_ptr<R>(ctx, "m"),
                   pst_ptr<R>(ctx, "inertia"), pst_ptr<R>(ctx, "x"), pst_ptr<R>(ctx, "y"), pst_ptr<R>(ctx, "z"), pst_ptr<R>(ctx, "bx0"),
                   pst_ptr<R>(ctx, "by0"), pst_ptr<R>(ctx, "bz0"), ctx->d_bodies, ctx->d_body_vals, ctx->body_members);
    }
    return launch_sum<6>(ctx, RbFields<6>{{RB_I0, RB_I0 + 1, RB_I0 + 2, RB_I0 + 3, RB_I0 + 4, RB_I0 + 5}});
}

template <class R>
pst_status launch_reduce(pst_ctx* ctx) {
    RbReduceArgs<R> A;
    A.body = pst_ptr<int32_t>(ctx, "body"); A.bpos = pst_ptr<int32_t>(ctx, "bpos");
    A.x = pst_ptr<R>(ctx, "x"); A.y = pst_ptr<R>(ctx, "y"); A.z = pst_ptr<R>(ctx, "z"); A.m = pst_ptr<R>(ctx, "m");
    A.fx = pst_ptr<R>(ctx, "fx"); A.fy = pst_ptr<R>(ctx, "fy"); A.fz = pst_ptr<R>(ctx, "fz");
    A.tx = pst_ptr<R>(ctx, "tx"); A.ty = pst_ptr<R>(ctx, "ty"); A.tz = pst_ptr<R>(ctx, "tz");
    A.au = pst_ptr<R>(ctx, "au"); A.av = pst_ptr<R>(ctx, "av"); A.aw = pst_ptr<R>(ctx, "aw");
    // the same rounding of the constants as k_coupled_integrate: formed in the context's real
    A.ratio = (double)((R)pst_param(ctx, "rho0") / (R)pst_param(ctx, "rho_solid", 1.0));
    A.g[0] = (double)(R)pst_param(ctx, "gx"); A.g[1] = (double)(R)pst_param(ctx, "gy"); A.g[2] = (double)(R)pst_param(ctx, "gz");
    A.n = (int)ctx->n; A.nb = (int)ctx->n_bodies;
    if (A.n > 0) PST_LAUNCH(ctx, k_rb_reduce_vals<R>, blocks_for(A.n, 256), 256, 0, A, ctx->d_bodies, ctx->d_body_vals, ctx->body_members);
    return launch_sum<6>(ctx, RbFields<6>{{RB_F, RB_F + 1, RB_F + 2, RB_T, RB_T + 1, RB_T + 2}});
}

struct RbName { const char* name; int field, width; bool writable; };
// mass and inertia0 are normally computed by pst_bodies_setup; they are writable so that a checkpoint can be restored
const RbName kRbNames[] = {{"mass", RB_M, 1, true}, {"cm", RB_X, 3, true},        {"vel", RB_V, 3, true},     {"omega", RB_W, 3, true},
                           {"rot", RB_R, 9, true},  {"inertia0", RB_I0, 6, true}, {"force", RB_F, 3, false}, {"torque", RB_T, 3, false}};

}  // namespace

pst_status pst_rb_reduce(pst_ctx* ctx) {
    if (!ctx->bodies_ready) return pst_fail(ctx, PST_ESTATE, "body_reduce: pst_bodies_setup has not run for this particle set");
    return ctx->f64 ? launch_reduce<double>(ctx) : launch_reduce<float>(ctx);
}

pst_status pst_rb_integrate(pst_ctx* ctx, double dt) {
    const int nb = (int)ctx->n_bodies;
    PST_LAUNCH(ctx, k_rb_integrate, blocks_for(nb, 128), 128, 0, nb, dt, ctx->d_bodies);
    return ctx->f64 ? launch_scatter<double>(ctx) : launch_scatter<float>(ctx);
}

extern "C" {

// Register `n_bodies` rigid bodies.  Creates the persistent particle arrays `body` (i32, -1 everywhere), the body-frame
// offsets `bx0 by0 bz0` and `bpos` (i32, position in the member list).  Coupled contexts on one GPU only.
pst_status pst_bodies_create(pst_ctx* ctx, uint32_t n_bodies) {
    if (!ctx || n_bodies == 0) return PST_EINVAL;
    if (!ctx->coupled) return pst_fail(ctx, PST_ESTATE, "rigid bodies need physics = PST_PHYS_WCSPH | PST_PHYS_DEM");
    if (ctx->comm) return pst_fail(ctx, PST_ESTATE, "rigid bodies are not supported with a communicator attached");
    if (ctx->d_bodies) return pst_fail(ctx, PST_ESTATE, "bodies already created");
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    PST_TRY(pst_array_create(ctx, "body", PST_I32, PST_ARRAY_PERSISTENT));
    PST_TRY(pst_array_create(ctx, "bx0", PST_REAL, PST_ARRAY_PERSISTENT));
    PST_TRY(pst_array_create(ctx, "by0", PST_REAL, PST_ARRAY_PERSISTENT));
    PST_TRY(pst_array_create(ctx, "bz0", PST_REAL, PST_ARRAY_PERSISTENT));
    PST_TRY(pst_array_create(ctx, "bpos", PST_I32, PST_ARRAY_PERSISTENT));   // position in the body-grouped member list
    for (const char* nm : {"body", "bpos"}) {
        PstArray* b = pst_find(ctx, nm);
        const size_t bytes = (ctx->capacity + 2 * ctx->ghost_cap) * b->esize;
        for (int k = 0; k < 2; ++k) PST_CUDA(ctx, cudaMemsetAsync(b->buf[k], 0xFF, bytes, ctx->stream));   // -1: no body
    }
    if (cudaMalloc((void**)&ctx->d_bodies, (size_t)RB_NF * n_bodies * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_body_start, ((size_t)n_bodies + 1) * sizeof(int32_t)) != cudaSuccess)
        return pst_fail(ctx, PST_ENOMEM, "body records (%u bodies)", n_bodies);
    PST_CUDA(ctx, cudaMemsetAsync(ctx->d_bodies, 0, (size_t)RB_NF * n_bodies * sizeof(double), ctx->stream));
    ctx->n_bodies = n_bodies;
    ctx->bodies_ready = false;
    return PST_OK;
}

// After `body`, x y z, u v w, m and inertia are uploaded: mass, centre of mass, mass-weighted velocity, body-frame
// offsets and inertia tensor of every body; orientation = identity, omega = 0.  Members then take the rigid motion
// (their individual u v w and spins are overwritten).
pst_status pst_bodies_setup(pst_ctx* ctx) {
    if (!ctx) return PST_EINVAL;
    if (!ctx->d_bodies) return pst_fail(ctx, PST_ESTATE, "pst_bodies_create first");
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    ctx->bodies_ready = false;
    PST_TRY(build_member_lists(ctx, true));
    PST_TRY(ctx->f64 ? launch_setup<double>(ctx) : launch_setup<float>(ctx));
    PST_TRY(ctx->f64 ? launch_scatter<double>(ctx) : launch_scatter<float>(ctx));
    ctx->bodies_ready = true;
    ctx->eos_valid = false; ctx->state_epoch++;
    return PST_OK;
}

// Restoring a checkpoint: `body`, `bpos` and `bx0 by0 bz0` have been uploaded as saved; rebuild the member-list ranges
// from `body` (they only depend on the member counts), keep the saved positions -- and with them the order of every
// body sum, so the resumed run is bit-identical -- and leave the records to pst_bodies_state writes.
pst_status pst_bodies_restore(pst_ctx* ctx) {
    if (!ctx) return PST_EINVAL;
    if (!ctx->d_bodies) return pst_fail(ctx, PST_ESTATE, "pst_bodies_create first");
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    ctx->bodies_ready = false;
    PST_TRY(build_member_lists(ctx, false));
    ctx->bodies_ready = true;
    ctx->eos_valid = false; ctx->state_epoch++;
    return PST_OK;
}

// Read (write = 0) or write (write = 1) one per-body quantity; host data is always double, body-major:
// "mass" [nb], "cm" "vel" "omega" "force" "torque" [nb][3], "rot" [nb][9] row-major, "inertia0" [nb][6] = xx yy zz xy xz yz.
// Writing cm / vel / omega / rot moves the member particles accordingly.
pst_status pst_bodies_state(pst_ctx* ctx, const char* name, double* host, size_t n, int write) {
    if (!ctx || !name || !host) return PST_EINVAL;
    if (!ctx->d_bodies) return pst_fail(ctx, PST_ESTATE, "pst_bodies_create first");
    const RbName* f = nullptr;
    for (const RbName& r : kRbNames)
        if (std::strcmp(r.name, name) == 0) f = &r;
    if (!f) return pst_fail(ctx, PST_EINVAL, "unknown body quantity '%s'", name);
    const size_t nb = ctx->n_bodies;
    if (n != nb * f->width) return pst_fail(ctx, PST_EINVAL, "body quantity '%s' has %zu values, not %zu", name, nb * f->width, n);
    if (write && !f->writable) return pst_fail(ctx, PST_EINVAL, "body quantity '%s' is read-only", name);
    if (write && !ctx->bodies_ready) return pst_fail(ctx, PST_ESTATE, "pst_bodies_setup first");
    PST_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    std::vector<double> tmp(n);                    // device layout is field-major: transpose on the host (small)
    double* dev = ctx->d_bodies + (size_t)f->field * nb;
    if (write) {
        for (size_t b = 0; b < nb; ++b)
            for (int k = 0; k < f->width; ++k) tmp[(size_t)k * nb + b] = host[b * f->width + k];
        PST_CUDA(ctx, cudaMemcpyAsync(dev, tmp.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // tmp goes out of scope
        return ctx->f64 ? launch_scatter<double>(ctx) : launch_scatter<float>(ctx);
    }
    PST_CUDA(ctx, cudaMemcpyAsync(tmp.data(), dev, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PST_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t b = 0; b < nb; ++b)
        for (int k = 0; k < f->width; ++k) host[b * f->width + k] = tmp[(size_t)k * nb + b];
    return PST_OK;
}

}  // extern "C"
