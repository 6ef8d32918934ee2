// eq1.cu -- the reference's own sample equation, `force[i] += mass[j]`
// (/root/reference/prestige/src/lib.rs:7-12), executed with exactly the loop
// generate_simple_cpu emits for it (prestige/src/codegen/simple_cpu.rs:7-16):
// every i, every j in 0..n including j == i, sequential accumulation into force[i].
// Each thread owns one i and walks j in storage order through shared-memory tiles,
// so the sum order -- and therefore the bits -- equal the CPU loop's when the device
// order is the id order (no re-sort has happened yet).
#include "pst_internal.h"

namespace {
constexpr int kTile = 256;

template <class R>
__global__ void __launch_bounds__(kTile) k_eq1(int n, const R* __restrict__ mass, R* __restrict__ force) {
    __shared__ R tile[kTile];
    const int i = blockIdx.x * kTile + threadIdx.x;
    R acc = i < n ? force[i] : (R)0;
    for (int j0 = 0; j0 < n; j0 += kTile) {
        const int j = j0 + threadIdx.x;
        tile[threadIdx.x] = j < n ? mass[j] : (R)0;
        __syncthreads();
        const int m = min(kTile, n - j0);
        for (int t = 0; t < m; ++t) acc += tile[t];   // broadcast reads, strictly ascending j
        __syncthreads();
    }
    if (i < n) force[i] = acc;
}
}  // namespace

pst_status pst_eq1_apply(pst_ctx* ctx) {
    PstArray *f = pst_find(ctx, "force"), *m = pst_find(ctx, "mass");
    if (!f || !m) return pst_fail(ctx, PST_ESTATE, "eq1 reads 'mass' and writes 'force': create both arrays first");
    const int want = ctx->f64 ? PST_F64 : PST_F32;
    if (f->dtype != want || m->dtype != want) return pst_fail(ctx, PST_EINVAL, "eq1 arrays must have the context's real type");
    const int n = (int)ctx->n;
    if (n == 0) return PST_OK;
    const unsigned grid = (n + kTile - 1) / kTile;
    if (ctx->f64) PST_LAUNCH(ctx, k_eq1<double>, grid, kTile, 0, n, pst_ptr<double>(ctx, m), pst_ptr<double>(ctx, f));
    else PST_LAUNCH(ctx, k_eq1<float>, grid, kTile, 0, n, pst_ptr<float>(ctx, m), pst_ptr<float>(ctx, f));
    return PST_OK;
}
