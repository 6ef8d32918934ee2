// Micro-benchmark: L1 cost of gathering a 64-byte f64 neighbour record per lane, lanes of a warp reading ~9 groups of 8
// neighbouring particles (what the pair kernel's phase 2 does), for three record layouts:
//   A  AoSoA-8, 4 x LDG.128 (rows of 8 x 16 B, 128 B apart)           -- shipped in round 2
//   B  AoSoA-4, 2 x LDG.256 (rows of 4 x 32 B, 128 B apart)           -- sm_100a 256-bit loads
//   C  AoS, 2 x LDG.256 off one 64-byte record
//   D  AoS, 4 x LDG.128 off one 64-byte record
// Window of 1536 records (96 KB: L1-resident like a tile's candidates).  Prints cycles per gathered record-warp per SM
// with 16 resident warps.   Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_width gather_width.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

struct D4 { double a, b, c, d; };
__device__ __forceinline__ D4 ld256(const void* p) {
    D4 r;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(256, 2) k(const double* __restrict__ rec, const int* __restrict__ pattern, long long* out, double* sink, int iters) {
    const int lane = threadIdx.x & 31;
    const int off = pattern[(threadIdx.x >> 5) * 32 + lane];
    const double* base = rec + (size_t)blockIdx.x * 1536 * 8;
    double acc = 0;
    int c = (threadIdx.x >> 5) * 97;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        int j = c + off; if (j >= 1536) j -= 1536;
        if (MODE == 0) {
            const double2* q = reinterpret_cast<const double2*>(base) + (j + (j & ~7) * 3);
            const double2 a = q[0], b = q[8], cc = q[16], d = q[24];
            acc += a.x + a.y + b.x + b.y + cc.x + cc.y + d.x + d.y;
        } else if (MODE == 1) {
            const char* q = reinterpret_cast<const char*>(base) + ((size_t)(j + (j & ~3)) * 32);
            const D4 a = ld256(q), b = ld256(q + 128);
            acc += a.a + a.b + a.c + a.d + b.a + b.b + b.c + b.d;
        } else if (MODE == 2) {
            const char* q = reinterpret_cast<const char*>(base) + (size_t)j * 64;
            const D4 a = ld256(q), b = ld256(q + 32);
            acc += a.a + a.b + a.c + a.d + b.a + b.b + b.c + b.d;
        } else {
            const double2* q = reinterpret_cast<const double2*>(base) + j * 4;
            const double2 a = q[0], b = q[1], cc = q[2], d = q[3];
            acc += a.x + a.y + b.x + b.y + cc.x + cc.y + d.x + d.y;
        }
        c += 13; if (c >= 1536) c -= 1536;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == -1.0) *sink = acc;
}

template <int MODE>
void run(const char* name, int spread) {
    int h[256];
    srand(7);
    for (int l = 0; l < 256; ++l) h[l] = rand() % spread;
    int* d; long long* o; double *sink, *rec;
    const int nb = 296;
    cudaMalloc(&d, sizeof h); cudaMalloc(&o, 8 * nb); cudaMalloc(&sink, 8); cudaMalloc(&rec, (size_t)nb * 1536 * 64);
    cudaMemset(rec, 0, (size_t)nb * 1536 * 64);
    cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
    const int iters = 20000;
    k<MODE><<<nb, 256>>>(rec, d, o, sink, iters);
    k<MODE><<<nb, 256>>>(rec, d, o, sink, iters);
    long long ho[296];
    cudaMemcpy(ho, o, sizeof ho, cudaMemcpyDeviceToHost);
    double cyc = (double)ho[0] / (iters * 16.0);
    printf("%-28s spread %4d  %.2f cycles per record-warp per SM (%s)\n", name, spread, cyc, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d); cudaFree(o); cudaFree(sink); cudaFree(rec);
}

int main() {
    for (int spread : {8, 32, 72, 200, 1536}) {
        run<0>("A AoSoA-8 4xLDG.128", spread);
        run<1>("B AoSoA-4 2xLDG.256", spread);
        run<2>("C AoS 2xLDG.256", spread);
        run<3>("D AoS 4xLDG.128", spread);
    }
    return 0;
}
