// Micro-benchmark: cost of LDS.32 / LDS.64 / LDS.128 when the lanes of a warp read (a) one common address,
// (b) three distinct addresses (a warp spanning three cells), (c) 32 distinct addresses.  Prints cycles per
// warp-instruction per SM with 8 resident warps (enough to saturate the shared-memory pipe).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lds_broadcast lds_broadcast.cu
#include <cstdio>
#include <cuda_runtime.h>

template <class V>
__global__ void k(const int* __restrict__ pattern, long long* out, float* sink, int iters) {
    extern __shared__ __align__(16) unsigned char smem[];
    V* s = reinterpret_cast<V*>(smem);
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)i;
    __syncthreads();
    int idx = pattern[threadIdx.x & 31];
    float acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            V v = s[(idx + u * 3) & 127];
            acc += *reinterpret_cast<float*>(&v);
        }
        idx = (idx + 1) & 127;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == -1.f) *sink = acc;
}

template <class V>
void run(const char* name, int mode) {
    int h[32];
    for (int l = 0; l < 32; ++l) h[l] = mode == 0 ? 5 : mode == 1 ? (l < 11 ? 5 : l < 22 ? 46 : 87) : l * 3 + 1;
    int* d; long long* o; float* sink;
    cudaMalloc(&d, sizeof h); cudaMalloc(&o, 8 * 148); cudaMalloc(&sink, 4);
    cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
    const int iters = 2000, warps = 8;
    k<V><<<148, warps * 32, 8192>>>(d, o, sink, iters);
    k<V><<<148, warps * 32, 8192>>>(d, o, sink, iters);
    long long ho[148];
    cudaMemcpy(ho, o, sizeof ho, cudaMemcpyDeviceToHost);
    double cyc = (double)ho[0] / (iters * 16.0 * warps);
    printf("%-8s %-22s %.2f cycles per warp-instruction per SM\n", name, mode == 0 ? "one address" : mode == 1 ? "three addresses" : "32 distinct addresses", cyc);
    cudaFree(d); cudaFree(o); cudaFree(sink);
}

int main() {
    for (int mode = 0; mode < 3; ++mode) {
        run<float>("LDS.32", mode);
        run<float2>("LDS.64", mode);
        run<float4>("LDS.128", mode);
    }
    return 0;
}
