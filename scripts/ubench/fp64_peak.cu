// Micro-benchmark: FP64 issue rate of one B200 ("FP64 FMA peak: not measured ... must be measured by the
// builder before quoting an FP64 fraction").  Every thread runs 8 independent DFMA (or DADD / DMUL) chains; with 8 resident
// warps per scheduler the pipe, not the dependency latency, is the limit.  Prints lane-operations per clock per SM and
// TFLOP/s (an FMA counted as 2 flop) from CUDA-event time, plus the SM clock implied by clock64().
// Result on this pool's B200 (profiles/ubench_fp64_peak.txt): DFMA 36.5 TFLOP/s, DADD and DMUL 18.5 Top/s each -- i.e.
// 62.8 lane-operations per clock per SM at the 1965 MHz maximum SM clock: the 64-lane FP64 pipe DESIGN.md assumes.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(256) k(double* out, double a, double b, int iters, long long* cycles) {
    double v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = a + (double)(threadIdx.x + c);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (OP == 0) v[c] = fma(v[c], a, b);
                else if (OP == 1) v[c] = __dadd_rn(v[c], b);
                else v[c] = __dmul_rn(v[c], a);
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += v[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char* name, int sms, double khz, double* out, long long* cyc) {
    const int iters = 4096, blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, 1.0000001, 1e-9, 64, cyc);       // warm-up
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * iters * 32.0;      // lane-operations
    const double per_clk_sm = ops / (ms * 1e-3) / (khz * 1e3) / sms;  // at the maximum SM clock (one block's clock64() span is not the kernel's)
    std::printf("%-5s %8.3f ms  %7.2f T%s/s  = %5.1f lane-ops/clk/SM at %.0f MHz\n", name, ms,
                ops * (OP == 0 ? 2.0 : 1.0) / (ms * 1e-3) / 1e12, OP == 0 ? "FLOP" : "OP", per_clk_sm, khz / 1e3);
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { std::printf("no CUDA device\n"); return 1; }
    double* out; long long* cyc;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    std::printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("DFMA", p.multiProcessorCount, (double)p.clockRate, out, cyc);
    run<1>("DADD", p.multiProcessorCount, (double)p.clockRate, out, cyc);
    run<2>("DMUL", p.multiProcessorCount, (double)p.clockRate, out, cyc);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
