// Micro-benchmark: can the 64-byte neighbour-record gathers of the pair kernel's phase 2 run through the TEXTURE path
// (tex1Dfetch of 16-byte texels) instead of the LSU path, and do the two paths overlap?  Same record layout and index
// pattern as gather_width.cu mode A (AoSoA-8, 4 x 16 B per lane).  Variants: LDG only, TEX only, each with and without a
// concurrent stream of shared-memory loads (what phase 1 of other warps does), and half the warps on each path.
// Prints cycles per gathered record-warp per SM (16 resident warps).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_tex gather_tex.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int MODE, bool WITH_LDS>
__global__ void __launch_bounds__(256, 2) k(const double2* __restrict__ rec, cudaTextureObject_t tex, const int* __restrict__ pattern, long long* out,
                                            double* sink, int iters) {
    __shared__ float4 sm[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sm[i] = make_float4(i, i, i, i);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int off = pattern[warp * 32 + lane];
    const int base = blockIdx.x * 1536;           // record window of this block
    double acc = 0;
    float facc = 0;
    int c = warp * 97;
    const bool use_tex = MODE == 1 || (MODE == 2 && (warp & 1));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        int j = c + off; if (j >= 1536) j -= 1536;
        const int e = (base + j) + ((base + j) & ~7) * 3;          // element (16-byte) index of row 0
        if (use_tex) {
            const uint4 a = tex1Dfetch<uint4>(tex, e), b = tex1Dfetch<uint4>(tex, e + 8), cc = tex1Dfetch<uint4>(tex, e + 16), d = tex1Dfetch<uint4>(tex, e + 24);
            acc += __hiloint2double(a.y, a.x) + __hiloint2double(b.w, b.z) + __hiloint2double(cc.y, cc.x) + __hiloint2double(d.w, d.z);
        } else {
            const double2 a = rec[e], b = rec[e + 8], cc = rec[e + 16], d = rec[e + 24];
            acc += a.x + b.y + cc.x + d.y;
        }
        if (WITH_LDS) {
#pragma unroll
            for (int u = 0; u < 12; ++u) {       // per-lane addresses, like the candidate scan
                const float4 v = sm[(j + off * 5 + u * 7) & 511];
                facc += v.x;
            }
        }
        c += 13; if (c >= 1536) c -= 1536;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == -1.0 || facc == -1.f) *sink = acc + facc;
}

template <int MODE, bool WITH_LDS>
void run(const char* name, const double2* rec, cudaTextureObject_t tex) {
    int h[256];
    srand(7);
    for (int l = 0; l < 256; ++l) h[l] = rand() % 72;
    int* d; long long* o; double* sink;
    const int nb = 296;
    cudaMalloc(&d, sizeof h); cudaMalloc(&o, 8 * nb); cudaMalloc(&sink, 8);
    cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
    const int iters = 20000;
    k<MODE, WITH_LDS><<<nb, 256>>>(rec, tex, d, o, sink, iters);
    k<MODE, WITH_LDS><<<nb, 256>>>(rec, tex, d, o, sink, iters);
    long long ho[296];
    cudaMemcpy(ho, o, sizeof ho, cudaMemcpyDeviceToHost);
    // two blocks of 8 warps share an SM: the block's clock spans 16 warps' worth of records
    printf("%-44s %.2f cycles per record-warp per SM (%s)\n", name, (double)ho[0] / (iters * 16.0), cudaGetErrorString(cudaGetLastError()));
    cudaFree(d); cudaFree(o); cudaFree(sink);
}

int main() {
    const size_t n16 = (size_t)296 * 1536 * 4 + 64;      // 16-byte elements
    double2* rec;
    cudaMalloc(&rec, n16 * 16);
    cudaMemset(rec, 0, n16 * 16);
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = rec;
    rd.res.linear.desc = cudaCreateChannelDesc<uint4>();
    rd.res.linear.sizeInBytes = n16 * 16;
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    cudaError_t e = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    printf("texture object: %s\n", cudaGetErrorString(e));
    run<0, false>("LDG gathers", rec, tex);
    run<1, false>("TEX gathers", rec, tex);
    run<2, false>("half the warps LDG, half TEX", rec, tex);
    run<0, true>("LDG gathers + 12 LDS.128 per record", rec, tex);
    run<1, true>("TEX gathers + 12 LDS.128 per record", rec, tex);
    run<2, true>("half LDG, half TEX + 12 LDS.128 per record", rec, tex);
    return 0;
}
