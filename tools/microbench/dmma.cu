// Calibration microbenchmark (not part of the product): fp64 tensor-core MMA (mma.sync m8n8k4 f64, SASS DMMA) on B200:
// throughput alone, beside DFMA in the same warp, beside shared-memory traffic, and warp-shuffle / LDS wavefront rates.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dmma dmma.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// NM independent DMMA chains + NF independent DFMA chains per thread
template <int NM, int NF>
__global__ void k_mix(double *out, int iters) {
    double c[NM > 0 ? NM : 1][2], f[NF > 0 ? NF : 1];
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < NM; ++i) c[i][0] = c[i][1] = i;
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = i + a;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (NM > NF ? NM : NF); ++i) {
            if (i < NM) dmma(c[i][0], c[i][1], a, b);
            if (i < NF) f[i] = fma(f[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NM; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
    if (s == 1.2345) *out = s;
}

// NM DMMA chains + NL conflict-free LDS.64 per iteration (2 wavefronts per warp instruction)
template <int NM, int NL>
__global__ void k_mma_lds(double *out, int iters) {
    __shared__ double sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i * 1e-9;
    __syncthreads();
    double c[NM > 0 ? NM : 1][2];
#pragma unroll
    for (int i = 0; i < NM; ++i) c[i][0] = c[i][1] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    int idx = threadIdx.x & 255;
    for (int it = 0; it < iters; ++it) {
        double s = 0;
#pragma unroll
        for (int i = 0; i < NL; ++i) s += sm[(idx + 256 * i) & 2047];
        b = s;
#pragma unroll
        for (int i = 0; i < NM; ++i) dmma(c[i][0], c[i][1], a, b);
        idx = (idx + 32) & 255;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NM; ++i) s += c[i][0] + c[i][1];
    if (s == 1.2345) *out = s;
}

template <int NS>
__global__ void k_shfl(double *out, int iters) {
    double v[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) v[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NS; ++i) v[i] = __shfl_xor_sync(0xffffffffu, v[i], 1 + (it & 15));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NS; ++i) s += v[i];
    if (s == 1.2345) *out = s;
}

int main() {
    double *out;
    CK(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    auto timeit = [&](auto f) {
        f(); f();
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        return (double)ms;
    };
    for (int warps : {4, 8, 16}) {
        const int g = 148 * 2, th = warps * 32;
        const double W = (double)g * warps;  // warps in flight (2 CTAs per SM)
        printf("--- %d CTAs x %d threads\n", g, th);
        double ms;
        ms = timeit([&] { k_mix<0, 8><<<g, th>>>(out, iters); });
        printf("DFMA only  (8 chains): %7.3f ms  %6.2f TFLOP/s\n", ms, 2.0 * 8 * iters * W * 32 / ms / 1e9);
        ms = timeit([&] { k_mix<8, 0><<<g, th>>>(out, iters); });
        printf("DMMA only  (8 chains): %7.3f ms  %6.2f TFLOP/s  (%.2f clk/DMMA/SM at 1.9 GHz)\n", ms, 512.0 * 8 * iters * W / ms / 1e9,
               ms * 1e-3 * 1.9e9 / (8.0 * iters * W / 148));
        ms = timeit([&] { k_mix<4, 0><<<g, th>>>(out, iters); });
        printf("DMMA only  (4 chains): %7.3f ms  %6.2f TFLOP/s\n", ms, 512.0 * 4 * iters * W / ms / 1e9);
        ms = timeit([&] { k_mix<2, 0><<<g, th>>>(out, iters); });
        printf("DMMA only  (2 chains): %7.3f ms  %6.2f TFLOP/s\n", ms, 512.0 * 2 * iters * W / ms / 1e9);
        ms = timeit([&] { k_mix<8, 8><<<g, th>>>(out, iters); });
        printf("DMMA+DFMA  (8+8)     : %7.3f ms  %6.2f TFLOP/s total\n", ms, (512.0 * 8 + 64.0 * 8) * iters * W / ms / 1e9);
        ms = timeit([&] { k_mix<8, 4><<<g, th>>>(out, iters); });
        printf("DMMA+DFMA  (8+4)     : %7.3f ms  %6.2f TFLOP/s total\n", ms, (512.0 * 8 + 64.0 * 4) * iters * W / ms / 1e9);
        ms = timeit([&] { k_mma_lds<8, 0><<<g, th>>>(out, iters); });
        printf("DMMA 8 + LDS 0       : %7.3f ms\n", ms);
        ms = timeit([&] { k_mma_lds<8, 8><<<g, th>>>(out, iters); });
        printf("DMMA 8 + LDS.64 8    : %7.3f ms  (%.2f clk/LDS/SM)\n", ms, ms * 1e-3 * 1.9e9 / (8.0 * iters * W / 148));
        ms = timeit([&] { k_mma_lds<0, 8><<<g, th>>>(out, iters); });
        printf("LDS.64 8 only        : %7.3f ms  (%.2f clk/LDS/SM)\n", ms, ms * 1e-3 * 1.9e9 / (8.0 * iters * W / 148));
        ms = timeit([&] { k_shfl<8><<<g, th>>>(out, iters); });
        printf("SHFL f64 x8 (16 SHFL): %7.3f ms  (%.2f clk/SHFL.32/SM)\n", ms, ms * 1e-3 * 1.9e9 / (16.0 * iters * W / 148));
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
