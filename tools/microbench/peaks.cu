// Calibration microbenchmarks (not part of the product): read-only stream bandwidth, copy bandwidth, fp64 DFMA peak.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int VEC, bool NA>
__global__ void k_read(const double *__restrict__ p, size_t n, double *out) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC, stride = (size_t)gridDim.x * blockDim.x * VEC;
    double s = 0;
    for (; i + VEC <= n; i += stride) {
        if (VEC == 2) {
            double2 v;
            if (NA) asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p + i));
            else v = *reinterpret_cast<const double2 *>(p + i);
            s += v.x + v.y;
        } else {
            double v;
            if (NA) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p + i));
            else v = p[i];
            s += v;
        }
    }
    if (s == 12345.678) *out = s;
}
__global__ void k_copy(const double2 *__restrict__ a, double2 *__restrict__ b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) b[i] = a[i];
}
__global__ void k_dfma(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 1.2345) *out = a0;
}
int main() {
    size_t n = (size_t)3 << 27;  // 3 GiB of doubles = 402M doubles
    double *a, *b, *out;
    CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&b, n * 8)); CK(cudaMalloc(&out, 8));
    CK(cudaMemset(a, 0, n * 8)); CK(cudaMemset(b, 0, n * 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](auto f, const char *name, double bytes) {
        for (int i = 0; i < 3; ++i) f();
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) f();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
        printf("%-44s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / ms / 1e6);
    };
    for (int g : {148 * 8, 148 * 16, 148 * 32}) {
        printf("grid %d x 256\n", g);
        timeit([&] { k_read<1, false><<<g, 256>>>(a, n, out); }, " read  8B/lane  default", n * 8.0);
        timeit([&] { k_read<1, true><<<g, 256>>>(a, n, out); }, " read  8B/lane  nc.L1::no_allocate", n * 8.0);
        timeit([&] { k_read<2, false><<<g, 256>>>(a, n, out); }, " read 16B/lane  default", n * 8.0);
        timeit([&] { k_read<2, true><<<g, 256>>>(a, n, out); }, " read 16B/lane  nc.L1::no_allocate", n * 8.0);
        timeit([&] { k_copy<<<g, 256>>>((const double2 *)a, (double2 *)b, n / 2); }, " copy 16B/lane (read+write bytes)", 2 * n * 8.0);
    }
    timeit([&] { cudaMemcpyAsync(b, a, n * 8, cudaMemcpyDeviceToDevice); }, "cudaMemcpy D2D (read+write bytes)", 2 * n * 8.0);
    {
        int iters = 20000, g = 148 * 8;
        for (int i = 0; i < 2; ++i) k_dfma<<<g, 256>>>(out, iters);
        cudaEventRecord(e0);
        k_dfma<<<g, 256>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 8 * iters * (double)g * 256;
        printf("DFMA peak: %.2f TFLOP/s fp64 (%.3f ms)\n", flops / ms / 1e9, ms);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
