// Development aid: dependent-chain latencies of the fp64 operations on the serial paths (1 warp).
#include <cstdio>
#include <cuda_runtime.h>
#define REP 256
template <int OP>
__global__ void chain(double* out, double seed, long long* clk) {
    double x = seed + threadIdx.x * 1e-9, y = 1.0000001, z = 0.9999;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REP; ++i) {
        if (OP == 0) x = fma(x, y, z);
        if (OP == 1) x = x * y;
        if (OP == 2) x = x + y;
        if (OP == 3) x = y / x + 1.0;
        if (OP == 4) x = sqrt(x) + 1.0;
        if (OP == 5) x = exp(x * 1e-3);
        if (OP == 6) { double s, c; sincos(x, &s, &c); x = s + c; }
        if (OP == 7) x = tan(x * 0.1) + 0.05;
        if (OP == 8) x = atan(x);
        if (OP == 9) x = hypot(x, y);
        if (OP == 10) x = 1.0 / x + 0.5;
        if (OP == 11) x = cos(x);
        if (OP == 12) x = __fma_rn(x, y, z) * y + z;  // 2 dependent
        if (OP == 13) x = sqrt(x * x + y * y);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *clk = t1 - t0;
}
__global__ void ldchain(const int* next, int* out, long long* clk, int n) {
    int j = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) j = next[j];
    long long t1 = clock64();
    *out = j;
    *clk = t1 - t0;
}
int main() {
    double* out; long long* clk; cudaMalloc(&out, 1024); cudaMalloc(&clk, 8);
    const char* names[] = {"dfma", "dmul", "dadd", "ddiv+add", "dsqrt+add", "exp", "sincos", "tan", "atan", "hypot", "drcp+add", "cos", "2xdfma", "sqrt(fma)"};
#define RUN(OP) { chain<OP><<<1, 32>>>(out, 0.7, clk); chain<OP><<<1, 32>>>(out, 0.7, clk); long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("%-10s %7.1f clk/op\n", names[OP], double(h) / REP); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13)
    // pointer chase through L2 (stride 4 KB over 64 MB) and L1 (small)
    for (int mode = 0; mode < 2; ++mode) {
        int n = mode == 0 ? (16 << 20) : 1024, stride = mode == 0 ? 1024 + 32 : 1;
        int* h = new int[n];
        for (int i = 0; i < n; ++i) h[i] = int((long long)(i + stride) % n);
        int* d; int* o; cudaMalloc(&d, n * 4L); cudaMalloc(&o, 4); cudaMemcpy(d, h, n * 4L, cudaMemcpyHostToDevice);
        ldchain<<<1, 1>>>(d, o, clk, 2000); ldchain<<<1, 1>>>(d, o, clk, 2000);
        long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
        printf("ld chain %s: %.1f clk/load\n", mode == 0 ? "L2/DRAM (64 MB, strided)" : "L1 (4 KB)", double(c) / 2000);
        cudaFree(d); delete[] h;
    }
    int dev; cudaGetDevice(&dev); int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev); printf("clock %d MHz\n", khz / 1000);
    return 0;
}
