// Dependent-chain latencies on the GPU at hand (one warp): DFMA, DMUL+DADD, LDS.64 pointer chase, sqrt(double), 1/x double.
#include <cuda_runtime.h>
#include <cstdio>
__global__ void probe(double* out, long long* cyc, int iters)
{
    __shared__ double sm[1024];
    __shared__ int chase[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) { sm[i] = 1.0 + 1e-9 * i; chase[i] = (i * 33 + 7) & 1023; }
    __syncthreads();
    double a = 1.0 + threadIdx.x * 1e-12, b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) a = fma(a, b, c);
    long long t1 = clock64();
    double m = a;
    for (int i = 0; i < iters; ++i) { m = __dmul_rn(m, b); m = __dadd_rn(m, c); }
    long long t2 = clock64();
    int p = threadIdx.x;
    for (int i = 0; i < iters; ++i) p = chase[p];
    long long t3 = clock64();
    double s = a;
    for (int i = 0; i < iters; ++i) s = sqrt(s + 1.0);
    long long t4 = clock64();
    double d = m;
    for (int i = 0; i < iters; ++i) d = __ddiv_rn(1.0, d + 1.5);
    long long t5 = clock64();
    double r = 0.0; int q = threadIdx.x & 1023;
    for (int i = 0; i < iters; ++i) { double v = sm[q]; v += 1.0; sm[q] = v; q = (q + 32) & 1023; }  // smem RMW, independent addresses
    long long t6 = clock64();
    q = threadIdx.x & 1023;
    for (int i = 0; i < iters; ++i) { double v = sm[q]; v += 1.0; sm[q] = v; }  // smem RMW, same address (true dependency)
    long long t7 = clock64();
    float f = (float)a;
    for (int i = 0; i < iters; ++i) f = fmaf(f, 1.0000001f, 1e-9f);
    long long t8 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; cyc[6] = t7 - t6; cyc[7] = t8 - t7; }
    out[threadIdx.x] = a + m + p + s + d + r + sm[q] + f;
}
int main()
{
    double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 64);
    const int iters = 4096;
    for (int warps = 1; warps <= 8; warps *= 8)
    {
        probe<<<1, 32 * warps>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        const char* nm[8] = { "DFMA", "DMUL+DADD", "LDS.32 chase", "sqrt(double)", "1/x (ddiv_rn)", "smem RMW (indep addr)", "smem RMW (same addr)", "FFMA" };
        printf("%d warp(s) in the CTA, cycles per dependent iteration:\n", warps);
        for (int k = 0; k < 8; ++k) printf("  %-24s %.1f\n", nm[k], (double)h[k] / iters);
    }
    return 0;
}
