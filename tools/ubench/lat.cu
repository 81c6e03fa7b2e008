// latency microbenchmark: single-thread dependent chains of FP64 ops on B200 (informs the fit-stage design)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double a, double b, double* out, long long* cyc) {
    __shared__ double sm[64];
    sm[threadIdx.x & 63] = a;
    __syncthreads();
    double x = a, y = b;
    long long t0, t1;
    const int N = 256;
    t0 = clock64(); for (int i = 0; i < N; ++i) x = x * y + b; t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0);
    t0 = clock64(); for (int i = 0; i < N; ++i) x = x / y + 1.0; t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0);
    t0 = clock64(); for (int i = 0; i < N; ++i) x = sqrt(x + y); t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0);
    int idx = (int)x & 31;
    t0 = clock64(); for (int i = 0; i < N; ++i) { idx = (int)sm[idx & 63] & 63; } t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0);
    t0 = clock64(); for (int i = 0; i < N; ++i) { __syncwarp(); sm[threadIdx.x & 63] = x; __syncwarp(); x += sm[(threadIdx.x + 1) & 63]; } t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0);
    t0 = clock64(); for (int i = 0; i < N; ++i) { __syncthreads(); x += 1.0; } t1 = clock64(); if (threadIdx.x == 0) cyc[5] = (t1 - t0);
    // givens-like chain
    double ww = a + 1.0, piv = b;
    t0 = clock64();
    for (int i = 0; i < N; ++i) {
        double dd = ww * sqrt(1.0 + (piv / ww) * (piv / ww));
        double cs = ww / dd, sn = piv / dd;
        ww = dd; piv = cs * 0.5 + sn;
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = (t1 - t0);
    t0 = clock64(); for (int i = 0; i < N; ++i) x = log(x + 2.0); t1 = clock64(); if (threadIdx.x == 0) cyc[7] = (t1 - t0);
    out[threadIdx.x] = x + idx + ww + piv;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64);
    for (int threads : {1, 32, 128}) {
        k<<<1, threads>>>(1.5, 1.0000001, out, cyc);
        cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("threads %3d: per-op cycles  dfma %.1f  ddiv+add %.1f  dsqrt+add %.1f  lds-chain %.1f  syncwarp+sts+lds %.1f  syncthreads %.1f  givens %.1f  log %.1f\n",
               threads, h[0] / 256.0, h[1] / 256.0, h[2] / 256.0, h[3] / 256.0, h[4] / 256.0, h[5] / 256.0, h[6] / 256.0, h[7] / 256.0);
    }
    return 0;
}
