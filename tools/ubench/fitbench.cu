// fitbench.cu - times the cooperative spline search (csrc/fit_coop.h) alone on one (x, y) set, with the
// per-phase cycle counters it keeps.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#ifndef NO_QR_PROFILE
#define BBK_QR_PROFILE 1
#endif
#include "../../blueberry_b200/csrc/fit_coop.h"

__global__ void __launch_bounds__(128, 1) k(const double* x, const double* y, int m, double s, double* ws, BbkCoopState* out, long long* cyc) {
    extern __shared__ double pool[];
    __shared__ BbkCoopState st;
    double* xs = pool;
    double* ys = pool + m;
    for (int j = threadIdx.x; j < m; j += blockDim.x) { xs[j] = x[j]; ys[j] = y[j]; }
    __syncthreads();
    BbkCoopWs cw;
    bbk_coop_ws_carve(pool + 2 * m, m, &cw);
    long long t0 = clock64();
    bbk_coop_univariate_spline_t<true>(xs, ys, m, s, &st, &cw);
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) { *out = st; cyc[0] = t1 - t0; }
}

int main(int argc, char** argv) {
    FILE* f = fopen(argc > 1 ? argv[1] : "tools/ubench/xy.bin", "rb");
    if (!f) { printf("no input\n"); return 1; }
    double dm; fread(&dm, 8, 1, f);
    int m = (int)dm;
    std::vector<double> x(m), y(m);
    fread(x.data(), 8, m, f); fread(y.data(), 8, m, f); fclose(f);
    double ymin = y[0]; for (double v : y) ymin = v < ymin ? v : ymin;
    double s = ymin * ymin;
    double *dx, *dy, *dws; BbkCoopState* dst; long long* dc;
    cudaMalloc(&dx, 8 * m); cudaMalloc(&dy, 8 * m); cudaMalloc(&dws, 1 << 20); cudaMalloc(&dst, sizeof(BbkCoopState)); cudaMalloc(&dc, 64);
    cudaMemcpy(dx, x.data(), 8 * m, cudaMemcpyHostToDevice); cudaMemcpy(dy, y.data(), 8 * m, cudaMemcpyHostToDevice);
    size_t smem = (bbk_coop_ws_doubles(m) + 2 * m + 16) * 8;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 3; ++rep) {
        k<<<1, 128, smem>>>(dx, dy, m, s, dws, dst, dc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        BbkCoopState st; long long c;
        cudaMemcpy(&st, dst, sizeof(st), cudaMemcpyDeviceToHost); cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        printf("m=%d n=%d ier=%d total %lld cyc | fits %lld piters %lld | rows+QR %lld backsub %lld resid+knots %lld sweep %lld f(p) %lld | per-fit QR %lld\n",
               m, st.n, st.ier, c, st.diag[0], st.diag[1], st.diag[2], st.diag[3], st.diag[4], st.diag[5], st.diag[6],
               st.diag[0] ? st.diag[2] / st.diag[0] : 0);
#ifdef BBK_QR_PROFILE
        long long pr[8];
        cudaMemcpyFromSymbol(pr, bbk_qr_prof, sizeof(pr));
        printf("   QR loops: %lld half-steps, %lld cycles -> %.0f per half-step\n", pr[0], pr[1], (double)pr[1] / (double)pr[0]);
        for (int w = 0; w < 2; ++w)
            printf("   warp %d per half-step: work on odd half-steps %.0f, on even %.0f, barrier wait %.0f\n", w,
                   2.0 * pr[2 + 3 * w] / pr[0], 2.0 * pr[3 + 3 * w] / pr[0], (double)pr[4 + 3 * w] / pr[0]);
        long long zero[8] = {0};
        cudaMemcpyToSymbol(bbk_qr_prof, zero, sizeof(zero));
#endif
    }
    return 0;
}
