// qrstep.cu - what one step of a one-warp systolic Givens sweep costs on B200, in a few code shapes
// (csrc/fit_coop.h register sweep).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o qrstep qrstep.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../blueberry_b200/csrc/fit_coop.h"

template <int F>
__device__ __forceinline__ void givens_f(double piv, double& ww, double& cs, double& sn) {
    const double store = fabs(piv);
    const bool big = store >= ww;
    const double mx = big ? store : ww, mn = big ? ww : store;
    const double r = (F & 1) ? mn * 0.5 : mn / mx;
    const double dd = (F & 2) ? mx * (1.0 + r * r) : mx * sqrt(1.0 + r * r);
    if (F & 4) { cs = ww * 0.5; sn = piv * 0.25; } else bbk_div2(ww, piv, dd, cs, sn);
    ww = dd;
}
#define G_(j, i) g[((j) - 1) * 5 + ((i) - 1)]
#define C_(j) c[(j) - 1]

// V0: the shape used in fit_coop.h today
template <int V>
__global__ void __launch_bounds__(128, 1) k(int n8, int nk1, int reps, long long* cyc, double* sink) {
    __shared__ double g[128 * 5];
    __shared__ double c[128];
    const int lane = threadIdx.x;
    if (threadIdx.x >= 32) return;
    double acc = 0.0;
    long long total = 0;
    for (int r = 0; r < reps; ++r) {
        for (int i = lane; i < 128 * 5; i += 32) g[i] = 1.0 + 0.001 * ((i * 7 + r) % 13);
        for (int i = lane; i < 128; i += 32) c[i] = 0.5 + 0.01 * (i % 5);
        double h[5];
        for (int i = 0; i < 5; ++i) h[i] = 0.1 * (lane + 1) + 0.01 * i;
        double yi = 0.0;
        const int it = lane + 1;
        __syncwarp();
        const int last_step = n8 + nk1 - 2;
        long long t0 = clock64();
        for (int step = 0; step <= last_step; ++step) {
            const int j = step - it + 2;
            const bool on = it <= n8 && j >= it && j <= nk1;
            if (V == 0) {
                if (on) {
                    double g1 = G_(j, 1), g2 = G_(j, 2), g3 = G_(j, 3), g4 = G_(j, 4), g5 = G_(j, 5), cj = C_(j), cs, sn;
                    bbk_givens_v(h[0], g1, cs, sn);
                    bbk_rotate_v(cs, sn, yi, cj);
                    G_(j, 1) = g1; C_(j) = cj;
                    if (j != nk1) {
                        const int i2 = j > n8 ? nk1 - j : 4;
                        if (i2 >= 1) { bbk_rotate_v(cs, sn, h[1], g2); h[0] = h[1]; G_(j, 2) = g2; }
                        if (i2 >= 2) { bbk_rotate_v(cs, sn, h[2], g3); h[1] = h[2]; G_(j, 3) = g3; }
                        if (i2 >= 3) { bbk_rotate_v(cs, sn, h[3], g4); h[2] = h[3]; G_(j, 4) = g4; }
                        if (i2 >= 4) { bbk_rotate_v(cs, sn, h[4], g5); h[3] = h[4]; G_(j, 5) = g5; }
                        if (i2 == 0) h[0] = 0.0;
                        if (i2 == 1) h[1] = 0.0;
                        if (i2 == 2) h[2] = 0.0;
                        if (i2 == 3) h[3] = 0.0;
                        if (i2 == 4) h[4] = 0.0;
                    }
                }
            } else {
                // V1: every lane runs the same straight line on safe operands; only the stores are predicated
                const int jj = on ? j : 1;
                double g1 = G_(jj, 1), g2 = G_(jj, 2), g3 = G_(jj, 3), g4 = G_(jj, 4), g5 = G_(jj, 5), cj = C_(jj), cs, sn;
                double piv = on ? h[0] : 1.0;
                if (!on) g1 = 1.0;
                givens_f<(V >> 4)>(piv, g1, cs, sn);
                const double yn = cs * yi - sn * cj, cn = cs * cj + sn * yi;
                const double h1n = cs * h[1] - sn * g2, g2n = cs * g2 + sn * h[1];
                const double h2n = cs * h[2] - sn * g3, g3n = cs * g3 + sn * h[2];
                const double h3n = cs * h[3] - sn * g4, g4n = cs * g4 + sn * h[3];
                const double h4n = cs * h[4] - sn * g5, g5n = cs * g5 + sn * h[4];
                const int i2 = (j == nk1) ? -1 : (j > n8 ? nk1 - j : 4);
                if (on) {
                    G_(jj, 1) = g1; C_(jj) = cn; yi = yn;
                    if (i2 >= 1) G_(jj, 2) = g2n;
                    if (i2 >= 2) G_(jj, 3) = g3n;
                    if (i2 >= 3) G_(jj, 4) = g4n;
                    if (i2 >= 4) G_(jj, 5) = g5n;
                    if (i2 >= 0) {
                        h[0] = i2 >= 1 ? h1n : 0.0;
                        h[1] = i2 >= 2 ? h2n : 0.0;
                        h[2] = i2 >= 3 ? h3n : 0.0;
                        h[3] = i2 >= 4 ? h4n : 0.0;
                        h[4] = 0.0;
                    }
                }
            }
            __syncwarp();
        }
        long long t1 = clock64();
        total += t1 - t0;
        acc += yi + h[0] + h[1] + h[2] + h[3] + h[4];
    }
    if (lane == 0) { cyc[0] = total; cyc[1] = (long long)reps * (n8 + nk1 - 1); }
    sink[lane] = acc + g[lane] + c[lane];
}

int main() {
    long long* cyc; double* sink;
    cudaMalloc(&cyc, 64); cudaMalloc(&sink, 1024);
    for (int v = 0; v < 10; ++v)
        for (int n8 : {15}) {
            int nk1 = n8 + 4;
            switch (v) {
                case 0: k<0><<<1, 128>>>(n8, nk1, 20, cyc, sink); break;
                case 1: k<1><<<1, 128>>>(n8, nk1, 20, cyc, sink); break;
                case 2: k<1 + 16><<<1, 128>>>(n8, nk1, 20, cyc, sink); break;     // no first division
                case 3: k<1 + 32><<<1, 128>>>(n8, nk1, 20, cyc, sink); break;     // no sqrt
                case 4: k<1 + 64><<<1, 128>>>(n8, nk1, 20, cyc, sink); break;     // no cos/sin division
                case 5: k<1 + 16 + 32><<<1, 128>>>(n8, nk1, 20, cyc, sink); break;
                case 6: k<1 + 16 + 32 + 64><<<1, 128>>>(n8, nk1, 20, cyc, sink); break;   // no div, no sqrt at all
                default: continue;
            }
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[2]; cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
            printf("variant %d n8=%d: %.0f cycles per step (%lld steps)\n", v, n8, (double)h[0] / (double)h[1], h[1]);
        }
    return 0;
}
