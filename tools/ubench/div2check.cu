// div2check.cu - bbk_div2 (csrc/fit_coop.h) against the compiler's own double division, bit for bit, on random
// and on awkward operands.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o div2check div2check.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../blueberry_b200/csrc/fit_coop.h"

__device__ unsigned long long mix(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31);
}
// mode 0: any bit pattern; 1: Givens-like (a1, a2 <= b in magnitude, moderate exponents); 2: exponents near the edges
__device__ double draw(unsigned long long r, int mode, int which) {
    if (mode == 0) return __longlong_as_double((long long)r);
    unsigned long long mant = r & 0xfffffffffffffull, sign = (r >> 63) << 63;
    int e;
    if (mode == 1) e = 1023 - 40 + (int)((r >> 52) % 80);
    else { int k = (int)((r >> 52) % 64); e = (k & 1) ? k / 2 : 2046 - k / 2; if (which == 2 && (k & 2)) e = 1023 + (k - 32); }
    return __longlong_as_double((long long)(sign | ((unsigned long long)e << 52) | mant));
}
__global__ void k(unsigned long long seed, int mode, long long per_thread, unsigned long long* bad, double* first) {
    unsigned long long id = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long local = 0;
    for (long long i = 0; i < per_thread; ++i) {
        unsigned long long s = mix(seed + id * 0x100000001b3ull + (unsigned long long)i * 0x9e3779b97f4a7c15ull);
        double a1 = draw(s, mode, 0), a2 = draw(mix(s), mode, 1), b = draw(mix(mix(s)), mode, 2);
        if (mode == 1) { double m = fmax(fabs(a1), fabs(a2)); if (fabs(b) < m) b = copysign(m * 1.0000001, b); }
        if ((i & 1023) == 7) a1 = 0.0;
        if ((i & 4095) == 9) a2 = -0.0;
        double q1, q2;
        bbk_div2(a1, a2, b, q1, q2);
        double w1 = a1 / b, w2 = a2 / b;
        bool same = (__double_as_longlong(q1) == __double_as_longlong(w1) || (w1 != w1 && q1 != q1)) &&
                    (__double_as_longlong(q2) == __double_as_longlong(w2) || (w2 != w2 && q2 != q2));
        if (!same) { if (atomicAdd(bad, 1ull) == 0) { first[0] = a1; first[1] = a2; first[2] = b; first[3] = q1; first[4] = w1; first[5] = q2; first[6] = w2; } }
        local += 1;
    }
    atomicAdd(bad + 1, local);
}
int main() {
    unsigned long long* bad; double* first;
    cudaMallocManaged(&bad, 16); cudaMallocManaged(&first, 64);
    for (int mode = 0; mode < 3; ++mode) {
        bad[0] = bad[1] = 0;
        k<<<148 * 8, 256>>>(12345 + mode, mode, 4096, bad, first);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        printf("mode %d: %llu triples, %llu mismatches", mode, bad[1], bad[0]);
        if (bad[0]) printf("  first: a1=%a a2=%a b=%a  div2 %a vs %a ; %a vs %a", first[0], first[1], first[2], first[3], first[4], first[5], first[6]);
        printf("\n");
    }
    return 0;
}
