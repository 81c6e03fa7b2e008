// synth.cu - synthetic Hi-C contact records of the BASELINE shapes, generated on the device
// (bench / smoke only; never on a parity path - tests upload host arrays).
// One chromosome of n_bins bins: every pair (i, i+d), 0 <= d <= K, i+d < n_bins, row-major,
// zeros kept.  count ~ Poisson(depth * b_i * b_j * (d+1)^-decay * (1 + 4*loop)), loop ~ Bernoulli(1e-3)
// for d >= 5.  RNG: counter-based (splitmix64 of (seed, record index)), so any shard is generable
// anywhere.
#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(unsigned long long x) { return ((x >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

__device__ int poisson_draw(double lam, unsigned long long& state) {
    if (lam <= 0.0) return 0;
    if (lam < 30.0) {                                   // inversion by sequential search
        double u = u01(state = mix64(state));
        double p = exp(-lam), cdf = p;
        int k = 0;
        while (u > cdf && k < 400) { ++k; p *= lam / k; cdf += p; }
        return k;
    }
    // normal approximation with continuity correction (only the bulk of near-diagonal counts)
    double u1 = u01(state = mix64(state)), u2 = u01(state = mix64(state));
    double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    double v = lam + sqrt(lam) * z + 0.5;
    return v < 0.0 ? 0 : (int)v;
}

__global__ void __launch_bounds__(256) synth_kernel(long long n_bins, long long K, long long R, double depth, double decay,
                                                    unsigned long long seed, const double* bias,
                                                    int* mid1, int* mid2, int* count, long long first, long long n_pairs) {
    // rows i < n_bins - K are full (K+1 records); the last K rows shrink by one each
    const long long full_rows = n_bins - K > 0 ? n_bins - K : 0;
    const long long full_pairs = full_rows * (K + 1);
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < n_pairs; o += stride) {
        const long long r = first + o;                  // record number inside the chromosome (the RNG counter)
        long long i, d;
        if (r < full_pairs) { i = r / (K + 1); d = r - i * (K + 1); }
        else {
            // tail rows: row full_rows + t has (K - t) records (t = 0 .. ), offsets t*(K) - t(t-1)/2 ...
            long long rem = r - full_pairs;
            long long Kt = n_bins - full_rows;          // number of tail rows (= min(K, n_bins))
            // row t holds (Kt - t) records; cumulative before row t: t*Kt - t(t-1)/2
            double disc = (2.0 * Kt + 1.0) * (2.0 * Kt + 1.0) - 8.0 * (double)rem;
            long long t = (long long)(((2.0 * Kt + 1.0) - sqrt(disc > 0 ? disc : 0)) * 0.5);
            while (t > 0 && t * Kt - t * (t - 1) / 2 > rem) --t;
            while ((t + 1) * Kt - (t + 1) * t / 2 <= rem) ++t;
            i = full_rows + t;
            d = rem - (t * Kt - t * (t - 1) / 2);
        }
        unsigned long long state = mix64(seed ^ mix64((unsigned long long)r));
        double lam = depth * pow((double)(d + 1), -decay);
        if (bias) lam *= bias[i] * bias[i + d];
        if (d >= 5 && u01(state = mix64(state)) < 1e-3) lam *= 5.0;
        mid1[o] = (int)(i * R + R / 2);
        mid2[o] = (int)((i + d) * R + R / 2);
        count[o] = poisson_draw(lam, state);
    }
}

}  // namespace

extern "C" int64_t bbk_synth_n_pairs(int64_t n_bins, int64_t K) {
    if (n_bins <= 0 || K < 0) return 0;
    if (K > n_bins - 1) K = n_bins - 1;
    return (K + 1) * n_bins - K * (K + 1) / 2;
}

extern "C" int bbk_synth_contacts_range(int64_t n_bins, int64_t K, int64_t resolution, double depth, double decay, uint64_t seed,
                                        const double* d_bias, int64_t first_record, int64_t n_records, int32_t* d_mid1,
                                        int32_t* d_mid2, int32_t* d_count, void* stream) {
    BBK_REQUIRE(n_bins > 0 && K >= 0 && resolution > 0, "bbk_synth_contacts: bad shape");
    BBK_REQUIRE((n_bins + 1) * resolution < (1ll << 31), "bbk_synth_contacts: coordinates overflow int32");
    if (K > n_bins - 1) K = n_bins - 1;
    long long n_pairs = bbk_synth_n_pairs(n_bins, K);
    BBK_REQUIRE(first_record >= 0 && n_records >= 0 && first_record + n_records <= n_pairs, "bbk_synth_contacts: record range outside the chromosome");
    if (n_records == 0) return BBK_OK;
    BBK_REQUIRE(d_mid1 && d_mid2 && d_count, "bbk_synth_contacts: null output");
    int grid = bbk_num_sms() * 8;
    synth_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n_bins, K, resolution, depth, decay, seed, d_bias,
                                                          d_mid1, d_mid2, d_count, first_record, n_records);
    BBK_CHECK_LAUNCH("synth_kernel");
    return BBK_OK;
}

extern "C" int bbk_synth_contacts(int64_t n_bins, int64_t K, int64_t resolution, double depth, double decay, uint64_t seed,
                                  const double* d_bias, int32_t* d_mid1, int32_t* d_mid2, int32_t* d_count, void* stream) {
    BBK_REQUIRE(n_bins > 0 && K >= 0, "bbk_synth_contacts: bad shape");
    return bbk_synth_contacts_range(n_bins, K, resolution, depth, decay, seed, d_bias, 0,
                                    bbk_synth_n_pairs(n_bins, K > n_bins - 1 ? n_bins - 1 : K), d_mid1, d_mid2, d_count, stream);
}
