// band.cu - K6: count_band_regions (reference: blueberry.pyx:77-91)
//     t = #{(i, j) : j < i, low <= regions[i] - regions[j] <= high}
// The reference is an O(n^2) double loop (3e10 iterations for chr1 at 1 kb).  For sorted input
// (what np.union1d hands it, datatypes.pyx:119) the FP64 difference regions[i] - regions[j] is
// monotone in j, so two binary searches per i give the exact count in O(n log n); unsorted input
// takes the exact tiled O(n^2) path.  Which path runs is decided on the device (no host sync).
#include "common.cuh"

namespace {

constexpr int BAND_THREADS = 256;

__global__ void band_init_kernel(long long* result, int* unsorted) { *result = 0; *unsorted = 0; }

__global__ void band_sorted_check_kernel(const double* r, long long n, int* unsorted) {
    long long stride = (long long)gridDim.x * blockDim.x;
    int bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += stride)
        bad |= !(r[i - 1] <= r[i]);                     // NaN counts as unsorted
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(unsorted, 1);
}

__device__ __forceinline__ void block_add(long long v, long long* result) {
    __shared__ long long red[BAND_THREADS / 32];
    v = warp_sum_ll(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < BAND_THREADS / 32; ++w) t += red[w];
        if (t) atomicAdd((unsigned long long*)result, (unsigned long long)t);
    }
}

__global__ void __launch_bounds__(BAND_THREADS) band_sorted_kernel(const double* r, long long n, double low, double high,
                                                                   const int* unsorted, long long* result) {
    if (*unsorted) return;
    long long stride = (long long)gridDim.x * blockDim.x, cnt = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double ri = r[i];
        // diff(j) = ri - r[j] is non-increasing in j on [0, i)
        long long lo = 0, hi = i;                       // first j with diff <= high
        while (lo < hi) { long long mid = (lo + hi) >> 1; if (ri - r[mid] <= high) hi = mid; else lo = mid + 1; }
        long long j_hi = lo;
        lo = j_hi; hi = i;                              // first j with diff < low
        while (lo < hi) { long long mid = (lo + hi) >> 1; if (ri - r[mid] < low) hi = mid; else lo = mid + 1; }
        cnt += lo - j_hi;
    }
    block_add(cnt, result);
}

// exact for any order: CTA (bi, bj) compares a tile of i against a tile of j <= i
__global__ void __launch_bounds__(BAND_THREADS) band_brute_kernel(const double* r, long long n, double low, double high,
                                                                  const int* unsorted, long long* result) {
    if (!*unsorted) return;
    __shared__ double sj[BAND_THREADS];
    const long long n_tiles = (n + BAND_THREADS - 1) / BAND_THREADS;
    const long long n_work = n_tiles * (n_tiles + 1) / 2;
    long long cnt = 0;
    for (long long w = blockIdx.x; w < n_work; w += gridDim.x) {
        // w -> (ti, tj) with tj <= ti
        long long ti = (long long)((sqrt(8.0 * (double)w + 1.0) - 1.0) * 0.5);
        while (ti * (ti + 1) / 2 > w) --ti;
        while ((ti + 1) * (ti + 2) / 2 <= w) ++ti;
        long long tj = w - ti * (ti + 1) / 2;
        long long i = ti * BAND_THREADS + threadIdx.x;
        long long j0 = tj * BAND_THREADS;
        __syncthreads();
        sj[threadIdx.x] = (j0 + threadIdx.x < n) ? r[j0 + threadIdx.x] : 0.0;
        __syncthreads();
        if (i < n) {
            double ri = r[i];
            long long jmax = i - j0;                    // j < i
            if (jmax > BAND_THREADS) jmax = BAND_THREADS;
            if (j0 + jmax > n) jmax = n - j0;
            for (int jj = 0; jj < jmax; ++jj) {
                double d = ri - sj[jj];
                cnt += (low <= d && d <= high) ? 1 : 0;
            }
        }
    }
    block_add(cnt, result);
}

}  // namespace

extern "C" int bbk_count_band(const double* d_regions, int64_t n, double low, double high, int64_t* d_result, void* stream) {
    BBK_REQUIRE(n >= 0 && d_result, "bbk_count_band: bad arguments");
    BBK_REQUIRE(n == 0 || d_regions, "bbk_count_band: null regions");
    cudaStream_t st = (cudaStream_t)stream;
    // the unsorted flag lives right after the result in a small static device buffer per call: keep it
    // inside the caller's result allocation instead -> caller provides 2 int64 (result, scratch)
    int* flag = (int*)(d_result + 1);
    band_init_kernel<<<1, 1, 0, st>>>((long long*)d_result, flag);
    BBK_CHECK_LAUNCH("band_init_kernel");
    if (n < 2) return BBK_OK;
    int sms = bbk_num_sms();
    long long want = (n + BAND_THREADS - 1) / BAND_THREADS;
    int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
    band_sorted_check_kernel<<<grid, BAND_THREADS, 0, st>>>(d_regions, n, flag);
    BBK_CHECK_LAUNCH("band_sorted_check_kernel");
    band_sorted_kernel<<<grid, BAND_THREADS, 0, st>>>(d_regions, n, low, high, flag, (long long*)d_result);
    BBK_CHECK_LAUNCH("band_sorted_kernel");
    band_brute_kernel<<<sms * 8, BAND_THREADS, 0, st>>>(d_regions, n, low, high, flag, (long long*)d_result);
    BBK_CHECK_LAUNCH("band_brute_kernel");
    return BBK_OK;
}
