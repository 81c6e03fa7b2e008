// contactmap.cu - K8: ContactMap ingest + normalize on band records (datatypes.pyx:88-171), SURVEY.md 8f row 3.
//
// The reference scatters the RAWobserved rows (pos1, pos2, count) into a dense (n_bins+1)^2 float64 matrix
// (j = int(pos1 / resolution), k = int(pos2 / resolution), both triangles, :110-116) and normalises it in place:
//     matrix[j][j+i] /= KRnorm[j] * KRnorm[j+i] * KRexpected[i]        for i < n_bins, j < n_bins - i      (:166-168)
// followed by numpy.nan_to_num (:171).  A dense matrix is impossible beyond ~25 kb on chr1, and all of this is
// elementwise on the records: the records stay records (bin1 <= bin2, value) and one pass does both steps.
#include "common.cuh"

namespace {

constexpr int CM_THREADS = 256;

__device__ __forceinline__ double nan_to_num(double v) {                 // numpy.nan_to_num defaults
    if (isnan(v)) return 0.0;
    if (isinf(v)) return v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    return v;
}

__global__ void __launch_bounds__(CM_THREADS) band_ingest_kernel(const double* pos1, const double* pos2, const double* count,
                                                                 long long n, double resolution, int* bin1, int* bin2,
                                                                 double* value, int* bad, int n_bins) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // data = nan_to_num(data) (:104), then C double -> int conversion of pos / resolution (:111-112)
        const double a = nan_to_num(pos1[i]) / resolution, b = nan_to_num(pos2[i]) / resolution;
        int j = 0, k = 0;
        if (a > -1.0 && a < 2147483647.0 && b > -1.0 && b < 2147483647.0) { j = (int)a; k = (int)b; } else *bad = 1;
        if (j > n_bins || k > n_bins) *bad = 1;                          // outside the reference's (n_bins+1)^2 matrix
        bin1[i] = j < k ? j : k;
        bin2[i] = j < k ? k : j;
        value[i] = nan_to_num(count[i]);
    }
}

__global__ void __launch_bounds__(CM_THREADS) band_normalize_kernel(const int* bin1, const int* bin2, const double* value, long long n,
                                                                    const double* kr, const double* kr_expected, int n_bins,
                                                                    double* out, int* bad) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int j = bin1[i], k = bin2[i];
        double v = value[i];
        if (j >= 0 && k < n_bins && j <= k) {                            // the loop never reaches row / column n_bins
            const double den = kr[j] * kr[k] * kr_expected[k - j];
            if (den == 0.0) *bad = 1;                                    // Cython's checked division raises ZeroDivisionError here
            v = v / den;
        }
        out[i] = nan_to_num(v);
    }
}

}  // namespace

extern "C" int bbk_contact_band_ingest(const double* d_pos1, const double* d_pos2, const double* d_count, int64_t n, int64_t resolution,
                                       int32_t n_bins, int32_t* d_bin1, int32_t* d_bin2, double* d_value, int32_t* d_bad, void* stream) {
    BBK_REQUIRE(n >= 0 && resolution > 0 && n_bins >= 0 && d_bad, "bbk_contact_band_ingest: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    BBK_CHECK_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int32_t), st));
    if (n == 0) return BBK_OK;
    BBK_REQUIRE(d_pos1 && d_pos2 && d_count && d_bin1 && d_bin2 && d_value, "bbk_contact_band_ingest: null column");
    long long want = (n + CM_THREADS - 1) / CM_THREADS;
    int grid = (int)(want < (long long)bbk_num_sms() * 8 ? want : (long long)bbk_num_sms() * 8);
    band_ingest_kernel<<<grid, CM_THREADS, 0, st>>>(d_pos1, d_pos2, d_count, n, (double)resolution, d_bin1, d_bin2, d_value, d_bad, n_bins);
    BBK_CHECK_LAUNCH("band_ingest_kernel");
    return BBK_OK;
}

extern "C" int bbk_contact_band_normalize(const int32_t* d_bin1, const int32_t* d_bin2, const double* d_value, int64_t n,
                                          const double* d_kr_norm, const double* d_kr_expected, int32_t n_bins, double* d_out,
                                          int32_t* d_bad, void* stream) {
    BBK_REQUIRE(n >= 0 && n_bins >= 0 && d_bad, "bbk_contact_band_normalize: bad arguments");
    BBK_CHECK_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int32_t), (cudaStream_t)stream));
    if (n == 0) return BBK_OK;
    BBK_REQUIRE(d_bin1 && d_bin2 && d_value && d_out && d_kr_norm && d_kr_expected, "bbk_contact_band_normalize: null pointer");
    long long want = (n + CM_THREADS - 1) / CM_THREADS;
    int grid = (int)(want < (long long)bbk_num_sms() * 8 ? want : (long long)bbk_num_sms() * 8);
    band_normalize_kernel<<<grid, CM_THREADS, 0, (cudaStream_t)stream>>>(d_bin1, d_bin2, d_value, n, d_kr_norm, d_kr_expected, n_bins, d_out, d_bad);
    BBK_CHECK_LAUNCH("band_normalize_kernel");
    return BBK_OK;
}
