// extract.cu - utils.extract_contacts (utils.py:31-90) on the device: the gather in front of the genome-wide q-value step.
//
// A Fit-Hi-C result table is (n, 5) float64 rows (mid1, mid2, contactCount, p, q) - FithicContactMap.map,
// datatypes.pyx:314.  The reference keeps the rows with p <= alpha (:72-73), shifts the columns right and puts the
// chromosome first (:76-77: chromosome, mid1, mid2, contactCount, p), and keeps the rows whose distance mid2 - mid1 lies
// in [LOW_FITHIC_CUTOFF, HIGH_FITHIC_CUTOFF] (:80-83).  Row order is kept, so this is an order-preserving stream compaction:
// per-chunk counts, one scan over the chunk counts, then the rows are written to their places.  40 B/row in, 40 B per kept row out.
#include "common.cuh"

namespace {

constexpr int EX_THREADS = 256;
constexpr int EX_CHUNK = 1024;          // rows per chunk: four per thread

struct ExParams {
    const double* map; long long n;
    double chromosome, alpha, low, high;
    int use_alpha;
    double* out; long long cap;
    unsigned* counts; unsigned long long* offsets; long long n_chunks;
    long long* n_out;
};

__device__ __forceinline__ bool ex_keep(const ExParams& P, long long r) {
    const double* row = P.map + 5 * r;
    const double pv = row[3];
    if (P.use_alpha && !(pv <= P.alpha)) return false;                       // utils.py:72-73 (NaN fails, as in numpy)
    const double d = row[1] - row[0];                                        // :80
    return d <= P.high && d >= P.low;                                        // :83
}

__global__ void __launch_bounds__(EX_THREADS) ex_count_kernel(ExParams P) {
    __shared__ unsigned s_w[EX_THREADS / 32];
    for (long long ch = blockIdx.x; ch < P.n_chunks; ch += gridDim.x) {
        unsigned c = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long r = ch * EX_CHUNK + (long long)threadIdx.x * 4 + j;
            c += r < P.n && ex_keep(P, r);
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned t = 0;
            for (int w = 0; w < EX_THREADS / 32; ++w) t += s_w[w];
            P.counts[ch] = t;
        }
        __syncthreads();
    }
}

// one CTA: exclusive prefix over the chunk counts
__global__ void __launch_bounds__(1024) ex_scan_kernel(ExParams P) {
    __shared__ unsigned long long s_part[1024];
    const int t = threadIdx.x;
    const long long per = (P.n_chunks + 1023) / 1024;
    const long long lo = (long long)t * per, hi = lo + per < P.n_chunks ? lo + per : P.n_chunks;
    unsigned long long sum = 0;
    for (long long i = lo; i < hi; ++i) sum += P.counts[i];
    s_part[t] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned long long v = t >= o ? s_part[t - o] : 0ull;
        __syncthreads();
        s_part[t] += v;
        __syncthreads();
    }
    unsigned long long run = s_part[t] - sum;
    for (long long i = lo; i < hi; ++i) { P.offsets[i] = run; run += P.counts[i]; }
    if (t == 1023) *P.n_out = (long long)s_part[1023];
}

__global__ void __launch_bounds__(EX_THREADS) ex_write_kernel(ExParams P) {
    __shared__ unsigned s_w[EX_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long ch = blockIdx.x; ch < P.n_chunks; ch += gridDim.x) {
        bool keep[4];
        unsigned c = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long r = ch * EX_CHUNK + (long long)threadIdx.x * 4 + j;
            keep[j] = r < P.n && ex_keep(P, r);
            c += keep[j];
        }
        unsigned inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        unsigned woff = 0;
        for (int w = 0; w < EX_THREADS / 32; ++w) woff += w < warp ? s_w[w] : 0u;
        unsigned long long at = P.offsets[ch] + woff + inc - c;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (keep[j]) {
                const long long r = ch * EX_CHUNK + (long long)threadIdx.x * 4 + j;
                if ((long long)at < P.cap) {
                    const double* row = P.map + 5 * r;
                    double* o = P.out + 5 * at;
                    o[0] = P.chromosome; o[1] = row[0]; o[2] = row[1]; o[3] = row[2]; o[4] = row[3];      // utils.py:76-77
                }
                at += 1;
            }
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" size_t bbk_extract_workspace_bytes(int64_t n) {
    const long long ch = n <= 0 ? 1 : (n + EX_CHUNK - 1) / EX_CHUNK;
    return (size_t)ch * 16 + 256;
}

extern "C" int bbk_extract_contacts(const double* d_map, int64_t n, double chromosome, double alpha, int32_t use_alpha, double low,
                                    double high, double* d_out, int64_t capacity, int64_t* d_n_out, void* d_workspace,
                                    size_t workspace_bytes, void* stream) {
    BBK_REQUIRE(n >= 0 && capacity >= 0, "bbk_extract_contacts: negative size");
    BBK_REQUIRE(d_n_out, "bbk_extract_contacts: null count");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { BBK_CHECK_CUDA(cudaMemsetAsync(d_n_out, 0, sizeof(int64_t), st)); return BBK_OK; }
    BBK_REQUIRE(d_map && d_workspace && (capacity == 0 || d_out), "bbk_extract_contacts: null pointer");
    if (workspace_bytes < bbk_extract_workspace_bytes(n)) {
        bbk_set_error("bbk_extract_contacts: workspace too small (%zu < %zu bytes)", workspace_bytes, bbk_extract_workspace_bytes(n));
        return BBK_E_WORKSPACE;
    }
    ExParams P = {};
    P.map = d_map; P.n = n; P.chromosome = chromosome; P.alpha = alpha; P.use_alpha = use_alpha ? 1 : 0; P.low = low; P.high = high;
    P.out = d_out; P.cap = capacity;
    P.n_chunks = (n + EX_CHUNK - 1) / EX_CHUNK;
    P.offsets = (unsigned long long*)d_workspace;
    P.counts = (unsigned*)((char*)d_workspace + (size_t)P.n_chunks * 8);
    P.n_out = (long long*)d_n_out;
    long long grid = (long long)bbk_num_sms() * 8;
    if (P.n_chunks < grid) grid = P.n_chunks;
    ex_count_kernel<<<(unsigned)grid, EX_THREADS, 0, st>>>(P);
    BBK_CHECK_LAUNCH("ex_count_kernel");
    ex_scan_kernel<<<1, 1024, 0, st>>>(P);
    BBK_CHECK_LAUNCH("ex_scan_kernel");
    ex_write_kernel<<<(unsigned)grid, EX_THREADS, 0, st>>>(P);
    BBK_CHECK_LAUNCH("ex_write_kernel");
    return BBK_OK;
}
