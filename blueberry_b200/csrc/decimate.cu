// decimate.cu - K7: FithicContactMap.decimate (datatypes.pyx:317-339) as a sort / reduce-by-key.
//
// The reference rounds both midpoints of every row to the coarser grid,
//     mid' = (int(mid) + r) / r * r - r/2          (Python-2 integer division: floor)
// and folds the rows that now share (mid1', mid2') in file order:
//     count = count_i + count,  p = p_i * p,  q = min(q_i, q)    starting from (0, 1, 1)      (:333-335)
// The floating-point sum and product are order-dependent, so a group is folded by ONE thread, member by member in
// file order: a stable radix sort by key (bin1' << 32 | bin2', the coarse bin numbers) brings a group's rows together in file order, each
// group head is sent, keyed by the file position of its first row, through a second sort, and the thread that
// owns output row k folds the k-th group to appear in the file (the order of a Python-3 dict; Python 2's was arbitrary).
#include "common.cuh"

namespace {

constexpr int DC_THREADS = 256;

__device__ __forceinline__ long long floordiv_ll(long long a, long long b) {      // b > 0
    long long q = a / b;
    return (a % b < 0) ? q - 1 : q;
}

__global__ void __launch_bounds__(DC_THREADS) dec_keys_kernel(const double* map, long long n, long long r,
                                                              unsigned long long* keys, unsigned* idx, unsigned long long* counters,
                                                              int* bad) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[0] = 0;                    // group counter of dec_heads_kernel
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double a = map[5 * i], b = map[5 * i + 1];
        // .astype('int') truncates toward zero; coordinates must be representable (and leave room for + r)
        if (!(a > -1.0 && a < 2147483647.0 && b > -1.0 && b < 2147483647.0)) { *bad = 1; keys[i] = 0; idx[i] = (unsigned)i; continue; }
        const long long ia = (long long)a, ib = (long long)b;
        // the key holds the coarse BIN numbers ka, kb (mid' = k r - r/2), not the midpoints: fewer significant bytes, and the
        // radix sort skips the passes whose digit is the same in every key (chr1 at 5 kb: 4 passes instead of 6)
        const long long ka = floordiv_ll(ia + r, r), kb = floordiv_ll(ib + r, r);
        const long long ra = ka * r - r / 2, rb = kb * r - r / 2;
        if (ra < 0 || ra > 0xffffffffll || rb < 0 || rb > 0xffffffffll) { *bad = 1; keys[i] = 0; idx[i] = (unsigned)i; continue; }
        keys[i] = ((unsigned long long)ka << 32) | (unsigned long long)kb;
        idx[i] = (unsigned)i;
    }
}

// heads of the runs of equal keys -> (file position of the run's first row, sorted position of the head), appended to the
// second sort's input
__global__ void __launch_bounds__(DC_THREADS) dec_heads_kernel(const unsigned long long* keys, const unsigned* idx, long long n,
                                                               unsigned long long* hkeys, unsigned* hpos, unsigned long long* counters) {
    // one global atomic per CTA step (a warp-aggregated atomic per warp was 1.5e6 same-address atomics on 5e7 rows: 1 ms)
    __shared__ unsigned s_warp[DC_THREADS / 32];
    __shared__ unsigned long long s_base;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n_iter = (n + stride - 1) / stride;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long it = 0; it < n_iter; ++it, i += stride) {
        const bool head = i < n && (i == 0 || keys[i - 1] != keys[i]);
        const unsigned vote = __ballot_sync(0xffffffffu, head);
        if (lane == 0) s_warp[warp] = __popc(vote);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned tot = 0;
            for (int w = 0; w < DC_THREADS / 32; ++w) { unsigned v = s_warp[w]; s_warp[w] = tot; tot += v; }
            s_base = tot ? atomicAdd(&counters[0], (unsigned long long)tot) : 0ull;
        }
        __syncthreads();
        if (head) {
            const unsigned long long pos = s_base + s_warp[warp] + __popc(vote & ((1u << lane) - 1));
            hkeys[pos] = (unsigned long long)idx[i];       // the sort is stable: the head is the run's first row in the file
            hpos[pos] = (unsigned)i;
        }
        __syncthreads();
    }
}

// output row k = the k-th group to appear in the file, folded member by member in file order
__global__ void __launch_bounds__(DC_THREADS) dec_fold_kernel(const unsigned long long* keys, const unsigned* idx, long long n,
                                                              const unsigned* hpos, const unsigned long long* counters,
                                                              const double* map, double* out, long long* n_out, long long r) {
    const long long groups = (long long)counters[0];
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_out = groups;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < groups; k += stride) {
        long long j = hpos[k];
        const unsigned long long key = keys[j];
        double c = 0.0, pp = 1.0, qq = 1.0;
        for (; j < n && keys[j] == key; ++j) {
            const double* row = map + 5ll * idx[j];             // count, p, q of a row sit together: one sector or two per member
            c = row[2] + c;                                     // contactCount + contact0      (:335)
            pp = row[3] * pp;                                   // p * p0
            const double qv = row[4];
            qq = (qq < qv) ? qq : qv;                           // Python's min(q, q0): q0 only when q0 < q
        }
        double* o = out + 5 * k;
        o[0] = (double)((long long)(key >> 32) * r - r / 2);
        o[1] = (double)((long long)(key & 0xffffffffull) * r - r / 2);
        o[2] = c; o[3] = pp; o[4] = qq;
    }
}

__global__ void dec_flag_kernel(const int* bad, long long* n_out) { if (*bad) *n_out = -1; }

}  // namespace

extern "C" size_t bbk_decimate_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    return 2 * bbk_bh_workspace_bytes(n) + 256;          // one sort workspace per sort (the second must not scratch the first's result)
}

extern "C" int bbk_decimate(const double* d_map, int64_t n, int64_t resolution, double* d_out_map, int64_t* d_n_out,
                            void* d_workspace, size_t workspace_bytes, void* stream) {
    BBK_REQUIRE(n >= 0 && n < (1ll << 32), "bbk_decimate: n must be in [0, 2^32)");
    BBK_REQUIRE(resolution > 0 && resolution <= (1ll << 30), "bbk_decimate: resolution must be in [1, 2^30]");
    BBK_REQUIRE(d_n_out && d_workspace, "bbk_decimate: null pointer");
    BBK_REQUIRE(((uintptr_t)d_workspace & 255) == 0, "bbk_decimate: workspace must be 256-byte aligned");
    BBK_REQUIRE(n == 0 || (d_map && d_out_map), "bbk_decimate: null map");
    if (workspace_bytes < bbk_decimate_workspace_bytes(n)) {
        bbk_set_error("bbk_decimate: workspace too small (%zu < %zu bytes)", workspace_bytes, bbk_decimate_workspace_bytes(n));
        return BBK_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        BBK_CHECK_CUDA(cudaMemsetAsync(d_n_out, 0, sizeof(int64_t), st));
        return BBK_OK;
    }
    unsigned long long *k0, *k1;
    unsigned *i0, *i1;
    void* ws_rows = d_workspace;                                             // sort 1: the rows
    void* ws_heads = (char*)d_workspace + bbk_bh_workspace_bytes(n);         // sort 2: the group heads
    bbk_sort_buffers(ws_rows, n, 0, &k0, &i0);
    bbk_sort_buffers(ws_heads, n, 0, &k1, &i1);
    unsigned long long* counters = bbk_sort_count_ptr(ws_heads);             // [0] = number of groups (device)
    int* bad = (int*)((char*)d_workspace + 2 * bbk_bh_workspace_bytes(n));   // the 256 spare bytes
    BBK_CHECK_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
    long long want = (n + DC_THREADS - 1) / DC_THREADS;
    int grid = (int)(want < (long long)bbk_num_sms() * 8 ? want : (long long)bbk_num_sms() * 8);
    dec_keys_kernel<<<grid, DC_THREADS, 0, st>>>(d_map, n, resolution, k0, i0, counters, bad);
    BBK_CHECK_LAUNCH("dec_keys_kernel");
    int rc = bbk_sort_pairs(ws_rows, n, n, 0, st);                           // rows by (mid1', mid2'), stable
    if (rc != BBK_OK) return rc;
    dec_heads_kernel<<<grid, DC_THREADS, 0, st>>>(k0, i0, n, k1, i1, counters);
    BBK_CHECK_LAUNCH("dec_heads_kernel");
    rc = bbk_sort_pairs(ws_heads, n, -1, 0, st);                             // groups by first appearance (count on the device)
    if (rc != BBK_OK) return rc;
    dec_fold_kernel<<<grid, DC_THREADS, 0, st>>>(k0, i0, n, i1, counters, d_map, d_out_map, (long long*)d_n_out, resolution);
    BBK_CHECK_LAUNCH("dec_fold_kernel");
    dec_flag_kernel<<<1, 1, 0, st>>>(bad, (long long*)d_n_out);              // coordinates out of range: *d_n_out = -1
    BBK_CHECK_LAUNCH("dec_flag_kernel");
    return BBK_OK;
}
