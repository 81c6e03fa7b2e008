// hist.cu - K1: contact histogram by genomic distance (reference: read_interactions, fithic.py:229-270).
//
// HBM-bound integer kernel: 12 B/pair in (mid1, mid2, count as int32 SoA), O(D) out.
//   * the usual case (no chromosome columns, distances below 2^31, a table of a few thousand keys) takes its records through
//     shared memory: every warp owns two stages of 256 records x 3 columns filled by bulk asynchronous copies (cp.async.bulk +
//     mbarrier, the TMA unit), so a warp has 6 KB in flight without holding a register and refills a stage as soon as its 8
//     records per lane are in registers; the other cases use 128-bit streaming loads (L1::no_allocate), two groups of 4
//     records in flight per thread;
//   * per-CTA shared-memory histogram (u32, flushed to the global int64 table before it can
//     overflow); warp-aggregated: when every active lane of a warp hits the same distance
//     (diagonal-major input) the counts are summed with REDUX and one lane issues the atomic,
//     otherwise lanes issue their own shared-memory atomics (row-major input touches 32
//     consecutive bins - the histogram is stored permuted so those land in 32 distinct banks);
//   * zero counts never touch the histogram (most records at 1 kb);
//   * scalar totals are carried in registers and reduced warp -> CTA -> one global atomic each;
//   * persistent grid: a multiple of the SM count.
#include "common.cuh"

namespace {

constexpr int HIST_THREADS = 512;
constexpr int SMALL_COUNT_LIMIT = 4096;          // counts below this go through the u32 shared histogram
constexpr long long FLUSH_PAIRS = 1ll << 20;     // 2^20 pairs * 4095 < 2^32: flush before a bin can overflow
constexpr int MAX_SMEM_KEYS = 40960;             // 160 KB of u32 bins

constexpr int HIST_WROWS = 256;                  // staged path: records of a warp tile (two int4 groups per lane)
constexpr int HIST_STAGE_INTS = 3 * HIST_WROWS;  // mid1 | mid2 | count
constexpr size_t HIST_STAGE_BYTES = (size_t)(HIST_THREADS / 32) * 2 * HIST_STAGE_INTS * sizeof(int);   // per CTA: 16 warps x 2 stages x 3 KB
static_assert((HIST_THREADS / 32) * HIST_WROWS == 2 * HIST_THREADS * 4, "a CTA tile (2 groups of 4 records per thread) is one warp tile per warp");

__device__ __forceinline__ unsigned h_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void h_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(h_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void h_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(h_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool h_mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(h_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// global -> shared bulk copy (TMA unit, no registers), completion counted in bytes on the mbarrier
__device__ __forceinline__ void h_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(h_smem_u32(dst)), "l"(src), "r"(bytes), "r"(h_smem_u32(bar)) : "memory");
}

struct HistParams {
    const int32_t* chr1;
    const int32_t* chr2;
    const int32_t* mid1;
    const int32_t* mid2;
    const int32_t* count;
    const double* p_excl;     // optional: records with p_excl[i] <= p_thr are skipped (second-pass outlier removal)
    double p_thr;
    long long n_pairs;
    long long min_dist, max_dist;
    long long lo_excl, hi_incl;   // in range  <=>  lo_excl < d <= hi_incl   (the -1 sentinels folded in)
    unsigned lo_u, span_u;        // FAST path: in range  <=>  m2 >= m1  &&  (u32)(m2 - m1) - lo_u <= span_u
    FastDiv div;
    int nkeys;        // len(mainDic)
    int nkeys_s;      // bins kept in shared memory (prefix of the table), 0 = none
    int quarter;      // permuted layout: phys(k) = (k >> 2) + (k & 3) * quarter
    long long* obs_sum;
    long long* totals;
};

struct Acc {
    long long S, in_range, intra_sum, intra_cnt, inter_sum, inter_cnt, dmin, dmax;
};

__device__ __forceinline__ int phys_bin(int k, int quarter) { return (k >> 2) + (k & 3) * quarter; }

template <bool HAS_CHR>
__device__ __forceinline__ void process_record(const HistParams& P, unsigned* sh, int m1, int m2, int c, int c1, int c2,
                                               bool live, Acc& a) {
    long long d = (long long)m2 - (long long)m1;                      // fithic.py:247 (no abs, no chromosome test)
    bool lo = (P.min_dist == -1) || (P.min_dist > -1 && d > P.min_dist);   // fithic.py:256
    bool hi = (P.max_dist == -1) || (P.max_dist > -1 && d <= P.max_dist);  // fithic.py:257
    bool in_range = live && lo && hi;
    if (HAS_CHR) {
        bool inter = live && (c1 != c2);                               // fithic.py:249-254
        bool intra = live && (c1 == c2);
        a.inter_sum += inter ? c : 0;
        a.inter_cnt += inter ? 1 : 0;
        a.intra_sum += intra ? c : 0;
        a.intra_cnt += intra ? 1 : 0;
    } else {
        a.intra_sum += live ? c : 0;
        a.intra_cnt += live ? 1 : 0;
    }
    if (in_range) {
        a.dmin = d < a.dmin ? d : a.dmin;                              // fithic.py:258-259
        a.dmax = d > a.dmax ? d : a.dmax;
        a.S += c;                                                      // fithic.py:262 (even when d is not a key)
        a.in_range += 1;                                               // fithic.py:263
    }
    // "if distance in mainDic" (fithic.py:260): d >= 0, a multiple of R, below the table end
    bool key_ok = false;
    unsigned k = 0;
    if (in_range && d >= 0 && c != 0) {
        unsigned ud = (unsigned)d;
        k = fastdiv(ud, P.div);
        key_ok = (k * P.div.R == ud) && (k < (unsigned)P.nkeys);
    }
    bool small = key_ok && (unsigned)c < (unsigned)SMALL_COUNT_LIMIT && k < (unsigned)P.nkeys_s;
    bool big = key_ok && !small;
    if (big) atomicAdd((unsigned long long*)&P.obs_sum[k], (unsigned long long)(long long)c);
    unsigned smask = __ballot_sync(0xffffffffu, small);
    if (smask == 0) return;
    int leader = __ffs(smask) - 1;
    unsigned k0 = __shfl_sync(0xffffffffu, k, leader);
    bool uniform = __all_sync(0xffffffffu, !small || k == k0);
    if (uniform) {
        unsigned tot = __reduce_add_sync(0xffffffffu, small ? (unsigned)c : 0u);
        if ((int)(threadIdx.x & 31) == leader) atomicAdd(&sh[phys_bin((int)k0, P.quarter)], tot);
    } else if (small) {
        atomicAdd(&sh[phys_bin((int)k, P.quarter)], (unsigned)c);
    }
}

// FAST path (0 <= d <= max_dist < 2^31 for every in-range record): one int4 group = 4
// consecutive records of a lane.  32-bit arithmetic throughout: with m2 >= m1 (signed) the difference is in [0, 2^32) and
// equal to the wrapped 32-bit subtraction, so "lo_excl < d <= hi_incl" is one unsigned span test; without it d < 0 is out of
// range anyway.  The warp-uniform test (diagonal-major input: all 128 records of the warp step on one distance) is made once
// per group, otherwise plain shared atomics (ptxas turns a predicated red.shared back into a branch, so there is no
// cheaper form of those).
struct FastAcc {
    long long S, intra_sum, inter_sum;
    int in_range, intra_cnt, inter_cnt, dmin, dmax;
};

template <bool FULL, bool HAS_CHR, bool EXCL = true>
__device__ __forceinline__ void process_group_fast(const HistParams& P, unsigned* sh, int4 m1, int4 m2, int4 c, int4 x1, int4 x2,
                                                   bool live_group, unsigned excl, FastAcc& a) {
    const int m1s[4] = {m1.x, m1.y, m1.z, m1.w}, m2s[4] = {m2.x, m2.y, m2.z, m2.w}, cs[4] = {c.x, c.y, c.z, c.w};
    const int x1s[4] = {x1.x, x1.y, x1.z, x1.w}, x2s[4] = {x2.x, x2.y, x2.z, x2.w};
    unsigned key[4];
    bool ok[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const bool live = (FULL || live_group) && (!EXCL || !((excl >> e) & 1u));   // (EXCL = false: the caller has no exclusion mask)
        const unsigned ud = (unsigned)m2s[e] - (unsigned)m1s[e];                      // fithic.py:247
        const bool in_range = live && m2s[e] >= m1s[e] && (ud - P.lo_u) <= P.span_u;  // fithic.py:256-257
        const int d = (int)ud;                                                        // < 2^31 when in_range
        if (HAS_CHR) {                                                                // fithic.py:249-254
            const bool inter = live && x1s[e] != x2s[e], intra = live && x1s[e] == x2s[e];
            a.inter_sum += inter ? cs[e] : 0;
            a.inter_cnt += inter ? 1 : 0;
            a.intra_sum += intra ? cs[e] : 0;
            a.intra_cnt += intra ? 1 : 0;
        } else {
            a.intra_sum += live ? cs[e] : 0;
            a.intra_cnt += live ? 1 : 0;
        }
        if (in_range) {
            a.dmin = min(a.dmin, d);                                                  // fithic.py:258-259
            a.dmax = max(a.dmax, d);
            a.S += cs[e];                                                             // fithic.py:262
            a.in_range += 1;                                                          // fithic.py:263
        }
        const unsigned k = fastdiv31(ud, P.div);
        key[e] = k;
        ok[e] = in_range && cs[e] != 0 && k * P.div.R == ud && k < (unsigned)P.nkeys;  // "distance in mainDic" (:260)
    }
    // big counts / keys beyond the shared table go straight to the global table
    bool small[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        small[e] = ok[e] && (unsigned)cs[e] < (unsigned)SMALL_COUNT_LIMIT && key[e] < (unsigned)P.nkeys_s;
        if (ok[e] && !small[e]) atomicAdd((unsigned long long*)&P.obs_sum[key[e]], (unsigned long long)(long long)cs[e]);
    }
    // warp-uniform distance?  (diagonal-major input: every lane's 4 keys equal lane 0's first key, all of them "small" or
    // zero-count.)  Row-major input fails the first, cheap test - a lane's first and last key against lane 0's - at once.
    const unsigned kref = __shfl_sync(0xffffffffu, key[0], 0);
    bool uniform = __all_sync(0xffffffffu, key[0] == kref && key[3] == kref);
    int ssum = 0;
    if (uniform) {
        bool mine_uniform = true;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            mine_uniform = mine_uniform && (!ok[e] || (small[e] && key[e] == kref));
            ssum += small[e] ? cs[e] : 0;
        }
        uniform = __all_sync(0xffffffffu, mine_uniform);
    }
    if (uniform) {
        unsigned tot = __reduce_add_sync(0xffffffffu, (unsigned)ssum);
        if ((threadIdx.x & 31) == 0 && tot) atomicAdd(&sh[phys_bin((int)kref, P.quarter)], tot);
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (small[e]) atomicAdd(&sh[phys_bin((int)key[e], P.quarter)], (unsigned)cs[e]);
    }
}

__device__ __forceinline__ void flush_hist(const HistParams& P, unsigned* sh) {
    __syncthreads();
    for (int k = threadIdx.x; k < P.nkeys_s; k += blockDim.x) {
        int ph = phys_bin(k, P.quarter);
        unsigned v = sh[ph];
        if (v) {
            atomicAdd((unsigned long long*)&P.obs_sum[k], (unsigned long long)v);
            sh[ph] = 0;
        }
    }
    __syncthreads();
}

// the records of a warp's part of CTA tile `tile` into its stage `stage` (one lane)
__device__ __forceinline__ void hist_issue_tile(const HistParams& P, int* wbuf, unsigned long long* bars, long long tile, int warp, int stage) {
    const long long r0 = tile * (2ll * HIST_THREADS * 4) + (long long)warp * HIST_WROWS;
    int* dst = wbuf + stage * HIST_STAGE_INTS;
    const unsigned bytes = HIST_WROWS * sizeof(int);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // the stage's last readers (generic proxy) are done
    h_mbar_expect_tx(&bars[stage], 3u * bytes);
    h_bulk_g2s(dst, P.mid1 + r0, bytes, &bars[stage]);
    h_bulk_g2s(dst + HIST_WROWS, P.mid2 + r0, bytes, &bars[stage]);
    h_bulk_g2s(dst + 2 * HIST_WROWS, P.count + r0, bytes, &bars[stage]);
}

template <bool HAS_CHR, bool FAST, bool STAGED>
__global__ void __launch_bounds__(HIST_THREADS, HAS_CHR ? 1 : 2) hist_pairs_kernel(HistParams P) {
    static_assert(!STAGED || (FAST && !HAS_CHR), "the staged path is the fast path without chromosome columns");
    extern __shared__ __align__(128) unsigned sh[];
    __shared__ long long red[8][HIST_THREADS / 32];
    __shared__ unsigned long long bars[HIST_THREADS / 32][2];              // staged path: a warp's two "stage has landed" barriers
    for (int i = threadIdx.x; i < 4 * P.quarter; i += blockDim.x) sh[i] = 0;
    if (STAGED && (threadIdx.x & 31) == 0) {
        h_mbar_init(&bars[threadIdx.x >> 5][0], 1); h_mbar_init(&bars[threadIdx.x >> 5][1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    Acc a = {0, 0, 0, 0, 0, 0, 500000000ll, 0ll};                      // fithic.py:40-41 initial min / max
    FastAcc fa = {0, 0, 0, 0, 0, 0, 500000000, 0};
    const long long n_groups = P.n_pairs >> 2;                         // groups of 4 records (one int4 per column)
    const int4* m1v = reinterpret_cast<const int4*>(P.mid1);
    const int4* m2v = reinterpret_cast<const int4*>(P.mid2);
    const int4* cv = reinterpret_cast<const int4*>(P.count);
    const int4* c1v = reinterpret_cast<const int4*>(P.chr1);
    const int4* c2v = reinterpret_cast<const int4*>(P.chr2);
    long long since_flush = 0;
    // CTA tiles of 2*blockDim groups: the trip count is uniform across the CTA, so the barrier in
    // flush_hist and the warp collectives in process_record are reached by every thread
    const long long tile_groups = 2ll * blockDim.x;
    const long long n_tiles = (n_groups + tile_groups - 1) / tile_groups;
    long long tile = blockIdx.x;
    if (STAGED) {
        // full tiles through the warp's two stages: wait for the stage, take the 8 records of the lane into registers, refill
        // the stage with the tile after next, then do the arithmetic
        const long long full_tiles = n_groups / tile_groups;
        const double2* pe = reinterpret_cast<const double2*>(P.p_excl);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        int* wbuf = reinterpret_cast<int*>(sh + 4 * P.quarter) + warp * (2 * HIST_STAGE_INTS);    // 16 * quarter bytes: 16-byte aligned
        if (lane == 0) {
            if (tile < full_tiles) hist_issue_tile(P, wbuf, bars[warp], tile, warp, 0);
            if (tile + gridDim.x < full_tiles) hist_issue_tile(P, wbuf, bars[warp], tile + gridDim.x, warp, 1);
        }
        for (int k = 0; tile < full_tiles; tile += gridDim.x, ++k) {
            const int stage = k & 1;
            const unsigned parity = (unsigned)(k >> 1) & 1u;
            while (!h_mbar_try_wait(&bars[warp][stage], parity)) { }
            const int4* sv = reinterpret_cast<const int4*>(wbuf + stage * HIST_STAGE_INTS);
            const int4 a1 = sv[lane], a2 = sv[HIST_WROWS / 4 + lane], ac = sv[2 * (HIST_WROWS / 4) + lane];
            const int4 b1 = sv[32 + lane], b2 = sv[HIST_WROWS / 4 + 32 + lane], bc = sv[2 * (HIST_WROWS / 4) + 32 + lane];
            __syncwarp();
            if (lane == 0 && tile + 2ll * gridDim.x < full_tiles) hist_issue_tile(P, wbuf, bars[warp], tile + 2ll * gridDim.x, warp, stage);
            const int4 z = make_int4(0, 0, 0, 0);
            if (pe) {                                  // (uniform: the refit pass)
                const long long g0 = tile * tile_groups + (long long)warp * (HIST_WROWS / 4) + lane, g1 = g0 + 32;
                const double2 u0 = ld_stream_double2(pe + 2 * g0), v0 = ld_stream_double2(pe + 2 * g0 + 1);
                const double2 u1 = ld_stream_double2(pe + 2 * g1), v1 = ld_stream_double2(pe + 2 * g1 + 1);
                const unsigned xa = (u0.x <= P.p_thr ? 1u : 0u) | (u0.y <= P.p_thr ? 2u : 0u) | (v0.x <= P.p_thr ? 4u : 0u) | (v0.y <= P.p_thr ? 8u : 0u);
                const unsigned xb = (u1.x <= P.p_thr ? 1u : 0u) | (u1.y <= P.p_thr ? 2u : 0u) | (v1.x <= P.p_thr ? 4u : 0u) | (v1.y <= P.p_thr ? 8u : 0u);
                process_group_fast<true, false, true>(P, sh, a1, a2, ac, z, z, true, xa, fa);
                process_group_fast<true, false, true>(P, sh, b1, b2, bc, z, z, true, xb, fa);
            } else {
                process_group_fast<true, false, false>(P, sh, a1, a2, ac, z, z, true, 0u, fa);
                process_group_fast<true, false, false>(P, sh, b1, b2, bc, z, z, true, 0u, fa);
            }
            since_flush += 4 * tile_groups;
            if (since_flush >= FLUSH_PAIRS) {      // uniform across the CTA
                flush_hist(P, sh);
                since_flush = 0;
            }
        }
    } else if (FAST) {
        // full tiles: no bounds predicates, the loads of a thread go out back to back
        const long long full_tiles = n_groups / tile_groups;
        const double2* pe = reinterpret_cast<const double2*>(P.p_excl);
        for (; tile < full_tiles; tile += gridDim.x) {
            const long long g0 = tile * tile_groups + threadIdx.x;
            const long long g1 = g0 + blockDim.x;
            const int4 a1 = ld_stream_int4(m1v + g0), a2 = ld_stream_int4(m2v + g0), ac = ld_stream_int4(cv + g0);
            const int4 b1 = ld_stream_int4(m1v + g1), b2 = ld_stream_int4(m2v + g1), bc = ld_stream_int4(cv + g1);
            int4 ax = make_int4(0, 0, 0, 0), ay = ax, bx = ax, by = ax;
            if (HAS_CHR) {
                ax = ld_stream_int4(c1v + g0); ay = ld_stream_int4(c2v + g0);
                bx = ld_stream_int4(c1v + g1); by = ld_stream_int4(c2v + g1);
            }
            unsigned xa = 0, xb = 0;
            if (pe) {
                const double2 u0 = ld_stream_double2(pe + 2 * g0), v0 = ld_stream_double2(pe + 2 * g0 + 1);
                const double2 u1 = ld_stream_double2(pe + 2 * g1), v1 = ld_stream_double2(pe + 2 * g1 + 1);
                xa = (u0.x <= P.p_thr ? 1u : 0u) | (u0.y <= P.p_thr ? 2u : 0u) | (v0.x <= P.p_thr ? 4u : 0u) | (v0.y <= P.p_thr ? 8u : 0u);
                xb = (u1.x <= P.p_thr ? 1u : 0u) | (u1.y <= P.p_thr ? 2u : 0u) | (v1.x <= P.p_thr ? 4u : 0u) | (v1.y <= P.p_thr ? 8u : 0u);
            }
            process_group_fast<true, HAS_CHR>(P, sh, a1, a2, ac, ax, ay, true, xa, fa);
            process_group_fast<true, HAS_CHR>(P, sh, b1, b2, bc, bx, by, true, xb, fa);
            since_flush += 4 * tile_groups;
            if (since_flush >= FLUSH_PAIRS) {      // uniform across the CTA
                flush_hist(P, sh);
                since_flush = 0;
            }
        }
        // at most one partial tile is left, and exactly one CTA arrives at it; it goes through the general loop below
    }
    for (; tile < n_tiles; tile += gridDim.x) {
        long long g0 = tile * tile_groups + threadIdx.x;
        long long g1 = g0 + blockDim.x;
        bool l0 = g0 < n_groups, l1 = g1 < n_groups;
        int4 z = make_int4(0, 0, 0, 0);
        int4 a1 = l0 ? ld_stream_int4(m1v + g0) : z, a2 = l0 ? ld_stream_int4(m2v + g0) : z, ac = l0 ? ld_stream_int4(cv + g0) : z;
        int4 b1 = l1 ? ld_stream_int4(m1v + g1) : z, b2 = l1 ? ld_stream_int4(m2v + g1) : z, bc = l1 ? ld_stream_int4(cv + g1) : z;
        int4 ax = z, ay = z, bx = z, by = z;
        if (HAS_CHR) {
            if (l0) { ax = ld_stream_int4(c1v + g0); ay = ld_stream_int4(c2v + g0); }
            if (l1) { bx = ld_stream_int4(c1v + g1); by = ld_stream_int4(c2v + g1); }
        }
        // outlier mask: bit e set = record e of the group is excluded
        unsigned xa = 0, xb = 0;
        if (P.p_excl) {
            const double2* pe = reinterpret_cast<const double2*>(P.p_excl);
            if (l0) { double2 u = ld_stream_double2(pe + 2 * g0), v = ld_stream_double2(pe + 2 * g0 + 1);
                      xa = (u.x <= P.p_thr ? 1u : 0u) | (u.y <= P.p_thr ? 2u : 0u) | (v.x <= P.p_thr ? 4u : 0u) | (v.y <= P.p_thr ? 8u : 0u); }
            if (l1) { double2 u = ld_stream_double2(pe + 2 * g1), v = ld_stream_double2(pe + 2 * g1 + 1);
                      xb = (u.x <= P.p_thr ? 1u : 0u) | (u.y <= P.p_thr ? 2u : 0u) | (v.x <= P.p_thr ? 4u : 0u) | (v.y <= P.p_thr ? 8u : 0u); }
        }
        if (FAST) {
            process_group_fast<false, HAS_CHR>(P, sh, a1, a2, ac, ax, ay, l0, xa, fa);
            process_group_fast<false, HAS_CHR>(P, sh, b1, b2, bc, bx, by, l1, xb, fa);
        } else {
            process_record<HAS_CHR>(P, sh, a1.x, a2.x, ac.x, ax.x, ay.x, l0 && !(xa & 1u), a);
            process_record<HAS_CHR>(P, sh, a1.y, a2.y, ac.y, ax.y, ay.y, l0 && !(xa & 2u), a);
            process_record<HAS_CHR>(P, sh, a1.z, a2.z, ac.z, ax.z, ay.z, l0 && !(xa & 4u), a);
            process_record<HAS_CHR>(P, sh, a1.w, a2.w, ac.w, ax.w, ay.w, l0 && !(xa & 8u), a);
            process_record<HAS_CHR>(P, sh, b1.x, b2.x, bc.x, bx.x, by.x, l1 && !(xb & 1u), a);
            process_record<HAS_CHR>(P, sh, b1.y, b2.y, bc.y, bx.y, by.y, l1 && !(xb & 2u), a);
            process_record<HAS_CHR>(P, sh, b1.z, b2.z, bc.z, bx.z, by.z, l1 && !(xb & 4u), a);
            process_record<HAS_CHR>(P, sh, b1.w, b2.w, bc.w, bx.w, by.w, l1 && !(xb & 8u), a);
        }
        since_flush += 4 * tile_groups;
        if (since_flush >= FLUSH_PAIRS) {      // uniform across the CTA
            flush_hist(P, sh);
            since_flush = 0;
        }
    }
    // tail: the last n_pairs % 4 records, by the first warp of block 0
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        long long i = (n_groups << 2) + threadIdx.x;
        bool live = threadIdx.x < (P.n_pairs & 3);
        if (live && P.p_excl && P.p_excl[i] <= P.p_thr) live = false;
        int m1 = live ? P.mid1[i] : 0, m2 = live ? P.mid2[i] : 0, c = live ? P.count[i] : 0;
        int c1 = 0, c2 = 0;
        if (HAS_CHR && live) { c1 = P.chr1[i]; c2 = P.chr2[i]; }
        process_record<HAS_CHR>(P, sh, m1, m2, c, c1, c2, live, a);
    }
    flush_hist(P, sh);
    if (FAST) {
        a.S += fa.S; a.in_range += fa.in_range; a.intra_sum += fa.intra_sum; a.intra_cnt += fa.intra_cnt;
        a.inter_sum += fa.inter_sum; a.inter_cnt += fa.inter_cnt;
        a.dmin = fa.dmin < a.dmin ? fa.dmin : a.dmin;
        a.dmax = fa.dmax > a.dmax ? fa.dmax : a.dmax;
    }

    // scalar totals: warp shuffle -> shared -> one global atomic per CTA and quantity
    long long v[8] = {a.S, a.in_range, a.intra_sum, a.intra_cnt, a.inter_sum, a.inter_cnt, a.dmin, a.dmax};
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        long long r = (q == 6) ? warp_min_ll(v[q]) : (q == 7) ? warp_max_ll(v[q]) : warp_sum_ll(v[q]);
        if (lane == 0) red[q][warp] = r;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        int q = threadIdx.x;
        long long r = red[q][0];
        for (int w = 1; w < HIST_THREADS / 32; ++w) {
            long long x = red[q][w];
            r = (q == 6) ? (x < r ? x : r) : (q == 7) ? (x > r ? x : r) : r + x;
        }
        if (q == 6) atomicMin(&P.totals[6], r);
        else if (q == 7) atomicMax(&P.totals[7], r);
        else if (r != 0) atomicAdd((unsigned long long*)&P.totals[q], (unsigned long long)r);
    }
}

__global__ void hist_init_kernel(long long* obs, int nkeys, long long* totals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nkeys) obs[i] = 0;
    if (i < 8) totals[i] = (i == 6) ? 500000000ll : 0ll;               // fithic.py:25-41
}

// one SUM all-reduce carries the whole of K1's output: min / max observed distance travel as one slot per rank
__global__ void stats_pack_kernel(const long long* totals, long long* ext, int world, int rank) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < world) { ext[i] = i == rank ? totals[6] : 0; ext[world + i] = i == rank ? totals[7] : 0; }
}
__global__ void stats_unpack_kernel(long long* totals, const long long* ext, int world) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    long long lo = ext[0], hi = ext[world];
    for (int r = 1; r < world; ++r) { lo = ext[r] < lo ? ext[r] : lo; hi = ext[world + r] > hi ? ext[world + r] : hi; }
    totals[6] = lo;                                                    // fithic.py:258-259 over all ranks
    totals[7] = hi;
}

}  // namespace

extern "C" int bbk_stats_pack(const int64_t* d_totals, int64_t* d_ext, int32_t world, int32_t rank, void* stream) {
    BBK_REQUIRE(d_totals && d_ext && world > 0 && rank >= 0 && rank < world, "bbk_stats_pack: bad arguments");
    stats_pack_kernel<<<(world + 63) / 64, 64, 0, (cudaStream_t)stream>>>((const long long*)d_totals, (long long*)d_ext, world, rank);
    BBK_CHECK_LAUNCH("stats_pack_kernel");
    return BBK_OK;
}

extern "C" int bbk_stats_unpack(int64_t* d_totals, const int64_t* d_ext, int32_t world, void* stream) {
    BBK_REQUIRE(d_totals && d_ext && world > 0, "bbk_stats_unpack: bad arguments");
    stats_unpack_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((long long*)d_totals, (const long long*)d_ext, world);
    BBK_CHECK_LAUNCH("stats_unpack_kernel");
    return BBK_OK;
}

extern "C" int bbk_hist_init(int64_t* d_obs_sum, int32_t nkeys, int64_t* d_totals, void* stream) {
    BBK_REQUIRE(d_obs_sum && d_totals && nkeys >= 0, "bbk_hist_init: null table or negative nkeys");
    int n = nkeys > 8 ? nkeys : 8;
    hist_init_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((long long*)d_obs_sum, nkeys, (long long*)d_totals);
    BBK_CHECK_LAUNCH("hist_init_kernel");
    return BBK_OK;
}

static int hist_pairs_impl(const int32_t* d_chr1, const int32_t* d_chr2, const int32_t* d_mid1, const int32_t* d_mid2,
                           const int32_t* d_count, const double* d_p_excl, double p_thr, int64_t n_pairs, int64_t resolution,
                           int64_t min_dist, int64_t max_dist, int32_t nkeys, int64_t* d_obs_sum, int64_t* d_totals, void* stream) {
    BBK_REQUIRE(n_pairs >= 0 && nkeys >= 0, "bbk_hist_pairs: negative size");
    BBK_REQUIRE(resolution > 0 && resolution < (1ll << 32), "bbk_hist_pairs: resolution must be in [1, 2^32)");
    BBK_REQUIRE((d_chr1 == nullptr) == (d_chr2 == nullptr), "bbk_hist_pairs: chr1/chr2 must both be given or both NULL");
    BBK_REQUIRE(d_obs_sum && d_totals, "bbk_hist_pairs: null output");
    if (n_pairs == 0) return BBK_OK;
    BBK_REQUIRE(d_mid1 && d_mid2 && d_count, "bbk_hist_pairs: null input column");
    uintptr_t align = (uintptr_t)d_mid1 | (uintptr_t)d_mid2 | (uintptr_t)d_count | (uintptr_t)d_chr1 | (uintptr_t)d_chr2 | (uintptr_t)d_p_excl;
    BBK_REQUIRE((align & 15) == 0, "bbk_hist_pairs: columns must be 16-byte aligned");

    HistParams P;
    P.chr1 = d_chr1; P.chr2 = d_chr2; P.mid1 = d_mid1; P.mid2 = d_mid2; P.count = d_count;
    P.p_excl = d_p_excl; P.p_thr = p_thr;
    P.n_pairs = n_pairs; P.min_dist = min_dist; P.max_dist = max_dist;
    P.div = make_fastdiv((uint64_t)resolution);
    P.lo_excl = (min_dist == -1) ? (-0x7fffffffffffffffll - 1) : (min_dist > -1 ? min_dist : 0x7fffffffffffffffll);
    P.hi_incl = (max_dist == -1) ? 0x7fffffffffffffffll : (max_dist > -1 ? max_dist : (-0x7fffffffffffffffll - 1));
    // 31-bit fast path: every in-range distance is in [0, 2^31)
    // (and a non-empty range, so that it is one unsigned span: lo_u <= d <= lo_u + span_u)
    const bool fast = min_dist >= -1 && max_dist >= 0 && max_dist < (1ll << 31) && P.lo_excl >= -1 && P.lo_excl < P.hi_incl;
    P.lo_u = fast ? (unsigned)(P.lo_excl + 1) : 0u;
    P.span_u = fast ? (unsigned)(P.hi_incl - (P.lo_excl + 1)) : 0u;
    P.nkeys = nkeys;
    // only keys that can be in range need a shared bin: k*R <= max_dist
    long long want = nkeys;
    if (max_dist > -1) { long long lim = max_dist / resolution + 1; if (lim < want) want = lim; }
    if (want > MAX_SMEM_KEYS) want = MAX_SMEM_KEYS;     // the rest goes straight to the global table
    P.nkeys_s = (int)want;
    P.quarter = ((P.nkeys_s + 3) / 4 + 31) / 32 * 32 + 1;   // odd multiple-of-32 offset keeps the 4 quarters on distinct banks
    if (P.nkeys_s == 0) P.quarter = 0;
    P.obs_sum = (long long*)d_obs_sum; P.totals = (long long*)d_totals;
    size_t smem = (size_t)4 * P.quarter * sizeof(unsigned);
    // records through shared memory (bulk copies) when two CTAs with their stages still fit an SM: 2 x (table + 96 KB + static + 1 KB) <= 228 KB
    const bool staged = fast && !d_chr1 && 2 * (smem + HIST_STAGE_BYTES + 4096) <= 228 * 1024;
    if (staged) smem += HIST_STAGE_BYTES;

    int sms = bbk_num_sms();
    long long groups = n_pairs >> 2;
    long long need = (groups + 2ll * HIST_THREADS - 1) / (2ll * HIST_THREADS);
    int per_sm = (!staged && smem > 100 * 1024) ? 1 : 2;
    long long grid = (long long)sms * per_sm;
    if (need < grid) grid = need > 0 ? need : 1;
    cudaStream_t st = (cudaStream_t)stream;
#define BBK_HIST_LAUNCH(C, F, G)                                                                                                    \
    do {                                                                                                                            \
        BBK_CHECK_CUDA(cudaFuncSetAttribute(hist_pairs_kernel<C, F, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        if (G) BBK_CHECK_CUDA(cudaFuncSetAttribute(hist_pairs_kernel<C, F, G>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)); \
        hist_pairs_kernel<C, F, G><<<(unsigned)grid, HIST_THREADS, smem, st>>>(P);                                                  \
    } while (0)
    if (d_chr1) { if (fast) BBK_HIST_LAUNCH(true, true, false); else BBK_HIST_LAUNCH(true, false, false); }
    else if (staged) BBK_HIST_LAUNCH(false, true, true);
    else        { if (fast) BBK_HIST_LAUNCH(false, true, false); else BBK_HIST_LAUNCH(false, false, false); }
#undef BBK_HIST_LAUNCH
    BBK_CHECK_LAUNCH("hist_pairs_kernel");
    return BBK_OK;
}

extern "C" int bbk_hist_pairs(const int32_t* d_chr1, const int32_t* d_chr2, const int32_t* d_mid1, const int32_t* d_mid2,
                              const int32_t* d_count, int64_t n_pairs, int64_t resolution, int64_t min_dist,
                              int64_t max_dist, int32_t nkeys, int64_t* d_obs_sum, int64_t* d_totals, void* stream) {
    return hist_pairs_impl(d_chr1, d_chr2, d_mid1, d_mid2, d_count, nullptr, 0.0, n_pairs, resolution, min_dist, max_dist,
                           nkeys, d_obs_sum, d_totals, stream);
}

extern "C" int bbk_hist_pairs_excluding(const int32_t* d_chr1, const int32_t* d_chr2, const int32_t* d_mid1, const int32_t* d_mid2,
                                        const int32_t* d_count, const double* d_p, double p_outlier, int64_t n_pairs,
                                        int64_t resolution, int64_t min_dist, int64_t max_dist, int32_t nkeys,
                                        int64_t* d_obs_sum, int64_t* d_totals, void* stream) {
    BBK_REQUIRE(d_p != nullptr || n_pairs == 0, "bbk_hist_pairs_excluding: null p");
    return hist_pairs_impl(d_chr1, d_chr2, d_mid1, d_mid2, d_count, d_p, p_outlier, n_pairs, resolution, min_dist, max_dist,
                           nkeys, d_obs_sum, d_totals, stream);
}
