// pvalue_tiles.inl - K4 as one streaming kernel over bulk-staged record tiles (included by pvalue.cu inside its
// anonymous namespace).  Reference: the scoring loop of fit_spline, fithic.py:413-435.
//
// What the records look like decides the design: most have count 0 (60 % at 5 kb, 93 % at 1 kb) and need no arithmetic,
// most of the rest have a count of 1..8, whose p-value is a handful of FP64 operations in the LOWER-tail form, and a few
// per cent (large counts next to the diagonal, the significant rows) need a tail sum of tens to hundreds of terms.  So:
//
//   score_tiles_kernel  streams every record once.  Every WARP runs on its own over tiles of 256 rows - no CTA barrier, no
//       shared state between warps.  The three record columns of a tile arrive in the warp's shared memory by bulk
//       asynchronous copies (cp.async.bulk + mbarrier, two stages: the warp's next two tiles are in flight while it works
//       on one, without holding a register).
//         decode   class of every row from its count and distance alone.  count <= 0: p = 1.0 unless one of the two loci
//                  carries a flag bit in the bias bit map (value < 0 or > 4: the prior has to be looked at) - one bit per
//                  locus instead of an 8-byte gather, which is what made the 1 kb workloads latency-bound.  The result of
//                  every such row (1.0 / NaN) replaces the row in the tile; rows that need arithmetic go on the warp's
//                  list, sorted count <= 1 | rest.
//         rounds   32 list entries at a time, every lane busy on the same path; the gathers of the NEXT round (spline value,
//                  two bias values) are in flight while a round computes.  prior = splineY[i] * (b1 * b2), fithic.py:429-431;
//                      count == 1     1 - (1-q)^S                                (bdtrc's closed form)
//                      2..LOW_C_MAX   1 - (1-q)^S (1 + r1 + r1 r2 + ...), count-1 terms (the lower tail; the terms written
//                                     out up to SMALL_C, a loop above)
//                  rows this cannot finish - larger counts, a prior outside (0, 2^-10), a lower-tail result below 1e-4
//                  (digits lost in the subtraction), tiny S - are DEFERRED: (row, count, prior) goes to a global list.
//         output   the warp writes its 256 rows of p (and q = 1.0 / NaN) with full coalesced 128-bit stores; no partial
//                  sector is ever written by this kernel.
//       It also fills the coarse p histogram and appends (key, row) of every p < BBK_SMALL_P to the q-value step's
//       candidate list; deferred rows and candidates collect in warp-private buffers and leave 32 or more at a time
//       (one global atomic per batch).
//   score_deferred_kernel  scores the deferred list with the tail-sum machinery of bbk_pvalues (bdtrc's case analysis,
//       upper-tail sum as a resumable state, saddle-point form outside the series' range) and patches p in place - a few
//       per cent of the rows, so the 8-byte scattered stores cost nothing next to the stream.
//   score_guard_kernel  (after the fit, before the two) checks what the count <= 0 shortcut assumes about the spline
//       (0 <= splineY, 16 max(splineY) <= 1); if that fails BbkScoreState.exact is raised and every in-range row goes
//       through its prior, so the result is exact in every case.

constexpr int ST_THREADS = 256;                    // (nine warps at 72 registers were measured: no gain, 0.99 vs 0.94 ms on cfg2)
constexpr int ST_WARPS = ST_THREADS / 32;
constexpr int ST_WROWS = 256;                      // rows of a warp tile: a lane owns two groups of four
constexpr int ST_CTAS_PER_SM = 3;
constexpr int ST_BUF = 64;                         // warp-private output buffers: flushed as soon as they hold 32
constexpr int SMALL_C = 8;                         // counts up to here: the lower tail with its factors written out
constexpr int LOW_C_MAX = 64;                      // counts up to here: the lower tail as a loop (count - 1 terms); above: deferred
constexpr double LOWER_MIN_P = 1e-4;               // below this the lower-tail form has lost digits: deferred
constexpr double LEAN_MAX_PRIOR = 9.765625e-4;     // 2^-10: ln(1-q) and q/(1-q) as short series
constexpr double BIAS_FLAG_MAX = 4.0;              // bias values in [0, 4] (and absent loci) carry no flag bit
constexpr int ST_HBASE = 959 * 2;                  // CTA-local histogram: buckets of p >= 2^-64, the rest goes straight to global
constexpr int ST_HBINS = 128;
constexpr int HI_ONE = 0x3ff00000, HI_NAN = 0x7ff80000;   // high words of 1.0 and of the one NaN this kernel writes
static_assert(ST_HBASE + ST_HBINS == 2046, "the local histogram ends with the bucket below 1.0");
static_assert(SMALL_C == 8, "the lower-tail factors are written out for counts up to 8");

struct StWarp {                                    // one warp's shared memory
    int m1[2][ST_WROWS];                           // mid1; after decode: low word of p (non-listed rows) / still mid1 (listed rows) / after its round: low word of p
    int m2[2][ST_WROWS];                           // mid2; high word likewise
    int cnt[2][ST_WROWS];
    unsigned long long c_key[ST_BUF]; unsigned c_row[ST_BUF];                     // candidates on their way out
    unsigned long long d_prior[ST_BUF]; unsigned d_row[ST_BUF]; int d_cnt[ST_BUF];  // deferred rows on their way out
    unsigned char list[ST_WROWS];                  // slots of the rows that need arithmetic
    unsigned long long full[2];                    // mbarriers: the stage's records have landed
};

struct StShared {
    StWarp w[ST_WARPS];
    double exp2[64];                               // 2^(j/64)
    double rcp[LOW_C_MAX];                         // 1/i
    unsigned hist[ST_HBINS];
};

struct StParams {
    const int* mid1; const int* mid2; const int* count; long long n_pairs;
    int shard_chrom; long long min_dist, max_dist; FastDiv div;
    FastDiv bdiv;                                  // the bias tables' grid step (= div unless BbkBiasTable.step says otherwise)
    const BbkFitResult* fit; const double* spline_y;
    const double* bias; const long long* chrom_base; const long long* mid0; int n_chrom; const unsigned* flags;
    long long out_base; double* p; double* q; long long* p_hist;
    unsigned long long* c_keys; unsigned* c_idx; long long c_cap;
    unsigned* d_row; int* d_cnt; double* d_prior; long long d_cap;
    BbkScoreState* st;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// global -> shared bulk copy (TMA unit, no registers), completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// the records of warp tile `wt` into stage `stage` (one lane): whole groups of four by bulk copy, a shard's last 1..3 by hand
__device__ __forceinline__ void st_issue_tile(StWarp& ws, const StParams& Q, long long wt, int stage) {
    const long long r0 = wt * ST_WROWS;
    const long long left = Q.n_pairs - r0;
    const int rows = left < ST_WROWS ? (int)left : ST_WROWS;
    const int r4 = rows & ~3;
    for (int e = r4; e < rows; ++e) {
        ws.m1[stage][e] = Q.mid1[r0 + e]; ws.m2[stage][e] = Q.mid2[r0 + e]; ws.cnt[stage][e] = Q.count[r0 + e];
    }
    if (r4) {
        const unsigned bytes = (unsigned)r4 * 4u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the stage's last readers (generic proxy) are done
        mbar_arrive_expect_tx(&ws.full[stage], 3u * bytes);
        bulk_g2s(ws.m1[stage], Q.mid1 + r0, bytes, &ws.full[stage]);
        bulk_g2s(ws.m2[stage], Q.mid2 + r0, bytes, &ws.full[stage]);
        bulk_g2s(ws.cnt[stage], Q.count + r0, bytes, &ws.full[stage]);
    } else {
        mbar_arrive(&ws.full[stage]);
    }
}

// biasDic[chr][mid] (default 1.0) for a shard whose table does not fit the 32-bit path: 64-bit arithmetic, rare
__device__ __noinline__ double st_bias_wide(const double* tab, long long nloc, long long mid0, long long R, int mid) {
    const long long off = (long long)mid - mid0;
    if (off < 0 || off % R) return 1.0;
    const long long idx = off / R;
    if (idx >= nloc) return 1.0;
    const double v = __ldg(tab + idx);
    return isnan(v) ? 1.0 : v;
}

// the shard's bias row as the 32-bit path sees it
struct StBias {
    const double* tab; const unsigned* flg;        // flg == nullptr: no flag lookups (no table for this chromosome, or the wide path)
    unsigned mid0u, spanu, fbit0;
    long long nloc, mid0;
    bool wide;                                     // the table does not fit the 32-bit path
};

// the flag bit of a locus (bias < 0 or > 4); loci outside the table have none.  An off-grid locus reads its floor entry's bit:
// at worst the row takes the exact path for nothing.
__device__ __forceinline__ bool st_flagged(const StBias& B, const FastDiv& div, int mid) {
    const unsigned o = (unsigned)mid - B.mid0u;
    if (o >= B.spanu) return false;
    const unsigned b = fastdiv31(o, div) + B.fbit0;
    return (__ldg(B.flg + (b >> 5)) >> (b & 31)) & 1u;
}

// exact floor(a / R) for a < 2^31, R >= 2 (FastDiv's 31-bit form without its R == 1 case: the host sends R == 1 elsewhere)
__device__ __forceinline__ unsigned st_div(unsigned a, const FastDiv& f) { return __umulhi(a, f.mul31) >> f.sh31; }

// e^u for -708 <= u <= 0: u = (64 k + j) ln2/64 + r, |r| <= ln2/128; 2^(j/64) from a table, e^r by its series (r^6/720 < 4e-17)
__device__ __forceinline__ double st_exp_neg(double u, const double* __restrict__ tab) {
    const double t = fma(u, 92.33248261689366, 6755399441055744.0);            // 64/ln2; 1.5 * 2^52: the integer lands in the low word
    const int k = __double2loint(t);
    const double kf = t - 6755399441055744.0;
    double r = fma(kf, -0.01083042469326756, u);                               // ln2/64, high 32 bits (kf * hi is exact)
    r = fma(kf, -2.9815858269852933e-12, r);
    double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double v = tab[k & 63] * p;                                          // in [1, 2)
    return __hiloint2double(__double2hiint(v) + (k >> 6) * 1048576, __double2loint(v));
}

// everything a list entry needs before its arithmetic: loaded one round ahead
struct StEntry { int slot, c; double y, v1, v2; bool active; };

template <bool HAS_BIAS>
__device__ __forceinline__ void st_fetch(StEntry& E, const StParams& Q, const StBias& B, const int* s_m1, const int* s_m2, const int* s_c,
                                         const unsigned char* s_list, int kk, int n_list, int k0R, int L) {
    E.active = kk < n_list;
    E.slot = E.active ? (int)s_list[kk] : 0;
    const int m1 = s_m1[E.slot], m2 = s_m2[E.slot];
    E.c = s_c[E.slot];
    const unsigned R = Q.div.R;
    const int t = (int)((unsigned)m2 - (unsigned)m1) - k0R;                                  // fithic.py:429-430 in closed form
    const unsigned qd = t > 0 ? st_div((unsigned)t + R - 1u, Q.div) : 0u;
    const unsigned i = qd > (unsigned)(L - 1) ? (unsigned)(L - 1) : qd;
    E.y = 0.0; E.v1 = 1.0; E.v2 = 1.0;
    if (E.active) {
        E.y = __ldg(Q.spline_y + i);
        if (HAS_BIAS) {
            if (B.wide) {
                E.v1 = st_bias_wide(B.tab, B.nloc, B.mid0, (long long)Q.bdiv.R, m1); E.v2 = st_bias_wide(B.tab, B.nloc, B.mid0, (long long)Q.bdiv.R, m2);
            } else if (B.spanu) {
                const unsigned o1 = (unsigned)m1 - B.mid0u, o2 = (unsigned)m2 - B.mid0u;
                const unsigned g = Q.bdiv.R;
                const unsigned x1 = st_div(o1 & 0x7fffffffu, Q.bdiv), x2 = st_div(o2 & 0x7fffffffu, Q.bdiv);
                if (o1 < B.spanu && x1 * g == o1) E.v1 = __ldg(B.tab + x1);                  // fithic.py:418-425: on the grid, inside the table
                if (o2 < B.spanu && x2 * g == o2) E.v2 = __ldg(B.tab + x2);
            }
        }
    }
}

// a warp-private buffer's content to its global list: one reservation
__device__ __forceinline__ void st_flush_cands(StWarp& ws, const StParams& Q, int lane, int& n) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd((unsigned long long*)&Q.st->n_cand, (unsigned long long)n);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int i = lane; i < n; i += 32) {
        const unsigned long long at = base + i;
        if ((long long)at < Q.c_cap) { Q.c_keys[at] = ws.c_key[i]; Q.c_idx[at] = ws.c_row[i]; } else Q.st->cand_overflow = 1;
    }
    n = 0;
    __syncwarp();
}
__device__ __forceinline__ void st_flush_deferred(StWarp& ws, const StParams& Q, int lane, int& n) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd((unsigned long long*)&Q.st->n_list, (unsigned long long)n);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int i = lane; i < n; i += 32) {
        const unsigned long long at = base + i;
        if ((long long)at < Q.d_cap) { Q.d_row[at] = ws.d_row[i]; Q.d_cnt[at] = ws.d_cnt[i]; Q.d_prior[at] = __longlong_as_double((long long)ws.d_prior[i]); }
        else Q.st->overflow = 1;
    }
    n = 0;
    __syncwarp();
}

// class of a group of four rows, 2 bits per row: 0 p = 1.0, 1 NaN (out of range), on the list: 2 count <= 1, 3 count >= 2.
// slow: 4 bits, row e's count <= 0 does not get the shortcut.
__device__ __forceinline__ unsigned st_classes(const int4& a1, const int4& a2, const int4& ac, unsigned lo_u, unsigned span_u, unsigned slow) {
    const int m1s[4] = {a1.x, a1.y, a1.z, a1.w}, m2s[4] = {a2.x, a2.y, a2.z, a2.w}, cs[4] = {ac.x, ac.y, ac.z, ac.w};
    unsigned cls = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const bool inr = ((unsigned)m2s[e] - (unsigned)m1s[e] - lo_u) <= span_u;                // fithic.py:416, :427 (coordinates >= 0 here)
        const int c = cs[e];
        unsigned cd = (unsigned)min(c, 2) + 1u;                                                 // 1 -> 2, >= 2 -> 3
        if (c <= 0) cd = ((slow >> e) & 1u) ? 2u : 0u;
        if (!inr) cd = 1u;
        cls |= cd << (2 * e);
    }
    return cls;
}

template <bool HAS_BIAS>
__global__ void __launch_bounds__(ST_THREADS, ST_CTAS_PER_SM) score_tiles_kernel(const __grid_constant__ StParams Q) {
    extern __shared__ __align__(128) unsigned char st_raw[];
    StShared& sh = *reinterpret_cast<StShared*>(st_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    StWarp& ws = sh.w[warp];
    const long long S = Q.fit->S;
    const int k0 = Q.fit->k0, L = Q.fit->L;
    if (!(Q.fit->status == BBK_FIT_OK && L > 0)) return;                       // failed fit: the host raises, p is never read
    const bool exact = Q.st->exact != 0;
    const bool all_defer = S < 64;                                             // tiny S: bdtrc's k >= n cases stay with the general code
    const double dn = (double)S;
    const unsigned R = Q.div.R;
    const int k0R = k0 * (int)R;                                               // k0 * R <= max_dist < 2^31
    const unsigned lo_u = (unsigned)Q.min_dist, span_u = (unsigned)(Q.max_dist - Q.min_dist);
    const long long n = Q.n_pairs;
    const long long n_wt = (n + ST_WROWS - 1) / ST_WROWS;
    const long long gw = (long long)blockIdx.x * ST_WARPS + warp, nw = (long long)gridDim.x * ST_WARPS;
    StBias B;
    B.tab = nullptr; B.flg = nullptr; B.mid0u = 0; B.spanu = 0; B.fbit0 = 0; B.nloc = 0; B.mid0 = 0; B.wide = false;
    if (HAS_BIAS && Q.shard_chrom >= 0 && Q.shard_chrom < Q.n_chrom) {
        const long long base = __ldg(&Q.chrom_base[Q.shard_chrom]);
        B.nloc = __ldg(&Q.chrom_base[Q.shard_chrom + 1]) - base;
        B.mid0 = __ldg(&Q.mid0[Q.shard_chrom]);
        B.tab = Q.bias + base;
        if (B.nloc > 0) {
            const unsigned long long span = (unsigned long long)B.nloc * Q.bdiv.R;
            if (B.mid0 >= 0 && B.mid0 < (1ll << 31) && span < (1ull << 31) && base + B.nloc < (1ll << 32)) {
                B.mid0u = (unsigned)B.mid0; B.spanu = (unsigned)span;
                B.flg = Q.flags + (base >> 5); B.fbit0 = (unsigned)(base & 31);
            } else B.wide = true;
        }
    }
    const bool zero_slow = exact || B.wide;                                    // count <= 0 rows go through their prior
    const bool use_flags = HAS_BIAS && B.flg != nullptr && !zero_slow;
    const bool same_grid = Q.bdiv.R == R;                                      // bias loci on the rows' own grid (the usual case)
    for (int i = tid; i < ST_HBINS; i += ST_THREADS) sh.hist[i] = 0;
    if (tid < 64) sh.exp2[tid] = exp2((double)tid * (1.0 / 64.0));
    if (tid >= 64 && tid < 64 + LOW_C_MAX) sh.rcp[tid - 64] = tid > 64 ? 1.0 / (double)(tid - 64) : 0.0;
    if (lane == 0) {
        mbar_init(&ws.full[0], 1); mbar_init(&ws.full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (lane == 0) {
        if (gw < n_wt) st_issue_tile(ws, Q, gw, 0);
        if (gw + nw < n_wt) st_issue_tile(ws, Q, gw + nw, 1);
    }
    unsigned ones = 0, nans = 0;
    int n_cb = 0, n_db = 0;                                                    // entries in the warp's two output buffers (warp-uniform)
    const unsigned lt = (1u << lane) - 1;

    for (int k = 0; ; ++k) {
        const long long wt = gw + (long long)k * nw;
        if (wt >= n_wt) break;
        const int stage = k & 1;
        const unsigned parity = (unsigned)(k >> 1) & 1u;
        const long long left = n - wt * ST_WROWS;
        const int wrows = left < ST_WROWS ? (int)left : ST_WROWS;
        int* const s_m1 = ws.m1[stage];
        int* const s_m2 = ws.m2[stage];
        const int* const s_c = ws.cnt[stage];
        const unsigned row0 = (unsigned)(Q.out_base + wt * ST_WROWS);
        if (!mbar_try_wait(&ws.full[stage], parity)) {
            while (!mbar_try_wait(&ws.full[stage], parity)) __nanosleep(64);
        }
        // ---- decode: class of every row; the result of the rows that need no arithmetic replaces them in the tile
        unsigned codes = 0;                                                    // 2 bits per row, the lane's eight rows
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int rb = u * 128 + lane * 4;
            int4 a1 = *reinterpret_cast<const int4*>(s_m1 + rb);
            int4 a2 = *reinterpret_cast<const int4*>(s_m2 + rb);
            const int4 ac = *reinterpret_cast<const int4*>(s_c + rb);
            unsigned slow = zero_slow ? 0xfu : 0u;
            if (use_flags && min(min(ac.x, ac.y), min(ac.z, ac.w)) <= 0) {
                // count <= 0: p = 1.0 unless a locus is flagged (bias < 0 or > 4); then the row goes through its prior.
                // Row-major input: the four rows share their first locus and their second loci are neighbours on the grid -
                // one division, and the four bits come out of two words.
                const unsigned t1 = (unsigned)((a1.x ^ a1.y) | (a1.x ^ a1.z) | (a1.x ^ a1.w));
                const unsigned t2 = ((unsigned)(a2.y - a2.x) - R) | ((unsigned)(a2.z - a2.y) - R) | ((unsigned)(a2.w - a2.z) - R);
                const unsigned o2 = (unsigned)a2.x - B.mid0u;
                if ((t1 | t2) == 0u && same_grid && o2 < B.spanu && o2 + 3u * R < B.spanu) {
                    const unsigned b = st_div(o2, Q.div) + B.fbit0;
                    const unsigned w0 = __ldg(B.flg + (b >> 5)), w1 = __ldg(B.flg + (b >> 5) + 1);   // (the bit map has a spare word at its end)
                    const bool f1 = st_flagged(B, Q.bdiv, a1.x);
                    slow = f1 ? 0xfu : (__funnelshift_r(w0, w1, b & 31) & 0xfu);
                } else {
                    const int m1s[4] = {a1.x, a1.y, a1.z, a1.w}, m2s[4] = {a2.x, a2.y, a2.z, a2.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) slow |= (st_flagged(B, Q.bdiv, m1s[e]) || st_flagged(B, Q.bdiv, m2s[e])) ? (1u << e) : 0u;
                }
            }
            unsigned cls = st_classes(a1, a2, ac, lo_u, span_u, slow);
            if ((a1.x | a1.y | a1.z | a1.w | a2.x | a2.y | a2.z | a2.w) < 0) {
                // a negative coordinate (never in real data): the wrapped subtraction is not the distance; rows with mid2 < mid1 are out of range
                const int m1s[4] = {a1.x, a1.y, a1.z, a1.w}, m2s[4] = {a2.x, a2.y, a2.z, a2.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) if (m2s[e] < m1s[e]) cls = (cls & ~(3u << (2 * e))) | (1u << (2 * e));
            }
            if (wrows < ST_WROWS) {
                // a shard's last tile: rows it does not have are NaN (and not counted)
#pragma unroll
                for (int e = 0; e < 4; ++e) if (rb + e >= wrows) { cls = (cls & ~(3u << (2 * e))) | (1u << (2 * e)); nans -= 1; }
            }
            int* los = reinterpret_cast<int*>(&a1);
            int* his = reinterpret_cast<int*>(&a2);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const unsigned cd = (cls >> (2 * e)) & 3u;
                los[e] = cd < 2u ? 0 : los[e];
                his[e] = cd < 2u ? (cd == 0u ? HI_ONE : HI_NAN) : his[e];
            }
            *reinterpret_cast<int4*>(s_m1 + rb) = a1;
            *reinterpret_cast<int4*>(s_m2 + rb) = a2;
            codes |= cls << (8 * u);
        }
        // ---- the warp's list: count <= 1 rows first, then the rest.  (Large counts sit next to the diagonal: in row-major input
        // they fill their own tiles, so the rounds that run the term loop are few without a third class.)
        {
            const unsigned b0 = codes & 0x5555u, b1 = (codes >> 1) & 0x5555u;
            ones += __popc(~b0 & ~b1 & 0x5555u); nans += __popc(b0 & ~b1);
        }
        const unsigned nA = __popc(~codes & (codes >> 1) & 0x5555u), nB = __popc(codes & (codes >> 1) & 0x5555u);
        const unsigned packed = nA | (nB << 16);
        unsigned inc = packed;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        const unsigned tot = __shfl_sync(0xffffffffu, inc, 31);
        const unsigned tA = tot & 0xffffu, tB = tot >> 16;
        const unsigned exc = inc - packed;
        unsigned at_a = exc & 0xffffu, at_b = tA + (exc >> 16);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const unsigned cd = (codes >> (2 * s)) & 3u;
            const unsigned pos = cd == 2u ? at_a : at_b;
            at_a += cd == 2u; at_b += cd == 3u;
            if (cd >= 2u) ws.list[pos] = (unsigned char)((s >> 2) * 128 + lane * 4 + (s & 3));
        }
        __syncwarp();
        // ---- rounds of 32 entries; the next round's gathers are in flight while one computes
        const int n_list = (int)(tA + tB);
        StEntry E;
        st_fetch<HAS_BIAS>(E, Q, B, s_m1, s_m2, s_c, ws.list, lane, n_list, k0R, L);
        for (int kb = 0; kb < n_list; kb += 32) {
            const bool active = E.active;
            const int c = E.c, slot = E.slot;
            double prior = E.y;
            if (HAS_BIAS) prior = prior * ((isnan(E.v1) ? 1.0 : E.v1) * (isnan(E.v2) ? 1.0 : E.v2));   // fithic.py:431 (NaN = locus absent)
            if (kb + 32 < n_list) st_fetch<HAS_BIAS>(E, Q, B, s_m1, s_m2, s_c, ws.list, kb + 32 + lane, n_list, k0R, L);
            const bool valid = active && prior >= 0.0 && prior <= 1.0;                          // bdtrc: NaN otherwise, before anything else
            const bool lean = valid && c >= 1 && c <= LOW_C_MAX && !all_defer && prior > 0.0 && prior < LEAN_MAX_PRIOR;
            bool defer = valid && c >= 1 && !lean;
            int hi = (valid && c <= 0) ? HI_ONE : HI_NAN, lo = 0;                               // k < 0 -> 1
            // which form a row takes depends on its own count only (never on its neighbours in the round): p is a function of
            // (count, prior, S), whatever the tiling
            const int cmax_s = __reduce_max_sync(0xffffffffu, lean && c <= SMALL_C ? c : 0);
            const int cmax_l = __reduce_max_sync(0xffffffffu, lean && c > SMALL_C ? c : 0);
            if ((cmax_s | cmax_l) > 0) {
                // P(X >= c) = 1 - pmf(0) (1 + r1 (1 + r2 (1 + ...))), c - 1 factors, r_i = (S - i + 1) q / (i (1 - q));
                // count == 1 is the same form without factors (bdtrc's 1 - (1-q)^S)
                const double q = lean ? prior : 0.0;
                const double l1m = -q * (1.0 + q * (0.5 + q * (1.0 / 3.0 + q * (0.25 + q * (0.2 + q * (1.0 / 6.0))))));
                const double u = dn * l1m;
                const double e0 = st_exp_neg(fmax(u, -708.0), sh.exp2);
                double sum = 1.0;
                if (cmax_s > 1 || cmax_l > 0) {
                    const double qr = q * (1.0 + q * (1.0 + q * (1.0 + q * (1.0 + q * (1.0 + q * (1.0 + q))))));   // q / (1 - q)
                    const double a = dn * qr;
#define BBK_ST_TERM(i) { double ri = fma(-(double)((i) - 1), qr, a) * (1.0 / (double)(i)); ri = (i) < c ? ri : 0.0; sum = fma(ri, sum, 1.0); }
                    switch (cmax_s) {                                        // warp-uniform
                        case 8: BBK_ST_TERM(7)
                        case 7: BBK_ST_TERM(6)
                        case 6: BBK_ST_TERM(5)
                        case 5: BBK_ST_TERM(4)
                        case 4: BBK_ST_TERM(3)
                        case 3: BBK_ST_TERM(2)
                        case 2: BBK_ST_TERM(1)
                        default: break;
                    }
#undef BBK_ST_TERM
                    if (cmax_l > 0) {                                        // warp-uniform: counts above SMALL_C, the terms as a loop
                        double ai = a, t = 1.0, sl = 1.0;
                        for (int i = 1; i < cmax_l; ++i) {
                            double ri = __dmul_rn(ai, sh.rcp[i]);             // (explicit roundings: however the loop is unrolled,
                            ri = i < c ? ri : 0.0;                            //  a row's p does not depend on the round's longest count)
                            t = __dmul_rn(t, ri); sl = __dadd_rn(sl, t); ai = __dsub_rn(ai, qr);
                        }
                        if (c > SMALL_C) sum = sl;
                    }
                }
                double pc = fma(-e0, sum, 1.0);
                const bool tiny = lean && c == 1 && u > -0.0078125;          // 1 - e^u loses digits: -expm1(u) by its series
                if (__any_sync(0xffffffffu, tiny)) {
                    if (tiny) pc = -u * (1.0 + u * 0.5 * (1.0 + u * (1.0 / 3.0) * (1.0 + u * 0.25 * (1.0 + u * 0.2 * (1.0 + u * (1.0 / 6.0) * (1.0 + u * (1.0 / 7.0)))))));
                }
                if (lean) {
                    if (c >= 2 && !(pc >= LOWER_MIN_P)) defer = true;        // digits lost (or the row is significant): the upper sum
                    else { hi = __double2hiint(pc); lo = __double2loint(pc); }
                }
            }
            if (defer) { hi = 0x3fe00000; lo = 0; }                          // placeholder 0.5: a number, so that q is pre-filled with 1.0
            if (active) { s_m1[slot] = lo; s_m2[slot] = hi; }
            const bool fin = active && !defer;
            ones += fin && hi == HI_ONE && lo == 0;
            nans += fin && hi == HI_NAN;
            const bool scored = fin && (unsigned)hi < (unsigned)HI_ONE;      // a p below 1.0 (p >= 0 here; false for NaN)
            if (Q.p_hist && scored) {
                const unsigned b = (unsigned)hi >> 19;                       // = (IEEE bits >> 51), below 2046
                if (b >= (unsigned)ST_HBASE) atomicAdd(&sh.hist[b - ST_HBASE], 1u);
                else atomicAdd((unsigned long long*)&Q.p_hist[b], 1ull);
            }
            const unsigned row = row0 + (unsigned)slot;
            if (Q.c_keys) {
                const bool cand = scored && (unsigned)hi < 0x3fa00000u;      // p < BBK_SMALL_P = 2^-5
                const unsigned m = __ballot_sync(0xffffffffu, cand);
                if (m) {
                    if (cand) { const int at = n_cb + __popc(m & lt); ws.c_key[at] = ((unsigned long long)(unsigned)hi << 32 | (unsigned)lo) | 0x8000000000000000ull; ws.c_row[at] = row; }
                    n_cb += __popc(m);
                    __syncwarp();
                    if (n_cb >= 32) st_flush_cands(ws, Q, lane, n_cb);
                }
            }
            {
                const unsigned m = __ballot_sync(0xffffffffu, defer);
                if (m) {
                    if (defer) { const int at = n_db + __popc(m & lt); ws.d_prior[at] = (unsigned long long)__double_as_longlong(prior); ws.d_row[at] = row; ws.d_cnt[at] = c; }
                    n_db += __popc(m);
                    __syncwarp();
                    if (n_db >= 32) st_flush_deferred(ws, Q, lane, n_db);
                }
            }
        }
        __syncwarp();
        // ---- the warp's rows out: coalesced 128-bit stores of p and of q = 1.0 / NaN
        const int wrows4 = (wrows + 3) & ~3;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int rb = u * 128 + lane * 4;
            if (rb < wrows4) {
                const int4 lo = *reinterpret_cast<const int4*>(s_m1 + rb);
                const int4 hi = *reinterpret_cast<const int4*>(s_m2 + rb);
                double2* pv2 = reinterpret_cast<double2*>(Q.p + row0 + rb);
                st_stream_double2(pv2, make_double2(__hiloint2double(hi.x, lo.x), __hiloint2double(hi.y, lo.y)));
                st_stream_double2(pv2 + 1, make_double2(__hiloint2double(hi.z, lo.z), __hiloint2double(hi.w, lo.w)));
                if (Q.q) {
                    double2* qv2 = reinterpret_cast<double2*>(Q.q + row0 + rb);
                    st_stream_double2(qv2, make_double2(__hiloint2double(hi.x == HI_NAN ? HI_NAN : HI_ONE, 0), __hiloint2double(hi.y == HI_NAN ? HI_NAN : HI_ONE, 0)));
                    st_stream_double2(qv2 + 1, make_double2(__hiloint2double(hi.z == HI_NAN ? HI_NAN : HI_ONE, 0), __hiloint2double(hi.w == HI_NAN ? HI_NAN : HI_ONE, 0)));
                }
            }
        }
        // ---- the stage is free: the tile after next
        __syncwarp();
        if (lane == 0 && wt + 2 * nw < n_wt) st_issue_tile(ws, Q, wt + 2 * nw, stage);
    }
    if (n_cb) st_flush_cands(ws, Q, lane, n_cb);
    if (n_db) st_flush_deferred(ws, Q, lane, n_db);
    if (Q.p_hist) {
        __syncthreads();
        for (int i = tid; i < ST_HBINS; i += ST_THREADS) {
            const unsigned v = sh.hist[i];
            if (v) atomicAdd((unsigned long long*)&Q.p_hist[ST_HBASE + i], (unsigned long long)v);
        }
        const unsigned o = __reduce_add_sync(0xffffffffu, ones), zn = __reduce_add_sync(0xffffffffu, nans);
        if (lane == 0) {
            if (o) atomicAdd((unsigned long long*)&Q.p_hist[BBK_PHIST_BINS], (unsigned long long)o);
            if (zn) atomicAdd((unsigned long long*)&Q.p_hist[BBK_PHIST_BINS + 1], (unsigned long long)zn);
        }
    }
}

// one flag bit per bias table entry: the value is < 0 or > 4 (NaN = absent locus = 1.0 carries none)
__global__ void bias_flags_kernel(const double* bias, long long n, unsigned* flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool f = false;
    if (i < n) { const double v = bias[i]; f = v < 0.0 || v > BIAS_FLAG_MAX; }
    const unsigned w = __ballot_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0 && i < n) flags[i >> 5] = w;
}

// after the fit: may the count <= 0 shortcut stand?
__global__ void __launch_bounds__(1024) score_guard_kernel(const BbkFitResult* fit, const double* spline_y, BbkScoreState* st) {
    __shared__ double s_lo[32], s_hi[32];
    __shared__ int s_nan[32];
    const int L = fit->status == BBK_FIT_OK ? fit->L : 0;
    double lo = INFINITY, hi = -INFINITY;
    int bad = 0;
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const double v = spline_y[i];
        if (isnan(v)) bad = 1;
        lo = fmin(lo, v); hi = fmax(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; s_nan[threadIdx.x >> 5] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w) { lo = fmin(lo, s_lo[w]); hi = fmax(hi, s_hi[w]); bad |= s_nan[w]; }
        // count <= 0 rows whose loci carry no flag have 0 <= b1*b2 <= 16 and are taken as p = 1.0: right iff every
        // prior splineY * b1*b2 lies in [0, 1]
        const bool ok = L > 0 && !bad && lo >= 0.0 && hi * (BIAS_FLAG_MAX * BIAS_FLAG_MAX) <= 1.0;
        if (L > 0 && !ok) st->exact = 1;
    }
}

struct DfParams {
    const unsigned* d_row; const int* d_cnt; const double* d_prior; long long d_cap;
    const BbkFitResult* fit;
    double* p; double* q; long long* p_hist;
    unsigned long long* c_keys; unsigned* c_idx; long long c_cap;
    BbkScoreState* st;
};

struct DfShared {
    double rcp[RCP_TAB];
    double lfact[LF_TAB];
    unsigned hist[BBK_PHIST_BINS];
};

__global__ void __launch_bounds__(PV_THREADS, 3) score_deferred_kernel(DfParams D) {
    extern __shared__ __align__(16) unsigned char df_raw[];
    DfShared& sh = *reinterpret_cast<DfShared*>(df_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!(D.fit->status == BBK_FIT_OK && D.fit->L > 0)) return;
    if (D.st->overflow) return;                                                 // the host repeats the pass with a larger list
    const long long n = (long long)D.st->n_list;
    if (n == 0) return;
    const long long S = D.fit->S;
    for (int j = tid; j < RCP_TAB; j += PV_THREADS) sh.rcp[j] = g_rcp[j];
    for (int j = tid; j < LF_TAB; j += PV_THREADS) sh.lfact[j] = g_lfact[j];
    for (int i = tid; i < BBK_PHIST_BINS; i += PV_THREADS) sh.hist[i] = 0;
    __syncthreads();
    TailConst K;
    K.dn = (double)S;
    K.inv_n = S > 0 ? 1.0 / K.dn : 0.0;
    K.c_max = 1e-4 * K.dn;
    K.c5_max = 2e-11 * (K.dn * K.dn) * (K.dn * K.dn);
    const bool s_fits = S <= 0x7fffffffll;
    const int s_cap = s_fits ? (int)S : 0x7fffffff;
    unsigned ones = 0, nans = 0;
    const long long rounds = (n + 31) >> 5;
    for (long long r = (long long)blockIdx.x * PV_WARPS + warp; r < rounds; r += (long long)gridDim.x * PV_WARPS) {
        const long long k = (r << 5) + lane;
        const bool active = k < n;
        unsigned row = 0; int c = 0; double prior = 0.0;
        if (active) { row = D.d_row[k]; c = D.d_cnt[k]; prior = D.d_prior[k]; }
        int cls = 0;
        double out = __longlong_as_double(0x7ff8000000000000ll);
        if (active) cls = bdtrc_class(c, s_cap, s_fits, prior, &out);
        const bool closed = cls == 1, upper = cls == 2;
        if (__any_sync(0xffffffffu, closed)) { if (closed) out = -expm1(K.dn * log1m(prior)); }    // bdtrc's closed form for k == 0
        if (__any_sync(0xffffffffu, upper)) {
            TailState T;
            T.lp = 0.0; T.term = 0.0; T.sum = 1.0; T.a = 0.0; T.step = 0.0; T.j = 0;
            bool running = false, fast = false;
            if (upper) {
                fast = fast_ok(c, K);
                if (fast) { tail_setup(c, prior, K, sh.lfact, T); running = true; }
                else out = tail_general(c, S, prior);
            }
            unsigned rmask = __ballot_sync(0xffffffffu, running);
            while (rmask) {
                if (running) {
                    const bool exhausted = tail_terms16(T, K, sh.rcp);
                    running = !(exhausted || T.term < TAIL_EPS * T.sum);
                }
                rmask = __ballot_sync(0xffffffffu, running);
            }
            if (fast) out = tail_finish(T.lp, T.sum);
        }
        const double pv = finish_p(out);                                        // fithic.py:434
        if (active) {
            D.p[row] = pv;
            if (D.q && isnan(pv)) D.q[row] = pv;                                // (the stream pre-filled 1.0)
            if (D.p_hist) hist_p(sh.hist, pv, ones, nans);
        }
        if (D.c_keys) {
            const bool cand = active && pv < BBK_SMALL_P;
            const unsigned m = __ballot_sync(0xffffffffu, cand);
            if (m) {
                const int leader = __ffs(m) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd((unsigned long long*)&D.st->n_cand, (unsigned long long)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (cand) {
                    const unsigned long long at = base + __popc(m & ((1u << lane) - 1));
                    if ((long long)at < D.c_cap) { D.c_keys[at] = bbk_key_of(pv); D.c_idx[at] = row; }
                    else D.st->cand_overflow = 1;
                }
            }
        }
    }
    if (D.p_hist) {
        __syncthreads();
        for (int i = tid; i < BBK_PHIST_BINS; i += PV_THREADS) {
            const unsigned v = sh.hist[i];
            if (v) atomicAdd((unsigned long long*)&D.p_hist[i], (unsigned long long)v);
        }
        const unsigned o = __reduce_add_sync(0xffffffffu, ones), zn = __reduce_add_sync(0xffffffffu, nans);
        if (lane == 0) {
            if (o) atomicAdd((unsigned long long*)&D.p_hist[BBK_PHIST_BINS], (unsigned long long)o);
            if (zn) atomicAdd((unsigned long long*)&D.p_hist[BBK_PHIST_BINS + 1], (unsigned long long)zn);
        }
    }
}

__global__ void score_begin_kernel(BbkScoreState* st, long long* p_hist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (p_hist && i < BBK_PHIST_LEN) p_hist[i] = 0;
    if (i == 0) { st->n_list = 0; st->n_cand = 0; st->overflow = 0; st->cand_overflow = 0; st->exact = 0; st->reserved = 0; }
}
