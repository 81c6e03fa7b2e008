// pvalue.cu - K4: per-pair binomial survival p-values (reference: fit_spline scoring loop,
// fithic.py:413-435):   p = scipy.special.bdtrc(count-1, S, newSplineY[i] * (bias1*bias2)).
//
// Fused elementwise kernel, 12 B/pair in (int32 mid1, mid2, count), 8 B/pair out (float64 p):
//   load (128-bit, streaming) -> distance -> spline index (closed form of the reference's bisect)
//   -> bias gathers (L2-resident tables) -> prior -> log-space binomial tail -> store (128-bit).
//
// The tail P(X >= c), X ~ Binomial(S, q), is the regularised incomplete beta I_q(c, S-c+1) that
// cephes' bdtrc evaluates.  Here it is computed as  pmf(c) * (1 + r_c + r_c r_{c+1} + ...)  from the
// mode outwards (upper tail when c >= (S+1)q, else 1 - lower tail), with pmf(c) in log space by the
// saddle-point form (Stirling error + deviance terms; C. Loader, "Fast and accurate computation of
// binomial probabilities", 2000) so nothing cancels even at S ~ 2^31.  Against 60-digit arithmetic
// this is good to ~3e-13 in log10 p; cephes itself is only good to ~3e-6 there, which is
// what bounds the parity tolerance.
// bdtrc's edge semantics are kept exactly: NaN for a prior outside [0,1] or NaN (checked BEFORE the
// count, so a count of 0 with a negative prior is dropped too), 1.0 for count <= 0, NaN for
// count-1 > S, 0.0 for count-1 == S, the count == 1 closed form, 0 / 1 at q == 0 / 1.
#include "common.cuh"

namespace {

constexpr int PV_THREADS = 256;

__device__ __constant__ double c_stirl[16] = {
    0.0, 0.08106146679532726, 0.04134069595540929, 0.02767792568499834, 0.02079067210376509,
    0.01664469118982119, 0.01387612882307075, 0.01189670994589177, 0.01041126526197209,
    0.009255462182712733, 0.008330563433362871, 0.007573675487951841, 0.006942840107209530,
    0.006408994188004207, 0.005951370112758848, 0.005554733551962801};

// log(n!) - [n log n - n + 0.5 log(2 pi n)]
__device__ __forceinline__ double stirlerr(double n) {
    if (n <= 15.0) return c_stirl[(int)n];
    const double S0 = 1.0 / 12.0, S1 = 1.0 / 360.0, S2 = 1.0 / 1260.0, S3 = 1.0 / 1680.0, S4 = 1.0 / 1188.0;
    double rn = 1.0 / n, rnn = rn * rn;
    if (n > 500.0) return (S0 - S1 * rnn) * rn;
    if (n > 80.0) return (S0 - (S1 - S2 * rnn) * rnn) * rn;
    if (n > 35.0) return (S0 - (S1 - (S2 - S3 * rnn) * rnn) * rnn) * rn;
    return (S0 - (S1 - (S2 - (S3 - S4 * rnn) * rnn) * rnn) * rnn) * rn;
}

// x log(x/M) + M - x, without cancellation when x ~ M
__device__ __forceinline__ double bd0(double x, double M) {
    double diff = x - M;
    if (fabs(diff) < 0.1 * (x + M)) {
        double v = diff / (x + M);
        double s = diff * v;
        double ej = 2.0 * x * v;
        double v2 = v * v;
        for (int j = 1; j < 1000; ++j) {
            ej *= v2;
            double s1 = s + ej / (double)(2 * j + 1);
            if (s1 == s) return s1;
            s = s1;
        }
        return s;
    }
    return x * log(x / M) + M - x;
}

struct TailConst {      // per-launch constants of the binomial tail (functions of S only)
    double dn;          // (double) S
    double inv_n;       // 1 / S
    double stirl_n;     // stirlerr(S)
};

// P(X >= c) for 2 <= c <= S - 1 ... and c == S handled by the caller; 0 < q < 1
__device__ double binom_upper_tail(int c, long long S, double q, const TailConst& K) {
    const double dn = K.dn;
    const double dc = (double)c;
    const double mu = dn * q;
    const double nmc = dn - dc;
    double lc = K.stirl_n - stirlerr(dc) - stirlerr(nmc) - bd0(dc, mu) - bd0(nmc, dn * (1.0 - q));
    double lf = 1.8378770664093454836 /* log(2 pi) */ + log(dc) + log1p(-dc * K.inv_n);
    double lp = lc - 0.5 * lf;                       // log pmf(c)
    const double qr = q / (1.0 - q);
    const double eps = 5.7e-14;                      // 2^-44: far inside the 1e-9 budget
    if (dc >= (dn + 1.0) * q) {
        // upper tail: pmf(j+1)/pmf(j) = (n-j)/(j+1) * q/(1-q), decreasing from j = c
        double term = 1.0, sum = 1.0, dj = dc, rem = nmc;
        while (rem > 0.0) {
            term *= rem * qr / (dj + 1.0);
            sum += term;
            dj += 1.0;
            rem -= 1.0;
            if (term < eps * sum) break;
        }
        if (lp < -690.0) return exp(lp + log(sum));  // keep the denormal range reachable
        return exp(lp) * sum;
    }
    // lower tail: pmf(j-1)/pmf(j) = j / ((n-j+1) * q/(1-q)), decreasing from j = c-1 downwards
    const double iqr = (1.0 - q) / q;
    double term = dc * iqr / (nmc + 1.0), sum = term, dj = dc - 1.0;
    while (dj > 0.0) {
        term *= dj * iqr / (dn - dj + 1.0);
        sum += term;
        dj -= 1.0;
        if (term < eps * sum) break;
    }
    return 1.0 - exp(lp) * sum;
}

// scipy.special.bdtrc(c - 1, S, q) semantics
__device__ __forceinline__ double bdtrc_like(int c, long long S, double q, const TailConst& K) {
    if (isnan(q)) return q;
    if (q < 0.0 || q > 1.0) return __longlong_as_double(0x7ff8000000000000ll);
    long long k = (long long)c - 1;
    if (k < 0) return 1.0;
    if (S < k) return __longlong_as_double(0x7ff8000000000000ll);
    if (k == S) return 0.0;
    if (k == 0) {
        if (q < 0.01) return -expm1(K.dn * log1p(-q));
        return 1.0 - pow(1.0 - q, K.dn);
    }
    if (q == 0.0) return 0.0;
    if (q == 1.0) return 1.0;
    if ((long long)c == S) return exp(K.dn * log(q));          // I_q(S, 1) = q^S
    return binom_upper_tail(c, S, q, K);
}

struct PvParams {
    const int32_t* chr1;
    const int32_t* chr2;
    const int32_t* mid1;
    const int32_t* mid2;
    const int32_t* count;
    long long n_pairs;
    int shard_chrom;
    long long R, min_dist, max_dist;
    FastDiv div;
    const BbkFitResult* fit;
    const double* spline_y;
    const double* bias;
    const long long* chrom_base;
    const long long* mid0;
    int n_chrom;
    double* p;
    long long* p_hist;
};

__device__ __forceinline__ double bias_lookup(const PvParams& P, int chrom, int mid) {
    // biasDic[chr][mid] with default 1.0 (fithic.py:418-425) on a dense per-chromosome grid
    if (chrom < 0 || chrom >= P.n_chrom) return 1.0;
    long long base = __ldg(&P.chrom_base[chrom]);
    long long nloc = __ldg(&P.chrom_base[chrom + 1]) - base;
    long long off = (long long)mid - __ldg(&P.mid0[chrom]);
    if (off < 0 || off >= (1ll << 32)) return 1.0;
    unsigned idx = fastdiv((unsigned)off, P.div);
    if ((long long)idx * P.div.R != off || (long long)idx >= nloc) return 1.0;
    double v = __ldg(&P.bias[base + idx]);
    return isnan(v) ? 1.0 : v;
}

template <bool HAS_CHR, bool HAS_BIAS>
__device__ __forceinline__ double score_record(const PvParams& P, int m1, int m2, int c, int c1, int c2,
                                               long long S, int k0, int L, const TailConst& K) {
    long long d = (long long)m2 - (long long)m1;                         // fithic.py:416
    if (!(P.min_dist <= d && d <= P.max_dist)) return __longlong_as_double(0x7ff8000000000000ll);   // :427 not scored
    // i = min(bisect_left(splineX, clamp(d, min_x, max_x)), L-1) == clamp(ceil((d - splineX[0]) / R), 0, L-1)
    long long t = d - (long long)k0 * P.R;
    int i = 0;
    if (t > 0) {
        long long tt = t + P.R - 1;
        long long qd = tt < (1ll << 32) ? (long long)fastdiv((unsigned)tt, P.div) : tt / P.R;
        i = qd > (long long)(L - 1) ? L - 1 : (int)qd;
    }
    double prior = __ldg(&P.spline_y[i]);
    if (HAS_BIAS) {
        double b1 = bias_lookup(P, HAS_CHR ? c1 : P.shard_chrom, m1);
        double b2 = bias_lookup(P, HAS_CHR ? c2 : P.shard_chrom, m2);
        prior = prior * (b1 * b2);                                        // :431
    }
    double pv = bdtrc_like(c, S, prior, K);                              // :432
    if (!(pv <= 1.0)) pv = __longlong_as_double(0x7ff8000000000000ll);   // :434 row dropped
    return pv;
}

__device__ __forceinline__ void hist_p(unsigned* sh_hist, double pv, unsigned& ones, unsigned& nans) {
    if (isnan(pv)) { nans += 1; return; }
    if (pv == 1.0) { ones += 1; return; }
    unsigned b = (unsigned)((unsigned long long)__double_as_longlong(pv) >> 51) & (BBK_PHIST_BINS - 1);
    atomicAdd(&sh_hist[b], 1u);
}

template <bool HAS_CHR, bool HAS_BIAS, bool HIST>
__global__ void __launch_bounds__(PV_THREADS, 3) pvalues_kernel(PvParams P) {
    __shared__ unsigned sh_hist[HIST ? BBK_PHIST_BINS : 1];
    if (HIST) {
        for (int i = threadIdx.x; i < BBK_PHIST_BINS; i += blockDim.x) sh_hist[i] = 0;
        __syncthreads();
    }
    const long long S = P.fit->S;
    const int k0 = P.fit->k0, L = P.fit->L;
    const bool fit_ok = P.fit->status == BBK_FIT_OK && L > 0;
    TailConst K;
    K.dn = (double)S;
    K.inv_n = S > 0 ? 1.0 / K.dn : 0.0;
    K.stirl_n = S > 0 ? stirlerr(K.dn) : 0.0;
    unsigned ones = 0, nans = 0;

    const long long n_groups = P.n_pairs >> 2;
    const int4* m1v = reinterpret_cast<const int4*>(P.mid1);
    const int4* m2v = reinterpret_cast<const int4*>(P.mid2);
    const int4* cv = reinterpret_cast<const int4*>(P.count);
    const int4* c1v = reinterpret_cast<const int4*>(P.chr1);
    const int4* c2v = reinterpret_cast<const int4*>(P.chr2);
    double2* pv2 = reinterpret_cast<double2*>(P.p);
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += (long long)gridDim.x * blockDim.x) {
        int4 a1 = ld_stream_int4(m1v + g), a2 = ld_stream_int4(m2v + g), ac = ld_stream_int4(cv + g);
        int4 x1 = make_int4(0, 0, 0, 0), x2 = x1;
        if (HAS_CHR) { x1 = ld_stream_int4(c1v + g); x2 = ld_stream_int4(c2v + g); }
        double r0 = qnan, r1 = qnan, r2 = qnan, r3 = qnan;
        if (fit_ok) {
            r0 = score_record<HAS_CHR, HAS_BIAS>(P, a1.x, a2.x, ac.x, x1.x, x2.x, S, k0, L, K);
            r1 = score_record<HAS_CHR, HAS_BIAS>(P, a1.y, a2.y, ac.y, x1.y, x2.y, S, k0, L, K);
            r2 = score_record<HAS_CHR, HAS_BIAS>(P, a1.z, a2.z, ac.z, x1.z, x2.z, S, k0, L, K);
            r3 = score_record<HAS_CHR, HAS_BIAS>(P, a1.w, a2.w, ac.w, x1.w, x2.w, S, k0, L, K);
        }
        st_stream_double2(pv2 + 2 * g, make_double2(r0, r1));
        st_stream_double2(pv2 + 2 * g + 1, make_double2(r2, r3));
        if (HIST) { hist_p(sh_hist, r0, ones, nans); hist_p(sh_hist, r1, ones, nans); hist_p(sh_hist, r2, ones, nans); hist_p(sh_hist, r3, ones, nans); }
    }
    if (blockIdx.x == 0 && threadIdx.x < (P.n_pairs & 3)) {
        long long i = (n_groups << 2) + threadIdx.x;
        int c1 = 0, c2 = 0;
        if (HAS_CHR) { c1 = P.chr1[i]; c2 = P.chr2[i]; }
        double r = fit_ok ? score_record<HAS_CHR, HAS_BIAS>(P, P.mid1[i], P.mid2[i], P.count[i], c1, c2, S, k0, L, K) : qnan;
        P.p[i] = r;
        if (HIST) hist_p(sh_hist, r, ones, nans);
    }
    if (HIST) {
        __syncthreads();
        for (int i = threadIdx.x; i < BBK_PHIST_BINS; i += blockDim.x) {
            unsigned v = sh_hist[i];
            if (v) atomicAdd((unsigned long long*)&P.p_hist[i], (unsigned long long)v);
        }
        unsigned o = __reduce_add_sync(0xffffffffu, ones), z = __reduce_add_sync(0xffffffffu, nans);
        if ((threadIdx.x & 31) == 0) {
            if (o) atomicAdd((unsigned long long*)&P.p_hist[BBK_PHIST_BINS], (unsigned long long)o);
            if (z) atomicAdd((unsigned long long*)&P.p_hist[BBK_PHIST_BINS + 1], (unsigned long long)z);
        }
    }
}

template <bool HAS_CHR, bool HAS_BIAS, bool HIST>
int launch_pv(const PvParams& P, cudaStream_t st) {
    long long groups = P.n_pairs >> 2;
    long long need = (groups + PV_THREADS - 1) / PV_THREADS;
    long long grid = (long long)bbk_num_sms() * 3 * 4;     // 4 waves of 3 CTAs/SM: evens out the divergent tails
    if (need < grid) grid = need > 0 ? need : 1;
    pvalues_kernel<HAS_CHR, HAS_BIAS, HIST><<<(unsigned)grid, PV_THREADS, 0, st>>>(P);
    BBK_CHECK_LAUNCH("pvalues_kernel");
    return BBK_OK;
}

}  // namespace

extern "C" int bbk_pvalues(const int32_t* d_chr1, const int32_t* d_chr2, const int32_t* d_mid1, const int32_t* d_mid2,
                           const int32_t* d_count, int64_t n_pairs, int32_t shard_chrom, int64_t resolution, int64_t min_dist,
                           int64_t max_dist, const BbkFitResult* d_fit, const double* d_spline_y, const BbkBiasTable* bias,
                           double* d_p, int64_t* d_p_hist, void* stream) {
    BBK_REQUIRE(n_pairs >= 0, "bbk_pvalues: negative size");
    BBK_REQUIRE(resolution > 0 && resolution < (1ll << 32), "bbk_pvalues: resolution must be in [1, 2^32)");
    BBK_REQUIRE((d_chr1 == nullptr) == (d_chr2 == nullptr), "bbk_pvalues: chr1/chr2 must both be given or both NULL");
    BBK_REQUIRE(d_fit && d_spline_y, "bbk_pvalues: null fit");
    if (n_pairs == 0) return BBK_OK;
    BBK_REQUIRE(d_mid1 && d_mid2 && d_count && d_p, "bbk_pvalues: null column");
    uintptr_t align = (uintptr_t)d_mid1 | (uintptr_t)d_mid2 | (uintptr_t)d_count | (uintptr_t)d_chr1 | (uintptr_t)d_chr2 | (uintptr_t)d_p;
    BBK_REQUIRE((align & 15) == 0, "bbk_pvalues: columns must be 16-byte aligned");
    PvParams P = {};
    P.chr1 = d_chr1; P.chr2 = d_chr2; P.mid1 = d_mid1; P.mid2 = d_mid2; P.count = d_count;
    P.n_pairs = n_pairs; P.shard_chrom = shard_chrom; P.R = resolution; P.min_dist = min_dist; P.max_dist = max_dist;
    P.div = make_fastdiv((uint64_t)resolution);
    P.fit = d_fit; P.spline_y = d_spline_y; P.p = d_p; P.p_hist = (long long*)d_p_hist;
    bool has_bias = bias && bias->d_bias;
    if (has_bias) {
        BBK_REQUIRE(bias->d_chrom_base && bias->d_mid0 && bias->n_chrom > 0, "bbk_pvalues: incomplete bias table");
        P.bias = bias->d_bias; P.chrom_base = (const long long*)bias->d_chrom_base; P.mid0 = (const long long*)bias->d_mid0;
        P.n_chrom = bias->n_chrom;
    }
    cudaStream_t st = (cudaStream_t)stream;
    bool chr = d_chr1 != nullptr, hist = d_p_hist != nullptr;
#define BBK_PV_CASE(C, B, H) if (chr == C && has_bias == B && hist == H) return launch_pv<C, B, H>(P, st);
    BBK_PV_CASE(false, false, false) BBK_PV_CASE(false, false, true)
    BBK_PV_CASE(false, true, false)  BBK_PV_CASE(false, true, true)
    BBK_PV_CASE(true, false, false)  BBK_PV_CASE(true, false, true)
    BBK_PV_CASE(true, true, false)   BBK_PV_CASE(true, true, true)
#undef BBK_PV_CASE
    return BBK_E_INVALID;
}
