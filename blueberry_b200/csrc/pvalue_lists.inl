// pvalue_lists.inl - K4 split around the fit (included by pvalue.cu inside its anonymous namespace).
//
// The scoring loop of fithic.py:413-435 needs the spline (hence the whole distance table) only for the prior of a record;
// everything else - the distance, the range test (:427), the two bias gathers (:418-425) and their product, the count - is
// known as soon as the records are on the device.  And most records need no arithmetic at all: a zero count gives p = 1
// (or NaN when the prior is outside [0, 1], which bdtrc checks first).  So the work is cut in two:
//
//   K4a classify_kernel (streams every record once, runs on a side stream WHILE the one-CTA fit kernel runs):
//       out of range                         -> p = q = NaN
//       count <= 0, 0 <= b1*b2 <= 16         -> p = q = 1.0      (speculative: needs 0 <= splineY, 16 max(splineY) <= 1)
//       count <= 0, b1*b2 < 0 or NaN         -> p = q = NaN      (speculative: needs splineY > 0)
//       count == 1                           -> appended to the front of the work list
//       everything else                      -> appended to the back of the work list
//     A work-list entry is (row, count, distance, b1*b2): 20 bytes, all the fit-independent state of the record.
//   score_guard_kernel (after the fit): checks the two conditions on splineY.  If one fails it raises
//       BbkScoreState::exact and K4a is run again, this time sending EVERY in-range record to the list (the launch is
//       always enqueued and returns at once when the flag is down), so the result is exact in every case.
//   K4b listed_kernel: dense lists, every lane busy on the same branch: prior = splineY[i] * (b1*b2), bdtrc's case
//       analysis, the closed form for count == 1, the tail sum otherwise; scatters p (and q = 1.0 / NaN), fills the coarse
//       p histogram and appends the rows with p < BBK_SMALL_P to the q-value step's candidate list.

constexpr int CL_THREADS = 256;
constexpr double CL_BB_MAX = 16.0;             // biases are in [0.5, 2] (fithic.py:147-149): products above 16 are deferred

struct ClsParams {
    PvParams pv;                                // records, bias table, range, divisor (fit / spline / p_hist unused)
    long long out_base;                         // row of record 0 in the rank-local p / q buffers (multiple of 4)
    unsigned* l_idx; int* l_cnt; int* l_dist; double* l_bb; long long cap;
    BbkScoreState* st;
    int exact_only;                             // 1: the post-fit relaunch (runs only when st->exact is up)
};

// class of one record before the fit.  0: p = 1.0, 1: p = NaN, 2: list front (count == 1), 3: list back
__device__ __forceinline__ int pre_class(bool in_range, int c, double bb, bool has_bias, bool exact) {
    if (!in_range) return 1;
    if (c == 1) return 2;
    if (c >= 2 || exact) return 3;
    if (!has_bias) return 0;
    if (bb >= 0.0 && bb <= CL_BB_MAX) return 0;
    if (bb < 0.0 || isnan(bb)) return 1;
    return 3;
}

template <bool HAS_CHR, bool HAS_BIAS>
__global__ void __launch_bounds__(CL_THREADS, 3) classify_kernel(ClsParams C) {
    if (C.exact_only && C.st->exact == 0) return;
    const bool exact = C.exact_only != 0;
    __shared__ unsigned s_w[CL_THREADS / 32];
    __shared__ unsigned long long s_base[2];
    const PvParams& P = C.pv;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    BiasRow shard_row = {0, 0, 0, 0};
    if (HAS_BIAS && !HAS_CHR) shard_row = bias_row(P, P.shard_chrom);
    const long long n_groups = P.n_pairs >> 2;
    const long long tile_groups = 2ll * CL_THREADS;
    const long long n_tiles = (n_groups + tile_groups - 1) / tile_groups;
    const int4* m1v = reinterpret_cast<const int4*>(P.mid1);
    const int4* m2v = reinterpret_cast<const int4*>(P.mid2);
    const int4* cv = reinterpret_cast<const int4*>(P.count);
    const int4* c1v = reinterpret_cast<const int4*>(P.chr1);
    const int4* c2v = reinterpret_cast<const int4*>(P.chr2);
    double2* pv2 = P.p ? reinterpret_cast<double2*>(P.p + C.out_base) : nullptr;
    double2* qv2 = P.q ? reinterpret_cast<double2*>(P.q + C.out_base) : nullptr;
    const unsigned lo_u = (unsigned)P.min_dist, span_u = (unsigned)(P.max_dist - P.min_dist);
    unsigned ones = 0, nans = 0;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int cnt[8], dist[8], code[8];
        double bb[8];
        unsigned n1 = 0, n2 = 0;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long g = tile * tile_groups + u * CL_THREADS + tid;
            const bool live = g < n_groups;
            const int4 z4 = make_int4(0, 0, 0, 0);
            int4 a1 = z4, a2 = z4, ac = z4, x1 = z4, x2 = z4;
            if (live) {
                a1 = ld_stream_int4(m1v + g); a2 = ld_stream_int4(m2v + g); ac = ld_stream_int4(cv + g);
                if (HAS_CHR) { x1 = ld_stream_int4(c1v + g); x2 = ld_stream_int4(c2v + g); }
            }
            const int m1s[4] = {a1.x, a1.y, a1.z, a1.w}, m2s[4] = {a2.x, a2.y, a2.z, a2.w}, cs[4] = {ac.x, ac.y, ac.z, ac.w};
            const int c1s[4] = {x1.x, x1.y, x1.z, x1.w}, c2s[4] = {x2.x, x2.y, x2.z, x2.w};
            // where the bias entries are, then all the loads, then the products: the gathers of a group are in flight together
            const bool same1 = !HAS_CHR && (a1.x == a1.y) & (a1.y == a1.z) & (a1.z == a1.w);
            bool ok1[4], ok2[4];
            long long at1[4], at2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                ok1[e] = false; ok2[e] = false; at1[e] = 0; at2[e] = 0;
                if (HAS_BIAS) {
                    if (HAS_CHR) {
                        ok1[e] = bias_index<true>(P, bias_row(P, c1s[e]), m1s[e], &at1[e]);
                        ok2[e] = bias_index<true>(P, bias_row(P, c2s[e]), m2s[e], &at2[e]);
                    } else {
                        if (e == 0 || !same1) ok1[e] = bias_index<true>(P, shard_row, m1s[e], &at1[e]);
                        ok2[e] = bias_index<true>(P, shard_row, m2s[e], &at2[e]);
                    }
                }
            }
            double v1[4], v2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                v1[e] = 1.0; v2[e] = 1.0;
                if (HAS_BIAS && live) {
                    if (e == 0 || !same1) v1[e] = __ldg(&P.bias[at1[e]]);
                    v2[e] = __ldg(&P.bias[at2[e]]);
                }
            }
            double pr[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int s = u * 4 + e;
                const unsigned ud = (unsigned)m2s[e] - (unsigned)m1s[e];                    // fithic.py:416
                const bool inr = live && m2s[e] >= m1s[e] && (ud - lo_u) <= span_u;         // fithic.py:427 (inclusive on both sides)
                double b = 1.0;
                if (HAS_BIAS) {
                    const double b1 = same1 ? bias_value(v1[0], ok1[0]) : bias_value(v1[e], ok1[e]);
                    b = b1 * bias_value(v2[e], ok2[e]);                                      // (bias1 * bias2) of :431
                }
                const int k = pre_class(inr, cs[e], b, HAS_BIAS, exact);
                code[s] = live ? k : 4;
                cnt[s] = cs[e]; dist[s] = (int)ud; bb[s] = b;
                n1 += k == 2 && live; n2 += k == 3 && live;
                ones += k == 0 && live; nans += k == 1 && live;
                pr[e] = k == 0 ? 1.0 : qnan;                                                 // list rows: NaN until K4b writes them
            }
            if (live) {
                if (pv2) { st_stream_double2(pv2 + 2 * g, make_double2(pr[0], pr[1])); st_stream_double2(pv2 + 2 * g + 1, make_double2(pr[2], pr[3])); }
                if (qv2) { st_stream_double2(qv2 + 2 * g, make_double2(pr[0], pr[1])); st_stream_double2(qv2 + 2 * g + 1, make_double2(pr[2], pr[3])); }
            }
        }
        // positions in the two lists: scan inside the warp, then over the warps, ONE global atomic per list and CTA tile
        unsigned packed = n1 | (n2 << 16), inc = packed;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        if (tid == 0) {
            unsigned run = 0;
#pragma unroll
            for (int w = 0; w < CL_THREADS / 32; ++w) { const unsigned v = s_w[w]; s_w[w] = run; run += v; }
            const unsigned t1 = run & 0xffffu, t2 = run >> 16;
            s_base[0] = t1 ? atomicAdd((unsigned long long*)&C.st->n_front, (unsigned long long)t1) : 0ull;
            s_base[1] = t2 ? atomicAdd((unsigned long long*)&C.st->n_back, (unsigned long long)t2) : 0ull;
        }
        __syncthreads();
        const unsigned pre = s_w[warp] + inc - packed;
        long long at_f = (long long)s_base[0] + (pre & 0xffffu);
        long long at_b = (long long)s_base[1] + (pre >> 16);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            if (code[s] == 2 || code[s] == 3) {
                const long long k = code[s] == 2 ? at_f++ : at_b++;
                const long long pos = code[s] == 2 ? k : C.cap - 1 - k;
                if (k < C.cap) {
                    const long long g = tile * tile_groups + (s >> 2) * CL_THREADS + tid;
                    C.l_idx[pos] = (unsigned)(C.out_base + 4 * g + (s & 3));
                    C.l_cnt[pos] = cnt[s];
                    C.l_dist[pos] = dist[s];
                    C.l_bb[pos] = bb[s];
                }
            }
        }
        __syncthreads();           // s_w / s_base are reused by the next tile
    }
    // the last n_pairs % 4 records: lanes 0..2 of the first warp of CTA 0, one record each
    if (blockIdx.x == 0 && warp == 0) {
        const long long i = (n_groups << 2) + lane;
        const bool live = lane < (int)(P.n_pairs & 3);
        int k = 4, c = 0;
        unsigned ud = 0;
        double b = 1.0;
        if (live) {
            const int m1 = P.mid1[i], m2 = P.mid2[i];
            c = P.count[i];
            ud = (unsigned)m2 - (unsigned)m1;
            const bool inr = m2 >= m1 && (ud - lo_u) <= span_u;
            if (HAS_BIAS) {
                const BiasRow r1 = HAS_CHR ? bias_row(P, P.chr1[i]) : shard_row, r2 = HAS_CHR ? bias_row(P, P.chr2[i]) : shard_row;
                b = bias_lookup<true>(P, r1, m1) * bias_lookup<true>(P, r2, m2);
            }
            k = pre_class(inr, c, b, HAS_BIAS, exact);
            ones += k == 0; nans += k == 1;
            if (P.p) P.p[C.out_base + i] = k == 0 ? 1.0 : qnan;
            if (P.q) P.q[C.out_base + i] = k == 0 ? 1.0 : qnan;
            if (k == 2 || k == 3) {
                const long long kk = (long long)atomicAdd((unsigned long long*)(k == 2 ? &C.st->n_front : &C.st->n_back), 1ull);
                const long long pos = k == 2 ? kk : C.cap - 1 - kk;
                if (kk < C.cap) { C.l_idx[pos] = (unsigned)(C.out_base + i); C.l_cnt[pos] = c; C.l_dist[pos] = (int)ud; C.l_bb[pos] = b; }
            }
        }
    }
    const unsigned o = __reduce_add_sync(0xffffffffu, ones), zn = __reduce_add_sync(0xffffffffu, nans);
    if (lane == 0) {
        if (o) atomicAdd((unsigned long long*)&C.st->n_ones, (unsigned long long)o);
        if (zn) atomicAdd((unsigned long long*)&C.st->n_nan, (unsigned long long)zn);
    }
}

// after the fit: may the speculative classification stand?
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
        // count <= 0 rows were written as 1.0 when 0 <= b1*b2 <= 16 and as NaN when b1*b2 < 0: right iff every
        // prior splineY * b1*b2 of the first kind lies in [0, 1] and every one of the second kind is negative
        const bool ok = L > 0 && !bad && lo > 0.0 && hi * CL_BB_MAX <= 1.0;
        if (L > 0 && !ok) {
            st->exact = 1;
            st->n_front = 0; st->n_back = 0; st->n_ones = 0; st->n_nan = 0;      // K4a starts over, in exact mode
        }
    }
}

struct LsParams {
    const unsigned* l_idx; const int* l_cnt; const int* l_dist; const double* l_bb; long long cap;
    const BbkFitResult* fit; const double* spline_y; PvParams pv;               // pv: divisor only
    double* p; double* q; long long* p_hist;
    unsigned long long* c_keys; unsigned* c_idx; long long c_cap;
    BbkScoreState* st;
};

struct LsShared {
    double rcp[RCP_TAB];
    double lfact[LF_TAB];
    unsigned hist[BBK_PHIST_BINS];
};

__global__ void __launch_bounds__(PV_THREADS, 3) listed_kernel(LsParams Q) {
    __shared__ LsShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long S = Q.fit->S;
    const int k0 = Q.fit->k0, L = Q.fit->L;
    if (!(Q.fit->status == BBK_FIT_OK && L > 0)) return;                       // failed fit: the host raises, p is never read
    const long long nA = (long long)Q.st->n_front, nB = (long long)Q.st->n_back;
    if (nA + nB > Q.cap) { if (tid == 0 && blockIdx.x == 0) Q.st->overflow = 1; return; }
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
    const long long roundsA = (nA + 31) >> 5, rounds = roundsA + ((nB + 31) >> 5);
    for (long long r = (long long)blockIdx.x * PV_WARPS + warp; r < rounds; r += (long long)gridDim.x * PV_WARPS) {
        const bool isB = r >= roundsA;
        const long long k = ((isB ? r - roundsA : r) << 5) + lane;
        const bool active = k < (isB ? nB : nA);
        const long long pos = isB ? Q.cap - 1 - k : k;
        unsigned row = 0;
        int c = 0, cls = 0;
        double prior = 0.0, out = __longlong_as_double(0x7ff8000000000000ll);
        if (active) {
            row = Q.l_idx[pos]; c = Q.l_cnt[pos];
            const int d = Q.l_dist[pos];
            const double b = Q.l_bb[pos];
            prior = __ldg(&Q.spline_y[spline_index<true>(Q.pv, (long long)d, k0, L)]) * b;    // fithic.py:429-431
            cls = bdtrc_class(c, s_cap, s_fits, prior, &out);
        }
        if (cls == 1) out = -expm1(K.dn * log1m(prior));                        // bdtrc's closed form for k == 0
        TailState T;
        T.lp = 0.0; T.term = 0.0; T.sum = 1.0; T.a = 0.0; T.step = 0.0; T.j = 0;
        bool running = false, fast = false;
        if (cls == 2) {
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
        const double pv = finish_p(out);                                        // fithic.py:434
        if (active) {
            Q.p[row] = pv;
            if (Q.q) Q.q[row] = prefill_q(pv);
            if (Q.p_hist) hist_p(sh.hist, pv, ones, nans);
        }
        if (Q.c_keys) {
            const bool cand = active && pv < BBK_SMALL_P;
            const unsigned m = __ballot_sync(0xffffffffu, cand);
            if (m) {
                const int leader = __ffs(m) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd((unsigned long long*)&Q.st->n_cand, (unsigned long long)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (cand) {
                    const unsigned long long at = base + __popc(m & ((1u << lane) - 1));
                    if ((long long)at < Q.c_cap) { Q.c_keys[at] = bbk_key_of(pv); Q.c_idx[at] = row; }
                    else Q.st->cand_overflow = 1;
                }
            }
        }
    }
    if (Q.p_hist) {
        __syncthreads();
        for (int i = tid; i < BBK_PHIST_BINS; i += PV_THREADS) {
            const unsigned v = sh.hist[i];
            if (v) atomicAdd((unsigned long long*)&Q.p_hist[i], (unsigned long long)v);
        }
        unsigned long long o = __reduce_add_sync(0xffffffffu, ones), zn = __reduce_add_sync(0xffffffffu, nans);
        if (blockIdx.x == 0 && tid == 0) { o += Q.st->n_ones; zn += Q.st->n_nan; }        // the rows K4a finished
        if (lane == 0) {
            if (o) atomicAdd((unsigned long long*)&Q.p_hist[BBK_PHIST_BINS], o);
            if (zn) atomicAdd((unsigned long long*)&Q.p_hist[BBK_PHIST_BINS + 1], zn);
        }
    }
}

__global__ void score_begin_kernel(BbkScoreState* st, long long* p_hist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (p_hist && i < BBK_PHIST_LEN) p_hist[i] = 0;
    if (i == 0) { st->n_front = 0; st->n_back = 0; st->n_ones = 0; st->n_nan = 0; st->n_cand = 0;
                  st->overflow = 0; st->cand_overflow = 0; st->exact = 0; st->reserved = 0; }
}
