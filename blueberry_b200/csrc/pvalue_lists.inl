// pvalue_lists.inl - K4 split around the fit (included by pvalue.cu inside its anonymous namespace).
//
// The scoring loop of fithic.py:413-435 needs the spline (hence the whole distance table) only for the prior of a record;
// everything else - the distance, the range test (:427), the two bias gathers (:418-425) and their product, the count - is
// known as soon as the records are on the device.  And most records need no arithmetic at all: a zero count gives p = 1
// (or NaN when the prior is outside [0, 1], which bdtrc checks first).  So the work is cut in two, tile by tile (a tile =
// BBK_TILE_ROWS consecutive rows of one shard):
//
//   K4a classify_kernel (streams every record once, on a side stream WHILE the one-CTA fit kernel runs).  Per tile:
//       out of range                         -> NaN bit set
//       count <= 0, 0 <= b1*b2 <= 16         -> p will be 1.0    (speculative: needs 0 < splineY, 16 max(splineY) <= 1)
//       count <= 0, b1*b2 < 0 or NaN         -> NaN bit set      (speculative: needs splineY > 0)
//       count == 1                           -> work-list entry, first part of the tile's block
//       2 <= count <= SMALL_C                -> work-list entry, second part
//       everything else                      -> work-list entry, third part
//     A work-list entry is (row, count, distance, b1*b2): 20 bytes, all the fit-independent state of the record.  A tile's
//     entries are staged in shared memory and written as ONE contiguous block (one global atomic per tile reserves it);
//     the tile directory says where.  K4a writes neither p nor q: 12 B/pair in, one bit per pair + the entries out.
//   score_guard_kernel (after the fit): checks the two conditions on splineY.  If one fails it raises
//       BbkScoreState::exact and K4a is run again, this time sending EVERY in-range record to the list (the launch is
//       always enqueued and returns at once when the flag is down), so the result is exact in every case.
//   K4b scored_tiles_kernel: a CTA owns a tile.  It rebuilds the tile's p column in shared memory from the NaN bits
//       (1.0 / NaN), scores the tile's entries - dense rounds of 32 handed out by a shared-memory counter, every lane
//       busy on the same branch: prior = splineY[i] * (b1*b2), bdtrc's case analysis, then
//         count == 1            1 - (1-q)^S                                              (bdtrc's own closed form)
//         2 <= count <= SMALL_C 1 - pmf(0) (1 + r1 + r1 r2 + ...), count-1 terms         (the LOWER tail: a handful of
//                               terms instead of a 16..32-term upper sum; when it comes out below 1e-4 the subtraction
//                               has cost digits and the lane takes the upper sum after all - the rare significant rows)
//         the rest              pmf(c) (1 + r(c+1) + r(c+1) r(c+2) + ...), the upper tail as bbk_pvalues sums it -
//       drops the results into the column and streams the whole column out with coalesced 128-bit stores, p and
//       q = 1.0 / NaN.  No partial sector is ever written (scattering 8-byte results into a column written earlier
//       costs a DRAM read-modify-write per row: 13 ms instead of 4 on BASELINE config 3).  It also fills the coarse
//       p histogram and appends the rows with p < BBK_SMALL_P to the q-value step's candidate list.

constexpr int CL_THREADS = 256;
constexpr int TILE_ROWS = BBK_TILE_ROWS;       // 8 records per thread
constexpr int TILE_WORDS = TILE_ROWS / 32;     // NaN-bit words per tile
constexpr double CL_BB_MAX = 16.0;             // biases are in [0.5, 2] (fithic.py:147-149): products above 16 are deferred
constexpr int SMALL_C = 8;                     // counts up to here are scored through the lower tail (count - 1 terms)
constexpr double LOWER_MIN_P = 1e-4;           // below this the lower-tail form has lost digits: upper sum instead
static_assert(TILE_ROWS == CL_THREADS * 8, "a thread classifies two groups of four records per tile");

struct ClsParams {
    PvParams pv;                                // records, bias table, range, divisor (fit / spline / p / q unused)
    long long out_base;                         // row of record 0 in the rank-local p / q buffers (multiple of 4)
    long long tile_base;                        // index of the shard's first tile in the directory / bit map
    unsigned* l_row; int* l_cnt; int* l_dist; double* l_bb; long long cap;
    BbkTileDir* dir; unsigned* nan_bits;
    BbkScoreState* st;
    int exact_only;                             // 1: the post-fit relaunch (runs only when st->exact is up)
};

// class of one record before the fit.  0: p = 1.0, 1: p = NaN, 2: list, count == 1, 3: list, small counts, 5: list, the rest
__device__ __forceinline__ int pre_class(bool in_range, int c, double bb, bool has_bias, bool exact) {
    if (!in_range) return 1;
    if (c == 1) return 2;
    if (c >= 2 && c <= SMALL_C) return 3;
    if (c > SMALL_C || exact) return 5;
    if (!has_bias) return 0;
    if (bb >= 0.0 && bb <= CL_BB_MAX) return 0;
    if (bb < 0.0 || isnan(bb)) return 1;
    return 5;
}

struct ClsShared {
    double bb[TILE_ROWS];
    unsigned row[TILE_ROWS];
    int cnt[TILE_ROWS];
    int dist[TILE_ROWS];
    unsigned wsum[CL_THREADS / 32], wsum3[CL_THREADS / 32];
    unsigned tot, tot3;
    unsigned long long base;
};

// the shard's own bias row in 32-bit arithmetic (coordinates are int32 >= 0 on this path): where the entry of `mid` is,
// and whether there is one (fithic.py:418-425: on the grid, inside the table)
struct BiasRow32 { const double* tab; unsigned mid0, span; bool usable; };
__device__ __forceinline__ bool bias_index32(const FastDiv& div, const BiasRow32& r, int mid, unsigned* at) {
    const unsigned off = (unsigned)mid - r.mid0;                 // mid < mid0 (or mid < 0) wraps far above span
    const unsigned idx = fastdiv31(off & 0x7fffffffu, div);
    const bool ok = off < r.span && idx * div.R == off;
    *at = ok ? idx : 0u;
    return ok;
}

template <bool HAS_CHR, bool HAS_BIAS>
__global__ void __launch_bounds__(CL_THREADS, 3) classify_kernel(ClsParams C) {
    if (C.exact_only && C.st->exact == 0) return;
    const bool exact = C.exact_only != 0;
    __shared__ ClsShared sh;
    const PvParams& P = C.pv;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    BiasRow shard_row = {0, 0, 0, 0};
    BiasRow32 row32 = {nullptr, 0u, 0u, false};
    if (HAS_BIAS && !HAS_CHR) {
        shard_row = bias_row(P, P.shard_chrom);
        if (shard_row.nloc > 0 && shard_row.mid0 >= 0 && shard_row.mid0 < (1ll << 31) && shard_row.span < (1ull << 31)) {
            row32.tab = P.bias + shard_row.base; row32.mid0 = (unsigned)shard_row.mid0; row32.span = (unsigned)shard_row.span; row32.usable = true;
        }
    }
    const bool use32 = row32.usable;
    const long long n = P.n_pairs;
    const long long n_tiles = (n + TILE_ROWS - 1) / TILE_ROWS;
    const int4* m1v = reinterpret_cast<const int4*>(P.mid1);
    const int4* m2v = reinterpret_cast<const int4*>(P.mid2);
    const int4* cv = reinterpret_cast<const int4*>(P.count);
    const int4* c1v = reinterpret_cast<const int4*>(P.chr1);
    const int4* c2v = reinterpret_cast<const int4*>(P.chr2);
    const unsigned lo_u = (unsigned)P.min_dist, span_u = (unsigned)(P.max_dist - P.min_dist);
    unsigned long long n_one = 0, n_small = 0, n_other = 0, n_final = 0;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int cnt[8], dist[8], code[8];
        double bb[8];
        unsigned n1 = 0, n2 = 0, n3 = 0;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long g = tile * (TILE_ROWS / 4) + u * CL_THREADS + tid;      // group of four rows
            const long long r0 = 4 * g;
            const bool full = r0 + 3 < n;
            const int4 z4 = make_int4(0, 0, 0, 0);
            int4 a1 = z4, a2 = z4, ac = z4, x1 = z4, x2 = z4;
            if (full) {
                a1 = ld_stream_int4(m1v + g); a2 = ld_stream_int4(m2v + g); ac = ld_stream_int4(cv + g);
                if (HAS_CHR) { x1 = ld_stream_int4(c1v + g); x2 = ld_stream_int4(c2v + g); }
            } else if (r0 < n) {                                                      // the shard's last, partial group
                int t1[4] = {0, 0, 0, 0}, t2[4] = {0, 0, 0, 0}, tc[4] = {0, 0, 0, 0}, y1[4] = {0, 0, 0, 0}, y2[4] = {0, 0, 0, 0};
                for (int e = 0; e < 4 && r0 + e < n; ++e) {
                    t1[e] = P.mid1[r0 + e]; t2[e] = P.mid2[r0 + e]; tc[e] = P.count[r0 + e];
                    if (HAS_CHR) { y1[e] = P.chr1[r0 + e]; y2[e] = P.chr2[r0 + e]; }
                }
                a1 = make_int4(t1[0], t1[1], t1[2], t1[3]); a2 = make_int4(t2[0], t2[1], t2[2], t2[3]);
                ac = make_int4(tc[0], tc[1], tc[2], tc[3]);
                x1 = make_int4(y1[0], y1[1], y1[2], y1[3]); x2 = make_int4(y2[0], y2[1], y2[2], y2[3]);
            }
            const int m1s[4] = {a1.x, a1.y, a1.z, a1.w}, m2s[4] = {a2.x, a2.y, a2.z, a2.w}, cs[4] = {ac.x, ac.y, ac.z, ac.w};
            const int c1s[4] = {x1.x, x1.y, x1.z, x1.w}, c2s[4] = {x2.x, x2.y, x2.z, x2.w};
            // where the bias entries are, then all the loads, then the products: the gathers of a group are in flight together.
            // Neighbours in memory usually share their first locus (row-major input): one gather serves the group then.
            const bool same1 = !HAS_CHR && ((a1.x == a1.y) & (a1.y == a1.z) & (a1.z == a1.w));
            bool ok1[4], ok2[4];
            double v1[4], v2[4];
            if (HAS_BIAS && !HAS_CHR && use32) {
                unsigned i1[4], i2[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    ok1[e] = false; i1[e] = 0;
                    if (e == 0 || !same1) ok1[e] = bias_index32(P.div, row32, m1s[e], &i1[e]);
                    ok2[e] = bias_index32(P.div, row32, m2s[e], &i2[e]);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    v1[e] = 1.0;
                    if (e == 0 || !same1) v1[e] = __ldg(row32.tab + i1[e]);
                    v2[e] = __ldg(row32.tab + i2[e]);
                }
            } else {
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
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    v1[e] = 1.0; v2[e] = 1.0;
                    if (HAS_BIAS && r0 < n) {
                        if (e == 0 || !same1) v1[e] = __ldg(&P.bias[at1[e]]);
                        v2[e] = __ldg(&P.bias[at2[e]]);
                    }
                }
            }
            bool nanbit[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int s = u * 4 + e;
                const bool live = r0 + e < n;
                const unsigned ud = (unsigned)m2s[e] - (unsigned)m1s[e];                    // fithic.py:416
                const bool inr = live && m2s[e] >= m1s[e] && (ud - lo_u) <= span_u;         // fithic.py:427 (inclusive on both sides)
                double b = 1.0;
                if (HAS_BIAS) {
                    const double b1 = same1 ? bias_value(v1[0], ok1[0]) : bias_value(v1[e], ok1[e]);
                    b = b1 * bias_value(v2[e], ok2[e]);                                      // (bias1 * bias2) of :431
                }
                const int k = live ? pre_class(inr, cs[e], b, HAS_BIAS, exact) : 4;          // 4: no such row
                code[s] = k;
                cnt[s] = cs[e]; dist[s] = (int)ud; bb[s] = b;
                n1 += k == 2; n2 += k == 3; n3 += k == 5;
                n_final += k <= 1;
                nanbit[e] = k == 1 || k == 4;
            }
            // one bit per row: NaN (or no row).  Word e of this warp and half-tile holds element e of every lane's group
            const unsigned w0 = __ballot_sync(0xffffffffu, nanbit[0]), w1 = __ballot_sync(0xffffffffu, nanbit[1]);
            const unsigned w2 = __ballot_sync(0xffffffffu, nanbit[2]), w3 = __ballot_sync(0xffffffffu, nanbit[3]);
            if (lane == 0)
                reinterpret_cast<uint4*>(C.nan_bits + (C.tile_base + tile) * TILE_WORDS)[u * (CL_THREADS / 32) + warp] = make_uint4(w0, w1, w2, w3);
        }
        // positions inside the tile's block: count == 1, then the small counts, then the rest; scan inside the warp, then over the warps
        const unsigned packed = n1 | (n2 << 16);
        unsigned inc = packed, inc3 = n3;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, inc, o), y3 = __shfl_up_sync(0xffffffffu, inc3, o);
            if (lane >= o) { inc += y; inc3 += y3; }
        }
        if (lane == 31) { sh.wsum[warp] = inc; sh.wsum3[warp] = inc3; }
        __syncthreads();
        if (tid == 0) {
            unsigned run = 0, run3 = 0;
#pragma unroll
            for (int w = 0; w < CL_THREADS / 32; ++w) {
                const unsigned v = sh.wsum[w], v3 = sh.wsum3[w];
                sh.wsum[w] = run; sh.wsum3[w] = run3;
                run += v; run3 += v3;
            }
            sh.tot = run; sh.tot3 = run3;
        }
        __syncthreads();
        const unsigned tot = sh.tot, tA = tot & 0xffffu, tS = tot >> 16, tO = sh.tot3, tAll = tA + tS + tO;
        const unsigned pre = sh.wsum[warp] + inc - packed;
        unsigned at_a = pre & 0xffffu, at_s = tA + (pre >> 16), at_o = tA + tS + sh.wsum3[warp] + inc3 - n3;
        if (tid == 0) {
            unsigned long long base = 0;
            bool fits = true;
            if (tAll) {
                base = atomicAdd((unsigned long long*)&C.st->n_list, (unsigned long long)tAll);
                if ((long long)(base + tAll) > C.cap) { C.st->overflow = 1; fits = false; }
            }
            sh.base = fits ? base : ~0ull;
            BbkTileDir d;
            d.base = fits ? base : 0ull;
            d.n_one = fits ? tA : 0u; d.n_small = fits ? tS : 0u; d.n_other = fits ? tO : 0u;
            d.row_base = (unsigned)(C.out_base + tile * TILE_ROWS);
            const long long left = n - tile * TILE_ROWS;
            d.n_rows = (unsigned)(left < TILE_ROWS ? left : TILE_ROWS);
            d.pad = 0;
            C.dir[C.tile_base + tile] = d;
            if (fits) { n_one += tA; n_small += tS; n_other += tO; }
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            if (code[s] == 2 || code[s] == 3 || code[s] == 5) {
                const unsigned k = code[s] == 2 ? at_a++ : (code[s] == 3 ? at_s++ : at_o++);
                sh.row[k] = (unsigned)(C.out_base + tile * TILE_ROWS + (s >> 2) * (CL_THREADS * 4) + tid * 4 + (s & 3));
                sh.cnt[k] = cnt[s];
                sh.dist[k] = dist[s];
                sh.bb[k] = bb[s];
            }
        }
        __syncthreads();
        const unsigned long long base = sh.base;
        if (base != ~0ull) {                                         // (an overflowed tile keeps nothing)
            for (unsigned i = tid; i < tAll; i += CL_THREADS) {
                C.l_row[base + i] = sh.row[i];
                C.l_cnt[base + i] = sh.cnt[i];
                C.l_dist[base + i] = sh.dist[i];
                C.l_bb[base + i] = sh.bb[i];
            }
        }
        __syncthreads();           // the staging arrays are reused by the next tile
    }
    n_final = (unsigned long long)warp_sum_ll((long long)n_final);
    if (lane == 0 && n_final) atomicAdd((unsigned long long*)&C.st->n_final, n_final);
    if (tid == 0) {
        if (n_one) atomicAdd((unsigned long long*)&C.st->n_one, n_one);
        if (n_small) atomicAdd((unsigned long long*)&C.st->n_small, n_small);
        if (n_other) atomicAdd((unsigned long long*)&C.st->n_other, n_other);
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
        // count <= 0 rows were taken as 1.0 when 0 <= b1*b2 <= 16 and as NaN when b1*b2 < 0: right iff every
        // prior splineY * b1*b2 of the first kind lies in [0, 1] and every one of the second kind is negative
        const bool ok = L > 0 && !bad && lo > 0.0 && hi * CL_BB_MAX <= 1.0;
        if (L > 0 && !ok) {
            st->exact = 1;
            st->n_list = 0; st->n_one = 0; st->n_small = 0; st->n_other = 0; st->n_final = 0;      // K4a starts over, in exact mode
        }
    }
}

struct TlParams {
    const unsigned* l_row; const int* l_cnt; const int* l_dist; const double* l_bb; long long cap;
    const BbkTileDir* dir; const unsigned* nan_bits; long long n_tiles;
    const BbkFitResult* fit; const double* spline_y; PvParams pv;               // pv: divisor only
    double* p; double* q; long long* p_hist;
    unsigned long long* c_keys; unsigned* c_idx; long long c_cap;
    BbkScoreState* st;
};

struct TlShared {
    double col[TILE_ROWS];                     // the tile's p column
    double rcp[RCP_TAB];
    double lfact[LF_TAB];
    unsigned hist[BBK_PHIST_BINS];
    unsigned next;                             // next round of the current tile to hand out
};

__global__ void __launch_bounds__(PV_THREADS, 3) scored_tiles_kernel(TlParams Q) {
    extern __shared__ __align__(16) unsigned char tl_raw[];
    TlShared& sh = *reinterpret_cast<TlShared*>(tl_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long S = Q.fit->S;
    const int k0 = Q.fit->k0, L = Q.fit->L;
    if (!(Q.fit->status == BBK_FIT_OK && L > 0)) return;                       // failed fit: the host raises, p is never read
    if (Q.st->overflow) return;                                                // the host repeats the pass with a larger list
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
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    unsigned ones = 0, nans = 0;                 // list rows that came out as exactly 1.0 / NaN
    long long adj_ones = 0, adj_nans = 0;        // the rows classify finished, from the bit map (lane 0 of every warp)
    for (long long tile = blockIdx.x; tile < Q.n_tiles; tile += gridDim.x) {
        const BbkTileDir D = Q.dir[tile];
        if (tid == 0) sh.next = 0;
        // ---- the column as classify left it: 1.0, or NaN where the bit is set (list rows are overwritten below)
        const uint4* bits = reinterpret_cast<const uint4*>(Q.nan_bits + tile * TILE_WORDS);
        unsigned tile_nan = 0;                   // NaN bits of this warp's two words (every lane holds the same count)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint4 w = bits[u * (PV_THREADS / 32) + warp];
            tile_nan += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
            double2* c2 = reinterpret_cast<double2*>(&sh.col[u * (PV_THREADS * 4) + tid * 4]);
            c2[0] = make_double2(((w.x >> lane) & 1u) ? qnan : 1.0, ((w.y >> lane) & 1u) ? qnan : 1.0);
            c2[1] = make_double2(((w.z >> lane) & 1u) ? qnan : 1.0, ((w.w >> lane) & 1u) ? qnan : 1.0);
        }
        // rows classify finished (never touched again): NaN = set bits minus the rows the tile does not have, 1.0 = the
        // other non-list rows.  Signed: a warp's share can be negative, the sum over the CTA's warps is not.
        if (Q.p_hist && lane == 0) {
            adj_nans += (long long)tile_nan;
            adj_ones -= (long long)tile_nan;
            if (warp == 0) {
                const long long pad = (long long)TILE_ROWS - (long long)D.n_rows;
                adj_nans -= pad;
                adj_ones += (long long)D.n_rows - (long long)(D.n_one + D.n_small + D.n_other) + pad;
            }
        }
        __syncthreads();
        // ---- the tile's entries, 32 per round, the expensive kind first; a warp takes the next round when it is done with
        // one, and has the entries of its NEXT round in flight while it computes (their first use waited ~1 us on DRAM)
        const unsigned nL = D.n_one + D.n_small, nO = D.n_other;             // [count == 1 | 2..SMALL_C] is one list for the rounds
        const unsigned rO = (nO + 31) >> 5, rounds = rO + ((nL + 31) >> 5);
        unsigned r_next = 0;
        if (lane == 0) r_next = atomicAdd(&sh.next, 1u);
        r_next = __shfl_sync(0xffffffffu, r_next, 0);
        unsigned e_row = 0; int e_c = 0, e_d = 0; double e_b = 0.0; bool e_act = false;
        if (r_next < rounds) {
            const bool big = r_next < rO;
            const unsigned k = ((big ? r_next : r_next - rO) << 5) + lane;
            e_act = k < (big ? nO : nL);
            if (e_act) { const unsigned long long pos = D.base + (big ? nL : 0u) + k;
                         e_row = Q.l_row[pos]; e_c = Q.l_cnt[pos]; e_d = Q.l_dist[pos]; e_b = Q.l_bb[pos]; }
        }
        while (r_next < rounds) {
            const bool big = r_next < rO;                                       // warp-uniform
            const unsigned row = e_row;
            const int c = e_c, d = e_d;
            const double b = e_b;
            const bool active = e_act;
            // next round: its number now, its entries in flight during this round's arithmetic
            if (lane == 0) r_next = atomicAdd(&sh.next, 1u);
            r_next = __shfl_sync(0xffffffffu, r_next, 0);
            e_act = false;
            if (r_next < rounds) {
                const bool nbig = r_next < rO;
                const unsigned k = ((nbig ? r_next : r_next - rO) << 5) + lane;
                e_act = k < (nbig ? nO : nL);
                if (e_act) { const unsigned long long pos = D.base + (nbig ? nL : 0u) + k;
                             e_row = Q.l_row[pos]; e_c = Q.l_cnt[pos]; e_d = Q.l_dist[pos]; e_b = Q.l_bb[pos]; }
            }
            int cls = 0;
            double prior = 0.0, out = qnan;
            if (active) {
                prior = __ldg(&Q.spline_y[spline_index<true>(Q.pv, (long long)d, k0, L)]) * b;    // fithic.py:429-431
                cls = bdtrc_class(c, s_cap, s_fits, prior, &out);
            }
            bool upper = cls == 2;
            bool closed = cls == 1;
            if (!big) {
                // counts 1 .. SMALL_C through the lower tail: P(X >= c) = 1 - pmf(0) (1 + r1 + r1 r2 + ...), c - 1 terms,
                // r_i = (S - i + 1) q / (i (1 - q)).  count == 1 is the same form without terms (bdtrc's 1 - (1-q)^S).
                const bool lower = cls != 0 && c >= 1 && c <= SMALL_C;
                const int cmax = __reduce_max_sync(0xffffffffu, lower ? c : 0);
                if (cmax > 0) {
                    const double e0 = exp(K.dn * log1m(prior));
                    double sum = 1.0;
                    if (cmax > 1) {
                        const double qr = prior / (1.0 - prior);
                        double a = K.dn * qr, t = 1.0;
#pragma unroll
                        for (int i = 1; i < SMALL_C; ++i) {
                            if (i < cmax) {                                      // warp-uniform
                                if (i < c) { t *= a * (1.0 / (double)i); sum += t; a -= qr; }
                            }
                        }
                    }
                    const double pc = 1.0 - e0 * sum;
                    if (lower && pc >= LOWER_MIN_P) { out = pc > 1.0 ? 1.0 : pc; upper = false; closed = false; }   // else: digits lost
                }
            }
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
            const double pv = finish_p(out);                                    // fithic.py:434
            if (active) {
                sh.col[row - D.row_base] = pv;
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
        __syncthreads();
        // ---- the whole column out: coalesced 128-bit stores of p and of q = 1.0 / NaN
        const unsigned rows4 = (D.n_rows + 3u) & ~3u;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const unsigned r0 = u * (PV_THREADS * 4) + tid * 4;
            if (r0 < rows4) {
                const double2* c2 = reinterpret_cast<const double2*>(&sh.col[r0]);
                const double2 r01 = c2[0], r23 = c2[1];
                double2* pv2 = reinterpret_cast<double2*>(Q.p + D.row_base + r0);
                st_stream_double2(pv2, r01);
                st_stream_double2(pv2 + 1, r23);
                if (Q.q) {
                    double2* qv2 = reinterpret_cast<double2*>(Q.q + D.row_base + r0);
                    st_stream_double2(qv2, make_double2(prefill_q(r01.x), prefill_q(r01.y)));
                    st_stream_double2(qv2 + 1, make_double2(prefill_q(r23.x), prefill_q(r23.y)));
                }
            }
        }
        __syncthreads();           // the column and the round counter are reused by the next tile
    }
    if (Q.p_hist) {
        __syncthreads();
        for (int i = tid; i < BBK_PHIST_BINS; i += PV_THREADS) {
            const unsigned v = sh.hist[i];
            if (v) atomicAdd((unsigned long long*)&Q.p_hist[i], (unsigned long long)v);
        }
        const unsigned o = __reduce_add_sync(0xffffffffu, ones), zn = __reduce_add_sync(0xffffffffu, nans);
        if (lane == 0) {
            const long long to = (long long)o + adj_ones, tn = (long long)zn + adj_nans;     // two's complement: the sums are exact
            if (to) atomicAdd((unsigned long long*)&Q.p_hist[BBK_PHIST_BINS], (unsigned long long)to);
            if (tn) atomicAdd((unsigned long long*)&Q.p_hist[BBK_PHIST_BINS + 1], (unsigned long long)tn);
        }
    }
}

__global__ void score_begin_kernel(BbkScoreState* st, long long* p_hist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (p_hist && i < BBK_PHIST_LEN) p_hist[i] = 0;
    if (i == 0) { st->n_list = 0; st->n_one = 0; st->n_small = 0; st->n_other = 0; st->n_final = 0; st->n_cand = 0;
                  st->overflow = 0; st->cand_overflow = 0; st->exact = 0; st->reserved = 0; }
}
