// fit_stage.h - the O(D) "fit" stage of the Fit-Hi-C pass, written once for device and host.
//
//   * equal-occupancy binning            (reference: blueberry/fithic.py:160-227)
//   * cubic smoothing spline, s=min(y)^2 (reference: fithic.py:340-343 -> scipy UnivariateSpline,
//                                          i.e. Dierckx's CURFIT/FPCURF algorithm; restated here
//                                          from the published algorithm - P. Dierckx, "Curve and
//                                          Surface Fitting with Splines", OUP 1993, ch. 5 - since
//                                          FITPACK is a third-party dependency absent from the
//                                          reference tree)
//   * spline evaluation on the distance grid (fithic.py:350-359 -> FITPACK SPLEV)
//   * antitonic regression by PAVA       (fithic.py:361-362 -> sklearn IsotonicRegression)
//
// The product compiles this for sm_100a only (fit_stage.cu, one CTA, FP64, FMA contraction OFF so
// that the discrete decisions of the knot search see the same roundings as the CPU library).
// tests/host_harness builds the very same source with g++ so the algorithm can be checked against
// scipy without a GPU; that build is test infrastructure and is never loaded by the package.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define BBK_HD __host__ __device__ __forceinline__
#define BBK_HD_NOINLINE __host__ __device__ __noinline__
#else
#define BBK_HD inline
#define BBK_HD_NOINLINE
#endif

// ------------------------------------------------------------------------------------------------
// error codes of the fit stage (mirrored in include/bbk.h)
// ------------------------------------------------------------------------------------------------
#define BBK_FIT_OK 0
#define BBK_FIT_ZERO_PAIRS_BIN (-11)   // a bin with 0 possible pairs: reference raises ZeroDivisionError (fithic.py:216)
#define BBK_FIT_TOO_FEW_BINS (-12)     // fewer than 4 bins: scipy raises "m > k must hold"
#define BBK_FIT_TOO_MANY_BINS (-13)    // more bins than the workspace was sized for
#define BBK_FIT_S_ZERO (-14)           // S == 0: reference raises ZeroDivisionError (fithic.py:216)
#define BBK_FIT_X_NOT_INCREASING (-15) // scipy raises "x must be increasing"

// ------------------------------------------------------------------------------------------------
// Equal-occupancy binning, fithic.py:160-227.
//   possible/observed: per-distance tables (index = distance / R), nkeys entries
//   S: observedIntraInRangeSum.  n_bins: requested number of bins (more can come out, :208).
//   Outputs x[], y[] (<= max_out), bin_of_key[nkeys] (-1 = none), *n_out.
// Integer comparisons against `desired` follow Python: the first threshold is an int floor (:167),
// later ones are doubles (:209); int-vs-double comparison in Python is exact, and so is the
// comparison below because every integer involved is < 2^53.
// ------------------------------------------------------------------------------------------------
BBK_HD bool bbk_in_range(int64_t d, int64_t min_dist, int64_t max_dist) {   // fithic.py:445-449
    bool lo = (min_dist == -1) || (min_dist > -1 && d > min_dist);
    bool hi = (max_dist == -1) || (max_dist > -1 && d <= max_dist);
    return lo && hi;
}

// Phase A (sequential): bin boundaries.  bin j covers the in-range keys bin_start[j]..bin_end[j].
BBK_HD int bbk_eo_boundaries(const int64_t* observed, int nkeys, int64_t S, int n_bins, int64_t R,
                             int64_t min_dist, int64_t max_dist, int32_t* bin_start, int32_t* bin_end,
                             int max_out, int* n_out) {
    *n_out = 0;
    bool desired_is_int = true;
    int64_t desired_i = 0;
    if (n_bins != 0) desired_i = (S >= 0 || S % n_bins == 0) ? S / n_bins : S / n_bins - 1;   // floor (:167)
    double desired_d = 0.0;
    int64_t acc = 0, total = 0;
    int n = 0, start = -1, nout = 0;
    for (int k = 0; k < nkeys; ++k) {
        int64_t obs = observed[k];
        total += obs;                                                        // :183
        if (!bbk_in_range((int64_t)k * R, min_dist, max_dist)) continue;     // :184
        if (start < 0) start = k;
        bool full;
        if (desired_is_int) full = (obs >= desired_i) || (acc + obs >= desired_i);                  // :188,:194
        else                full = ((double)obs >= desired_d) || ((double)(acc + obs) >= desired_d);
        if (!full) { acc += obs; continue; }                                 // :199-201
        acc = 0;
        n += 1;                                                              // :206
        if (n < n_bins) {                                                    // :208-209
            desired_d = 1.0 * (double)(S - total) / (double)(n_bins - n);
            desired_is_int = false;
        }
        if (nout >= max_out) return BBK_FIT_TOO_MANY_BINS;
        bin_start[nout] = start;
        bin_end[nout] = k;
        nout += 1;
        start = -1;
    }
    *n_out = nout;
    return BBK_FIT_OK;
}

// Phase B (one bin, independent of the others): the bin's mean distance x and mean probability y.
// Members are summed in ascending distance like the reference loop (:211-214); only in-range keys
// are members (out-of-range keys inside [start, end] cannot exist: the range is one interval).
BBK_HD int bbk_eo_bin_stats(const int64_t* possible, const int64_t* observed, int start, int end, int64_t S,
                            int64_t R, double* x_out, double* y_out) {
    double n_pairs = 0.0, n_inter = 0.0, avg = 0.0;
    for (int b = start; b <= end; ++b) {
        n_pairs += (double)possible[b];
        n_inter += (double)observed[b];
        avg += 1.0 * (double)possible[b] * ((double)((int64_t)b * R) / 10000.0);
    }
    if (n_pairs == 0.0 || S == 0) return n_pairs == 0.0 ? BBK_FIT_ZERO_PAIRS_BIN : BBK_FIT_S_ZERO;
    *y_out = (n_inter / n_pairs) / (double)S;                                // :216
    *x_out = 10000.0 * (avg / n_pairs);                                      // :217
    return BBK_FIT_OK;
}

// ------------------------------------------------------------------------------------------------
// Smoothing spline (degree 3, unit weights): Dierckx's algorithm.
// Arrays are addressed 1-based through the macros below to keep the index arithmetic of the
// published algorithm recognisable; storage is row-major C.
// ------------------------------------------------------------------------------------------------
struct BbkSplineWs {
    double* t;       // [nest]   knots
    double* c;       // [nest]   B-spline coefficients
    double* fpint;   // [nest]   residual sum per knot interval
    double* z;       // [nest]   rotated right-hand side
    double* a;       // [nest*4] triangularised observation matrix (band)
    double* b;       // [nest*5] discontinuity jumps of the 3rd derivative
    double* g;       // [nest*5] work copy of a, extended
    double* q;       // [m*4]    B-spline values at the data points
    int32_t* nrdata; // [nest]   number of data points strictly inside each knot interval
};

BBK_HD size_t bbk_spline_ws_doubles(int m) {
    size_t nest = (size_t)m + 4;
    return nest * 4 /*t c fpint z*/ + nest * 4 + nest * 5 + nest * 5 + (size_t)m * 4 + (nest + 1) / 2 /*nrdata as int32*/ + 8;
}

BBK_HD void bbk_spline_ws_carve(double* base, int m, BbkSplineWs* ws) {
    size_t nest = (size_t)m + 4;
    ws->t = base;            base += nest;
    ws->c = base;            base += nest;
    ws->fpint = base;        base += nest;
    ws->z = base;            base += nest;
    ws->a = base;            base += nest * 4;
    ws->b = base;            base += nest * 5;
    ws->g = base;            base += nest * 5;
    ws->q = base;            base += (size_t)m * 4;
    ws->nrdata = (int32_t*)base;
}

#define T_(i) t[(i) - 1]
#define C_(i) c[(i) - 1]
#define Z_(i) z[(i) - 1]
#define X_(i) x[(i) - 1]
#define Y_(i) y[(i) - 1]
#define FPINT_(i) fpint[(i) - 1]
#define NRDATA_(i) nrdata[(i) - 1]
#define A_(i, j) a[((i) - 1) * 4 + ((j) - 1)]
#define B_(i, j) b[((i) - 1) * 5 + ((j) - 1)]
#define G_(i, j) g[((i) - 1) * 5 + ((j) - 1)]
#define Q_(i, j) q[((i) - 1) * 4 + ((j) - 1)]

// Non-zero cubic B-splines at x with t(l) <= x < t(l+1) (de Boor-Cox, stable recurrence).
BBK_HD void bbk_bspl3(const double* t, double x, int l, double* h /*[4], 1-based via h[i-1]*/) {
    double hh[3];
    h[0] = 1.0;
    for (int j = 1; j <= 3; ++j) {
        for (int i = 1; i <= j; ++i) hh[i - 1] = h[i - 1];
        h[0] = 0.0;
        for (int i = 1; i <= j; ++i) {
            int li = l + i, lj = li - j;
            if (T_(li) == T_(lj)) { h[i] = 0.0; continue; }
            double f = hh[i - 1] / (T_(li) - T_(lj));
            h[i - 1] = h[i - 1] + f * (T_(li) - x);
            h[i] = f * (x - T_(lj));
        }
    }
}

// Givens rotation parameters that annihilate piv against the diagonal element ww (updated in place).
BBK_HD void bbk_givens(double piv, double* ww, double* cs, double* sn) {
    double store = fabs(piv), dd;
    if (store >= *ww) dd = store * sqrt(1.0 + (*ww / piv) * (*ww / piv));
    else              dd = *ww * sqrt(1.0 + (piv / *ww) * (piv / *ww));
    *cs = *ww / dd;
    *sn = piv / dd;
    *ww = dd;
}

BBK_HD void bbk_rotate(double cs, double sn, double* a, double* b) {
    double s1 = *a, s2 = *b;
    *b = cs * s2 + sn * s1;
    *a = cs * s1 - sn * s2;
}

// Back substitution for an upper triangular band matrix of bandwidth k stored row-wise with row stride `ld`.
BBK_HD void bbk_backsub(const double* a, int ld, const double* z, int n, int k, double* c) {
    int k1 = k - 1;
    c[n - 1] = z[n - 1] / a[(n - 1) * ld];
    int i = n - 1;
    for (int j = 2; j <= n; ++j) {
        double store = z[i - 1];
        int i1 = (j <= k1) ? j - 1 : k1;
        int mm = i;
        for (int l = 1; l <= i1; ++l) {
            mm += 1;
            store = store - c[mm - 1] * a[(i - 1) * ld + l];
        }
        c[i - 1] = store / a[(i - 1) * ld];
        i -= 1;
    }
}

// Jumps of the 3rd derivative of the B-splines at the interior knots t(5..n-4), scaled by
// (nrint/(t(n-3)-t(4)))^3 per factor as in Dierckx (k2 = 5).
BBK_HD void bbk_discontinuity(const double* t, int n, double* b) {
    const int k2 = 5, k1 = 4, k = 3;
    int nk1 = n - k1, nrint = nk1 - k;
    double h[12];
    double an = (double)nrint;
    double fac = an / (T_(nk1 + 1) - T_(k1));
    for (int l = k2; l <= nk1; ++l) {
        int lmk = l - k1;
        for (int j = 1; j <= k1; ++j) {
            int ik = j + k1, lj = l + j, lk = lj - k2;
            h[j - 1] = T_(l) - T_(lk);
            h[ik - 1] = T_(l) - T_(lj);
        }
        int lp = lmk;
        for (int j = 1; j <= k2; ++j) {
            int jk = j;
            double prod = h[j - 1];
            for (int i = 1; i <= k; ++i) {
                jk += 1;
                prod = prod * h[jk - 1] * fac;
            }
            int lk = lp + k1;
            B_(lmk, j) = (T_(lk) - T_(lp)) / prod;
            lp += 1;
        }
    }
}

// Insert one knot in the interval with the largest residual sum (that still holds data points).
BBK_HD void bbk_add_knot(const double* x, double* t, int* n_io, double* fpint, int32_t* nrdata, int* nrint_io) {
    int n = *n_io, nrint = *nrint_io;
    int k = (n - nrint - 1) / 2;
    double fpmax = 0.0;
    int jbegin = 1, number = 0, maxpt = 0, maxbeg = 0;
    for (int j = 1; j <= nrint; ++j) {
        int jpoint = NRDATA_(j);
        if (!(fpmax >= FPINT_(j) || jpoint == 0)) {
            fpmax = FPINT_(j);
            number = j;
            maxpt = jpoint;
            maxbeg = jbegin;
        }
        jbegin = jbegin + jpoint + 1;
    }
    if (number == 0) {   // every candidate interval has zero residual or no interior point: pick the
        jbegin = 1;      // first interval that still holds a data point (keeps the search well defined)
        for (int j = 1; j <= nrint; ++j) {
            int jpoint = NRDATA_(j);
            if (jpoint != 0) { number = j; maxpt = jpoint; maxbeg = jbegin; break; }
            jbegin = jbegin + jpoint + 1;
        }
        if (number == 0) return;
    }
    int ihalf = maxpt / 2 + 1;
    int nrx = maxbeg + ihalf;
    int next = number + 1;
    if (next <= nrint) {
        for (int j = next; j <= nrint; ++j) {
            int jj = next + nrint - j;
            FPINT_(jj + 1) = FPINT_(jj);
            NRDATA_(jj + 1) = NRDATA_(jj);
            int jk = jj + k;
            T_(jk + 1) = T_(jk);
        }
    }
    NRDATA_(number) = ihalf - 1;
    NRDATA_(next) = maxpt - ihalf;
    double am = (double)maxpt;
    double an = (double)NRDATA_(number);
    FPINT_(number) = fpmax * an / am;
    an = (double)NRDATA_(next);
    FPINT_(next) = fpmax * an / am;
    int jk = next + k;
    T_(jk) = X_(nrx);
    *n_io = n + 1;
    *nrint_io = nrint + 1;
}

BBK_HD double bbk_rational_root(double* p1, double* f1, double p2, double f2, double* p3, double* f3) {
    double p;
    if (*p3 > 0.0) {
        double h1 = *f1 * (f2 - *f3), h2 = f2 * (*f3 - *f1), h3 = *f3 * (*f1 - f2);
        p = -(*p1 * p2 * h3 + p2 * *p3 * h1 + *p3 * *p1 * h2) / (*p1 * h1 + p2 * h2 + *p3 * h3);
    } else {
        p = (*p1 * (*f1 - *f3) * f2 - p2 * (f2 - *f3) * *f1) / ((*f1 - f2) * *f3);   // p3 = infinity
    }
    if (f2 < 0.0) { *p3 = p2; *f3 = f2; }
    else          { *p1 = p2; *f1 = f2; }
    return p;
}

// One run of the smoothing-spline search with storage for `nest` knots.
// Returns ier: 0 ok, -1 interpolating spline, -2 least-squares polynomial, 1 nest too small,
//              2 non-monotone f(p), 3 maxit reached.  n, t, c, fp are outputs.
BBK_HD_NOINLINE int bbk_smoothing_spline_run(const double* x, const double* y, int m, double s, int nest,
                                             int* n_out, double* fp_out, BbkSplineWs* ws) {
    const int k = 3, k1 = 4, k2 = 5, maxit = 20;
    const double tol = 0.001, con1 = 0.1, con9 = 0.9, con4 = 0.04, half = 0.5;
    double *t = ws->t, *c = ws->c, *fpint = ws->fpint, *z = ws->z, *a = ws->a, *b = ws->b, *g = ws->g, *q = ws->q;
    int32_t* nrdata = ws->nrdata;
    const double xb = x[0], xe = x[m - 1];
    const int nmin = 2 * k1;
    const double acc = tol * s;
    const int nmax = m + k1;
    int n, nplus = 0, nrint = 0, nk1 = 0, ier = 0;
    double fp = 0.0, fpold = 0.0, fp0 = 0.0, fpms = 0.0;
    double h[7];
    bool interpolate = false;

    if (s > 0.0) {
        n = nmin;
        fpold = 0.0;
        nplus = 0;
        NRDATA_(1) = m - 2;
    } else {
        n = nmax;
        if (nmax > nest) { *n_out = n; *fp_out = 0.0; return 1; }
        interpolate = true;
    }

    for (;;) {   // (re)entry point for "locate the knots as for interpolation"
        if (interpolate) {
            // k = 3 (odd): interior knots coincide with x(3..m-2)
            int mk1 = m - k1, i = k2, j = k / 2 + 2;
            for (int l = 1; l <= mk1; ++l) { T_(i) = X_(j); i += 1; j += 1; }
            interpolate = false;
        }
        bool restart = false;
        for (int iter = 1; iter <= m; ++iter) {
            if (n == nmin) ier = -2;
            nrint = n - nmin + 1;
            nk1 = n - k1;
            {
                int i = n;
                for (int j = 1; j <= k1; ++j) { T_(j) = xb; T_(i) = xe; i -= 1; }
            }
            // least-squares spline for the current knots: QR by Givens rotations, row by row
            fp = 0.0;
            for (int i = 1; i <= nk1; ++i) { Z_(i) = 0.0; for (int j = 1; j <= k1; ++j) A_(i, j) = 0.0; }
            int l = k1;
            for (int it = 1; it <= m; ++it) {
                double xi = X_(it), yi = Y_(it);
                while (!(xi < T_(l + 1) || l == nk1)) l += 1;
                bbk_bspl3(t, xi, l, h);
                for (int i = 1; i <= k1; ++i) Q_(it, i) = h[i - 1];
                int j = l - k1;
                for (int i = 1; i <= k1; ++i) {
                    j += 1;
                    double piv = h[i - 1];
                    if (piv == 0.0) continue;
                    double cs, sn;
                    bbk_givens(piv, &A_(j, 1), &cs, &sn);
                    bbk_rotate(cs, sn, &yi, &Z_(j));
                    if (i == k1) break;
                    int i2 = 1;
                    for (int i1 = i + 1; i1 <= k1; ++i1) {
                        i2 += 1;
                        bbk_rotate(cs, sn, &h[i1 - 1], &A_(j, i2));
                    }
                }
                fp = fp + yi * yi;
            }
            if (ier == -2) fp0 = fp;
            FPINT_(n) = fp0;
            FPINT_(n - 1) = fpold;
            NRDATA_(n) = nplus;
            bbk_backsub(a, 4, z, nk1, k1, c);
            fpms = fp - s;
            if (fabs(fpms) < acc) goto done;
            if (fpms < 0.0) goto smoothing;
            if (n == nmax) { ier = -1; goto done; }
            if (n == nest) { ier = 1; goto done; }
            if (ier != 0) {
                nplus = 1;
                ier = 0;
            } else {
                int npl1 = nplus * 2;
                double rn = (double)nplus;
                if (fpold - fp > acc) npl1 = (int)(rn * fpms / (fpold - fp));
                int mx = npl1 > nplus / 2 ? npl1 : nplus / 2;
                if (mx < 1) mx = 1;
                nplus = nplus * 2 < mx ? nplus * 2 : mx;
            }
            fpold = fp;
            // residual sum per knot interval t(j+k) <= x <= t(j+k+1)
            {
                double fpart = 0.0;
                int i = 1, newk = 0;
                l = k2;
                for (int it = 1; it <= m; ++it) {
                    if (!(X_(it) < T_(l) || l > nk1)) { newk = 1; l += 1; }
                    double term = 0.0;
                    int l0 = l - k2;
                    for (int j = 1; j <= k1; ++j) { l0 += 1; term = term + C_(l0) * Q_(it, j); }
                    term = (term - Y_(it)) * (term - Y_(it));
                    fpart = fpart + term;
                    if (newk == 0) continue;
                    double store = term * half;
                    FPINT_(i) = fpart - store;
                    i += 1;
                    fpart = store;
                    newk = 0;
                }
                FPINT_(nrint) = fpart;
            }
            for (int lk = 1; lk <= nplus; ++lk) {
                bbk_add_knot(x, t, &n, fpint, nrdata, &nrint);
                if (n == nmax) { interpolate = true; restart = true; break; }
                if (n == nest) break;
            }
            if (restart) break;
        }
        if (!restart) break;   // trial budget exhausted: fall through to the smoothing step like the published loop
    }

smoothing:
    if (ier == -2) goto done;
    {
        bbk_discontinuity(t, n, b);
        double p1 = 0.0, f1 = fp0 - s, p3 = -1.0, f3 = fpms, p = 0.0;
        for (int i = 1; i <= nk1; ++i) p = p + A_(i, 1);
        double rn = (double)nk1;
        p = rn / p;
        int ich1 = 0, ich3 = 0;
        int n8 = n - nmin;
        for (int iter = 1; iter <= maxit; ++iter) {
            double pinv = 1.0 / p;
            for (int i = 1; i <= nk1; ++i) {
                C_(i) = Z_(i);
                G_(i, k2) = 0.0;
                for (int j = 1; j <= k1; ++j) G_(i, j) = A_(i, j);
            }
            for (int it = 1; it <= n8; ++it) {
                for (int i = 1; i <= k2; ++i) h[i - 1] = B_(it, i) * pinv;
                double yi = 0.0;
                for (int j = it; j <= nk1; ++j) {
                    double piv = h[0], cs, sn;
                    bbk_givens(piv, &G_(j, 1), &cs, &sn);
                    bbk_rotate(cs, sn, &yi, &C_(j));
                    if (j == nk1) break;
                    int i2 = k1;
                    if (j > n8) i2 = nk1 - j;
                    for (int i = 1; i <= i2; ++i) {
                        int i1 = i + 1;
                        bbk_rotate(cs, sn, &h[i1 - 1], &G_(j, i1));
                        h[i - 1] = h[i1 - 1];
                    }
                    h[i2] = 0.0;
                }
            }
            bbk_backsub(g, 5, c, nk1, k2, c);
            fp = 0.0;
            int l = k2;
            for (int it = 1; it <= m; ++it) {
                if (!(X_(it) < T_(l) || l > nk1)) l += 1;
                int l0 = l - k2;
                double term = 0.0;
                for (int j = 1; j <= k1; ++j) { l0 += 1; term = term + C_(l0) * Q_(it, j); }
                fp = fp + (term - Y_(it)) * (term - Y_(it));
            }
            fpms = fp - s;
            if (fabs(fpms) < acc) goto done;
            if (iter == maxit) { ier = 3; goto done; }
            double p2 = p, f2 = fpms;
            if (ich3 == 0) {
                if (!((f2 - f3) > acc)) {          // initial p too large
                    p3 = p2; f3 = f2;
                    p = p * con4;
                    if (p <= p1) p = p1 * con9 + p2 * con1;
                    continue;
                }
                if (f2 < 0.0) ich3 = 1;
            }
            if (ich1 == 0) {
                if (!((f1 - f2) > acc)) {          // initial p too small
                    p1 = p2; f1 = f2;
                    p = p / con4;
                    if (p3 < 0.0) continue;
                    if (p >= p3) p = p2 * con1 + p3 * con9;
                    continue;
                }
                if (f2 > 0.0) ich1 = 1;
            }
            if (f2 >= f1 || f2 <= f3) { ier = 2; goto done; }
            p = bbk_rational_root(&p1, &f1, p2, f2, &p3, &f3);
        }
    }
done:
    *n_out = n;
    *fp_out = fp;
    return ier;
}

// UnivariateSpline(x, y, s=s) as scipy 1.18 drives it (_fitpack2.py:559-572): first with storage for
// max(m/2, 8) knots; when that is too small (ier == 1) the fit is redone with the maximal storage m+4.
BBK_HD int bbk_univariate_spline(const double* x, const double* y, int m, double s,
                                 int* n_out, double* fp_out, BbkSplineWs* ws) {
    int nest = m / 2 > 8 ? m / 2 : 8;
    int ier = bbk_smoothing_spline_run(x, y, m, s, nest, n_out, fp_out, ws);
    if (ier == 1) ier = bbk_smoothing_spline_run(x, y, m, s, m + 4, n_out, fp_out, ws);
    return ier;
}

// Spline value at arg (extrapolating with the end polynomial pieces, ext=0). *l_io carries the
// knot-interval cursor between calls (start with 4); arguments may come in any order.
BBK_HD double bbk_spline_eval(const double* t, int n, const double* c, double arg, int* l_io) {
    const int k1 = 4, k2 = 5;
    int nk1 = n - k1;
    int l = *l_io, l1 = l + 1;
    while (!(arg >= T_(l) || l1 == k2)) { l1 = l; l = l - 1; }
    while (!(arg < T_(l1) || l == nk1)) { l = l1; l1 = l + 1; }
    double h[4];
    bbk_bspl3(t, arg, l, h);
    double sp = 0.0;
    int ll = l - k1;
    for (int j = 1; j <= k1; ++j) { ll += 1; sp = sp + C_(ll) * h[j - 1]; }
    *l_io = l;
    return sp;
}

// Antitonic (non-increasing) least-squares regression with unit weights, pool-adjacent-violators.
// sklearn fits the increasing problem on the reversed sequence and reverses the result
// (sklearn/isotonic.py isotonic_regression(..., increasing=False) -> y[::-1]); block means are
// kept as (weighted-mean) running values, merged backwards, exactly in that order.
// v[L] in, out[L]; work arrays wmean[L], wcount[L] (doubles), start[L+1] (int32).
BBK_HD int bbk_antitonic_pava_blocks(const double* v, int L, double* wmean, double* wcount, int32_t* start) {
    // Pool-adjacent-violators on the reversed view u[i] = v[L-1-i] (non-decreasing fit), the current block
    // carried in registers.  Returns the number of blocks; block b covers u[start[b] .. start[b+1]-1].
    if (L <= 0) return 0;
    int b = 0;
    double xb_prev = v[L - 1], wb_prev = 1.0;
    wmean[0] = xb_prev;
    wcount[0] = 1.0;
    if (start) { start[0] = 0; start[1] = 1; }
    double nxt = L > 1 ? v[L - 2] : 0.0;          // element i of the reversed view, loaded one step ahead
    for (int i = 1; i < L; ++i) {
        b += 1;
        double xb = nxt, wb = 1.0;
        if (i + 1 < L) nxt = v[L - 2 - i];
        if (xb_prev >= xb) {
            // violation (or tie): pool with the previous block, then look ahead and behind
            b -= 1;
            double sb = wb_prev * xb_prev + wb * xb;
            wb += wb_prev;
            xb = sb / wb;
            while (i < L - 1 && xb >= nxt) {
                i += 1;
                sb += nxt;
                wb += 1.0;
                xb = sb / wb;
                if (i + 1 < L) nxt = v[L - 2 - i];
            }
            while (b > 0 && wmean[b - 1] >= xb) {
                b -= 1;
                sb += wcount[b] * wmean[b];
                wb += wcount[b];
                xb = sb / wb;
            }
        }
        wmean[b] = xb_prev = xb;
        wcount[b] = wb_prev = wb;
        if (start) start[b + 1] = i + 1;
    }
    return b + 1;
}

// The same regression cut into independent pieces.  In the reversed (non-decreasing) view u, pool-adjacent-
// violators never merges across a position where every value on the left is below every value on the right, so
// the pieces between such cuts can be run separately - each with exactly the operations, in exactly the order,
// that the one-pass algorithm would have spent on it - and a smooth, mostly monotone curve falls into thousands
// of one-element pieces plus a few wiggles.  A block mean can exceed the largest pooled value by accumulated
// rounding, so a cut needs a relative gap of 1e-9 (the rounding of <= 1e5 additions is below 2e-11); where the
// gap is smaller no cut is made, which is always exact.
// `nt` workers each own a contiguous chunk of u; phases are separated by barriers on the device (the host
// driver below runs the workers of a phase one after the other).
//   phase 1: chunk extrema                    -> cmax[t], cmin[t]
//   phase 2: flags[i] = 1 where a piece starts (uses wcount as scratch)
//   phase 3: every worker runs the pieces that START in its chunk and writes out[]
BBK_HD void bbk_pava_chunk(int L, int nt, int t, int* a, int* b) {
    const int per = (L + nt - 1) / nt;
    int lo = t * per, hi = lo + per;
    if (lo > L) lo = L;
    if (hi > L) hi = L;
    *a = lo; *b = hi;
}
BBK_HD void bbk_pava_phase1(const double* v, int L, int nt, int t, double* cmax, double* cmin) {
    int a, b;
    bbk_pava_chunk(L, nt, t, &a, &b);
    double mx = -INFINITY, mn = INFINITY;
    for (int i = a; i < b; ++i) { double u = v[L - 1 - i]; mx = u > mx ? u : mx; mn = u < mn ? u : mn; }
    cmax[t] = mx; cmin[t] = mn;
}
BBK_HD void bbk_pava_phase2(const double* v, int L, int nt, int t, const double* cmax, const double* cmin, double* wcount, int32_t* flags) {
    int a, b;
    bbk_pava_chunk(L, nt, t, &a, &b);
    if (a >= b) return;
    double pm = -INFINITY, sm = INFINITY;
    for (int k = 0; k < t; ++k) pm = cmax[k] > pm ? cmax[k] : pm;
    for (int k = t + 1; k < nt; ++k) sm = cmin[k] < sm ? cmin[k] : sm;
    for (int i = b - 1; i >= a; --i) {            // wcount[i] = min of u[i+1 ..]
        wcount[i] = sm;
        double u = v[L - 1 - i];
        sm = u < sm ? u : sm;
    }
    if (a == 0) flags[0] = 1;
    for (int i = a; i < b; ++i) {                 // pm = max of u[.. i]
        double u = v[L - 1 - i];
        pm = u > pm ? u : pm;
        if (i + 1 < L) {
            double right = wcount[i];
            flags[i + 1] = (right - pm > 1e-9 * (fabs(pm) + fabs(right))) ? 1 : 0;
        }
    }
}
BBK_HD void bbk_pava_phase3(const double* v, int L, int nt, int t, const int32_t* flags, double* wmean, double* wcount, double* out) {
    int a, b;
    bbk_pava_chunk(L, nt, t, &a, &b);
    for (int s = a; s < b; ++s) {
        if (!flags[s]) continue;
        int e = s + 1;
        while (e < L && !flags[e]) ++e;
        if (e == s + 1) { out[L - 1 - s] = v[L - 1 - s]; continue; }
        const int nb = bbk_antitonic_pava_blocks(v + (L - e), e - s, wmean + s, wcount + s, (int32_t*)0);
        int j = s;
        for (int blk = 0; blk < nb; ++blk) {
            const double val = wmean[s + blk];
            const int len = (int)wcount[s + blk];
            for (int k = 0; k < len; ++k, ++j) out[L - 1 - j] = val;
        }
    }
}
// host driver (tests/host_harness): cmax / cmin hold nt doubles each
BBK_HD void bbk_antitonic_pava_segmented(const double* v, int L, double* out, double* wmean, double* wcount, int32_t* flags,
                                         double* cmax, double* cmin, int nt) {
    for (int t = 0; t < nt; ++t) bbk_pava_phase1(v, L, nt, t, cmax, cmin);
    for (int t = 0; t < nt; ++t) bbk_pava_phase2(v, L, nt, t, cmax, cmin, wcount, flags);
    for (int t = 0; t < nt; ++t) bbk_pava_phase3(v, L, nt, t, flags, wmean, wcount, out);
}

BBK_HD void bbk_antitonic_pava(const double* v, int L, double* out, double* wmean, double* wcount, int32_t* start) {
    int nb = bbk_antitonic_pava_blocks(v, L, wmean, wcount, start);
    for (int blk = 0; blk < nb; ++blk)
        for (int j = start[blk]; j < start[blk + 1]; ++j) out[L - 1 - j] = wmean[blk];
}

#undef T_
#undef C_
#undef Z_
#undef X_
#undef Y_
#undef FPINT_
#undef NRDATA_
#undef A_
#undef B_
#undef G_
#undef Q_
