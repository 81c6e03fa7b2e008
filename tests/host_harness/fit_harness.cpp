// Host build of blueberry_b200/csrc/fit_stage.h for CPU-side unit tests (TEST INFRASTRUCTURE ONLY:
// the package never loads this library; the product runs the same source as an sm_100a kernel).
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../blueberry_b200/csrc/fit_coop.h"

extern "C" {

int th_equal_occupancy(const int64_t* possible, const int64_t* observed, int nkeys, int64_t S, int n_bins,
                       int64_t R, int64_t min_dist, int64_t max_dist, double* x, double* y, int max_out,
                       int32_t* bin_of_key, int* n_out) {
    std::vector<int32_t> bs(max_out), be(max_out);
    for (int k = 0; k < nkeys; ++k) bin_of_key[k] = -1;
    int st = bbk_eo_boundaries(observed, nkeys, S, n_bins, R, min_dist, max_dist, bs.data(), be.data(), max_out, n_out);
    if (st != BBK_FIT_OK) return st;
    for (int j = 0; j < *n_out; ++j) {
        st = bbk_eo_bin_stats(possible, observed, bs[j], be[j], S, R, &x[j], &y[j]);
        if (st != BBK_FIT_OK) return st;
        for (int k = bs[j]; k <= be[j]; ++k) bin_of_key[k] = j;
    }
    return BBK_FIT_OK;
}

// returns ier; t_out/c_out sized m+4
int th_univariate_spline(const double* x, const double* y, int m, double s, int* n_out, double* fp_out,
                         double* t_out, double* c_out) {
    std::vector<double> buf(bbk_spline_ws_doubles(m), 0.0);
    BbkSplineWs ws;
    bbk_spline_ws_carve(buf.data(), m, &ws);
    int ier = bbk_univariate_spline(x, y, m, s, n_out, fp_out, &ws);
    memcpy(t_out, ws.t, sizeof(double) * (m + 4));
    memcpy(c_out, ws.c, sizeof(double) * (m + 4));
    return ier;
}

// the block-cooperative schedule, threads emulated by loops (BBK_COOP_HOST_NT of them)
int th_coop_univariate_spline(const double* x, const double* y, int m, double s, int* n_out, double* fp_out,
                              double* t_out, double* c_out) {
    std::vector<double> buf(bbk_coop_ws_doubles(m), 0.0);
    BbkCoopWs cw;
    bbk_coop_ws_carve(buf.data(), m, &cw);
    BbkCoopState st;
    int ier = bbk_coop_univariate_spline(x, y, m, s, &st, &cw);
    *n_out = st.n;
    *fp_out = st.fp;
    memcpy(t_out, cw.w.t, sizeof(double) * (m + 4));
    memcpy(c_out, cw.w.c, sizeof(double) * (m + 4));
    return ier;
}

void th_spline_eval(const double* t, int n, const double* c, const double* args, int na, double* out) {
    int l = 4;
    for (int i = 0; i < na; ++i) out[i] = bbk_spline_eval(t, n, c, args[i], &l);
}

void th_antitonic(const double* v, int L, double* out) {
    std::vector<double> wm(L + 1), wc(L + 1);
    std::vector<int32_t> st(L + 2);
    bbk_antitonic_pava(v, L, out, wm.data(), wc.data(), st.data());
}

// the cut-into-pieces variant the kernel runs, `nt` emulated workers
void th_antitonic_segmented(const double* v, int L, int nt, double* out) {
    std::vector<double> wm(L + 1), wc(L + 1), cmax(nt), cmin(nt);
    std::vector<int32_t> flags(L + 2, 0);
    bbk_antitonic_pava_segmented(v, L, out, wm.data(), wc.data(), flags.data(), cmax.data(), cmin.data(), nt);
}

}
