// io_host.cpp - libbbkio.so: the significances writer of include/bbk_io.h (host only).
#include <charconv>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>

#include "../../include/bbk_io.h"

namespace {

thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

inline char* put_int(char* o, long long v) {
    if (v < 0) { *o++ = '-'; }
    unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
    char tmp[24];
    int k = 0;
    do { tmp[k++] = (char)('0' + u % 10); u /= 10; } while (u);
    while (k) *o++ = tmp[--k];
    return o;
}

// Python / numpy float64 str: shortest round-trip digits; positional when the decimal point position is in (-4, 16]
int format_double(double x, char* out) {
    char* o = out;
    if (x != x) { memcpy(o, "nan", 3); return 3; }
    if (x < 0 || (x == 0 && std::signbit(x))) { *o++ = '-'; x = -x; }
    if (x == 0) { memcpy(o, "0.0", 3); return (int)(o - out) + 3; }
    if (x > 1.7976931348623157e308) { memcpy(o, "inf", 3); return (int)(o - out) + 3; }
    char sci[40];
    auto r = std::to_chars(sci, sci + sizeof(sci), x, std::chars_format::scientific);     // d[.ddd]e[+-]XX, shortest
    char digits[24];
    int nd = 0;
    const char* s = sci;
    for (; s < r.ptr && *s != 'e'; ++s) if (*s != '.') digits[nd++] = *s;
    int e10 = 0;
    {
        ++s;                                     // 'e'
        bool neg = *s == '-';
        ++s;
        for (; s < r.ptr; ++s) e10 = e10 * 10 + (*s - '0');
        if (neg) e10 = -e10;
    }
    const int decpt = e10 + 1;                   // digits * 10^(decpt - nd)
    if (decpt > -4 && decpt <= 16) {
        if (decpt <= 0) {
            *o++ = '0'; *o++ = '.';
            for (int i = 0; i < -decpt; ++i) *o++ = '0';
            memcpy(o, digits, nd); o += nd;
        } else if (decpt >= nd) {
            memcpy(o, digits, nd); o += nd;
            for (int i = nd; i < decpt; ++i) *o++ = '0';
            *o++ = '.'; *o++ = '0';
        } else {
            memcpy(o, digits, decpt); o += decpt;
            *o++ = '.';
            memcpy(o, digits + decpt, nd - decpt); o += nd - decpt;
        }
    } else {
        *o++ = digits[0];
        if (nd > 1) { *o++ = '.'; memcpy(o, digits + 1, nd - 1); o += nd - 1; }
        *o++ = 'e';
        int e = decpt - 1;
        *o++ = e < 0 ? '-' : '+';
        if (e < 0) e = -e;
        if (e < 10) *o++ = '0';
        o = put_int(o, e);
    }
    return (int)(o - out);
}

struct Job {
    const char* const* names;
    const size_t* name_len;
    const int32_t *chr1, *chr2;
    const int64_t *mid1, *mid2, *count;
    const double *p, *q;
};

// rows [lo, hi) -> text -> one gzip member appended to `out`; returns rows kept, or -1
long long format_and_deflate(const Job& J, long long lo, long long hi, bool header, int level, std::string& text, std::string& out) {
    text.clear();
    if (header) text += "chr1\tfragmentMid1\tchr2\tfragmentMid2\tcontactCount\tp-value\tq-value\n";
    long long kept = 0;
    char line[256];
    for (long long i = lo; i < hi; ++i) {
        const double pv = J.p[i];
        if (!(pv <= 1.0)) continue;
        const int a = J.chr1 ? J.chr1[i] : 0, b = J.chr2 ? J.chr2[i] : 0;
        char* o = line;
        memcpy(o, J.names[a], J.name_len[a]); o += J.name_len[a]; *o++ = '\t';
        o = put_int(o, J.mid1[i]); *o++ = '\t';
        memcpy(o, J.names[b], J.name_len[b]); o += J.name_len[b]; *o++ = '\t';
        o = put_int(o, J.mid2[i]); *o++ = '\t';
        o = put_int(o, J.count[i]); *o++ = '\t';
        o += format_double(pv, o); *o++ = '\t';
        if (J.q) o += format_double(J.q[i], o); else { *o++ = '-'; *o++ = '1'; }
        *o++ = '\n';
        text.append(line, (size_t)(o - line));
        kept += 1;
    }
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return -1;
    const size_t bound = deflateBound(&zs, (uLong)text.size()) + 64;
    const size_t at = out.size();
    out.resize(at + bound);
    zs.next_in = (Bytef*)text.data();
    zs.avail_in = (uInt)text.size();
    zs.next_out = (Bytef*)&out[at];
    zs.avail_out = (uInt)bound;
    const int rc = deflate(&zs, Z_FINISH);
    const size_t produced = bound - zs.avail_out;
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) return -1;
    out.resize(at + produced);
    return kept;
}

}  // namespace

extern "C" void bbkio_last_error(char* buf, size_t len) {
    if (!buf || !len) return;
    strncpy(buf, g_err, len - 1);
    buf[len - 1] = 0;
}

extern "C" int bbkio_format_double(double x, char* buf) { return format_double(x, buf); }

extern "C" int bbkio_write_significances(const char* path, const char* const* chrom_names, int32_t n_chrom, const int32_t* chr1,
                                         const int64_t* mid1, const int32_t* chr2, const int64_t* mid2, const int64_t* count,
                                         const double* p, const double* q, int64_t n, int32_t threads, int32_t level,
                                         int64_t* rows_written) {
    if (!path || !chrom_names || n_chrom <= 0 || n < 0 || (n > 0 && (!mid1 || !mid2 || !count || !p)) || ((chr1 == nullptr) != (chr2 == nullptr))) {
        set_error("bbkio_write_significances: bad arguments");
        return BBKIO_E_INVALID;
    }
    if (level < 1 || level > 9) level = 1;
    std::vector<size_t> name_len(n_chrom);
    for (int c = 0; c < n_chrom; ++c) {
        if (!chrom_names[c] || strlen(chrom_names[c]) > 100) { set_error("bbkio_write_significances: bad chromosome name %d", c); return BBKIO_E_INVALID; }
        name_len[c] = strlen(chrom_names[c]);
    }
    if (chr1) {
        for (int64_t i = 0; i < n; ++i)
            if (chr1[i] < 0 || chr1[i] >= n_chrom || chr2[i] < 0 || chr2[i] >= n_chrom) {
                set_error("bbkio_write_significances: chromosome id out of range at row %lld", (long long)i);
                return BBKIO_E_INVALID;
            }
    }
    int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    const long long block = 1 << 17;                                  // rows per gzip member (~6 MB of text)
    const long long n_blocks = n == 0 ? 1 : (n + block - 1) / block;
    if (T > n_blocks) T = (int)n_blocks;
    FILE* fh = fopen(path, "wb");
    if (!fh) { set_error("bbkio_write_significances: cannot open %s", path); return BBKIO_E_IO; }
    Job J = {chrom_names, name_len.data(), chr1, chr2, mid1, mid2, count, p, q};
    std::vector<std::string> text(T), out(T);
    std::vector<long long> kept(T);
    long long total = 0;
    int rc = BBKIO_OK;
    for (long long wave = 0; wave < n_blocks && rc == BBKIO_OK; wave += T) {
        const int live = (int)((n_blocks - wave) < T ? (n_blocks - wave) : T);
        std::vector<std::thread> pool;
        for (int t = 0; t < live; ++t) {
            pool.emplace_back([&, t]() {
                const long long b = wave + t, lo = b * block, hi = (lo + block < n) ? lo + block : n;
                out[t].clear();
                kept[t] = format_and_deflate(J, lo, hi, b == 0, level, text[t], out[t]);
            });
        }
        for (auto& th : pool) th.join();
        for (int t = 0; t < live; ++t) {
            if (kept[t] < 0) { set_error("bbkio_write_significances: zlib failed"); rc = BBKIO_E_ZLIB; break; }
            if (fwrite(out[t].data(), 1, out[t].size(), fh) != out[t].size()) { set_error("bbkio_write_significances: short write to %s", path); rc = BBKIO_E_IO; break; }
            total += kept[t];
        }
    }
    if (fclose(fh) != 0 && rc == BBKIO_OK) { set_error("bbkio_write_significances: close failed on %s", path); rc = BBKIO_E_IO; }
    if (rows_written) *rows_written = total;
    return rc;
}
