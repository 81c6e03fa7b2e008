// api.cu - version, error reporting, device query.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void bbk_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int bbk_num_sms() {
    static thread_local int cached_dev = -1, cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        cached_dev = dev;
        cached_sms = sms;
    }
    return cached_sms;
}

extern "C" int bbk_version(void) { return BBK_VERSION; }

extern "C" int bbk_last_error(char* buf, size_t buflen) {
    size_t n = strlen(g_err);
    if (buf && buflen) {
        size_t c = n < buflen - 1 ? n : buflen - 1;
        memcpy(buf, g_err, c);
        buf[c] = 0;
    }
    return (int)n;
}

extern "C" int bbk_sm_count(void) { return bbk_num_sms(); }
