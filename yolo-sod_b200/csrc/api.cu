// C-ABI housekeeping: version, thread-local last error.
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void ysod_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {
int ysod_version(void) { return 100; }
const char* ysod_last_error(void) { return g_err; }
// Which device architecture the library was compiled for (sm_100a only).
int ysod_compiled_arch(void) { return 100; }
// 16-bit storage / tensor-core input type of this build: 1 = bf16 (libysod.so), 2 = IEEE fp16 (libysod_f16.so, -DYSOD_HALF=1).
int ysod_storage_dtype(void) {
#ifdef YSOD_HALF
    return 2;
#else
    return 1;
#endif
}
}
