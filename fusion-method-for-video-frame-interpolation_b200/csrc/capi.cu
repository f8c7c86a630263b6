// capi.cu -- error plumbing and misc entry points of the C-ABI (include/fvfi.h).
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace fvfi {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}
}  // namespace fvfi

extern "C" {
int fvfi_version(void) { return 100; }
const char* fvfi_last_error(void) { return fvfi::g_err; }
int fvfi_device_sm_count(void) { return fvfi::sm_count(); }
unsigned long long fvfi_launch_count(void) { return fvfi::g_launches.load(); }
}
