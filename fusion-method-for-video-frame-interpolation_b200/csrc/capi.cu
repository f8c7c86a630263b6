// capi.cu -- error plumbing and misc entry points of the C-ABI (include/fvfi.h).
#include <stdarg.h>

#include <atomic>
#include <map>
#include <mutex>
#include <utility>

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

int current_device() {
    int dev = 0;
    return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

int sm_count() {      // queried once per device (this sits on the launch path of every kernel)
    static std::atomic<int> cached[FVFI_MAX_DEVICES];
    const int dev = current_device();
    if (dev < 0 || dev >= FVFI_MAX_DEVICES) return -1;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n > 0) return n;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached[dev].store(n, std::memory_order_relaxed);
    return n;
}
int smem_opt_in(const void* kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> granted;
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = granted[std::make_pair(kernel, dev)];
    if (have >= bytes && have > 0) return FVFI_OK;
    FVFI_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
    return FVFI_OK;
}
}  // namespace fvfi

extern "C" {
int fvfi_version(void) { return 100; }
const char* fvfi_last_error(void) { return fvfi::g_err; }
int fvfi_device_sm_count(void) { return fvfi::sm_count(); }
unsigned long long fvfi_launch_count(void) { return fvfi::g_launches.load(); }
}
