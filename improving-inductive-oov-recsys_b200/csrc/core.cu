// Library-wide state: error string, launch counter, device check.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace oov {

std::atomic<uint64_t> g_launches{0};
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cur_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    return dev;
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace oov

extern "C" {

const char* oov_version(void) { return "oov_b200 0.1 (sm_100a)"; }
const char* oov_last_error(void) { return oov::g_err; }
uint64_t oov_launch_count(void) { return oov::g_launches.load(std::memory_order_relaxed); }

int oov_check_device(int device) {
    int major = 0, minor = 0;
    cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
    if (e != cudaSuccess) {
        oov::set_error("oov_check_device: %s", cudaGetErrorString(e));
        return OOV_ERR_CUDA;
    }
    if (major != 10) {
        oov::set_error("oov_check_device: device %d is sm_%d%d; this library is built for sm_100a only", device, major, minor);
        return OOV_ERR_ARCH;
    }
    return OOV_OK;
}

}  // extern "C"
