// C-ABI housekeeping: version, last-error string, device probe.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace lpf {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace lpf

extern "C" int lpf_abi_version(void) { return 1; }
extern "C" const char* lpf_last_error(void) { return lpf::g_err; }
extern "C" int lpf_device_ok(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return major == 10 ? 1 : 0;
}
