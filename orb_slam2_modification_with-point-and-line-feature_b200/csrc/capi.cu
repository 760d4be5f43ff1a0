// capi.cu — library-wide pieces of the C ABI: error reporting and device discovery.
#include <stdarg.h>

#include "pl_common.cuh"

namespace pl {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace pl

extern "C" {
PL_API const char* pl_last_error(void) { return pl::g_err; }

PL_API int pl_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        pl::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return 0;
    }
    return n;
}

PL_API const char* pl_build_info(void) { return "plslam sm_100a (CUDA " PL_STR(CUDART_VERSION) ")"; }
}
