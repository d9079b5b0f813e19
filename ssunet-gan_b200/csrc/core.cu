// Error plumbing, device query.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace ssg {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace ssg

extern "C" {

int ssg_version(void) { return 100; }

const char* ssg_last_error(void) { return ssg::g_err; }

int ssg_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    SSG_CHECK_CUDA(cudaGetDevice(&dev));
    int sms = 0, maj = 0, min = 0;
    SSG_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SSG_CHECK_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    SSG_CHECK_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    return SSG_OK;
}

}  // extern "C"
