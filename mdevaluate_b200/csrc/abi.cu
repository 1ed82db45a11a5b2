// C-ABI plumbing shared by the kernels: error text, version, device query.
#include <stdarg.h>

#include "common.cuh"

namespace mdb {

char *error_buffer() {
    static thread_local char buf[512] = "";
    return buf;
}

int fail(int status, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return status;
}

int &last_launch_count() {
    static thread_local int n = 0;
    return n;
}

}  // namespace mdb

extern "C" const char *mdb_last_error(void) { return mdb::error_buffer(); }

extern "C" int mdb_abi_version(void) { return MDB_ABI_VERSION; }

extern "C" int mdb_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    MDB_CUDA_CHECK(cudaGetDevice(&dev));
    int sms = 0, major = 0, minor = 0;
    MDB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MDB_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    MDB_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = major;
    if (cc_minor) *cc_minor = minor;
    return MDB_OK;
}
