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

// Strided host -> device copy of frame rows (one cudaMemcpy2DAsync): the columns of an atom slab out of a
// page-locked (frames x 3N) trajectory go straight into the slab's rows at PCIe speed, with no staging copy
// on the host.  `src_host` should be page-locked for the copy to be asynchronous.
extern "C" int mdb_copy_rows_h2d(const void *src_host, int64_t src_pitch_bytes, void *dst_dev, int64_t dst_pitch_bytes,
                                 int64_t width_bytes, int64_t n_rows, void *stream) {
    if (!src_host || !dst_dev) return mdb::fail(MDB_ERR_ARGUMENT, "mdb_copy_rows_h2d: NULL pointer");
    if (width_bytes < 0 || n_rows < 0 || src_pitch_bytes < width_bytes || dst_pitch_bytes < width_bytes)
        return mdb::fail(MDB_ERR_ARGUMENT, "mdb_copy_rows_h2d: pitches must be >= width >= 0");
    if (width_bytes == 0 || n_rows == 0) return MDB_OK;
    MDB_CUDA_CHECK(cudaMemcpy2DAsync(dst_dev, (size_t)dst_pitch_bytes, src_host, (size_t)src_pitch_bytes,
                                     (size_t)width_bytes, (size_t)n_rows, cudaMemcpyHostToDevice,
                                     static_cast<cudaStream_t>(stream)));
    return MDB_OK;
}
