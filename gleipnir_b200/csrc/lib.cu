// Library plumbing: last-error string, device info.
#include "common.cuh"

extern "C" {

int gl_abi_version(void) { return GL_ABI_VERSION; }

const char *gl_last_error(void) { return gl::error_buffer(); }

int gl_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    GL_CUDA(cudaGetDevice(&dev));
    int n = 0, maj = 0, min = 0;
    GL_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    GL_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    GL_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = n;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    return GL_OK;
}

}  // extern "C"
