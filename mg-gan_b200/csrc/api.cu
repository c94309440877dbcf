// C-ABI plumbing: version, device check, per-thread error string.
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>

namespace {
thread_local char g_err[512] = "";
}

int mggan_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int mggan_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return mggan_set_error(MGGAN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return MGGAN_OK;
}

extern "C" const char* mggan_last_error(void) { return g_err; }

extern "C" int mggan_version(void) { return 100; }   // 0.1.0

// Fails unless the current device is compute capability 10.x (the library holds sm_100a code only).
extern "C" int mggan_device_check(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return mggan_set_error(MGGAN_ERR_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return mggan_set_error(MGGAN_ERR_DEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (p.major != 10)
        return mggan_set_error(MGGAN_ERR_DEVICE, "device %d (%s) is sm_%d%d; this library is built for sm_100a only", dev,
                               p.name, p.major, p.minor);
    return MGGAN_OK;
}
