// api.cu -- error plumbing shared by every entry point of libcnc_b200.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace cnc {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char *what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
        return CNC_ECUDA;
    }
    return CNC_OK;
}

}  // namespace cnc

extern "C" {
int cnc_version(void) { return 100; /* 0.1.0 */ }
const char *cnc_last_error(void) { return cnc::g_err; }
}
