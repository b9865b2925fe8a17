// libdfm: version, error plumbing.
#include "dfm_common.cuh"

namespace dfm {

static thread_local char g_err[512] = "";

char *err_buf() { return g_err; }

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char *what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();   // clear the sticky launch-configuration error
        return fail(DFM_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    }
    return DFM_OK;
}

}  // namespace dfm

extern "C" int dfm_version(void) { return DFM_VERSION; }
extern "C" int dfm_exact_order(void) { return DFM_EXACT_ORDER; }
extern "C" const char *dfm_last_error(void) { return dfm::err_buf(); }

