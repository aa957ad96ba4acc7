// api.cu — error reporting and library-level entry points of libntx.
#include "common.cuh"

#include <cstdlib>
#include <cstring>

namespace ntx {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const Tunables& tunables() {
    static Tunables t = [] {
        Tunables v;
        const char* e;
        v.field_ctas = (e = getenv("NTX_FIELD_CTAS")) ? atoi(e) : 0;
        v.pair_ctas = (e = getenv("NTX_PAIR_CTAS")) ? atoi(e) : 0;
        return v;
    }();
    return t;
}
}  // namespace ntx

extern "C" const char* ntx_last_error(void) { return ntx::g_err; }
extern "C" int ntx_version(void) { return 100; }

extern "C" int ntx_device_ok(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}

// kept for drop-in parity with ffmlp.py:126,133 (the reference creates side streams for its split-K GEMMs)
extern "C" int ntx_allocate_splitk(size_t) { return NTX_OK; }
extern "C" int ntx_free_splitk(void) { return NTX_OK; }
