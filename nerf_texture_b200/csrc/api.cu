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
        v.frame_ahead = (e = getenv("NTX_FRAME_AHEAD")) ? atoi(e) : 0;
        v.mesh_block = (e = getenv("NTX_MESH_BLOCK")) ? std::min(128, std::max(32, atoi(e) / 32 * 32)) : 128;
        return v;
    }();
    return t;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point: libntx links only the (static) runtime, libcuda is the
// process's own.  2-D row-major fp16 tensor [rows x inner]; a box is `box_rows` rows of `box_inner` elements.
int make_tensor_map_2d_f16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows) {
    using encode_fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); fn = nullptr; }
        return reinterpret_cast<encode_fn>(fn);
    }();
    if (!encode) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return NTX_ERR_CUDA; }
    const cuuint64_t dims[2] = {inner, rows};
    const cuuint64_t strides[1] = {inner * 2};       // bytes between rows
    const cuuint32_t box[2] = {box_inner, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for a [%llu x %llu] fp16 tensor", (int)r, (unsigned long long)rows, (unsigned long long)inner); return NTX_ERR_CUDA; }
    return NTX_OK;
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
