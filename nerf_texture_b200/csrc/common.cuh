// common.cuh — shared host/device helpers for libntx (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <algorithm>

#include "../../include/ntx.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libntx is written for sm_100a (B200) only"
#endif

namespace ntx {

constexpr int kNumSMs = 148;  // B200

void set_error(const char* fmt, ...);

// run-time tuning knobs (environment variables, read once): experiments without rebuilding
struct Tunables {
    int field_ctas;          // NTX_FIELD_CTAS: resident CTAs per SM of the fused field kernel (0 = as many as fit)
    int pair_ctas;           // NTX_PAIR_CTAS: same for the stand-alone pair gather kernel
    int mesh_block;          // NTX_MESH_BLOCK: threads per block of the mesh kernels (32..128, default 128)
    int frame_ahead;         // NTX_FRAME_AHEAD: iterations ntx_render_rays queues ahead of the mailbox it has read (0 = by frame size)
};
const Tunables& tunables();

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return NTX_ERR_CUDA;
    }
    return NTX_OK;
}

#define NTX_REQUIRE(cond, code, ...)        \
    do {                                    \
        if (!(cond)) {                      \
            ::ntx::set_error(__VA_ARGS__);  \
            return (code);                  \
        }                                   \
    } while (0)

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

// Resident CTAs per SM of a kernel, from registers / shared memory / threads / tensor-memory columns.
// (cudaOccupancyMaxActiveBlocksPerMultiprocessor reports 1 for kernels containing tcgen05.alloc — measured on B200 — although
// several such CTAs do co-reside as long as their TMEM columns sum to <= 512, so persistent grids are sized with this instead.)
inline int resident_ctas_per_sm(const void* func, int threads, size_t dyn_smem, int tmem_cols) {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, func) != cudaSuccess) { cudaGetLastError(); return 1; }
    int dev = 0, regs_sm = 65536, smem_sm = 233472, thr_sm = 2048;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&thr_sm, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
    const int regs = ((fa.numRegs + 7) / 8) * 8;  // allocation granularity
    int n = regs_sm / std::max(1, regs * threads);
    n = std::min(n, (int)(smem_sm / (dyn_smem + fa.sharedSizeBytes + 1024)));  // 1 KB reserved per CTA
    n = std::min(n, thr_sm / threads);
    if (tmem_cols > 0) n = std::min(n, 512 / tmem_cols);
    n = std::min(n, 32);
    return std::max(n, 1);
}
// field.cu: the fused field kernel, callable from the device-driven frame loop in raymarch.cu
int launch_ngp_field(const float* xyz, const float* dirs, const float* deltas, uint32_t M, const int* M_dev, const int* rows, float bound, const void* embeddings_f16,
                     const int* offsets, uint32_t L, float S, uint32_t H, int align_corners, const void* w_sigma_f16, const void* w_color_f16,
                     float density_scale, float* sigmas, float* rgbs, cudaStream_t stream);

// api.cu: TMA descriptor of a row-major fp16 matrix (driver entry point resolved at run time, no link-time libcuda dependency)
int make_tensor_map_2d_f16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows);

// field.cu: the same kernel over the device-driven frame's live list of (sample row, ray) entries
int launch_ngp_field_rays(const int2* live, const int* n_live_dev, uint32_t M_bound, const float* ts, const float* rays_o, const float* rays_d, float bound,
                          const void* embeddings_f16, const int* offsets, uint32_t L, float S, uint32_t H, int align_corners, const void* w_sigma_f16,
                          const void* w_color_f16, float density_scale, float* sigmas, float* rgbs, cudaStream_t stream);
// field.cu: the same kernel in density-grid maintenance mode (positions generated from Morton cell indices, sigma net only)
int launch_density_query(uint32_t n_cells, const int* cells, const float* noise, uint32_t grid_size, float cascade_bound, float bound, const void* embeddings_f16,
                         const int* offsets, uint32_t L, float S, uint32_t H, int align_corners, const void* w_sigma_f16, float density_scale,
                         float* tmp_grid_cascade, cudaStream_t stream);

// per-device caches of launch configuration (one process may drive several GPUs)
constexpr int kMaxDevices = 64;
inline int current_device() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < kMaxDevices) ? d : 0; }

inline int device_sm_count() {
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

// ---- cache-hinted accesses ---------------------------------------------------------------------------
// streaming (read-once) 128-bit load that does not pollute L1
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_u2(void* p, const uint2& v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream_f32(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// table gather: read-only path, keep in L1 (tables are re-read by neighbouring samples)
__device__ __forceinline__ uint32_t ld_table_u32(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

__device__ __forceinline__ float2 half2_bits_to_float2(uint32_t v) {
    __half2 h = *reinterpret_cast<__half2*>(&v);
    return __half22float2(h);
}
__device__ __forceinline__ uint32_t float2_to_half2_bits(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

}  // namespace ntx
