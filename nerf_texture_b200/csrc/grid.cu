// grid.cu — multi-resolution hash / tiled grid encoder for sm_100a.
//
// Replaces gridencoder/src/gridencoder.cu of the reference (kernel_grid :76, kernel_grid_backward :228,
// kernel_input_backward :318).  Same arithmetic, different machine mapping:
//   * one launch covers all L levels and writes [B, L*C] directly (the reference launches a (B/512, L) grid into
//     [L,B,C] and then pays a permute copy in Python, grid.py:52);
//   * hot instantiation (D=3, C=2): a *lane pair* owns one sample — lane 0 gathers the four x0 corners, lane 1 the
//     four x0+1 corners.  x-neighbours are adjacent in memory for dense levels and (prime[0]==1) almost always in the
//     same 32-byte sector for hashed levels, so pairing them in one load instruction halves the L1 wavefronts
//     (distinct 128-byte lines per instruction), which is what bounds this kernel on B200 — the 23 MiB fp16 table
//     is L2-resident, DRAM is not the limiter;
//   * consecutive lane pairs hold consecutive samples of the *same* level, so ray-coherent inputs share lines
//     inside one instruction as well;
//   * 16 independent gathers are in flight per lane (4 levels x 4 corners) before the first use;
//   * the 8-corner accumulation is done in exactly the reference's order and rounding (fp16 tables round after
//     every corner like c10::Half, see oracle/grid_impl.inc), so fp16 results are bit-identical to the reference.
#include "grid_common.cuh"

#include <type_traits>

namespace ntx {

// ---------------------------------------------------------------------------------------------------- scalar traits
// rounding points: float/double = FMA (nvcc -fmad contraction in the reference), half = c10::Half operator semantics
template <typename T> struct Num;
template <> struct Num<float> {
    using vec2 = float2;
    static __device__ __forceinline__ float zero() { return 0.f; }
    static __device__ __forceinline__ void acc(float& a, float w, float v) { a = __fmaf_rn(w, v, a); }
    static __device__ __forceinline__ float sub(float r, float l) { return r - l; }
    static __device__ __forceinline__ float wmul(float w, float v) { return w * v; }
    static __device__ __forceinline__ void mulacc(float& a, float x, float y) { a = __fmaf_rn(x, y, a); }
};
template <> struct Num<double> {
    static __device__ __forceinline__ double zero() { return 0.0; }
    static __device__ __forceinline__ void acc(double& a, float w, double v) { a = fma((double)w, v, a); }
    static __device__ __forceinline__ double sub(double r, double l) { return r - l; }
    static __device__ __forceinline__ double wmul(float w, double v) { return (double)w * v; }
    static __device__ __forceinline__ void mulacc(double& a, double x, double y) { a = fma(x, y, a); }
};
template <> struct Num<__half> {
    static __device__ __forceinline__ __half zero() { return __float2half_rn(0.f); }
    static __device__ __forceinline__ void acc(__half& a, float w, __half v) {
        const __half p = __float2half_rn(w * __half2float(v));
        a = __float2half_rn(__half2float(a) + __half2float(p));
    }
    static __device__ __forceinline__ __half sub(__half r, __half l) { return __float2half_rn(__half2float(r) - __half2float(l)); }
    static __device__ __forceinline__ __half wmul(float w, __half v) { return __float2half_rn(w * __half2float(v)); }
    static __device__ __forceinline__ void mulacc(__half& a, __half x, __half y) {
        const __half p = __float2half_rn(__half2float(x) * __half2float(y));
        a = __float2half_rn(__half2float(a) + __half2float(p));
    }
};

// ---------------------------------------------------------------------------------------------------- generic forward
// one thread per (sample, level); any D in {2,3}, C in {1,2,4,8}, any table dtype, optional dy_dx.
template <typename scalar_t, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) grid_fwd_generic_kernel(
    const float* __restrict__ inputs, const scalar_t* __restrict__ grid, const int* __restrict__ offsets,
    scalar_t* __restrict__ outputs, const uint32_t B, const uint32_t L, const float S, const uint32_t H,
    const bool calc_grad_inputs, scalar_t* __restrict__ dy_dx, const uint32_t gridtype, const bool align, const int layout) {
    const uint32_t level = blockIdx.y;
    __shared__ GridLevel sg;
    if (threadIdx.x == 0) sg = make_level<D>(offsets, level, S, H, gridtype, align);
    __syncthreads();
    const GridLevel g = sg;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        scalar_t* out = (layout == NTX_LAYOUT_LBC) ? outputs + ((size_t)level * B + b) * C : outputs + ((size_t)b * L + level) * C;
        scalar_t* dd = calc_grad_inputs ? dy_dx + (size_t)b * D * L * C + (size_t)level * D * C : nullptr;
        float x[D];
        bool oob = false;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) { x[d] = inputs[(size_t)b * D + d]; if (x[d] < 0 || x[d] > 1) oob = true; }
        if (oob) {
#pragma unroll
            for (uint32_t c = 0; c < C; c++) out[c] = Num<scalar_t>::zero();
            if (dd) for (uint32_t i = 0; i < D * C; i++) dd[i] = Num<scalar_t>::zero();
            continue;
        }
        float pos[D]; uint32_t pg[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            pos[d] = __fmaf_rn(x[d], g.scale, align ? 0.0f : 0.5f);
            const float fl = floorf(pos[d]);
            pg[d] = (uint32_t)fl;
            pos[d] -= fl;
        }
        const scalar_t* gl = grid + (size_t)g.offset * C;
        scalar_t res[C];
#pragma unroll
        for (uint32_t c = 0; c < C; c++) res[c] = Num<scalar_t>::zero();
#pragma unroll
        for (uint32_t idx = 0; idx < (1u << D); idx++) {
            float w = 1; uint32_t pl[D];
#pragma unroll
            for (uint32_t d = 0; d < D; d++) {
                if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; }
                else { w *= pos[d]; pl[d] = pg[d] + 1; }
            }
            const uint32_t index = corner_index<D>(g, pl) * C;
#pragma unroll
            for (uint32_t c = 0; c < C; c++) Num<scalar_t>::acc(res[c], w, gl[index + c]);
        }
#pragma unroll
        for (uint32_t c = 0; c < C; c++) out[c] = res[c];

        if (dd) {  // gridencoder.cu:180-222
#pragma unroll
            for (uint32_t gd = 0; gd < D; gd++) {
                scalar_t rg[C];
#pragma unroll
                for (uint32_t c = 0; c < C; c++) rg[c] = Num<scalar_t>::zero();
#pragma unroll
                for (uint32_t idx = 0; idx < (1u << (D - 1)); idx++) {
                    float w = g.scale; uint32_t pl[D];
#pragma unroll
                    for (uint32_t nd = 0; nd < D - 1; nd++) {
                        const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
                        if ((idx & (1u << nd)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; }
                        else { w *= pos[d]; pl[d] = pg[d] + 1; }
                    }
                    pl[gd] = pg[gd];
                    const uint32_t il = corner_index<D>(g, pl) * C;
                    pl[gd] = pg[gd] + 1;
                    const uint32_t ir = corner_index<D>(g, pl) * C;
#pragma unroll
                    for (uint32_t c = 0; c < C; c++) Num<scalar_t>::acc(rg[c], w, Num<scalar_t>::sub(gl[ir + c], gl[il + c]));
                }
#pragma unroll
                for (uint32_t c = 0; c < C; c++) dd[gd * C + c] = rg[c];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------- pair forward
constexpr int kPairThreads = 256;
constexpr int kPairMaxLevels = 32;

template <typename scalar_t>
__global__ void __launch_bounds__(kPairThreads, 3) grid_fwd_pair_kernel(
    const float* __restrict__ inputs, const scalar_t* __restrict__ grid, const int* __restrict__ offsets,
    scalar_t* __restrict__ outputs, const uint32_t B, const uint32_t L, const float S, const uint32_t H,
    const uint32_t gridtype, const bool align, const int layout) {
    using E = Elem2<scalar_t>;
    using raw = typename E::raw;
    __shared__ PairLevel lv[kPairMaxLevels];
    if (threadIdx.x < L) lv[threadIdx.x] = make_pair_level(make_level<3>(offsets, threadIdx.x, S, H, gridtype, align), grid);
    __syncthreads();

    const uint32_t p = threadIdx.x & 1u;  // which x corner this lane gathers
    const float half_off = align ? 0.0f : 0.5f;
    const uint32_t groups = ceil_div<uint32_t>(B, kPairThreads / 2);
    for (uint32_t grp = blockIdx.x; grp < groups; grp += gridDim.x) {
        const uint32_t b = grp * (kPairThreads / 2) + (threadIdx.x >> 1);
        const bool valid = b < B;
        float x = 0.f, y = 0.f, z = 0.f;
        if (valid) { x = inputs[(size_t)b * 3]; y = inputs[(size_t)b * 3 + 1]; z = inputs[(size_t)b * 3 + 2]; }
        const uint32_t cls = valid ? sample_class(x, y, z) : 0u;
        const bool live = cls == 1u;
        if (!live) { x = 0.f; y = 0.f; z = 0.f; }  // keeps the (discarded) loads of dead lanes in bounds

        for (uint32_t l0 = 0; l0 < L; l0 += 4) {
            raw packed[2];
            pair_gather4<scalar_t>(x, y, z, live, p, lv + l0, half_off, packed);
            if (cls == 2u) { packed[0] = nan_features<raw>(); packed[1] = nan_features<raw>(); }
            if (valid) {
                const uint32_t la = l0 + 2 * p;
                if (layout == NTX_LAYOUT_BLC) {
                    raw* dst = reinterpret_cast<raw*>(outputs + ((size_t)b * L + la) * 2);
                    if (sizeof(raw) == 4) { st_stream_u2(dst, make_uint2((uint32_t)packed[0], (uint32_t)packed[1])); }
                    else { st_stream_u4(dst, make_uint4((uint32_t)packed[0], (uint32_t)((uint64_t)packed[0] >> 32), (uint32_t)packed[1], (uint32_t)((uint64_t)packed[1] >> 32))); }
                } else {
                    *reinterpret_cast<raw*>(outputs + ((size_t)la * B + b) * 2) = packed[0];
                    *reinterpret_cast<raw*>(outputs + ((size_t)(la + 1) * B + b) * 2) = packed[1];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------- backward
template <typename T> struct AtomicAdd2;  // add two adjacent channels
// add_pair: two adjacent 2-channel entries (an aligned 4-scalar slot) in one vector reduction (sm_90+ `red.v2.f16x2` / `red.v4.f32`)
template <> struct AtomicAdd2<float> {
    static constexpr bool kHasPairAdd = true;
    static __device__ __forceinline__ void add(float* p, float a, float b) { atomicAdd(reinterpret_cast<float2*>(p), make_float2(a, b)); }
    static __device__ __forceinline__ void add1(float* p, float a) { atomicAdd(p, a); }
    static __device__ __forceinline__ void add_pair(float* p, float a0, float b0, float a1, float b1) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a0), "f"(b0), "f"(a1), "f"(b1) : "memory");
    }
};
template <> struct AtomicAdd2<double> {
    static constexpr bool kHasPairAdd = false;
    static __device__ __forceinline__ void add(double* p, double a, double b) { atomicAdd(p, a); atomicAdd(p + 1, b); }
    static __device__ __forceinline__ void add1(double* p, double a) { atomicAdd(p, a); }
    static __device__ __forceinline__ void add_pair(double* p, double a0, double b0, double a1, double b1) { add(p, a0, b0); add(p + 2, a1, b1); }
};
template <> struct AtomicAdd2<__half> {
    static constexpr bool kHasPairAdd = true;
    static __device__ __forceinline__ void add(__half* p, __half a, __half b) { atomicAdd(reinterpret_cast<__half2*>(p), __halves2half2(a, b)); }
    static __device__ __forceinline__ void add1(__half* p, __half a) { atomicAdd(p, a); }
    static __device__ __forceinline__ void add_pair(__half* p, __half a0, __half b0, __half a1, __half b1) {
        const __half2 lo = __halves2half2(a0, b0), hi = __halves2half2(a1, b1);
        asm volatile("red.global.add.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(p), "r"(*reinterpret_cast<const uint32_t*>(&lo)), "r"(*reinterpret_cast<const uint32_t*>(&hi)) : "memory");
    }
};

// one thread per (sample, level, channel pair); level is the fastest index inside a warp's sample so that a warp's
// grad read ([B, L*C] layout) is one contiguous segment.
template <typename scalar_t, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) grid_bwd_kernel(
    const scalar_t* __restrict__ grad, const float* __restrict__ inputs, const int* __restrict__ offsets,
    scalar_t* __restrict__ grad_grid, const uint32_t B, const uint32_t L, const float S, const uint32_t H,
    const uint32_t gridtype, const bool align, const int layout, const uint32_t level0) {
    constexpr uint32_t NC = C >= 2 ? 2 : 1;  // channels per thread (gridencoder.cu:378)
    constexpr uint32_t CP = C / NC;
    extern __shared__ GridLevel lvb[];
    for (uint32_t l = threadIdx.x; l < L; l += blockDim.x) lvb[l] = make_level<D>(offsets, l, S, H, gridtype, align);
    __syncthreads();
    const uint32_t LN = L - level0;          // levels [level0, L): the coarse ones may have been taken by grid_bwd_small_levels_kernel
    const uint64_t total = (uint64_t)B * LN * CP;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t cp = (uint32_t)(t % CP);
        const uint32_t level = level0 + (uint32_t)((t / CP) % LN);
        const uint32_t b = (uint32_t)(t / ((uint64_t)CP * LN));
        const uint32_t ch = cp * NC;
        const GridLevel g = lvb[level];
        float x[D];
        bool oob = false;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) { x[d] = inputs[(size_t)b * D + d]; if (x[d] < 0 || x[d] > 1) oob = true; }
        if (oob) continue;
        float pos[D]; uint32_t pg[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            pos[d] = __fmaf_rn(x[d], g.scale, align ? 0.0f : 0.5f);
            const float fl = floorf(pos[d]);
            pg[d] = (uint32_t)fl;
            pos[d] -= fl;
        }
        const scalar_t* gp = (layout == NTX_LAYOUT_LBC) ? grad + ((size_t)level * B + b) * C + ch : grad + ((size_t)b * L + level) * C + ch;
        scalar_t gc[NC];
#pragma unroll
        for (uint32_t c = 0; c < NC; c++) gc[c] = gp[c];
        scalar_t* gg = grad_grid + (size_t)g.offset * C + ch;
        // corners are visited in x-pairs (idx, idx + 1): the two entries of a pair differ in the x coordinate only
#pragma unroll
        for (uint32_t idx = 0; idx < (1u << D); idx += 2) {
            float w0 = 1 - pos[0], w1 = pos[0]; uint32_t pl[D];
            pl[0] = pg[0];
#pragma unroll
            for (uint32_t d = 1; d < D; d++) {
                if ((idx & (1u << d)) == 0) { w0 *= 1 - pos[d]; w1 *= 1 - pos[d]; pl[d] = pg[d]; }
                else { w0 *= pos[d]; w1 *= pos[d]; pl[d] = pg[d] + 1; }
            }
            const uint32_t i0 = corner_index<D>(g, pl);
            pl[0] = pg[0] + 1;
            const uint32_t i1 = corner_index<D>(g, pl);
            if (NC == 2) {
                const scalar_t a0 = Num<scalar_t>::wmul(w0, gc[0]), b0 = Num<scalar_t>::wmul(w0, gc[NC - 1]);
                const scalar_t a1 = Num<scalar_t>::wmul(w1, gc[0]), b1 = Num<scalar_t>::wmul(w1, gc[NC - 1]);
                // The two table entries of an x-pair share an aligned 2-entry slot whenever the cell's x is even (dense levels: they are
                // neighbours; hashed levels: prime[0] == 1, so x and x + 1 differ in index bit 0 only): ONE vector reduction then does the
                // work of two — the scatter is bound by the L2's reduction rate (~128 G red/s measured), not by bytes.
                if (C == 2 && (i0 ^ i1) == 1u && AtomicAdd2<scalar_t>::kHasPairAdd) {
                    if (i0 < i1) AtomicAdd2<scalar_t>::add_pair(gg + (size_t)i0 * C, a0, b0, a1, b1);
                    else AtomicAdd2<scalar_t>::add_pair(gg + (size_t)i1 * C, a1, b1, a0, b0);
                } else {
                    AtomicAdd2<scalar_t>::add(gg + (size_t)i0 * C, a0, b0);
                    AtomicAdd2<scalar_t>::add(gg + (size_t)i1 * C, a1, b1);
                }
            } else {
                AtomicAdd2<scalar_t>::add1(gg + (size_t)i0 * C, Num<scalar_t>::wmul(w0, gc[0]));
                AtomicAdd2<scalar_t>::add1(gg + (size_t)i1 * C, Num<scalar_t>::wmul(w1, gc[0]));
            }
        }
    }
}

// Coarse levels: a level-0 table of 4 096 entries receives 8*B updates — with one global atomic per update (the reference's
// scheme, gridencoder.cu:296-311) the L2 serialises them per address.  Levels whose whole gradient table fits in shared memory are
// accumulated there instead, in fp32: a CTA owns one such level and a contiguous share of the batch, adds its updates with
// shared-memory atomics and flushes the non-zero entries with ONE global atomic each at the end (fp32 sums rounded to the table
// dtype once per CTA: strictly more accurate than a chain of fp16 atomics).  D = 3, C = 2.
constexpr uint32_t kBwdSmallThreads = 512;
constexpr uint32_t kBwdSmallMaxBytes = 100u * 1024u;      // fp32 accumulators: levels with hs * 8 bytes <= this are privatised
template <typename scalar_t>
__global__ void __launch_bounds__(kBwdSmallThreads) grid_bwd_small_levels_kernel(
    const scalar_t* __restrict__ grad, const float* __restrict__ inputs, const int* __restrict__ offsets, scalar_t* __restrict__ grad_grid,
    const uint32_t B, const uint32_t L, const float S, const uint32_t H, const uint32_t gridtype, const bool align, const int layout) {
    extern __shared__ __align__(16) float2 acc[];           // [hs] (dC0, dC1)
    __shared__ GridLevel sg;
    const uint32_t level = blockIdx.y;
    if (threadIdx.x == 0) sg = make_level<3>(offsets, level, S, H, gridtype, align);
    __syncthreads();
    const GridLevel g = sg;
    // the launcher sizes levels from (S, H, align); the kernel trusts only the real offsets: a level that does not fit after all
    // is accumulated with global atomics like the fine levels (same result, just slower)
    const bool priv = (size_t)g.hs * sizeof(float2) <= kBwdSmallMaxBytes;
    scalar_t* gg = grad_grid + (size_t)g.offset * 2;
    if (priv) for (uint32_t e = threadIdx.x; e < g.hs; e += blockDim.x) acc[e] = make_float2(0.f, 0.f);
    __syncthreads();
    const uint32_t per = ceil_div<uint32_t>(B, gridDim.x), b0 = blockIdx.x * per, b1 = min(B, b0 + per);
    // a thread owns one (sample, corner): the 8 corners of a sample sit in 8 consecutive lanes and share its loads through L1
    for (uint32_t t = b0 * 8u + threadIdx.x; t < b1 * 8u; t += blockDim.x) {
        const uint32_t b = t >> 3, idx = t & 7u;
        const float x[3] = {inputs[(size_t)b * 3], inputs[(size_t)b * 3 + 1], inputs[(size_t)b * 3 + 2]};
        if ((x[0] < 0 || x[0] > 1) || (x[1] < 0 || x[1] > 1) || (x[2] < 0 || x[2] > 1)) continue;     // gridencoder.cu:262-270
        float w = 1; uint32_t pl[3];
#pragma unroll
        for (uint32_t d = 0; d < 3; d++) {
            const float pos = __fmaf_rn(x[d], g.scale, align ? 0.0f : 0.5f);
            const float fl = floorf(pos);
            const float fr = pos - fl;
            if ((idx & (1u << d)) == 0) { w *= 1 - fr; pl[d] = (uint32_t)fl; }
            else { w *= fr; pl[d] = (uint32_t)fl + 1; }
        }
        const scalar_t* gp = (layout == NTX_LAYOUT_LBC) ? grad + ((size_t)level * B + b) * 2 : grad + ((size_t)b * L + level) * 2;
        const uint32_t index = corner_index<3>(g, pl);
        // the reference multiplies in the table dtype (w * grad rounded to scalar_t, gridencoder.cu:306); keep that rounding point
        const float g0 = (float)Num<scalar_t>::wmul(w, gp[0]), g1 = (float)Num<scalar_t>::wmul(w, gp[1]);
        if (priv) { atomicAdd(&acc[index].x, g0); atomicAdd(&acc[index].y, g1); }
        else AtomicAdd2<scalar_t>::add(gg + (size_t)index * 2, (scalar_t)g0, (scalar_t)g1);
    }
    __syncthreads();
    if (priv) for (uint32_t e = threadIdx.x; e < g.hs; e += blockDim.x) {
        const float2 v = acc[e];
        if (v.x != 0.f || v.y != 0.f) AtomicAdd2<scalar_t>::add(gg + (size_t)e * 2, (scalar_t)v.x, (scalar_t)v.y);
    }
}

// kernel_input_backward (gridencoder.cu:318-343)
template <typename scalar_t>
__global__ void __launch_bounds__(256) grid_input_bwd_kernel(const scalar_t* __restrict__ grad, const scalar_t* __restrict__ dy_dx,
                                                             scalar_t* __restrict__ grad_inputs, uint32_t B, uint32_t D, uint32_t C,
                                                             uint32_t L, const int layout) {
    const uint64_t total = (uint64_t)B * D;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t)(t / D), d = (uint32_t)(t - (uint64_t)b * D);
        const scalar_t* dd = dy_dx + (size_t)b * L * D * C;
        scalar_t r = Num<scalar_t>::zero();
        for (uint32_t l = 0; l < L; l++)
            for (uint32_t c = 0; c < C; c++) {
                const scalar_t gv = (layout == NTX_LAYOUT_LBC) ? grad[((size_t)l * B + b) * C + c] : grad[((size_t)b * L + l) * C + c];
                Num<scalar_t>::mulacc(r, gv, dd[l * D * C + d * C + c]);
            }
        grad_inputs[t] = r;
    }
}

__global__ void grid_level_scales_kernel(float S, uint32_t H, uint32_t L, float* out) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < L) out[l] = level_scale(l, S, H);
}

template <uint32_t D>
__global__ void grid_debug_indices_kernel(const float* __restrict__ inputs, const int* __restrict__ offsets, uint32_t B, uint32_t level,
                                          float S, uint32_t H, uint32_t gridtype, bool align, uint32_t* __restrict__ out) {
    const GridLevel g = make_level<D>(offsets, level, S, H, gridtype, align);
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        float x[D]; bool oob = false; uint32_t pg[D];
        for (uint32_t d = 0; d < D; d++) { x[d] = inputs[(size_t)b * D + d]; if (x[d] < 0 || x[d] > 1) oob = true; }
        for (uint32_t d = 0; d < D; d++) pg[d] = oob ? 0u : (uint32_t)floorf(__fmaf_rn(x[d], g.scale, align ? 0.0f : 0.5f));
        for (uint32_t c = 0; c < (1u << D); c++) {
            uint32_t pl[D];
            for (uint32_t d = 0; d < D; d++) pl[d] = pg[d] + ((c >> d) & 1u);
            out[(size_t)b * (1u << D) + c] = oob ? 0xffffffffu : corner_index<D>(g, pl);
        }
    }
}

// ---------------------------------------------------------------------------------------------------- host side
static int persistent_grid(const void* kernel, int threads, size_t smem) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ < 1) occ = 1;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return occ * sms;
}

template <typename scalar_t, uint32_t D>
static int launch_fwd_generic(const float* inputs, const scalar_t* emb, const int* offsets, scalar_t* out, uint32_t B, uint32_t C,
                              uint32_t L, float S, uint32_t H, bool cgi, scalar_t* dy_dx, uint32_t gridtype, bool align, int layout,
                              cudaStream_t st) {
    const dim3 grid(min(ceil_div<uint32_t>(B, 256), 148u * 8u), L);
#define NTX_FWD(CC) grid_fwd_generic_kernel<scalar_t, D, CC><<<grid, 256, 0, st>>>(inputs, emb, offsets, out, B, L, S, H, cgi, dy_dx, gridtype, align, layout)
    switch (C) {
        case 1: NTX_FWD(1); break;
        case 2: NTX_FWD(2); break;
        case 4: NTX_FWD(4); break;
        case 8: NTX_FWD(8); break;
        default: set_error("GridEncoding: C must be 1, 2, 4, or 8."); return NTX_ERR_UNSUPPORTED;
    }
#undef NTX_FWD
    return check_launch("grid_encode_forward");
}

template <typename scalar_t>
static int launch_fwd_pair(const float* inputs, const scalar_t* emb, const int* offsets, scalar_t* out, uint32_t B, uint32_t L, float S,
                           uint32_t H, uint32_t gridtype, bool align, int layout, cudaStream_t st) {
    static int grid_cap_dev[kMaxDevices] = {};
    int& grid_cap = grid_cap_dev[current_device()];
    if (!grid_cap) {
        grid_cap = persistent_grid((const void*)grid_fwd_pair_kernel<scalar_t>, kPairThreads, 0);
        if (tunables().pair_ctas > 0) grid_cap = tunables().pair_ctas * device_sm_count();
        if (const char* e = getenv("NTX_PAIR_CARVEOUT"))   // experiment: shrink L1 by forcing a shared-memory carve-out (percent)
            cudaFuncSetAttribute((const void*)grid_fwd_pair_kernel<scalar_t>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
    }
    const uint32_t groups = ceil_div<uint32_t>(B, kPairThreads / 2);
    grid_fwd_pair_kernel<scalar_t><<<min(groups, (uint32_t)grid_cap), kPairThreads, 0, st>>>(inputs, emb, offsets, out, B, L, S, H, gridtype, align, layout);
    return check_launch("grid_encode_forward(pair)");
}

template <typename scalar_t>
static int fwd_dispatch(const float* inputs, const void* emb_, const int* offsets, void* out_, uint32_t B, uint32_t D, uint32_t C,
                        uint32_t L, float S, uint32_t H, bool cgi, void* dy_dx_, uint32_t gridtype, bool align, int layout, cudaStream_t st) {
    auto emb = static_cast<const scalar_t*>(emb_);
    auto out = static_cast<scalar_t*>(out_);
    auto dy_dx = static_cast<scalar_t*>(dy_dx_);
    if (D == 2) return launch_fwd_generic<scalar_t, 2>(inputs, emb, offsets, out, B, C, L, S, H, cgi, dy_dx, gridtype, align, layout, st);
    return launch_fwd_generic<scalar_t, 3>(inputs, emb, offsets, out, B, C, L, S, H, cgi, dy_dx, gridtype, align, layout, st);
}

// number of leading levels (all of them must qualify) whose gradient table is accumulated in shared memory
static uint32_t count_small_levels(const int* offsets_host, uint32_t L) {
    uint32_t n = 0;
    while (n < L && (size_t)(offsets_host[n + 1] - offsets_host[n]) * sizeof(float2) <= kBwdSmallMaxBytes) n++;
    return n;
}

template <typename scalar_t, uint32_t D>
static int launch_bwd(const scalar_t* grad, const float* inputs, const int* offsets, scalar_t* gg, uint32_t B, uint32_t C, uint32_t L, float S,
                      uint32_t H, uint32_t gridtype, bool align, int layout, cudaStream_t st) {
    uint32_t level0 = 0;
    if constexpr (D == 3 && !std::is_same<scalar_t, double>::value) {
        if (C == 2 && B >= 4096) {
            // Level sizes follow from (S, H, align, log2_hashmap_size) only, but the offsets live in device memory: recompute the
            // first levels' sizes on the host the way GridEncoder builds them (grid.py:113-124) — dense coarse levels are
            // min(cap, R^3) rounded up to 8 with R = ceil(2^(l*S) * H - 1) + 1 (+1 without align_corners); a level is privatised only if
            // its dense size fits, and the kernel itself uses the real offsets, so a mismatch can only cost performance.
            int sizes[9] = {0};
            uint32_t n = 0;
            for (; n < 8 && n < L; n++) {
                const float scale = exp2f(n * S) * H - 1.0f;
                const uint64_t res = (uint64_t)ceilf(scale) + 1, R = align ? res : res + 1;
                const uint64_t dense = ((R * R * R + 7) / 8) * 8;
                if (dense * sizeof(float2) > kBwdSmallMaxBytes) break;
                sizes[n + 1] = sizes[n] + (int)dense;
            }
            level0 = count_small_levels(sizes, n);
            if (level0) {
                static bool configured_dev[kMaxDevices] = {};
                bool& configured = configured_dev[current_device()];
                if (!configured) {
                    cudaFuncSetAttribute(grid_bwd_small_levels_kernel<scalar_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmallMaxBytes);
                    configured = true;
                }
                const uint32_t ctas = std::max<uint32_t>(1u, (uint32_t)device_sm_count() * 2u / level0);
                const dim3 grid(std::min<uint32_t>(ctas, ceil_div<uint32_t>(B, 256)), level0);
                grid_bwd_small_levels_kernel<scalar_t><<<grid, kBwdSmallThreads, kBwdSmallMaxBytes, st>>>(grad, inputs, offsets, gg, B, L, S, H, gridtype, align, layout);
            }
        }
    }
    if (level0 >= L) return check_launch("grid_encode_backward");
    const uint64_t total = (uint64_t)B * (L - level0) * (C >= 2 ? C / 2 : 1);
    const uint32_t blocks = (uint32_t)std::min<uint64_t>(ceil_div<uint64_t>(total, 256), (uint64_t)device_sm_count() * 16ull);
    const size_t smem = sizeof(GridLevel) * L;
#define NTX_BWD(CC) grid_bwd_kernel<scalar_t, D, CC><<<blocks, 256, smem, st>>>(grad, inputs, offsets, gg, B, L, S, H, gridtype, align, layout, level0)
    switch (C) {
        case 1: NTX_BWD(1); break;
        case 2: NTX_BWD(2); break;
        case 4: NTX_BWD(4); break;
        case 8: NTX_BWD(8); break;
        default: set_error("GridEncoding: C must be 1, 2, 4, or 8."); return NTX_ERR_UNSUPPORTED;
    }
#undef NTX_BWD
    return check_launch("grid_encode_backward");
}

template <typename scalar_t>
static int bwd_dispatch(const void* grad_, const float* inputs, const int* offsets, void* gg_, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                        float S, uint32_t H, bool cgi, const void* dy_dx_, void* gi_, uint32_t gridtype, bool align, int layout, cudaStream_t st) {
    auto grad = static_cast<const scalar_t*>(grad_);
    auto gg = static_cast<scalar_t*>(gg_);
    int rc = (D == 2) ? launch_bwd<scalar_t, 2>(grad, inputs, offsets, gg, B, C, L, S, H, gridtype, align, layout, st)
                      : launch_bwd<scalar_t, 3>(grad, inputs, offsets, gg, B, C, L, S, H, gridtype, align, layout, st);
    if (rc != NTX_OK) return rc;
    if (cgi) {
        const uint32_t blocks = (uint32_t)std::min<uint64_t>(ceil_div<uint64_t>((uint64_t)B * D, 256), 148ull * 16ull);
        grid_input_bwd_kernel<scalar_t><<<blocks, 256, 0, st>>>(grad, static_cast<const scalar_t*>(dy_dx_), static_cast<scalar_t*>(gi_), B, D, C, L, layout);
        rc = check_launch("grid_encode_backward(inputs)");
    }
    return rc;
}

}  // namespace ntx

using namespace ntx;

extern "C" int ntx_grid_encode_forward(const float* inputs, const void* embeddings, const int* offsets, void* outputs, uint32_t B, uint32_t D,
                                       uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs, void* dy_dx, uint32_t gridtype,
                                       int align_corners, int dtype, int out_layout, ntx_stream_t stream) {
    NTX_REQUIRE(D == 2 || D == 3, NTX_ERR_UNSUPPORTED, "GridEncoding: D must be 2 or 3.");
    NTX_REQUIRE(C == 1 || C == 2 || C == 4 || C == 8, NTX_ERR_UNSUPPORTED, "GridEncoding: C must be 1, 2, 4, or 8.");
    NTX_REQUIRE(inputs && embeddings && offsets && outputs, NTX_ERR_INVALID_ARGUMENT, "grid_encode_forward: null pointer");
    NTX_REQUIRE(!calc_grad_inputs || dy_dx, NTX_ERR_INVALID_ARGUMENT, "grid_encode_forward: dy_dx required when calc_grad_inputs");
    NTX_REQUIRE(L >= 1 && L <= 65535, NTX_ERR_INVALID_ARGUMENT, "grid_encode_forward: bad level count %u", L);
    NTX_REQUIRE(out_layout == NTX_LAYOUT_LBC || out_layout == NTX_LAYOUT_BLC, NTX_ERR_INVALID_ARGUMENT, "grid_encode_forward: bad layout");
    if (B == 0) return NTX_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool align = align_corners != 0, cgi = calc_grad_inputs != 0;
    const bool pair_ok = (D == 3 && C == 2 && !cgi && (L % 4 == 0) && L <= kPairMaxLevels);
    switch (dtype) {
        case NTX_F16:
            if (pair_ok) return launch_fwd_pair<__half>(inputs, (const __half*)embeddings, offsets, (__half*)outputs, B, L, S, H, gridtype, align, out_layout, st);
            return fwd_dispatch<__half>(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, cgi, dy_dx, gridtype, align, out_layout, st);
        case NTX_F32:
            if (pair_ok) return launch_fwd_pair<float>(inputs, (const float*)embeddings, offsets, (float*)outputs, B, L, S, H, gridtype, align, out_layout, st);
            return fwd_dispatch<float>(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, cgi, dy_dx, gridtype, align, out_layout, st);
        case NTX_F64:
            return fwd_dispatch<double>(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, cgi, dy_dx, gridtype, align, out_layout, st);
        default:
            set_error("grid_encode_forward: embeddings must be a floating tensor (f16/f32/f64)");
            return NTX_ERR_INVALID_ARGUMENT;
    }
}

extern "C" int ntx_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings, const int* offsets, void* grad_embeddings,
                                        uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs,
                                        const void* dy_dx, void* grad_inputs, uint32_t gridtype, int align_corners, int dtype, int grad_layout,
                                        ntx_stream_t stream) {
    (void)embeddings;
    NTX_REQUIRE(D == 2 || D == 3, NTX_ERR_UNSUPPORTED, "GridEncoding: D must be 2 or 3.");
    NTX_REQUIRE(C == 1 || C == 2 || C == 4 || C == 8, NTX_ERR_UNSUPPORTED, "GridEncoding: C must be 1, 2, 4, or 8.");
    NTX_REQUIRE(grad && inputs && offsets && grad_embeddings, NTX_ERR_INVALID_ARGUMENT, "grid_encode_backward: null pointer");
    NTX_REQUIRE(!calc_grad_inputs || (dy_dx && grad_inputs), NTX_ERR_INVALID_ARGUMENT, "grid_encode_backward: dy_dx/grad_inputs required");
    NTX_REQUIRE(L >= 1 && L <= 1024, NTX_ERR_INVALID_ARGUMENT, "grid_encode_backward: bad level count %u", L);
    if (B == 0) return NTX_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool align = align_corners != 0, cgi = calc_grad_inputs != 0;
    switch (dtype) {
        case NTX_F16: return bwd_dispatch<__half>(grad, inputs, offsets, grad_embeddings, B, D, C, L, S, H, cgi, dy_dx, grad_inputs, gridtype, align, grad_layout, st);
        case NTX_F32: return bwd_dispatch<float>(grad, inputs, offsets, grad_embeddings, B, D, C, L, S, H, cgi, dy_dx, grad_inputs, gridtype, align, grad_layout, st);
        case NTX_F64: return bwd_dispatch<double>(grad, inputs, offsets, grad_embeddings, B, D, C, L, S, H, cgi, dy_dx, grad_inputs, gridtype, align, grad_layout, st);
        default: set_error("grid_encode_backward: grad must be a floating tensor (f16/f32/f64)"); return NTX_ERR_INVALID_ARGUMENT;
    }
}

extern "C" int ntx_grid_level_scales(float S, uint32_t H, uint32_t L, float* scales_out, ntx_stream_t stream) {
    NTX_REQUIRE(scales_out && L >= 1, NTX_ERR_INVALID_ARGUMENT, "grid_level_scales: bad arguments");
    grid_level_scales_kernel<<<ceil_div<uint32_t>(L, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(S, H, L, scales_out);
    return check_launch("grid_level_scales");
}

extern "C" int ntx_grid_debug_indices(const float* inputs, const int* offsets, uint32_t B, uint32_t D, uint32_t level, float S, uint32_t H,
                                      uint32_t gridtype, int align_corners, uint32_t* idx_out, ntx_stream_t stream) {
    NTX_REQUIRE(D == 2 || D == 3, NTX_ERR_UNSUPPORTED, "GridEncoding: D must be 2 or 3.");
    if (B == 0) return NTX_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const uint32_t blocks = min(ceil_div<uint32_t>(B, 256), 148u * 8u);
    if (D == 2) grid_debug_indices_kernel<2><<<blocks, 256, 0, st>>>(inputs, offsets, B, level, S, H, gridtype, align_corners != 0, idx_out);
    else grid_debug_indices_kernel<3><<<blocks, 256, 0, st>>>(inputs, offsets, B, level, S, H, gridtype, align_corners != 0, idx_out);
    return check_launch("grid_debug_indices");
}
