// grid_common.cuh — level table, index arithmetic and the lane-pair gather of the multi-resolution grid encoder,
// shared by grid.cu (stand-alone encoder) and field.cu (encoder fused with the MLPs).
#pragma once

#include "common.cuh"

namespace ntx {

// ---------------------------------------------------------------------------------------------------- level table
struct GridLevel {
    float scale;         // exp2f(level*S)*H - 1                      gridencoder.cu:126
    uint32_t res;        // ceil(scale)+1                              :127
    uint32_t hs;         // hashmap_size = offsets[l+1]-offsets[l]     :125
    uint32_t offset;     // offsets[l]
    uint32_t sy, sz;     // dense strides of y and z (0 if that dimension is dropped by the early loop exit, :60)
    uint32_t use_hash;   // gridtype==0 && stride>hashmap_size         :67
    uint32_t mask;       // hs-1 if hs is a power of two else 0 (then a real modulo is used)
};

__device__ __forceinline__ float level_scale(uint32_t level, float S, uint32_t H) {
    // written exactly as the reference writes it so that nvcc contracts it the same way (one FMA after exp2f)
    return exp2f(level * S) * H - 1.0f;
}

template <uint32_t D>
__device__ __forceinline__ GridLevel make_level(const int* __restrict__ offsets, uint32_t level, float S, uint32_t H,
                                                uint32_t gridtype, bool align) {
    GridLevel g;
    g.scale = level_scale(level, S, H);
    g.res = (uint32_t)ceilf(g.scale) + 1;
    g.offset = (uint32_t)offsets[level];
    g.hs = (uint32_t)offsets[level + 1] - g.offset;
    const uint32_t R = align ? g.res : g.res + 1;
    uint32_t stride = 1;  // uint32 wrap-around on purpose: mirrors get_grid_index (:56-63)
    g.sy = g.sz = 0;
    // d = 0 is always taken (stride 1 <= hs)
    stride *= R;
    if (D > 1 && stride <= g.hs) { g.sy = stride; stride *= R; if (D > 2 && stride <= g.hs) { g.sz = stride; stride *= R; } }
    g.use_hash = (gridtype == 0 && stride > g.hs) ? 1u : 0u;
    g.mask = ((g.hs & (g.hs - 1)) == 0) ? g.hs - 1 : 0u;
    return g;
}

__device__ __forceinline__ uint32_t wrap_index(const GridLevel& g, uint32_t index) {
    return g.mask ? (index & g.mask) : (index % g.hs);
}

template <uint32_t D>
__device__ __forceinline__ uint32_t corner_index(const GridLevel& g, const uint32_t* p) {
    uint32_t index;
    if (g.use_hash) {
        index = p[0];  // primes[0] == 1
        if (D > 1) index ^= p[1] * 2654435761u;
        if (D > 2) index ^= p[2] * 805459861u;
    } else {
        index = p[0];
        if (D > 1) index += p[1] * g.sy;
        if (D > 2) index += p[2] * g.sz;
    }
    return wrap_index(g, index);
}

// ---------------------------------------------------------------------------------------------------- lane-pair gather
// D = 3, C = 2.  Raw table element = 2 scalars: u32 for half2, u64 for float2.
template <typename scalar_t> struct Elem2;
template <> struct Elem2<__half> {
    using raw = uint32_t;
    static __device__ __forceinline__ raw load(const void* p) { return ld_table_u32(p); }
    static __device__ __forceinline__ raw xchg(raw v) { return __shfl_xor_sync(0xffffffffu, v, 1); }
    struct accum { __half a, b; };
    static __device__ __forceinline__ accum zero() { return {__float2half_rn(0.f), __float2half_rn(0.f)}; }
    static __device__ __forceinline__ void add(accum& r, float w, raw v) {
        const float2 f = half2_bits_to_float2(v);
        // p = half(w*g); r = half(float(r)+float(p))  per channel (c10::Half rounding points)
        const __half2 p = __floats2half2_rn(w * f.x, w * f.y);
        const float2 pf = __half22float2(p);
        r.a = __float2half_rn(__half2float(r.a) + pf.x);
        r.b = __float2half_rn(__half2float(r.b) + pf.y);
    }
    static __device__ __forceinline__ raw pack(const accum& r) { __half2 h = __halves2half2(r.a, r.b); return *reinterpret_cast<const uint32_t*>(&h); }
};
template <> struct Elem2<float> {
    using raw = uint64_t;
    static __device__ __forceinline__ raw load(const void* p) {
        raw r; asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(r) : "l"(p)); return r;
    }
    static __device__ __forceinline__ raw xchg(raw v) { return __shfl_xor_sync(0xffffffffu, v, 1); }
    struct accum { float a, b; };
    static __device__ __forceinline__ accum zero() { return {0.f, 0.f}; }
    static __device__ __forceinline__ void add(accum& r, float w, raw v) {
        const float lo = __uint_as_float((uint32_t)v), hi = __uint_as_float((uint32_t)(v >> 32));
        r.a = __fmaf_rn(w, lo, r.a);
        r.b = __fmaf_rn(w, hi, r.b);
    }
    static __device__ __forceinline__ raw pack(const accum& r) { return (uint64_t)__float_as_uint(r.a) | ((uint64_t)__float_as_uint(r.b) << 32); }
};

// A lane pair (lanes 2i, 2i+1) owns one sample; lane p (= lane & 1) gathers the four corners whose x bit is p.
// Processes the 4 consecutive levels lv[0..3]; on return lane p holds the finished features of levels 2p and 2p+1
// (packed[0], packed[1]), accumulated over the 8 corners in exactly the reference's order and rounding.
// Must be called by all 32 lanes of the warp (uses shuffles).  `live` = sample valid and inside [0,1]^3.
template <typename scalar_t>
__device__ __forceinline__ void pair_gather4(const float x, const float y, const float z, const bool live, const uint32_t p, const GridLevel* __restrict__ lv,
                                             const scalar_t* __restrict__ grid, const float half_off, typename Elem2<scalar_t>::raw packed[2]) {
    using E = Elem2<scalar_t>;
    using raw = typename E::raw;
    raw v[4][4];
    float fx[4], fy[4], fz[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const GridLevel g = lv[j];
        const float px = __fmaf_rn(x, g.scale, half_off), py = __fmaf_rn(y, g.scale, half_off), pz = __fmaf_rn(z, g.scale, half_off);
        const float flx = floorf(px), fly = floorf(py), flz = floorf(pz);
        fx[j] = px - flx; fy[j] = py - fly; fz[j] = pz - flz;
        const uint32_t cx = (uint32_t)flx + p, iy = (uint32_t)fly, iz = (uint32_t)flz;
        const scalar_t* gl = grid + (size_t)g.offset * 2;
        uint32_t ty0, ty1, tz0, tz1;
        if (g.use_hash) { ty0 = iy * 2654435761u; ty1 = ty0 + 2654435761u; tz0 = iz * 805459861u; tz1 = tz0 + 805459861u; }
        else { ty0 = iy * g.sy; ty1 = ty0 + g.sy; tz0 = iz * g.sz; tz1 = tz0 + g.sz; }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint32_t ty = (c & 1) ? ty1 : ty0, tz = (c & 2) ? tz1 : tz0;
            const uint32_t index = wrap_index(g, g.use_hash ? (cx ^ ty ^ tz) : (cx + ty + tz));
            // out-of-range / padding lanes read entry 0 of the level (always mapped) and discard it
            v[j][c] = E::load(gl + (size_t)(live ? index : 0u) * 2);
        }
    }
    // lane p accumulates levels 2p and 2p+1 of the batch; it needs the partner's corners of those levels
    raw o[2][4];
#pragma unroll
    for (int jj = 0; jj < 2; jj++)
#pragma unroll
        for (int c = 0; c < 4; c++) o[jj][c] = E::xchg(p ? v[jj][c] : v[2 + jj][c]);
#pragma unroll
    for (int jj = 0; jj < 2; jj++) {
        const float wfx = p ? fx[2 + jj] : fx[jj], wfy = p ? fy[2 + jj] : fy[jj], wfz = p ? fz[2 + jj] : fz[jj];
        typename E::accum r = E::zero();
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float wy = (c & 1) ? wfy : 1 - wfy, wz = (c & 2) ? wfz : 1 - wfz;
            const raw mine = p ? v[2 + jj][c] : v[jj][c];
            const raw x0v = p ? o[jj][c] : mine;   // corner with x bit 0
            const raw x1v = p ? mine : o[jj][c];   // corner with x bit 1
            // reference order: idx = bx + 2*by + 4*bz, w = ((1*wx)*wy)*wz   (gridencoder.cu:146-167)
            E::add(r, ((1 - wfx) * wy) * wz, x0v);
            E::add(r, (wfx * wy) * wz, x1v);
        }
        packed[jj] = live ? E::pack(r) : (raw)0;
    }
}

}  // namespace ntx
