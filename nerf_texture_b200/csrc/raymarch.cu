// raymarch.cu — occupancy-grid ray marching, compositing and ray compaction for sm_100a.
// Replaces raymarching/src/raymarching.cu of the reference.  Arithmetic per ray is the reference's; what changes:
//   * every place the reference claims output space with a global atomicAdd per ray (march_rays_train :408-409,
//     compact_rays :1130) uses a warp-shuffle / block scan + decoupled look-back across blocks instead, so segment
//     offsets and compacted rays come out in ascending order: deterministic, and adjacent slots stay adjacent
//     rays, which keeps the downstream hash-grid gather coherent;
//   * march_rays can zero its own unused slots (saves three cudaMemset per render-loop iteration);
//   * launches go to the caller's stream and are error-checked.
#include "common.cuh"

namespace ntx {

__device__ __forceinline__ float clampf(const float x, const float lo, const float hi) { return fminf(hi, fmaxf(lo, x)); }
__device__ __forceinline__ float signf1(const float x) { return copysignf(1.0f, x); }
constexpr float kSqrt3 = 1.7320508075688772f;
constexpr float kRPi = 0.3183098861837907f;

// raymarching.cu:44-56
__device__ __forceinline__ int mip_from_pos(const float x, const float y, const float z, const float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int exponent;
    frexpf(mx, &exponent);
    return (int)fminf(max_cascade - 1, fmaxf(0, exponent));
}
__device__ __forceinline__ int mip_from_dt(const float dt, const float H, const float max_cascade) {
    const float mx = dt * H * 0.5;
    int exponent;
    frexpf(mx, &exponent);
    return (int)fminf(max_cascade - 1, fmaxf(0, exponent));
}
// raymarching.cu:58-83
__host__ __device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t morton3D_1(uint32_t x, uint32_t y, uint32_t z) { return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2); }
__host__ __device__ __forceinline__ uint32_t morton3D_invert_1(uint32_t x) {
    x = x & 0x49249249;
    x = (x | (x >> 2)) & 0xc30c30c3;
    x = (x | (x >> 4)) & 0x0f00f00f;
    x = (x | (x >> 8)) & 0xff0000ff;
    x = (x | (x >> 16)) & 0x0000ffff;
    return x;
}

// PCG-XSH-RR 64/32 (pcg32.h:44-170)
struct Pcg32 {
    uint64_t state, inc;
    static constexpr uint64_t kMult = 0x5851f42d4c957f2dULL;
    __device__ __forceinline__ uint32_t next_uint() {
        const uint64_t old = state;
        state = old * kMult + inc;
        const uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        const uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    __device__ __forceinline__ void seed(uint64_t initstate, uint64_t initseq = 1) {
        state = 0U; inc = (initseq << 1u) | 1u;
        next_uint(); state += initstate; next_uint();
    }
    __device__ __forceinline__ float next_float() { return __uint_as_float((next_uint() >> 9) | 0x3f800000u) - 1.0f; }
    __device__ __forceinline__ void advance(uint64_t delta) {
        uint64_t cur_mult = kMult, cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
        while (delta > 0) {
            if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
            cur_plus = (cur_mult + 1) * cur_plus;
            cur_mult *= cur_mult;
            delta /= 2;
        }
        state = acc_mult * state + acc_plus;
    }
};

// ---------------------------------------------------------------------------------------------------- DDA probe
struct Ray {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
};
struct MarchParams {
    float bound, dt_gamma, dt_min, dt_max, rH;
    uint32_t C, H;
    const uint8_t* __restrict__ grid;
    // single cascade + power-of-two grid (NeRF-Texture: bound 1, H 128): the cascade level is always 0 and the voxel index can be
    // computed in fp32 with bit-identical results (multiplying by 0.5*H is an exact scaling, so the reference's detour through
    // double precision in `0.5 * (x * rbound + 1) * H` changes nothing)
    bool fast;
    float mb0, rb0, halfH;
};

__device__ __forceinline__ MarchParams make_march_params(float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid) {
    MarchParams p;
    p.bound = bound; p.dt_gamma = dt_gamma;
    p.dt_min = 2 * kSqrt3 / max_steps;               // raymarching.cu:346
    p.dt_max = 2 * kSqrt3 * (1 << (C - 1)) / H;      // :347
    p.rH = 1 / (float)H;
    p.C = C; p.H = H; p.grid = grid;
    p.fast = (C == 1) && ((H & (H - 1)) == 0);
    p.mb0 = fminf(1.0f, bound);
    p.rb0 = 1 / p.mb0;
    p.halfH = 0.5f * (float)H;
    return p;
}
__device__ __forceinline__ Ray load_ray(const float* __restrict__ o, const float* __restrict__ d) {
    Ray r;
    r.ox = o[0]; r.oy = o[1]; r.oz = o[2]; r.dx = d[0]; r.dy = d[1]; r.dz = d[2];
    r.rdx = 1 / r.dx; r.rdy = 1 / r.dy; r.rdz = 1 / r.dz;
    return r;
}

// One step of the marcher (raymarching.cu:362-402).  Returns true if the voxel containing t is occupied (x,y,z,dt valid);
// otherwise advances t to the first dt-quantised position beyond the voxel.
__device__ __forceinline__ bool probe(const Ray& r, const MarchParams& p, float& t, float& x, float& y, float& z, float& dt) {
    x = clampf(r.ox + t * r.dx, -p.bound, p.bound);
    y = clampf(r.oy + t * r.dy, -p.bound, p.bound);
    z = clampf(r.oz + t * r.dz, -p.bound, p.bound);
    dt = clampf(t * p.dt_gamma, p.dt_min, p.dt_max);
    if (p.fast) {   // level == 0, fp32 index math (see MarchParams::fast)
        const int nx = clampf((x * p.rb0 + 1) * p.halfH, 0.0f, (float)(p.H - 1));
        const int ny = clampf((y * p.rb0 + 1) * p.halfH, 0.0f, (float)(p.H - 1));
        const int nz = clampf((z * p.rb0 + 1) * p.halfH, 0.0f, (float)(p.H - 1));
        const uint32_t index = morton3D_1(nx, ny, nz);
        if (p.grid[index >> 3] & (1u << (index & 7u))) return true;
        const float tx = (((nx + 0.5f + 0.5f * signf1(r.dx)) * p.rH * 2 - 1) * p.mb0 - x) * r.rdx;
        const float ty = (((ny + 0.5f + 0.5f * signf1(r.dy)) * p.rH * 2 - 1) * p.mb0 - y) * r.rdy;
        const float tz = (((nz + 0.5f + 0.5f * signf1(r.dz)) * p.rH * 2 - 1) * p.mb0 - z) * r.rdz;
        const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
        // dt_gamma == 0: clamp(t * 0, dt_min, dt_max) is dt_min for every finite t — same additions, two instructions less per lattice step
        if (p.dt_gamma == 0.0f) { do { t += p.dt_min; } while (t < tt); }
        else { do { t += clampf(t * p.dt_gamma, p.dt_min, p.dt_max); } while (t < tt); }
        return false;
    }
    const int level = max(mip_from_pos(x, y, z, p.C), mip_from_dt(dt, p.H, p.C));
    const float mip_bound = fminf((float)(1 << level), p.bound);
    const float mip_rbound = 1 / mip_bound;
    const int nx = clampf(0.5 * (x * mip_rbound + 1) * p.H, 0.0f, (float)(p.H - 1));
    const int ny = clampf(0.5 * (y * mip_rbound + 1) * p.H, 0.0f, (float)(p.H - 1));
    const int nz = clampf(0.5 * (z * mip_rbound + 1) * p.H, 0.0f, (float)(p.H - 1));
    const uint32_t index = level * p.H * p.H * p.H + morton3D_1(nx, ny, nz);
    const bool occ = p.grid[index / 8] & (1 << (index % 8));
    if (occ) return true;
    const float tx = (((nx + 0.5f + 0.5f * signf1(r.dx)) * p.rH * 2 - 1) * mip_bound - x) * r.rdx;
    const float ty = (((ny + 0.5f + 0.5f * signf1(r.dy)) * p.rH * 2 - 1) * mip_bound - y) * r.rdy;
    const float tz = (((nz + 0.5f + 0.5f * signf1(r.dz)) * p.rH * 2 - 1) * mip_bound - z) * r.rdz;
    const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    do { t += clampf(t * p.dt_gamma, p.dt_min, p.dt_max); } while (t < tt);
    return false;
}

// The address part of probe() only: sample position, step length and the occupancy bit index at parameter t (same
// expressions as probe(), so the values are bit-identical).  Used to issue several occupancy loads back to back.
__device__ __forceinline__ uint32_t locate(const Ray& r, const MarchParams& p, const float t, float& x, float& y, float& z, float& dt) {
    x = clampf(r.ox + t * r.dx, -p.bound, p.bound);
    y = clampf(r.oy + t * r.dy, -p.bound, p.bound);
    z = clampf(r.oz + t * r.dz, -p.bound, p.bound);
    dt = clampf(t * p.dt_gamma, p.dt_min, p.dt_max);
    if (p.fast) {
        const int nx = clampf((x * p.rb0 + 1) * p.halfH, 0.0f, (float)(p.H - 1));
        const int ny = clampf((y * p.rb0 + 1) * p.halfH, 0.0f, (float)(p.H - 1));
        const int nz = clampf((z * p.rb0 + 1) * p.halfH, 0.0f, (float)(p.H - 1));
        return morton3D_1(nx, ny, nz);
    }
    const int level = max(mip_from_pos(x, y, z, p.C), mip_from_dt(dt, p.H, p.C));
    const float mip_bound = fminf((float)(1 << level), p.bound);
    const float mip_rbound = 1 / mip_bound;
    const int nx = clampf(0.5 * (x * mip_rbound + 1) * p.H, 0.0f, (float)(p.H - 1));
    const int ny = clampf(0.5 * (y * mip_rbound + 1) * p.H, 0.0f, (float)(p.H - 1));
    const int nz = clampf(0.5 * (z * mip_rbound + 1) * p.H, 0.0f, (float)(p.H - 1));
    return level * p.H * p.H * p.H + morton3D_1(nx, ny, nz);
}

// ---------------------------------------------------------------------------------------------------- conservative occupancy mip
// A ray that has left the object keeps marching to `far` voxel by voxel in the reference (one dependent probe per empty voxel);
// it emits nothing on the way, and because it then comes back with fewer than n_step samples composite_rays marks it dead, so
// the exact t it stops at is unobservable.  The mip lets the marcher prove "no occupied voxel can be probed on [t, far)" and stop.
//   coarse bit (cascade k, cell c) = OR of the 4x4x4 fine voxels of that cell = OR of 8 consecutive bytes (Morton order),
//   then dilated by one coarse cell in every direction, so that testing sample points spaced half a coarse cell apart is
//   conservative for every point in between (and for the marcher's clamping at cascade borders).
// Buffer: [header 512 B: int bbox[C][6] of the occupied coarse cells][dilated: C*n/8 bytes][raw: C*n/8 bytes], n = (H/4)^3 (Morton order).
constexpr uint32_t kMipShift = 2;                       // coarse cell = 4 fine voxels per side
constexpr uint32_t kMipCell = 1u << kMipShift;
constexpr uint32_t kMipHeaderBytes = 512;               // 16 cascades x 6 ints

__global__ void occupancy_mip_init_kernel(int* __restrict__ bbox, const uint32_t C) {
    const uint32_t i = threadIdx.x;
    if (i < C * 6) bbox[i] = (i % 6 < 3) ? 0x7fffffff : -0x7fffffff;   // min x,y,z | max x,y,z
}
__global__ void __launch_bounds__(128) occupancy_mip_raw_kernel(const uint8_t* __restrict__ grid, const uint32_t cells_total, const uint32_t cells_per_cascade,
                                                                uint8_t* __restrict__ raw, int* __restrict__ bbox) {
    const uint32_t byte = blockIdx.x * blockDim.x + threadIdx.x;       // one output byte = 8 coarse cells = 8 x 8 fine bytes
    if (byte >= cells_total / 8) return;
    const uint4* src = reinterpret_cast<const uint4*>(grid + (size_t)byte * 64);
    uint32_t bits = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint4 v = src[q];
        bits |= ((v.x | v.y) ? 1u : 0u) << (2 * q);
        bits |= ((v.z | v.w) ? 1u : 0u) << (2 * q + 1);
    }
    raw[byte] = (uint8_t)bits;
    if (bits) {
        const uint32_t k = (byte * 8) / cells_per_cascade;
        int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-1, -1, -1};
        for (uint32_t c = 0; c < 8; c++)
            if (bits & (1u << c)) {
                const uint32_t m = byte * 8 + c - k * cells_per_cascade;
                const int q[3] = {(int)morton3D_invert_1(m), (int)morton3D_invert_1(m >> 1), (int)morton3D_invert_1(m >> 2)};
                for (int a = 0; a < 3; a++) { lo[a] = min(lo[a], q[a]); hi[a] = max(hi[a], q[a]); }
            }
        for (int a = 0; a < 3; a++) { atomicMin(&bbox[k * 6 + a], lo[a]); atomicMax(&bbox[k * 6 + 3 + a], hi[a]); }
    }
}
__global__ void __launch_bounds__(128) occupancy_mip_dilate_kernel(const uint8_t* __restrict__ raw, const uint32_t C, const uint32_t Hc, uint8_t* __restrict__ dil) {
    const uint32_t n = Hc * Hc * Hc;
    const uint32_t byte = blockIdx.x * blockDim.x + threadIdx.x;
    if (byte >= C * n / 8) return;
    const uint32_t k = (byte * 8) / n;
    uint32_t bits = 0;
    for (uint32_t c = 0; c < 8; c++) {
        const uint32_t m = byte * 8 + c - k * n;
        const int x = morton3D_invert_1(m), y = morton3D_invert_1(m >> 1), z = morton3D_invert_1(m >> 2);
        uint32_t any = 0;
        for (int dz = -1; dz <= 1; dz++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    const int xx = x + dx, yy = y + dy, zz = z + dz;
                    if (xx < 0 || yy < 0 || zz < 0 || xx >= (int)Hc || yy >= (int)Hc || zz >= (int)Hc) continue;
                    const uint32_t idx = k * n + morton3D_1(xx, yy, zz);
                    any |= raw[idx >> 3] & (1u << (idx & 7u));
                }
        bits |= (any ? 1u : 0u) << c;
    }
    dil[byte] = (uint8_t)bits;
}

// Parameter interval [*t_lo, *t_hi] of the ray inside the union of the per-cascade bounding boxes of occupied coarse cells, each
// grown by one coarse cell (conservative w.r.t. clamping and rounding).  Returns false if the ray misses all of them: then no
// occupied voxel can ever be probed on this ray.
__device__ __forceinline__ bool clip_to_occupied(const Ray& r, const MarchParams& p, const uint8_t* __restrict__ mip, float& t_lo, float& t_hi) {
    const int* bbox = reinterpret_cast<const int*>(mip);
    const float Hc = (float)(p.H >> kMipShift);
    bool any = false;
    t_lo = 3.402823466e+38f; t_hi = -3.402823466e+38f;
    for (uint32_t k = 0; k < p.C; k++) {
        const int* b = bbox + k * 6;
        if (b[0] > b[3]) continue;                                // cascade has no occupied cell
        const float mb = fminf((float)(1 << k), p.bound), cell = 2.0f * mb / Hc;
        float lo = -3.402823466e+38f, hi = 3.402823466e+38f;
        const float o[3] = {r.ox, r.oy, r.oz}, rd[3] = {r.rdx, r.rdy, r.rdz}, d[3] = {r.dx, r.dy, r.dz};
        bool miss = false;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            float wmin = ((float)(b[a] - 1)) * cell - mb, wmax = ((float)(b[3 + a] + 2)) * cell - mb;   // grown by one cell each side
            if (k + 1 == p.C || (float)(1 << k) >= p.bound) { if (b[a] == 0) wmin = -3.402823466e+38f; if (b[3 + a] == (int)Hc - 1) wmax = 3.402823466e+38f; }  // outermost cascade: clamped points land in border cells
            if (d[a] == 0.0f) { if (o[a] < wmin || o[a] > wmax) miss = true; }
            else {
                float t0 = (wmin - o[a]) * rd[a], t1 = (wmax - o[a]) * rd[a];
                if (t0 > t1) { const float c = t0; t0 = t1; t1 = c; }
                lo = fmaxf(lo, t0); hi = fminf(hi, t1);
            }
        }
        if (miss || lo > hi) continue;
        any = true;
        t_lo = fminf(t_lo, lo); t_hi = fmaxf(t_hi, hi);
    }
    return any;
}

// true if an occupied voxel MAY be probed by the marcher anywhere on the ray segment [t, far)
__device__ __forceinline__ bool maybe_occupied_ahead(const Ray& r, const MarchParams& p, const uint8_t* __restrict__ mip, float t, const float far) {
    const uint8_t* __restrict__ coarse = mip + kMipHeaderBytes;
    const uint32_t Hc = p.H >> kMipShift, n = Hc * Hc * Hc;
    const float inv_speed = 1.0f / fmaxf(fmaxf(fabsf(r.dx), fabsf(r.dy)), fmaxf(fabsf(r.dz), 1e-20f));
    if (p.C == 1) {
        // single cascade: cell coordinates are affine in t; test 4 sample points per round so that their loads overlap
        const float mb = p.mb0, s = 0.5f * (float)Hc / mb, top = (float)(Hc - 1);
        const float ax = (r.ox + mb) * s, ay = (r.oy + mb) * s, az = (r.oz + mb) * s, bx = r.dx * s, by = r.dy * s, bz = r.dz * s;
        const float dt = ((float)kMipCell * mb / (float)p.H) * inv_speed;
        while (t < far) {
            uint32_t hit = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float tj = t + (float)j * dt;
                const int cx = (int)clampf(ax + tj * bx, 0.0f, top), cy = (int)clampf(ay + tj * by, 0.0f, top), cz = (int)clampf(az + tj * bz, 0.0f, top);
                const uint32_t idx = morton3D_1(cx, cy, cz);
                hit |= (tj < far) ? (coarse[idx >> 3] & (1u << (idx & 7u))) : 0u;
            }
            if (hit) return true;
            t += 4.0f * dt;
        }
        return false;
    }
    while (t < far) {
        const float x = r.ox + t * r.dx, y = r.oy + t * r.dy, z = r.oz + t * r.dz;
        for (uint32_t k = 0; k < p.C; k++) {
            const float mb = fminf((float)(1 << k), p.bound), s = 0.5f * (float)Hc / mb;
            const int cx = min(max((int)floorf((clampf(x, -mb, mb) + mb) * s), 0), (int)Hc - 1);
            const int cy = min(max((int)floorf((clampf(y, -mb, mb) + mb) * s), 0), (int)Hc - 1);
            const int cz = min(max((int)floorf((clampf(z, -mb, mb) + mb) * s), 0), (int)Hc - 1);
            const uint32_t idx = k * n + morton3D_1(cx, cy, cz);
            if (coarse[idx >> 3] & (1u << (idx & 7u))) return true;
        }
        // advance half a coarse cell (of the cascade this point lies in) along the fastest axis
        const int lp = mip_from_pos(clampf(x, -p.bound, p.bound), clampf(y, -p.bound, p.bound), clampf(z, -p.bound, p.bound), p.C);
        t += ((float)kMipCell * fminf((float)(1 << lp), p.bound) / (float)p.H) * inv_speed;   // half of (kMipCell * 2*mb/H)
    }
    return false;
}

// The ray parameter beyond which the marcher can no longer probe an occupied voxel (same sampling and the same conservativeness
// argument as maybe_occupied_ahead: the dilated mip is tested at points half a coarse cell apart, so if neither of two consecutive
// points is set there is no occupied voxel between them), or -1 if there is none on [t, far) at all.  Evaluated ONCE per ray and
// frame (frame_init_kernel): every lane of a warp runs the same loop, instead of the marcher interrupting its walk for a scan
// whenever one of its 32 rays is due for one.
__device__ __forceinline__ float last_maybe_occupied(const Ray& r, const MarchParams& p, const uint8_t* __restrict__ mip, float t, const float far) {
    const uint8_t* __restrict__ coarse = mip + kMipHeaderBytes;
    const uint32_t Hc = p.H >> kMipShift, n = Hc * Hc * Hc;
    const float inv_speed = 1.0f / fmaxf(fmaxf(fabsf(r.dx), fabsf(r.dy)), fmaxf(fabsf(r.dz), 1e-20f));
    float last = -1.0f;
    if (p.C == 1) {
        const float mb = p.mb0, s = 0.5f * (float)Hc / mb, top = (float)(Hc - 1);
        const float ax = (r.ox + mb) * s, ay = (r.oy + mb) * s, az = (r.oz + mb) * s, bx = r.dx * s, by = r.dy * s, bz = r.dz * s;
        const float dt = ((float)kMipCell * mb / (float)p.H) * inv_speed;
        while (t < far) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float tj = t + (float)j * dt;
                const int cx = (int)clampf(ax + tj * bx, 0.0f, top), cy = (int)clampf(ay + tj * by, 0.0f, top), cz = (int)clampf(az + tj * bz, 0.0f, top);
                const uint32_t idx = morton3D_1(cx, cy, cz);
                if (tj < far && (coarse[idx >> 3] & (1u << (idx & 7u)))) last = tj + dt;
            }
            t += 4.0f * dt;
        }
        return last;
    }
    while (t < far) {
        const float x = r.ox + t * r.dx, y = r.oy + t * r.dy, z = r.oz + t * r.dz;
        const int lp = mip_from_pos(clampf(x, -p.bound, p.bound), clampf(y, -p.bound, p.bound), clampf(z, -p.bound, p.bound), p.C);
        const float step = ((float)kMipCell * fminf((float)(1 << lp), p.bound) / (float)p.H) * inv_speed;
        for (uint32_t k = 0; k < p.C; k++) {
            const float mb = fminf((float)(1 << k), p.bound), s = 0.5f * (float)Hc / mb;
            const int cx = min(max((int)floorf((clampf(x, -mb, mb) + mb) * s), 0), (int)Hc - 1);
            const int cy = min(max((int)floorf((clampf(y, -mb, mb) + mb) * s), 0), (int)Hc - 1);
            const int cz = min(max((int)floorf((clampf(z, -mb, mb) + mb) * s), 0), (int)Hc - 1);
            const uint32_t idx = k * n + morton3D_1(cx, cy, cz);
            if (coarse[idx >> 3] & (1u << (idx & 7u))) last = t + step;
        }
        t += step;
    }
    return last;
}

// ---------------------------------------------------------------------------------------------------- ordered grid-wide scan
// Workspace layout: u32 ticket, u32 done, then one u64 status word per block:
//   bits 63..62: 0 = nothing yet, 1 = block aggregate, 2 = inclusive prefix;  bits 61..0: value.
// Blocks take dynamic tickets so that a block only ever waits on blocks that are already running.
struct ScanWS { uint32_t ticket, done; unsigned long long state[1]; };
constexpr unsigned long long kScanAgg = 1ull << 62, kScanIncl = 2ull << 62, kScanMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// block-wide exclusive scan of one value per thread (blockDim.x <= 1024, multiple of 32); returns exclusive prefix, total in `total`
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t& total, uint32_t* warp_sums /* smem[32] */) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < nw ? warp_sums[lane] : 0, si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, si, o); if (lane >= o) si += n; }
        warp_sums[lane] = si - s;  // exclusive prefix of warp sums
        if (lane == 31) warp_sums[32] = si;
    }
    __syncthreads();
    total = warp_sums[32];
    const uint32_t r = warp_sums[wid] + inc - v;
    __syncthreads();
    return r;
}

// called by all threads of the block; returns the exclusive prefix of `agg` over blocks in ticket order (+ base for everything).
// `base` only needs to be valid in the block holding ticket 0.
__device__ __forceinline__ unsigned long long chained_prefix(ScanWS* ws, uint32_t ticket, unsigned long long agg, unsigned long long base,
                                                             unsigned long long* smem_bcast) {
    if (threadIdx.x < 32) {
        const uint32_t lane = threadIdx.x;
        unsigned long long excl = 0;
        if (ticket == 0) {
            excl = base;
            if (lane == 0) st_release_u64(&ws->state[0], kScanIncl | (base + agg));
        } else {
            if (lane == 0) st_release_u64(&ws->state[ticket], kScanAgg | agg);
            int j = (int)ticket - 1;
            while (true) {
                const int idx = j - (int)lane;
                unsigned long long s = kScanIncl;  // virtual predecessor of block 0 (never selected: block 0 is inclusive)
                if (idx >= 0) { do { s = ld_acquire_u64(&ws->state[idx]); } while ((s >> 62) == 0); }
                const uint32_t incl = __ballot_sync(0xffffffffu, (s >> 62) == 2);
                const int first = incl ? (__ffs(incl) - 1) : 31;
                unsigned long long val = ((int)lane <= first) ? (s & kScanMask) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                excl += val;
                if (incl) break;
                j -= 32;
            }
            if (lane == 0) st_release_u64(&ws->state[ticket], kScanIncl | (excl + agg));
        }
        if (lane == 0) *smem_bcast = excl;
    }
    __syncthreads();
    const unsigned long long r = *smem_bcast;
    __syncthreads();
    return r;
}

// every block calls this once at its very end; the last block to arrive restores the workspace to all-zero
__device__ __forceinline__ void scan_ws_release(ScanWS* ws, uint32_t nblocks) {
    __syncthreads();
    __shared__ uint32_t s_last;
    if (threadIdx.x == 0) { __threadfence(); s_last = (atomicAdd(&ws->done, 1u) == nblocks - 1) ? 1u : 0u; }
    __syncthreads();
    if (s_last) {
        for (uint32_t i = threadIdx.x; i < nblocks; i += blockDim.x) ws->state[i] = 0ull;
        if (threadIdx.x == 0) { ws->ticket = 0; ws->done = 0; }
    }
}

// ---------------------------------------------------------------------------------------------------- small utilities
__global__ void __launch_bounds__(128) near_far_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ aabb,
                                                       const uint32_t N, const float min_near, float* __restrict__ nears, float* __restrict__ fars) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float rdx = 1 / rays_d[n * 3], rdy = 1 / rays_d[n * 3 + 1], rdz = 1 / rays_d[n * 3 + 2];
    float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx;
    if (near > far) { const float c = near; near = far; far = c; }
    float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
    if (near_y > far_y) { const float c = near_y; near_y = far_y; far_y = c; }
    if (near > far_y || near_y > far) { nears[n] = fars[n] = 3.402823466e+38f; return; }
    if (near_y > near) near = near_y;
    if (far_y < far) far = far_y;
    float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
    if (near_z > far_z) { const float c = near_z; near_z = far_z; far_z = c; }
    if (near > far_z || near_z > far) { nears[n] = fars[n] = 3.402823466e+38f; return; }
    if (near_z > near) near = near_z;
    if (far_z < far) far = far_z;
    if (near < min_near) near = min_near;
    nears[n] = near;
    fars[n] = far;
}

__global__ void __launch_bounds__(128) polar_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float radius, const uint32_t N,
                                                    float* __restrict__ coords) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float Bq = ox * dx + oy * dy + oz * dz;
    const float Cq = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-Bq + sqrtf(Bq * Bq - A * Cq)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    const float theta = atan2(sqrtf(x * x + z * z), y);
    const float phi = atan2(z, x);
    coords[n * 2] = 2 * theta * kRPi - 1;
    coords[n * 2 + 1] = phi * kRPi;
}

__global__ void __launch_bounds__(128) morton3D_kernel(const int* __restrict__ coords, const uint32_t N, int* __restrict__ indices) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    indices[n] = morton3D_1(coords[n * 3], coords[n * 3 + 1], coords[n * 3 + 2]);
}
__global__ void __launch_bounds__(128) morton3D_invert_kernel(const int* __restrict__ indices, const uint32_t N, int* __restrict__ coords) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const int ind = indices[n];
    coords[n * 3] = morton3D_invert_1(ind >> 0);
    coords[n * 3 + 1] = morton3D_invert_1(ind >> 1);
    coords[n * 3 + 2] = morton3D_invert_1(ind >> 2);
}

// 8 densities -> 1 byte (raymarching.cu:270-291); each thread packs 32 densities (one 32-bit word) with 128-bit loads
__global__ void __launch_bounds__(256) packbits_kernel(const float* __restrict__ grid, const uint32_t N, float thresh, uint8_t* __restrict__ bitfield,
                                                       const float* __restrict__ thresh_dev) {
    if (thresh_dev) thresh = *thresh_dev;       // threshold produced on the device by the kernel before (density-grid maintenance)
    const uint32_t words = N / 4;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += gridDim.x * blockDim.x) {
        const uint4* src = reinterpret_cast<const uint4*>(grid + (size_t)w * 32);
        uint32_t bits = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint4 v = ld_stream_u4(src + q);
            bits |= (__uint_as_float(v.x) > thresh ? 1u : 0u) << (q * 4 + 0);
            bits |= (__uint_as_float(v.y) > thresh ? 1u : 0u) << (q * 4 + 1);
            bits |= (__uint_as_float(v.z) > thresh ? 1u : 0u) << (q * 4 + 2);
            bits |= (__uint_as_float(v.w) > thresh ? 1u : 0u) << (q * 4 + 3);
        }
        reinterpret_cast<uint32_t*>(bitfield)[w] = bits;
    }
    // tail bytes (N not a multiple of 4) and unaligned inputs are handled by the scalar kernel below
}
__global__ void __launch_bounds__(128) packbits_scalar_kernel(const float* __restrict__ grid, const uint32_t n0, const uint32_t N, float thresh,
                                                              uint8_t* __restrict__ bitfield, const float* __restrict__ thresh_dev) {
    if (thresh_dev) thresh = *thresh_dev;
    const uint32_t n = n0 + threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    uint8_t bits = 0;
#pragma unroll
    for (uint8_t i = 0; i < 8; i++) bits |= (grid[(size_t)n * 8 + i] > thresh) ? ((uint8_t)1 << i) : 0;
    bitfield[n] = bits;
}

// ---------------------------------------------------------------------------------------------------- training marcher
constexpr int kTrainThreads = 128;

__global__ void __launch_bounds__(kTrainThreads) march_rays_train_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, const uint8_t* __restrict__ grid, const float bound, const float dt_gamma,
    const uint32_t max_steps, const uint32_t N, const uint32_t C, const uint32_t H, const uint32_t M, const float* __restrict__ nears,
    const float* __restrict__ fars, float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas, float* __restrict__ rays_ts,
    int* __restrict__ rays, int* __restrict__ counter, const uint32_t perturb, ScanWS* ws) {
    __shared__ uint32_t s_ticket;
    __shared__ uint32_t s_warp[33];
    __shared__ unsigned long long s_bcast;
    if (threadIdx.x == 0) s_ticket = atomicAdd(&ws->ticket, 1u);
    __syncthreads();
    const uint32_t ticket = s_ticket;
    const uint32_t n = ticket * kTrainThreads + threadIdx.x;
    const bool valid = n < N;
    const MarchParams p = make_march_params(bound, dt_gamma, max_steps, C, H, grid);

    Ray r{};
    float far = 0.f, t0 = 0.f;
    uint32_t num_steps = 0;
    if (valid) {
        r = load_ray(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
        far = fars[n];
        t0 = nears[n];
        if (perturb) {
            Pcg32 rng; rng.seed(42);          // hard-coded seed, raymarching.cu:488
            rng.advance(n);
            t0 += p.dt_min * rng.next_float();
        }
        // pass 1: count the occupied steps (:362-403)
        float t = t0, x, y, z, dt;
        while (t < far && num_steps < max_steps) {
            if (probe(r, p, t, x, y, z, dt)) { num_steps++; t += dt; }
        }
    }
    // ordered allocation of the output segment (replaces the two atomicAdds at :408-409)
    uint32_t block_total;
    const uint32_t local = block_exclusive_scan(num_steps, block_total, s_warp);
    const unsigned long long base = (ticket == 0) ? (unsigned long long)(uint32_t)counter[0] : 0ull;
    const unsigned long long excl = chained_prefix(ws, ticket, block_total, base, &s_bcast);
    const uint32_t nblocks = gridDim.x;
    if (ticket == nblocks - 1 && threadIdx.x == 0) {
        // excl already contains the caller's starting value of counter[0]
        const uint32_t ray_base = (uint32_t)counter[1];
        counter[0] = (int)(uint32_t)(excl + block_total);
        counter[1] = (int)(ray_base + N);
    }
    if (valid) {
        const uint32_t point_index = (uint32_t)excl + local;
        // NOTE: the reference adds counter[1]'s starting value to the row index; callers always pass 0 (renderer.py:367),
        // and a non-zero start would index rays[] out of bounds there, so row == n here.
        rays[(size_t)n * 3] = (int)n;
        rays[(size_t)n * 3 + 1] = (int)point_index;
        rays[(size_t)n * 3 + 2] = (int)num_steps;
        if (num_steps != 0 && point_index + num_steps < M) {   // :418-419
            float* px = xyzs + (size_t)point_index * 3;
            float* pd = dirs + (size_t)point_index * 3;
            float* pl = deltas + (size_t)point_index * 2;
            float* pt = rays_ts ? rays_ts + point_index : nullptr;
            float t = t0, last_t = t0, x, y, z, dt;
            uint32_t step = 0;
            while (t < far && step < num_steps) {               // pass 2 (:430-482)
                if (probe(r, p, t, x, y, z, dt)) {
                    px[0] = x; px[1] = y; px[2] = z;
                    pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
                    t += dt;
                    pl[0] = dt; pl[1] = t - last_t;
                    if (pt) { pt[0] = t; pt++; }
                    last_t = t;
                    px += 3; pd += 3; pl += 2; step++;
                }
            }
        }
    }
    scan_ws_release(ws, nblocks);
}

// raymarching.cu:700-777
__global__ void __launch_bounds__(128) composite_train_fwd_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
                                                                  const int* __restrict__ rays, const uint32_t M, const uint32_t N,
                                                                  float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const uint32_t index = rays[n * 3], offset = rays[n * 3 + 1], num_steps = rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps >= M) {
        weights_sum[index] = 0; depth[index] = 0;
        image[index * 3] = 0; image[index * 3 + 1] = 0; image[index * 3 + 2] = 0;
        return;
    }
    sigmas += offset; rgbs += (size_t)offset * 3; deltas += (size_t)offset * 2;
    float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
    for (uint32_t step = 0; step < num_steps; step++) {
        const float alpha = 1.0f - __expf(-sigmas[0] * deltas[0]);
        const float weight = alpha * T;
        r += weight * rgbs[0]; g += weight * rgbs[1]; b += weight * rgbs[2];
        t += deltas[1];
        d += weight * t;
        ws += weight;
        T *= 1.0f - alpha;
        sigmas++; rgbs += 3; deltas += 2;
    }
    weights_sum[index] = ws; depth[index] = d;
    image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
}

// raymarching.cu:802-881
__global__ void __launch_bounds__(128) composite_train_bwd_kernel(const float* __restrict__ grad_weights_sum, const float* __restrict__ grad_image,
                                                                  const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
                                                                  const int* __restrict__ rays, const float* __restrict__ weights_sum, const float* __restrict__ image,
                                                                  const uint32_t M, const uint32_t N, float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const uint32_t index = rays[n * 3], offset = rays[n * 3 + 1], num_steps = rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps >= M) return;
    const float gws = grad_weights_sum[index];
    const float gi0 = grad_image[index * 3], gi1 = grad_image[index * 3 + 1], gi2 = grad_image[index * 3 + 2];
    const float r_final = image[index * 3], g_final = image[index * 3 + 1], b_final = image[index * 3 + 2], ws_final = weights_sum[index];
    sigmas += offset; rgbs += (size_t)offset * 3; deltas += (size_t)offset * 2; grad_sigmas += offset; grad_rgbs += (size_t)offset * 3;
    float T = 1.0f, r = 0, g = 0, b = 0, ws = 0;
    for (uint32_t step = 0; step < num_steps; step++) {
        const float alpha = 1.0f - __expf(-sigmas[0] * deltas[0]);
        const float weight = alpha * T;
        r += weight * rgbs[0]; g += weight * rgbs[1]; b += weight * rgbs[2];
        ws += weight;
        T *= 1.0f - alpha;
        grad_rgbs[0] = gi0 * weight; grad_rgbs[1] = gi1 * weight; grad_rgbs[2] = gi2 * weight;
        grad_sigmas[0] = deltas[0] * (gi0 * (T * rgbs[0] - (r_final - r)) + gi1 * (T * rgbs[1] - (g_final - g)) + gi2 * (T * rgbs[2] - (b_final - b)) +
                                      gws * (T - (ws_final - ws)));
        sigmas++; rgbs += 3; deltas += 2; grad_sigmas++; grad_rgbs += 3;
    }
}

// ---------------------------------------------------------------------------------------------------- inference loop body
// raymarching.cu:900-1006
__global__ void __launch_bounds__(128) march_rays_kernel(const uint32_t n_alive, const uint32_t n_step, const int* __restrict__ rays_alive,
                                                         const float* __restrict__ rays_t, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                         const float bound, const float dt_gamma, const uint32_t max_steps, const uint32_t C, const uint32_t H,
                                                         const uint8_t* __restrict__ grid, const float* __restrict__ nears, const float* __restrict__ fars,
                                                         float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas, const uint32_t perturb,
                                                         const int zero_fill, const uint32_t M_padded) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (zero_fill) {
        // padding rows beyond n_alive*n_step (align-to-128 tail, raymarching.py:386-387)
        const uint32_t row = n_alive * n_step + n;
        if (row < M_padded) {
            xyzs[(size_t)row * 3] = 0; xyzs[(size_t)row * 3 + 1] = 0; xyzs[(size_t)row * 3 + 2] = 0;
            dirs[(size_t)row * 3] = 0; dirs[(size_t)row * 3 + 1] = 0; dirs[(size_t)row * 3 + 2] = 0;
            deltas[(size_t)row * 2] = 0; deltas[(size_t)row * 2 + 1] = 0;
        }
    }
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    float t = rays_t[n];
    const Ray r = load_ray(rays_o + (size_t)index * 3, rays_d + (size_t)index * 3);
    const MarchParams p = make_march_params(bound, dt_gamma, max_steps, C, H, grid);
    float* px = xyzs + (size_t)n * n_step * 3;
    float* pd = dirs + (size_t)n * n_step * 3;
    float* pl = deltas + (size_t)n * n_step * 2;
    const float far = fars[index];
    if (perturb) {
        Pcg32 rng; rng.seed((uint64_t)perturb);   // :1011
        rng.advance(n);
        t += p.dt_min * rng.next_float();
    }
    float last_t = t, x, y, z, dt;
    uint32_t step = 0;
    while (t < far && step < n_step) {
        if (probe(r, p, t, x, y, z, dt)) {
            px[0] = x; px[1] = y; px[2] = z;
            pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
            t += dt;
            pl[0] = dt; pl[1] = t - last_t;
            last_t = t;
            px += 3; pd += 3; pl += 2; step++;
        }
    }
    if (zero_fill) {
        for (; step < n_step; step++) {
            px[0] = 0; px[1] = 0; px[2] = 0; pd[0] = 0; pd[1] = 0; pd[2] = 0; pl[0] = 0; pl[1] = 0;
            px += 3; pd += 3; pl += 2;
        }
    }
}

// Same marcher, but every thread parks its (at most 8) samples in shared memory first and the block then writes its three
// contiguous output segments with coalesced 16-byte stores.  (One thread per ray emits 32 B per sample at a stride of
// n_step*32 B between lanes: written directly that is ~29 scattered sectors per warp store and the LSU queue becomes the
// bottleneck — 106 us per loop iteration in the first ncu capture, profiles/r01_frame_kernels.md.)
constexpr int kMarchThreads = 128;
constexpr uint32_t kMarchMaxStagedSteps = 8;

__device__ __forceinline__ void block_copy_out(float* __restrict__ dst, const float* __restrict__ src, uint32_t nfloats) {
    // dst is 16-byte aligned for every block (segment starts at a multiple of 128 rows)
    const uint32_t nvec = nfloats >> 2;
    for (uint32_t i = threadIdx.x; i < nvec; i += blockDim.x) st_stream_u4(dst + 4 * i, *reinterpret_cast<const uint4*>(src + 4 * i));
    for (uint32_t i = (nvec << 2) + threadIdx.x; i < nfloats; i += blockDim.x) dst[i] = src[i];
}

// Device-driven frames (ntx_render_rays): n_alive / n_step / M_padded of an iteration are only known on the device.  The
// kernels of the inference loop then take them from a FrameState and the by-value arguments are just the launch bound.
struct FrameState { int n_alive, n_step, m_padded, step, n_live, pad0, pad1, pad2; };   // n_live: rows the marcher filled this iteration
// "Paused" sentinel of the device-driven loop: a ray that has walked `walk_budget` empty voxels in one launch without filling
// its n_step slots stops there and writes (delta, delta_t) = (0, -t) into its next slot instead of (0, 0).  composite_rays then
// keeps the ray alive with rays_t = t (exactly the t the marcher would have probed next), so the walk continues in the next
// iteration instead of making the whole launch wait for the few rays that graze the object (the launch time of a steady-state
// iteration was the latency of its longest walk).  The reference's marcher never writes a negative delta_t.

__global__ void __launch_bounds__(kMarchThreads, 6) march_rays_staged_kernel(
    uint32_t n_alive, uint32_t n_step, const int* __restrict__ rays_alive, const float* __restrict__ rays_t, const float* __restrict__ rays_o,
    const float* __restrict__ rays_d, const float bound, const float dt_gamma, const uint32_t max_steps, const uint32_t C, const uint32_t H,
    const uint8_t* __restrict__ grid, const float* __restrict__ nears, const float* __restrict__ fars, float* __restrict__ xyzs, float* __restrict__ dirs,
    float* __restrict__ deltas, const uint32_t perturb, uint32_t M_padded, const uint8_t* __restrict__ coarse, const FrameState* __restrict__ state,
    unsigned long long* __restrict__ sample_counter, uint32_t walk_budget, int* __restrict__ live_rows, int* __restrict__ live_counter) {
    extern __shared__ __align__(16) float stage[];          // [128*nc*3] xyz | [128*nc*3] dir | [128*nc*2] delta, nc = min(n_step, 8)
    if (state) { n_alive = (uint32_t)state->n_alive; n_step = (uint32_t)state->n_step; M_padded = (uint32_t)state->m_padded; }
    if (!state || state->step == 0 || walk_budget == 0) walk_budget = 0xffffffffu;   // the approach to the object (first iteration) is walked in one go
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    {   // padding rows beyond n_alive*n_step (align-to-128 tail, raymarching.py:386-387)
        const uint32_t row = n_alive * n_step + n;
        if (row < M_padded) {
            xyzs[(size_t)row * 3] = 0; xyzs[(size_t)row * 3 + 1] = 0; xyzs[(size_t)row * 3 + 2] = 0;
            dirs[(size_t)row * 3] = 0; dirs[(size_t)row * 3 + 1] = 0; dirs[(size_t)row * 3 + 2] = 0;
            deltas[(size_t)row * 2] = 0; deltas[(size_t)row * 2 + 1] = 0;
        }
    }
    const uint32_t first = blockIdx.x * kMarchThreads;
    if (first >= n_alive) return;                            // block only had padding rows to clear
    const bool mine = n < n_alive;
    const uint32_t rays_here = min((uint32_t)kMarchThreads, n_alive - first);
    Ray r;
    float t = 0.f, far = 0.f, last_t = 0.f;
    const MarchParams p = make_march_params(bound, dt_gamma, max_steps, C, H, grid);
    bool exhausted = !mine;                                  // nothing (more) to emit on this ray
    if (mine) {
        const int index = rays_alive[n];
        t = rays_t[n];
        r = load_ray(rays_o + (size_t)index * 3, rays_d + (size_t)index * 3);
        far = fars[index];
        if (perturb) {
            Pcg32 rng; rng.seed((uint64_t)perturb);   // raymarching.cu:1011
            rng.advance(n);
            t += p.dt_min * rng.next_float();
        }
        if (coarse) {
            // beyond the (grown) bounding box of everything occupied no sample can be emitted: shorten the walk, or skip it
            // altogether for rays that miss the box.  Samples in front of `far` are unaffected.
            float b_lo, b_hi;
            if (clip_to_occupied(r, p, coarse, b_lo, b_hi)) far = fminf(far, b_hi); else far = t;
        }
        last_t = t;
    }
    uint32_t emitted = 0, empties = 0;
    bool paused = false;
    // The ray's n_step output rows are produced in chunks of at most 8 (the reference never asks for more than 8; wider
    // schedules of ntx_render_rays do): a chunk is staged in shared memory and written out before the next one starts.
    for (uint32_t c0 = 0; c0 < n_step; c0 += kMarchMaxStagedSteps) {
        const uint32_t nc = min(kMarchMaxStagedSteps, n_step - c0);
        float* sx = stage;
        float* sd = stage + kMarchThreads * nc * 3;
        float* sl = sd + kMarchThreads * nc * 3;
        float* px = sx + threadIdx.x * nc * 3;
        float* pd = sd + threadIdx.x * nc * 3;
        float* pl = sl + threadIdx.x * nc * 2;
        uint32_t step = 0, filled = 0;                           // filled: real samples of this chunk (no sentinels)
        if (mine) {
            float x, y, z, dt;
            if (!exhausted) {
                // Speculation: inside the object every probe is occupied, so the next nc sample positions are t, t+dt, ...
                // (the same float additions the sequential marcher performs).  Issue all their occupancy loads at once instead
                // of one dependent load per step, then accept the longest all-occupied prefix; the first empty voxel falls back
                // to the sequential loop below at exactly the state the reference would be in.
                // The speculated samples are written into their slots right away (a slot that turns out not to be reached is
                // overwritten by the sequential loop or the zero fill), so accepting a prefix costs one bit test per sample.
                float tq[kMarchMaxStagedSteps];
                uint32_t bitidx[kMarchMaxStagedSteps], byte[kMarchMaxStagedSteps];
                float tcur = t;
                uint32_t nspec = 0;
#pragma unroll
                for (uint32_t s = 0; s < kMarchMaxStagedSteps; s++) {
                    tq[s] = tcur; bitidx[s] = 0; byte[s] = 0;
                    if (s < nc && tcur < far) {
                        const uint32_t index = locate(r, p, tcur, x, y, z, dt);
                        bitidx[s] = index & 7u;
                        byte[s] = p.grid[index >> 3];
                        px[3 * s] = x; px[3 * s + 1] = y; px[3 * s + 2] = z;
                        pd[3 * s] = r.dx; pd[3 * s + 1] = r.dy; pd[3 * s + 2] = r.dz;
                        const float t_end = tcur + dt;                 // == the sequential marcher's `t += dt`
                        pl[2 * s] = dt; pl[2 * s + 1] = t_end - (s == 0 ? last_t : tcur);   // last_t of sample s > 0 is the end of sample s-1 = tcur
                        tcur = t_end;
                        nspec = s + 1;
                    }
                }
                bool run = true;
#pragma unroll
                for (uint32_t s = 0; s < kMarchMaxStagedSteps; s++) {
                    if (s < nspec && run) {                            // run: all earlier speculated probes were occupied
                        if (byte[s] & (1u << bitidx[s])) { step = s + 1; t = (s + 1 < kMarchMaxStagedSteps) ? tq[s + 1 < kMarchMaxStagedSteps ? s + 1 : s] : tcur; }
                        else run = false;
                    }
                }
                if (step) { last_t = t; px += 3 * step; pd += 3 * step; pl += 2 * step; }
                while (t < far && step < nc) {
                    if (probe(r, p, t, x, y, z, dt)) {
                        px[0] = x; px[1] = y; px[2] = z;
                        pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
                        t += dt;
                        pl[0] = dt; pl[1] = t - last_t;
                        last_t = t;
                        px += 3; pd += 3; pl += 2; step++;
                    } else {
                        // (probe() has advanced t to the next voxel)
                        if (coarse && ((empties & 3u) == 0u)) {
                            // in empty space: if nothing occupied can be reached any more, this ray emits no further sample and is
                            // dead after composite_rays whatever t it stops at — skip the voxel-by-voxel walk to `far`
                            if (!maybe_occupied_ahead(r, p, coarse, t, far)) { far = t; break; }
                        }
                        if (++empties > walk_budget && t < far) { paused = true; break; }
                    }
                }
                if (step < nc) exhausted = true;                 // ran out of ray (or paused): every later slot is a sentinel
                emitted += step;
                filled = step;
                if (paused) {                                    // step < nc here: the pause happens on an empty probe
                    px[0] = 0; px[1] = 0; px[2] = 0; pd[0] = 0; pd[1] = 0; pd[2] = 0; pl[0] = 0; pl[1] = -t;
                    px += 3; pd += 3; pl += 2; step++;
                    paused = false;
                }
            }
            for (; step < nc; step++) {                          // unused slots: zero (delta == 0 is composite_rays' stop sentinel)
                px[0] = 0; px[1] = 0; px[2] = 0; pd[0] = 0; pd[1] = 0; pd[2] = 0; pl[0] = 0; pl[1] = 0;
                px += 3; pd += 3; pl += 2;
            }
        }
        __shared__ uint8_t s_rows[kMarchThreads];                // rows of this chunk somebody will read: filled (+1 for the sentinel)
        s_rows[threadIdx.x] = (uint8_t)min(filled + 1u, nc);
        if (live_rows) {
            // Row list for the field kernel: the rows this block filled, in ray order; blocks append in arrival order (the list is
            // consumed 128 rows at a time, so locality between consecutive samples of a ray and neighbouring rays is preserved).
            __shared__ uint32_t s_scan[33];
            __shared__ uint32_t s_base;
            uint32_t total;
            const uint32_t excl = block_exclusive_scan(filled, total, s_scan);
            if (threadIdx.x == 0) s_base = total ? (uint32_t)atomicAdd(live_counter, (int)total) : 0u;
            __syncthreads();
            const uint32_t row = n * n_step + c0;
            for (uint32_t j = 0; j < filled; j++) live_rows[s_base + excl + j] = (int)(row + j);
        }
        __syncthreads();
        if (nc == n_step) {
            // the block's rows form one contiguous segment of each output
            const size_t row0 = (size_t)first * n_step;
            block_copy_out(xyzs + row0 * 3, sx, rays_here * n_step * 3);
            block_copy_out(dirs + row0 * 3, sd, rays_here * n_step * 3);
            block_copy_out(deltas + row0 * 2, sl, rays_here * n_step * 2);
        } else {
            // one segment of nc rows per ray, n_step rows apart: a warp writes one ray's segments (<= 96 contiguous bytes each)
            // per round (an index/division-based mapping of this copy was 37 % of the kernel's instructions)
            // In a device-driven frame nothing reads past a ray's first sentinel (the field kernel runs over the live-row list,
            // composite_rays stops at the sentinel): those rows are not written at all — at 64 samples per ray and iteration the
            // zero fill was a quarter of the marcher's HBM writes.
            const uint32_t lane = threadIdx.x & 31u;
            for (uint32_t ray = threadIdx.x >> 5; ray < rays_here; ray += kMarchThreads / 32) {
                const size_t row = (size_t)(first + ray) * n_step + c0;
                const uint32_t keep = live_rows ? s_rows[ray] : nc;
                if (lane < keep * 3) { xyzs[row * 3 + lane] = sx[ray * nc * 3 + lane]; dirs[row * 3 + lane] = sd[ray * nc * 3 + lane]; }
                if (lane < keep * 2) deltas[row * 2 + lane] = sl[ray * nc * 2 + lane];
            }
        }
        __syncthreads();
    }
    if (sample_counter && mine) {                                // bench / statistics only: samples emitted in this launch
        if (__activemask() == 0xffffffffu) {                     // full warp: one atomic per warp
            uint32_t cnt = emitted;
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            if ((threadIdx.x & 31) == 0) atomicAdd(sample_counter, (unsigned long long)cnt);
        } else {
            atomicAdd(sample_counter, (unsigned long long)emitted);
        }
    }
}

// ---------------------------------------------------------------------------------------------------- compact marcher
// The marcher of the device-driven frame (ntx_render_rays).  Same walk as march_rays_staged_kernel — same probes in the same order,
// same float additions, so every sample has bit-identical (t, dt, delta_t) — but a sample leaves the kernel as 12 bytes instead of 32:
//   ts[row] = the sample's ray parameter t, deltas[row] = (dt, t_end - last_t)        row = ray_slot * n_step + k
// and the field kernel (MODE_RAYS) rebuilds xyz = clamp(o + t d) itself and reads the view direction from the ray.  With 3 floats
// per sample a thread keeps a whole speculation window of 8 samples in registers and writes it with six 16-byte stores (a ray's
// rows are contiguous and 16-byte aligned when n_step % 4 == 0), so the shared-memory staging, its bank conflicts, the block
// barriers around the copy-out (44 % of the staged kernel's stall samples) and the copy-out itself are gone.
// live[i] = (row, ray index) of every filled row, appended per warp (warp scan + one atomic per warp: no barrier in this kernel).
__global__ void __launch_bounds__(kMarchThreads, 6) march_rays_compact_kernel(
    const int* __restrict__ rays_alive, const float* __restrict__ rays_t, const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float bound,
    const float dt_gamma, const uint32_t max_steps, const uint32_t C, const uint32_t H, const uint8_t* __restrict__ grid, const float* __restrict__ fars,
    float* __restrict__ ts, float* __restrict__ deltas, const uint32_t perturb, const uint8_t* __restrict__ coarse, const FrameState* __restrict__ state,
    unsigned long long* __restrict__ sample_counter, uint32_t walk_budget, int2* __restrict__ live, int* __restrict__ live_counter) {
    const uint32_t n_alive = (uint32_t)state->n_alive, n_step = (uint32_t)state->n_step;
    if (state->step == 0 || walk_budget == 0) walk_budget = 0xffffffffu;   // the approach to the object (first iteration) is walked in one go
    if (blockIdx.x * kMarchThreads >= n_alive) return;
    const uint32_t n = threadIdx.x + blockIdx.x * kMarchThreads;
    const bool mine = n < n_alive;
    const MarchParams p = make_march_params(bound, dt_gamma, max_steps, C, H, grid);
    Ray r;
    float t = 0.f, far = 0.f, last_t = 0.f;
    int index = 0;
    if (mine) {
        index = rays_alive[n];
        t = rays_t[n];
        r = load_ray(rays_o + (size_t)index * 3, rays_d + (size_t)index * 3);
        far = fars[index];
        if (perturb) {
            Pcg32 rng; rng.seed((uint64_t)perturb);   // raymarching.cu:1011
            rng.advance(n);
            t += p.dt_min * rng.next_float();
        }
        if (coarse) {
            float b_lo, b_hi;
            if (clip_to_occupied(r, p, coarse, b_lo, b_hi)) far = fminf(far, b_hi); else far = t;
        }
        last_t = t;
    }
    const size_t row0 = (size_t)n * n_step;
    float* __restrict__ pl = deltas + row0 * 2;
    float* __restrict__ pt = ts + row0;
    const bool vec_ok = (n_step & 3u) == 0;                      // then row0 * 4 bytes is a multiple of 16
    uint32_t step = 0, empties = 0;                              // step: rows filled so far
    bool exhausted = !mine, paused = false;
    while (!exhausted && step < n_step) {
        // Speculation window (see march_rays_staged_kernel): the next nc lattice samples t, t+dt, ... with all their occupancy loads
        // in flight at once; the longest all-occupied prefix is accepted, the first empty voxel falls back to the sequential loop.
        const uint32_t nc = min(kMarchMaxStagedSteps, n_step - step);
        float tq[kMarchMaxStagedSteps], dq[kMarchMaxStagedSteps], lq[kMarchMaxStagedSteps];
        uint32_t bitidx[kMarchMaxStagedSteps], byte[kMarchMaxStagedSteps];
        float tcur = t, x, y, z, dt;
        uint32_t nspec = 0;
#pragma unroll
        for (uint32_t s = 0; s < kMarchMaxStagedSteps; s++) {
            tq[s] = tcur; dq[s] = 0.f; lq[s] = 0.f; bitidx[s] = 0; byte[s] = 0;
            if (s < nc && tcur < far) {
                const uint32_t vox = locate(r, p, tcur, x, y, z, dt);
                bitidx[s] = vox & 7u;
                byte[s] = p.grid[vox >> 3];
                const float t_end = tcur + dt;                   // == the sequential marcher's `t += dt`
                dq[s] = dt;
                lq[s] = t_end - (s == 0 ? last_t : tcur);        // last_t of sample s > 0 is the end of sample s-1 = tcur
                tcur = t_end;
                nspec = s + 1;
            }
        }
        uint32_t acc = 0;
        bool run = true;
#pragma unroll
        for (uint32_t s = 0; s < kMarchMaxStagedSteps; s++) {
            if (s < nspec && run) {
                if (byte[s] & (1u << bitidx[s])) acc = s + 1; else run = false;
            }
        }
        if (acc == kMarchMaxStagedSteps && vec_ok && (step & 3u) == 0) {
            // a full window: 8 samples = 64 bytes of deltas + 32 bytes of ts, all 16-byte aligned
            float4* d4 = reinterpret_cast<float4*>(pl + 2 * step);
            d4[0] = make_float4(dq[0], lq[0], dq[1], lq[1]); d4[1] = make_float4(dq[2], lq[2], dq[3], lq[3]);
            d4[2] = make_float4(dq[4], lq[4], dq[5], lq[5]); d4[3] = make_float4(dq[6], lq[6], dq[7], lq[7]);
            float4* t4 = reinterpret_cast<float4*>(pt + step);
            t4[0] = make_float4(tq[0], tq[1], tq[2], tq[3]); t4[1] = make_float4(tq[4], tq[5], tq[6], tq[7]);
        } else {
#pragma unroll
            for (uint32_t s = 0; s < kMarchMaxStagedSteps; s++)
                if (s < acc) { pl[2 * (step + s)] = dq[s]; pl[2 * (step + s) + 1] = lq[s]; pt[step + s] = tq[s]; }
        }
        if (acc) {
            // t after the accepted samples: the start of the next speculated one (tq[acc]), or the end of the window
            float tn = tcur;
#pragma unroll
            for (uint32_t s = 1; s < kMarchMaxStagedSteps; s++) if (s == acc) tn = tq[s];
            t = tn; last_t = tn;
        }
        const uint32_t window_end = step + nc;
        step += acc;
        while (t < far && step < window_end) {
            if (probe(r, p, t, x, y, z, dt)) {
                pt[step] = t;
                t += dt;
                pl[2 * step] = dt; pl[2 * step + 1] = t - last_t;
                last_t = t;
                step++;
            } else {
                // (probe() has advanced t to the next voxel)
                if (coarse && ((empties & 3u) == 0u)) {
                    if (!maybe_occupied_ahead(r, p, coarse, t, far)) { far = t; break; }
                }
                if (++empties > walk_budget && t < far) { paused = true; break; }
            }
        }
        if (step < window_end) exhausted = true;                 // ran out of ray (or paused): the next slot is the sentinel
    }
    if (mine && step < n_step) { pl[2 * step] = 0.f; pl[2 * step + 1] = paused ? -t : 0.f; }   // composite_rays stops at (0, .); (0, -t) = paused at t
    const uint32_t filled = mine ? step : 0u;
    {
        // Append this warp's rows to the live list: warp scan + ONE atomic per warp, no block barrier — a warp that has finished its
        // rays retires instead of waiting for the slowest ray of its block (a third of the staged design's and of the first compact
        // version's stall samples sat on that barrier).  Rows of a warp stay together and in ray order; warps append in arrival order.
        __syncwarp();
        const uint32_t lane = threadIdx.x & 31u;
        uint32_t inc = filled;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
        const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
        uint32_t base = 0;
        if (lane == 31 && total) base = (uint32_t)atomicAdd(live_counter, (int)total);
        base = __shfl_sync(0xffffffffu, base, 31);
        int2* __restrict__ dst = live + base + (inc - filled);
        for (uint32_t j = 0; j < filled; j++) dst[j] = make_int2((int)(row0 + j), index);
    }
    if (sample_counter && mine) {                                // bench / statistics only: samples emitted in this launch
        if (__activemask() == 0xffffffffu) {
            uint32_t cnt = filled;
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            if ((threadIdx.x & 31) == 0) atomicAdd(sample_counter, (unsigned long long)cnt);
        } else {
            atomicAdd(sample_counter, (unsigned long long)filled);
        }
    }
}

// raymarching.cu:1021-1104
__global__ void __launch_bounds__(128) composite_rays_kernel(uint32_t n_alive, uint32_t n_step, const int* __restrict__ rays_alive,
                                                             float* __restrict__ rays_t, const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                             const float* __restrict__ deltas, float* __restrict__ weights_sum, float* __restrict__ depth,
                                                             float* __restrict__ image, const FrameState* __restrict__ state) {
    if (state) { n_alive = (uint32_t)state->n_alive; n_step = (uint32_t)state->n_step; }
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    float t = rays_t[n];
    sigmas += (size_t)n * n_step; rgbs += (size_t)n * n_step * 3; deltas += (size_t)n * n_step * 2;
    float weight_sum = weights_sum[index], d = depth[index];
    float r = image[(size_t)index * 3], g = image[(size_t)index * 3 + 1], b = image[(size_t)index * 3 + 2];
    uint32_t step = 0;
    bool paused = false, stop = false;
    // one sample of the reference's loop body (raymarching.cu:1060-1090); returns false when the ray stops at this slot
    auto body = [&](const float sigma, const float dl0, const float dl1, const float cr, const float cg, const float cb) -> bool {
        if (dl0 == 0) {
            if (dl1 < 0) { t = -dl1; paused = true; }   // paused by the device-driven marcher: resume at exactly this t
            return false;
        }
        const float alpha = 1.0f - __expf(-sigma * dl0);
        const float T = 1 - weight_sum;
        const float weight = alpha * T;
        weight_sum += weight;
        t += dl1;
        d += weight * t;
        r += weight * cr; g += weight * cg; b += weight * cb;
        if (T < 1e-4) return false;   // double literal on purpose (:1081); the reference breaks BEFORE counting this step
        step++;
        return true;
    };
    if ((n_step & 3u) == 0 && ((reinterpret_cast<uintptr_t>(sigmas) | reinterpret_cast<uintptr_t>(rgbs) | reinterpret_cast<uintptr_t>(deltas)) & 15u) == 0) {
        // four samples per round through 128-bit loads (a ray's n_step rows are contiguous and, with n_step % 4 == 0, 16-byte
        // aligned): 6 load instructions per 4 samples instead of 24.  Rows behind a ray's first sentinel may be unwritten —
        // they are loaded but never looked at.
        const float4* s4 = reinterpret_cast<const float4*>(sigmas);
        const float4* c4 = reinterpret_cast<const float4*>(rgbs);
        const float4* d4 = reinterpret_cast<const float4*>(deltas);
        for (uint32_t q = 0; q < n_step / 4 && !stop; q++) {
            const float4 sg = s4[q], ca = c4[3 * q], cb4 = c4[3 * q + 1], cc = c4[3 * q + 2], da = d4[2 * q], db = d4[2 * q + 1];
            stop = !body(sg.x, da.x, da.y, ca.x, ca.y, ca.z) || !body(sg.y, da.z, da.w, ca.w, cb4.x, cb4.y) ||
                   !body(sg.z, db.x, db.y, cb4.z, cb4.w, cc.x) || !body(sg.w, db.z, db.w, cc.y, cc.z, cc.w);
        }
    } else {
        while (step < n_step && body(sigmas[step], deltas[2 * step], deltas[2 * step + 1], rgbs[3 * step], rgbs[3 * step + 1], rgbs[3 * step + 2])) {}
    }
    rays_t[n] = (step < n_step && !paused) ? -1.0f : t;
    weights_sum[index] = weight_sum; depth[index] = d;
    image[(size_t)index * 3] = r; image[(size_t)index * 3 + 1] = g; image[(size_t)index * 3 + 2] = b;
}

// samples per ray of an iteration of the device-driven frame: the reference's clamp(budget // n_alive, 1, cap) (renderer.py:464 with
// budget = N, cap = 8).  Wide schedules (cap > 8) round down to a multiple of 4 so that a ray's rows start 16-byte aligned in
// ts / deltas (vector stores in the marcher, vector loads in composite_rays); the image does not depend on n_step.
__device__ __forceinline__ uint32_t schedule_n_step(uint32_t budget, uint32_t n_alive, uint32_t cap) {
    uint32_t n = max(min(budget / n_alive, cap), 1u);
    if (cap > 8u && n >= 4u) n &= ~3u;
    return n;
}

// ordered stream compaction (raymarching.cu:1117-1134, atomics replaced by scans)
constexpr int kCompactThreads = 256;
constexpr int kCompactItems = 4;

// Device-driven mode (state_in != nullptr): the number of rays to compact comes from state_in, the result goes to state_out
// together with the next iteration's n_step / padded sample count / step counter (the bookkeeping of renderer.py:456-475),
// and to a pinned host mailbox the launching thread reads a few iterations later.
__global__ void __launch_bounds__(kCompactThreads) compact_rays_kernel(uint32_t n_alive, int* __restrict__ rays_alive, const int* __restrict__ rays_alive_old,
                                                                       float* __restrict__ rays_t, const float* __restrict__ rays_t_old, int* __restrict__ alive_counter,
                                                                       ScanWS* ws, const FrameState* __restrict__ state_in, FrameState* __restrict__ state_out,
                                                                       const uint32_t budget, const uint32_t max_n_step, const uint32_t max_steps,
                                                                       volatile int* host_mailbox) {
    if (state_in) n_alive = (uint32_t)state_in->n_alive;
    __shared__ uint32_t s_ticket;
    __shared__ uint32_t s_warp[33];
    __shared__ unsigned long long s_bcast;
    if (threadIdx.x == 0) s_ticket = atomicAdd(&ws->ticket, 1u);
    __syncthreads();
    const uint32_t ticket = s_ticket;
    const uint32_t first = (ticket * kCompactThreads + threadIdx.x) * kCompactItems;  // each thread owns 4 consecutive slots
    int id[kCompactItems];
    float tv[kCompactItems];
    uint32_t cnt = 0;
#pragma unroll
    for (int i = 0; i < kCompactItems; i++) {
        const uint32_t n = first + i;
        tv[i] = -1.0f; id[i] = 0;
        if (n < n_alive) { tv[i] = rays_t_old[n]; id[i] = rays_alive_old[n]; }
        cnt += (tv[i] >= 0) ? 1u : 0u;
    }
    uint32_t block_total;
    const uint32_t local = block_exclusive_scan(cnt, block_total, s_warp);
    const unsigned long long base = (ticket == 0 && !state_in) ? (unsigned long long)(uint32_t)alive_counter[0] : 0ull;
    const unsigned long long excl = chained_prefix(ws, ticket, block_total, base, &s_bcast);
    const uint32_t nblocks = gridDim.x;
    if (ticket == nblocks - 1 && threadIdx.x == 0) {
        const uint32_t total = (uint32_t)(excl + block_total);
        if (state_in) {
            FrameState s;
            s.step = state_in->step + state_in->n_step;                       // step += n_step   (renderer.py:475)
            s.n_alive = ((uint32_t)s.step < max_steps) ? (int)total : 0;       // while step < max_steps
            s.n_step = s.n_alive ? (int)schedule_n_step(budget, (uint32_t)s.n_alive, max_n_step) : 0;
            const uint32_t m = (uint32_t)s.n_alive * (uint32_t)s.n_step;
            s.m_padded = s.n_alive ? (int)(m + 128u - (m % 128u)) : 0;         // raymarching.py:386-387 (align = 128)
            s.n_live = 0; s.pad0 = s.pad1 = s.pad2 = 0;
            *state_out = s;
            if (host_mailbox) { *host_mailbox = s.n_alive; __threadfence_system(); }
        } else {
            alive_counter[0] = (int)total;
        }
    }
    uint32_t dst = (uint32_t)excl + local;
#pragma unroll
    for (int i = 0; i < kCompactItems; i++) {
        if (tv[i] >= 0) { rays_alive[dst] = id[i]; rays_t[dst] = tv[i]; dst++; }
    }
    scan_ws_release(ws, nblocks);
}

}  // namespace ntx

using namespace ntx;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int ntx_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near, float* nears, float* fars,
                                      ntx_stream_t stream) {
    NTX_REQUIRE(rays_o && rays_d && aabb && nears && fars, NTX_ERR_INVALID_ARGUMENT, "near_far_from_aabb: null pointer");
    if (N == 0) return NTX_OK;
    near_far_kernel<<<ceil_div<uint32_t>(N, 128), 128, 0, ST(stream)>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    return check_launch("near_far_from_aabb");
}

extern "C" int ntx_polar_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords, ntx_stream_t stream) {
    NTX_REQUIRE(rays_o && rays_d && coords, NTX_ERR_INVALID_ARGUMENT, "polar_from_ray: null pointer");
    if (N == 0) return NTX_OK;
    polar_kernel<<<ceil_div<uint32_t>(N, 128), 128, 0, ST(stream)>>>(rays_o, rays_d, radius, N, coords);
    return check_launch("polar_from_ray");
}

extern "C" int ntx_morton3D(const int* coords, uint32_t N, int* indices, ntx_stream_t stream) {
    NTX_REQUIRE(coords && indices, NTX_ERR_INVALID_ARGUMENT, "morton3D: null pointer");
    if (N == 0) return NTX_OK;
    morton3D_kernel<<<ceil_div<uint32_t>(N, 128), 128, 0, ST(stream)>>>(coords, N, indices);
    return check_launch("morton3D");
}

extern "C" int ntx_morton3D_invert(const int* indices, uint32_t N, int* coords, ntx_stream_t stream) {
    NTX_REQUIRE(coords && indices, NTX_ERR_INVALID_ARGUMENT, "morton3D_invert: null pointer");
    if (N == 0) return NTX_OK;
    morton3D_invert_kernel<<<ceil_div<uint32_t>(N, 128), 128, 0, ST(stream)>>>(indices, N, coords);
    return check_launch("morton3D_invert");
}

static int launch_packbits(const float* grid, uint32_t N, float density_thresh, const float* thresh_dev, uint8_t* bitfield, cudaStream_t st) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(grid) & 15) == 0) && ((reinterpret_cast<uintptr_t>(bitfield) & 3) == 0);
    const uint32_t vec_bytes = aligned ? (N / 4) * 4 : 0;
    if (vec_bytes) packbits_kernel<<<min(ceil_div<uint32_t>(vec_bytes / 4, 256), (uint32_t)device_sm_count() * 8u), 256, 0, st>>>(grid, vec_bytes, density_thresh, bitfield, thresh_dev);
    if (vec_bytes < N) packbits_scalar_kernel<<<ceil_div<uint32_t>(N - vec_bytes, 128), 128, 0, st>>>(grid, vec_bytes, N, density_thresh, bitfield, thresh_dev);
    return check_launch("packbits");
}

extern "C" int ntx_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield, ntx_stream_t stream) {
    NTX_REQUIRE(grid && bitfield, NTX_ERR_INVALID_ARGUMENT, "packbits: null pointer");
    if (N == 0) return NTX_OK;
    return launch_packbits(grid, N, density_thresh, nullptr, bitfield, ST(stream));
}

// ---------------------------------------------------------------------------------------------------- density-grid maintenance
// NeRFRenderer.update_extra_state (nerf/renderer.py:567-660) as one launch chain on the stream, no host round trip:
//   tmp = -1  ->  per cascade: fused field kernel in density mode (positions from Morton cell indices, hash-grid gather + sigma MLP,
//   sigma * density_scale -> tmp[cell])  ->  EMA-max into density_grid + sum of clamp(grid, 0)  ->  threshold = min(mean, density_thresh)
//   ->  packbits with that device-side threshold.
namespace ntx {
__global__ void __launch_bounds__(256) fill_f32_kernel(float* __restrict__ p, const uint32_t n, const float v) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}
// density_grid[valid] = max(density_grid[valid] * decay, tmp[valid]), valid = (density_grid >= 0) & (tmp >= 0)  (renderer.py:637-641);
// stats[0] += sum(clamp(density_grid, min=0)) in double
__global__ void __launch_bounds__(256) density_ema_kernel(float* __restrict__ grid, const float* __restrict__ tmp, const uint32_t n, const float decay,
                                                          const bool force_full_grid, double* __restrict__ stats) {
    __shared__ double s_sum[8];
    double local = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float g = grid[i];
        const float t = tmp[i];
        if (force_full_grid || (g >= 0 && t >= 0)) { g = fmaxf(__fmul_rn(g, decay), t); grid[i] = g; }
        local += (double)fmaxf(g, 0.0f);
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += s_sum[w];
        atomicAdd(stats, t);
    }
}
// stats_out = {mean_density, threshold = min(mean_density, density_thresh)}   (renderer.py:642-647)
__global__ void density_finalize_kernel(const double* __restrict__ stats, const uint32_t n, const float density_thresh, float* __restrict__ out2) {
    const float mean = (float)(stats[0] / (double)n);
    out2[0] = mean;
    out2[1] = fminf(mean, density_thresh);
}
}  // namespace ntx

extern "C" size_t ntx_update_density_grid_workspace_bytes(uint32_t C, uint32_t H) { return sizeof(float) * (size_t)C * H * H * H + 256; }

extern "C" int ntx_update_density_grid(float* density_grid, uint8_t* density_bitfield, uint32_t C, uint32_t H, float bound, float density_scale, float decay,
                                       float density_thresh, const void* embeddings_f16, const int* offsets, uint32_t L, float S, uint32_t base_resolution,
                                       int align_corners, const void* w_sigma_f16, const int* cells, uint32_t n_cells, const float* noise, int force_full_grid,
                                       void* workspace, float* stats_out, ntx_stream_t stream) {
    NTX_REQUIRE(density_grid && density_bitfield && workspace && stats_out, NTX_ERR_INVALID_ARGUMENT, "update_density_grid: null pointer");
    NTX_REQUIRE(C >= 1 && C <= 16 && H >= 2 && H <= 1024 && (H & (H - 1)) == 0, NTX_ERR_INVALID_ARGUMENT, "update_density_grid: C in [1,16], H a power of two <= 1024");
    NTX_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, NTX_ERR_INVALID_ARGUMENT, "update_density_grid: workspace must be 256-byte aligned");
    cudaStream_t st = ST(stream);
    const uint32_t H3 = H * H * H, total = C * H3;
    double* stats = static_cast<double*>(workspace);                       // [1] sum of clamp(grid, 0)
    float* tmp = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
    cudaMemsetAsync(stats, 0, sizeof(double), st);
    fill_f32_kernel<<<min(ceil_div<uint32_t>(total, 256), (uint32_t)device_sm_count() * 8u), 256, 0, st>>>(tmp, total, -1.0f);   // tmp_grid = -ones_like (renderer.py:575)
    const uint32_t per = cells ? n_cells : H3;
    for (uint32_t cas = 0; cas < C; cas++) {
        const float cb = fminf((float)(1u << cas), bound);                // bound = min(2 ** cas, self.bound)   (renderer.py:592)
        const int rc = launch_density_query(per, cells ? cells + (size_t)cas * n_cells : nullptr, noise ? noise + (size_t)cas * per * 3 : nullptr, H, cb, bound,
                                            embeddings_f16, offsets, L, S, base_resolution, align_corners, w_sigma_f16, density_scale, tmp + (size_t)cas * H3, st);
        if (rc != NTX_OK) return rc;
    }
    density_ema_kernel<<<min(ceil_div<uint32_t>(total, 256), (uint32_t)device_sm_count() * 8u), 256, 0, st>>>(density_grid, tmp, total, decay, force_full_grid != 0, stats);
    density_finalize_kernel<<<1, 1, 0, st>>>(stats, total, density_thresh, stats_out);
    return launch_packbits(density_grid, total / 8, 0.0f, stats_out + 1, density_bitfield, st);
}

static size_t scan_ws_bytes(uint32_t nblocks) { return sizeof(uint32_t) * 2 + sizeof(unsigned long long) * (size_t)(nblocks ? nblocks : 1); }

extern "C" size_t ntx_march_rays_train_workspace_bytes(uint32_t N) { return scan_ws_bytes(ceil_div<uint32_t>(N, kTrainThreads)); }

extern "C" int ntx_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma, uint32_t max_steps, uint32_t N,
                                    uint32_t C, uint32_t H, uint32_t M, const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                                    float* rays_ts, int* rays, int* counter, uint32_t perturb, void* workspace, ntx_stream_t stream) {
    NTX_REQUIRE(rays_o && rays_d && grid && nears && fars && xyzs && dirs && deltas && rays && counter, NTX_ERR_INVALID_ARGUMENT, "march_rays_train: null pointer");
    NTX_REQUIRE(workspace, NTX_ERR_WORKSPACE, "march_rays_train: workspace of ntx_march_rays_train_workspace_bytes(N) bytes required");
    NTX_REQUIRE(C >= 1 && C <= 16 && H >= 1 && H <= 1024 && max_steps >= 1, NTX_ERR_INVALID_ARGUMENT, "march_rays_train: bad C/H/max_steps");
    if (N == 0) return NTX_OK;
    march_rays_train_kernel<<<ceil_div<uint32_t>(N, kTrainThreads), kTrainThreads, 0, ST(stream)>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars,
                                                                                                     xyzs, dirs, deltas, rays_ts, rays, counter, perturb,
                                                                                                     static_cast<ScanWS*>(workspace));
    return check_launch("march_rays_train");
}

extern "C" int ntx_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int* rays, uint32_t M, uint32_t N,
                                                float* weights_sum, float* depth, float* image, ntx_stream_t stream) {
    NTX_REQUIRE(sigmas && rgbs && deltas && rays && weights_sum && depth && image, NTX_ERR_INVALID_ARGUMENT, "composite_rays_train_forward: null pointer");
    if (N == 0) return NTX_OK;
    composite_train_fwd_kernel<<<ceil_div<uint32_t>(N, 128), 128, 0, ST(stream)>>>(sigmas, rgbs, deltas, rays, M, N, weights_sum, depth, image);
    return check_launch("composite_rays_train_forward");
}

extern "C" int ntx_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas, const float* rgbs, const float* deltas,
                                                 const int* rays, const float* weights_sum, const float* image, uint32_t M, uint32_t N, float* grad_sigmas,
                                                 float* grad_rgbs, ntx_stream_t stream) {
    NTX_REQUIRE(grad_weights_sum && grad_image && sigmas && rgbs && deltas && rays && weights_sum && image && grad_sigmas && grad_rgbs, NTX_ERR_INVALID_ARGUMENT,
                "composite_rays_train_backward: null pointer");
    if (N == 0) return NTX_OK;
    composite_train_bwd_kernel<<<ceil_div<uint32_t>(N, 128), 128, 0, ST(stream)>>>(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N,
                                                                                   grad_sigmas, grad_rgbs);
    return check_launch("composite_rays_train_backward");
}

extern "C" int ntx_march_rays(uint32_t n_alive, uint32_t n_step, const int* rays_alive, const float* rays_t, const float* rays_o, const float* rays_d, float bound,
                              float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars, float* xyzs,
                              float* dirs, float* deltas, uint32_t perturb, int zero_fill, uint32_t M_padded, const uint8_t* occupancy_mip, ntx_stream_t stream) {
    NTX_REQUIRE(rays_alive && rays_t && rays_o && rays_d && grid && nears && fars && xyzs && dirs && deltas, NTX_ERR_INVALID_ARGUMENT, "march_rays: null pointer");
    if (occupancy_mip && ((H & (H - 1)) != 0 || H < 8)) occupancy_mip = nullptr;   // the mip needs a power-of-two grid
    NTX_REQUIRE(C >= 1 && C <= 16 && H >= 1 && H <= 1024 && max_steps >= 1, NTX_ERR_INVALID_ARGUMENT, "march_rays: bad C/H/max_steps");
    NTX_REQUIRE(!zero_fill || M_padded >= n_alive * n_step, NTX_ERR_INVALID_ARGUMENT, "march_rays: M_padded smaller than n_alive*n_step");
    uint32_t threads = n_alive;
    if (zero_fill) threads = max(threads, M_padded - n_alive * n_step);
    if (threads == 0) return NTX_OK;
    if (zero_fill && n_step >= 1 && ((reinterpret_cast<uintptr_t>(xyzs) | reinterpret_cast<uintptr_t>(dirs) | reinterpret_cast<uintptr_t>(deltas)) & 15) == 0) {
        const size_t smem = (size_t)kMarchThreads * min(n_step, kMarchMaxStagedSteps) * 8 * sizeof(float);
        march_rays_staged_kernel<<<ceil_div<uint32_t>(threads, kMarchThreads), kMarchThreads, smem, ST(stream)>>>(
            n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears, fars, xyzs, dirs, deltas, perturb, M_padded,
            occupancy_mip, nullptr, nullptr, 0, nullptr, nullptr);
        return check_launch("march_rays");
    }
    march_rays_kernel<<<ceil_div<uint32_t>(threads, 128), 128, 0, ST(stream)>>>(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid,
                                                                               nears, fars, xyzs, dirs, deltas, perturb, zero_fill, M_padded);
    return check_launch("march_rays");
}

extern "C" size_t ntx_occupancy_mip_bytes(uint32_t C, uint32_t H) { return kMipHeaderBytes + 2 * (size_t)C * (H >> kMipShift) * (H >> kMipShift) * (H >> kMipShift) / 8; }

extern "C" int ntx_build_occupancy_mip(const uint8_t* grid, uint32_t C, uint32_t H, uint8_t* mip, ntx_stream_t stream) {
    NTX_REQUIRE(grid && mip, NTX_ERR_INVALID_ARGUMENT, "build_occupancy_mip: null pointer");
    NTX_REQUIRE(C >= 1 && C <= 16 && H >= 16 && H <= 1024 && (H & (H - 1)) == 0, NTX_ERR_UNSUPPORTED, "build_occupancy_mip: H must be a power of two in [16, 1024]");
    NTX_REQUIRE((reinterpret_cast<uintptr_t>(grid) & 15) == 0, NTX_ERR_INVALID_ARGUMENT, "build_occupancy_mip: bit-field must be 16-byte aligned");
    const uint32_t Hc = H >> kMipShift, cells = C * Hc * Hc * Hc;
    uint8_t* dil = mip + kMipHeaderBytes;
    uint8_t* raw = dil + cells / 8;
    occupancy_mip_init_kernel<<<1, 128, 0, ST(stream)>>>(reinterpret_cast<int*>(mip), C);
    occupancy_mip_raw_kernel<<<ceil_div<uint32_t>(cells / 8, 128), 128, 0, ST(stream)>>>(grid, cells, Hc * Hc * Hc, raw, reinterpret_cast<int*>(mip));
    occupancy_mip_dilate_kernel<<<ceil_div<uint32_t>(cells / 8, 128), 128, 0, ST(stream)>>>(raw, C, Hc, dil);
    return check_launch("build_occupancy_mip");
}

extern "C" int ntx_composite_rays(uint32_t n_alive, uint32_t n_step, const int* rays_alive, float* rays_t, const float* sigmas, const float* rgbs, const float* deltas,
                                  float* weights_sum, float* depth, float* image, ntx_stream_t stream) {
    NTX_REQUIRE(rays_alive && rays_t && sigmas && rgbs && deltas && weights_sum && depth && image, NTX_ERR_INVALID_ARGUMENT, "composite_rays: null pointer");
    if (n_alive == 0) return NTX_OK;
    composite_rays_kernel<<<ceil_div<uint32_t>(n_alive, 128), 128, 0, ST(stream)>>>(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image,
                                                                                    nullptr);
    return check_launch("composite_rays");
}

extern "C" size_t ntx_compact_rays_workspace_bytes(uint32_t n_alive) { return scan_ws_bytes(ceil_div<uint32_t>(n_alive, kCompactThreads * kCompactItems)); }

extern "C" int ntx_compact_rays(uint32_t n_alive, int* rays_alive, const int* rays_alive_old, float* rays_t, const float* rays_t_old, int* alive_counter,
                                void* workspace, ntx_stream_t stream) {
    NTX_REQUIRE(rays_alive && rays_alive_old && rays_t && rays_t_old && alive_counter, NTX_ERR_INVALID_ARGUMENT, "compact_rays: null pointer");
    NTX_REQUIRE(workspace, NTX_ERR_WORKSPACE, "compact_rays: workspace of ntx_compact_rays_workspace_bytes(n_alive) bytes required");
    if (n_alive == 0) return NTX_OK;
    compact_rays_kernel<<<ceil_div<uint32_t>(n_alive, kCompactThreads * kCompactItems), kCompactThreads, 0, ST(stream)>>>(n_alive, rays_alive, rays_alive_old, rays_t,
                                                                                                                         rays_t_old, alive_counter,
                                                                                                                         static_cast<ScanWS*>(workspace), nullptr, nullptr,
                                                                                                                         0, 0, 0, nullptr);
    return check_launch("compact_rays");
}

// ---------------------------------------------------------------------------------------------------- device-driven frame
namespace ntx {
// rays_alive = 0..N-1, rays_t = nears (renderer.py:449-451).  With `prekill` (schedules other than the reference's) the occupancy
// mip is consulted ONCE per ray here instead of inside the marcher's walk: a ray that cannot reach anything occupied starts dead
// (rays_t = -1: it could not emit a sample; the reference finds that out by marching it through the cube and drops it in its first
// composite_rays with nothing accumulated — same image), and the others get `far` shortened to the end of their last maybe-occupied
// stretch, beyond which the marcher would only cross empty voxels and die.  The marcher then runs without any mip test in its loop:
// with 32 rays per warp whose tests fell due at different iterations, almost every iteration of every warp used to pay for a scan
// (first launch 485 us).  The compaction that follows is the ordinary compact_rays_kernel.
__global__ void __launch_bounds__(256) frame_init_kernel(const uint32_t N, const uint32_t budget, const uint32_t max_n_step, const float* __restrict__ nears,
                                                         float* __restrict__ fars, int* __restrict__ rays_alive, float* __restrict__ rays_t,
                                                         FrameState* __restrict__ state, volatile int* host_mailbox, const bool prekill, const float* __restrict__ rays_o,
                                                         const float* __restrict__ rays_d, const float bound, const float dt_gamma, const uint32_t max_steps,
                                                         const uint32_t C, const uint32_t H, const uint8_t* __restrict__ coarse) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) {
        float t = nears[n];
        if (prekill) {
            const MarchParams p = make_march_params(bound, dt_gamma, max_steps, C, H, nullptr);
            const Ray r = load_ray(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
            float b_lo, b_hi;
            if (!clip_to_occupied(r, p, coarse, b_lo, b_hi)) t = -1.0f;
            else {
                // the end of the ray's useful part, once: beyond it the marcher could only walk empty voxels and die
                const float far0 = fminf(fars[n], b_hi);
                const float t_end = last_maybe_occupied(r, p, coarse, fmaxf(t, b_lo), far0);
                if (t_end < 0.0f) t = -1.0f; else fars[n] = fminf(far0, t_end);
            }
        }
        rays_alive[n] = (int)n; rays_t[n] = t;
    }
    if (n == 0) {
        FrameState s;
        s.step = 0; s.n_alive = (int)N;
        s.n_step = prekill ? 0 : (int)schedule_n_step(budget, N, max_n_step);   // prekill: a compaction follows and plans iteration 0 (step += 0)
        const uint32_t m = N * (uint32_t)s.n_step;
        s.m_padded = (int)(m + 128u - (m % 128u));
        s.n_live = 0; s.pad0 = s.pad1 = s.pad2 = 0;
        *state = s;
        if (host_mailbox && !prekill) { *host_mailbox = (int)N; __threadfence_system(); }
    }
}

struct FrameWorkspace {
    float *nears, *fars, *rays_t[2], *ts, *deltas, *sigmas, *rgbs;
    int* rays_alive[2];
    int2* live;            // (row, ray) of every row the marcher filled this iteration
    FrameState* state;     // [2]
    ScanWS* scan;
    size_t bytes;
};
static FrameWorkspace carve_frame_workspace(void* base, uint32_t N, uint32_t budget) {
    FrameWorkspace w;
    size_t off = 0;
    auto take = [&](size_t bytes) { void* p = base ? static_cast<char*>(base) + off : nullptr; off += (bytes + 255) & ~size_t(255); return p; };
    const size_t Mmax = (size_t)max(N, budget) + 128;        // n_alive * n_step <= max(budget, n_alive)
    w.state = static_cast<FrameState*>(take(2 * sizeof(FrameState)));
    w.scan = static_cast<ScanWS*>(take(scan_ws_bytes(ceil_div<uint32_t>(N, kCompactThreads * kCompactItems))));
    w.nears = static_cast<float*>(take(sizeof(float) * N));
    w.fars = static_cast<float*>(take(sizeof(float) * N));
    for (int i = 0; i < 2; i++) { w.rays_alive[i] = static_cast<int*>(take(sizeof(int) * N)); w.rays_t[i] = static_cast<float*>(take(sizeof(float) * N)); }
    w.ts = static_cast<float*>(take(sizeof(float) * Mmax));
    w.deltas = static_cast<float*>(take(sizeof(float) * 2 * Mmax));
    w.sigmas = static_cast<float*>(take(sizeof(float) * Mmax));
    w.rgbs = static_cast<float*>(take(sizeof(float) * 3 * Mmax));
    w.live = static_cast<int2*>(take(sizeof(int2) * Mmax));
    w.bytes = off;
    return w;
}
}  // namespace ntx

extern "C" size_t ntx_render_rays_workspace_bytes(uint32_t N, uint32_t sample_budget) {
    return carve_frame_workspace(nullptr, N, sample_budget ? sample_budget : N).bytes;
}

#include <mutex>

extern "C" int ntx_render_rays(const float* rays_o, const float* rays_d, uint32_t N, const float* aabb, float min_near, float bound, float dt_gamma,
                               uint32_t max_steps, uint32_t perturb, uint32_t sample_budget, uint32_t max_n_step, uint32_t walk_budget, uint32_t C, uint32_t H,
                               const uint8_t* grid,
                               const uint8_t* occupancy_mip,
                               const void* embeddings_f16, const int* offsets, uint32_t L, float S, uint32_t base_resolution, int align_corners,
                               const void* w_sigma_f16, const void* w_color_f16, float density_scale, float* weights_sum, float* depth, float* image,
                               void* workspace, int* host_mailbox, unsigned long long* sample_counter, uint32_t* stats_out, float* kernel_ms_out,
                               ntx_stream_t stream) {
    NTX_REQUIRE(rays_o && rays_d && aabb && grid && weights_sum && depth && image, NTX_ERR_INVALID_ARGUMENT, "render_rays: null pointer");
    NTX_REQUIRE(workspace, NTX_ERR_WORKSPACE, "render_rays: workspace of ntx_render_rays_workspace_bytes(N) bytes required");
    NTX_REQUIRE(host_mailbox, NTX_ERR_INVALID_ARGUMENT, "render_rays: host_mailbox must point to max_steps + 2*H*C + 8 ints of pinned (mapped) host memory");
    NTX_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, NTX_ERR_INVALID_ARGUMENT, "render_rays: workspace must be 256-byte aligned");
    NTX_REQUIRE(C >= 1 && C <= 16 && H >= 1 && H <= 1024 && max_steps >= 1, NTX_ERR_INVALID_ARGUMENT, "render_rays: bad C/H/max_steps");
    if (occupancy_mip && ((H & (H - 1)) != 0 || H < 16)) occupancy_mip = nullptr;
    if (stats_out) stats_out[0] = stats_out[1] = 0;
    if (N == 0) return NTX_OK;
    if (sample_budget == 0) sample_budget = N;     // the reference's schedule: n_step = clamp(N // n_alive, 1, 8)   (renderer.py:464)
    if (max_n_step == 0) max_n_step = 8;
    NTX_REQUIRE(max_n_step <= 1024 && (uint64_t)max(N, sample_budget) + 128 < (1ull << 31), NTX_ERR_INVALID_ARGUMENT, "render_rays: bad sample_budget / max_n_step");
    cudaStream_t st = ST(stream);
    const FrameWorkspace w = carve_frame_workspace(workspace, N, sample_budget);
    // One frame at a time per device: the run-ahead events (and the optional profiling events) below are per-device state, so
    // concurrent calls from several threads on one device take turns here; calls on different devices do not interact.
    static std::mutex frame_mutex[kMaxDevices];
    std::lock_guard<std::mutex> frame_lock(frame_mutex[current_device()]);
    constexpr int kEvents = 8;
    static cudaEvent_t ev_dev[kMaxDevices][kEvents];
    static bool ev_ready_dev[kMaxDevices] = {};
    cudaEvent_t* ev = ev_dev[current_device()];
    bool& ev_ready = ev_ready_dev[current_device()];
    if (!ev_ready) {
        for (int i = 0; i < kEvents; i++)
            if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); set_error("render_rays: cannot create events"); return NTX_ERR_CUDA; }
        ev_ready = true;
    }
    // state and scan workspace are adjacent (carve_frame_workspace): one memset; the three result arrays are one when they form the planar
    // block [weights_sum | depth | rgb] of render.py / ntx_unshard_frame
    cudaMemsetAsync(w.state, 0, (size_t)(reinterpret_cast<char*>(w.scan) - reinterpret_cast<char*>(w.state)) + scan_ws_bytes(ceil_div<uint32_t>(N, kCompactThreads * kCompactItems)), st);
    if (depth == weights_sum + N && image == depth + N) {
        cudaMemsetAsync(weights_sum, 0, sizeof(float) * 5 * N, st);
    } else {
        cudaMemsetAsync(weights_sum, 0, sizeof(float) * N, st);
        cudaMemsetAsync(depth, 0, sizeof(float) * N, st);
        cudaMemsetAsync(image, 0, sizeof(float) * 3 * N, st);
    }
    near_far_kernel<<<ceil_div<uint32_t>(N, 128), 128, 0, st>>>(rays_o, rays_d, aabb, N, min_near, w.nears, w.fars);
    // Loop bounds.  Without pauses every iteration gives each living ray n_step samples, and `step` (their sum) reaching max_steps
    // ends the loop like the reference's `while step < max_steps`.  A paused ray spends iterations without sampling — at most
    // (voxels on a ray) / walk_budget < 2*H*C / walk_budget of them — so the bounds grow by that much; rays still end at t >= far.
    // (With pauses a ray that never pauses may therefore take up to step_limit > max_steps samples where the reference truncates at
    // max_steps: only reachable when a ray crosses more than max_steps occupied lattice points, i.e. not with dt_gamma = 0 and
    // max_steps >= 2*sqrt(3)*bound/dt_min; use walk_budget = 0 where that cap matters.)
    const uint32_t pause_iters = walk_budget ? 2u * H * C / walk_budget + 2u : 0u;
    const uint32_t step_limit = max_steps + pause_iters * max_n_step, iter_limit = max_steps + pause_iters;
    // rays that cannot hit anything occupied start dead (see frame_init_kernel); not for the reference's exact iteration structure
    // (perturb != 0: the jitter of a ray is drawn from its slot number in the first march, raymarching.cu:1011 — slots must stay ray ids)
    const bool prekill = walk_budget != 0 && occupancy_mip != nullptr && perturb == 0;
    // How many iterations the launching thread runs ahead of the one whose n_alive it has read.  1: grids shrink as soon as possible and
    // at most two empty iterations are queued at the end.  Queueing 3 ahead was measured on an 8-GPU box to check whether the host's
    // wake-up after each event limits a 1/8 shard: 1.45 -> 1.85 ms per frame — it does not; the extra empty launches and full-size
    // grids cost more (NTX_FRAME_AHEAD overrides for experiments).
    const uint32_t ahead = tunables().frame_ahead > 0 ? (uint32_t)std::min(tunables().frame_ahead, kEvents - 1) : 1u;
    uint32_t bound_rays = N, iterations = 0, kernels = 1;   // near_far
    // optional per-kernel timing (bench.py's roofline): CUDA events around every march and field launch of this frame
    constexpr uint32_t kProfIters = 256;
    static cudaEvent_t* prof_dev[kMaxDevices] = {};
    cudaEvent_t* prof = nullptr;
    uint32_t prof_iters = 0;
    if (kernel_ms_out) {
        kernel_ms_out[0] = kernel_ms_out[1] = 0.f;
        cudaEvent_t*& pool = prof_dev[current_device()];
        if (!pool) {
            pool = new cudaEvent_t[3 * kProfIters];
            for (uint32_t e = 0; e < 3 * kProfIters; e++) cudaEventCreate(&pool[e]);
        }
        prof = pool;
    }
    for (uint32_t i = 0; i < iter_limit; i++) {
        const int cur = i & 1, old = cur ^ 1;
        FrameState* s_cur = w.state + cur;
        if (i == 0) {
            // without prekill the initial state IS iteration 0's; with it, iteration 0 starts with a compaction like every other one
            frame_init_kernel<<<ceil_div<uint32_t>(N, 256), 256, 0, st>>>(N, sample_budget, max_n_step, w.nears, w.fars, w.rays_alive[prekill ? old : cur],
                                                                          w.rays_t[prekill ? old : cur], prekill ? w.state + old : s_cur, host_mailbox, prekill,
                                                                          rays_o, rays_d, bound, dt_gamma, max_steps, C, H, occupancy_mip);
            kernels += 1;
        }
        if (i > 0 || prekill) {
            compact_rays_kernel<<<ceil_div<uint32_t>(bound_rays, kCompactThreads * kCompactItems), kCompactThreads, 0, st>>>(
                bound_rays, w.rays_alive[cur], w.rays_alive[old], w.rays_t[cur], w.rays_t[old], nullptr, w.scan, w.state + old, s_cur, sample_budget, max_n_step, step_limit,
                host_mailbox + i);
            kernels += 1;
        }
        const bool timed = prof && i < kProfIters;
        if (timed) { cudaEventRecord(prof[3 * i], st); prof_iters = i + 1; }
        march_rays_compact_kernel<<<ceil_div<uint32_t>(bound_rays, kMarchThreads), kMarchThreads, 0, st>>>(
            w.rays_alive[cur], w.rays_t[cur], rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, w.fars, w.ts, w.deltas, perturb, prekill ? nullptr : occupancy_mip, s_cur, sample_counter,
            walk_budget, w.live, &s_cur->n_live);
        const uint32_t m_bound = (uint32_t)min((uint64_t)max(N, sample_budget), (uint64_t)bound_rays * max_n_step) + 128u;
        if (timed) cudaEventRecord(prof[3 * i + 1], st);
        // the field runs over the list of rows the marcher filled (no tile is spent on sentinel rows); positions and view directions
        // are rebuilt from (ray, t)
        const int rc = launch_ngp_field_rays(w.live, &s_cur->n_live, m_bound, w.ts, rays_o, rays_d, bound, embeddings_f16, offsets, L, S, base_resolution, align_corners,
                                             w_sigma_f16, w_color_f16, density_scale, w.sigmas, w.rgbs, st);
        if (rc != NTX_OK) return rc;
        if (timed) cudaEventRecord(prof[3 * i + 2], st);
        composite_rays_kernel<<<ceil_div<uint32_t>(bound_rays, 128), 128, 0, st>>>(bound_rays, 1, w.rays_alive[cur], w.rays_t[cur], w.sigmas, w.rgbs, w.deltas, weights_sum,
                                                                                  depth, image, s_cur);
        kernels += 3;                                       // march, field, composite
        cudaEventRecord(ev[i % kEvents], st);
        if (i >= ahead) {
            // iteration i-ahead has certainly been planned once its event fires; the iterations queued behind it keep the device busy
            // while this thread looks at the mailbox
            const uint32_t j = i - ahead;
            if (cudaEventSynchronize(ev[j % kEvents]) != cudaSuccess) return check_launch("render_rays");
            const int alive = host_mailbox[j];
            if (alive <= 0) break;            // iteration j found nothing alive; the iterations queued after it are no-ops
            iterations = j + 1;               // iterations 0 .. j did work
            bound_rays = (uint32_t)alive;     // n_alive never grows
        }
    }
    if (stats_out) { stats_out[0] = iterations; stats_out[1] = kernels; }
    if (prof && prof_iters) {
        if (cudaStreamSynchronize(st) != cudaSuccess) return check_launch("render_rays");
        for (uint32_t i = 0; i < prof_iters; i++) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, prof[3 * i], prof[3 * i + 1]);
            cudaEventElapsedTime(&b, prof[3 * i + 1], prof[3 * i + 2]);
            kernel_ms_out[0] += a; kernel_ms_out[1] += b;
        }
    }
    return check_launch("render_rays");
}

// ---------------------------------------------------------------------------------------------------- sharded frame assembly
// The N>1 path's epilogue in one kernel ("one exchange step at the end ... followed by a local un-permute"): every rank
// renders interleaved tiles of `tile` consecutive rays (tile k -> rank k % world) straight into a planar block
//   [weights_sum (n_max) | depth (n_max) | rgb (3 n_max)]
// the blocks are all-gathered as they are, and this kernel puts every ray back at its image position and adds the background term
// image + (1 - weights_sum) * bg (renderer.py:485) on the way.  world = 1 is the single-GPU epilogue (identity permutation).
namespace ntx {
__global__ void __launch_bounds__(256) unshard_frame_kernel(const float* __restrict__ gathered, const uint32_t world, const uint32_t n_max, const uint32_t tile,
                                                            const uint32_t N, const float bg, float* __restrict__ image, float* __restrict__ depth,
                                                            float* __restrict__ weights_sum) {
    const uint32_t ray = blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= N) return;
    const uint32_t k = ray / tile, rank = k % world, j = (k / world) * tile + (ray - k * tile);
    const float* blk = gathered + (size_t)rank * 5u * n_max;
    const float ws = blk[j];
    weights_sum[ray] = ws;
    depth[ray] = blk[n_max + j];
    const float b = __fmul_rn(__fsub_rn(1.0f, ws), bg);       // one rounding per torch op of `image + (1 - weights_sum).unsqueeze(-1) * bg_color`: no FMA
    const float* c = blk + 2u * (size_t)n_max + 3u * (size_t)j;
    image[(size_t)ray * 3] = __fadd_rn(c[0], b); image[(size_t)ray * 3 + 1] = __fadd_rn(c[1], b); image[(size_t)ray * 3 + 2] = __fadd_rn(c[2], b);
}
}  // namespace ntx

// The same assembly without the collective: every rank's planar block lives in NVLink-mapped symmetric memory, and this kernel READS
// the other ranks' blocks in place (ld.global on peer addresses) while it un-permutes — exchange, un-permute and background term in
// one kernel, 20 B per ray over NVLink, no staging copy.  `peers` = device array of `world` base pointers (this rank's view of every
// rank's buffer, e.g. torch.distributed._symmetric_memory's buffer_ptrs_dev); the caller orders it after a cross-rank barrier.
namespace ntx {
__global__ void __launch_bounds__(256) unshard_frame_peers_kernel(const float* const* __restrict__ peers, const size_t offset_floats, const uint32_t world,
                                                                  const uint32_t n_max, const uint32_t tile, const uint32_t N, const float bg,
                                                                  float* __restrict__ image, float* __restrict__ depth, float* __restrict__ weights_sum) {
    const uint32_t ray = blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= N) return;
    const uint32_t k = ray / tile, rank = k % world, j = (k / world) * tile + (ray - k * tile);
    const float* blk = peers[rank] + offset_floats;
    const float ws = blk[j];
    weights_sum[ray] = ws;
    depth[ray] = blk[n_max + j];
    const float b = __fmul_rn(__fsub_rn(1.0f, ws), bg);
    const float* c = blk + 2u * (size_t)n_max + 3u * (size_t)j;
    image[(size_t)ray * 3] = __fadd_rn(c[0], b); image[(size_t)ray * 3 + 1] = __fadd_rn(c[1], b); image[(size_t)ray * 3 + 2] = __fadd_rn(c[2], b);
}
}  // namespace ntx

extern "C" int ntx_unshard_frame_peers(const void* peers_dev, size_t offset_floats, uint32_t world, uint32_t n_max, uint32_t tile, uint32_t N, float bg,
                                       float* image, float* depth, float* weights_sum, ntx_stream_t stream) {
    NTX_REQUIRE(peers_dev && image && depth && weights_sum, NTX_ERR_INVALID_ARGUMENT, "unshard_frame_peers: null pointer");
    NTX_REQUIRE(world >= 1 && tile >= 1 && n_max >= 1, NTX_ERR_INVALID_ARGUMENT, "unshard_frame_peers: bad world / tile / n_max");
    if (N == 0) return NTX_OK;
    unshard_frame_peers_kernel<<<ceil_div<uint32_t>(N, 256), 256, 0, ST(stream)>>>(static_cast<const float* const*>(peers_dev), offset_floats, world, n_max, tile, N, bg,
                                                                                   image, depth, weights_sum);
    return check_launch("unshard_frame_peers");
}

extern "C" int ntx_unshard_frame(const float* gathered, uint32_t world, uint32_t n_max, uint32_t tile, uint32_t N, float bg, float* image, float* depth,
                                 float* weights_sum, ntx_stream_t stream) {
    NTX_REQUIRE(gathered && image && depth && weights_sum, NTX_ERR_INVALID_ARGUMENT, "unshard_frame: null pointer");
    NTX_REQUIRE(world >= 1 && tile >= 1 && n_max >= 1, NTX_ERR_INVALID_ARGUMENT, "unshard_frame: bad world / tile / n_max");
    if (N == 0) return NTX_OK;
    unshard_frame_kernel<<<ceil_div<uint32_t>(N, 256), 256, 0, ST(stream)>>>(gathered, world, n_max, tile, N, bg, image, depth, weights_sum);
    return check_launch("unshard_frame");
}
