// mlp_wgrad.cuh — weight gradients of the fully-fused MLP on tcgen05 (included by mlp_bwd.cu).
//
// G[p x q] = P[B x p]^T . Q[B x q]: a skinny GEMM whose reduction dimension is the batch (the reference: CUTLASS split-K with
// fp16 accumulation on side streams, ffmlp.cu:804-875, cutlass_matmul.h:459).  Here a CTA walks its share of 128-row batch slabs;
// each slab of P and Q is staged TRANSPOSED into shared memory (batch = K becomes the contiguous dimension of a no-swizzle K-major
// operand: 8 two-byte stores per 16-byte global load, bank-conflict free thanks to a 16-byte pad between the 8-row groups) and
// eight K=16 UMMAs accumulate it into ONE fp32 accumulator in tensor memory that lives for the whole kernel — the batch reduction
// never leaves the tensor core.  Two slab stages: the loads of slab i+1 overlap the MMAs of slab i.
// M is always 128: rows p..127 of the A window are whatever lies behind the staged rows (an accumulator row depends on its own A
// row only, and rows >= p are never read back), so p = 16 .. 128 all take the same M=128 path as the forward kernels.
// Every CTA writes its partial G to its own slice of the workspace; wgrad_reduce_kernel sums the slices in a fixed order and rounds
// to fp16 once: deterministic, no atomics.
#pragma once

#include "mlp_tile.cuh"

namespace ntx {

constexpr int kWgThreads = 256;
constexpr uint32_t kWgSlab = 128;                               // batch rows per stage = K of 8 UMMAs
constexpr uint32_t kWgSbo = kWgSlab * 16u + 16u;                // bytes between 8-row groups (+16: transposed stores hit 32 distinct banks)
constexpr uint32_t kWgMaxParts = 480;                           // partial-sum slices per job in the workspace (three CTAs per SM of a B200)
constexpr int kWgStages = 2;

struct WgPlan { uint32_t a_bytes, b_bytes, stage_bytes, misc_off, total; };
__host__ __device__ inline WgPlan wg_plan(uint32_t p, uint32_t q) {
    WgPlan w;
    w.a_bytes = (p >> 3) * kWgSbo;
    w.b_bytes = (q >> 3) * kWgSbo;
    w.stage_bytes = (w.a_bytes + w.b_bytes + 127u) & ~127u;
    // the A window of an M=128 UMMA spans 16 row groups from a stage's start: keep the last stage's window inside the allocation
    const uint32_t stages_end = kWgStages * w.stage_bytes;
    const uint32_t window_end = (kWgStages - 1) * w.stage_bytes + 16u * kWgSbo;
    w.misc_off = ((stages_end > window_end ? stages_end : window_end) + 127u) & ~127u;
    w.total = w.misc_off + 64u;
    return w;
}

// [128 batch rows x cols] (row-major, leading dimension ld) -> K-major core-matrix tile [cols x 128]; rows >= rows_valid are zero
__device__ __forceinline__ void stage_slab_transposed(uint8_t* dst, const __half* __restrict__ src, uint32_t ld, uint32_t cols, uint32_t rows_valid,
                                                      uint32_t tid) {
    const uint32_t chunks = cols >> 3, total = kWgSlab * chunks;
    for (uint32_t e = tid; e < total; e += kWgThreads) {
        const uint32_t k = e / chunks, mc = e - k * chunks;      // consecutive lanes: consecutive 16-byte pieces of one batch row
        const uint4 v = (k < rows_valid) ? ld_stream_u4(src + (size_t)k * ld + mc * 8) : make_uint4(0u, 0u, 0u, 0u);
        uint8_t* d = dst + mc * kWgSbo + (k >> 3) * 128u + (k & 7u) * 2u;
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            *reinterpret_cast<uint16_t*>(d + (2 * j) * 16) = (uint16_t)(w[j] & 0xffffu);
            *reinterpret_cast<uint16_t*>(d + (2 * j + 1) * 16) = (uint16_t)(w[j] >> 16);
        }
    }
}

// one gradient block G = P^T Q of the backward pass; all blocks of a backward run in ONE launch (blockIdx.y = job) and one reduce
struct WgJob {
    const __half* P; const __half* Q;   // [B x ldp], [B x ldq] row-major; the first p / q columns are used
    float* partials;                    // [gridDim.x][p * q] scratch
    __half* dst;                        // fp16 result, leading dimension ld_dst; transposed: dst is [q x p]
    uint32_t ldp, p, ldq, q, ld_dst, transpose, tmem_cols, pad;
};
constexpr int kWgMaxJobs = 12;
struct WgJobs { WgJob j[kWgMaxJobs]; };

__global__ void __launch_bounds__(kWgThreads) mlp_wgrad_tc_kernel(const __grid_constant__ WgJobs jobs, const uint32_t B) {
    const WgJob& job = jobs.j[blockIdx.y];
    const __half* __restrict__ P = job.P;
    const __half* __restrict__ Q = job.Q;
    float* __restrict__ partials = job.partials;
    const uint32_t ldp = job.ldp, p = job.p, ldq = job.ldq, q = job.q, tmem_cols = job.tmem_cols;
    extern __shared__ __align__(1024) uint8_t smem[];
    const WgPlan plan = wg_plan(p, q);
    uint64_t* empty_bar = reinterpret_cast<uint64_t*>(smem + plan.misc_off);    // [kWgStages] the MMAs of a stage have read it
    uint64_t* acc_bar = empty_bar + kWgStages;                                  // all MMAs of this CTA have completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) { for (int s = 0; s < kWgStages; s++) tc::mbar_init(&empty_bar[s], 1); tc::mbar_init(acc_bar, 1); tc::fence_mbar_init(); }
    if (warp == 0) {   // power-of-two column count >= q (run-time: q depends on the layer)
        if (tmem_cols == 32) tc::tmem_alloc<32>(tmem_slot); else if (tmem_cols == 64) tc::tmem_alloc<64>(tmem_slot);
        else if (tmem_cols == 128) tc::tmem_alloc<128>(tmem_slot); else tc::tmem_alloc<256>(tmem_slot);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = tc::idesc_f16_f32(128, q);
    const uint32_t nslabs = ceil_div<uint32_t>(B, kWgSlab);

    uint32_t it = 0;
    for (uint32_t slab = blockIdx.x; slab < nslabs; slab += gridDim.x, it++) {
        const uint32_t s = it % kWgStages, use = it / kWgStages;
        tc::mbar_wait(&empty_bar[s], (use & 1u) ^ 1u);                          // first use of a stage passes at once
        uint8_t* a_smem = smem + s * plan.stage_bytes;
        uint8_t* b_smem = a_smem + plan.a_bytes;
        const size_t row0 = (size_t)slab * kWgSlab;
        const uint32_t rows_valid = (uint32_t)min((size_t)kWgSlab, (size_t)B - row0);
        stage_slab_transposed(a_smem, P + row0 * ldp, ldp, p, rows_valid, tid);
        stage_slab_transposed(b_smem, Q + row0 * ldq, ldq, q, rows_valid, tid);
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0 && tc::elect_one()) {
            tc::tc_fence_after_sync();
            const uint32_t a_addr = tc::smem_u32(a_smem), b_addr = tc::smem_u32(b_smem);
#pragma unroll
            for (uint32_t ks = 0; ks < kWgSlab / 16; ks++) {
                const uint64_t da = tc::smem_desc_kmajor_noswz(a_addr + ks * 256u, 128u, kWgSbo);
                const uint64_t db = tc::smem_desc_kmajor_noswz(b_addr + ks * 256u, 128u, kWgSbo);
                tc::mma_f16_ss(tmem_base, da, db, idesc, (it > 0 || ks > 0) ? 1u : 0u);
            }
            tc::mma_commit(&empty_bar[s]);
        }
    }
    if (warp == 0 && tc::elect_one()) tc::mma_commit(acc_bar);                 // arrives when every MMA issued above has completed
    tc::mbar_wait(acc_bar, 0);
    tc::tc_fence_after_sync();
    float* out = partials + (size_t)blockIdx.x * p * q;
    if (warp < 4) {
        const uint32_t row = warp * 32 + lane;
        for (uint32_t c0 = 0; c0 < q; c0 += 16) {
            uint32_t v[16];
            tc::tmem_ld_x16(tmem_base + ((warp * 32u) << 16) + c0, v);
            tc::tmem_wait_ld();
            if (row < p) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) *reinterpret_cast<uint4*>(out + (size_t)row * q + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        if (tmem_cols == 32) tc::tmem_dealloc<32>(tmem_base); else if (tmem_cols == 64) tc::tmem_dealloc<64>(tmem_base);
        else if (tmem_cols == 128) tc::tmem_dealloc<128>(tmem_base); else tc::tmem_dealloc<256>(tmem_base);
    }
}

// job.dst (fp16) = sum over the CTAs' partial G[p x q] slices of that job; transpose: dst is [q x p] with leading dimension ld_dst (the output
// layer's gradient is produced as activations^T . dY and stored as dW[16 x hidden]).  A block owns 32 consecutive entries: warp w
// sums slices w, w+8, ... (coalesced 128-byte reads, independent loads in flight), the eight per-warp sums are added in warp order —
// a fixed summation tree, so the result does not depend on scheduling.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const __grid_constant__ WgJobs jobs, const uint32_t n_parts) {
    const WgJob& job = jobs.j[blockIdx.y];
    const float* __restrict__ partials = job.partials;
    const uint32_t p = job.p, q = job.q;
    __shared__ float s_part[8][33];
    const uint32_t n = p * q, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t i = blockIdx.x * 32u + lane;
    if (blockIdx.x * 32u >= n) return;
    float acc = 0.f;
    if (i < n) {
#pragma unroll 4
        for (uint32_t c = warp; c < n_parts; c += 8) acc += partials[(size_t)c * n + i];
    }
    s_part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && i < n) {
        float t = s_part[0][lane];
#pragma unroll
        for (int w = 1; w < 8; w++) t += s_part[w][lane];
        const uint32_t r = i / q, col = i - r * q;
        job.dst[job.transpose ? (size_t)col * job.ld_dst + r : (size_t)r * job.ld_dst + col] = __float2half_rn(t);
    }
}

// number of batch-slab CTAs per job (= partial-sum slices per job in the workspace)
static uint32_t wgrad_parts(uint32_t B) {
    const uint32_t nslabs = ceil_div<uint32_t>(B, kWgSlab);
    return std::min<uint32_t>(nslabs, std::min<uint32_t>(3u * (uint32_t)device_sm_count(), kWgMaxParts));
}

// all jobs of one backward: ONE tcgen05 launch (grid = parts x jobs) + ONE reduce launch
static int launch_wgrad_jobs(WgJobs& jobs, uint32_t njobs, float* ws, uint32_t B, cudaStream_t st) {
    const uint32_t parts = wgrad_parts(B);
    uint32_t smem = 0, max_n = 0;
    size_t off = 0;
    for (uint32_t i = 0; i < njobs; i++) {
        WgJob& j = jobs.j[i];
        smem = std::max(smem, wg_plan(j.p, j.q).total);
        max_n = std::max(max_n, j.p * j.q);
        j.tmem_cols = j.q <= 32 ? 32u : j.q <= 64 ? 64u : j.q <= 128 ? 128u : 256u;
        j.partials = ws + off;
        off += (size_t)parts * j.p * j.q;
    }
    static uint32_t configured_dev[kMaxDevices] = {};
    uint32_t& configured = configured_dev[current_device()];
    if (smem > configured) {
        if (cudaFuncSetAttribute(mlp_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            set_error("FullyFusedMLP backward: cannot reserve %u bytes of shared memory for the weight-gradient kernel", smem);
            return NTX_ERR_CUDA;
        }
        configured = smem;
    }
    mlp_wgrad_tc_kernel<<<dim3(parts, njobs), kWgThreads, smem, st>>>(jobs, B);
    wgrad_reduce_kernel<<<dim3(ceil_div<uint32_t>(max_n, 32), njobs), 256, 0, st>>>(jobs, parts);
    return check_launch("ffmlp_backward(wgrad)");
}

}  // namespace ntx
