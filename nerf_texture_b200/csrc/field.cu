// field.cu — the whole per-sample field of nerf/network_ff.py:85-101 in ONE kernel (fp16 inference):
//
//   xyz -> [0,1]^3 -> hash-grid gather (L levels x 2 features, fp16 table, lane-pair gather of grid_common.cuh)
//       -> sigma MLP  (2L -> 64 -> .. -> 16)  on tcgen05      -> sigma = exp(h[0]) * density_scale
//   dir -> SH degree 4 (16) ++ h[1..15] ++ 0  = 32
//       -> colour MLP (32 -> 64 -> .. -> 16)  on tcgen05      -> rgb = sigmoid(h[0..2])
//
// The reference runs this as 4 extension calls + ~10 torch glue kernels with every intermediate (features 64 B, padded
// copies, geo_feat, SH 64 B, concatenated colour input 64 B ...) making a round trip through HBM.
// Here a CTA owns 128-sample tiles and is warp-specialised:
//   * 12 producer warps gather features straight into the shared-memory A operand of the first UMMA and evaluate SH into
//     the colour net's A operand.  A warp (16 lane pairs) takes 16-row slices of the CTA's tile sequence round-robin, so
//     the producers run up to two tiles ahead of the consumers;
//   * 4 consumer warps (one per TMEM lane quarter) run the two MLPs of the previous tile: one thread issues the
//     tcgen05.mma of a layer, the warps pull the accumulator out of TMEM, apply ReLU, round to fp16 and write the next
//     layer's A operand; geo_feat goes from the sigma net's accumulator into the colour net's A tile;
//   * a 3-slot ring of A tiles with full/empty mbarriers decouples them, so the latency-bound MMA->epilogue chain (7
//     dependent stages per tile) hides behind the gather instead of serialising with it;
//   * the gather is bound by loads in flight and issue slots, i.e. by resident producer warps (tools/field_probe.py): the
//     kernel is held to 64 registers so that 2 CTAs x (12 producer + 4 consumer) warps fit on an SM.
// Only xyz/dir (24 B) come in and sigma/rgb (16 B) go out per sample.
#include "grid_common.cuh"
#include "mlp_tile.cuh"
#include "sh_poly.cuh"

namespace ntx {

constexpr int kConsumerWarps = 4;
constexpr int kProducerWarps = 12;
constexpr int kTaskRows = 16;                                            // rows one producer warp gathers at a time
constexpr int kTasksPerTile = 8;                                         // kTileRows / kTaskRows
constexpr int kFieldThreads = 32 * (kConsumerWarps + kProducerWarps);   // 512
constexpr int kFW = 64;          // hidden width of both MLPs
constexpr int kColorIn = 32;     // SH(16) + geo_feat(15) + zero pad (network_ff.py:42,95-97)
constexpr int kShDim = 16;       // SH degree 4
constexpr int kFieldMaxLevels = 32;
constexpr int kStages = 3;

struct FieldPlan {
    uint32_t k0;                 // sigma-net input width = 2L
    uint32_t ws_bytes, wc_bytes; // weight bytes of the two nets
    uint32_t ws_off, wc_off, a0_off, h_off, lv_off, misc_off, total;
    uint32_t a0_stage;
};
__host__ __device__ inline FieldPlan field_plan(uint32_t L, uint32_t ns, uint32_t nc) {
    FieldPlan p;
    p.k0 = 2 * L;
    p.ws_bytes = 2u * (kFW * p.k0 + (ns - 1) * kFW * kFW + 16u * kFW);
    p.wc_bytes = 2u * (kFW * kColorIn + (nc - 1) * kFW * kFW + 16u * kFW);
    p.ws_off = 0;
    p.wc_off = p.ws_off + p.ws_bytes;
    p.a0_stage = kTileRows * p.k0 * 2u;
    p.a0_off = (p.wc_off + p.wc_bytes + 127u) & ~127u;
    p.h_off = p.a0_off + kStages * p.a0_stage;
    p.lv_off = p.h_off + kTileRows * kFW * 2u;
    p.misc_off = p.lv_off + (uint32_t)sizeof(PairLevel) * kFieldMaxLevels;
    p.total = p.misc_off + 128u;
    return p;
}

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kConsumerWarps) : "memory"); }

// hidden-layer epilogue (4 consumer warps, warp w owns TMEM lanes 32w..32w+31 = tile rows): four 16-column passes
// (keeps the consumer inside the kernel's 64-register budget)
__device__ __forceinline__ void field_hidden_epilogue(uint32_t tmem_base, uint8_t* h_smem, uint32_t warp, uint32_t lane) {
    const uint32_t row = warp * 32 + lane;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t col = q * 16;
        uint32_t v[16];
        tc::tmem_ld_x16(tmem_base + ((warp * 32u) << 16) + col, v);
        tc::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
            uint4 o;
            o.x = act_pack2(0, v[j + 0], v[j + 1]);
            o.y = act_pack2(0, v[j + 2], v[j + 3]);
            o.z = act_pack2(0, v[j + 4], v[j + 5]);
            o.w = act_pack2(0, v[j + 6], v[j + 7]);
            *reinterpret_cast<uint4*>(h_smem + kmajor_chunk_off(row, (col + j) >> 3, kFW)) = o;
        }
    }
}

__global__ void __launch_bounds__(kFieldThreads, 2) ngp_field_kernel(
    const float* __restrict__ xyz, const float* __restrict__ dirs, const float* __restrict__ deltas, const uint32_t M, const float bound,
    const __half* __restrict__ table, const int* __restrict__ offsets, const uint32_t L, const float S, const uint32_t H, const bool align,
    const __half* __restrict__ w_sigma, const __half* __restrict__ w_color, const uint32_t ns, const uint32_t nc, const float density_scale,
    float* __restrict__ sigmas, float* __restrict__ rgbs, const uint32_t dbg) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const FieldPlan plan = field_plan(L, ns, nc);
    uint8_t* ws_smem = smem + plan.ws_off;
    uint8_t* wc_smem = smem + plan.wc_off;
    uint8_t* a0_smem = smem + plan.a0_off;   // kStages tiles
    uint8_t* h_smem = smem + plan.h_off;
    PairLevel* lv = reinterpret_cast<PairLevel*>(smem + plan.lv_off);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + plan.misc_off);        // [kStages] producers -> consumer
    uint64_t* empty_bar = full_bar + kStages;                                      // [kStages] consumer (MMA completion) -> producers
    uint64_t* mma_bar = empty_bar + kStages;                                       // layer done -> consumer warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t K0 = plan.k0;

    // ---- one-time setup (all 12 warps) ---------------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < kStages; s++) { tc::mbar_init(&full_bar[s], kTasksPerTile); tc::mbar_init(&empty_bar[s], 1); }
        tc::mbar_init(mma_bar, 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc<64>(tmem_slot);
    if (tid < L) lv[tid] = make_pair_level(make_level<3>(offsets, tid, S, H, /*gridtype=*/0, align), table);
    {
        const __half* w = w_sigma; uint8_t* dst = ws_smem;
        load_matrix_kmajor(dst, w, kFW, K0, tid, kFieldThreads); w += (size_t)kFW * K0; dst += (size_t)kFW * K0 * 2;
        for (uint32_t k = 0; k + 1 < ns; k++) { load_matrix_kmajor(dst, w, kFW, kFW, tid, kFieldThreads); w += kFW * kFW; dst += kFW * kFW * 2; }
        load_matrix_kmajor(dst, w, 16, kFW, tid, kFieldThreads);
        w = w_color; dst = wc_smem;
        load_matrix_kmajor(dst, w, kFW, kColorIn, tid, kFieldThreads); w += kFW * kColorIn; dst += kFW * kColorIn * 2;
        for (uint32_t k = 0; k + 1 < nc; k++) { load_matrix_kmajor(dst, w, kFW, kFW, tid, kFieldThreads); w += kFW * kFW; dst += kFW * kFW * 2; }
        load_matrix_kmajor(dst, w, 16, kFW, tid, kFieldThreads);
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t ntiles = ceil_div<uint32_t>(M, kTileRows);

    if (warp >= kConsumerWarps) {
        // =============================== PRODUCERS: gather + SH =====================================================
        const uint32_t pw = warp - kConsumerWarps;             // 0..11
        const uint32_t p = lane & 1u;                          // which x corner this lane gathers
        const float half_off = align ? 0.0f : 0.5f;
        // GridEncoder.forward: inputs = (inputs + bound) / (2 * bound) (grid.py:143).  torch's CUDA true-division by a Python
        // scalar multiplies by the fp32 reciprocal (BinaryDivTrueKernel.cu), so do exactly that.
        const float inv2b = 1.0f / (2.0f * bound);
        for (uint32_t task = pw;; task += kProducerWarps) {
            const uint32_t k = task / kTasksPerTile;            // k-th tile of this CTA
            const uint32_t tile = blockIdx.x + k * gridDim.x;
            if (tile >= ntiles) break;
            const uint32_t s = k % kStages, use = k / kStages;
            const uint32_t srow = (task % kTasksPerTile) * kTaskRows + (lane >> 1);   // row of this lane pair inside the tile
            const uint32_t b = tile * kTileRows + srow;
            const bool valid = b < M;
            float x = 0.f, y = 0.f, z = 0.f;
            bool skip = !valid;
            if (valid) {
                x = __fmul_rn(__fadd_rn(xyz[(size_t)b * 3], bound), inv2b);
                y = __fmul_rn(__fadd_rn(xyz[(size_t)b * 3 + 1], bound), inv2b);
                z = __fmul_rn(__fadd_rn(xyz[(size_t)b * 3 + 2], bound), inv2b);
                if (deltas && deltas[(size_t)b * 2] == 0.0f) skip = true;  // sentinel slot of march_rays
            }
            const bool oob = (x < 0 || x > 1) || (y < 0 || y > 1) || (z < 0 || z > 1);
            const bool live = !skip && !oob;
            if (!live) { x = 0.f; y = 0.f; z = 0.f; }  // keeps the (discarded) loads of dead lanes in bounds
            tc::mbar_wait_relaxed(&empty_bar[s], (use & 1u) ^ 1u);   // slot free? (first use of a slot passes immediately)
            uint8_t* a0 = a0_smem + s * plan.a0_stage;
            for (uint32_t l0 = 0; l0 < L; l0 += 4) {
                uint32_t packed[2];
#ifdef NTX_DEV_PROBES
                if (dbg & 1u) { packed[0] = packed[1] = 0; } else
#endif
                pair_gather4<__half>(x, y, z, live, p, lv + l0, half_off, packed);
                // lane p owns levels l0+2p, l0+2p+1 -> 4 consecutive halfs (8 bytes) of the row
                const uint32_t kcol = 2 * (l0 + 2 * p);
                *reinterpret_cast<uint2*>(a0 + kmajor_off(srow, kcol, K0)) = make_uint2(packed[0], packed[1]);
            }
            tc::fence_proxy_async_smem();          // my generic-proxy writes -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full_bar[s]);
        }
    } else {
        // =============================== CONSUMERS: the two MLPs =====================================================
        const uint32_t ws_addr = tc::smem_u32(ws_smem), wc_addr = tc::smem_u32(wc_smem), h_addr = tc::smem_u32(h_smem);
        uint32_t mma_phase = 0;
        uint32_t k = 0;
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++k) {
            const uint32_t s = k % kStages, use = k / kStages;
            const uint32_t a0_addr = tc::smem_u32(a0_smem + s * plan.a0_stage);
            const uint32_t row = warp * 32 + lane;
            const uint32_t b = tile * kTileRows + row;
            const bool dead = (b >= M) || (deltas && deltas[(size_t)b * 2] == 0.0f);
            float dx = 0.f, dy = 0.f, dz = 0.f;     // view direction of this row: its SH basis is evaluated between the two nets
            if (b < M) { dx = dirs[(size_t)b * 3]; dy = dirs[(size_t)b * 3 + 1]; dz = dirs[(size_t)b * 3 + 2]; }

            tc::mbar_wait(&full_bar[s], use & 1u);
            tc::tc_fence_after_sync();
#ifdef NTX_DEV_PROBES
            if (dbg & 2u) {   // producer-only timing: release the slot at once, skip both MLPs
                if (tid == 0) tc::mbar_arrive(&empty_bar[s]);
                if (b < M) { st_stream_f32(sigmas + b, 0.f); for (int c = 0; c < 3; c++) st_stream_f32(rgbs + (size_t)b * 3 + c, 0.f); }
                continue;
            }
#endif
            // ---------------- sigma net ------------------------------------------------------------------------
            if (tid == 0) {
                issue_layer(a0_addr, ws_addr, K0, kFW, tmem_base);
                tc::mma_commit(mma_bar);
                tc::mma_commit(&empty_bar[s]);   // the feature tile is free again as soon as this MMA has read it
            }
            tc::mbar_wait(mma_bar, mma_phase); mma_phase ^= 1;
            tc::tc_fence_after_sync();
            field_hidden_epilogue(tmem_base, h_smem, warp, lane);
            tc::fence_proxy_async_smem(); tc::tc_fence_before_sync(); consumer_sync();
            uint32_t w_off = kFW * K0 * 2u;
            for (uint32_t j = 0; j + 1 < ns; j++) {
                if (tid == 0) { tc::tc_fence_after_sync(); issue_layer(h_addr, ws_addr + w_off, kFW, kFW, tmem_base); tc::mma_commit(mma_bar); }
                tc::mbar_wait(mma_bar, mma_phase); mma_phase ^= 1;
                tc::tc_fence_after_sync();
                field_hidden_epilogue(tmem_base, h_smem, warp, lane);
                tc::fence_proxy_async_smem(); tc::tc_fence_before_sync(); consumer_sync();
                w_off += kFW * kFW * 2u;
            }
            if (tid == 0) { tc::tc_fence_after_sync(); issue_layer(h_addr, ws_addr + w_off, kFW, 16, tmem_base); tc::mma_commit(mma_bar); }
            tc::mbar_wait(mma_bar, mma_phase); mma_phase ^= 1;
            tc::tc_fence_after_sync();
            {
                uint32_t v[16];
                tc::tmem_ld_x16(tmem_base + ((warp * 32u) << 16), v);
                tc::tmem_wait_ld();
                // h = fp16(W2 . a)   (FFMLP output is fp16); sigma = trunc_exp(h[0]) = exp(float(h[0]))   (network_ff.py:88)
                uint32_t hb[8];
#pragma unroll
                for (int j = 0; j < 8; j++) hb[j] = float2_to_half2_bits(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                const float h0 = __low2float(*reinterpret_cast<const __half2*>(&hb[0]));
                if (b < M) st_stream_f32(sigmas + b, dead ? 0.0f : density_scale * expf(h0));
                // colour-net input [128 x 32] = SH(dir) ++ h[1..15] ++ 0   (network_ff.py:93-97), built at the start of h_smem: the
                // hidden activations there are dead (the output-layer MMA has completed), and the colour net's first epilogue
                // overwrites it only after its MMA has read it.
                uint4 c2, c3;
                c2.x = __byte_perm(hb[0], hb[1], 0x5432); c2.y = __byte_perm(hb[1], hb[2], 0x5432);
                c2.z = __byte_perm(hb[2], hb[3], 0x5432); c2.w = __byte_perm(hb[3], hb[4], 0x5432);
                c3.x = __byte_perm(hb[4], hb[5], 0x5432); c3.y = __byte_perm(hb[5], hb[6], 0x5432);
                c3.z = __byte_perm(hb[6], hb[7], 0x5432); c3.w = __byte_perm(hb[7], 0u, 0x5432);
                *reinterpret_cast<uint4*>(h_smem + kmajor_chunk_off(row, 2, kColorIn)) = c2;
                *reinterpret_cast<uint4*>(h_smem + kmajor_chunk_off(row, 3, kColorIn)) = c3;
                // SH of the view direction (fp32, rounded to fp16 when it enters the fp16 MLP)
                float sh[16];
                sh_basis<4, false>(dx, dy, dz, sh, nullptr, nullptr, nullptr);
                uint4 s0, s1;
                s0.x = float2_to_half2_bits(sh[0], sh[1]); s0.y = float2_to_half2_bits(sh[2], sh[3]);
                s0.z = float2_to_half2_bits(sh[4], sh[5]); s0.w = float2_to_half2_bits(sh[6], sh[7]);
                s1.x = float2_to_half2_bits(sh[8], sh[9]); s1.y = float2_to_half2_bits(sh[10], sh[11]);
                s1.z = float2_to_half2_bits(sh[12], sh[13]); s1.w = float2_to_half2_bits(sh[14], sh[15]);
                *reinterpret_cast<uint4*>(h_smem + kmajor_chunk_off(row, 0, kColorIn)) = s0;
                *reinterpret_cast<uint4*>(h_smem + kmajor_chunk_off(row, 1, kColorIn)) = s1;
            }
            tc::fence_proxy_async_smem(); tc::tc_fence_before_sync(); consumer_sync();

            // ---------------- colour net -----------------------------------------------------------------------
            if (tid == 0) {
                tc::tc_fence_after_sync();
                issue_layer(h_addr, wc_addr, kColorIn, kFW, tmem_base);
                tc::mma_commit(mma_bar);
            }
            tc::mbar_wait(mma_bar, mma_phase); mma_phase ^= 1;
            tc::tc_fence_after_sync();
            field_hidden_epilogue(tmem_base, h_smem, warp, lane);
            tc::fence_proxy_async_smem(); tc::tc_fence_before_sync(); consumer_sync();
            w_off = kFW * kColorIn * 2u;
            for (uint32_t j = 0; j + 1 < nc; j++) {
                if (tid == 0) { tc::tc_fence_after_sync(); issue_layer(h_addr, wc_addr + w_off, kFW, kFW, tmem_base); tc::mma_commit(mma_bar); }
                tc::mbar_wait(mma_bar, mma_phase); mma_phase ^= 1;
                tc::tc_fence_after_sync();
                field_hidden_epilogue(tmem_base, h_smem, warp, lane);
                tc::fence_proxy_async_smem(); tc::tc_fence_before_sync(); consumer_sync();
                w_off += kFW * kFW * 2u;
            }
            if (tid == 0) { tc::tc_fence_after_sync(); issue_layer(h_addr, wc_addr + w_off, kFW, 16, tmem_base); tc::mma_commit(mma_bar); }
            tc::mbar_wait(mma_bar, mma_phase); mma_phase ^= 1;
            tc::tc_fence_after_sync();
            {
                uint32_t v[8];
                tc::tmem_ld_x8(tmem_base + ((warp * 32u) << 16), v);
                tc::tmem_wait_ld();
                if (b < M) {
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        // rgb = sigmoid(h) evaluated on the fp16 output in fp32 and rounded back to fp16 (torch.sigmoid on a half tensor)
                        const float hc = __half2float(__float2half_rn(__uint_as_float(v[c])));
                        const float sg = 1.0f / (1.0f + expf(-hc));
                        st_stream_f32(rgbs + (size_t)b * 3 + c, dead ? 0.0f : __half2float(__float2half_rn(sg)));
                    }
                }
            }
            tc::tc_fence_before_sync();
            consumer_sync();   // TMEM accumulator is reused by the next tile's first MMA
        }
    }

    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<64>(tmem_base);
}

}  // namespace ntx

using namespace ntx;

// development probes (only in builds with -DNTX_DEV_PROBES; tools/field_probe.py): bit 0 = producers skip the gather,
// bit 1 = consumers skip the MLPs.  Product builds always pass 0 and the branches do not exist.
static uint32_t dev_probe_flags() {
#ifdef NTX_DEV_PROBES
    const char* e = getenv("NTX_FIELD_DEBUG");
    return e ? (uint32_t)atoi(e) : 0u;
#else
    return 0u;
#endif
}

extern "C" int ntx_ngp_field_forward(const float* xyz, const float* dirs, const float* deltas, uint32_t M, float bound, const void* embeddings_f16,
                                     const int* offsets, uint32_t L, float S, uint32_t H, int align_corners, const void* w_sigma_f16,
                                     const void* w_color_f16, float density_scale, float* sigmas, float* rgbs, ntx_stream_t stream) {
    NTX_REQUIRE(xyz && dirs && embeddings_f16 && offsets && w_sigma_f16 && w_color_f16 && sigmas && rgbs, NTX_ERR_INVALID_ARGUMENT, "ngp_field_forward: null pointer");
    NTX_REQUIRE(L >= 8 && L <= kFieldMaxLevels && L % 8 == 0, NTX_ERR_UNSUPPORTED, "ngp_field_forward: num_levels must be 8, 16, 24 or 32 (got %u)", L);
    NTX_REQUIRE(((uintptr_t)w_sigma_f16 & 15) == 0 && ((uintptr_t)w_color_f16 & 15) == 0, NTX_ERR_INVALID_ARGUMENT, "ngp_field_forward: weights must be 16-byte aligned");
    NTX_REQUIRE(bound > 0, NTX_ERR_INVALID_ARGUMENT, "ngp_field_forward: bound must be positive");
    if (M == 0) return NTX_OK;
    const uint32_t ns = 2, nc = 3;  // FFMLP(num_layers=2) and FFMLP(num_layers=3): nerf/network_ff.py:31-49
    const FieldPlan plan = field_plan(L, ns, nc);
    static int occ = 0;
    static uint32_t configured = 0;
    if (plan.total > configured) {
        if (cudaFuncSetAttribute(ngp_field_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total) != cudaSuccess) {
            cudaGetLastError();
            set_error("ngp_field_forward: cannot reserve %u bytes of shared memory", plan.total);
            return NTX_ERR_CUDA;
        }
        configured = plan.total;
        occ = 0;
    }
    if (!occ) {
        occ = resident_ctas_per_sm((const void*)ngp_field_kernel, kFieldThreads, plan.total, 64);
        if (tunables().field_ctas > 0) occ = std::min(occ, tunables().field_ctas);
    }
    const int sms = device_sm_count();
    const uint32_t ntiles = ceil_div<uint32_t>(M, kTileRows);
    const uint32_t grid = std::min<uint32_t>(ntiles, (uint32_t)(occ * sms));
    ngp_field_kernel<<<grid, kFieldThreads, plan.total, reinterpret_cast<cudaStream_t>(stream)>>>(
        xyz, dirs, deltas, M, bound, static_cast<const __half*>(embeddings_f16), offsets, L, S, H, align_corners != 0,
        static_cast<const __half*>(w_sigma_f16), static_cast<const __half*>(w_color_f16), ns, nc, density_scale, sigmas, rgbs, dev_probe_flags());
    return check_launch("ngp_field_forward");
}
