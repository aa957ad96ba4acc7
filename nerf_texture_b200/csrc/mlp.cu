// mlp.cu — fully-fused fp16 MLP on tcgen05 tensor cores (sm_100a).
// Replaces ffmlp/src/ffmlp.cu: kernel_mlp_fused (:332, wmma m16n16k16 with fp16 accumulators, weights re-read from
// L2 by each of 8192 CTAs) and kernel_mlp_fused_backward (:411) + the CUTLASS split-K weight-gradient GEMMs (:804-875).
//
// Forward / inference: persistent CTAs, all weight matrices resident in shared memory for the CTA's lifetime; input tiles arrive
// through a TMA ring, hidden activations stay in tensor memory between the layers (tcgen05.mma with the A operand in TMEM), two
// 128-row tiles in flight per CTA, everything synchronised with mbarriers (see mlp_pipe_kernel).  Accumulation is fp32 (TMEM);
// activations are rounded to fp16 once per layer, as in the reference.
#include "mlp_tile.cuh"

namespace ntx {

// ==================================================================================================== the kernel
// mlp_pipe_kernel: organised so that nothing waits for anything it does not depend on.
//   * input tiles arrive through a TMA ring (cp.async.bulk.tensor, one producer lane): in_dim/8 box loads of [128 rows x 16 B]
//     land a tile in "chunk-major" order, byte(r,k) = (k/8)*2048 + r*16 + (k%8)*2 — a no-swizzle K-major UMMA operand with
//     LBO = 2048 (K-adjacent core matrices) and SBO = 128 (M-adjacent); rows past B are zero-filled by the TMA unit;
//   * hidden activations never leave tensor memory: the epilogue of layer i reads the fp32 accumulator with tcgen05.ld,
//     applies the activation, rounds to fp16 and writes the tile back with tcgen05.st as the A operand of layer i+1
//     (tcgen05.mma with A in TMEM) — no STS, no generic->async proxy fence, no CTA-wide barrier per layer;
//   * up to two tile contexts per CTA, each with its own accumulator + operand columns, its own four epilogue warps and its
//     own MMA-issuing warp, synchronised by mbarriers only: while one tile's epilogue runs, the other tile's MMA does.
// TMEM per context: WIDTH accumulator columns + WIDTH/2 operand columns.
constexpr uint32_t kChunkBytes = kTileRows * 16u;           // one [128 x 8 halfs] chunk of a TMA-landed tile

template <int WIDTH> struct PipeCfg {
    static constexpr int kCtx = WIDTH <= 128 ? 2 : 1;
    static constexpr int kThreads = 32 * (5 * kCtx + 1);                        // 4 epilogue warps + 1 MMA warp per context, 1 TMA warp
    static constexpr uint32_t kCtxCols = (WIDTH < 32 ? 32 : WIDTH) + (WIDTH < 32 ? 16 : WIDTH / 2);
    static constexpr uint32_t kUsedCols = kCtx * kCtxCols;
    static constexpr uint32_t kTmemCols = kUsedCols <= 32 ? 32 : kUsedCols <= 64 ? 64 : kUsedCols <= 128 ? 128 : kUsedCols <= 256 ? 256 : 512;
    static constexpr uint32_t kAOff = WIDTH < 32 ? 32 : WIDTH;                  // operand columns follow the accumulator
    static constexpr int kMinCtas = (512 / kTmemCols) < 2 ? 1 : 2;              // register budget: two CTAs per SM whenever tensor memory allows it
};

struct PipePlan { uint32_t w_bytes, in_off, stage_bytes, misc_off, total; };
__host__ __device__ inline PipePlan pipe_plan(uint32_t in_dim, uint32_t hidden, uint32_t num_layers, uint32_t stages) {
    PipePlan p;
    p.w_bytes = 2u * (hidden * in_dim + (num_layers - 1) * hidden * hidden + 16u * hidden);
    p.in_off = (p.w_bytes + 127u) & ~127u;
    p.stage_bytes = kTileRows * in_dim * 2u;
    p.misc_off = p.in_off + stages * p.stage_bytes;
    p.total = p.misc_off + 256u;
    return p;
}
constexpr uint32_t kPipeMaxStages = 8;

// D[128 x N] = A[128 x Kdim](TMA-landed, chunk-major) . W[N x Kdim]^T
__device__ __forceinline__ void issue_layer_chunked(uint32_t a_smem, uint32_t w_smem, uint32_t Kdim, uint32_t N, uint32_t tmem_d) {
    const uint32_t idesc = tc::idesc_f16_f32(kTileRows, N);
    for (uint32_t ks = 0; ks < (Kdim >> 4); ks++) {
        const uint64_t da = tc::smem_desc_kmajor_noswz(a_smem + ks * 2u * kChunkBytes, kChunkBytes, 128u);
        const uint64_t db = tc::smem_desc_kmajor_noswz(w_smem + ks * 256u, 128u, Kdim * 16u);
        tc::mma_f16_ss(tmem_d, da, db, idesc, ks > 0 ? 1u : 0u);
    }
}
// D[128 x N] = A[128 x Kdim](tensor memory, 8 columns per K=16 step) . W[N x Kdim]^T.  The weight descriptor of K step ks is the first
// one + 16 * ks in its address field (256 bytes >> 4): one add per MMA on the issuing lane instead of a descriptor build.
__device__ __forceinline__ void issue_layer_tmem(uint32_t tmem_a, uint32_t w_smem, uint32_t Kdim, uint32_t N, uint32_t tmem_d) {
    const uint32_t idesc = tc::idesc_f16_f32(kTileRows, N);
    const uint64_t db0 = tc::smem_desc_kmajor_noswz(w_smem, 128u, Kdim * 16u);
    for (uint32_t ks = 0; ks < (Kdim >> 4); ks++) tc::mma_f16_ts(tmem_d, tmem_a + ks * 8u, db0 + (uint64_t)(ks * 16u), idesc, ks > 0 ? 1u : 0u);
}

// one pass of a hidden-layer epilogue: NCOL fp32 accumulator columns of this thread's row -> activation -> fp16 ->
// NCOL/2 operand columns in tensor memory (+ the training copy of the activations in global memory)
template <int NCOL, bool TRAIN>
__device__ __forceinline__ void pipe_epilogue_pass(uint32_t acc, uint32_t opnd, uint32_t act, __half* __restrict__ gdst, bool row_ok) {
    uint32_t v[NCOL];
    if (NCOL == 64) { tc::tmem_ld_x32(acc, v); tc::tmem_ld_x32(acc + 32, v + (NCOL == 64 ? 32 : 0)); }
    else if (NCOL == 32) tc::tmem_ld_x32(acc, v);
    else tc::tmem_ld_x16(acc, v);
    tc::tmem_wait_ld();
#pragma unroll
    for (int h = 0; h < NCOL; h += 32) {
        constexpr int N2 = (NCOL < 32 ? NCOL : 32) / 2;
        uint32_t o[N2];
#pragma unroll
        for (int j = 0; j < N2; j++) o[j] = act_pack2(act, v[h + 2 * j], v[h + 2 * j + 1]);
        if (N2 == 16) tc::tmem_st_x16(opnd + (h >> 1), o); else tc::tmem_st_x8(opnd + (h >> 1), o);
        if (TRAIN && row_ok) {
#pragma unroll
            for (int j = 0; j < N2; j += 4) st_stream_u4(gdst + h + 2 * j, make_uint4(o[j], o[j + 1], o[j + 2], o[j + 3]));
        }
    }
}

template <int WIDTH, bool TRAIN>
__global__ void __launch_bounds__(PipeCfg<WIDTH>::kThreads, PipeCfg<WIDTH>::kMinCtas) mlp_pipe_kernel(const __grid_constant__ CUtensorMap in_map, const __half* __restrict__ weights,
                                                                            __half* __restrict__ outputs, __half* __restrict__ fwd_buf, const uint32_t B,
                                                                            const uint32_t in_dim, const uint32_t num_layers, const uint32_t act,
                                                                            const uint32_t out_act, const uint32_t stages) {
    using Cfg = PipeCfg<WIDTH>;
    constexpr int NCTX = Cfg::kCtx;
    extern __shared__ __align__(1024) uint8_t smem[];
    const PipePlan plan = pipe_plan(in_dim, WIDTH, num_layers, stages);
    uint8_t* w_smem = smem;
    uint8_t* in_smem = smem + plan.in_off;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + plan.misc_off);   // [stages]  TMA -> MMA
    uint64_t* empty_bar = full_bar + kPipeMaxStages;                          // [stages]  MMA (tcgen05.commit) -> TMA
    uint64_t* mma_bar = empty_bar + kPipeMaxStages;                           // [NCTX]    MMA -> epilogue: accumulator ready
    uint64_t* epi_bar = mma_bar + 2;                                          // [NCTX]    epilogue -> MMA: operand written / accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_bar + 2);

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t n_mm = num_layers + 1;                                     // matmuls per tile (ffmlp.cu:574: num_layers - 1 hidden x hidden)

    if (tid == 0) {
        for (uint32_t s = 0; s < stages; s++) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
        for (int c = 0; c < NCTX; c++) { tc::mbar_init(&mma_bar[c], 1); tc::mbar_init(&epi_bar[c], 4); }
        tc::fence_mbar_init();
        tc::prefetch_tensormap(&in_map);
    }
    if (warp == 0) tc::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    {
        const __half* w = weights;
        uint8_t* dst = w_smem;
        load_matrix_kmajor(dst, w, WIDTH, in_dim, tid, Cfg::kThreads);
        w += (size_t)WIDTH * in_dim; dst += (size_t)WIDTH * in_dim * 2;
        for (uint32_t k = 0; k + 1 < num_layers; k++) {
            load_matrix_kmajor(dst, w, WIDTH, WIDTH, tid, Cfg::kThreads);
            w += (size_t)WIDTH * WIDTH; dst += (size_t)WIDTH * WIDTH * 2;
        }
        load_matrix_kmajor(dst, w, 16, WIDTH, tid, Cfg::kThreads);
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t ntiles = ceil_div<uint32_t>(B, kTileRows);

    if (warp < 4 * NCTX) {
        // ============================ epilogue warps: context c, TMEM lane quarter q ====================================
        const uint32_t c = warp >> 2, q = warp & 3u, row = q * 32 + lane;
        const uint32_t acc = tmem_base + c * Cfg::kCtxCols + ((q * 32u) << 16);
        const uint32_t opnd = acc + Cfg::kAOff;
        uint32_t ph = 0;
        for (uint32_t k = c;; k += NCTX) {
            const uint32_t tile = blockIdx.x + k * gridDim.x;
            if (tile >= ntiles) break;
            const size_t row0 = (size_t)tile * kTileRows;
            const bool row_ok = row0 + row < B;
            for (uint32_t st = 0; st < n_mm; st++) {
                tc::mbar_wait(&mma_bar[c], ph); ph ^= 1;
                tc::tc_fence_after_sync();
                if (st + 1 < n_mm) {
                    __half* gdst = TRAIN ? fwd_buf + ((size_t)st * B + row0 + row) * WIDTH : nullptr;
                    constexpr int PASS = WIDTH < 64 ? WIDTH : 64;             // accumulator columns per pass: both loads in flight before the first use
#pragma unroll
                    for (int c0 = 0; c0 < WIDTH; c0 += PASS) pipe_epilogue_pass<PASS, TRAIN>(acc + c0, opnd + (c0 >> 1), act, gdst + c0, row_ok);
                    tc::tmem_wait_st();
                } else {
                    uint32_t v[16];
                    tc::tmem_ld_x16(acc, v);
                    tc::tmem_wait_ld();
                    if (row_ok) {
                        uint4 o0, o1;
                        o0.x = act_pack2(out_act, v[0], v[1]);   o0.y = act_pack2(out_act, v[2], v[3]);
                        o0.z = act_pack2(out_act, v[4], v[5]);   o0.w = act_pack2(out_act, v[6], v[7]);
                        o1.x = act_pack2(out_act, v[8], v[9]);   o1.y = act_pack2(out_act, v[10], v[11]);
                        o1.z = act_pack2(out_act, v[12], v[13]); o1.w = act_pack2(out_act, v[14], v[15]);
                        __half* dst = outputs + (row0 + row) * 16;
                        st_stream_u4(dst, o0);
                        st_stream_u4(dst + 8, o1);
                    }
                }
                // operand written (or accumulator drained): the context's MMA warp may issue the next layer / next tile
                tc::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&epi_bar[c]);
            }
        }
    } else if (warp < 5 * NCTX) {
        // ============================ MMA warp of context c =============================================================
        const uint32_t c = warp - 4 * NCTX;
        const uint32_t acc = tmem_base + c * Cfg::kCtxCols, opnd = acc + Cfg::kAOff;
        const uint32_t w_addr = tc::smem_u32(w_smem);
        const uint32_t w_hidden_addr = w_addr + WIDTH * in_dim * 2u;
        const uint32_t w_last_addr = w_hidden_addr + (num_layers - 1) * WIDTH * WIDTH * 2u;
        uint32_t j = 0;                                                       // waits on epi_bar so far
        for (uint32_t k = c;; k += NCTX) {
            const uint32_t tile = blockIdx.x + k * gridDim.x;
            if (tile >= ntiles) break;
            const uint32_t s = k % stages, use = k / stages;
            tc::mbar_wait(&full_bar[s], use & 1u);                            // the tile has landed
            tc::mbar_wait(&epi_bar[c], (j & 1u) ^ 1u); j++;                   // the previous tile's output has left the accumulator (first: passes)
            tc::tc_fence_after_sync();
            if (tc::elect_one()) {
                issue_layer_chunked(tc::smem_u32(in_smem + s * plan.stage_bytes), w_addr, in_dim, WIDTH, acc);
                tc::mma_commit(&mma_bar[c]);
                tc::mma_commit(&empty_bar[s]);                                // the input stage is free once this MMA has read it
            }
            __syncwarp();
            for (uint32_t st = 1; st < n_mm; st++) {
                tc::mbar_wait(&epi_bar[c], (j & 1u) ^ 1u); j++;
                tc::tc_fence_after_sync();
                if (tc::elect_one()) {
                    const bool last = st + 1 == n_mm;
                    issue_layer_tmem(opnd, last ? w_last_addr : w_hidden_addr + (st - 1) * WIDTH * WIDTH * 2u, WIDTH, last ? 16u : (uint32_t)WIDTH, acc);
                    tc::mma_commit(&mma_bar[c]);
                }
                __syncwarp();
            }
        }
    } else {
        // ============================ TMA producer ======================================================================
        if (tc::elect_one()) {
            const uint32_t chunks = in_dim >> 3;
            for (uint32_t k = 0;; k++) {
                const uint32_t tile = blockIdx.x + k * gridDim.x;
                if (tile >= ntiles) break;
                const uint32_t s = k % stages, use = k / stages;
                tc::mbar_wait_relaxed(&empty_bar[s], (use & 1u) ^ 1u, 64u);
                tc::mbar_arrive_expect_tx(&full_bar[s], plan.stage_bytes);
                uint8_t* dst = in_smem + s * plan.stage_bytes;
                for (uint32_t ch = 0; ch < chunks; ch++) tc::tma_load_2d(dst + ch * kChunkBytes, &in_map, (int32_t)(ch * 8u), (int32_t)(tile * kTileRows), &full_bar[s]);
            }
        }
    }

    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

template <int WIDTH, bool TRAIN>
static int launch_mlp_pipe(const __half* in, const __half* w, __half* out, __half* fwd, uint32_t B, uint32_t in_dim, uint32_t num_layers, uint32_t act,
                           uint32_t out_act, cudaStream_t st) {
    using Cfg = PipeCfg<WIDTH>;
    // input ring depth: as many stages as fit next to the weights, at most 4 (two per tile context)
    uint32_t stages = 4;
    while (stages > 1 && pipe_plan(in_dim, WIDTH, num_layers, stages).total > 100u * 1024u) stages--;
    const PipePlan plan = pipe_plan(in_dim, WIDTH, num_layers, stages);
    NTX_REQUIRE(plan.total <= 227u * 1024u, NTX_ERR_UNSUPPORTED,
                "FullyFusedMLP: %u bytes of shared memory needed (hidden=%d, input_dim=%u, num_layers=%u) exceed the 227 KB of a B200 SM", plan.total,
                WIDTH, in_dim, num_layers);
    auto kern = mlp_pipe_kernel<WIDTH, TRAIN>;
    static int configured_dev[kMaxDevices] = {};
    int& configured_smem = configured_dev[current_device()];
    if ((int)plan.total > configured_smem) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total) != cudaSuccess) {
            cudaGetLastError();
            set_error("FullyFusedMLP: insufficient shared memory available on the GPU.");
            return NTX_ERR_CUDA;
        }
        configured_smem = (int)plan.total;
    }
    CUtensorMap in_map;
    const int rc = make_tensor_map_2d_f16(&in_map, in, in_dim, B, 8, kTileRows);
    if (rc != NTX_OK) return rc;
    const int occ = resident_ctas_per_sm((const void*)kern, Cfg::kThreads, plan.total, Cfg::kTmemCols);
    const uint32_t ntiles = ceil_div<uint32_t>(B, kTileRows);
    const uint32_t grid = std::min<uint32_t>(ntiles, (uint32_t)(occ * device_sm_count()));
    kern<<<grid, Cfg::kThreads, plan.total, st>>>(in_map, w, out, fwd, B, in_dim, num_layers, act, out_act, stages);
    return check_launch("ffmlp_forward");
}

static int mlp_forward_dispatch(const void* inputs, const void* weights, uint32_t B, uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim,
                                uint32_t num_layers, uint32_t activation, uint32_t output_activation, void* forward_buffer, void* outputs, bool train,
                                ntx_stream_t stream) {
    NTX_REQUIRE(inputs && weights && outputs, NTX_ERR_INVALID_ARGUMENT, "ffmlp: null pointer");
    NTX_REQUIRE(!train || forward_buffer, NTX_ERR_INVALID_ARGUMENT, "ffmlp_forward: forward_buffer required");
    NTX_REQUIRE(input_dim > 0 && input_dim % 16 == 0, NTX_ERR_INVALID_ARGUMENT, "FFMLP input_dim should be 16 * m (m > 0), but got %u", input_dim);
    NTX_REQUIRE(output_dim == 16, NTX_ERR_UNSUPPORTED, "FFMLP current only supports output dim <= 16 (padded to 16), but got %u", output_dim);
    // FFMLP (ffmlp.py:115) asserts num_layers >= 2; the kernel itself also handles one hidden layer (2 matmuls), which the
    // tinycudann shim needs for n_hidden_layers == 1 networks (network_curvedfield.py:173)
    NTX_REQUIRE(num_layers >= 1, NTX_ERR_INVALID_ARGUMENT, "FFMLP num_layers must be positive, but got %u", num_layers);
    NTX_REQUIRE(((uintptr_t)inputs & 15) == 0 && ((uintptr_t)weights & 15) == 0 && ((uintptr_t)outputs & 15) == 0, NTX_ERR_INVALID_ARGUMENT,
                "ffmlp: inputs / weights / outputs must be 16-byte aligned");
    if (B == 0) return NTX_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    auto in = static_cast<const __half*>(inputs);
    auto w = static_cast<const __half*>(weights);
    auto out = static_cast<__half*>(outputs);
    auto fwd = static_cast<__half*>(forward_buffer);
#define NTX_MLP(WD)                                                                                                        \
    return train ? launch_mlp_pipe<WD, true>(in, w, out, fwd, B, input_dim, num_layers, activation, output_activation, st) \
                 : launch_mlp_pipe<WD, false>(in, w, out, nullptr, B, input_dim, num_layers, activation, output_activation, st)
    switch (hidden_dim) {
        case 16: NTX_MLP(16);
        case 32: NTX_MLP(32);
        case 64: NTX_MLP(64);
        case 128: NTX_MLP(128);
        case 256: NTX_MLP(256);
        default: set_error("hidden_dim should in [16, 32, 64, 128, 256]"); return NTX_ERR_UNSUPPORTED;
    }
#undef NTX_MLP
}

}  // namespace ntx

using namespace ntx;

extern "C" int ntx_ffmlp_forward(const void* inputs, const void* weights, uint32_t B, uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim,
                                 uint32_t num_layers, uint32_t activation, uint32_t output_activation, void* forward_buffer, void* outputs,
                                 ntx_stream_t stream) {
    return mlp_forward_dispatch(inputs, weights, B, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, forward_buffer, outputs, true, stream);
}

extern "C" int ntx_ffmlp_inference(const void* inputs, const void* weights, uint32_t B, uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim,
                                   uint32_t num_layers, uint32_t activation, uint32_t output_activation, void* inference_buffer, void* outputs,
                                   ntx_stream_t stream) {
    (void)inference_buffer;
    return mlp_forward_dispatch(inputs, weights, B, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, nullptr, outputs, false, stream);
}
