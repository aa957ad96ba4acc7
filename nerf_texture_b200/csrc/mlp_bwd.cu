// mlp_bwd.cu — backward of the fully-fused MLP (sm_100a).
// Replaces kernel_mlp_fused_backward (ffmlp.cu:411) and the CUTLASS split-K weight-gradient GEMMs issued on side streams
// (ffmlp.cu:804-886, cutlass_matmul.h:459).
//
//  * activation gradients ("dgrad"): one persistent tcgen05 kernel walks the layers backwards per 128-row tile with the
//    TRANSPOSED weight matrices resident in shared memory: dH = (dOut . W) * act'(fwd), all in one kernel like the forward;
//    every layer's dH goes to backward_buffer (the reference's layout) and optionally dL/dinput comes out at the end;
//  * weight gradients ("wgrad"): dW[out,in] = dpre^T . act_in summed over the batch — a skinny GEMM (64x64 output, K = B) on
//    tcgen05 with the batch as the UMMA reduction dimension and one fp32 accumulator in tensor memory per CTA (mlp_wgrad.cuh);
//    per-CTA partial sums are reduced in a fixed order and rounded to fp16 once.  No side streams, no events, no atomics.
#include "mlp_tile.cuh"
#include "mlp_wgrad.cuh"

namespace ntx {

constexpr int kBwdThreads = 256;

struct BwdPlan { uint32_t w_bytes, g_off, h_off, misc_off, total; };
__host__ __device__ inline BwdPlan bwd_plan(uint32_t in_dim, uint32_t hidden, uint32_t num_layers, bool grad_in) {
    BwdPlan p;
    // transposed matrices: W_last^T [hidden x 16], (num_layers-1) x W_hidden^T [hidden x hidden], optionally W_0^T [in_dim x hidden]
    p.w_bytes = 2u * (hidden * 16u + (num_layers - 1) * hidden * hidden + (grad_in ? in_dim * hidden : 0u));
    p.g_off = (p.w_bytes + 127u) & ~127u;
    p.h_off = p.g_off + kTileRows * 16u * 2u;
    p.misc_off = p.h_off + kTileRows * hidden * 2u;
    p.total = p.misc_off + 64u;
    return p;
}

// W [rows_out x cols_in] row-major in global  ->  W^T as an [N = cols_in x K = rows_out] K-major core-matrix tile
__device__ __forceinline__ void load_matrix_transposed_kmajor(uint8_t* smem_dst, const __half* __restrict__ W, uint32_t rows_out, uint32_t cols_in,
                                                              uint32_t tid, uint32_t nthreads) {
    const uint32_t total = rows_out * cols_in;
    for (uint32_t e = tid; e < total; e += nthreads) {
        const uint32_t o = e / cols_in, i = e - o * cols_in;          // coalesced read of W[o][i]
        *reinterpret_cast<__half*>(smem_dst + kmajor_off(i, o, rows_out)) = W[e];
    }
}

// activation backward on the stored post-activation value (utils.h:538-588)
__device__ __forceinline__ float act_backward(uint32_t act, float g, float fwd) {
    constexpr float K_ACT = 10.0f;
    switch (act) {
        case 0: return g * (float)(fwd > 0.0f);
        case 1: return g * fwd;
        case 3: return g * __half2float(__float2half_rn(fwd * __half2float(__float2half_rn(1.0f - fwd))));
        case 4: { const float y = fwd * K_ACT; return g * __half2float(__float2half_rn(y * y / (y * y + 1))); }
        case 5: return g * __half2float(__float2half_rn(1.0f - expf(-fwd * K_ACT)));
        default: return g;
    }
}

// epilogue: TMEM [128 x WIDTH] fp32 -> fp16 -> * act'(fwd tile, read from global) -> fp16 -> smem A tile + backward_buffer
template <int WIDTH>
__device__ __forceinline__ void dgrad_epilogue(uint32_t tmem_base, uint8_t* h_smem, uint32_t act, const __half* __restrict__ fwd_tile,
                                               __half* __restrict__ bwd_tile, uint32_t rows_valid) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q = warp & 3, half_sel = warp >> 2;
    const uint32_t row = q * 32 + lane;
    constexpr int CPT = WIDTH / 2;
    constexpr int CH = CPT < 32 ? CPT : 32;
#pragma unroll
    for (int c0 = 0; c0 < CPT; c0 += CH) {
        const uint32_t col = half_sel * CPT + c0;
        uint32_t v[CH];
        const uint32_t taddr = tmem_base + ((q * 32u) << 16) + col;
        if (CH == 32) tc::tmem_ld_x32(taddr, v);
        else if (CH == 16) tc::tmem_ld_x16(taddr, v);
        else tc::tmem_ld_x8(taddr, v);
        tc::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < CH; j += 8) {
            uint4 f = make_uint4(0u, 0u, 0u, 0u);
            if (row < rows_valid) f = ld_stream_u4(fwd_tile + (size_t)row * WIDTH + col + j);
            const uint32_t fw[4] = {f.x, f.y, f.z, f.w};
            uint32_t ov[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const float2 ff = half2_bits_to_float2(fw[t]);
                // the accumulator is rounded to fp16 first (the reference's result fragment is fp16), then masked / scaled
                const float2 gg = __half22float2(__floats2half2_rn(__uint_as_float(v[j + 2 * t]), __uint_as_float(v[j + 2 * t + 1])));
                ov[t] = float2_to_half2_bits(act_backward(act, gg.x, ff.x), act_backward(act, gg.y, ff.y));
            }
            const uint4 o = make_uint4(ov[0], ov[1], ov[2], ov[3]);
            *reinterpret_cast<uint4*>(h_smem + kmajor_chunk_off(row, (col + j) >> 3, WIDTH)) = o;
            if (row < rows_valid) st_stream_u4(bwd_tile + (size_t)row * WIDTH + col + j, o);
        }
    }
}

template <int WIDTH>
__global__ void __launch_bounds__(kBwdThreads) mlp_dgrad_kernel(const __half* __restrict__ grad, const __half* __restrict__ weights,
                                                                const __half* __restrict__ fwd_buf, __half* __restrict__ bwd_buf,
                                                                __half* __restrict__ grad_inputs, const uint32_t B, const uint32_t in_dim,
                                                                const uint32_t num_layers, const uint32_t act) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const bool grad_in = grad_inputs != nullptr;
    const BwdPlan plan = bwd_plan(in_dim, WIDTH, num_layers, grad_in);
    uint8_t* w_smem = smem;
    uint8_t* g_smem = smem + plan.g_off;
    uint8_t* h_smem = smem + plan.h_off;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + plan.misc_off);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + plan.misc_off + 16);
    constexpr uint32_t TM_COLS = WIDTH < 32 ? 32 : WIDTH;   // in_dim <= WIDTH*? : the dL/dinput layer is issued in <=WIDTH-column pieces

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t n_hidden = num_layers - 1;
    const __half* W0 = weights;
    const __half* Wh = weights + (size_t)WIDTH * in_dim;
    const __half* Wl = Wh + (size_t)n_hidden * WIDTH * WIDTH;

    if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc<TM_COLS>(tmem_slot);
    {
        uint8_t* dst = w_smem;
        load_matrix_transposed_kmajor(dst, Wl, 16, WIDTH, tid, kBwdThreads);                 // W_last^T: [WIDTH x 16]
        dst += WIDTH * 16 * 2;
        for (uint32_t k = 0; k < n_hidden; k++) {                                             // W_hidden[j]^T: [WIDTH x WIDTH]
            load_matrix_transposed_kmajor(dst, Wh + (size_t)k * WIDTH * WIDTH, WIDTH, WIDTH, tid, kBwdThreads);
            dst += WIDTH * WIDTH * 2;
        }
        if (grad_in) load_matrix_transposed_kmajor(dst, W0, WIDTH, in_dim, tid, kBwdThreads);  // W_0^T: [in_dim x WIDTH]
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t w_addr = tc::smem_u32(w_smem), g_addr = tc::smem_u32(g_smem), h_addr = tc::smem_u32(h_smem);
    const uint32_t wh_addr = w_addr + WIDTH * 16 * 2u;
    const uint32_t w0_addr = wh_addr + n_hidden * WIDTH * WIDTH * 2u;
    uint32_t phase = 0;

    const uint32_t ntiles = ceil_div<uint32_t>(B, kTileRows);
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t row0 = (size_t)tile * kTileRows;
        const uint32_t rows_valid = (uint32_t)min((size_t)kTileRows, (size_t)B - row0);
        load_rows_kmajor_stream(g_smem, grad + row0 * 16, kTileRows, rows_valid, 16, tid, kBwdThreads);
        tc::fence_proxy_async_smem();
        __syncthreads();

        // through the output layer: dH = (G . W_last) * act'(fwd[n-1])      (ffmlp.cu:452-500)
        if (warp == 0 && tc::elect_one()) { tc::tc_fence_after_sync(); issue_layer(g_addr, w_addr, 16, WIDTH, tmem_base); tc::mma_commit(bar); }
        tc::mbar_wait(bar, phase); phase ^= 1;
        tc::tc_fence_after_sync();
        dgrad_epilogue<WIDTH>(tmem_base, h_smem, act, fwd_buf + ((size_t)(num_layers - 1) * B + row0) * WIDTH, bwd_buf + ((size_t)0 * B + row0) * WIDTH, rows_valid);
        tc::fence_proxy_async_smem(); tc::tc_fence_before_sync(); __syncthreads();

        // hidden layers, last to first (ffmlp.cu:508-510)
        for (uint32_t k = 0; k < n_hidden; k++) {
            const uint32_t j = n_hidden - 1 - k;   // hidden matrix index
            if (warp == 0 && tc::elect_one()) { tc::tc_fence_after_sync(); issue_layer(h_addr, wh_addr + j * WIDTH * WIDTH * 2u, WIDTH, WIDTH, tmem_base); tc::mma_commit(bar); }
            tc::mbar_wait(bar, phase); phase ^= 1;
            tc::tc_fence_after_sync();
            dgrad_epilogue<WIDTH>(tmem_base, h_smem, act, fwd_buf + ((size_t)j * B + row0) * WIDTH, bwd_buf + ((size_t)(k + 1) * B + row0) * WIDTH, rows_valid);
            tc::fence_proxy_async_smem(); tc::tc_fence_before_sync(); __syncthreads();
        }

        // dL/dinput = dH_0 . W_0     (ffmlp.cu:515-517 / :880-886), issued in pieces of at most TM_COLS output columns
        if (grad_in) {
            for (uint32_t n0 = 0; n0 < in_dim; n0 += TM_COLS) {
                const uint32_t ncols = min(TM_COLS, in_dim - n0);
                if (warp == 0 && tc::elect_one()) { tc::tc_fence_after_sync(); issue_layer(h_addr, w0_addr + (n0 >> 3) * (WIDTH * 16u), WIDTH, ncols, tmem_base); tc::mma_commit(bar); }
                tc::mbar_wait(bar, phase); phase ^= 1;
                tc::tc_fence_after_sync();
                if (warp < 4) {
                    const uint32_t row = warp * 32 + lane;
                    for (uint32_t c = 0; c < ncols; c += 8) {
                        uint32_t v[8];
                        tc::tmem_ld_x8(tmem_base + ((warp * 32u) << 16) + c, v);
                        tc::tmem_wait_ld();
                        if (row < rows_valid) {
                            const uint4 o = make_uint4(float2_to_half2_bits(__uint_as_float(v[0]), __uint_as_float(v[1])), float2_to_half2_bits(__uint_as_float(v[2]), __uint_as_float(v[3])),
                                                       float2_to_half2_bits(__uint_as_float(v[4]), __uint_as_float(v[5])), float2_to_half2_bits(__uint_as_float(v[6]), __uint_as_float(v[7])));
                            st_stream_u4(grad_inputs + (row0 + row) * in_dim + n0 + c, o);
                        }
                    }
                }
                tc::tc_fence_before_sync();
                __syncthreads();
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<TM_COLS>(tmem_base);
}

template <int WIDTH>
static int launch_dgrad(const __half* grad, const __half* w, const __half* fwd, __half* bwd, __half* gi, uint32_t B, uint32_t in_dim, uint32_t num_layers,
                        uint32_t act, cudaStream_t st) {
    const BwdPlan plan = bwd_plan(in_dim, WIDTH, num_layers, gi != nullptr);
    NTX_REQUIRE(plan.total <= 227u * 1024u, NTX_ERR_UNSUPPORTED, "FullyFusedMLP backward: %u bytes of shared memory needed exceed the 227 KB of a B200 SM", plan.total);
    auto kern = mlp_dgrad_kernel<WIDTH>;
    static int configured_dev[kMaxDevices] = {};
    int& configured = configured_dev[current_device()];
    if ((int)plan.total > configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total) != cudaSuccess) {
            cudaGetLastError();
            set_error("FullyFusedMLP: insufficient shared memory available on the GPU.");
            return NTX_ERR_CUDA;
        }
        configured = (int)plan.total;
    }
    constexpr int TM_COLS = WIDTH < 32 ? 32 : WIDTH;
    const int occ = resident_ctas_per_sm((const void*)kern, kBwdThreads, plan.total, TM_COLS);
    const uint32_t ntiles = ceil_div<uint32_t>(B, kTileRows);
    kern<<<std::min<uint32_t>(ntiles, (uint32_t)(occ * device_sm_count())), kBwdThreads, plan.total, st>>>(grad, w, fwd, bwd, gi, B, in_dim, num_layers, act);
    return check_launch("ffmlp_backward");
}

}  // namespace ntx

using namespace ntx;

extern "C" size_t ntx_ffmlp_backward_workspace_bytes(uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers) {
    // scratch of the weight-gradient kernel: kWgMaxParts partial sums of every parameter (all gradient blocks are computed in one launch)
    const size_t params = (size_t)hidden_dim * ((size_t)input_dim + (size_t)hidden_dim * (num_layers - 1) + output_dim);
    return sizeof(float) * (size_t)kWgMaxParts * params;
}

extern "C" int ntx_ffmlp_backward(const void* grad_, const void* inputs_, const void* weights_, const void* forward_buffer_, uint32_t B, uint32_t input_dim,
                                  uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers, uint32_t activation, uint32_t output_activation,
                                  int calc_grad_inputs, void* backward_buffer_, void* grad_inputs_, void* grad_weights_, void* workspace_, ntx_stream_t stream) {
    (void)output_activation;   // the reference ignores it in the backward as well (ffmlp.cu:781)
    NTX_REQUIRE(grad_ && inputs_ && weights_ && forward_buffer_ && backward_buffer_ && grad_weights_, NTX_ERR_INVALID_ARGUMENT, "ffmlp_backward: null pointer");
    NTX_REQUIRE(workspace_, NTX_ERR_WORKSPACE, "ffmlp_backward: workspace of ntx_ffmlp_backward_workspace_bytes() bytes required");
    NTX_REQUIRE(!calc_grad_inputs || grad_inputs_, NTX_ERR_INVALID_ARGUMENT, "ffmlp_backward: grad_inputs required");
    NTX_REQUIRE(input_dim > 0 && input_dim % 16 == 0, NTX_ERR_INVALID_ARGUMENT, "FFMLP input_dim should be 16 * m (m > 0), but got %u", input_dim);
    NTX_REQUIRE(output_dim == 16, NTX_ERR_UNSUPPORTED, "FFMLP current only supports output dim <= 16 (padded to 16), but got %u", output_dim);
    NTX_REQUIRE(num_layers >= 1, NTX_ERR_INVALID_ARGUMENT, "FFMLP num_layers must be positive");
    if (B == 0) return NTX_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    auto grad = static_cast<const __half*>(grad_);
    auto inputs = static_cast<const __half*>(inputs_);
    auto w = static_cast<const __half*>(weights_);
    auto fwd = static_cast<const __half*>(forward_buffer_);
    auto bwd = static_cast<__half*>(backward_buffer_);
    auto gi = calc_grad_inputs ? static_cast<__half*>(grad_inputs_) : nullptr;
    auto gw = static_cast<__half*>(grad_weights_);
    auto ws = static_cast<float*>(workspace_);
    int rc;
    switch (hidden_dim) {
        case 16: rc = launch_dgrad<16>(grad, w, fwd, bwd, gi, B, input_dim, num_layers, activation, st); break;
        case 32: rc = launch_dgrad<32>(grad, w, fwd, bwd, gi, B, input_dim, num_layers, activation, st); break;
        case 64: rc = launch_dgrad<64>(grad, w, fwd, bwd, gi, B, input_dim, num_layers, activation, st); break;
        case 128: rc = launch_dgrad<128>(grad, w, fwd, bwd, gi, B, input_dim, num_layers, activation, st); break;
        case 256: rc = launch_dgrad<256>(grad, w, fwd, bwd, gi, B, input_dim, num_layers, activation, st); break;
        default: set_error("hidden_dim should in [16, 32, 64, 128, 256]"); return NTX_ERR_UNSUPPORTED;
    }
    if (rc != NTX_OK) return rc;
    // weight gradients (layouts as ffmlp.cu:742-748).  dpre[j] = backward_buffer[n-1-j], n = num_layers.  Every [<=128 x <=256]
    // gradient block is a job; all jobs run in ONE tcgen05 launch + ONE reduce.
    const uint32_t n = num_layers, Hd = hidden_dim;
    __half* gw0 = gw;                                            // [hidden, in]
    __half* gwh = gw + (size_t)Hd * input_dim;                   // (n-1) x [hidden, hidden]
    __half* gwl = gwh + (size_t)(n - 1) * Hd * Hd;               // [16, hidden]
    WgJobs jobs;
    uint32_t nj = 0;
    auto add_job = [&](const __half* P, uint32_t ldp, uint32_t p, const __half* Q, uint32_t ldq, uint32_t q, bool transpose, __half* dst, uint32_t ld_dst) {
        if (nj == kWgMaxJobs) return false;
        WgJob& j = jobs.j[nj++];
        j.P = P; j.Q = Q; j.partials = nullptr; j.dst = dst; j.ldp = ldp; j.p = p; j.ldq = ldq; j.q = q; j.ld_dst = ld_dst; j.transpose = transpose ? 1u : 0u;
        j.tmem_cols = 0; j.pad = 0;
        return true;
    };
    bool fits = true;
    for (uint32_t r0 = 0; r0 < Hd; r0 += 128) {
        const uint32_t pr = std::min<uint32_t>(128u, Hd - r0);
        // output layer: dW_last[16 x hidden] = grad^T . act[n-1], computed as (act[n-1][:, r0:r0+pr])^T . grad and stored transposed
        fits = fits && add_job(fwd + (size_t)(n - 1) * B * Hd + r0, Hd, pr, grad, 16, 16, true, gwl + r0, Hd);
        // hidden layers: dW_h[j][hidden x hidden] = dpre[j+1]^T . act[j]
        for (uint32_t j = 0; j + 1 < n; j++)
            for (uint32_t c0 = 0; c0 < Hd; c0 += 256)
                fits = fits && add_job(bwd + (size_t)(n - 2 - j) * B * Hd + r0, Hd, pr, fwd + (size_t)j * B * Hd + c0, Hd, std::min<uint32_t>(256u, Hd - c0), false,
                                       gwh + (size_t)j * Hd * Hd + (size_t)r0 * Hd + c0, Hd);
        // first layer: dW_0[hidden x in] = dpre[0]^T . inputs
        for (uint32_t c0 = 0; c0 < input_dim; c0 += 256)
            fits = fits && add_job(bwd + (size_t)(n - 1) * B * Hd + r0, Hd, pr, inputs + c0, input_dim, std::min<uint32_t>(256u, input_dim - c0), false,
                                   gw0 + (size_t)r0 * input_dim + c0, input_dim);
    }
    NTX_REQUIRE(fits, NTX_ERR_UNSUPPORTED, "ffmlp_backward: more than %d weight-gradient blocks (hidden_dim %u, input_dim %u, num_layers %u)", kWgMaxJobs, Hd, input_dim, n);
    return launch_wgrad_jobs(jobs, nj, ws, B, st);
}
