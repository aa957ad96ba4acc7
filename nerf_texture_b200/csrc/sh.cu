// sh.cu — real spherical-harmonics direction encoder (degree <= 8) for sm_100a.
// Replaces shencoder/src/shencoder.cu (kernel_sh :28, kernel_sh_backward :360).  The basis polynomials live in the
// generated sh_poly.cuh (tools/gen_sh.py); each thread evaluates one direction in registers and the warp then writes
// whole 16-byte vectors (the reference issues one 4-byte store per coefficient).
#include "common.cuh"
#include "sh_poly.cuh"

namespace ntx {

template <int DEG, bool GRAD>
__global__ void __launch_bounds__(128) sh_fwd_kernel(const float* __restrict__ inputs, float* __restrict__ outputs, float* __restrict__ dy_dx, uint32_t B) {
    constexpr int C2 = DEG * DEG;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const float x = inputs[(size_t)b * 3], y = inputs[(size_t)b * 3 + 1], z = inputs[(size_t)b * 3 + 2];
        float o[C2], gx[GRAD ? C2 : 1], gy[GRAD ? C2 : 1], gz[GRAD ? C2 : 1];
        sh_basis<DEG, GRAD>(x, y, z, o, gx, gy, gz);
        float* out = outputs + (size_t)b * C2;
        if (C2 % 4 == 0) {
#pragma unroll
            for (int k = 0; k < C2 / 4; k++) st_stream_u4(out + 4 * k, make_uint4(__float_as_uint(o[4 * k]), __float_as_uint(o[4 * k + 1]), __float_as_uint(o[4 * k + 2]), __float_as_uint(o[4 * k + 3])));
        } else {
#pragma unroll
            for (int k = 0; k < C2; k++) out[k] = o[k];
        }
        if (GRAD) {
            float* d = dy_dx + (size_t)b * 3 * C2;
#pragma unroll
            for (int k = 0; k < C2; k++) { d[k] = gx[k]; d[C2 + k] = gy[k]; d[2 * C2 + k] = gz[k]; }
        }
    }
}

__global__ void __launch_bounds__(256) sh_bwd_kernel(const float* __restrict__ grad, const float* __restrict__ dy_dx, float* __restrict__ grad_inputs,
                                                     uint32_t B, uint32_t D, uint32_t C2) {
    const uint64_t total = (uint64_t)B * D;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t)(t / D), d = (uint32_t)(t - (uint64_t)b * D);
        const float* g = grad + (size_t)b * C2;
        const float* dd = dy_dx + ((size_t)b * D + d) * C2;
        float r = grad_inputs[t];  // accumulated into, like shencoder.cu:377
        for (uint32_t ch = 0; ch < C2; ch++) r = __fmaf_rn(g[ch], dd[ch], r);
        grad_inputs[t] = r;
    }
}

template <int DEG>
static void launch_sh(const float* in, float* out, float* dy_dx, uint32_t B, bool grad, cudaStream_t st) {
    const uint32_t blocks = min(ceil_div<uint32_t>(B, 128), 148u * 16u);
    if (grad) sh_fwd_kernel<DEG, true><<<blocks, 128, 0, st>>>(in, out, dy_dx, B);
    else sh_fwd_kernel<DEG, false><<<blocks, 128, 0, st>>>(in, out, dy_dx, B);
}

}  // namespace ntx

using namespace ntx;

extern "C" int ntx_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t C, int calc_grad_inputs, float* dy_dx,
                                     ntx_stream_t stream) {
    NTX_REQUIRE(D == 3, NTX_ERR_UNSUPPORTED, "SH encoder only support input dim == 3");
    NTX_REQUIRE(C >= 1 && C <= 8, NTX_ERR_UNSUPPORTED, "SH encoder only supports degree in [1, 8]");
    NTX_REQUIRE(inputs && outputs && (!calc_grad_inputs || dy_dx), NTX_ERR_INVALID_ARGUMENT, "sh_encode_forward: null pointer");
    if (B == 0) return NTX_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool g = calc_grad_inputs != 0;
    switch (C) {
        case 1: launch_sh<1>(inputs, outputs, dy_dx, B, g, st); break;
        case 2: launch_sh<2>(inputs, outputs, dy_dx, B, g, st); break;
        case 3: launch_sh<3>(inputs, outputs, dy_dx, B, g, st); break;
        case 4: launch_sh<4>(inputs, outputs, dy_dx, B, g, st); break;
        case 5: launch_sh<5>(inputs, outputs, dy_dx, B, g, st); break;
        case 6: launch_sh<6>(inputs, outputs, dy_dx, B, g, st); break;
        case 7: launch_sh<7>(inputs, outputs, dy_dx, B, g, st); break;
        default: launch_sh<8>(inputs, outputs, dy_dx, B, g, st); break;
    }
    return check_launch("sh_encode_forward");
}

extern "C" int ntx_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t C, const float* dy_dx, float* grad_inputs,
                                      ntx_stream_t stream) {
    (void)inputs;
    NTX_REQUIRE(D == 3 && C >= 1 && C <= 8, NTX_ERR_UNSUPPORTED, "SH encoder only supports input dim 3 and degree in [1, 8]");
    NTX_REQUIRE(grad && dy_dx && grad_inputs, NTX_ERR_INVALID_ARGUMENT, "sh_encode_backward: null pointer");
    if (B == 0) return NTX_OK;
    const uint32_t blocks = (uint32_t)std::min<uint64_t>(ceil_div<uint64_t>((uint64_t)B * D, 256), 148ull * 16ull);
    sh_bwd_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(grad, dy_dx, grad_inputs, B, D, C * C);
    return check_launch("sh_encode_backward");
}
