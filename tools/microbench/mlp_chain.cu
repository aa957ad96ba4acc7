// Latency anatomy of one layer of the TMEM-resident MLP chain (the dependent loop of mlp_pipe_kernel / ngp_field_kernel):
//   MMA warp:      wait(epi_bar) -> fence -> 4 x tcgen05.mma (A from TMEM, B = weights in smem) -> commit(mma_bar)
//   epilogue warp: wait(mma_bar) -> fence -> 2 x tcgen05.ld.x32 -> wait::ld -> 32 x cvt.relu.f16x2 -> 2 x tcgen05.st.x16 -> wait::st -> fence -> arrive(epi_bar)
// One CTA, nothing else on the SM: the numbers are the chain's floor; what the full kernels add is contention.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../nerf_texture_b200/csrc -o mlp_chain mlp_chain.cu && ./mlp_chain
#include <cstdio>
#include "tc05.cuh"
#include <cuda_fp16.h>
using namespace ntx;

__global__ void __launch_bounds__(160) chain_kernel(int iters, unsigned long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_smem = smem;                                    // [64 x 64] fp16, K-major core-matrix layout (contents irrelevant)
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(smem + 8192);
    uint64_t* epi_bar = mma_bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(epi_bar + 1);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (uint32_t i = tid; i < 2048; i += blockDim.x) reinterpret_cast<uint32_t*>(w_smem)[i] = 0x2c002c00u;   // 0.0625 halves
    if (tid == 0) { tc::mbar_init(mma_bar, 1); tc::mbar_init(epi_bar, 4); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc<128>(slot);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = *slot;
    unsigned long long acc[6] = {0, 0, 0, 0, 0, 0};
    if (warp < 4) {
        const uint32_t accum = tmem + ((warp * 32u) << 16), opnd = accum + 64;
        uint32_t ph = 0;
        for (int it = 0; it < iters; it++) {
            unsigned long long t0 = clock64();
            tc::mbar_wait(mma_bar, ph); ph ^= 1;
            tc::tc_fence_after_sync();
            unsigned long long t1 = clock64();
            uint32_t v[64];
            tc::tmem_ld_x32(accum, v); tc::tmem_ld_x32(accum + 32, v + 32);
            tc::tmem_wait_ld();
            unsigned long long t2 = clock64();
            uint32_t o[32];
#pragma unroll
            for (int j = 0; j < 32; j++) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(o[j]) : "f"(__uint_as_float(v[2 * j + 1])), "f"(__uint_as_float(v[2 * j])));
            tc::tmem_st_x16(opnd, o); tc::tmem_st_x16(opnd + 16, o + 16);
            unsigned long long t3 = clock64();
            tc::tmem_wait_st();
            unsigned long long t4 = clock64();
            tc::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(epi_bar);
            unsigned long long t5 = clock64();
            acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2; acc[3] += t4 - t3; acc[4] += t5 - t4;
        }
        if (tid == 0) for (int i = 0; i < 5; i++) out[i] = acc[i];
    } else {
        const uint32_t w_addr = tc::smem_u32(w_smem);
        const uint32_t idesc = tc::idesc_f16_f32(128, 64);
        for (int it = 0; it < iters; it++) {
            unsigned long long t0 = clock64();
            tc::mbar_wait(epi_bar, (it & 1u) ^ 1u);
            tc::tc_fence_after_sync();
            unsigned long long t1 = clock64();
            if (tc::elect_one()) {
                for (uint32_t ks = 0; ks < 4; ks++) {
                    const uint64_t db = tc::smem_desc_kmajor_noswz(w_addr + ks * 256u, 128u, 64 * 16u);
                    tc::mma_f16_ts(tmem, tmem + 64 + ks * 8u, db, idesc, ks > 0 ? 1u : 0u);
                }
                tc::mma_commit(mma_bar);
            }
            __syncwarp();
            unsigned long long t2 = clock64();
            acc[0] += t1 - t0; acc[1] += t2 - t1;
        }
        if (lane == 0) { out[8] = acc[0]; out[9] = acc[1]; }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<128>(tmem);
}

int main() {
    unsigned long long* d;
    cudaMalloc(&d, 16 * sizeof(unsigned long long));
    cudaMemset(d, 0, 16 * sizeof(unsigned long long));
    const int iters = 2000;
    cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    chain_kernel<<<1, 160, 16384>>>(iters, d);
    cudaEventRecord(e0);
    chain_kernel<<<1, 160, 16384>>>(iters, d);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[5] = {"wait mma_bar (MMA + commit + wake-up)", "2 x tcgen05.ld.x32 + wait::ld", "32 cvt + 2 x tcgen05.st.x16 (issue)", "wait::st", "fence + syncwarp + arrive"};
    double sum = 0;
    for (int i = 0; i < 5; i++) { printf("epilogue warp: %-42s %7.1f cycles\n", names[i], (double)h[i] / iters); sum += (double)h[i] / iters; }
    printf("epilogue warp: one layer of the chain                      %7.1f cycles  (kernel: %.1f cycles per layer at the event clock)\n", sum, ms * 1e-3 * 1.965e9 / iters);
    printf("MMA warp: wait epi_bar %.1f cycles, fence + issue 4 MMAs + commit %.1f cycles\n", (double)h[8] / iters, (double)h[9] / iters);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
