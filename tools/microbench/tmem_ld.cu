// TMEM -> register read bandwidth of tcgen05.ld (32x32b.x32) on one SM, as a function of how many warps read at once.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld tmem_ld.cu && ./tmem_ld
// Prints cycles per x32 load (4 KB per warp) and bytes/clk for 1, 2, 4 (one per sub-partition), 8 warps in ONE CTA, and for
// 4 + 4 warps in two co-resident CTAs of one SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) tmem_ld_kernel(int active_warps, int iters, unsigned long long* out, int wait_each) {
    __shared__ uint32_t slot;
    const uint32_t warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + (((warp & 3u) * 32u) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const unsigned long long t0 = clock64();
    if ((int)warp < active_warps) {
        for (int i = 0; i < iters; i++) {
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                  "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                  "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(base + (uint32_t)((i & 3) * 32)) : "memory");
            if (wait_each || (i & 3) == 3) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= v[0] ^ v[31];
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    const unsigned long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0 + (acc == 0x12345u);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(slot) : "memory");
}

int main() {
    unsigned long long* d;
    cudaMalloc(&d, 16 * sizeof(unsigned long long));
    const int iters = 4096;
    for (int wait_each = 0; wait_each < 2; wait_each++) {
        for (int nw : {1, 2, 4, 8}) {
            tmem_ld_kernel<<<1, 256>>>(nw, iters, d, wait_each);
            unsigned long long h[2] = {0, 0};
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            const double cyc = (double)h[0];
            printf("1 CTA, %d warps, wait %s: %.1f cycles per x32 load per warp, %.1f B/clk per SM\n", nw, wait_each ? "after every load" : "after every 4th",
                   cyc / iters, (double)nw * iters * 4096.0 / cyc);
        }
    }
    // two CTAs on one SM cannot be forced; launch 2 x SM-count CTAs and report the slowest (co-residency is what the occupancy gives)
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned long long* d2;
    cudaMalloc(&d2, 2 * sms * sizeof(unsigned long long));
    tmem_ld_kernel<<<2 * sms, 256>>>(4, iters, d2, 0);
    unsigned long long* h2 = new unsigned long long[2 * sms];
    cudaMemcpy(h2, d2, 2 * sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0;
    for (int i = 0; i < 2 * sms; i++) mx = h2[i] > mx ? h2[i] : mx;
    printf("2 CTAs per SM x 4 warps (grid %d): %.1f cycles per x32 load per warp, %.1f B/clk per SM\n", 2 * sms, (double)mx / iters, 8.0 * iters * 4096.0 / (double)mx);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
