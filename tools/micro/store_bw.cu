// Microbenchmark (development tool): how fast can an SM push data from shared memory to HBM/L2?
//   mode 0: TMA bulk store (cp.async.bulk.global.shared::cta) of `bytes` per warp per iteration
//   mode 1: LDS.128 + STG.128 by the warp's lanes
//   mode 2: STG.128 straight from registers
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bw store_bw.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void k(uint8_t *out, int bytes, int iters, int warps_total) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * (blockDim.x >> 5) + warp;
    uint8_t *stage = smem + warp * bytes;
    for (int i = lane * 16; i < bytes; i += 512) *(uint4 *)(stage + i) = make_uint4(i, gw, 3, 4);
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int it = 0; it < iters; it++) {
        uint8_t *dst = out + ((size_t)it * warps_total + gw) * bytes;
        if (MODE == 0) {
            if (lane == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(stage)), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncwarp();
        } else if (MODE == 1) {
            for (int i = lane * 16; i < bytes; i += 512) *(uint4 *)(dst + i) = *(uint4 *)(stage + i);
        } else {
            const uint4 v = make_uint4(it, gw, lane, 7);
            for (int i = lane * 16; i < bytes; i += 512) *(uint4 *)(dst + i) = v;
        }
    }
}

int main(int argc, char **argv) {
    const int sizes[] = {1024, 2560, 5120, 10240};
    uint8_t *out;
    const size_t cap = 2ull << 30;
    cudaMalloc(&out, cap);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wps : {8, 14, 28}) {
        for (int bytes : sizes) {
            const int wpb = 2, blocks = 148 * wps / wpb, warps = blocks * wpb;
            int iters = (int)(cap / ((size_t)warps * bytes));
            if (iters > 64) iters = 64;
            for (int mode = 0; mode < 3; mode++) {
                auto fn = mode == 0 ? k<0> : (mode == 1 ? k<1> : k<2>);
                cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                fn<<<blocks, wpb * 32, wpb * bytes>>>(out, bytes, iters, warps);
                cudaDeviceSynchronize();
                cudaEventRecord(e0);
                fn<<<blocks, wpb * 32, wpb * bytes>>>(out, bytes, iters, warps);
                cudaEventRecord(e1);
                cudaDeviceSynchronize();
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                const double gb = (double)warps * bytes * iters / 1e9;
                printf("warps/SM %2d bytes %5d mode %d: %.1f us, %.0f GB/s (%s) err=%d\n", wps, bytes, mode, ms * 1e3, gb / (ms * 1e-3),
                       mode == 0 ? "TMA bulk" : mode == 1 ? "LDS+STG" : "STG regs", (int)cudaGetLastError());
            }
        }
    }
    return 0;
}
