// Micro-benchmarks for epilogue building blocks on sm_100a (debug aid, not part of the library):
// tcgen05.ld throughput by shape and warp count, fence.proxy.async cost, STS/LDS tile traffic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb tools/microbench_tmem.cu && /tmp/mb
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ld_tmem(uint32_t taddr, uint32_t* v);
template <>
__device__ __forceinline__ void ld_tmem<32>(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void ld_tmem<16>(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void ld_tmem<8>(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}

// mode 0: ld + wait each time (latency-bound); mode 1: 4 loads in flight then wait
template <int X>
__global__ void k_ldtm(long long* out, int iters, int mode) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tptr + (((uint32_t)(warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    if (mode == 0) {
        for (int i = 0; i < iters; ++i) {
            uint32_t v[X];
            ld_tmem<X>(base + ((i * X) & 255), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < X; ++j) acc += v[j];
        }
    } else {
        for (int i = 0; i < iters; i += 4) {
            uint32_t v[4][X];
#pragma unroll
            for (int u = 0; u < 4; ++u) ld_tmem<X>(base + (((i + u) * X) & 255), v[u]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < X; ++j) acc += v[u][j];
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (acc == 0x12345678u) out[1] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(512));
}

__global__ void k_fence(long long* out, int iters) {
    __shared__ uint4 buf[256];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        buf[threadIdx.x] = make_uint4(i, i, i, i);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
}

int main() {
    long long* d;
    cudaMalloc(&d, 64);
    long long h[2];
    const int iters = 4096;
    for (int warps : {4, 8}) {
        for (int mode : {0, 1}) {
            auto run = [&](auto kern, int X) {
                kern<<<1, warps * 32>>>(d, iters, mode);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                double bytes = (double)iters * X * 4 * 32 * warps;
                printf("ldtm 32x32b.x%-2d warps=%d mode=%d: %8lld cyc, %.1f cyc/ld/warp, %.1f B/clk/SM (%s)\n", X, warps, mode, h[0],
                       (double)h[0] / iters, bytes / h[0], cudaGetErrorString(e));
            };
            run(k_ldtm<32>, 32);
            run(k_ldtm<16>, 16);
            run(k_ldtm<8>, 8);
        }
    }
    k_fence<<<1, 256>>>(d, 1000);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("sts.128 + fence.proxy.async + syncwarp: %.1f cyc/iter (8 warps)\n", h[0] / 1000.0);
    return 0;
}
