// microbench_ingest.cu -- how fast can TMA land operand tiles in an SM's shared memory, and does
// cp.async.bulk.tensor ... .multicast::cluster change it?  (Decides whether sharing the W K-block between the CTA pairs
// of a cluster can move the decode GEMM: DESIGN.md 4, "operand ingest".)
//
// Every CTA runs a producer thread (issues 128-row x 128-byte boxes, SWIZZLE_128B, into an S-slot ring) and a consumer
// thread (waits for the slot's bytes, hands the slot back); there is no math.  Modes:
//   0  unicast, every CTA streams its OWN rows            (distinct data)
//   1  unicast, the C CTAs of a cluster stream the SAME rows (what two CTA pairs sharing a W tile do today)
//   2  multicast: each CTA of the cluster loads 1/C of the box and multicasts it to all C CTAs
// Reported: bytes LANDED per SM per microsecond (mode 2: each SM lands a full box per step but requests 1/C of it).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_ingest tools/microbench_ingest.cu -lcuda
//   tools/microbench_ingest [buffer_MB=64] [steps=2048] [boxes_per_slot=2]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mixq_tensorrt_llm_b200/csrc/ptx.cuh"

using namespace mixq;

constexpr int kRows = 128, kRowBytes = 128, kBoxBytes = kRows * kRowBytes;   // 16 KB
constexpr int kPitch = 4096;                                               // bytes per tensor row (a K = 4096 operand)
constexpr int kSlots = 6;
constexpr int kMaxBoxes = 2;   // boxes per ring slot (the GEMM stages an A box and a W box per slot)

__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        :
        : "r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}

__global__ void __launch_bounds__(64, 1)
ingest_kernel(const __grid_constant__ CUtensorMap tm_full, const __grid_constant__ CUtensorMap tm_slice, int mode, int steps,
              int n_regions, int nbox, unsigned long long* t_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + kSlots * kMaxBoxes * kBoxBytes);
    uint64_t* empty_bar = full_bar + kSlots;
    const uint32_t C = cluster_nctarank();
    const uint32_t rank = ptx::cluster_ctarank();
    const int cluster_id = blockIdx.x / C, n_clusters = gridDim.x / C;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], mode == 2 ? C : 1);
        }
        ptx::fence_barrier_init();
    }
    if (C > 1) ptx::cluster_sync(); else __syncthreads();
    unsigned long long t0 = 0;
    if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const int kb_per_row = kPitch / kRowBytes;   // 32 boxes along a row band
    if (threadIdx.x == 0) {
        // producer
        for (int i = 0; i < steps; ++i) {
            const int s = i % kSlots;
            const uint32_t ph = (i / kSlots) & 1;
            ptx::mbar_wait(&empty_bar[s], ph ^ 1);
            ptx::mbar_arrive_expect_tx(&full_bar[s], kBoxBytes * nbox);
            const int band = i / kb_per_row, kb = i % kb_per_row;
            const int who = mode == 0 ? blockIdx.x : cluster_id;
            const int nwho = mode == 0 ? gridDim.x : n_clusters;
            const int region = (who + band * nwho) % n_regions;
            for (int b = 0; b < nbox; ++b) {
                uint8_t* dst = ring + (s * kMaxBoxes + b) * kBoxBytes;
                const int reg_b = (region + b * (n_regions / 2)) % n_regions;   // the second box streams another row band
                if (mode == 2) {
                    const int slice = kRows / C;
                    tma_load_2d_mc(dst + rank * slice * kRowBytes, &tm_slice, &full_bar[s], kb * kRowBytes, reg_b * kRows + rank * slice,
                                   static_cast<uint16_t>((1u << C) - 1));
                } else {
                    ptx::tma_load_2d(dst, &tm_full, &full_bar[s], kb * kRowBytes, reg_b * kRows, ptx::kEvictNormal);
                }
            }
        }
    } else if (threadIdx.x == 32) {
        // consumer
        for (int i = 0; i < steps; ++i) {
            const int s = i % kSlots;
            const uint32_t ph = (i / kSlots) & 1;
            ptx::mbar_wait(&full_bar[s], ph);
            if (mode == 2) {
                for (uint32_t c = 0; c < C; ++c) ptx::mbar_arrive_cluster(&empty_bar[s], c);
            } else {
                ptx::mbar_arrive(&empty_bar[s]);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        t_out[blockIdx.x] = t1 - t0;
    }
    if (C > 1) ptx::cluster_sync();
}

static CUtensorMap make_map(void* base, uint64_t rows, uint32_t box_rows) {
    using Fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                            CUtensorMapFloatOOBfill);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t gdim[2] = {kPitch, rows};
    cuuint64_t gstride[1] = {kPitch};
    cuuint32_t box[2] = {kRowBytes, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<Fn>(p)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("encode failed %d\n", (int)r);
        exit(1);
    }
    return m;
}

int main(int argc, char** argv) {
    const size_t mb = argc > 1 ? atoi(argv[1]) : 64;
    const int steps = argc > 2 ? atoi(argv[2]) : 2048;
    const uint64_t rows = mb * 1024 * 1024 / kPitch;
    const int n_regions = static_cast<int>(rows / kRows);
    void* buf;
    cudaMalloc(&buf, rows * kPitch);
    cudaMemset(buf, 1, rows * kPitch);
    unsigned long long* t_dev;
    cudaMalloc(&t_dev, 256 * sizeof(unsigned long long));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int nbox = argc > 3 ? atoi(argv[3]) : 2;
    const size_t smem = 1024 + kSlots * kMaxBoxes * kBoxBytes + 2 * kSlots * 8;
    cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    printf("buffer %zu MB (%s L2), %d steps of %d x 16 KB per CTA, ring %d slots, %d SMs\n", mb, mb <= 96 ? "fits" : "exceeds", steps, nbox, kSlots,
           prop.multiProcessorCount);
    printf("%-5s %-8s %-6s %12s %14s %14s\n", "mode", "cluster", "CTAs", "us", "GB/s per SM", "TB/s landed");
    for (int mode = 0; mode < 3; ++mode) {
        for (int C : {1, 2, 4, 8}) {
            if (mode == 0 && C != 1) continue;
            if (mode != 0 && C == 1) continue;
            for (int ctas : {8, 64, 148}) {
                const int grid = ctas / C * C;
                if (grid == 0) continue;
                CUtensorMap tm_full = make_map(buf, rows, kRows);
                CUtensorMap tm_slice = make_map(buf, rows, kRows / C);
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(grid);
                cfg.blockDim = dim3(64);
                cfg.dynamicSmemBytes = smem;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = C;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                float best = 1e30f;
                for (int rep = 0; rep < 4; ++rep) {
                    cudaEvent_t e0, e1;
                    cudaEventCreate(&e0);
                    cudaEventCreate(&e1);
                    cudaEventRecord(e0);
                    cudaError_t e = cudaLaunchKernelEx(&cfg, ingest_kernel, tm_full, tm_slice, mode, steps, n_regions, nbox, t_dev);
                    cudaEventRecord(e1);
                    cudaError_t e2 = cudaDeviceSynchronize();
                    if (e != cudaSuccess || e2 != cudaSuccess) {
                        printf("mode %d C %d grid %d: %s / %s\n", mode, C, grid, cudaGetErrorString(e), cudaGetErrorString(e2));
                        return 1;
                    }
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (rep > 0 && ms < best) best = ms;
                }
                std::vector<unsigned long long> t(grid);
                cudaMemcpy(t.data(), t_dev, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
                unsigned long long mx = 0;
                for (auto v : t) mx = v > mx ? v : mx;
                const double us = mx / 1e3;   // in-kernel time of the slowest CTA (excludes launch overhead)
                const double per_sm = static_cast<double>(steps) * kBoxBytes * nbox / us / 1e3;
                printf("%-5d %-8d %-6d %12.1f %14.1f %14.2f   (event %.1f us)\n", mode, C, grid, us, per_sm, per_sm * grid / 1e3, best * 1e3);
            }
        }
    }
    return 0;
}
