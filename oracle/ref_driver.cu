// ref_driver.cu -- TEST INFRASTRUCTURE: replays the reference plugin's M>4 launch sequence
// (TsinghuaMixQPlugin.cpp:406-421 workspace carve, :518-532 launches) using the REFERENCE'S OWN
// kernels.  Those kernels are compiled, unmodified, from /root/reference/kernel/i8gemm.cu into
// oracle/_ref/libref_i8gemm.so (see oracle/Makefile); this file only declares them through the
// reference's header (found with -I /root/reference/kernel at build time, never copied) and
// calls them plus cuBLAS in the order enqueueImpl does.  Used as
//   * the GPU-side oracle the -m gpu parity tests compare against, and
//   * the same-box baseline bench.py reports as `ref_gpu`.
// Nothing under mixq_tensorrt_llm_b200/ links or loads this.
#include <cublas_v2.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "int8FusedDequantizeCUDA.h"  // reference header: int8quant, ExtractOutliersAndSetToZeros, int8FusedDequantizeCUDA

namespace {
cublasHandle_t g_handle = nullptr;

constexpr uintptr_t kCudaMemAlign = 128;  // TsinghuaMixQPlugin.cpp:204
int8_t* next_ws(int8_t* ptr, uintptr_t prev) {  // nextWorkspacePtr, :206-215
    uintptr_t a = reinterpret_cast<uintptr_t>(ptr) + prev;
    if (a % kCudaMemAlign) a += kCudaMemAlign - a % kCudaMemAlign;
    return reinterpret_cast<int8_t*>(a);
}

__global__ void rcp_table_kernel(uint32_t* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 65536u) {
        const float f = __half2float(__ushort_as_half(static_cast<unsigned short>(i)));
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f));
        out[i] = __float_as_uint(r);
    }
}

// device __hdiv over arbitrary (a, b) pairs -- lets the tests check the oracle's emulation directly
__global__ void hdiv_kernel(const __half* a, const __half* b, __half* q, int* qi, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const __half v = __hdiv(a[i], b[i]);
        q[i] = v;
        qi[i] = __half2int_rn(v);
    }
}
}  // namespace

extern "C" {

int ref_init() {
    if (!g_handle) return cublasCreate(&g_handle) == CUBLAS_STATUS_SUCCESS ? 0 : -1;  // MixQPlugin::initialize, :792-799
    return 0;
}

size_t ref_workspace_size(int64_t M, int64_t N, int64_t K) {
    (void)N;
    return static_cast<size_t>(M) * K + 2 * M + 256 * M + 3 * 128 + 128;
}

// int8quant alone (kernel/i8gemm.cu:139-150)
int ref_int8quant(const void* A, int M, int K, void* q, void* sa, void* stream) {
    int8quant(M, K, static_cast<const half*>(A), static_cast<int8_t*>(q), static_cast<half*>(sa),
              static_cast<cudaStream_t>(stream));
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ExtractOutliersAndSetToZeros alone (kernel/i8gemm.cu:225-244)
int ref_extract(const void* A, int M, int K, const void* ind, int n_ind, void* fp_A, void* stream) {
    ExtractOutliersAndSetToZeros(M, K, static_cast<const half*>(A), static_cast<half*>(fp_A),
                                 static_cast<const int*>(ind), n_ind, static_cast<cudaStream_t>(stream));
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// int8FusedDequantizeCUDA alone, C = D = Out (kernel/i8gemm.cu:151-194)
int ref_fused_dequant(const void* A8, const void* W8, const void* sa, const void* sb, void* Out, int M, int N, int K,
                      void* workspace, void* stream) {
    int8FusedDequantizeCUDA(static_cast<const int8_t*>(A8), static_cast<const int8_t*>(W8),
                            static_cast<const half*>(sa), static_cast<const half*>(sb), static_cast<half*>(Out),
                            static_cast<half*>(Out), M, N, K, static_cast<char*>(workspace),
                            static_cast<cudaStream_t>(stream));
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// The whole M>4 branch, in the reference's order.
int ref_enqueue(const void* A_, const void* W_, const void* scale_b_, const void* fp_weight_, const void* ind_,
                void* Out_, int M, int N, int K, void* workspace, void* stream_) {
    if (ref_init()) return -1;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    half* Out = static_cast<half*>(Out_);
    half* actPtr = static_cast<half*>(workspace);
    int8_t* int8_out = next_ws(reinterpret_cast<int8_t*>(actPtr), 0);                                    // :409
    half* scale_a = reinterpret_cast<half*>(next_ws(int8_out, sizeof(int8_t) * (size_t)M * K));          // :411-413
    half* fp_activation = reinterpret_cast<half*>(next_ws(reinterpret_cast<int8_t*>(scale_a), sizeof(half) * M));  // :416
    const half* A = static_cast<const half*>(A_);
    const int8_t* W = static_cast<const int8_t*>(W_);
    const half* scale_b = static_cast<const half*>(scale_b_);
    const half* fp_weight = static_cast<const half*>(fp_weight_);
    const int* ind = static_cast<const int*>(ind_);

    const int num_ind = 128;                                                                             // :518
    ExtractOutliersAndSetToZeros(M, K, A, fp_activation, ind, num_ind, stream);                          // :519
    cublasSetStream(g_handle, stream);                                                                   // :520
    {                                                                                                    // gemmfp16, :122-161
        const float alpha = 1.0f, beta = 0.0f;
        cublasStatus_t st = cublasGemmEx(g_handle, CUBLAS_OP_T, CUBLAS_OP_N, N, M, num_ind, &alpha, fp_weight,
                                         CUDA_R_16F, num_ind, fp_activation, CUDA_R_16F, num_ind, &beta, Out,
                                         CUDA_R_16F, N, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT_TENSOR_OP);
        if (st != CUBLAS_STATUS_SUCCESS) return -2;
    }
    int8quant(M, K, A, int8_out, scale_a, stream);                                                       // :522
    int8FusedDequantizeCUDA(int8_out, W, scale_a, scale_b, Out, Out, M, N, K,                            // :529-532
                            reinterpret_cast<char*>(workspace), stream);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

// rcp.approx.ftz.f32 over all 65536 fp16 inputs -> tests/golden/rcp_approx_f16.bin
int ref_rcp_table(void* out_dev, void* stream) {
    rcp_table_kernel<<<256, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<uint32_t*>(out_dev));
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int ref_hdiv(const void* a, const void* b, void* q, void* qi, int n, void* stream) {
    hdiv_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __half*>(a), static_cast<const __half*>(b), static_cast<__half*>(q), static_cast<int*>(qi), n);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
