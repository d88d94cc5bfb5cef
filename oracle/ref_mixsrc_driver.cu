// ref_mixsrc_driver.cu -- TEST INFRASTRUCTURE: C entry points over two pieces of the reference's MixQ/src torch extension
// ("mixlib") that SURVEY.md 8f names as the producers / fused epilogues next to the hot path:
//   * generalT5LayerNorm_extract_outliers (MixQ/src/kernel/mix_cuda/layernorm/layernorm.cu:121-198, launcher :289-313):
//     RMSNorm -> outlier extract (+ zeroing) -> per-token INT8 quantise;
//   * cutlass::gemm::device::symmetric::GemmDequantSilu (MixQ/src/kernel/symmetric/gemm/device/gemm_dequantsilu.h, epilogue
//     functor symmetric/epilogue/thread/linear_combination_dequant.h:167-272), instantiated exactly as
//     int8FusedDequantizeSiluCUDA does (MixQ/src/kernel/mix_cuda/cult.cu:2248-2273).
// The reference sources are compiled UNMODIFIED from where they lie (this file #includes layernorm.cu through the -I path
// given by oracle/Makefile; nothing is copied) with the flags of the reference's own setup.py (MixQ/src/kernel/setup.py:
// -O3 --use_fast_math, the __CUDA_NO_HALF_* undefines), retargeted to sm_100a.  layernorm.cu needs the torch headers
// (it launches on at::cuda::getCurrentCUDAStream()); the tests load this library into a process that already holds
// libtorch.  Nothing under mixq_tensorrt_llm_b200/ links or loads this.
#include "mix_cuda/layernorm/layernorm.cu"   // the reference translation unit, in place

#include "symmetric/gemm/device/gemm_dequantsilu.h"

extern "C" {

// invokeGeneralT5LayerNorm_extract_outliers<half> on raw device pointers (what layernorm_forward_cuda_extract_outliers does
// after unpacking its tensors, layernorm.cu:314-346).  Runs on torch's current CUDA stream, like the reference.
int ref_rmsnorm_extract_quant(const void* input, const void* gamma, void* out, float eps, int m, int n, void* outliers,
                              const void* ind, int len_ind, void* out_i8, void* scales) {
    invokeGeneralT5LayerNorm_extract_outliers<half>(static_cast<half*>(out), static_cast<const half*>(input),
                                                    static_cast<const half*>(gamma), eps, m, n, static_cast<half*>(outliers),
                                                    const_cast<int*>(static_cast<const int*>(ind)), len_ind,
                                                    static_cast<int8_t*>(out_i8), static_cast<half*>(scales));
    return static_cast<int>(cudaGetLastError());
}

// int8FusedDequantizeSiluCUDA (cult.cu:2234-2281) on raw pointers: D = fp16(silu(float(acc) * (sc_col * sc_row) + float(y)))
int ref_int8_fused_dequant_silu(const void* A, const void* B, const void* scale_row, const void* scale_col, const void* y,
                                void* D, int M, int N, int K, void* stream) {
    using Gemm = cutlass::gemm::device::symmetric::GemmDequantSilu<int8_t, cutlass::layout::RowMajor, int8_t,
                                                                   cutlass::layout::ColumnMajor, cutlass::half_t,
                                                                   cutlass::layout::RowMajor, int32_t,
                                                                   cutlass::arch::OpClassTensorOp, cutlass::arch::Sm80>;
    Gemm gemmOp;
    using GemmCoord = cutlass::gemm::GemmCoord;
    typename Gemm::Arguments arguments{
        {static_cast<GemmCoord::Index>(M), static_cast<GemmCoord::Index>(N), static_cast<GemmCoord::Index>(K)},
        {static_cast<const int8_t*>(A), K},
        {static_cast<const int8_t*>(B), K},
        {static_cast<cutlass::half_t*>(const_cast<void*>(y)), N},
        {static_cast<cutlass::half_t*>(D), N},
        {static_cast<cutlass::half_t*>(const_cast<void*>(scale_col)), N},
        {static_cast<cutlass::half_t*>(const_cast<void*>(scale_row)), M},
        Gemm::ElementC(1)};
    auto status = gemmOp(arguments, nullptr, static_cast<cudaStream_t>(stream));
    return status == cutlass::Status::kSuccess ? 0 : -1;
}

}  // extern "C"
