// ref_gemv_driver.cu -- TEST INFRASTRUCTURE ONLY (GPU-side checker for the M <= 4 branch).
// Calls the reference's own weight-only GEMV kernels, compiled unmodified from
// /root/reference/weightonlykernel/weightOnlyBatchedGemv/*.cu into oracle/_ref/libref_gemv.so, with exactly the
// parameters fpA_intB_gemm_wrapper.cu:46-56 builds for w8_a16_gemm_forward_cuda (Int8b, PerChannel, FP16, Identity,
// no zeros / bias / act_scale).  The CUTLASS fpA_intB GEMM behind m > 4 in that wrapper is not on the plugin's
// path (the plugin only calls it for M <= 4, TsinghuaMixQPlugin.cpp:472) and is not built.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "weightOnlyBatchedGemv/kernelLauncher.h"

extern "C" int ref_w8a16_gemv(const void* in, const void* qweight, const void* scale, void* out, int m, int n, int k,
                              void* stream) {
    using namespace tensorrt_llm::kernels;
    WeightOnlyParams params{reinterpret_cast<const uint8_t*>(qweight), scale, nullptr, in, nullptr, nullptr, out, m, n, k, 0,
                            WeightOnlyQuantType::Int8b, WeightOnlyType::PerChannel,
                            WeightOnlyActivationFunctionType::Identity, WeightOnlyActivationType::FP16};
    weight_only_batched_gemv_launcher(params, static_cast<cudaStream_t>(stream));
    return static_cast<int>(cudaGetLastError());
}
