// ref_preprocess_shim.cc -- TEST INFRASTRUCTURE ONLY.  C entry points over the reference's CPU weight packer
// (weightonlykernel/cutlass_kernels/cutlass_preprocessors.cc, compiled unmodified into oracle/_ref/libref_preprocess.so).
//   ref_preprocess_int8      = preprocess_weights(..., is_int4 = false, arch)       (:536-545)
//   ref_symmetric_quantize   = symmetric_quantize<half, half> up to, not including, its final layout step: that step
//                              asks the CUDA runtime for the SM version (:676) and throws where there is no device or
//                              the device is newer than sm_90 (:113-128), AFTER the plain codes and the scales have
//                              been written; the layout step is then run with arch = 80 through ref_preprocess_int8.
#include <cuda_fp16.h>

#include <cstdint>
#include <vector>

#include "cutlass_preprocessors.h"

extern "C" int ref_preprocess_int8(int8_t* out, const int8_t* row_major_kn, size_t K, size_t N, int arch) {
    try {
        fastertransformer::preprocess_weights(out, row_major_kn, K, N, false, arch);
    } catch (...) {
        return 1;
    }
    return 0;
}

extern "C" int ref_symmetric_quantize_half(int8_t* processed, int8_t* unprocessed, void* scales_f16, const void* weight_f16_kn,
                                           size_t K, size_t N) {
    try {
        fastertransformer::symmetric_quantize<half, half>(processed, unprocessed, static_cast<half*>(scales_f16),
                                                          static_cast<const half*>(weight_f16_kn), std::vector<size_t>{K, N},
                                                          fastertransformer::QuantType::INT8_WEIGHT_ONLY);
    } catch (...) {
        return 1;   // the codes and scales are complete; only the arch-dependent layout step did not run
    }
    return 0;
}
