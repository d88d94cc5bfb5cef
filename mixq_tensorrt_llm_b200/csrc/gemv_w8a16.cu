// gemv_w8a16.cu -- the M <= 4 branch of MixQPlugin::enqueueImpl (TsinghuaMixQPlugin.cpp:472, 641-647):
//     Out[m, n] = sum_k A[m, k] * fp16( (code[k, n] - 128) * scales[n] )        (W8A16, weight only)
// over the second copy of the weights (`q_weight`, plugin input 5) in the layout EETQ.quant_weights produces
// (weightonlykernel/cutlass_kernels/cutlass_preprocessors.cc:497-533 with the Sm80 layout: K permuted inside
// groups of 16, transposed to [N, K], two output channels interleaved in runs of 64 codes, +128 bias, bytes 1 and 2
// of every four swapped).  It replaces w8_a16_gemm_forward_cuda -> weight_only_batched_gemv
// (weightonlykernel/fpA_intB_gemm_wrapper.cu:29-57, weightOnlyBatchedGemv/kernel.h:285-438).
//
// The reference accumulates in fp16 per thread and in fp32 across threads, so its rounding sequence is part of
// the result.  This kernel keeps that sequence -- which code goes through which fp16 FMA chain, the xor-16/8/2/1
// butterfly, the warp-ordered final sum -- and is bit-identical to the reference kernel (tests/golden/
// ref_gemv_b200.npz).  What is left to choose is how the bytes reach the SM, and on B200 the measured answer is
// occupancy (profiles/r2_gemv_variants.txt): one CTA of 256 threads per group of four output channels (= two contiguous
// rows of 2K bytes of the interleaved layout), 16-byte streaming loads with no L1 allocation, as few registers as the
// chains allow (32 at M = 1: eight CTAs per SM, 64 KB of loads in flight per SM) and the hardware CTA scheduler for the
// tail; activations come through the read-only path and stay in L1.  Persistent CTAs with a register prefetch, a
// statically balanced grid and a shared-memory ring fed by bulk asynchronous copies were all measured slower.  The
// dequantisation interleaves the two channels on the code BYTES (half the permutes of the reference's converter).
// Bytes per call: N*K (weights) + 2*M*K (activations, L2-resident) + 2*M*N; HBM roofline.
#include <cuda_fp16.h>

#include "mixq_internal.h"

namespace mixq {
namespace {

constexpr int kGemvThreads = 256;

__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

// Biased codes -> fp16, exactly: 0x6400 | byte is the fp16 value 1024 + byte; subtracting 1152 leaves byte - 128.
// A stored word holds four codes of ONE channel as bytes [e0, e2, e1, e3]; the FMA chains want, for each k, the pair
// (channel 0, channel 1) in one half2 register (one HFMA2 then advances both channels' chains).  The interleave is done on
// the BYTES, before the conversion: two PRMTs gather (a.e0, b.e0, a.e1, b.e1) and (a.e2, b.e2, a.e3, b.e3) from the two
// channels' words, four more widen them -- half the permutes of converting per channel and transposing the half2 values
// afterwards, and none inside the per-token loop.  The values, and so every rounding, are unchanged.
__device__ __forceinline__ void cvt8(uint32_t a, uint32_t b, __half2 (&out)[4]) {   // out[j] = (a.e_j - 128, b.e_j - 128)
    const uint32_t g0 = __byte_perm(a, b, 0x6240);    // a.e0 b.e0 a.e1 b.e1
    const uint32_t g1 = __byte_perm(a, b, 0x7351);    // a.e2 b.e2 a.e3 b.e3
    const uint32_t h[4] = {__byte_perm(g0, 0x64646464u, 0x5140), __byte_perm(g0, 0x64646464u, 0x5342),
                           __byte_perm(g1, 0x64646464u, 0x5140), __byte_perm(g1, 0x64646464u, 0x5342)};
    const __half2 bias = __halves2half2(__ushort_as_half(0x6480), __ushort_as_half(0x6480));
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = __hsub2(*reinterpret_cast<const __half2*>(&h[j]), bias);
}

// One group of four output channels per CTA; OCC = CTAs per SM the register budget is held to.
template <int M, int OCC>
__global__ void __launch_bounds__(kGemvThreads, OCC)
mixq_gemv_w8a16_kernel(const __half* __restrict__ in, const uint8_t* __restrict__ qweight,
                              const __half* __restrict__ scales, __half* __restrict__ out, int N, int K,
                              const __half* __restrict__ bias, int act) {
    __shared__ float sm[kGemvThreads / 32][M * 4];
    const int t = threadIdx.x;
    const int r = (t >> 2) & 1;
    const int total = 2 * K;
    const int n0 = blockIdx.x * 4;
    const uint8_t* qw = qweight + static_cast<size_t>(blockIdx.x) * 4 * K + t * 16;
    const __half2 zero = __float2half2_rn(0.0f);
    const __half2 s01 = __halves2half2(scales[n0 + r], scales[n0 + 2 + r]);
    __half2 acc[M];
#pragma unroll
    for (int m = 0; m < M; ++m) acc[m] = zero;
#pragma unroll 1
    for (int lk = t * 16; lk < total; lk += 4096, qw += 4096) {
        const uint4 qa = ld_stream_v4(qw), qb = ld_stream_v4(qw + total);
        const int kb = (lk >> 7) * 64 + (lk & 63);
        __half2 wk[16];
        const uint32_t a[4] = {qa.x, qa.y, qa.z, qa.w};
        const uint32_t b[4] = {qb.x, qb.y, qb.z, qb.w};
#pragma unroll
        for (int wd = 0; wd < 4; ++wd) {
            __half2 c[4];
            cvt8(a[wd], b[wd], c);
            wk[2 * wd] = __hfma2(c[0], s01, zero);
            wk[2 * wd + 1] = __hfma2(c[1], s01, zero);
            wk[8 + 2 * wd] = __hfma2(c[2], s01, zero);
            wk[8 + 2 * wd + 1] = __hfma2(c[3], s01, zero);
        }
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const uint4* ap = reinterpret_cast<const uint4*>(in + static_cast<size_t>(m) * K + kb);
            const uint4 x0 = __ldg(ap), x1 = __ldg(ap + 1);
            const uint32_t xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int yy = 0; yy < 8; ++yy) {
                const __half2 x = *reinterpret_cast<const __half2*>(&xs[yy]);
                acc[m] = __hfma2(wk[2 * yy], __low2half2(x), acc[m]);
                acc[m] = __hfma2(wk[2 * yy + 1], __high2half2(x), acc[m]);
            }
        }
    }
    float res[M * 2];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        res[m * 2] = __low2float(acc[m]);
        res[m * 2 + 1] = __high2float(acc[m]);
    }
#pragma unroll
    for (int i = 0; i < M * 2; ++i) {
        res[i] += __shfl_xor_sync(0xffffffffu, res[i], 16);
        res[i] += __shfl_xor_sync(0xffffffffu, res[i], 8);
        res[i] += __shfl_xor_sync(0xffffffffu, res[i], 2);
        res[i] += __shfl_xor_sync(0xffffffffu, res[i], 1);
    }
    const int warp = t >> 5, lane = t & 31;
    if (lane == 0 || lane == 4) {
#pragma unroll
        for (int i = 0; i < M * 2; ++i) sm[warp][i * 2 + (lane >> 2)] = res[i];
    }
    __syncthreads();
    if (t < M * 4) {
        float v = 0.0f;
#pragma unroll
        for (int j = 0; j < kGemvThreads / 32; ++j) v += sm[j][t];
        if (act == MIXQ_ACT_SILU) v = __fdividef(v, 1.0f + __expf(-v));
        __half h = __float2half_rn(v);
        if (bias) h = __float2half_rn(__half2float(h) + __half2float(bias[n0 + (t & 3)]));
        out[static_cast<size_t>(t >> 2) * N + n0 + (t & 3)] = h;
    }
}

}  // namespace

int launch_gemv_w8a16(const void* A, const void* q_weight, const void* scales, void* Out, int64_t M, int64_t N,
                      int64_t K, cudaStream_t stream, const void* bias, int act) {
    if (M == 0 || N == 0) return MIXQ_OK;
    if (!A || !q_weight || !scales || !Out) return set_error(MIXQ_ERR_BAD_ARG, "gemv_w8a16: null pointer");
    if (M < 0 || M > 4) return set_error(MIXQ_ERR_UNSUPPORTED, "gemv_w8a16: the weight-only branch serves 1 <= M <= 4");
    if (N <= 0 || K <= 0 || N > INT32_MAX || K > (INT32_MAX / 2))
        return set_error(MIXQ_ERR_BAD_ARG, "gemv_w8a16: bad dimensions");
    if ((N & 3) != 0 || (K & 63) != 0)
        return set_error(MIXQ_ERR_UNSUPPORTED, "gemv_w8a16: the interleaved layout needs N % 4 == 0 and K % 64 == 0");
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(q_weight)) & 15)
        return set_error(MIXQ_ERR_BAD_ARG, "gemv_w8a16: A and q_weight must be 16-byte aligned");
    if (!device_info().ok) return set_error(MIXQ_ERR_CUDA, "no usable sm_100 device");
    const dim3 grid(static_cast<unsigned>(N / 4)), block(kGemvThreads);
    const __half* a = static_cast<const __half*>(A);
    const uint8_t* q = static_cast<const uint8_t*>(q_weight);
    const __half* s = static_cast<const __half*>(scales);
    const __half* b = static_cast<const __half*>(bias);
    __half* o = static_cast<__half*>(Out);
    const int n = static_cast<int>(N), k = static_cast<int>(K);
    switch (M) {   // CTAs per SM: measured best of 4 / 6 / 8 for each batch (the fp16 chains of M tokens need M-dependent registers)
        case 1: mixq_gemv_w8a16_kernel<1, 8><<<grid, block, 0, stream>>>(a, q, s, o, n, k, b, act); break;
        case 2: mixq_gemv_w8a16_kernel<2, 6><<<grid, block, 0, stream>>>(a, q, s, o, n, k, b, act); break;
        case 3: mixq_gemv_w8a16_kernel<3, 6><<<grid, block, 0, stream>>>(a, q, s, o, n, k, b, act); break;
        default: mixq_gemv_w8a16_kernel<4, 4><<<grid, block, 0, stream>>>(a, q, s, o, n, k, b, act); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "launch gemv_w8a16");
    count_launch();
    return MIXQ_OK;
}

}  // namespace mixq
