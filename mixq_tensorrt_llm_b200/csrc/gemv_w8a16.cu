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
// ref_gemv_b200.npz); everything else is laid out for an HBM-bound stream on B200: a CTA of 256 threads owns four
// output channels at a time (= two contiguous rows of 2K bytes of the interleaved layout), CTAs are persistent
// (3 per SM) and stride over the channel groups, the 16-byte streaming weight loads (no L1 allocation) of the next
// step -- of the next group, at a group's end -- are in flight while the current ones go through the FMA chains,
// activations come through the read-only path and stay in L1.
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

// four biased codes (bytes [e0, e2, e1, e3]) -> (e0 - 128, e1 - 128), (e2 - 128, e3 - 128) as fp16, exactly:
// 0x6400 | byte is the fp16 value 1024 + byte; subtracting 1152 leaves byte - 128.
__device__ __forceinline__ void cvt4(uint32_t w, __half2& lo, __half2& hi) {
    const uint32_t a = __byte_perm(w, 0x64646464u, 0x5250);
    const uint32_t b = __byte_perm(w, 0x64646464u, 0x5351);
    const __half2 bias = __halves2half2(__ushort_as_half(0x6480), __ushort_as_half(0x6480));
    lo = __hsub2(*reinterpret_cast<const __half2*>(&a), bias);
    hi = __hsub2(*reinterpret_cast<const __half2*>(&b), bias);
}

template <int M>
__global__ void __launch_bounds__(kGemvThreads, (M == 1 ? 4 : 3))
mixq_gemv_w8a16_kernel(const __half* __restrict__ in, const uint8_t* __restrict__ qweight,
                       const __half* __restrict__ scales, __half* __restrict__ out, int N, int K, int groups,
                       const __half* __restrict__ bias, int act) {
    // Persistent CTAs stride over the groups of four output channels.  The weights of step s+1 (two 4 KB passes over
    // the group's two interleaved rows, possibly of the NEXT group) are in flight while step s is being multiplied.
    __shared__ float sm[2][kGemvThreads / 32][M * 4];
    const int t = threadIdx.x;
    const int r = (t >> 2) & 1;                                  // which channel of an interleaved pair this slot reads
    const int total = 2 * K;
    const int passes = (total + 4095) >> 12;
    const int steps = (passes + 1) >> 1;
    const __half2 zero = __float2half2_rn(0.0f);

    auto load_step = [&](int g, int st, uint4 (&q)[2][2]) {
        const uint8_t* qw = qweight + static_cast<size_t>(g) * 4 * K;   // (4 g / 2) rows of 2K bytes
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int lk = (st * 2 + p) * 4096 + t * 16;
            if (lk < total) {
                q[p][0] = ld_stream_v4(qw + lk);
                q[p][1] = ld_stream_v4(qw + total + lk);
            }
        }
    };

    int g = blockIdx.x;
    if (g >= groups) return;
    uint4 qc[2][2], qn[2][2];
    load_step(g, 0, qc);
    int par = 0;
    for (; g < groups; g += gridDim.x, par ^= 1) {
        const int n0 = g * 4;
        const __half2 s0 = __half2half2(scales[n0 + r]);         // channel n0 + 2*idx + r, idx = 0, 1
        const __half2 s1 = __half2half2(scales[n0 + 2 + r]);
        __half2 acc[M];                                          // (.x, .y) = (idx 0, idx 1)
#pragma unroll
        for (int m = 0; m < M; ++m) acc[m] = zero;
        for (int st = 0; st < steps; ++st) {
            {
                int ng = g, nst = st + 1;
                if (nst == steps) {
                    ng = g + gridDim.x;
                    nst = 0;
                }
                if (ng < groups) load_step(ng, nst, qn);
            }
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int lk = (st * 2 + p) * 4096 + t * 16;
                if (lk < total) {
                    const int kb = (lk >> 7) * 64 + (lk & 63);   // first k of this slot's 16 codes
                    // dequantise: pair pp = i + 2 j of the stored order is k-pair 4 i + j (undoes permute_B_rows)
                    __half2 w0[8], w1[8];
                    const uint32_t a[4] = {qc[p][0].x, qc[p][0].y, qc[p][0].z, qc[p][0].w};
                    const uint32_t b[4] = {qc[p][1].x, qc[p][1].y, qc[p][1].z, qc[p][1].w};
#pragma unroll
                    for (int wd = 0; wd < 4; ++wd) {
                        __half2 lo, hi;
                        cvt4(a[wd], lo, hi);                      // stored pairs 2 wd, 2 wd + 1
                        w0[wd] = __hfma2(lo, s0, zero);           // pp = 2 wd     -> k-pair wd
                        w0[4 + wd] = __hfma2(hi, s0, zero);       // pp = 2 wd + 1 -> k-pair 4 + wd
                        cvt4(b[wd], lo, hi);
                        w1[wd] = __hfma2(lo, s1, zero);
                        w1[4 + wd] = __hfma2(hi, s1, zero);
                    }
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const uint4* ap = reinterpret_cast<const uint4*>(in + static_cast<size_t>(m) * K + kb);
                        const uint4 x0 = __ldg(ap), x1 = __ldg(ap + 1);
                        const uint32_t xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                        for (int yy = 0; yy < 8; ++yy) {
                            const __half2 x = *reinterpret_cast<const __half2*>(&xs[yy]);
                            acc[m] = __hfma2(__lows2half2(w0[yy], w1[yy]), __low2half2(x), acc[m]);     // k = kb + 2 yy
                            acc[m] = __hfma2(__highs2half2(w0[yy], w1[yy]), __high2half2(x), acc[m]);   // k = kb + 2 yy + 1
                        }
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                qc[p][0] = qn[p][0];
                qc[p][1] = qn[p][1];
            }
        }

        // fp32 across slots: butterfly inside the warp (never crosses bit 2 = r), then the 8 warps in order
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
            for (int i = 0; i < M * 2; ++i) sm[par][warp][i * 2 + (lane >> 2)] = res[i];
        }
        __syncthreads();   // sm[par] is rewritten two groups later, after the next group's barrier
        if (t < M * 4) {
            float v = 0.0f;
#pragma unroll
            for (int j = 0; j < kGemvThreads / 32; ++j) v += sm[par][j][t];
            // optional fused epilogue, same convention as the GEMM kernels: activation in fp32 before the rounding,
            // bias added to the fp16 result
            if (act == MIXQ_ACT_SILU) v = __fdividef(v, 1.0f + __expf(-v));
            __half h = __float2half_rn(v);
            if (bias) h = __float2half_rn(__half2float(h) + __half2float(bias[n0 + (t & 3)]));
            out[static_cast<size_t>(t >> 2) * N + n0 + (t & 3)] = h;
        }
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
    const int groups = static_cast<int>(N / 4);
    const int max_ctas = device_info().num_sms * (M == 1 ? 4 : 3);   // persistent: 3-4 CTAs of 256 threads per SM
    const dim3 grid(static_cast<unsigned>(groups < max_ctas ? groups : max_ctas)), block(kGemvThreads);
    const __half* a = static_cast<const __half*>(A);
    const uint8_t* q = static_cast<const uint8_t*>(q_weight);
    const __half* s = static_cast<const __half*>(scales);
    __half* o = static_cast<__half*>(Out);
    const int n = static_cast<int>(N), k = static_cast<int>(K);
    switch (M) {
        case 1: mixq_gemv_w8a16_kernel<1><<<grid, block, 0, stream>>>(a, q, s, o, n, k, groups, static_cast<const __half*>(bias), act); break;
        case 2: mixq_gemv_w8a16_kernel<2><<<grid, block, 0, stream>>>(a, q, s, o, n, k, groups, static_cast<const __half*>(bias), act); break;
        case 3: mixq_gemv_w8a16_kernel<3><<<grid, block, 0, stream>>>(a, q, s, o, n, k, groups, static_cast<const __half*>(bias), act); break;
        default: mixq_gemv_w8a16_kernel<4><<<grid, block, 0, stream>>>(a, q, s, o, n, k, groups, static_cast<const __half*>(bias), act); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "launch gemv_w8a16");
    count_launch();
    return MIXQ_OK;
}

}  // namespace mixq
