// quant_extract.cu -- stage 1 of the W8A8O16 path: one pass over the activations that
//   * gathers the 128 outlier columns        fp_A[m, j] = A[m, ind[j]]
//   * finds the per-token scale              sa[m] = hdiv(max_k |A[m,k]|, 127)
//   * writes the INT8 codes                  A8[m,k] = int8(half2int_rn(hdiv(A[m,k], sa[m])))
//
// It replaces two reference launches:
//   ExtractOutliersAndSetToZeros / FindOutliersAndSetToZeros_kernel  (kernel/i8gemm.cu:198-244)
//   int8quant / FindRowScaleKernel<256>                               (kernel/i8gemm.cu:66-107,139-150)
// which read A three times with 2-byte accesses (the gather with one 32-byte sector per
// element).  Here each row is streamed once from HBM into shared memory with 16-byte
// cp.async (double buffered across rows), everything else works out of shared memory, and the
// INT8 codes leave in 8-byte coalesced stores: HBM traffic is the algorithmic minimum
// 2*M*K (read) + M*K + 256*M + 2*M (write).
//
// Bit-exactness with the reference kernel:
//   * the max is order-free; it is taken on the |x| bit patterns as unsigned integers, which
//     orders finite fp16 values exactly like __hmax and makes Inf/NaN the largest codes, so a
//     single reduction both yields the max and tells us whether the row is "abnormal";
//   * sa uses the real __hdiv intrinsic;
//   * device __hdiv(x, s) is fp16_rn(float(x) * rcp.approx.ftz.f32(float(s))) plus a fix-up that
//     only touches results below 2^-17 (cuda_fp16.hpp:2723-2746), all of which round to the
//     integer 0 either way.  The fast path therefore hoists the one rcp per row, multiplies in
//     fp32, rounds to fp16 (the reference's intermediate rounding), and rounds to integer with
//     the 1.5*2^23 magic-number add (round-to-nearest-even, identical to cvt.rni for |v| < 2^22).
//   * rows that are abnormal (Inf/NaN present, or sa == 0 so that x/sa is Inf/NaN) take a slow
//     path built from the very intrinsics the reference uses, so that even the saturating
//     conversions agree.
#include "mixq_internal.h"
#include "ptx.cuh"

namespace mixq {

namespace {

constexpr int kQuantThreads = 256;

// 4 halves -> 4 int8 (packed little-endian), the reference's arithmetic on the fast path:
//   q16 = fp16_rn(float(x) * rcp)           (what device __hdiv computes, cuda_fp16.hpp:2723-2746)
//   i   = rint_even(q16)                     (cvt.rni.s32.f16)
// rint is done in fp16 with the 1.5 * 2^10 magic add: for |q16| < 512 the sum lands in [1024, 2048)
// where the fp16 ulp is 1, so the add itself rounds to nearest-even and the low byte of each
// half is the two's-complement int8 code (|q16| <= ~190 on this path, see the file comment).
__device__ __forceinline__ uint32_t quant4_fast(uint32_t h01, uint32_t h23, float rcp) {
    const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&h01));
    const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&h23));
    const __half2 qa = __floats2half2_rn(__fmul_rn(fa.x, rcp), __fmul_rn(fa.y, rcp));
    const __half2 qb = __floats2half2_rn(__fmul_rn(fb.x, rcp), __fmul_rn(fb.y, rcp));
    const __half2 magic = __half2half2(__ushort_as_half(0x6600));  // 1536.0
    const __half2 ra = __hadd2(qa, magic);
    const __half2 rb = __hadd2(qb, magic);
    return __byte_perm(*reinterpret_cast<const uint32_t*>(&ra), *reinterpret_cast<const uint32_t*>(&rb), 0x6420);
}

__device__ __forceinline__ uint32_t quant4_exact(uint32_t h01, uint32_t h23, __half scale) {
    // the reference's own sequence (kernel/i8gemm.cu:103-104), saturating cvt included
    const __half2 a = *reinterpret_cast<const __half2*>(&h01);
    const __half2 b = *reinterpret_cast<const __half2*>(&h23);
    const uint32_t q0 = static_cast<uint8_t>(static_cast<int8_t>(__half2int_rn(__hdiv(__low2half(a), scale))));
    const uint32_t q1 = static_cast<uint8_t>(static_cast<int8_t>(__half2int_rn(__hdiv(__high2half(a), scale))));
    const uint32_t q2 = static_cast<uint8_t>(static_cast<int8_t>(__half2int_rn(__hdiv(__low2half(b), scale))));
    const uint32_t q3 = static_cast<uint8_t>(static_cast<int8_t>(__half2int_rn(__hdiv(__high2half(b), scale))));
    return q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
}

__device__ __forceinline__ __half2 absmax2(__half2 m, uint32_t v) {
    // NaN-propagating so that a NaN anywhere in the row is seen by the row classification below
    return __hmax2_nan(m, __habs2(*reinterpret_cast<const __half2*>(&v)));
}

// grid: persistent CTAs striding over rows; block: 256 threads; dynamic smem: 2 * K * 2 bytes.
// VPT = 16-byte vectors per thread (K <= VPT * 2048): the row lives in registers between the max
// and the quantise pass and every loop has a compile-time trip count.
// NORM = true fuses the producer of the activations into the same pass (SURVEY.md 8f "next #1", the
// reference's generalT5LayerNorm_extract_outliers, MixQ/src/kernel/mix_cuda/layernorm/layernorm.cu:121-198):
//   y[k] = fp16( clamp( (float(x[k]) * rsqrtf(sum_k x[k]^2 / K + eps)) * float(gamma[k]) ) )
// and the gather / scale / INT8 codes are taken from y (optionally also written out).
template <int VPT, bool NORM>
__global__ void __launch_bounds__(kQuantThreads)
mixq_quant_extract_kernel(const __half* __restrict__ A, int64_t M, int K, const int* __restrict__ ind, int n_ind,
                          int8_t* __restrict__ A8, __half* __restrict__ scale_a, __half* __restrict__ fp_A,
                          int mask_outliers, uint32_t* __restrict__ clear_words, int n_clear,
                          const __half* __restrict__ gamma, float eps, __half* __restrict__ y_out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ uint32_t s_warp_max[kQuantThreads / 32];
    __shared__ float s_warp_sum[kQuantThreads / 32];
    const int tid = threadIdx.x;
    const int vec_per_row = K >> 3;  // 16-byte vectors (K % 8 == 0 is checked on the host)
    uint4* const buf0 = reinterpret_cast<uint4*>(smem_raw);

    // The activations are produced by the previous kernel in the stream.
    ptx::pdl_wait_prior_grid();
    ptx::pdl_launch_dependents();   // dependents may be scheduled as our CTAs retire; they wait for this grid's completion themselves

    // kernel 2's stream-K flags live in the same workspace; clearing them here costs no extra launch
    if (blockIdx.x == 0)
        for (int i = tid; i < n_clear; i += kQuantThreads) clear_words[i] = 0u;

    auto prefetch = [&](int64_t r, int which) {
        const uint4* src = reinterpret_cast<const uint4*>(A + r * K) + tid;
        const uint32_t dst = ptx::smem_u32(buf0 + which * vec_per_row + tid);
#pragma unroll
        for (int j = 0; j < VPT; ++j)
            if (tid + j * kQuantThreads < vec_per_row) ptx::cp_async_16(dst + j * kQuantThreads * 16, src + j * kQuantThreads);
    };

    int64_t row = blockIdx.x;
    if (row < M) prefetch(row, 0);
    ptx::cp_async_commit();

    int cur = 0;
    for (; row < M; row += gridDim.x, cur ^= 1) {
        const int64_t next = row + gridDim.x;
        if (next < M) prefetch(next, cur ^ 1);
        ptx::cp_async_commit();
        ptx::cp_async_wait<1>();  // the current row has landed (the prefetch may still be in flight)
        __syncthreads();

        uint4* rowv = buf0 + cur * vec_per_row;
        __half* rowh = reinterpret_cast<__half*>(rowv);

        if constexpr (NORM) {
            // RMSNorm in place (shared memory): sum of squares in fp32, one rsqrt per token
            uint4 x[VPT];
            float ss = 0.0f;
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                x[j] = (tid + j * kQuantThreads < vec_per_row) ? rowv[tid + j * kQuantThreads] : make_uint4(0u, 0u, 0u, 0u);
                const uint32_t w[4] = {x[j].x, x[j].y, x[j].z, x[j].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
                    ss = fmaf(f.x, f.x, ss);
                    ss = fmaf(f.y, f.y, ss);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
            if ((tid & 31) == 0) s_warp_sum[tid >> 5] = ss;
            __syncthreads();
            float tot = 0.0f;
#pragma unroll
            for (int w = 0; w < kQuantThreads / 32; ++w) tot += s_warp_sum[w];
            const float rs = rsqrtf(tot / static_cast<float>(K) + eps);
            const uint4* gv = reinterpret_cast<const uint4*>(gamma);
            uint4* yv = y_out ? reinterpret_cast<uint4*>(y_out + row * K) : nullptr;
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const int idx = tid + j * kQuantThreads;
                if (idx < vec_per_row) {
                    const uint4 g4 = __ldg(gv + idx);
                    const uint32_t xw[4] = {x[j].x, x[j].y, x[j].z, x[j].w};
                    const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w};
                    uint32_t yw[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&xw[q]));
                        const float2 g = __half22float2(*reinterpret_cast<const __half2*>(&gw[q]));
                        float y0 = __fmul_rn(__fmul_rn(f.x, rs), g.x), y1 = __fmul_rn(__fmul_rn(f.y, rs), g.y);
                        // clamp_inf_for_half (reduction.cuh:111-115)
                        y0 = y0 > 0.0f ? fminf(y0, 64504.0f) : fmaxf(y0, -64504.0f);
                        y1 = y1 > 0.0f ? fminf(y1, 64504.0f) : fmaxf(y1, -64504.0f);
                        const __half2 h = __floats2half2_rn(y0, y1);
                        yw[q] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    const uint4 y4 = make_uint4(yw[0], yw[1], yw[2], yw[3]);
                    rowv[idx] = y4;
                    if (yv) yv[idx] = y4;
                }
            }
            __syncthreads();  // the normalised row is what the gather and the quantiser see
        }

        // outlier gather (and, in MixQ/src mode, zeroing: cult.cu:1588)
        if (tid < n_ind) fp_A[row * n_ind + tid] = rowh[ind[tid]];
        if (mask_outliers) {
            __syncthreads();
            if (tid < n_ind) rowh[ind[tid]] = __ushort_as_half(0);
            __syncthreads();
        }

        // row -> registers, per-token max of |x|
        uint4 v[VPT];
        __half2 m2 = __half2half2(__ushort_as_half(0));
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            v[j] = (tid + j * kQuantThreads < vec_per_row) ? rowv[tid + j * kQuantThreads] : make_uint4(0u, 0u, 0u, 0u);
            m2 = absmax2(absmax2(absmax2(absmax2(m2, v[j].x), v[j].y), v[j].z), v[j].w);
        }
        // |x| bit patterns order like unsigned integers; NaN > Inf > every finite value
        const uint32_t mb = *reinterpret_cast<const uint32_t*>(&m2);
        uint32_t m = max(mb & 0xFFFFu, mb >> 16);
        m = __reduce_max_sync(0xFFFFFFFFu, m);
        if ((tid & 31) == 0) s_warp_max[tid >> 5] = m;
        __syncthreads();  // also orders every read of this buffer before the prefetch two rows ahead
        uint32_t mx_bits = 0;
#pragma unroll
        for (int w = 0; w < kQuantThreads / 32; ++w) mx_bits = max(mx_bits, s_warp_max[w]);

        // NaN elements do not take part in the reference's __hmax chain; if the row holds any
        // Inf/NaN we recompute the max the slow, faithful way.
        __half mx = __ushort_as_half(static_cast<unsigned short>(mx_bits));
        const bool nonfinite = mx_bits >= 0x7C00u;
        if (nonfinite) {
            __half t = __ushort_as_half(0);
            for (int i = 0; i < K; ++i) t = __hmax(__habs(rowh[i]), t);  // redundant per thread; rare path
            mx = t;
        }
        const __half scale = __hdiv(mx, __float2half(127.0f));
        if (tid == 0) scale_a[row] = scale;

        uint2* dst = reinterpret_cast<uint2*>(A8 + row * K) + tid;
        if (!nonfinite && __half_as_ushort(scale) != 0) {
            float rcp;
            const float fs = __half2float(scale);
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(fs));
#pragma unroll
            for (int j = 0; j < VPT; ++j)
                if (tid + j * kQuantThreads < vec_per_row)
                    dst[j * kQuantThreads] = make_uint2(quant4_fast(v[j].x, v[j].y, rcp), quant4_fast(v[j].z, v[j].w, rcp));
        } else {
#pragma unroll
            for (int j = 0; j < VPT; ++j)
                if (tid + j * kQuantThreads < vec_per_row)
                    dst[j * kQuantThreads] = make_uint2(quant4_exact(v[j].x, v[j].y, scale), quant4_exact(v[j].z, v[j].w, scale));
            __syncthreads();  // the slow max above read the buffer after the barrier: re-order before reuse
        }
    }
    ptx::cp_async_wait<0>();
}

using QuantKernel = void (*)(const __half*, int64_t, int, const int*, int, int8_t*, __half*, __half*, int, uint32_t*, int,
                             const __half*, float, __half*);
struct QuantVariant {
    int vpt;
    QuantKernel fn, fn_norm;
};
#define MIXQ_QV(n) {n, mixq_quant_extract_kernel<n, false>, mixq_quant_extract_kernel<n, true>}
const QuantVariant kQuantVariants[] = {MIXQ_QV(1),  MIXQ_QV(2),  MIXQ_QV(3),  MIXQ_QV(4),  MIXQ_QV(6),  MIXQ_QV(8),
                                       MIXQ_QV(10), MIXQ_QV(12), MIXQ_QV(14), MIXQ_QV(16), MIXQ_QV(24), MIXQ_QV(32)};
#undef MIXQ_QV

}  // namespace

int launch_quant_extract(const void* A, int64_t M, int64_t K, const void* ind, int n_ind, void* A8, void* scale_a,
                         void* fp_A, unsigned flags, cudaStream_t stream, bool pdl, void* clear_words, int n_clear,
                         const void* gamma, float eps, void* y_out, LaunchOpts opts) {
    if (M == 0) return MIXQ_OK;
    if (K <= 0 || (K & 7) != 0) return set_error(MIXQ_ERR_BAD_ARG, "quant_extract: K must be a positive multiple of 8");
    if (n_ind < 0 || n_ind > kQuantThreads) return set_error(MIXQ_ERR_BAD_ARG, "quant_extract: n_ind must be in [0,256]");
    if (n_ind > 0 && (!ind || !fp_A)) return set_error(MIXQ_ERR_BAD_ARG, "quant_extract: ind/fp_A null");
    if (!A || !A8 || !scale_a) return set_error(MIXQ_ERR_BAD_ARG, "quant_extract: null pointer");
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(A8) & 7) ||
        (reinterpret_cast<uintptr_t>(gamma) & 15) || (reinterpret_cast<uintptr_t>(y_out) & 15))
        return set_error(MIXQ_ERR_BAD_ARG, "quant_extract: A/gamma/Y must be 16-byte and A8 8-byte aligned");

    const DeviceInfo& dev = device_info();
    if (!dev.ok) return set_error(MIXQ_ERR_CUDA, "no usable sm_100 device");
    const size_t smem = static_cast<size_t>(K) * 2 * 2;
    if (smem + 1024 > dev.max_smem_optin) return set_error(MIXQ_ERR_UNSUPPORTED, "quant_extract: K too large for shared memory staging");
    const int need_vpt = static_cast<int>((K / 8 + kQuantThreads - 1) / kQuantThreads);
    QuantKernel kern = nullptr;
    for (const QuantVariant& qv : kQuantVariants)
        if (qv.vpt >= need_vpt) {
            kern = gamma ? qv.fn_norm : qv.fn;
            break;
        }
    if (!kern) return set_error(MIXQ_ERR_UNSUPPORTED, "quant_extract: K too large (max 65536)");
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(quant_extract)");
    }
    // resident CTAs per SM: limited by threads (2048/256 = 8) and by shared memory
    int per_sm = static_cast<int>(dev.smem_per_sm / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
    int64_t grid = static_cast<int64_t>(usable_sms(opts)) * per_sm;
    if (grid > M) grid = M;

    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(kQuantThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const int mask = (flags & MIXQ_FLAG_MASK_OUTLIERS) ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<const __half*>(A), M,
                                       static_cast<int>(K), static_cast<const int*>(ind), n_ind,
                                       static_cast<int8_t*>(A8), static_cast<__half*>(scale_a),
                                       static_cast<__half*>(fp_A), mask, static_cast<uint32_t*>(clear_words), n_clear,
                                       static_cast<const __half*>(gamma), eps, static_cast<__half*>(y_out));
    if (e != cudaSuccess) return set_cuda_error(e, "launch quant_extract");
    count_launch();
    return MIXQ_OK;
}

}  // namespace mixq
