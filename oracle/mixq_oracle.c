/*
 * mixq_oracle.c -- CPU restatement of the reference's W8A8O16 hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mixq_tensorrt_llm_b200/ may link,
 * import or call this file; it is the checker used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference leg.
 *
 * Each function restates one step of MixQPlugin::enqueueImpl's M>4 branch
 * (reference TsinghuaMixQPlugin.cpp:472-532) in the order that function runs it:
 *
 *   mixq_oracle_gather        <- kernel/i8gemm.cu:198-224  (FindOutliersAndSetToZeros_kernel)
 *   mixq_oracle_outlier_gemm  <- TsinghuaMixQPlugin.cpp:122-161 (cublasGemmEx fp16, fp32 accumulate)
 *   mixq_oracle_quant         <- kernel/i8gemm.cu:66-107   (FindRowScaleKernel<256>)
 *   mixq_oracle_igemm         <- kernel/symmetric/gemm/kernel/gemm_dequant.h:224-292 (int8 x int8 -> int32)
 *   mixq_oracle_epilogue      <- kernel/symmetric/epilogue/thread/linear_combination_dequant.h:152-157
 *
 * Device arithmetic that has no host equivalent is emulated bit-for-bit:
 *   __hdiv(a,b) on the device is fp16_rn( float(a) * rcp.approx.ftz.f32(float(b)) ) with a
 *   refinement only for results in the fp16 denormal range
 *   (/usr/local/cuda/include/cuda_fp16.hpp:2723-2746).  rcp.approx is a hardware table
 *   look-up; since b is always an fp16 value there are only 65536 possible inputs, so the
 *   table is captured once on a B200 (tests/golden/make_gpu_golden.py ->
 *   tests/golden/rcp_approx_f16.bin) and passed in here.  Without the table the oracle
 *   falls back to the correctly rounded 1/b (differs from the device by <=1 ulp of the
 *   reciprocal, which changes the int8 code of ~1e-5 of the elements).
 *
 * Parity status: the reference ships no test or golden vector for this path
 * (SURVEY.md section 4).  The oracle is pinned instead against the reference's own kernels
 * compiled from /root/reference/kernel/i8gemm.cu (oracle/_ref) and run on a B200; their
 * outputs are committed under tests/golden/ (see tests/golden/README.md).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef _Float16 f16;

static inline f16 bits_to_f16(uint16_t b) {
    f16 h;
    memcpy(&h, &b, 2);
    return h;
}
static inline uint16_t f16_to_bits(f16 h) {
    uint16_t b;
    memcpy(&b, &h, 2);
    return b;
}
static inline float bits_to_f32(uint32_t b) {
    float f;
    memcpy(&f, &b, 4);
    return f;
}

/* rcp.approx.ftz.f32 of an fp16-valued input; `bits` is the fp16 bit pattern. */
static inline float rcp_approx(uint16_t bits, const uint32_t* table) {
    if (table) return bits_to_f32(table[bits]);
    return 1.0f / (float)bits_to_f16(bits);
}

/* Device __hdiv (cuda_fp16.hpp:2723-2746). */
static inline f16 dev_hdiv(f16 a, f16 b, const uint32_t* table) {
    const float fa = (float)a;
    const float fb = (float)b;
    const float rcp = rcp_approx(f16_to_bits(b), table);
    float fv = rcp * fa;
    f16 v = (f16)fv;
    const f16 den = bits_to_f16(0x008F);
    f16 av = bits_to_f16((uint16_t)(f16_to_bits(v) & 0x7FFF));
    /* __hlt is false for NaN operands */
    if (av < den && (f16)0.0f < av) {
        const float err = fmaf(-fb, fv, fa);
        fv = fmaf(rcp, err, fv);
        v = (f16)fv;
    }
    return v;
}

/* __hmax: returns the non-NaN operand when exactly one is NaN (PTX max.f16). */
static inline f16 dev_hmax(f16 a, f16 b) {
    if (a != a) return b;
    if (b != b) return a;
    return a > b ? a : b;
}

/* cvt.rni.s32.f16 (what __half2int_rn lowers to): NaN -> 0, saturating. */
static inline int32_t dev_half2int_rn(f16 h) {
    if (h != h) return 0;
    const float f = (float)h;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)nearbyintf(f); /* default rounding mode = nearest-even */
}

/* kernel/i8gemm.cu:198-224: fp_A[i, j] = A[i, ind[j]]; the zeroing store is commented
 * out in the plugin's copy (line 218) and live in MixQ/src (cult.cu:1588) -- see `mask`
 * in mixq_oracle_quant. */
void mixq_oracle_gather(const uint16_t* A, int64_t M, int64_t K, const int32_t* ind, int n_ind,
                        uint16_t* fp_A) {
    for (int64_t m = 0; m < M; ++m)
        for (int j = 0; j < n_ind; ++j) fp_A[m * n_ind + j] = A[m * K + ind[j]];
}

/* kernel/i8gemm.cu:66-107 FindRowScaleKernel<256>:
 *   max   = reduce(__hmax, __habs(row))                       (:76-88, order-free)
 *   scale = __hdiv(max, 127.0)                                (:96)
 *   q[k]  = (int8_t)__half2int_rn(__hdiv(row[k], scale))      (:103-104)
 * mask != 0 restates MixQ/src (cult.cu:1588): the n_ind outlier columns are treated as
 * zero for both the max and the quantised value. */
void mixq_oracle_quant(const uint16_t* A, int64_t M, int64_t K, const uint32_t* rcp_table,
                       const int32_t* ind, int n_ind, int mask, int8_t* q, uint16_t* scale_a) {
    uint8_t* is_out = NULL;
    if (mask && n_ind > 0) {
        is_out = (uint8_t*)calloc((size_t)K, 1);
        for (int j = 0; j < n_ind; ++j) is_out[ind[j]] = 1;
    }
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < M; ++m) {
        const uint16_t* row = A + m * K;
        f16 mx = (f16)0.0f;
        for (int64_t k = 0; k < K; ++k) {
            uint16_t b = (is_out && is_out[k]) ? 0 : (uint16_t)(row[k] & 0x7FFF); /* __habs */
            mx = dev_hmax(bits_to_f16(b), mx);
        }
        const f16 scale = dev_hdiv(mx, (f16)127.0f, rcp_table);
        scale_a[m] = f16_to_bits(scale);
        for (int64_t k = 0; k < K; ++k) {
            const f16 x = (is_out && is_out[k]) ? (f16)0.0f : bits_to_f16(row[k]);
            q[m * K + k] = (int8_t)dev_half2int_rn(dev_hdiv(x, scale, rcp_table));
        }
    }
    free(is_out);
}

/* TsinghuaMixQPlugin.cpp:122-161: Out0 = fp_A [M,kf] x fp_weight[N,kf]^T, fp16 in/out,
 * CUBLAS_COMPUTE_32F.  cuBLAS' accumulation order is unspecified; the products are exact
 * in fp32, so the oracle accumulates in double and rounds once to fp32 and once to fp16
 * (the reference rounds the fp32 accumulator to fp16 on store). */
void mixq_oracle_outlier_gemm(const uint16_t* fp_A, const uint16_t* fp_W, int64_t M, int64_t N,
                              int kf, uint16_t* out0) {
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < M; ++m) {
        float a[1024];
        for (int j = 0; j < kf; ++j) a[j] = (float)bits_to_f16(fp_A[m * kf + j]);
        for (int64_t n = 0; n < N; ++n) {
            const uint16_t* w = fp_W + n * kf;
            double acc = 0.0;
            for (int j = 0; j < kf; ++j) acc += (double)a[j] * (double)(float)bits_to_f16(w[j]);
            out0[m * N + n] = f16_to_bits((f16)(float)acc);
        }
    }
}

/* The O(MNK) term: acc[m,n] = sum_k q[m,k] * w[n,k] in exact int32
 * (kernel/symmetric/gemm/kernel/gemm_dequant.h:224-292; OpMultiplyAddSaturate cannot
 * saturate for K*127*128 < 2^31). */
__attribute__((target_clones("arch=skylake-avx512", "avx2", "default"))) static int32_t dot_i8(
    const int8_t* __restrict a, const int8_t* __restrict b, int64_t K) {
    int32_t s = 0;
    for (int64_t k = 0; k < K; ++k) s += (int32_t)a[k] * (int32_t)b[k];
    return s;
}

void mixq_oracle_igemm(const int8_t* q, const int8_t* w, int64_t M, int64_t N, int64_t K,
                       int32_t* acc) {
#pragma omp parallel for schedule(static) collapse(2)
    for (int64_t mb = 0; mb < M; mb += 8)
        for (int64_t n = 0; n < N; ++n) {
            const int64_t me = mb + 8 < M ? mb + 8 : M;
            for (int64_t m = mb; m < me; ++m) acc[m * N + n] = dot_i8(q + m * K, w + n * K, K);
        }
}

/* linear_combination_dequant.h:152-157 with C = Out0:
 *   D = __float2half( float(acc) * (float(sb[n]) * float(sa[m])) + float(C) )
 * compiled by nvcc to I2FP.F32.S32, FMUL (scale product), FFMA, F2FP.F16.F32,
 * i.e. a single-rounded fused multiply-add followed by one RN-even fp16 rounding. */
void mixq_oracle_epilogue(const int32_t* acc, const uint16_t* scale_a, const uint16_t* scale_b,
                          const uint16_t* out0, int64_t M, int64_t N, uint16_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < M; ++m) {
        const float sa = (float)bits_to_f16(scale_a[m]);
        for (int64_t n = 0; n < N; ++n) {
            const float p = (float)bits_to_f16(scale_b[n]) * sa;
            const float c = out0 ? (float)bits_to_f16(out0[m * N + n]) : 0.0f;
            out[m * N + n] = f16_to_bits((f16)fmaf((float)acc[m * N + n], p, c));
        }
    }
}

/* "next #1" producer fusion -- MixQ/src/kernel/mix_cuda/layernorm/layernorm.cu:121-157
 * (generalT5LayerNorm_extract_outliers, RMSNorm part):
 *   s = rsqrtf(sum_k x^2 / n + eps);  y = clamp_inf_for_half((float(x) * s) * float(gamma))
 * The device sums in fp32 in tree order and uses the approximate rsqrt; the oracle sums in double and
 * uses 1/sqrtf, so y may differ from the device in the last fp16 bit on a small fraction of elements. */
void mixq_oracle_rmsnorm(const uint16_t* X, const uint16_t* gamma, float eps, int64_t M, int64_t K, uint16_t* Y) {
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < M; ++m) {
        double ss = 0.0;
        for (int64_t k = 0; k < K; ++k) {
            const double x = (double)(float)bits_to_f16(X[m * K + k]);
            ss += x * x;
        }
        const float s = 1.0f / sqrtf((float)ss / (float)K + eps);
        for (int64_t k = 0; k < K; ++k) {
            float y = ((float)bits_to_f16(X[m * K + k]) * s) * (float)bits_to_f16(gamma[k]);
            y = y > 0.0f ? fminf(y, 65504.0f - 1000.0f) : fmaxf(y, -65504.0f + 1000.0f); /* reduction.cuh:111-115 */
            Y[m * K + k] = f16_to_bits((f16)y);
        }
    }
}

/* MixQPlugin::enqueueImpl M>4 branch, TsinghuaMixQPlugin.cpp:518-532, in call order.
 * scratch buffers are caller-provided so the timing legs do not measure malloc:
 *   fp_A [M,128] u16 | out0 [M,N] u16 | q [M,K] i8 | sa [M] u16 | acc [M,N] i32 */
void mixq_oracle_forward(const uint16_t* A, const int8_t* W8, const uint16_t* scale_b,
                         const uint16_t* fp_weight, const int32_t* ind, int n_ind, int64_t M,
                         int64_t N, int64_t K, const uint32_t* rcp_table, int mask,
                         uint16_t* fp_A, uint16_t* out0, int8_t* q, uint16_t* sa, int32_t* acc,
                         uint16_t* out) {
    mixq_oracle_gather(A, M, K, ind, n_ind, fp_A);                       /* :519 */
    mixq_oracle_outlier_gemm(fp_A, fp_weight, M, N, n_ind, out0);        /* :521 */
    mixq_oracle_quant(A, M, K, rcp_table, ind, n_ind, mask, q, sa);      /* :522 */
    mixq_oracle_igemm(q, W8, M, N, K, acc);                              /* :529 mainloop */
    mixq_oracle_epilogue(acc, sa, scale_b, out0, M, N, out);             /* :529 epilogue */
}

/* torchrun exports OMP_NUM_THREADS=1; the timing legs ask for all host cores explicitly. */
/* ------------------------------------------------------------------------------------------------
 * M <= 4 branch: W8A16 weight-only GEMV (reference TsinghuaMixQPlugin.cpp:472,641-647 ->
 * weightonlykernel/fpA_intB_gemm_wrapper.cu:29-57 -> weightOnlyBatchedGemv/kernelLauncher.cu:179-206
 * (Int8b, PerChannel: NPerBlock = 2, Batch = M, BlockSize = 256) -> weightOnlyBatchedGemv/kernel.h:285-438).
 *
 * The kernel's arithmetic is restated slot by slot, because its rounding sequence IS the specification:
 *   - a block of 256 thread slots owns 4 consecutive output channels n0..n0+3 (two interleaved row pairs);
 *   - slot t reads 16 codes of channel n0 + 2*idx + r, r = (t / 4) % 2, idx in {0, 1}, at
 *       k = (t / 8 + 32 * it) * 64 + (t % 4) * 16 + y,  y = 0..15,   while t * 16 + it * 4096 < 2 * K   (kernel.h:322-323);
 *   - w = fp16( (code - 128) * scale[n] )   (__hfma2(v, scale, 0), kernel.h:355-356);
 *   - per slot and token, an fp16 chain  acc = fp16_fma(w, in[m, k], acc)  over y ascending, over `it` ascending
 *     (__hfma2 with single rounding, kernel.h:401-411);
 *   - the slot sums are converted to fp32 and combined with the butterfly xor 16, 8, 2, 1 inside each warp
 *     (kernel.h:155-161), the 8 warp results are added in warp order starting from 0.f (kernel.h:420-425),
 *     and the fp32 sum is rounded to fp16 (kernel.h:432).
 * `qweight` is the processed (permuted / transposed / 2-column interleaved / +128 biased / byte-swizzled) tensor
 * EETQ.quant_weights produces (cutlass_preprocessors.cc:497-533); the byte fetched for (channel, k) follows from
 * undoing those steps exactly as the kernel does (kernel.h:336-368):
 *   byte address = (n / 2) * 2K + (k / 64) * 128 + (n % 2) * 64 + (k % 64) / 16 * 16 + pos(k % 16)
 *   pos(y): p = (y / 8) * 2 + ((y % 8) / 2) * 4 + y % 2  (position after permute_B_rows),  then bytes 1 and 2 of each
 *           aligned group of four are swapped (add_bias_and_interleave_int8s_inplace).
 */
static inline f16 dev_hfma(f16 a, f16 b, f16 c) {
    /* exact product (22 significant bits), one rounding to fp16 of product + c: the double sum is made
       round-to-odd when it is inexact, so the final rounding sees the sticky information */
    const double p = (double)a * (double)b;
    const double cc = (double)c;
    double s = p + cc;
    if (isfinite(s)) {
        const double bb = s - p;
        const double err = (p - (s - bb)) + (cc - bb);
        if (err != 0.0) {
            uint64_t u;
            memcpy(&u, &s, 8);
            if ((u & 1) == 0) {
                if ((err > 0) == (s > 0)) u += 1; else u -= 1;
                memcpy(&s, &u, 8);
            }
        }
    }
    return (f16)s;
}

/* test hook: the emulated device __hfma on raw fp16 bit patterns (tests/test_oracle.py checks it against exact rational
   arithmetic) */
uint16_t mixq_oracle_hfma(uint16_t a, uint16_t b, uint16_t c) {
    return f16_to_bits(dev_hfma(bits_to_f16(a), bits_to_f16(b), bits_to_f16(c)));
}

static inline int gemv_byte_pos(int y) {
    const int p = (y / 8) * 2 + ((y % 8) / 2) * 4 + (y % 2);
    const int q = p & 3;
    return (p & ~3) | (q == 1 ? 2 : q == 2 ? 1 : q);
}

void mixq_oracle_gemv_w8a16(const uint16_t* in, const uint8_t* qweight, const uint16_t* scales, int64_t M,
                            int64_t N, int64_t K, uint16_t* out) {
    int pos[16];
    for (int y = 0; y < 16; ++y) pos[y] = gemv_byte_pos(y);
#pragma omp parallel for schedule(static)
    for (int64_t blk = 0; blk < N / 4; ++blk) {
        const int64_t n0 = blk * 4;
        float slot[256][8];                       /* [slot][m * 2 + idx] as fp32 */
        for (int t = 0; t < 256; ++t) {
            const int r = (t / 4) % 2;
            f16 acc[8];
            for (int i = 0; i < 8; ++i) acc[i] = (f16)0.0f;
            for (int64_t it = 0; (int64_t)t * 16 + it * 4096 < 2 * K; ++it) {
                const int64_t kb = ((int64_t)(t / 8) + 32 * it) * 64 + (t % 4) * 16;
                f16 w[2][16];
                for (int idx = 0; idx < 2; ++idx) {
                    const int64_t n = n0 + 2 * idx + r;
                    const f16 sc = bits_to_f16(scales[n]);
                    const uint8_t* src = qweight + (n / 2) * 2 * K + (kb / 64) * 128 + (n % 2) * 64 + (kb % 64);
                    for (int y = 0; y < 16; ++y) {
                        const f16 code = (f16)(float)((int)src[pos[y]] - 128);
                        w[idx][y] = (f16)((float)code * (float)sc + 0.0f);   /* __hfma2(v, scale, 0): product exact in fp32 */
                    }
                }
                for (int64_t m = 0; m < M; ++m)
                    for (int y = 0; y < 16; ++y) {
                        const f16 x = bits_to_f16(in[m * K + kb + y]);
                        acc[m * 2 + 0] = dev_hfma(w[0][y], x, acc[m * 2 + 0]);
                        acc[m * 2 + 1] = dev_hfma(w[1][y], x, acc[m * 2 + 1]);
                    }
            }
            for (int i = 0; i < 8; ++i) slot[t][i] = (float)acc[i];
        }
        for (int64_t m = 0; m < M; ++m)
            for (int idx = 0; idx < 2; ++idx)
                for (int r = 0; r < 2; ++r) {
                    float v = 0.0f;
                    for (int wp = 0; wp < 8; ++wp) {
                        float lane[32];
                        for (int l = 0; l < 32; ++l) lane[l] = slot[wp * 32 + l][m * 2 + idx];
                        static const int steps[4] = {16, 8, 2, 1};
                        for (int st = 0; st < 4; ++st) {
                            float nx[32];
                            for (int l = 0; l < 32; ++l) nx[l] = lane[l] + lane[l ^ steps[st]];
                            memcpy(lane, nx, sizeof(lane));
                        }
                        v += lane[r * 4];              /* lanes 0 and 4 publish rows r = 0 and 1 */
                    }
                    out[m * N + n0 + 2 * idx + r] = f16_to_bits((f16)v);
                }
    }
}

void mixq_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int mixq_oracle_num_threads(void) {
    int n = 1;
#ifdef _OPENMP
#pragma omp parallel
    {
#pragma omp master
        n = omp_get_num_threads();
    }
#endif
    return n;
}
