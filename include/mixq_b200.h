/*
 * mixq_b200.h -- C ABI of the B200-native W8A8O16 mixed-precision linear.
 *
 * This is the drop-in boundary for the hot path behind
 * TsinghuaMixQPlugin::enqueue (reference TsinghuaMixQPlugin.cpp:384-765).
 * Everything above it (the IPluginV2DynamicExt adapter in
 * mixq_tensorrt_llm_b200/csrc/mixq_plugin.{h,cpp}, the Python MixQLinear mirror,
 * the tests and bench.py) calls ONLY these entry points.  No torch / TensorRT
 * types appear in any signature: plain pointers, sizes and a cudaStream_t
 * passed as void*.
 *
 * Conventions
 *   - every function returns 0 on success and a negative mixq_status on error,
 *     never throws, never calls exit() (the reference ignores cuBLAS/CUTLASS
 *     status and always returns 0: TsinghuaMixQPlugin.cpp:402,752;
 *     kernel/i8gemm.cu:190-192 -- we report instead);
 *   - device functions are asynchronous on `stream`, perform no allocation and
 *     no host synchronisation, and are CUDA-graph-capture safe;
 *   - there is NO CPU fallback: without a CUDA device these calls fail with
 *     MIXQ_ERR_CUDA.
 */
#ifndef MIXQ_B200_H_
#define MIXQ_B200_H_

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIXQ_NUM_OUTLIERS 128 /* num_ind: TsinghuaMixQPlugin.cpp:518; fp_features: model_config_utils.py:446 */

typedef enum mixq_status {
    MIXQ_OK = 0,
    MIXQ_ERR_BAD_ARG = -1,    /* null pointer, non-positive dim, misaligned pointer/stride */
    MIXQ_ERR_WORKSPACE = -2,  /* workspace too small */
    MIXQ_ERR_CUDA = -3,       /* a CUDA runtime/driver call or launch failed (see mixq_last_error) */
    MIXQ_ERR_UNSUPPORTED = -4 /* shape outside what the kernels support (see mixq_enqueue) */
} mixq_status;

/* Flags for mixq_enqueue / mixq_quant_extract. Default (0) is the TensorRT plugin's behaviour. */
enum {
    /* Zero A[:, ind] before the per-token amax/quantize, as MixQ/src does
     * (MixQ/src/kernel/mix_cuda/cult.cu:1588); the plugin copy has that store
     * commented out (kernel/i8gemm.cu:218). */
    MIXQ_FLAG_MASK_OUTLIERS = 1u << 0,
    /* Skip the M<=4 weight-only branch (TsinghuaMixQPlugin.cpp:472,641-647) and
     * run the mixed W8A8O16 path for every M (also what happens when q_weight is NULL). */
    MIXQ_FLAG_FORCE_MIXED = 1u << 1,
    /* Host-buffer calls only (mixq_linear_host, mixq_linears_host, mixq_gated_host): enqueue the upload, the kernels
     * and the download and return without waiting.  A_host must stay untouched and Out_host is not valid until
     * mixq_host_drain returns.  See mixq_host_drain. */
    MIXQ_FLAG_HOST_ASYNC = 1u << 8
};

/* The seven plugin inputs + one output, in the order of plugin.py:142-150 /
 * TsinghuaMixQPlugin.cpp:425-445. All are device pointers. The TensorRT side
 * types every tensor as kHALF; the true element types are given here. */
typedef struct mixq_tensors {
    const void* A;               /* [M, K]  fp16 activations (input 0)                           */
    const void* W8;              /* [N, K]  int8 weights, outlier columns zeroed (input 1)       */
    const void* scale_b;         /* [N]     fp16 per-output-channel weight scale (input 2)       */
    const void* fp_weight;       /* [N,128] fp16 outlier weight columns (input 3)                */
    const void* ind;             /* [128]   int32 outlier column indices (input 4)               */
    const void* q_weight;        /* [K, N]  int8 weight-only layout, M<=4 branch only (input 5)  */
    const void* scaling_factors; /* [N]     fp16 scales for q_weight (input 6)                   */
    void* Out;                   /* [M, N]  fp16 output                                          */
} mixq_tensors;

/* Library identity. */
const char* mixq_version(void);
/* Human-readable text for the last failing call on this thread ("" if none). */
const char* mixq_last_error(void);
/* 1 when a CUDA device of compute capability 10.x is usable, else 0. */
int mixq_device_ok(void);

/* Bytes of device workspace mixq_enqueue needs for (M, N, K). Layout (each
 * block 128-byte aligned, as nextWorkspacePtr does, TsinghuaMixQPlugin.cpp:206-215):
 *   A8 int8 [M,K] | scale_a fp16 [M] | fp_A fp16 [M,128]
 * i.e. M*K + 258*M bytes plus alignment; non-decreasing in M, so the size for a profile's maximum covers every
 * smaller batch.  Replaces the reference's max(M*K + 2M + 2*K*N, 16*M*N) (TsinghuaMixQPlugin.cpp:342-346).
 * The opt-in split-K tile configurations (ids 8, 11, 12: never picked automatically) need
 * mixq_decode_workspace_size(M, N) more bytes behind it: mixq_workspace_size_opt() includes them for such a
 * mixq_options; mixq_enqueue_opt uses that scratch when the workspace it is given is large enough. */
size_t mixq_workspace_size(int64_t M, int64_t N, int64_t K);

/* The hot path: replaces MixQPlugin::enqueueImpl's M>4 branch
 * (TsinghuaMixQPlugin.cpp:472-532: ExtractOutliersAndSetToZeros -> cublasGemmEx
 * fp16 -> int8quant -> int8FusedDequantizeCUDA) with two launches:
 *   1. mixq_quant_extract   (per-token INT8 quantize + outlier gather)
 *   2. mixq_gemm_dequant    (tcgen05 INT8 GEMM + fused FP16 outlier GEMM + dequant)
 *
 *   Out[m,n] = fp16( float(sum_k A8[m,k]*W8[n,k]) * (float(sb[n])*float(sa[m]))
 *                    + float(fp16(sum_j fp_A[m,j]*fp_weight[n,j])) )
 *
 * Requirements: K % 16 == 0, N % 8 == 0, all pointers 16-byte aligned, `ind`
 * values in [0, K). M may be any positive value (M == 0 is a no-op).
 * For M <= 4 the reference switches to a weight-only GEMV over q_weight
 * (TsinghuaMixQPlugin.cpp:472,641-647).  So does this call when t->q_weight is given
 * (then only A, q_weight, scaling_factors and Out are read, N % 4 == 0 and K % 64 == 0
 * are required and the workspace is not used): see mixq_gemv_w8a16.  With q_weight ==
 * NULL or MIXQ_FLAG_FORCE_MIXED the mixed path runs for every M.  */
int mixq_enqueue(const mixq_tensors* t, int64_t M, int64_t N, int64_t K, void* workspace,
                 size_t workspace_bytes, unsigned flags, void* stream);

/* The same call with a fused epilogue:
 *   y   = fp16( act( fma(float(acc), sb*sa, float(out0)) ) )     act in fp32 BEFORE the rounding, as the reference's
 *                                                                 LinearCombinationDequantSilu (linear_combination_dequant.h:167-272,
 *                                                                 silu(x) = x / (1 + expf(-x)), fast-math build)
 *   Out = bias ? fp16( float(y) + float(bias[n]) ) : y           the bias add the reference does after the plugin
 *                                                                 (plugin.py:158-160; MixQ/src/mixquant/modules/linear.py:368-369)
 * epi == NULL is mixq_enqueue.  Also applied on the M <= 4 weight-only branch. */
enum { MIXQ_ACT_NONE = 0, MIXQ_ACT_SILU = 1 };
typedef struct mixq_epilogue {
    const void* bias; /* fp16 [N] device pointer, or NULL */
    int activation;   /* MIXQ_ACT_*                        */
} mixq_epilogue;
int mixq_enqueue_ex(const mixq_tensors* t, int64_t M, int64_t N, int64_t K, void* workspace,
                    size_t workspace_bytes, const mixq_epilogue* epi, unsigned flags, void* stream);
int mixq_gemm_dequant_ex(const void* A8, const void* W8, const void* scale_a, const void* scale_b,
                         const void* fp_A, const void* fp_weight, void* Out, int64_t M, int64_t N,
                         int64_t K, const mixq_epilogue* epi, void* stream);

/* ---- gated MLP input half in one call -------------------------------------------------------
 * The gate and up projections of one MLP consume the same activations; the reference's fused MLP quantises them once
 * (MixGemmCache), runs the gate projection with SiLU fused into the dequant epilogue and multiplies the two fp16 results
 * (MixQ/src/mixquant/modules/fused/mlp.py:57-70; modules/linear.py:288-373; epilogue
 * kernel/symmetric/epilogue/thread/linear_combination_dequant.h:167-272):
 *   g   = fp16( silu( fma(float(acc_gate), sb_gate*sa, float(out0_gate)) ) )
 *   u   = fp16(       fma(float(acc_up),   sb_up*sa,   float(out0_up))    )
 *   Out = fp16( float(g) * float(u) )                                       [M, N], written to gate->Out
 * `gate` and `up` are the two linears' plugin tensor tables: same A, same `ind` (both see the same activation
 * statistics; gate->ind is the one read), each its own W8 / scale_b / fp_weight of N output channels; up->Out and the
 * q_weight fields are ignored (the mixed path serves every M).  For M <= 1024 this is ONE quantise launch and ONE GEMM
 * launch whose CTA pairs hold a gate tile and the matching up tile side by side in TMEM; the two [M, N] intermediates
 * never reach memory.  Larger M runs the two GEMMs over the shared quantised A plus one elementwise pass (same roundings)
 * and needs M*N*2 more bytes of workspace than mixq_workspace_size: mixq_gated_workspace_size includes them. */
size_t mixq_gated_workspace_size(int64_t M, int64_t N, int64_t K);
struct mixq_options;
int mixq_enqueue_gated(const mixq_tensors* gate, const mixq_tensors* up, int64_t M, int64_t N, int64_t K,
                       void* workspace, size_t workspace_bytes, const struct mixq_options* opt, unsigned flags,
                       void* stream);
/* Stage 2 of it alone (A already quantised); scratch: M*N*2 bytes, only read for M > 1024 (may be NULL otherwise). */
int mixq_gemm_dequant_gated(const void* A8, const void* scale_a, const void* fp_A, const void* W8_gate,
                            const void* scale_b_gate, const void* fp_weight_gate, const void* W8_up,
                            const void* scale_b_up, const void* fp_weight_up, void* Out, int64_t M, int64_t N,
                            int64_t K, const struct mixq_options* opt, void* scratch, size_t scratch_bytes,
                            void* stream);

/* Stage 1 alone. Replaces int8quant (kernel/i8gemm.cu:66-107,139-150) and
 * ExtractOutliersAndSetToZeros (kernel/i8gemm.cu:198-244) in one pass over A.
 *   sa[m]   = hdiv(max_k |A[m,k]|, 127)
 *   A8[m,k] = int8(half2int_rn(hdiv(A[m,k], sa[m])))
 *   fp_A[m,j] = A[m, ind[j]]   (j < n_ind; fp_A row stride = n_ind)
 * fp_A/ind may be NULL (n_ind = 0) to quantize only. */
int mixq_quant_extract(const void* A, int64_t M, int64_t K, const void* ind, int n_ind, void* A8,
                       void* scale_a, void* fp_A, unsigned flags, void* stream);

/* Stage 1 fused with its producer (SURVEY.md 8f "next #1"): RMSNorm -> outlier extract -> per-token INT8
 * quantise in one pass over the hidden state X [M,K].  Restates generalT5LayerNorm_extract_outliers
 * (MixQ/src/kernel/mix_cuda/layernorm/layernorm.cu:121-198):
 *   y[m,k] = fp16( clamp( (float(X[m,k]) * rsqrtf(sum_k X[m,k]^2 / K + eps)) * float(gamma[k]) ) )
 * then exactly mixq_quant_extract on y.  The reference kernel always zeroes the outlier columns of y
 * (pass MIXQ_FLAG_MASK_OUTLIERS for that); Y (fp16 [M,K], may be NULL) receives y before that zeroing. */
int mixq_rmsnorm_quant_extract(const void* X, const void* gamma, float eps, int64_t M, int64_t K,
                               const void* ind, int n_ind, void* A8, void* scale_a, void* fp_A, void* Y,
                               unsigned flags, void* stream);

/* Stage 2 alone. Replaces cublasGemmEx fp16 (TsinghuaMixQPlugin.cpp:122-161) +
 * int8FusedDequantizeCUDA (kernel/i8gemm.cu:151-194) in one kernel.
 * fp_A/fp_weight may both be NULL: then the addend is 0 (plain W8A8 dequant GEMM). */
int mixq_gemm_dequant(const void* A8, const void* W8, const void* scale_a, const void* scale_b,
                      const void* fp_A, const void* fp_weight, void* Out, int64_t M, int64_t N,
                      int64_t K, void* stream);

/* The M <= 4 branch alone.  Replaces w8_a16_gemm_forward_cuda -> weight_only_batched_gemv
 * (weightonlykernel/fpA_intB_gemm_wrapper.cu:29-57; weightOnlyBatchedGemv/kernel.h:285-438):
 *   Out[m,n] = sum_k A[m,k] * fp16((code[k,n] - 128) * scales[n]),  1 <= M <= 4,
 * fp16 accumulation per thread and fp32 across threads in the reference kernel's order (bit-identical to it).
 * q_weight is the processed int8 tensor EETQ.quant_weights returns for W^T [K,N]
 * (cutlass_preprocessors.cc:497-533, Sm80 layout); scales is its fp16 [N] companion. */
int mixq_gemv_w8a16(const void* A, const void* q_weight, const void* scales, void* Out, int64_t M,
                    int64_t N, int64_t K, void* stream);

/* Stage 2 with scratch for the stream-K schedule the library prefers for decode-sized M (tiles that do
 * not fill the GPU evenly are cut along K; int32 partial sums meet in `workspace`).  Results are
 * bit-identical to mixq_gemm_dequant.  workspace_bytes >= mixq_gemm_workspace_size(), 16-byte aligned. */
size_t mixq_gemm_workspace_size(void);
int mixq_gemm_dequant_ws(const void* A8, const void* W8, const void* scale_a, const void* scale_b,
                         const void* fp_A, const void* fp_weight, void* Out, int64_t M, int64_t N,
                         int64_t K, void* workspace, size_t workspace_bytes, void* stream);

/* ---- row-parallel linear with the all-reduce fused into the GEMM kernel ----
 * Replaces "plugin, then allreduce(x, tp_group)" (reference plugin.py:152-156) for the row-parallel shards of a
 * tensor-parallel layer by ONE kernel per rank: every rank's epilogue pushes its fp16 partial tiles straight into
 * the staging area of the rank that owns the tile (peer memory over NVLink), owners sum the `world` partials in
 * fp32 in rank order (deterministic; one rounding to fp16) and write the result into every rank's Out.  When the
 * call's kernel completes on a rank, that rank's Out holds the reduced [M,N] result.
 *
 * All ranks must issue the same sequence of fused calls with the same (M, N), and on each rank the calls that share
 * a peer group must be ordered (one stream, or event-ordered streams): two of them in flight at once on one rank
 * would share the staging area and the counters.  The calls are CUDA-graph capturable (no NCCL, no host state:
 * the launch epoch lives in the counter block).  The pointers are addresses valid in
 * THIS process for every rank's buffers (CUDA IPC / cuMem fabric handles / torch symmetric memory: whatever the host
 * uses to map peer memory); index `rank` is the local buffer.  `counters` must be zeroed once after allocation and is
 * re-armed by the kernel itself.  Buffers are M/N dependent only through the two size functions.
 *
 * Decode-sized results take a two-launch variant of the same exchange while the bytes a rank must receive,
 * (world-1) * M*N*2, stay within 4 MB (8 MB at 2 ranks): the GEMM (whatever tile configuration `auto` picks) writes the
 * partial into staging[rank] and a small kernel lets every rank pull all partials with peer loads and reduce them in rank
 * order (csrc/allreduce_pull.cu): the same arithmetic, bit-identical results, one NVLink hop instead of two.
 * mixq_options.gemm_config = 9 keeps the one-kernel path; MIXQ_PULL_MAX_INGRESS_MB (environment) moves the limit. */
#define MIXQ_MAX_RANKS 8
typedef struct mixq_peer_group {
    int world, rank;
    void* out[MIXQ_MAX_RANKS];      /* every rank's Out [M,N] fp16                                        */
    void* staging[MIXQ_MAX_RANKS];  /* every rank's staging area, >= mixq_allreduce_staging_size() bytes  */
    void* counters[MIXQ_MAX_RANKS]; /* every rank's counter block, >= mixq_allreduce_counter_size() bytes */
    size_t staging_bytes, counter_bytes;
    void* out_multicast;            /* optional: NVSwitch multicast mapping of the Out buffers (a store to it lands in every
                                       rank's Out): the result is then broadcast with one multimem store per row instead of
                                       one TMA store per rank -- (world-1)x fewer bytes leave the GPU in that phase; the
                                       arithmetic, and so the result, is unchanged.  NULL = per-rank stores.               */
} mixq_peer_group;
size_t mixq_allreduce_staging_size(int64_t M, int64_t N, int world);
size_t mixq_allreduce_counter_size(int64_t M, int64_t N, int world);
/* Peer waits inside the fused kernel poll with back-off and never trap: ranks may be skewed by seconds (first call after a
 * per-rank checkpoint load, a debugger, a long kernel ahead of the call on one rank).  Callers should still barrier once
 * before the FIRST fused call (every rank's counter block must have been zeroed).  After MIXQ_AR_TIMEOUT_MS (environment,
 * default 60000, 0 = wait for ever) without progress a rank gives up, finishes the launch with an invalid result and raises
 * a sticky error word in its counter block; mixq_allreduce_check copies that word back (synchronising `stream`) and
 * returns MIXQ_ERR_CUDA if it is set, 0 otherwise.  `clear` != 0 resets it. */
int mixq_allreduce_check(void* counters_local, int clear, void* stream);
/* mixq_enqueue with the reduction fused in: t->Out is ignored, the result lands in g->out[i] on every rank i. */
int mixq_enqueue_allreduce(const mixq_tensors* t, int64_t M, int64_t N, int64_t K, void* workspace,
                           size_t workspace_bytes, const mixq_peer_group* g, unsigned flags, void* stream);
/* Stage 2 alone with the reduction fused in (tests, callers that share one quantised A). */
int mixq_gemm_dequant_allreduce(const void* A8, const void* W8, const void* scale_a, const void* scale_b,
                                const void* fp_A, const void* fp_weight, int64_t M, int64_t N, int64_t K,
                                const mixq_peer_group* g, void* stream);

/* End-to-end call with HOST buffers for the per-call tensors: copies A (host,
 * fp16 [M,K]) to the device, runs mixq_enqueue with the device-resident weights
 * in `t` (t->A and t->Out are ignored), copies Out back to `Out_host`
 * (fp16 [M,N]) and synchronises `stream`.  `dev_scratch` must hold
 * mixq_host_scratch_size(M,N,K) bytes of device memory.  This is the call
 * bench.py times for its `e2e` figure. */
size_t mixq_host_scratch_size(int64_t M, int64_t N, int64_t K);
int mixq_linear_host(const mixq_tensors* t, const void* A_host, void* Out_host, int64_t M,
                     int64_t N, int64_t K, void* dev_scratch, size_t dev_scratch_bytes,
                     unsigned flags, void* stream);

/* The same for several linears that consume the SAME host activations (the fused q/k/v projection, or the gate and up
 * projections of one MLP: MixQ/src/mixquant/modules/fused/mlp.py:57-70 shares one quantised input between them): A
 * crosses PCIe once, every linear i (tensors t[i], N[i] output channels, 1 <= count <= 8) writes its own Out_host[i]
 * [M, N[i]].  The three-stage pipeline (H2D of row slab c+1 | kernels of slab c | D2H of slab c-1) is shared. */
size_t mixq_linears_host_scratch_size(int64_t M, const int64_t* N, int count, int64_t K);
int mixq_linears_host(const mixq_tensors* const* t, int count, const void* A_host, void* const* Out_host,
                      int64_t M, const int64_t* N, int64_t K, void* dev_scratch, size_t dev_scratch_bytes,
                      unsigned flags, void* stream);

/* mixq_enqueue_gated with HOST buffers: A crosses PCIe once, one [M, N] result comes back (half the bytes of the two
 * projections' outputs). */
size_t mixq_gated_host_scratch_size(int64_t M, int64_t N, int64_t K);
int mixq_gated_host(const mixq_tensors* gate, const mixq_tensors* up, const void* A_host, void* Out_host,
                    int64_t M, int64_t N, int64_t K, void* dev_scratch, size_t dev_scratch_bytes,
                    unsigned flags, void* stream);

/* Pipelined use of the host-buffer calls.  A decode step is a chain of small linears (4 MB in, 4-12 MB out each): run one
 * at a time, PCIe carries the upload, then idles during the kernels, then carries the download.  With
 * MIXQ_FLAG_HOST_ASYNC the calls of one thread queue up behind one another instead: the upload of call i+1 and the download
 * of call i use the link's two directions at once and the kernels of call i+1 start as soon as their operands are resident.
 * A `dev_scratch` of k (2 to 4) times what the *_scratch_size function reports is used as k parts that consecutive calls
 * take in turn (pass k times the size of the largest call of the sequence): a call then waits only for the calls that
 * used the same bytes, k calls earlier; with the plain size call i+1 starts when call i's results have left the device.
 * The same `dev_scratch` may be passed to every call of a sequence, calls of different shapes included (the library
 * tracks the byte ranges in flight).  mixq_host_drain makes `stream` wait for everything those calls queued, synchronises
 * it and closes the sequence; a call without the flag drains first by itself.  State is per calling thread. */
int mixq_host_drain(void* stream);

/* Kernel launches issued by this library since load (all threads); bench.py
 * reads it to fill `gpu_launches`. */
uint64_t mixq_launch_count(void);

/* Debug only: device buffer of 8 x uint64 %globaltimer stamps per CTA written by the stream-K GEMM
 * kernel (entry, prologue done, first TMA, first tile issued, MMA done, first/last accumulator ready,
 * epilogue done); NULL (default) disables it. */
int mixq_debug_set_trace(void* dev_buf);

/* Debug only (host arithmetic, no device needed): the schedule the fat-tile decode kernel would use for an M x N
 * problem on `pairs` CTA pairs (N = output channels of ONE projection; gated != 0: gate and up side by side),
 * out5 = {tile width Nt in accumulator columns, column tiles, row tiles, ring stages, waves}.  tests/ checks its
 * invariants (Nt + Nt/2 <= 512 TMEM columns, the tiles cover N, the ring fits shared memory) over random shapes. */
int mixq_debug_fat_plan(int64_t M, int64_t N, int pairs, int gated, int epi_warps, int* out5);

/* Debug only (host arithmetic): byte offset inside `dev_scratch` of the part the `call_index`-th queued host-buffer
 * call of a sequence takes, for a call that needs `need` bytes of a scratch of `scratch_bytes`; -1 if it does not fit. */
int64_t mixq_debug_host_part_offset(size_t need, size_t scratch_bytes, unsigned call_index);

/* ---- per-call tuning ------------------------------------------------------------------------------------------
 * The library keeps no mutable process-wide state (SURVEY.md 8b "no static mutable state"): a caller that wants a
 * particular tile configuration, or wants to leave SMs free for a concurrent communication kernel, says so on the
 * call.  Nothing in mixq_options changes a result bit.  opt == NULL (and every entry point without an `opt`
 * argument) means {0, 0}. */
typedef struct mixq_options {
    int gemm_config; /* 0 = pick by shape; tile ids are listed in DESIGN.md (tests pin every id against the oracle);
                        + 100 x s (s = 2, 4, 8) on the one-CTA ids 1 / 3 / 15 and M <= 128 splits K over a cluster of s CTAs */
    int sm_limit;    /* SMs the persistent kernels may occupy, 0 = all of the current device                    */
} mixq_options;
/* mixq_enqueue_ex / mixq_gemm_dequant_ws+_ex / mixq_enqueue_allreduce / mixq_gemm_dequant_allreduce with options.
 * Unknown config ids fail with MIXQ_ERR_BAD_ARG. */
size_t mixq_workspace_size_opt(int64_t M, int64_t N, int64_t K, const mixq_options* opt);
int mixq_enqueue_opt(const mixq_tensors* t, int64_t M, int64_t N, int64_t K, void* workspace,
                     size_t workspace_bytes, const mixq_epilogue* epi, const mixq_options* opt,
                     unsigned flags, void* stream);
int mixq_gemm_dequant_opt(const void* A8, const void* W8, const void* scale_a, const void* scale_b,
                          const void* fp_A, const void* fp_weight, void* Out, int64_t M, int64_t N,
                          int64_t K, const mixq_epilogue* epi, const mixq_options* opt, void* workspace,
                          size_t workspace_bytes, void* stream);
int mixq_enqueue_allreduce_opt(const mixq_tensors* t, int64_t M, int64_t N, int64_t K, void* workspace,
                               size_t workspace_bytes, const mixq_peer_group* g, const mixq_options* opt,
                               unsigned flags, void* stream);
int mixq_gemm_dequant_allreduce_opt(const void* A8, const void* W8, const void* scale_a,
                                    const void* scale_b, const void* fp_A, const void* fp_weight,
                                    int64_t M, int64_t N, int64_t K, const mixq_peer_group* g,
                                    const mixq_options* opt, void* stream);
/* Scratch stage 2 needs to run the decode-batch kernel (M <= 1024: split-K partial sums + the fp16 outlier product
 * of every tile): mixq_gemm_workspace_size() plus an (M, N)-dependent part (31.4 MB + 2 * Mpad256 * Npad256
 * bytes).  Only the opt-in configurations 8 / 11 / 12 read it. */
size_t mixq_decode_workspace_size(int64_t M, int64_t N);

/* ---- TensorRT plugin surface through C handles (for ctypes / C callers) ----
 * Mirrors MixQPluginCreator / MixQPlugin (TsinghuaMixQPlugin.h:34-115). */
typedef struct mixq_plugin_s mixq_plugin_t;

/* Same symbol and signature as the reference (MixQPlugins.cpp:126-132); loaded by
 * ctypes.CDLL(...).initOpenAiTritonPlugins(None, b"tensorrt_llm") in plugin.py:34-43. */
bool initOpenAiTritonPlugins(void* logger, const char* libNamespace);

/* creator.createPlugin(name, {m,n,k}) -- TsinghuaMixQPlugin.cpp:895-933. NULL if the
 * ("MixQ","1",ns) creator is not registered. */
mixq_plugin_t* mixq_plugin_create(const char* ns, int m, int n, int k);
/* creator.deserializePlugin -- TsinghuaMixQPlugin.cpp:935-952 (12 bytes: mm,mn,mk int32 LE). */
mixq_plugin_t* mixq_plugin_deserialize(const char* ns, const void* data, size_t len);
mixq_plugin_t* mixq_plugin_clone(const mixq_plugin_t* p);
void mixq_plugin_destroy(mixq_plugin_t* p);
const char* mixq_plugin_type(const mixq_plugin_t* p);      /* "MixQ" */
const char* mixq_plugin_version(const mixq_plugin_t* p);   /* "1"    */
const char* mixq_plugin_namespace(const mixq_plugin_t* p);
int mixq_plugin_nb_outputs(const mixq_plugin_t* p);
size_t mixq_plugin_serialization_size(const mixq_plugin_t* p);
void mixq_plugin_serialize(const mixq_plugin_t* p, void* buffer);
/* supportsFormatCombination(pos, ...) with type/format codes of nvinfer1 (kHALF=1, kLINEAR=0). */
int mixq_plugin_supports_format(const mixq_plugin_t* p, int pos, int dtype_code, int format_code);
/* configurePlugin + getWorkspaceSize for the given maxima. */
size_t mixq_plugin_workspace_size(mixq_plugin_t* p, const int64_t* a_max_dims, int a_nb_dims,
                                  int64_t n);
/* enqueue(inputDesc, outputDesc, inputs, outputs, workspace, stream):
 * a_dims/a_nb_dims describe input 0 ([..., K]); w_dim0 = inputDesc[1].dims.d[0] = N. */
int mixq_plugin_enqueue(mixq_plugin_t* p, const int64_t* a_dims, int a_nb_dims, int64_t w_dim0,
                        const void* const* inputs, void* const* outputs, void* workspace,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MIXQ_B200_H_ */
