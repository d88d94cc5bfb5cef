// mixq_internal.h -- declarations shared by the translation units of libmixq_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/mixq_b200.h"

namespace mixq {

struct DeviceInfo {
    bool ok = false;         // a CUDA device with compute capability 10.x is current
    int device = -1;
    int cc_major = 0, cc_minor = 0;
    int num_sms = 0;
    size_t max_smem_optin = 0;  // per-block opt-in limit (227 KB on B200)
    size_t smem_per_sm = 0;
};

// Properties of the CURRENT device (cudaGetDevice at the time of the call), cached per device ordinal:
// a process that drives several GPUs gets each one's own SM count and its own one-time function attributes.
const DeviceInfo& device_info();
constexpr int kMaxDevices = 64;

// Per-call tuning (mixq_options in the C ABI).  Nothing here changes a result bit; defaults = automatic.
struct LaunchOpts {
    int cfg = 0;        // GemmConfig id, 0 = auto
    int sm_limit = 0;   // SMs the persistent kernels may occupy, 0 = all
};
inline LaunchOpts make_opts(const mixq_options* o) { return o ? LaunchOpts{o->gemm_config, o->sm_limit} : LaunchOpts{}; }

// Error bookkeeping (thread-local text behind mixq_last_error()).
int set_error(int status, const char* msg);
int set_cuda_error(cudaError_t e, const char* what);
void count_launch();

// stage 1 (quant_extract.cu)
// clear_words: optional, `n_clear` 32-bit words zeroed by CTA 0 (the stream-K flags of kernel 2)
int launch_quant_extract(const void* A, int64_t M, int64_t K, const void* ind, int n_ind, void* A8, void* scale_a,
                         void* fp_A, unsigned flags, cudaStream_t stream, bool pdl, void* clear_words = nullptr,
                         int n_clear = 0, const void* gamma = nullptr, float eps = 0.0f, void* y_out = nullptr,
                         LaunchOpts opts = LaunchOpts{});

// stage 2 (gemm_i8_tcgen05.cu)
// sk_ws: optional scratch of the split-K schedules (flags first, then partial-sum slots, then the decode kernel's
// outlier-product area): streamk_workspace_bytes() serves config 8, decode_workspace_bytes(M, N) also the decode kernel;
// sk_flags_clean: the flags are already zero (mixq_enqueue lets kernel 1 clear them).
int launch_gemm_dequant(const void* A8, const void* W8, const void* scale_a, const void* scale_b, const void* fp_A,
                        const void* fp_weight, void* Out, int64_t M, int64_t N, int64_t K, cudaStream_t stream,
                        bool pdl, void* sk_ws = nullptr, size_t sk_ws_bytes = 0, bool sk_flags_clean = false,
                        const void* bias = nullptr, int act = 0,   // fused epilogue: see mixq_epilogue
                        LaunchOpts opts = LaunchOpts{});
// gate and up projections of one MLP over one quantised A: Out = fp16(silu(gate)) * fp16(up) (gemm_fat.cuh, gated mode);
// scratch (M*N*2 bytes) is only needed for M > 1024
int launch_gemm_dequant_gated(const void* A8, const void* scale_a, const void* fp_A, const void* W8_gate, const void* sb_gate,
                              const void* fpw_gate, const void* W8_up, const void* sb_up, const void* fpw_up, void* Out, int64_t M,
                              int64_t N, int64_t K, cudaStream_t stream, bool pdl, void* scratch, size_t scratch_bytes,
                              LaunchOpts opts = LaunchOpts{});
size_t streamk_workspace_bytes();
size_t decode_workspace_bytes(int64_t M, int64_t N);
// stage 2 with the row-parallel all-reduce fused in (peer memory; see ArParams in gemm_i8_tcgen05.cu)
int launch_gemm_dequant_allreduce(const void* A8, const void* W8, const void* scale_a, const void* scale_b,
                                  const void* fp_A, const void* fp_weight, int64_t M, int64_t N, int64_t K,
                                  const mixq_peer_group* pg, cudaStream_t stream, bool pdl, LaunchOpts opts = LaunchOpts{});
// decode-sized results on few ranks: every rank pulls all partials over peer loads and reduces in rank order (allreduce_pull.cu)
int launch_allreduce_pull(void* const* partials, void* const* counters, void* const* outs, int world, int rank, size_t n_elems,
                          bool two_shot, cudaStream_t stream, bool pdl, LaunchOpts opts = LaunchOpts{});
size_t allreduce_staging_bytes(int64_t M, int64_t N, int world);
size_t allreduce_counter_bytes(int64_t M, int64_t N, int world);
int set_trace_buffer(void* dev_buf);
// the fat-tile schedule of a decode batch (host arithmetic only): out = {Nt, n_tiles, m_tiles, stages, waves}
void fat_plan_for(int64_t M, int64_t N, int pairs, int gated, int epi_warps, int* out5);

// M <= 4 branch (gemv_w8a16.cu): weight-only GEMV over the EETQ-interleaved q_weight
int launch_gemv_w8a16(const void* A, const void* q_weight, const void* scales, void* Out, int64_t M, int64_t N, int64_t K,
                      cudaStream_t stream, const void* bias = nullptr, int act = 0);

// GEMM tile configuration ids (mixq_set_gemm_config); 0 = pick automatically.
enum GemmConfig { kCfgAuto = 0, kCfgN128x2 = 1, kCfgN256x1 = 2, kCfgN64x2 = 3, kCfg2CtaN256x1 = 4, kCfg2CtaN128x2 = 5, kCfg2CtaN256Stash = 6, kCfgN256Stash = 7, kCfg2CtaN256StreamK = 8, kCfg2CtaN256Tma = 9, kCfg2CtaN192Tma = 10, kCfg2CtaN256Decode = 11, kCfg2CtaN256DecodeNoSplit = 12, kCfg2CtaFat = 13, kCfg2CtaFatSplitK = 14, kCfgN32x2 = 15, kCfg2CtaFatEpi8 = 16 /* the fat tile with 8 epilogue warps (13: 12) */, kCfgCount,
                  kCfgGatedUnfused = 100 /* mixq_*_gated only: the two-GEMM + multiply composition for every M (tests) */ };
// SMs the persistent kernels may occupy on the current device (all, or the caller's per-call limit)
int usable_sms(const LaunchOpts& opts);

}  // namespace mixq
