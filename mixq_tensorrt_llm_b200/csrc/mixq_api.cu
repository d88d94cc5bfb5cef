// mixq_api.cu -- the extern "C" entry points of include/mixq_b200.h (everything except the
// plugin-handle functions, which live in mixq_plugin.cpp).  Host logic only: argument checks,
// workspace carving (reference TsinghuaMixQPlugin.cpp:406-421), the two launches.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "mixq_internal.h"

namespace mixq {

namespace {
thread_local char g_err[256] = "";
std::atomic<uint64_t> g_launches{0};   // statistics only (mixq_launch_count); no call reads it back

constexpr size_t kAlign = 128;  // kCudaMemAlign, TsinghuaMixQPlugin.cpp:204
inline size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

struct Carve {
    size_t off_a8, off_sa, off_fpa, off_sk, total;
};
// int8_out | scale_a | fp_activation, in the reference's order (TsinghuaMixQPlugin.cpp:410-421)
inline Carve carve(int64_t M, int64_t N, int64_t K) {
    Carve c;
    c.off_a8 = 0;
    c.off_sa = align_up(static_cast<size_t>(M) * static_cast<size_t>(K));
    c.off_fpa = c.off_sa + align_up(static_cast<size_t>(M) * 2);
    c.off_sk = c.off_fpa + align_up(static_cast<size_t>(M) * MIXQ_NUM_OUTLIERS * 2);
    c.total = c.off_sk;   // what the default path needs; the opt-in split-K configurations want decode_workspace_bytes() more
    (void)N;
    return c;
}
}  // namespace

int set_error(int status, const char* msg) {
    std::snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "");
    return status;
}
int set_cuda_error(cudaError_t e, const char* what) {
    std::snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return MIXQ_ERR_CUDA;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int usable_sms(const LaunchOpts& opts) {
    const int n = device_info().num_sms;
    return (opts.sm_limit > 0 && opts.sm_limit < n) ? opts.sm_limit : n;
}

const DeviceInfo& device_info() {
    // one entry per device ordinal, filled on first use with that device current
    static DeviceInfo infos[kMaxDevices];
    static std::once_flag once[kMaxDevices];
    static const DeviceInfo none;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        cudaGetLastError();
        return none;
    }
    std::call_once(once[dev], [dev] {
        DeviceInfo& info = infos[dev];
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        info.device = dev;
        info.cc_major = p.major;
        info.cc_minor = p.minor;
        info.num_sms = p.multiProcessorCount;
        info.max_smem_optin = p.sharedMemPerBlockOptin;
        info.smem_per_sm = p.sharedMemPerMultiprocessor;
        info.ok = (p.major == 10);  // the cubin is sm_100a only
    });
    return infos[dev];
}

}  // namespace mixq

using namespace mixq;

extern "C" {

const char* mixq_version(void) { return "mixq-b200 0.1.0 (sm_100a; tcgen05 W8A8O16)"; }
const char* mixq_last_error(void) { return g_err; }
int mixq_device_ok(void) { return device_info().ok ? 1 : 0; }
uint64_t mixq_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

size_t mixq_workspace_size(int64_t M, int64_t N, int64_t K) {
    if (M <= 0 || K <= 0) return 0;
    return carve(M, N < 0 ? 0 : N, K).total + kAlign;  // + slack to align the base like nextWorkspacePtr does
}
size_t mixq_decode_workspace_size(int64_t M, int64_t N) { return decode_workspace_bytes(M, N); }
size_t mixq_workspace_size_opt(int64_t M, int64_t N, int64_t K, const mixq_options* opt) {
    const size_t base = mixq_workspace_size(M, N, K);
    if (!base || !opt) return base;
    const int c = opt->gemm_config;
    const bool scratch = c == kCfg2CtaN256StreamK || c == kCfg2CtaN256Decode || c == kCfg2CtaN256DecodeNoSplit;
    return base + (scratch ? align_up(decode_workspace_bytes(M, N < 0 ? 0 : N)) : 0);
}

int mixq_quant_extract(const void* A, int64_t M, int64_t K, const void* ind, int n_ind, void* A8, void* scale_a,
                       void* fp_A, unsigned flags, void* stream) {
    if (M < 0) return set_error(MIXQ_ERR_BAD_ARG, "quant_extract: M < 0");
    return launch_quant_extract(A, M, K, ind, n_ind, A8, scale_a, fp_A, flags, static_cast<cudaStream_t>(stream),
                                /*pdl=*/false);
}

int mixq_rmsnorm_quant_extract(const void* X, const void* gamma, float eps, int64_t M, int64_t K, const void* ind,
                               int n_ind, void* A8, void* scale_a, void* fp_A, void* Y, unsigned flags, void* stream) {
    if (M < 0) return set_error(MIXQ_ERR_BAD_ARG, "rmsnorm_quant_extract: M < 0");
    if (!gamma) return set_error(MIXQ_ERR_BAD_ARG, "rmsnorm_quant_extract: gamma is null");
    return launch_quant_extract(X, M, K, ind, n_ind, A8, scale_a, fp_A, flags, static_cast<cudaStream_t>(stream),
                                /*pdl=*/false, nullptr, 0, gamma, eps, Y);
}

int mixq_gemm_dequant(const void* A8, const void* W8, const void* scale_a, const void* scale_b, const void* fp_A,
                      const void* fp_weight, void* Out, int64_t M, int64_t N, int64_t K, void* stream) {
    return launch_gemm_dequant(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K,
                               static_cast<cudaStream_t>(stream), /*pdl=*/false);
}

int mixq_gemv_w8a16(const void* A, const void* q_weight, const void* scales, void* Out, int64_t M, int64_t N, int64_t K,
                    void* stream) {
    return launch_gemv_w8a16(A, q_weight, scales, Out, M, N, K, static_cast<cudaStream_t>(stream));
}

int mixq_debug_set_trace(void* dev_buf) { return set_trace_buffer(dev_buf); }
int mixq_debug_fat_plan(int64_t M, int64_t N, int pairs, int gated, int epi_warps, int* out5) {
    if (!out5 || M <= 0 || N <= 0 || pairs <= 0 || (epi_warps != 8 && epi_warps != 12))
        return set_error(MIXQ_ERR_BAD_ARG, "debug_fat_plan: bad argument");
    fat_plan_for(M, N, pairs, gated, epi_warps, out5);
    return MIXQ_OK;
}

size_t mixq_gemm_workspace_size(void) { return streamk_workspace_bytes(); }

int mixq_gemm_dequant_ws(const void* A8, const void* W8, const void* scale_a, const void* scale_b, const void* fp_A,
                         const void* fp_weight, void* Out, int64_t M, int64_t N, int64_t K, void* workspace,
                         size_t workspace_bytes, void* stream) {
    return launch_gemm_dequant(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K,
                               static_cast<cudaStream_t>(stream), /*pdl=*/false, workspace, workspace_bytes, false);
}

int mixq_enqueue(const mixq_tensors* t, int64_t M, int64_t N, int64_t K, void* workspace, size_t workspace_bytes,
                 unsigned flags, void* stream) {
    return mixq_enqueue_opt(t, M, N, K, workspace, workspace_bytes, nullptr, nullptr, flags, stream);
}

int mixq_gemm_dequant_opt(const void* A8, const void* W8, const void* scale_a, const void* scale_b, const void* fp_A,
                          const void* fp_weight, void* Out, int64_t M, int64_t N, int64_t K, const mixq_epilogue* epi,
                          const mixq_options* opt, void* workspace, size_t workspace_bytes, void* stream) {
    return launch_gemm_dequant(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, static_cast<cudaStream_t>(stream),
                               /*pdl=*/false, workspace, workspace_bytes, false, epi ? epi->bias : nullptr,
                               epi ? epi->activation : 0, make_opts(opt));
}

int mixq_gemm_dequant_ex(const void* A8, const void* W8, const void* scale_a, const void* scale_b, const void* fp_A,
                         const void* fp_weight, void* Out, int64_t M, int64_t N, int64_t K, const mixq_epilogue* epi,
                         void* stream) {
    return launch_gemm_dequant(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, static_cast<cudaStream_t>(stream),
                               /*pdl=*/false, nullptr, 0, false, epi ? epi->bias : nullptr, epi ? epi->activation : 0);
}

int mixq_enqueue_ex(const mixq_tensors* t, int64_t M, int64_t N, int64_t K, void* workspace, size_t workspace_bytes,
                    const mixq_epilogue* epi, unsigned flags, void* stream) {
    return mixq_enqueue_opt(t, M, N, K, workspace, workspace_bytes, epi, nullptr, flags, stream);
}

int mixq_enqueue_opt(const mixq_tensors* t, int64_t M, int64_t N, int64_t K, void* workspace, size_t workspace_bytes,
                     const mixq_epilogue* epi, const mixq_options* opt, unsigned flags, void* stream) {
    const LaunchOpts lo = make_opts(opt);
    const void* bias = epi ? epi->bias : nullptr;
    const int act = epi ? epi->activation : 0;
    if (act != MIXQ_ACT_NONE && act != MIXQ_ACT_SILU) return set_error(MIXQ_ERR_BAD_ARG, "enqueue: unknown activation");
    if (!t) return set_error(MIXQ_ERR_BAD_ARG, "enqueue: null tensor table");
    if (M < 0 || N <= 0 || K <= 0) return set_error(MIXQ_ERR_BAD_ARG, "enqueue: bad dimensions");
    if (M == 0) return MIXQ_OK;
    // decode with at most 4 tokens: the reference switches to the weight-only GEMV over q_weight
    // (TsinghuaMixQPlugin.cpp:472, 641-647).  Taken when the caller provides that second weight copy.
    if (M <= 4 && t->q_weight && !(flags & MIXQ_FLAG_FORCE_MIXED)) {
        if (!t->A || !t->scaling_factors || !t->Out)
            return set_error(MIXQ_ERR_BAD_ARG, "enqueue: null tensor (A, q_weight, scaling_factors and Out are required for M <= 4)");
        return launch_gemv_w8a16(t->A, t->q_weight, t->scaling_factors, t->Out, M, N, K, static_cast<cudaStream_t>(stream), bias, act);
    }
    if (!t->A || !t->W8 || !t->scale_b || !t->fp_weight || !t->ind || !t->Out)
        return set_error(MIXQ_ERR_BAD_ARG, "enqueue: null tensor (A, W8, scale_b, fp_weight, ind and Out are required)");
    if (!workspace) return set_error(MIXQ_ERR_WORKSPACE, "enqueue: null workspace");
    const Carve c = carve(M, N, K);
    // align the base the way nextWorkspacePtr(ptr, 0) does (TsinghuaMixQPlugin.cpp:206-215)
    uintptr_t base = reinterpret_cast<uintptr_t>(workspace);
    const uintptr_t aligned = (base + kAlign - 1) / kAlign * kAlign;
    if (workspace_bytes < (aligned - base) + c.total) return set_error(MIXQ_ERR_WORKSPACE, "enqueue: workspace too small");
    uint8_t* ws = reinterpret_cast<uint8_t*>(aligned);
    void* A8 = ws + c.off_a8;
    void* sa = ws + c.off_sa;
    void* fpA = ws + c.off_fpa;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // split-K scratch of the opt-in configurations 8 / 11 / 12: present only when the caller sized the workspace with
    // mixq_workspace_size_opt; the quantise kernel then also clears its flags
    const size_t sk_bytes = decode_workspace_bytes(M, N);
    void* sk = workspace_bytes >= (aligned - base) + c.total + sk_bytes ? ws + c.off_sk : nullptr;
    int rc = launch_quant_extract(t->A, M, K, t->ind, MIXQ_NUM_OUTLIERS, A8, sa, fpA, flags, s, /*pdl=*/true, sk, sk ? 1024 : 0, nullptr,
                                  0.0f, nullptr, lo);
    if (rc) return rc;
    return launch_gemm_dequant(A8, t->W8, sa, t->scale_b, fpA, t->fp_weight, t->Out, M, N, K, s, /*pdl=*/true, sk,
                               sk ? sk_bytes : 0, /*sk_flags_clean=*/sk != nullptr, bias, act, lo);
}

size_t mixq_gated_workspace_size(int64_t M, int64_t N, int64_t K) {
    if (M <= 0 || K <= 0 || N <= 0) return 0;
    return mixq_workspace_size(M, N, K) + align_up(static_cast<size_t>(M) * N * 2);
}

int mixq_gemm_dequant_gated(const void* A8, const void* scale_a, const void* fp_A, const void* W8_gate, const void* scale_b_gate,
                            const void* fp_weight_gate, const void* W8_up, const void* scale_b_up, const void* fp_weight_up, void* Out,
                            int64_t M, int64_t N, int64_t K, const mixq_options* opt, void* scratch, size_t scratch_bytes, void* stream) {
    return launch_gemm_dequant_gated(A8, scale_a, fp_A, W8_gate, scale_b_gate, fp_weight_gate, W8_up, scale_b_up, fp_weight_up, Out, M,
                                     N, K, static_cast<cudaStream_t>(stream), /*pdl=*/false, scratch, scratch_bytes, make_opts(opt));
}

int mixq_enqueue_gated(const mixq_tensors* gate, const mixq_tensors* up, int64_t M, int64_t N, int64_t K, void* workspace,
                       size_t workspace_bytes, const mixq_options* opt, unsigned flags, void* stream) {
    const LaunchOpts lo = make_opts(opt);
    if (!gate || !up) return set_error(MIXQ_ERR_BAD_ARG, "enqueue_gated: null tensor table");
    if (M < 0 || N <= 0 || K <= 0) return set_error(MIXQ_ERR_BAD_ARG, "enqueue_gated: bad dimensions");
    if (M == 0) return MIXQ_OK;
    if (!gate->A || !gate->W8 || !gate->scale_b || !gate->fp_weight || !gate->ind || !gate->Out || !up->W8 || !up->scale_b || !up->fp_weight)
        return set_error(MIXQ_ERR_BAD_ARG, "enqueue_gated: null tensor (gate: A, W8, scale_b, fp_weight, ind, Out; up: W8, scale_b, fp_weight)");
    if (up->A && up->A != gate->A) return set_error(MIXQ_ERR_BAD_ARG, "enqueue_gated: gate and up must read the same activations");
    if (!workspace) return set_error(MIXQ_ERR_WORKSPACE, "enqueue_gated: null workspace");
    const Carve c = carve(M, N, K);
    uintptr_t base = reinterpret_cast<uintptr_t>(workspace);
    const uintptr_t aligned = (base + kAlign - 1) / kAlign * kAlign;
    if (workspace_bytes < (aligned - base) + c.total) return set_error(MIXQ_ERR_WORKSPACE, "enqueue_gated: workspace too small");
    // the [M, N] scratch of the two-GEMM composition (M > 1024) follows the plugin workspace when the caller provided it
    size_t extra = align_up(static_cast<size_t>(M) * N * 2);
    if (workspace_bytes < (aligned - base) + c.total + extra) extra = 0;
    uint8_t* ws = reinterpret_cast<uint8_t*>(aligned);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = launch_quant_extract(gate->A, M, K, gate->ind, MIXQ_NUM_OUTLIERS, ws + c.off_a8, ws + c.off_sa, ws + c.off_fpa, flags, s,
                                  /*pdl=*/true, nullptr, 0, nullptr, 0.0f, nullptr, lo);
    if (rc) return rc;
    return launch_gemm_dequant_gated(ws + c.off_a8, ws + c.off_sa, ws + c.off_fpa, gate->W8, gate->scale_b, gate->fp_weight, up->W8,
                                     up->scale_b, up->fp_weight, gate->Out, M, N, K, s, /*pdl=*/true, extra ? ws + c.total : nullptr, extra,
                                     lo);
}

size_t mixq_allreduce_staging_size(int64_t M, int64_t N, int world) { return allreduce_staging_bytes(M, N, world); }
size_t mixq_allreduce_counter_size(int64_t M, int64_t N, int world) { return allreduce_counter_bytes(M, N, world); }

int mixq_allreduce_check(void* counters_local, int clear, void* stream) {
    if (!counters_local) return set_error(MIXQ_ERR_BAD_ARG, "allreduce_check: null counter block");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint32_t word = 0;
    uint32_t* dev_word = static_cast<uint32_t*>(counters_local) + 4;   // kArWordError
    cudaError_t e = cudaMemcpyAsync(&word, dev_word, sizeof(word), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess && word && clear) e = cudaMemsetAsync(dev_word, 0, sizeof(word), s);
    if (e != cudaSuccess) return set_cuda_error(e, "allreduce_check");
    return word ? set_error(MIXQ_ERR_CUDA, "fused all-reduce: a peer did not arrive within MIXQ_AR_TIMEOUT_MS; the result of that call is invalid") : MIXQ_OK;
}

int mixq_gemm_dequant_allreduce(const void* A8, const void* W8, const void* scale_a, const void* scale_b, const void* fp_A,
                                const void* fp_weight, int64_t M, int64_t N, int64_t K, const mixq_peer_group* g, void* stream) {
    return mixq_gemm_dequant_allreduce_opt(A8, W8, scale_a, scale_b, fp_A, fp_weight, M, N, K, g, nullptr, stream);
}
namespace {
// Decode-sized result: the GEMM (any tile configuration) writes this rank's partial into its staging area and a small kernel
// pulls every rank's partial and reduces in rank order (allreduce_pull.cu) -- same arithmetic as the one-kernel path,
// bit-identical results.  Every rank then RECEIVES (world - 1) partials, so it pays while that ingress is small: up to 4 MB
// (measured: 2 ranks 4 MB 1.3x faster in the graph-replayed decode step, 4 ranks 12 MB 1.2x slower, 4 ranks 0.75 MB faster),
// 8 MB at 2 ranks (equal).  gemm_config 9 keeps the one-kernel path; MIXQ_PULL_MAX_INGRESS_MB overrides the limit.
// returns 0 = one-kernel path, 1 = one-shot pull, 2 = two-shot pull
int use_pull(const mixq_peer_group* g, int64_t M, int64_t N, const LaunchOpts& lo) {
    static const long long max_ingress = [] {
        const char* e = std::getenv("MIXQ_PULL_MAX_INGRESS_MB");
        return (e ? std::atoll(e) : 4ll) << 20;
    }();
    // two-shot pull (each rank reduces its slice, then copies the others'): bit-exact and tested, but measured slower than the
    // one-kernel path in the graph-replayed decode step on 4 ranks (1457 vs 1873 TFLOPS), so it is opt-in: results up to
    // MIXQ_PULL_TWO_SHOT_MAX_MB megabytes take it
    static const long long max_two_shot = [] {
        const char* e = std::getenv("MIXQ_PULL_TWO_SHOT_MAX_MB");
        return (e ? std::atoll(e) : 0ll) << 20;
    }();
    if (!g || g->world < 2 || g->world > MIXQ_MAX_RANKS || g->rank < 0 || g->rank >= g->world || lo.cfg == kCfg2CtaN256Tma) return 0;
    const long long bytes = static_cast<long long>(M) * N * 2;
    if (static_cast<size_t>(bytes) > g->staging_bytes) return 0;
    if (bytes * (g->world - 1) <= max_ingress || (g->world == 2 && bytes <= 2 * max_ingress)) return 1;
    return bytes <= max_two_shot ? 2 : 0;
}
int pull_reduce(const mixq_peer_group* g, int64_t M, int64_t N, int mode, cudaStream_t s, const LaunchOpts& lo) {
    for (int i = 0; i < g->world; ++i)
        if (!g->out[i] || !g->staging[i] || !g->counters[i]) return set_error(MIXQ_ERR_BAD_ARG, "enqueue_allreduce: null peer pointer");
    return launch_allreduce_pull(g->staging, g->counters, g->out, g->world, g->rank, static_cast<size_t>(M) * N, mode == 2, s, /*pdl=*/true, lo);
}
}  // namespace

int mixq_gemm_dequant_allreduce_opt(const void* A8, const void* W8, const void* scale_a, const void* scale_b, const void* fp_A,
                                    const void* fp_weight, int64_t M, int64_t N, int64_t K, const mixq_peer_group* g,
                                    const mixq_options* opt, void* stream) {
    const LaunchOpts lo0 = make_opts(opt);
    if (const int mode = (M > 0 && N > 0) ? use_pull(g, M, N, lo0) : 0) {
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const int rc = launch_gemm_dequant(A8, W8, scale_a, scale_b, fp_A, fp_weight, g->staging[g->rank], M, N, K, s, /*pdl=*/false, nullptr,
                                           0, false, nullptr, 0, lo0);
        return rc ? rc : pull_reduce(g, M, N, mode, s, lo0);
    }
    return launch_gemm_dequant_allreduce(A8, W8, scale_a, scale_b, fp_A, fp_weight, M, N, K, g,
                                         static_cast<cudaStream_t>(stream), /*pdl=*/false, make_opts(opt));
}

int mixq_enqueue_allreduce(const mixq_tensors* t, int64_t M, int64_t N, int64_t K, void* workspace, size_t workspace_bytes,
                           const mixq_peer_group* g, unsigned flags, void* stream) {
    return mixq_enqueue_allreduce_opt(t, M, N, K, workspace, workspace_bytes, g, nullptr, flags, stream);
}
int mixq_enqueue_allreduce_opt(const mixq_tensors* t, int64_t M, int64_t N, int64_t K, void* workspace, size_t workspace_bytes,
                               const mixq_peer_group* g, const mixq_options* opt, unsigned flags, void* stream) {
    const LaunchOpts lo = make_opts(opt);
    if (!t || !g) return set_error(MIXQ_ERR_BAD_ARG, "enqueue_allreduce: null tensor table / peer group");
    if (M < 0 || N <= 0 || K <= 0) return set_error(MIXQ_ERR_BAD_ARG, "enqueue_allreduce: bad dimensions");
    if (M == 0) return MIXQ_OK;
    if (!t->A || !t->W8 || !t->scale_b || !t->fp_weight || !t->ind)
        return set_error(MIXQ_ERR_BAD_ARG, "enqueue_allreduce: null tensor (A, W8, scale_b, fp_weight and ind are required)");
    if (!workspace) return set_error(MIXQ_ERR_WORKSPACE, "enqueue_allreduce: null workspace");
    const Carve c = carve(M, N, K);
    uintptr_t base = reinterpret_cast<uintptr_t>(workspace);
    const uintptr_t aligned = (base + kAlign - 1) / kAlign * kAlign;
    if (workspace_bytes < (aligned - base) + c.total) return set_error(MIXQ_ERR_WORKSPACE, "enqueue_allreduce: workspace too small");
    uint8_t* ws = reinterpret_cast<uint8_t*>(aligned);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = launch_quant_extract(t->A, M, K, t->ind, MIXQ_NUM_OUTLIERS, ws + c.off_a8, ws + c.off_sa, ws + c.off_fpa, flags, s,
                                  /*pdl=*/true, nullptr, 0, nullptr, 0.0f, nullptr, lo);
    if (rc) return rc;
    if (const int mode = use_pull(g, M, N, lo)) {
        rc = launch_gemm_dequant(ws + c.off_a8, t->W8, ws + c.off_sa, t->scale_b, ws + c.off_fpa, t->fp_weight, g->staging[g->rank], M, N, K, s,
                                 /*pdl=*/true, nullptr, 0, false, nullptr, 0, lo);
        return rc ? rc : pull_reduce(g, M, N, mode, s, lo);
    }
    return launch_gemm_dequant_allreduce(ws + c.off_a8, t->W8, ws + c.off_sa, t->scale_b, ws + c.off_fpa, t->fp_weight, M, N, K, g,
                                         s, /*pdl=*/true, lo);
}

size_t mixq_linears_host_scratch_size(int64_t M, const int64_t* N, int count, int64_t K) {
    if (M <= 0 || K <= 0 || count <= 0 || !N) return 0;
    size_t total = align_up(static_cast<size_t>(M) * K * 2), ws = 0;
    for (int i = 0; i < count; ++i) {
        if (N[i] <= 0) return 0;
        total += align_up(static_cast<size_t>(M) * N[i] * 2);
        const size_t w = mixq_workspace_size(M, N[i], K);
        ws = w > ws ? w : ws;
    }
    return total + ws + kAlign;
}
size_t mixq_host_scratch_size(int64_t M, int64_t N, int64_t K) { return mixq_linears_host_scratch_size(M, &N, 1, K); }

}  // extern "C" (part 1)

// Host-buffer path.  The rows are cut into slabs and pipelined over three streams -- H2D of slab c+1, the two
// kernels of slab c (on the caller's stream) and D2H of slab c-1 run concurrently -- so a call costs
// max(H2D, D2H) over PCIe instead of their sum plus the compute.
namespace {
struct HostPipe {
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaEvent_t entry = nullptr, up[64] = {}, done[64] = {}, drained = nullptr;
    // MIXQ_FLAG_HOST_ASYNC: the last kInFlight asynchronous calls, each with the byte range of the scratch it uses and two
    // events: `read` = its kernels have run (its activations and workspace may be overwritten), `drained` = its results
    // have left the device.  A new call waits for every recorded call whose range overlaps its own.
    static constexpr int kInFlight = 8;
    struct InFlight {
        uintptr_t begin = 0, end = 0;
        cudaEvent_t read = nullptr, drained = nullptr;
        bool used = false;
    } inflight[kInFlight];
    unsigned n_async = 0;   // asynchronous calls since the last drain: picks the part of the scratch and the record
    bool pending = false;   // asynchronous calls issued since the last mixq_host_drain
    bool ok = false;
    HostPipe() {
        ok = cudaStreamCreateWithFlags(&h2d, cudaStreamNonBlocking) == cudaSuccess &&
             cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&entry, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&drained, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; ok && i < kInFlight; ++i)
            ok = cudaEventCreateWithFlags(&inflight[i].read, cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&inflight[i].drained, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; ok && i < 64; ++i)
            ok = cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) == cudaSuccess;
    }
};
HostPipe& host_pipe() {
    static thread_local HostPipe p;  // streams/events are cheap to keep; one set per calling thread
    return p;
}
}  // namespace

namespace {
// Queued host-buffer calls: a scratch of k (2..4) times the call's size is cut into k equal parts, taken in turn
size_t host_part_offset(size_t need, size_t scratch_bytes, unsigned call_index) {
    size_t k = scratch_bytes / align_up(need);
    k = k > 4 ? 4 : k;
    return k >= 2 ? (call_index % k) * (scratch_bytes / k / kAlign * kAlign) : 0;
}

// gated: t[0] / t[1] are the gate / up projections of one MLP, ONE output Out_host[0] [M, N[0]] (mixq_enqueue_gated per slab)
int linears_host_impl(const mixq_tensors* const* t, int count, const void* A_host, void* const* Out_host, int64_t M,
                      const int64_t* N, int64_t K, void* dev_scratch, size_t dev_scratch_bytes, unsigned flags, void* stream,
                      bool gated) {
    const int n_tab = gated ? 2 : count;
    if (!t || !N || !Out_host || !A_host || !dev_scratch || count <= 0 || count > 8)
        return set_error(MIXQ_ERR_BAD_ARG, "linears_host: null pointer / count not in [1, 8]");
    if (M <= 0 || K <= 0) return set_error(MIXQ_ERR_BAD_ARG, "linears_host: bad dimensions");
    for (int i = 0; i < n_tab; ++i)
        if (!t[i]) return set_error(MIXQ_ERR_BAD_ARG, "linears_host: null tensor table");
    for (int i = 0; i < count; ++i)
        if (!Out_host[i] || N[i] <= 0) return set_error(MIXQ_ERR_BAD_ARG, "linears_host: null output or bad N");
    const size_t need = gated ? mixq_gated_host_scratch_size(M, N[0], K) : mixq_linears_host_scratch_size(M, N, count, K);
    if (dev_scratch_bytes < need) return set_error(MIXQ_ERR_WORKSPACE, "linears_host: scratch too small");
    if (!device_info().ok) return set_error(MIXQ_ERR_CUDA, "no usable sm_100 device");
    HostPipe& hp = host_pipe();
    if (!hp.ok) return set_error(MIXQ_ERR_CUDA, "linears_host: could not create the copy streams");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // Asynchronous form: nothing is waited for here.  A scratch of k (up to 4) times the call's size is used as k parts that
    // consecutive calls take in turn, so the upload of a call and the downloads of the calls before it share the PCIe
    // link's two directions while its kernels wait only for their own operands.  Calls of different shapes may cut the same
    // scratch differently: what a call waits for is decided by byte ranges, not by part numbers.
    const bool async = (flags & MIXQ_FLAG_HOST_ASYNC) != 0;
    flags &= ~static_cast<unsigned>(MIXQ_FLAG_HOST_ASYNC);
    const size_t part_off = async ? host_part_offset(need, dev_scratch_bytes, hp.n_async) : 0;
    uintptr_t base = (reinterpret_cast<uintptr_t>(dev_scratch) + part_off + kAlign - 1) / kAlign * kAlign;
    const uintptr_t range_begin = base, range_end = base + need - kAlign;   // `need` carries kAlign bytes of slack for the alignment of base
    uint8_t* dA = reinterpret_cast<uint8_t*>(base);
    uint8_t* dOut[8];
    uint8_t* cur = dA + align_up(static_cast<size_t>(M) * K * 2);
    size_t ws_bytes = 0;
    for (int i = 0; i < count; ++i) {
        dOut[i] = cur;
        cur += align_up(static_cast<size_t>(M) * N[i] * 2);
        const size_t w = gated ? mixq_gated_workspace_size(M, N[i], K) : mixq_workspace_size(M, N[i], K);
        ws_bytes = w > ws_bytes ? w : ws_bytes;
    }
    uint8_t* ws = cur;

    // ~32 slabs of >= 1024 rows (a multiple of the 256-row tile), at most 64 of them: the call is PCIe-bound, so short
    // slabs (small fill / drain bubbles of the three-stage pipeline) matter more than GEMM efficiency per slab
    int64_t rows = (M + 31) / 32;
    rows = rows < 1024 ? 1024 : (rows + 255) / 256 * 256;
    while ((M + rows - 1) / rows > 64) rows *= 2;
    const int n_slabs = static_cast<int>((M + rows - 1) / rows);

    cudaError_t e = cudaSuccess;
    HostPipe::InFlight* rec = nullptr;
    if (async) {
        // earlier asynchronous calls whose bytes this call reuses, possibly of another shape (their Out area may lie where this
        // call's activations go): the upload waits for their kernels AND their downloads, the kernels for their downloads.
        // The record about to be recycled is the oldest one: waited for unconditionally.
        rec = &hp.inflight[hp.n_async % HostPipe::kInFlight];
        for (int i = 0; i < HostPipe::kInFlight && e == cudaSuccess; ++i) {
            HostPipe::InFlight& f = hp.inflight[i];
            if (!f.used || (&f != rec && (f.end <= range_begin || range_end <= f.begin))) continue;
            e = cudaStreamWaitEvent(hp.h2d, f.read, 0);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(hp.h2d, f.drained, 0);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(s, f.drained, 0);
        }
    }
    if (async && hp.pending) {
        // nothing else to wait for: earlier work on the caller's stream was joined when the sequence began
    } else {
        if (hp.pending) {   // undrained asynchronous calls: their downloads may still read the scratch
            e = cudaEventRecord(hp.drained, hp.d2h);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(s, hp.drained, 0);
            if (e != cudaSuccess) return set_cuda_error(e, "linears_host: join earlier downloads");
        }
        e = cudaEventRecord(hp.entry, s);  // earlier work on the caller's stream may still use the scratch
        if (e == cudaSuccess) e = cudaStreamWaitEvent(hp.h2d, hp.entry, 0);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(hp.d2h, hp.entry, 0);
    }
    if (e != cudaSuccess) return set_cuda_error(e, "linears_host: stream setup");
    const uint8_t* hA = static_cast<const uint8_t*>(A_host);
    // the activations cross PCIe ONCE, whatever the number of linears that consume them
    for (int c = 0; c < n_slabs; ++c) {
        const int64_t r0 = c * rows, nr = (r0 + rows <= M) ? rows : M - r0;
        e = cudaMemcpyAsync(dA + r0 * K * 2, hA + r0 * K * 2, static_cast<size_t>(nr) * K * 2, cudaMemcpyHostToDevice, hp.h2d);
        if (e == cudaSuccess) e = cudaEventRecord(hp.up[c], hp.h2d);
        if (e != cudaSuccess) return set_cuda_error(e, "H2D activations");
    }
    for (int c = 0; c < n_slabs; ++c) {
        const int64_t r0 = c * rows, nr = (r0 + rows <= M) ? rows : M - r0;
        e = cudaStreamWaitEvent(s, hp.up[c], 0);
        if (e != cudaSuccess) return set_cuda_error(e, "linears_host: wait H2D");
        if (gated) {
            mixq_tensors g = *t[0], u = *t[1];
            g.A = u.A = dA + r0 * K * 2;
            g.Out = dOut[0] + r0 * N[0] * 2;
            const int rc = mixq_enqueue_gated(&g, &u, nr, N[0], K, ws, ws_bytes, nullptr, flags, stream);
            if (rc) return rc;
        }
        for (int i = 0; i < count && !gated; ++i) {
            mixq_tensors d = *t[i];
            d.A = dA + r0 * K * 2;
            d.Out = dOut[i] + r0 * N[i] * 2;
            const int rc = mixq_enqueue(&d, nr, N[i], K, ws, ws_bytes, flags, stream);
            if (rc) return rc;
        }
        e = cudaEventRecord(hp.done[c], s);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(hp.d2h, hp.done[c], 0);
        for (int i = 0; i < count && e == cudaSuccess; ++i)
            e = cudaMemcpyAsync(static_cast<uint8_t*>(Out_host[i]) + r0 * N[i] * 2, dOut[i] + r0 * N[i] * 2,
                                static_cast<size_t>(nr) * N[i] * 2, cudaMemcpyDeviceToHost, hp.d2h);
        if (e != cudaSuccess) return set_cuda_error(e, "D2H output");
    }
    if (async) {
        e = cudaEventRecord(rec->read, s);
        if (e == cudaSuccess) e = cudaEventRecord(rec->drained, hp.d2h);
        if (e != cudaSuccess) return set_cuda_error(e, "linears_host: record");
        rec->begin = range_begin;
        rec->end = range_end;
        rec->used = true;
        hp.pending = true;
        ++hp.n_async;
        return MIXQ_OK;
    }
    e = cudaEventRecord(hp.drained, hp.d2h);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s, hp.drained, 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return set_cuda_error(e, "stream synchronize");
    // the download stream is in order: asynchronous calls queued before this one have finished too
    hp.pending = false;
    for (auto& f : hp.inflight) f.used = false;
    hp.n_async = 0;
    return MIXQ_OK;
}
}  // namespace

extern "C" int64_t mixq_debug_host_part_offset(size_t need, size_t scratch_bytes, unsigned call_index) {
    if (need == 0 || scratch_bytes < need) return -1;
    return static_cast<int64_t>(host_part_offset(need, scratch_bytes, call_index));
}

extern "C" int mixq_host_drain(void* stream) {
    HostPipe& hp = host_pipe();
    if (!hp.ok) return set_error(MIXQ_ERR_CUDA, "host_drain: could not create the copy streams");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaSuccess;
    if (hp.pending) {
        e = cudaEventRecord(hp.drained, hp.d2h);   // the download stream is in order: everything issued so far
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s, hp.drained, 0);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    hp.pending = false;
    for (auto& f : hp.inflight) f.used = false;
    hp.n_async = 0;
    if (e != cudaSuccess) return set_cuda_error(e, "host_drain");
    return MIXQ_OK;
}

extern "C" int mixq_linears_host(const mixq_tensors* const* t, int count, const void* A_host, void* const* Out_host, int64_t M,
                                 const int64_t* N, int64_t K, void* dev_scratch, size_t dev_scratch_bytes, unsigned flags,
                                 void* stream) {
    return linears_host_impl(t, count, A_host, Out_host, M, N, K, dev_scratch, dev_scratch_bytes, flags, stream, false);
}

extern "C" size_t mixq_gated_host_scratch_size(int64_t M, int64_t N, int64_t K) {
    if (M <= 0 || K <= 0 || N <= 0) return 0;
    return align_up(static_cast<size_t>(M) * K * 2) + align_up(static_cast<size_t>(M) * N * 2) + mixq_gated_workspace_size(M, N, K) + kAlign;
}

extern "C" int mixq_gated_host(const mixq_tensors* gate, const mixq_tensors* up, const void* A_host, void* Out_host, int64_t M,
                               int64_t N, int64_t K, void* dev_scratch, size_t dev_scratch_bytes, unsigned flags, void* stream) {
    const mixq_tensors* t[2] = {gate, up};
    if (!gate || !up) return set_error(MIXQ_ERR_BAD_ARG, "gated_host: null tensor table");
    return linears_host_impl(t, 1, A_host, &Out_host, M, &N, K, dev_scratch, dev_scratch_bytes, flags, stream, true);
}

extern "C" int mixq_linear_host(const mixq_tensors* t, const void* A_host, void* Out_host, int64_t M, int64_t N, int64_t K,
                                void* dev_scratch, size_t dev_scratch_bytes, unsigned flags, void* stream) {
    if (!t || !A_host || !Out_host || !dev_scratch) return set_error(MIXQ_ERR_BAD_ARG, "linear_host: null pointer");
    if (M <= 0 || N <= 0 || K <= 0) return set_error(MIXQ_ERR_BAD_ARG, "linear_host: bad dimensions");
    return mixq_linears_host(&t, 1, A_host, &Out_host, M, &N, K, dev_scratch, dev_scratch_bytes, flags, stream);
}
