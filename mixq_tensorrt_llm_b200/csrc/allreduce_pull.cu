// allreduce_pull.cu -- the exchange step of a row-parallel linear for DECODE-SIZED results on a FEW ranks (SURVEY.md 8e;
// replaces "plugin, then allreduce(x, tp_group)", reference plugin.py:152-156).
//
// A few MB of fp16 partial sums are latency-bound: the one-kernel path (gemm_i8_tcgen05.cu: partial tiles pushed to their owner,
// sums broadcast back) pays two NVLink round trips behind two system-scope fences, each as long as the remote stores need to
// land (~20-25 us in all), and has to run the 256x256 bulk tile configuration, 6 us slower at this size than what `auto` picks.
// Here every byte crosses NVLink through peer LOADS instead -- nothing to fence, only flags to wait for:
//   GEMM     any tile configuration, partial [M, N] fp16 into this rank's staging area (local stores)
//   arrive   one remote atomic per peer ("my partial is complete"), then wait for theirs
//   one-shot (few ranks, ingress (world-1)*M*N*2 small): every rank pulls ALL partials, sums them in fp32 in RANK ORDER (the
//            arithmetic of the one-kernel path: bit-identical on every rank), rounds once, stores its own Out
//   two-shot (more ranks): rank r reduces only the r-th 1/world of the vectors the same way into its own Out, announces it,
//            and then copies the other ranks' reduced slices from THEIR Out buffers: 2 (world-1)/world * M*N*2 bytes come in
//   leave    "I have read everything" to every peer; the rank's last CTA waits for the peers' so that nobody's next GEMM (or next
//            exchange) can overwrite data that is still being read
// Counters only grow (targets scale with an epoch word), nothing is re-armed.
#include <cstdlib>

#include "mixq_internal.h"
#include "ptx.cuh"

namespace mixq {
namespace {

constexpr int kPullThreads = 512;
constexpr int kPullUnroll = 2;
// words of the rank's counter block used here (0..4 belong to the bulk kernel)
constexpr int kWordArrive = 8, kWordReadDone = 9, kWordEpoch = 10, kWordTicket = 11, kWordSliceReady = 12, kWordTicket2 = 13, kWordEpoch2 = 14, kWordError = 4;

struct PullParams {
    const uint4* partial[MIXQ_MAX_RANKS];   // every rank's partial result (peer-mapped), index = rank
    uint32_t* counters[MIXQ_MAX_RANKS];     // every rank's counter block
    uint4* out[MIXQ_MAX_RANKS];             // every rank's Out (two-shot: reduced slices are copied from the peers' Out)
    int world, rank, two_shot;
    unsigned long long timeout_ns;
};

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_sys_add(uint32_t* p, uint32_t v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_sys_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// poll with back-off; after timeout_ns (0 = never) raise the error word and give up (see wait_counter_sys in gemm_i8_tcgen05.cu)
__device__ __forceinline__ bool wait_at_least(const uint32_t* p, uint32_t expect, unsigned long long timeout_ns, uint32_t* err_word) {
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (static_cast<int32_t>(ld_acquire_sys_u32(p) - expect) < 0) {      // wrap-safe: the counters only grow
        if (++spins < 2048u) continue;
        __nanosleep(spins < (1u << 16) ? 64 : 1000);
        if ((spins & 0x3FFu) == 0u && timeout_ns != 0ull) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0ull) t0 = now;
            else if (now - t0 > timeout_ns) {
                atomicExch(err_word, 1u);
                return false;
            }
        }
    }
    return true;
}

__global__ void __launch_bounds__(kPullThreads, 1)
mixq_allreduce_pull_kernel(const __grid_constant__ PullParams p, uint4* __restrict__ out, size_t n_vec) {
    ptx::pdl_wait_prior_grid();          // this rank's partial is the previous kernel's output
    ptx::pdl_launch_dependents();
    uint32_t* cnt = p.counters[p.rank];
    __shared__ uint32_t s_epoch, s_epoch2;
    if (threadIdx.x == 0) {
        const uint32_t epoch = cnt[kWordEpoch] + 1u;     // calls so far + 1: the same number on every rank
        s_epoch = epoch;
        s_epoch2 = cnt[kWordEpoch2] + 1u;                // two-shot calls so far + 1 (slice-ready is only bumped by those)
        if (blockIdx.x == 0)                             // release: the partial (written by the prior grid) is visible system-wide
            for (int r = 1; r < p.world; ++r) red_release_sys_add(p.counters[(p.rank + r) % p.world] + kWordArrive, 1u);
        wait_at_least(cnt + kWordArrive, epoch * static_cast<uint32_t>(p.world - 1), p.timeout_ns, cnt + kWordError);
    }
    __syncthreads();
    const uint32_t epoch = s_epoch, epoch2 = s_epoch2;

    const size_t stride = static_cast<size_t>(gridDim.x) * kPullThreads;
    // the vectors this rank reduces: all of them (one-shot) or its 1/world slice (two-shot)
    const size_t per = p.two_shot ? (n_vec + p.world - 1) / p.world : n_vec;
    const size_t lo = p.two_shot ? per * p.rank : 0, hi = lo + per < n_vec ? lo + per : n_vec;
    for (size_t base = lo + static_cast<size_t>(blockIdx.x) * kPullThreads + threadIdx.x; base < hi; base += stride * kPullUnroll) {
        uint4 v[kPullUnroll][MIXQ_MAX_RANKS];
#pragma unroll
        for (int u = 0; u < kPullUnroll; ++u)
#pragma unroll
            for (int r = 0; r < MIXQ_MAX_RANKS; ++r)
                if (r < p.world && base + u * stride < hi) v[u][r] = ld_sys_v4(p.partial[r] + base + u * stride);
#pragma unroll
        for (int u = 0; u < kPullUnroll; ++u) {
            if (base + u * stride >= hi) continue;
            float acc[8];
#pragma unroll
            for (int r = 0; r < MIXQ_MAX_RANKS; ++r) {
                if (r < p.world) {
                    const uint32_t w[4] = {v[u][r].x, v[u][r].y, v[u][r].z, v[u][r].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                        acc[2 * e] = r == 0 ? f.x : acc[2 * e] + f.x;            // rank order: deterministic, identical on every rank
                        acc[2 * e + 1] = r == 0 ? f.y : acc[2 * e + 1] + f.y;
                    }
                }
            }
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const __half2 h = __floats2half2_rn(acc[2 * e], acc[2 * e + 1]);
                pk[e] = *reinterpret_cast<const uint32_t*>(&h);
            }
            out[base + u * stride] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
    if (p.two_shot) {
        // ---- second shot: this rank's slice is final once all its CTAs are here; tell the peers, wait for theirs, copy their slices
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(cnt + kWordTicket2, 1u) == gridDim.x - 1u) {
                cnt[kWordTicket2] = 0u;
                for (int r = 0; r < p.world; ++r) red_release_sys_add(p.counters[(p.rank + r) % p.world] + kWordSliceReady, 1u);   // r = 0: ourselves
            }
            wait_at_least(cnt + kWordSliceReady, epoch2 * static_cast<uint32_t>(p.world), p.timeout_ns, cnt + kWordError);
        }
        __syncthreads();
        for (int r = 1; r < p.world; ++r) {
            const int src = (p.rank + r) % p.world;
            const size_t slo = per * src, shi = slo + per < n_vec ? slo + per : n_vec;
            const uint4* from = p.out[src];
            for (size_t base = slo + static_cast<size_t>(blockIdx.x) * kPullThreads + threadIdx.x; base < shi; base += stride * 4) {
                uint4 c[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (base + u * stride < shi) c[u] = ld_sys_v4(from + base + u * stride);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (base + u * stride < shi) out[base + u * stride] = c[u];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        // the last CTA of this rank tells every peer that their partials have been read, and waits for the same from them: after
        // that nobody's next GEMM can overwrite a partial that is still being read
        if (atomicAdd(cnt + kWordTicket, 1u) == gridDim.x - 1u) {
            cnt[kWordTicket] = 0u;
            for (int r = 1; r < p.world; ++r) red_release_sys_add(p.counters[(p.rank + r) % p.world] + kWordReadDone, 1u);
            wait_at_least(cnt + kWordReadDone, epoch * static_cast<uint32_t>(p.world - 1), p.timeout_ns, cnt + kWordError);
            cnt[kWordEpoch] = epoch;
            if (p.two_shot) cnt[kWordEpoch2] = epoch2;
        }
    }
}

}  // namespace

// Out[rank] = fp16(sum over ranks, in rank order, of fp32(partial_r)), n_elems fp16 values (% 8 == 0); `partials` / `counters` / `outs`
// hold peer-mapped addresses of every rank's buffers.  Every rank must call it with the same n_elems in the same order.
int launch_allreduce_pull(void* const* partials, void* const* counters, void* const* outs, int world, int rank, size_t n_elems,
                          bool two_shot, cudaStream_t stream, bool pdl, LaunchOpts opts) {
    if (!partials || !counters || !outs) return set_error(MIXQ_ERR_BAD_ARG, "allreduce: null pointer");
    if (world < 1 || world > MIXQ_MAX_RANKS || rank < 0 || rank >= world) return set_error(MIXQ_ERR_BAD_ARG, "allreduce: bad world/rank");
    if (n_elems == 0) return MIXQ_OK;
    if (n_elems & 7) return set_error(MIXQ_ERR_BAD_ARG, "allreduce: element count must be a multiple of 8");
    if (!device_info().ok) return set_error(MIXQ_ERR_CUDA, "no usable sm_100 device");
    static const long long timeout_ms = [] {
        const char* e = std::getenv("MIXQ_AR_TIMEOUT_MS");
        return e ? std::atoll(e) : 60000ll;
    }();
    PullParams p{};
    p.world = world;
    p.rank = rank;
    p.two_shot = two_shot ? 1 : 0;
    p.timeout_ns = timeout_ms > 0 ? static_cast<unsigned long long>(timeout_ms) * 1000000ull : 0ull;
    for (int i = 0; i < world; ++i) {
        if (!partials[i] || !counters[i] || !outs[i] || ((reinterpret_cast<uintptr_t>(partials[i]) | reinterpret_cast<uintptr_t>(outs[i])) & 15))
            return set_error(MIXQ_ERR_BAD_ARG, "allreduce: null or misaligned peer pointer");
        p.partial[i] = static_cast<const uint4*>(partials[i]);
        p.counters[i] = static_cast<uint32_t*>(counters[i]);
        p.out[i] = static_cast<uint4*>(outs[i]);
    }
    const size_t n_vec = n_elems / 8;
    const size_t work = two_shot ? (n_vec + world - 1) / world : n_vec;
    int grid = static_cast<int>((work + static_cast<size_t>(kPullThreads) * kPullUnroll - 1) / (static_cast<size_t>(kPullThreads) * kPullUnroll));
    const int sms = usable_sms(opts);
    if (grid > sms) grid = sms;
    if (grid < 1) grid = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kPullThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, mixq_allreduce_pull_kernel, p, p.out[rank], n_vec);
    if (e != cudaSuccess) return set_cuda_error(e, "launch allreduce_pull");
    count_launch();
    return MIXQ_OK;
}

}  // namespace mixq
