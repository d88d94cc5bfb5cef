// gemm_i8_tcgen05.cu -- stage 2 of the W8A8O16 path: ONE kernel that
//   (a) contracts A8[M,K] (int8) with W8[N,K]^T (int8) on the tcgen05 INT8 tensor cores
//       (tcgen05.mma.kind::i8, 128 x BLOCK_N x 32 per instruction, int32 accumulators in TMEM),
//   (b) contracts the 128 outlier columns fp_A[M,128] x fp_weight[N,128]^T with
//       tcgen05.mma.kind::f16 into a second (fp32) TMEM accumulator, and
//   (c) dequantises in the epilogue:
//          Out[m,n] = fp16( fma( float(acc_i32), float(sb[n]) * float(sa[m]), float(fp16(acc_f32)) ) )
//
// It replaces two reference launches and the fp16 round trip between them:
//   gemmfp16 / cublasGemmEx            (TsinghuaMixQPlugin.cpp:122-161, called at :521)
//   int8FusedDequantizeCUDA / CUTLASS  (kernel/i8gemm.cu:151-194; mainloop
//       kernel/symmetric/gemm/kernel/gemm_dequant.h:224-292, epilogue
//       kernel/symmetric/epilogue/thread/linear_combination_dequant.h:152-157)
// The inner fp16() in (c) reproduces the reference's rounding of the outlier product when
// cuBLAS stores it to `Out` before the CUTLASS epilogue reads it back as the addend.
//
// Structure (persistent, warp specialised, 256 threads, 1 CTA / SM):
//   warp 0   TMA producer: fills a ring of kStages smem slots, each one K-block of
//            A (128 rows x 128 B) and of W (BLOCK_N rows x 128 B), 128B-swizzled.  The outlier
//            slab rides the same ring as two extra K-blocks of 64 fp16 columns (same footprint).
//   warp 1   MMA issuer: one elected thread issues 4 tcgen05.mma per slot and commits the slot's
//            "empty" barrier; after the last K-block it commits the accumulator's "full" barrier.
//   warp 2   TMEM allocator (alloc before, dealloc after).
//   warps 4-7 epilogue: tcgen05.ld both accumulators (32 lanes x 32 columns per warp and step),
//            dequantise in registers, 16-byte global stores.  With kAccStages = 2 the epilogue of
//            tile i overlaps the main loop of tile i+1.
// Barriers: full[kStages] (TMA -> MMA, tx bytes), empty[kStages] (MMA -> TMA, tcgen05.commit),
//           tmem_full[kAccStages] (MMA -> epilogue), tmem_empty[kAccStages] (epilogue -> MMA).
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "mixq_internal.h"
#include "ptx.cuh"

namespace mixq {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockKBytes = 128;  // one 128B swizzle atom per row per K-block
constexpr int kUmmaKBytes = 32;    // K=32 int8 / K=16 fp16 per tcgen05.mma
constexpr int kGemmThreads = 256;
constexpr int kEpilogueWarp0 = 4;
constexpr int kNumEpilogueThreads = 128;
constexpr int kOutlierKBlocks = (MIXQ_NUM_OUTLIERS * 2) / kBlockKBytes;  // 2
constexpr size_t kStreamKFlagBytes = 4096;                                 // stream-K: one flag per (worker, CTA rank)
constexpr size_t kStreamKSlotBytes = 2 * (128 * 256 * 4 + 128 * 256 * 2);  // int32 partial tile + fp16 outlier product of one CTA pair
constexpr int kStreamKMaxWorkers = 80;                                     // CTA pairs (148 SMs -> 74)
constexpr int kStashEpiThreads = 256;                                      // wide-tile kernel: 8 epilogue warps
constexpr int kStashThreads = kEpilogueWarp0 * 32 + kStashEpiThreads;      // 384

// CTA  = 1: one CTA per 128 x BLOCK_N tile (tcgen05 cta_group::1, UMMA M = 128)
// CTA  = 2: a CTA pair (cluster of 2 on one TPC) per 256 x BLOCK_N tile (cta_group::2, UMMA M = 256):
//           each CTA stages its own 128 rows of A and HALF of the W rows, the leader issues the MMAs
//           for both, each CTA's TMEM holds its 128 accumulator rows.  Halves the shared-memory
//           and L2 operand traffic per MAC.
// A_ROWS: rows of the A box staged per K-block: 128, or 32 / 64 for a batch that small.  The tensor core still reads 128 rows
//           (whatever follows the A box in shared memory) into accumulator rows nobody stores; the smaller stage buys a deeper
//           ring, and a decode-sized call is bound by the weight bytes a CTA has in flight (HBM latency x bandwidth).
template <int CTA, int BLOCK_N, int ACC_STAGES, int STAGES, int A_ROWS = 128>
struct GemmTraits {
    static constexpr int kCta = CTA;
    static constexpr int kARows = A_ROWS;
    static constexpr int kBlockN = BLOCK_N;            // UMMA N = output columns per tile
    static constexpr int kLoadN = BLOCK_N / CTA;       // W rows each CTA stages per K-block
    static constexpr int kTileM = kBlockM * CTA;       // output rows per tile
    static constexpr int kAccStages = ACC_STAGES;
    static constexpr int kStages = STAGES;
    static constexpr int kABytes = A_ROWS * kBlockKBytes;
    static constexpr int kBBytes = kLoadN * kBlockKBytes;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static_assert(A_ROWS == 128 || (CTA == 1 && (A_ROWS == 32 || A_ROWS == 64)), "short A boxes: one-CTA tiles only");
    static_assert(kABytes % 1024 == 0, "A box must be whole 8-row swizzle groups");
    static constexpr int kAccCols = 2 * BLOCK_N;  // int32 accumulator | fp32 outlier accumulator
    static constexpr int kTmemColsRaw = kAccCols * ACC_STAGES;
    static constexpr int kTmemCols = kTmemColsRaw <= 32 ? 32 : kTmemColsRaw <= 64 ? 64 : kTmemColsRaw <= 128 ? 128
                                     : kTmemColsRaw <= 256 ? 256 : 512;
    static_assert(CTA == 1 || CTA == 2, "CTA group size");
    static_assert(kTmemColsRaw <= 512, "TMEM has 512 columns");
    static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "BLOCK_N");
    static_assert(kLoadN % 8 == 0 && kBBytes % 1024 == 0, "W tile must be whole 8-row swizzle groups");
    // dynamic smem: ring | sb and bias staging (2 x 2 x BLOCK_N floats) | barriers | tmem ptr  (+1024 alignment slack)
    static constexpr int kMaxKSplit = 8;                                       // CTAs of a cluster that share one tile along K (CTA == 1)
    static constexpr int kNumBarriers = 2 * STAGES + 2 * ACC_STAGES + 1 + (kMaxKSplit - 1) * 4;   // + ring_free, partial sums landed
    // (+ the bytes the tensor core over-reads behind the last stage's short A box)
    static constexpr size_t kSmemBytes =
        1024 + static_cast<size_t>(STAGES) * kStageBytes + 4 * BLOCK_N * sizeof(float) + kNumBarriers * 8 + 16 +
        (A_ROWS < 128 ? (128 - A_ROWS) * kBlockKBytes : 0);
    static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

struct TileCoord {
    int m_blk, n_blk;
};
// Grouped rasterisation: tiles are walked in bands of `group_m` row-blocks, M fastest inside a
// band, so the CTAs of one wave share a few A row-blocks and a few W row-blocks in L2.  Odd bands walk the
// column tiles backwards: the weights a band touched last are the ones the next band starts with, so the part of W
// that is still in L2 is re-used instead of being re-streamed from HBM (W alone does not fit beside a band's A and
// results when N x K is tens of MB).
__device__ __forceinline__ TileCoord tile_coord(int tile, int m_tiles, int n_tiles, int group_m) {
    const int per_group = group_m * n_tiles;
    const int g = tile / per_group;
    const int first_m = g * group_m;
    const int gsz = min(m_tiles - first_m, group_m);
    const int r = tile - g * per_group;
    const int n = r / gsz;
    return {first_m + r % gsz, (g & 1) ? n_tiles - 1 - n : n};
}

// Optional fused epilogue: activation in fp32 BEFORE the output rounding, as the reference's
// LinearCombinationDequantSilu does (kernel/symmetric/epilogue/thread/linear_combination_dequant.h:167-272,
// silu(x) = x / (1 + expf(-x)); the reference extension is built with --use_fast_math, hence the fast intrinsics),
// then the bias added to the fp16 result with a second rounding, as the reference adds it outside the kernel
// (plugin.py:158-160, MixQ/src/mixquant/modules/linear.py:368-369).
struct EpiArgs {
    const __half* bias;   // fp16 [N] or null
    int act;              // MIXQ_ACT_NONE / MIXQ_ACT_SILU
};
__device__ __forceinline__ __half2 epi_finish(float r0, float r1, const float* bias_s, const EpiArgs& e) {
    if (e.act == MIXQ_ACT_NONE && e.bias == nullptr) return __floats2half2_rn(r0, r1);   // the plugin's epilogue (uniform branch)
    if (e.act == MIXQ_ACT_SILU) {
        r0 = __fdividef(r0, 1.0f + __expf(-r0));
        r1 = __fdividef(r1, 1.0f + __expf(-r1));
    }
    __half2 h = __floats2half2_rn(r0, r1);
    if (e.bias) {
        const float2 f = __half22float2(h);
        h = __floats2half2_rn(f.x + bias_s[0], f.y + bias_s[1]);
    }
    return h;
}

template <class T>
__global__ void __launch_bounds__(kGemmThreads, 1)
mixq_gemm_dequant_kernel(const __grid_constant__ CUtensorMap tm_a8, const __grid_constant__ CUtensorMap tm_w8,
                         const __grid_constant__ CUtensorMap tm_fa, const __grid_constant__ CUtensorMap tm_fw,
                         const __half* __restrict__ scale_a, const __half* __restrict__ scale_b,
                         __half* __restrict__ Out, int M, int N, int K, int has_outlier, int m_tiles, int n_tiles,
                         int group_m, int ksplit, EpiArgs epi) {
    // ksplit > 1 (one-CTA tiles, one row-block of tokens): a cluster of `ksplit` CTAs shares ONE tile, each reducing a slice of K.
    // A decode-sized call with few tokens is bound by the number of tcgen05.mma it issues along K (~93 clk each below N = 128,
    // whatever the tile), so the K loop is what has to be parallelised.  Rank 0 also does the outlier K-blocks and finishes the
    // tile; the others ship the int32 partial sums of the rows that exist (32 per epilogue warp) into rank 0's idle ring with
    // st.async, counted on mbarriers there.  Integer sums: the result is bit-identical to the unsplit kernel.
    constexpr int BLOCK_N = T::kBlockN;
    constexpr int CTA = T::kCta;
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled tiles need 1024-byte alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    float* sb_s = reinterpret_cast<float*>(ring + static_cast<size_t>(T::kStages) * T::kStageBytes);
    float* bias_sm = sb_s + 2 * BLOCK_N;   // staged bias, double-buffered like sb_s
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sb_s + 4 * BLOCK_N);
    uint64_t* empty_bar = full_bar + T::kStages;
    uint64_t* tmem_full_bar = empty_bar + T::kStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + T::kAccStages;
    uint64_t* ring_free_bar = tmem_empty_bar + T::kAccStages;   // K-split, in a sending CTA: rank 0's ring may be overwritten
    uint64_t* part_bar = ring_free_bar + 1;                     // K-split, in rank 0: [sender 1..7][epilogue warp] partial sums landed
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(part_bar + (T::kMaxKSplit - 1) * 4);

    const int warp_idx = threadIdx.x >> 5;  // warp-uniform
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CTA == 2) ? ptx::cluster_ctarank() : 0u;
    const bool is_leader = cta_rank == 0;
    const int ks = (CTA == 1 && ksplit > 1) ? ksplit : 1;
    const uint32_t krank = ks > 1 ? ptx::cluster_ctarank() : 0u;   // which K slice this CTA reduces
    if (krank) has_outlier = 0;
    const int row_warps = min(4, (M + 31) >> 5);                  // epilogue warps that hold real token rows (K-split: one row-block)
    // tiles are distributed over CTA groups (a single CTA, a pair working on one 256-row tile, or a K-split cluster)
    const int group_id = blockIdx.x / (CTA * ks);
    const int num_groups = gridDim.x / (CTA * ks);

    if (warp_idx == 0 && ptx::elect_one()) {
        ptx::prefetch_tensormap(&tm_a8);
        ptx::prefetch_tensormap(&tm_w8);
        if (has_outlier) {
            ptx::prefetch_tensormap(&tm_fa);
            ptx::prefetch_tensormap(&tm_fw);
        }
    }
    if (warp_idx == 1 && ptx::elect_one()) {
        for (int i = 0; i < T::kStages; ++i) {
            ptx::mbar_init(&full_bar[i], 1);   // leader's arrive.expect_tx; TMA bytes of the whole group
            ptx::mbar_init(&empty_bar[i], 1);  // one tcgen05.commit (multicast to both CTAs of a pair)
        }
        for (int i = 0; i < T::kAccStages; ++i) {
            ptx::mbar_init(&tmem_full_bar[i], 1);
            ptx::mbar_init(&tmem_empty_bar[i], CTA * kNumEpilogueThreads / 32);  // every epilogue warp of the group
        }
        ptx::mbar_init(ring_free_bar, 1);
        for (int i = 0; i < (T::kMaxKSplit - 1) * 4; ++i) ptx::mbar_init(&part_bar[i], 1);
        if (ks > 1 && krank == 0)   // arm the landing barriers: sender sd, warp q brings 32 lanes x BLOCK_N columns x 4 bytes
            for (int sd = 0; sd < ks - 1; ++sd)
                for (int q = 0; q < row_warps; ++q) ptx::mbar_arrive_expect_tx(&part_bar[sd * 4 + q], 32u * BLOCK_N * 4u);
        ptx::fence_barrier_init();
    }
    if (warp_idx == 2) {
        if constexpr (CTA == 2) {
            ptx::tmem_alloc_2cta(tmem_ptr_s, T::kTmemCols);
            ptx::tmem_relinquish_2cta();
        } else {
            ptx::tmem_alloc(tmem_ptr_s, T::kTmemCols);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before_sync();
    if (CTA == 2 || ks > 1) ptx::cluster_sync(); else __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr_s;

    // Everything above overlaps the tail of the previous kernel (PDL); A8/sa/fp_A are produced by it.
    ptx::pdl_wait_prior_grid();
    ptx::pdl_launch_dependents();   // dependents may be scheduled as our CTAs retire; they wait for this grid's completion themselves

    const int num_tiles = m_tiles * n_tiles;
    const int num_kb_all = (K + kBlockKBytes - 1) / kBlockKBytes;
    // K-split: rank 0 (outlier K-blocks, finishing epilogue) gets about two K-blocks less than the others
    const int kb_cut0 = ks > 1 ? max(1, min(num_kb_all - (ks - 1), num_kb_all / ks - 2)) : num_kb_all;
    const int kb_first = krank == 0 ? 0 : kb_cut0 + static_cast<int>((static_cast<long long>(krank - 1) * (num_kb_all - kb_cut0)) / (ks - 1));
    const int kb_last = krank == 0 ? kb_cut0
                                   : kb_cut0 + static_cast<int>((static_cast<long long>(krank) * (num_kb_all - kb_cut0)) / (ks - 1));
    const int num_kb = kb_last - kb_first;      // K-blocks of THIS CTA
    const int n_f = has_outlier ? kOutlierKBlocks : 0;
    const int num_items = n_f + num_kb;

    if (warp_idx == 0) {
        if (ptx::elect_one()) {
            // ===================== TMA producer (every CTA) =====================
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = group_id; tile < num_tiles; tile += num_groups) {
                const TileCoord tc = tile_coord(tile, m_tiles, n_tiles, group_m);
                const int m0 = tc.m_blk * T::kTileM + static_cast<int>(cta_rank) * kBlockM;
                const int n0 = tc.n_blk * BLOCK_N + static_cast<int>(cta_rank) * T::kLoadN;
                for (int it = 0; it < num_items; ++it) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (is_leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], T::kStageBytes * CTA);
                    uint8_t* sA = ring + static_cast<size_t>(stage) * T::kStageBytes;
                    uint8_t* sB = sA + T::kABytes;
                    const CUtensorMap* ma = it < n_f ? &tm_fa : &tm_a8;
                    const CUtensorMap* mb = it < n_f ? &tm_fw : &tm_w8;
                    const int k0 = it < n_f ? it * (kBlockKBytes / 2) : (kb_first + it - n_f) * kBlockKBytes;
                    if constexpr (CTA == 2) {
                        ptx::tma_load_2d_2cta(sA, ma, &full_bar[stage], k0, m0, ptx::kEvictNormal);
                        ptx::tma_load_2d_2cta(sB, mb, &full_bar[stage], k0, n0, ptx::kEvictNormal);
                    } else {
                        ptx::tma_load_2d(sA, ma, &full_bar[stage], k0, m0, ptx::kEvictNormal);
                        ptx::tma_load_2d(sB, mb, &full_bar[stage], k0, n0, ptx::kEvictNormal);
                    }
                    if (++stage == T::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp_idx == 1) {
        if (is_leader && ptx::elect_one()) {
            // ===================== MMA issuer (leader CTA only) =====================
            // The loop body runs once per 128-byte K-block and must stay well below the time the tensor
            // core needs for it (4 MMAs = 256 cycles at N = 128), so everything loop-invariant is
            // hoisted: descriptors are a per-stage constant plus 2 (= 32 B >> 4) per MMA, the outlier
            // and int8 K-blocks run in separate loops, and the barrier of the NEXT stage is probed
            // before the MMAs of this one are issued so its latency overlaps the issue.
            constexpr uint32_t idesc_i8 = ptx::make_idesc_i8(T::kTileM, BLOCK_N);
            constexpr uint32_t idesc_f16 = ptx::make_idesc_f16(T::kTileM, BLOCK_N);
            constexpr uint32_t kDescStep = kUmmaKBytes >> 4;
            const uint64_t desc_a0 = ptx::make_smem_desc_sw128(ptx::smem_u32(ring));
            const uint64_t desc_b0 = ptx::make_smem_desc_sw128(ptx::smem_u32(ring) + T::kABytes);
            int stage = 0;
            uint32_t phase = 0;
            int acc_stage = 0;
            uint32_t acc_phase = 0;
            bool ready = false;
            auto issue_block = [&](auto kind_tag, uint32_t tmem_d, bool first) {
                if (!ready) ptx::mbar_wait(&full_bar[stage], phase);
                const int nstage = (stage + 1 == T::kStages) ? 0 : stage + 1;
                const uint32_t nphase = (stage + 1 == T::kStages) ? phase ^ 1 : phase;
                ready = ptx::mbar_try_wait(&full_bar[nstage], nphase);  // early probe of the next slot
                ptx::tc_fence_after_sync();
                const uint64_t da = desc_a0 + static_cast<uint64_t>(stage) * (T::kStageBytes >> 4);
                const uint64_t db = desc_b0 + static_cast<uint64_t>(stage) * (T::kStageBytes >> 4);
#pragma unroll
                for (int k = 0; k < kBlockKBytes / kUmmaKBytes; ++k) {
                    const uint32_t acc = (first && k == 0) ? 0u : 1u;
                    if constexpr (decltype(kind_tag)::value == 0) {
                        if constexpr (CTA == 2) ptx::umma_f16_2cta(tmem_d, da + k * kDescStep, db + k * kDescStep, idesc_f16, acc);
                        else ptx::umma_f16(tmem_d, da + k * kDescStep, db + k * kDescStep, idesc_f16, acc);
                    } else {
                        if constexpr (CTA == 2) ptx::umma_i8_2cta(tmem_d, da + k * kDescStep, db + k * kDescStep, idesc_i8, acc);
                        else ptx::umma_i8(tmem_d, da + k * kDescStep, db + k * kDescStep, idesc_i8, acc);
                    }
                }
                // slot reusable (in both CTAs of a pair) once these MMAs have read it
                if constexpr (CTA == 2) ptx::umma_commit_2cta(&empty_bar[stage]); else ptx::umma_commit(&empty_bar[stage]);
                stage = nstage;
                phase = nphase;
            };
            for (int tile = group_id; tile < num_tiles; tile += num_groups) {
                ptx::mbar_wait(&tmem_empty_bar[acc_stage], acc_phase ^ 1);
                ptx::tc_fence_after_sync();
                const uint32_t tmem_i = tmem_base + acc_stage * T::kAccCols;
                const uint32_t tmem_f = tmem_i + BLOCK_N;
                for (int it = 0; it < n_f; ++it) issue_block(std::integral_constant<int, 0>{}, tmem_f, it == 0);
                for (int kb = 0; kb < num_kb; ++kb) issue_block(std::integral_constant<int, 1>{}, tmem_i, kb == 0);
                // accumulators complete (signalled to the epilogue warps of both CTAs)
                if constexpr (CTA == 2) ptx::umma_commit_2cta(&tmem_full_bar[acc_stage]);
                else ptx::umma_commit(&tmem_full_bar[acc_stage]);
                if (++acc_stage == T::kAccStages) {
                    acc_stage = 0;
                    acc_phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp_idx >= kEpilogueWarp0) {
        // ===================== epilogue (every CTA: its own 128 accumulator rows) =====================
        const int quarter = warp_idx - kEpilogueWarp0;  // == warp_idx % 4: the TMEM lane quarter this warp may read
        const int et = threadIdx.x - kEpilogueWarp0 * 32;
        const int row = quarter * 32 + lane;
        int acc_stage = 0;
        uint32_t acc_phase = 0;
        int local_tile = 0;
        for (int tile = group_id; tile < num_tiles; tile += num_groups, ++local_tile) {
            const TileCoord tc = tile_coord(tile, m_tiles, n_tiles, group_m);
            const int m0 = tc.m_blk * T::kTileM + static_cast<int>(cta_rank) * kBlockM;
            const int n0 = tc.n_blk * BLOCK_N;
            float* sbt = sb_s + (local_tile & 1) * BLOCK_N;
            float* bt = bias_sm + (local_tile & 1) * BLOCK_N;
            for (int j = et; j < BLOCK_N; j += kNumEpilogueThreads) {
                sbt[j] = (n0 + j < N) ? __half2float(scale_b[n0 + j]) : 0.0f;
                bt[j] = (epi.bias && n0 + j < N) ? __half2float(epi.bias[n0 + j]) : 0.0f;
            }
            const int gm = m0 + row;
            const bool row_ok = gm < M;
            const float sa_f = row_ok ? __half2float(scale_a[gm]) : 0.0f;
            ptx::named_bar_sync(1, kNumEpilogueThreads);

            ptx::mbar_wait(&tmem_full_bar[acc_stage], acc_phase);
            ptx::tc_fence_after_sync();
            const uint32_t t_i = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc_stage * T::kAccCols;
            __half* out_row = Out + static_cast<size_t>(gm) * N + n0;
            // K-split landing zone in rank 0's ring: [sender][row warp][chunk of 32 columns][16-byte vector 0..7][lane] (4 KB a chunk)
            const uint32_t zone_q = ptx::smem_u32(ring) + static_cast<uint32_t>(quarter) * (BLOCK_N * 128u) + lane * 16u;
            if (ks > 1 && krank != 0) {
                // ---- sending CTA: ship the int32 partial sums of the rows that exist to rank 0
                ptx::mbar_wait(ring_free_bar, 0);                 // rank 0's tensor core has read its whole ring
                if (quarter < row_warps) {
                    const uint32_t r_zone = ptx::mapa_shared(zone_q + (krank - 1u) * static_cast<uint32_t>(row_warps) * (BLOCK_N * 128u), 0u);
                    const uint32_t r_bar = ptx::mapa_shared(ptx::smem_u32(&part_bar[(krank - 1u) * 4u + quarter]), 0u);
#pragma unroll 1
                    for (int c = 0; c < BLOCK_N / 32; ++c) {
                        uint32_t vi[32];
                        ptx::tmem_ld_32x32(t_i + c * 32, vi);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            ptx::st_async_v4(r_zone + c * 4096u + g * 512u, vi[4 * g], vi[4 * g + 1], vi[4 * g + 2], vi[4 * g + 3], r_bar);
                    }
                }
                ptx::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc_stage]);
                if (++acc_stage == T::kAccStages) {
                    acc_stage = 0;
                    acc_phase ^= 1;
                }
                continue;
            }
            if (ks > 1 && et == 0)
                for (int r = 1; r < ks; ++r) ptx::mbar_arrive_cluster(ring_free_bar, static_cast<uint32_t>(r));   // senders may overwrite our ring
            if (ks > 1 && quarter < row_warps)
                for (int sd = 0; sd < ks - 1; ++sd) ptx::mbar_wait(&part_bar[sd * 4 + quarter], 0);
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t vi[32], vf[32];
                ptx::tmem_ld_32x32(t_i + c * 32, vi);
                if (has_outlier) {
                    ptx::tmem_ld_32x32(t_i + BLOCK_N + c * 32, vf);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) vf[j] = 0u;
                }
                ptx::tmem_ld_wait();
                if (ks > 1 && quarter < row_warps) {
                    for (int sd = 0; sd < ks - 1; ++sd) {
                        const uint32_t src = zone_q + static_cast<uint32_t>(sd * row_warps) * (BLOCK_N * 128u) + c * 4096u;
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            const uint4 pv = ptx::ld_shared_u4(src + g * 512u);
                            vi[4 * g] += pv.x; vi[4 * g + 1] += pv.y; vi[4 * g + 2] += pv.z; vi[4 * g + 3] += pv.w;
                        }
                    }
                }
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float p0 = __fmul_rn(sbt[c * 32 + j], sa_f);
                    const float p1 = __fmul_rn(sbt[c * 32 + j + 1], sa_f);
                    // outlier product: fp32 accumulator -> fp16 (the reference's store to Out) -> fp32 addend
                    const __half2 o = __floats2half2_rn(__uint_as_float(vf[j]), __uint_as_float(vf[j + 1]));
                    const float2 of = __half22float2(o);
                    const float r0 = __fmaf_rn(__int2float_rn(static_cast<int>(vi[j])), p0, of.x);
                    const float r1 = __fmaf_rn(__int2float_rn(static_cast<int>(vi[j + 1])), p1, of.y);
                    const __half2 r = epi_finish(r0, r1, bt + c * 32 + j, epi);
                    packed[j >> 1] = *reinterpret_cast<const uint32_t*>(&r);
                }
                if (row_ok) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (n0 + c * 32 + g * 8 + 8 <= N)
                            ptx::st_global_v4(out_row + c * 32 + g * 8, packed[g * 4], packed[g * 4 + 1],
                                              packed[g * 4 + 2], packed[g * 4 + 3]);
                    }
                }
            }
            // hand the accumulator stage back to the MMA warp (which lives in the leader CTA)
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CTA == 2) ptx::mbar_arrive_cluster(&tmem_empty_bar[acc_stage], 0);
                else ptx::mbar_arrive(&tmem_empty_bar[acc_stage]);
            }
            if (++acc_stage == T::kAccStages) {
                acc_stage = 0;
                acc_phase ^= 1;
            }
        }
    }

    // teardown
    ptx::tc_fence_before_sync();
    if (CTA == 2 || ks > 1) ptx::cluster_sync(); else __syncthreads();
    if (warp_idx == 2) {
        if constexpr (CTA == 2) ptx::tmem_dealloc_2cta(tmem_base, T::kTmemCols);
        else ptx::tmem_dealloc(tmem_base, T::kTmemCols);
    }
}

// =================================================================================================
// Wide-tile variant ("stash" kernel): 256 output columns per tile with DOUBLE-BUFFERED int32
// accumulators (2 x 256 TMEM columns = all of TMEM), so the epilogue of tile t overlaps the main
// loop of tile t+1 while the operand traffic per MAC is the lowest the hardware allows
// (with CTA = 2: 64 B/clk/SM from L2 and from shared memory).
// There is no TMEM left for the fp32 outlier accumulator, so it borrows the *other* accumulator
// buffer while that one is idle:
//     MMA thread, tile t in buffer b:   I-MMAs [0, split) -> wait E(t-1) done -> F-MMAs into buffer b^1
//                                       -> commit f_full -> I-MMAs [split, end) -> commit tmem_full(b)
//     epilogue warps, tile t:           wait f_full -> drain F (fp32 -> fp16, the reference's rounding)
//                                       into a thread-private shared-memory stash -> arrive f_drained(b^1)
//                                       -> wait tmem_full(b) -> out = fma(float(I), sa*sb, stash) -> release b
// f_drained(b^1) gates the main loop of tile t+1 (which accumulates into b^1); both happen half a
// main loop earlier than needed, so in steady state the tensor core never waits.
template <int CTA, int STAGES, int BLOCK_N = 256>
struct StashTraits {
    static constexpr int kCta = CTA;
    static constexpr int kBlockN = BLOCK_N;   // 256, or 192 for decode-sized M (more, smaller tiles per wave)
    static_assert(BLOCK_N % 64 == 0 && BLOCK_N <= 256 && (BLOCK_N / CTA) % 8 == 0, "two column halves of whole 32-column chunks");
    static constexpr int kLoadN = kBlockN / CTA;
    static constexpr int kTileM = kBlockM * CTA;
    static constexpr int kStages = STAGES;
    static constexpr int kABytes = kBlockM * kBlockKBytes;
    static constexpr int kBBytes = kLoadN * kBlockKBytes;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = 512;
    static constexpr int kStashBytes = kBlockM * kBlockN * 2;  // fp16 outlier product of this CTA's 128 rows
    static constexpr int kNumBarriers = 2 * STAGES + 8;
    static constexpr size_t kSmemBytes = 1024 + static_cast<size_t>(STAGES) * kStageBytes + kStashBytes +
                                         4 * kBlockN * sizeof(float) + kNumBarriers * 8 + 16;
    static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

template <class T>
__global__ void __launch_bounds__(kStashThreads, 1)
mixq_gemm_dequant_stash_kernel(const __grid_constant__ CUtensorMap tm_a8, const __grid_constant__ CUtensorMap tm_w8,
                               const __grid_constant__ CUtensorMap tm_fa, const __grid_constant__ CUtensorMap tm_fw,
                               const __half* __restrict__ scale_a, const __half* __restrict__ scale_b,
                               __half* __restrict__ Out, int M, int N, int K, int has_outlier, int m_tiles,
                               int n_tiles, int group_m, int /*ksplit: plain one-CTA tiles only*/, EpiArgs epi) {
    constexpr int BLOCK_N = T::kBlockN;
    constexpr int CTA = T::kCta;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    uint4* stash = reinterpret_cast<uint4*>(ring + static_cast<size_t>(T::kStages) * T::kStageBytes);
    float* sb_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(stash) + T::kStashBytes);
    float* bias_sm = sb_s + 2 * BLOCK_N;   // staged bias, double-buffered like sb_s
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sb_s + 4 * BLOCK_N);
    uint64_t* empty_bar = full_bar + T::kStages;
    uint64_t* tmem_full_bar = empty_bar + T::kStages;  // [2] int32 accumulators of buffer b complete
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2] epilogue has read buffer b's int32 accumulators
    uint64_t* f_full_bar = tmem_empty_bar + 2;         // [2] outlier accumulators in buffer b complete
    uint64_t* f_drained_bar = f_full_bar + 2;          // [2] outlier accumulators of buffer b moved to the stash
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(f_drained_bar + 2);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CTA == 2) ? ptx::cluster_ctarank() : 0u;
    const bool is_leader = cta_rank == 0;
    const int group_id = blockIdx.x / CTA;
    const int num_groups = gridDim.x / CTA;

    if (warp_idx == 0 && ptx::elect_one()) {
        ptx::prefetch_tensormap(&tm_a8);
        ptx::prefetch_tensormap(&tm_w8);
        if (has_outlier) {
            ptx::prefetch_tensormap(&tm_fa);
            ptx::prefetch_tensormap(&tm_fw);
        }
    }
    if (warp_idx == 1 && ptx::elect_one()) {
        for (int i = 0; i < T::kStages; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tmem_full_bar[i], 1);
            ptx::mbar_init(&f_full_bar[i], 1);
            ptx::mbar_init(&tmem_empty_bar[i], CTA * kStashEpiThreads / 32);
            ptx::mbar_init(&f_drained_bar[i], CTA * kStashEpiThreads / 32);
        }
        ptx::fence_barrier_init();
    }
    if (warp_idx == 2) {
        if constexpr (CTA == 2) {
            ptx::tmem_alloc_2cta(tmem_ptr_s, T::kTmemCols);
            ptx::tmem_relinquish_2cta();
        } else {
            ptx::tmem_alloc(tmem_ptr_s, T::kTmemCols);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before_sync();
    if constexpr (CTA == 2) ptx::cluster_sync(); else __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr_s;

    ptx::pdl_wait_prior_grid();
    ptx::pdl_launch_dependents();   // dependents may be scheduled as our CTAs retire; they wait for this grid's completion themselves

    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = (K + kBlockKBytes - 1) / kBlockKBytes;
    const int n_f = has_outlier ? kOutlierKBlocks : 0;
    const int kb_split = num_kb / 2;  // the outlier K-blocks are slotted in after this many int8 K-blocks

    if (warp_idx == 0) {
        if (ptx::elect_one()) {
            // ===================== TMA producer (every CTA) =====================
            int stage = 0;
            uint32_t phase = 0;
            auto load_block = [&](const CUtensorMap* ma, const CUtensorMap* mb, int k0, int m0, int n0) {
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                if (is_leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], T::kStageBytes * CTA);
                uint8_t* sA = ring + static_cast<size_t>(stage) * T::kStageBytes;
                uint8_t* sB = sA + T::kABytes;
                if constexpr (CTA == 2) {
                    ptx::tma_load_2d_2cta(sA, ma, &full_bar[stage], k0, m0, ptx::kEvictNormal);
                    ptx::tma_load_2d_2cta(sB, mb, &full_bar[stage], k0, n0, ptx::kEvictNormal);
                } else {
                    ptx::tma_load_2d(sA, ma, &full_bar[stage], k0, m0, ptx::kEvictNormal);
                    ptx::tma_load_2d(sB, mb, &full_bar[stage], k0, n0, ptx::kEvictNormal);
                }
                if (++stage == T::kStages) {
                    stage = 0;
                    phase ^= 1;
                }
            };
            for (int tile = group_id; tile < num_tiles; tile += num_groups) {
                const TileCoord tc = tile_coord(tile, m_tiles, n_tiles, group_m);
                const int m0 = tc.m_blk * T::kTileM + static_cast<int>(cta_rank) * kBlockM;
                const int n0 = tc.n_blk * BLOCK_N + static_cast<int>(cta_rank) * T::kLoadN;
                for (int kb = 0; kb < kb_split; ++kb) load_block(&tm_a8, &tm_w8, kb * kBlockKBytes, m0, n0);
                for (int it = 0; it < n_f; ++it) load_block(&tm_fa, &tm_fw, it * (kBlockKBytes / 2), m0, n0);
                for (int kb = kb_split; kb < num_kb; ++kb) load_block(&tm_a8, &tm_w8, kb * kBlockKBytes, m0, n0);
            }
        }
        __syncwarp();
    } else if (warp_idx == 1) {
        if (is_leader && ptx::elect_one()) {
            // ===================== MMA issuer (leader CTA only) =====================
            constexpr uint32_t idesc_i8 = ptx::make_idesc_i8(T::kTileM, BLOCK_N);
            constexpr uint32_t idesc_f16 = ptx::make_idesc_f16(T::kTileM, BLOCK_N);
            constexpr uint32_t kDescStep = kUmmaKBytes >> 4;
            const uint64_t desc_a0 = ptx::make_smem_desc_sw128(ptx::smem_u32(ring));
            const uint64_t desc_b0 = ptx::make_smem_desc_sw128(ptx::smem_u32(ring) + T::kABytes);
            int stage = 0;
            uint32_t phase = 0;
            bool ready = false;
            auto issue_block = [&](auto kind_tag, uint32_t tmem_d, bool first) {
                if (!ready) ptx::mbar_wait(&full_bar[stage], phase);
                const int nstage = (stage + 1 == T::kStages) ? 0 : stage + 1;
                const uint32_t nphase = (stage + 1 == T::kStages) ? phase ^ 1 : phase;
                ready = ptx::mbar_try_wait(&full_bar[nstage], nphase);
                ptx::tc_fence_after_sync();
                const uint64_t da = desc_a0 + static_cast<uint64_t>(stage) * (T::kStageBytes >> 4);
                const uint64_t db = desc_b0 + static_cast<uint64_t>(stage) * (T::kStageBytes >> 4);
#pragma unroll
                for (int k = 0; k < kBlockKBytes / kUmmaKBytes; ++k) {
                    const uint32_t acc = (first && k == 0) ? 0u : 1u;
                    if constexpr (decltype(kind_tag)::value == 0) {
                        if constexpr (CTA == 2) ptx::umma_f16_2cta(tmem_d, da + k * kDescStep, db + k * kDescStep, idesc_f16, acc);
                        else ptx::umma_f16(tmem_d, da + k * kDescStep, db + k * kDescStep, idesc_f16, acc);
                    } else {
                        if constexpr (CTA == 2) ptx::umma_i8_2cta(tmem_d, da + k * kDescStep, db + k * kDescStep, idesc_i8, acc);
                        else ptx::umma_i8(tmem_d, da + k * kDescStep, db + k * kDescStep, idesc_i8, acc);
                    }
                }
                if constexpr (CTA == 2) ptx::umma_commit_2cta(&empty_bar[stage]); else ptx::umma_commit(&empty_bar[stage]);
                stage = nstage;
                phase = nphase;
            };
            auto commit = [&](uint64_t* bar) {
                if constexpr (CTA == 2) ptx::umma_commit_2cta(bar); else ptx::umma_commit(bar);
            };
            int local_tile = 0;
            for (int tile = group_id; tile < num_tiles; tile += num_groups, ++local_tile) {
                const int b = local_tile & 1;
                const uint32_t use = static_cast<uint32_t>(local_tile >> 1);  // how often buffer b was used before
                const uint32_t tmem_i = tmem_base + b * BLOCK_N;
                const uint32_t tmem_f = tmem_base + (b ^ 1) * BLOCK_N;
                // buffer b last held the outlier accumulator of tile t-1 (drained) and, before that, the
                // int32 accumulator of tile t-2 (released before that drain by the same epilogue warps)
                if (has_outlier) {
                    if (local_tile >= 1) ptx::mbar_wait(&f_drained_bar[b], ((local_tile - 1) >> 1) & 1);
                } else {
                    ptx::mbar_wait(&tmem_empty_bar[b], (use & 1) ^ 1);
                }
                ptx::tc_fence_after_sync();
                for (int kb = 0; kb < kb_split; ++kb) issue_block(std::integral_constant<int, 1>{}, tmem_i, kb == 0);
                if (has_outlier) {
                    // the other buffer is free once the epilogue of tile t-1 has read its int32 accumulators
                    if (local_tile >= 1) ptx::mbar_wait(&tmem_empty_bar[b ^ 1], ((local_tile - 1) >> 1) & 1);
                    ptx::tc_fence_after_sync();
                    for (int it = 0; it < n_f; ++it) issue_block(std::integral_constant<int, 0>{}, tmem_f, it == 0);
                    commit(&f_full_bar[b ^ 1]);
                }
                for (int kb = kb_split; kb < num_kb; ++kb) issue_block(std::integral_constant<int, 1>{}, tmem_i, kb == 0);
                commit(&tmem_full_bar[b]);
            }
        }
        __syncwarp();
    } else if (warp_idx >= kEpilogueWarp0) {
        // ===================== epilogue (every CTA: its own 128 accumulator rows) =====================
        // 8 warps = two per TMEM lane quarter (a warp may only touch lanes 32*(warp_idx % 4)...): the pair
        // splits the 256 columns, which doubles the latency hiding of the LDTM / LDS / convert chains.
        const int quarter = warp_idx & 3;
        const int half = (warp_idx - kEpilogueWarp0) >> 2;  // 0: columns [0,128)   1: columns [128,256)
        const int et = threadIdx.x - kEpilogueWarp0 * 32;   // 0..255
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        constexpr int kCols = BLOCK_N / 2;                  // columns per epilogue thread
        const int col0 = half * kCols;
        uint4* my_stash = stash + et;  // thread-private, conflict-free: vector v lives at stash[v * 256 + et]
        auto arrive = [&](uint64_t* bar) {
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CTA == 2) ptx::mbar_arrive_cluster(bar, 0); else ptx::mbar_arrive(bar);
            }
        };
        int local_tile = 0;
        for (int tile = group_id; tile < num_tiles; tile += num_groups, ++local_tile) {
            const int b = local_tile & 1;
            const uint32_t use = static_cast<uint32_t>(local_tile >> 1);
            const TileCoord tc = tile_coord(tile, m_tiles, n_tiles, group_m);
            const int m0 = tc.m_blk * T::kTileM + static_cast<int>(cta_rank) * kBlockM;
            const int n0 = tc.n_blk * BLOCK_N;
            float* sbt = sb_s + b * BLOCK_N;
            float* bt = bias_sm + b * BLOCK_N;
            sbt[et] = (n0 + et < N) ? __half2float(scale_b[n0 + et]) : 0.0f;
            bt[et] = (epi.bias && n0 + et < N) ? __half2float(epi.bias[n0 + et]) : 0.0f;
            const int gm = m0 + row;
            const bool row_ok = gm < M;
            const float sa_f = row_ok ? __half2float(scale_a[gm]) : 0.0f;
            ptx::named_bar_sync(1, kStashEpiThreads);

            if (has_outlier) {
                // ---- drain the outlier accumulator (parked in the other buffer) into the stash as fp16
                ptx::mbar_wait(&f_full_bar[b ^ 1], use & 1);
                ptx::tc_fence_after_sync();
                const uint32_t t_f = tmem_base + lane_base + (b ^ 1) * BLOCK_N + col0;
                uint32_t va[32], vb[32];
                ptx::tmem_ld_32x32(t_f, va);
#pragma unroll
                for (int c = 0; c < kCols / 32; c += 2) {
                    ptx::tmem_ld_wait();
                    ptx::tmem_ld_32x32(t_f + (c + 1) * 32, vb);
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        uint32_t h[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const __half2 o = __floats2half2_rn(__uint_as_float(va[v * 8 + q * 2]), __uint_as_float(va[v * 8 + q * 2 + 1]));
                            h[q] = *reinterpret_cast<const uint32_t*>(&o);
                        }
                        my_stash[(c * 4 + v) * kStashEpiThreads] = make_uint4(h[0], h[1], h[2], h[3]);
                    }
                    ptx::tmem_ld_wait();
                    if (c + 2 < kCols / 32) ptx::tmem_ld_32x32(t_f + (c + 2) * 32, va);
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        uint32_t h[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const __half2 o = __floats2half2_rn(__uint_as_float(vb[v * 8 + q * 2]), __uint_as_float(vb[v * 8 + q * 2 + 1]));
                            h[q] = *reinterpret_cast<const uint32_t*>(&o);
                        }
                        my_stash[((c + 1) * 4 + v) * kStashEpiThreads] = make_uint4(h[0], h[1], h[2], h[3]);
                    }
                }
                arrive(&f_drained_bar[b ^ 1]);
            }

            // ---- dequantise the int32 accumulators of buffer b
            ptx::mbar_wait(&tmem_full_bar[b], use & 1);
            ptx::tc_fence_after_sync();
            const uint32_t t_i = tmem_base + lane_base + b * BLOCK_N + col0;
            __half* out_row = Out + static_cast<size_t>(gm) * N + n0 + col0;
            const float4* sb4 = reinterpret_cast<const float4*>(sbt + col0);
            uint32_t vi[2][32];
            ptx::tmem_ld_32x32(t_i, vi[0]);
#pragma unroll
            for (int c = 0; c < kCols / 32; ++c) {
                ptx::tmem_ld_wait();
                if (c + 1 < kCols / 32) ptx::tmem_ld_32x32(t_i + (c + 1) * 32, vi[(c + 1) & 1]);
                const uint32_t* v = vi[c & 1];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 f = make_uint4(0u, 0u, 0u, 0u);
                    if (has_outlier) f = my_stash[(c * 4 + g) * kStashEpiThreads];
                    const uint32_t fw[4] = {f.x, f.y, f.z, f.w};
                    const float4 s0 = sb4[c * 8 + g * 2], s1 = sb4[c * 8 + g * 2 + 1];
                    const float sbv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                    uint32_t packed[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int j = g * 8 + q * 2;
                        const float p0 = __fmul_rn(sbv[q * 2], sa_f);
                        const float p1 = __fmul_rn(sbv[q * 2 + 1], sa_f);
                        const float2 of = __half22float2(*reinterpret_cast<const __half2*>(&fw[q]));
                        const float r0 = __fmaf_rn(__int2float_rn(static_cast<int>(v[j])), p0, of.x);
                        const float r1 = __fmaf_rn(__int2float_rn(static_cast<int>(v[j + 1])), p1, of.y);
                        const __half2 r = epi_finish(r0, r1, bt + col0 + c * 32 + j, epi);
                        packed[q] = *reinterpret_cast<const uint32_t*>(&r);
                    }
                    if (row_ok && n0 + col0 + c * 32 + g * 8 + 8 <= N)
                        ptx::st_global_v4(out_row + c * 32 + g * 8, packed[0], packed[1], packed[2], packed[3]);
                }
            }
            arrive(&tmem_empty_bar[b]);
        }
    }

    ptx::tc_fence_before_sync();
    if constexpr (CTA == 2) ptx::cluster_sync(); else __syncthreads();
    if (warp_idx == 2) {
        if constexpr (CTA == 2) ptx::tmem_dealloc_2cta(tmem_base, T::kTmemCols);
        else ptx::tmem_dealloc(tmem_base, T::kTmemCols);
    }
}

// =================================================================================================
// Stream-K variant of the wide-tile kernel, for shapes whose tile count does not fill the machine
// evenly (decode batches: M = 512 gives 96 tiles of 256x256 for 74 CTA pairs; M = 32 gives 48).
// The iteration space (tile, K-block) is cut into one contiguous, equally long span per CTA pair.
// A span may start in the middle of a tile and end in the middle of another one:
//   * the segment that contains K-block 0 of a tile is the tile's FINISHER: it also runs the outlier
//     MMAs, collects the int32 partial sums of the other segments and does the dequant epilogue;
//   * every other segment is a PEER: its epilogue dumps the raw int32 accumulators to a per-worker
//     slot in the workspace and raises a flag.
// Integer accumulation is associative, so the result is bit-identical to the unsplit kernel.  The
// finisher's segment is the LAST thing its worker does while peer segments are the FIRST thing their
// workers do, so the finisher practically never waits.  All workers are co-resident (grid <= #SMs).
// With stream_k = 0 the kernel degenerates to the strided whole-tile schedule of the stash kernel.
// Debug timeline (mixq_debug_set_trace): 16 x uint64 nanosecond stamps per CTA; null in production.
__device__ unsigned long long* g_trace = nullptr;
__device__ __forceinline__ void trace_stamp(int slot) {
    unsigned long long* t = g_trace;
    if (t) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        t[blockIdx.x * 16 + slot] = now;
    }
}

struct SegIter {
    int stream_k, num_tiles, num_kb, num_groups;
    long long u, u_end;
    int tile_cursor;
    __device__ void init(int sk, int nt, int nkb, int gid, int ng) {
        stream_k = sk; num_tiles = nt; num_kb = nkb; num_groups = ng;
        const long long total = static_cast<long long>(nt) * nkb;
        u = total * gid / ng;
        u_end = total * (gid + 1) / ng;
        tile_cursor = gid;
    }
    __device__ bool next(int& tile, int& kb0, int& kb1) {
        if (stream_k) {
            if (u >= u_end) return false;
            tile = static_cast<int>(u / num_kb);
            kb0 = static_cast<int>(u - static_cast<long long>(tile) * num_kb);
            const long long rem = u_end - u;
            kb1 = static_cast<int>(rem < num_kb - kb0 ? kb0 + rem : num_kb);
            u += kb1 - kb0;
            return true;
        }
        if (tile_cursor >= num_tiles) return false;
        tile = tile_cursor;
        tile_cursor += num_groups;
        kb0 = 0;
        kb1 = num_kb;
        return true;
    }
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// publish everything this thread has observed/written (cumulativity) and bump a counter that may live on a peer GPU
__device__ __forceinline__ void signal_add_sys(uint32_t* p, uint32_t v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void signal_add_relaxed_sys(uint32_t* p, uint32_t v) {
    asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
// Wait until a counter that peers on other GPUs bump reaches `expect`.  Ranks may be skewed by seconds (first call after a
// per-rank checkpoint load, a debugger, a long kernel ahead of this one on one rank), so this never traps: it polls, backs
// off with nanosleep, and only after `timeout_ns` (0 = wait for ever) gives up, raising the error word of this rank's counter
// block -- the launch then completes with an invalid result that mixq_allreduce_check() reports to the host, instead of a
// sticky context error.
__device__ __forceinline__ bool wait_counter_sys(const uint32_t* p, uint32_t expect, unsigned long long timeout_ns, uint32_t* err_word) {
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (ld_acquire_sys(p) < expect) {
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
// 16-byte store to an NVSwitch multicast address: the switch replicates it into every rank's copy of the buffer
__device__ __forceinline__ void multimem_st_v4(void* mc_ptr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_ptr), "r"(a), "r"(b), "r"(c), "r"(d)
                 : "memory");
}

#include "gemm_decode.cuh"
#include "gemm_fat.cuh"

// ---- fused all-reduce of a row-parallel linear -------------------------------------------------
// Every rank computes the fp16 partial of every output tile; tile t is OWNED by rank t % world.
//   phase A (GEMM epilogue):  the partial tile leaves shared memory through 2 KB bulk copies (one per 32x32 chunk)
//                             straight into the OWNER's staging area (peer memory over NVLink).  No per-tile
//                             signalling: a system-scope fence after remote stores costs as long as the stores need to
//                             land (measured ~16 us for 4 MB), so it is paid ONCE per CTA, after its last tile, followed
//                             by a release bump of every rank's `pushed` counter.
//   phase B (same kernel):    once `pushed` shows every CTA of every rank, the 32-row groups of the tiles this rank
//                             owns are spread over all CTAs: fp32 sum in rank order 0..world-1 (deterministic and
//                             identical on every rank), one rounding to fp16, the finished [32 x BLOCK_N] slab goes
//                             through shared memory and one TMA store per rank into EVERY rank's Out, then a release
//                             bump of every rank's `done` counter by the number of slabs delivered.
//   exit:                     CTA 0 leaves only when its own `done` counter shows every row group of every tile,
//                             i.e. when this rank's Out is complete.  It re-arms `done` and the `pushed` word of this
//                             launch and advances the launch epoch; consecutive launches alternate between two
//                             `pushed` words so a fast peer's next launch can never race with the re-arming.
// Counter block of a rank (uint32 words): 0 = done, 1 = launch epoch, 2..3 = pushed[epoch & 1], 4 = error flag.
constexpr int kArWordDone = 0, kArWordEpoch = 1, kArWordPushed = 2, kArWordError = 4;   // 4: a peer wait timed out (sticky until cleared by the host)
constexpr size_t kArCounterBytes = 128;
struct alignas(64) ArParams {
    CUtensorMap tm_out[MIXQ_MAX_RANKS];     // every rank's Out [M, N], box 32 rows x BLOCK_N columns, no swizzle
    uint32_t* counters[MIXQ_MAX_RANKS];     // every rank's counter block
    uint8_t* stage_push[MIXQ_MAX_RANKS];    // region of rank i's staging area reserved for THIS rank's partials; a unit
                                            // (slot, 32-row group) is BLOCK_N/32 chunks of 32 rows x 32 columns, each the
                                            // 2 KB SWIZZLE_64B image of the shared-memory tile it was copied from
    const __half* stage_local;              // this rank's whole staging area: [world][slots][256][BLOCK_N] fp16
    __half* out_mc;                         // NVSwitch multicast mapping of Out (a store lands in every rank's Out), or null
    int world, rank, slots;
    unsigned long long timeout_ns;          // give up waiting for a peer after this long (0 = never), see wait_counter_sys
};
struct ArNone {
    int unused;
};

template <class T, bool AR>
__global__ void __launch_bounds__(kStashThreads, 1)
mixq_gemm_dequant_streamk_kernel(const __grid_constant__ CUtensorMap tm_a8, const __grid_constant__ CUtensorMap tm_w8,
                                 const __grid_constant__ CUtensorMap tm_fa, const __grid_constant__ CUtensorMap tm_fw,
                                 const __grid_constant__ CUtensorMap tm_out,
                                 const __half* __restrict__ scale_a, const __half* __restrict__ scale_b,
                                 __half* __restrict__ Out, int M, int N, int K, int has_outlier, int m_tiles,
                                 int n_tiles, int group_m, int stream_k, uint4* __restrict__ sk_slots,
                                 uint32_t* __restrict__ sk_flags,
                                 const __grid_constant__ std::conditional_t<AR, ArParams, ArNone> ar, EpiArgs epi) {
    constexpr int BLOCK_N = T::kBlockN;
    constexpr int CTA = T::kCta;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    uint4* stash = reinterpret_cast<uint4*>(ring + static_cast<size_t>(T::kStages) * T::kStageBytes);
    float* sb_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(stash) + T::kStashBytes);
    float* bias_sm = sb_s + 2 * BLOCK_N;   // staged bias, double-buffered like sb_s
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sb_s + 4 * BLOCK_N);
    uint64_t* empty_bar = full_bar + T::kStages;
    uint64_t* tmem_full_bar = empty_bar + T::kStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint64_t* f_full_bar = tmem_empty_bar + 2;
    uint64_t* f_drained_bar = f_full_bar + 2;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(f_drained_bar + 2);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CTA == 2) ? ptx::cluster_ctarank() : 0u;
    const bool is_leader = cta_rank == 0;
    const int group_id = blockIdx.x / CTA;
    const int num_groups = gridDim.x / CTA;
    if (threadIdx.x == 0) trace_stamp(0);

    if (warp_idx == 0 && ptx::elect_one()) {
        ptx::prefetch_tensormap(&tm_a8);
        ptx::prefetch_tensormap(&tm_w8);
        ptx::prefetch_tensormap(&tm_out);
        if (has_outlier) {
            ptx::prefetch_tensormap(&tm_fa);
            ptx::prefetch_tensormap(&tm_fw);
        }
    }
    if (warp_idx == 1 && ptx::elect_one()) {
        for (int i = 0; i < T::kStages; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tmem_full_bar[i], 1);
            ptx::mbar_init(&f_full_bar[i], 1);
            ptx::mbar_init(&tmem_empty_bar[i], CTA * kStashEpiThreads / 32);
            ptx::mbar_init(&f_drained_bar[i], CTA * kStashEpiThreads / 32);
        }
        ptx::fence_barrier_init();
    }
    if (warp_idx == 2) {
        if constexpr (CTA == 2) {
            ptx::tmem_alloc_2cta(tmem_ptr_s, T::kTmemCols);
            ptx::tmem_relinquish_2cta();
        } else {
            ptx::tmem_alloc(tmem_ptr_s, T::kTmemCols);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before_sync();
    if constexpr (CTA == 2) ptx::cluster_sync(); else __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (threadIdx.x == 0) trace_stamp(1);
    ptx::pdl_wait_prior_grid();
    ptx::pdl_launch_dependents();   // dependents may be scheduled as our CTAs retire; they wait for this grid's completion themselves

    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = (K + kBlockKBytes - 1) / kBlockKBytes;
    SegIter seg;
    seg.init(stream_k, num_tiles, num_kb, group_id, num_groups);
    int tile, kb0, kb1;

    if (warp_idx == 0) {
        if (ptx::elect_one()) {
            // ===================== TMA producer (every CTA) =====================
            int stage = 0;
            uint32_t phase = 0;
            auto load_block = [&](const CUtensorMap* ma, const CUtensorMap* mb, int k0, int m0, int n0) {
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                if (is_leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], T::kStageBytes * CTA);
                uint8_t* sA = ring + static_cast<size_t>(stage) * T::kStageBytes;
                uint8_t* sB = sA + T::kABytes;
                if constexpr (CTA == 2) {
                    ptx::tma_load_2d_2cta(sA, ma, &full_bar[stage], k0, m0, ptx::kEvictNormal);
                    ptx::tma_load_2d_2cta(sB, mb, &full_bar[stage], k0, n0, ptx::kEvictNormal);
                } else {
                    ptx::tma_load_2d(sA, ma, &full_bar[stage], k0, m0, ptx::kEvictNormal);
                    ptx::tma_load_2d(sB, mb, &full_bar[stage], k0, n0, ptx::kEvictNormal);
                }
                if (++stage == T::kStages) {
                    stage = 0;
                    phase ^= 1;
                }
            };
            trace_stamp(2);
            int s = 0;
            while (seg.next(tile, kb0, kb1)) {
                const TileCoord tc = tile_coord(tile, m_tiles, n_tiles, group_m);
                const int m0 = tc.m_blk * T::kTileM + static_cast<int>(cta_rank) * kBlockM;
                const int n0 = tc.n_blk * BLOCK_N + static_cast<int>(cta_rank) * T::kLoadN;
                // the outlier K-blocks belong to the segment that holds the LAST K-block of the tile; they come
                // first in a worker's first segment (both accumulator buffers are free) and mid-way otherwise
                const int n_f = (has_outlier && kb1 == num_kb) ? kOutlierKBlocks : 0;
                const int split = (s++ == 0) ? kb0 : kb0 + (kb1 - kb0) / 2;
                for (int kb = kb0; kb < split; ++kb) load_block(&tm_a8, &tm_w8, kb * kBlockKBytes, m0, n0);
                for (int it = 0; it < n_f; ++it) load_block(&tm_fa, &tm_fw, it * (kBlockKBytes / 2), m0, n0);
                for (int kb = split; kb < kb1; ++kb) load_block(&tm_a8, &tm_w8, kb * kBlockKBytes, m0, n0);
            }
        }
        __syncwarp();
    } else if (warp_idx == 1) {
        if (is_leader && ptx::elect_one()) {
            // ===================== MMA issuer (leader CTA only) =====================
            constexpr uint32_t idesc_i8 = ptx::make_idesc_i8(T::kTileM, BLOCK_N);
            constexpr uint32_t idesc_f16 = ptx::make_idesc_f16(T::kTileM, BLOCK_N);
            constexpr uint32_t kDescStep = kUmmaKBytes >> 4;
            const uint64_t desc_a0 = ptx::make_smem_desc_sw128(ptx::smem_u32(ring));
            const uint64_t desc_b0 = ptx::make_smem_desc_sw128(ptx::smem_u32(ring) + T::kABytes);
            int stage = 0;
            uint32_t phase = 0;
            bool ready = false;
            auto issue_block = [&](auto kind_tag, uint32_t tmem_d, bool first) {
                if (!ready) ptx::mbar_wait(&full_bar[stage], phase);
                const int nstage = (stage + 1 == T::kStages) ? 0 : stage + 1;
                const uint32_t nphase = (stage + 1 == T::kStages) ? phase ^ 1 : phase;
                ready = ptx::mbar_try_wait(&full_bar[nstage], nphase);
                ptx::tc_fence_after_sync();
                const uint64_t da = desc_a0 + static_cast<uint64_t>(stage) * (T::kStageBytes >> 4);
                const uint64_t db = desc_b0 + static_cast<uint64_t>(stage) * (T::kStageBytes >> 4);
#pragma unroll
                for (int k = 0; k < kBlockKBytes / kUmmaKBytes; ++k) {
                    const uint32_t acc = (first && k == 0) ? 0u : 1u;
                    if constexpr (decltype(kind_tag)::value == 0) {
                        if constexpr (CTA == 2) ptx::umma_f16_2cta(tmem_d, da + k * kDescStep, db + k * kDescStep, idesc_f16, acc);
                        else ptx::umma_f16(tmem_d, da + k * kDescStep, db + k * kDescStep, idesc_f16, acc);
                    } else {
                        if constexpr (CTA == 2) ptx::umma_i8_2cta(tmem_d, da + k * kDescStep, db + k * kDescStep, idesc_i8, acc);
                        else ptx::umma_i8(tmem_d, da + k * kDescStep, db + k * kDescStep, idesc_i8, acc);
                    }
                }
                if constexpr (CTA == 2) ptx::umma_commit_2cta(&empty_bar[stage]); else ptx::umma_commit(&empty_bar[stage]);
                stage = nstage;
                phase = nphase;
            };
            auto commit = [&](uint64_t* bar) {
                if constexpr (CTA == 2) ptx::umma_commit_2cta(bar); else ptx::umma_commit(bar);
            };
            // completions this thread has caused / must have seen, per accumulator buffer
            uint32_t n_int[2] = {0, 0};     // segments that used buffer x for int32 accumulators (= tmem_full commits)
            uint32_t n_f[2] = {0, 0};       // outlier accumulators parked in buffer x (= f_full commits)
            bool f_pending[2] = {false, false};
            int s = 0;
            while (seg.next(tile, kb0, kb1)) {
                const int b = s & 1;
                const bool has_f = has_outlier && kb1 == num_kb;
                const uint32_t tmem_i = tmem_base + b * BLOCK_N;
                const uint32_t tmem_f = tmem_base + (b ^ 1) * BLOCK_N;
                // buffer b must be free: its previous int32 tenant read by the epilogue, a parked outlier
                // accumulator drained
                if (n_int[b] > 0) ptx::mbar_wait(&tmem_empty_bar[b], (n_int[b] - 1) & 1);
                if (f_pending[b]) {
                    ptx::mbar_wait(&f_drained_bar[b], (n_f[b] - 1) & 1);
                    f_pending[b] = false;
                }
                ptx::tc_fence_after_sync();
                const int split = (s == 0) ? kb0 : kb0 + (kb1 - kb0) / 2;
                for (int kb = kb0; kb < split; ++kb) issue_block(std::integral_constant<int, 1>{}, tmem_i, kb == kb0);
                if (has_f) {
                    if (n_int[b ^ 1] > 0) ptx::mbar_wait(&tmem_empty_bar[b ^ 1], (n_int[b ^ 1] - 1) & 1);
                    ptx::tc_fence_after_sync();
                    for (int it = 0; it < kOutlierKBlocks; ++it) issue_block(std::integral_constant<int, 0>{}, tmem_f, it == 0);
                    commit(&f_full_bar[b ^ 1]);
                    ++n_f[b ^ 1];
                    f_pending[b ^ 1] = true;
                }
                for (int kb = split; kb < kb1; ++kb) issue_block(std::integral_constant<int, 1>{}, tmem_i, kb == kb0);
                commit(&tmem_full_bar[b]);
                if (s == 0) trace_stamp(3);
                ++n_int[b];
                ++s;
            }
            trace_stamp(4);
        }
        __syncwarp();
    } else if (warp_idx >= kEpilogueWarp0) {
        // ===================== epilogue (8 warps; every CTA: its own 128 accumulator rows) =====================
        // Each warp owns 32 rows x 128 columns = 4 chunks of 32 rows x 32 columns.  A chunk has a 2 KB
        // shared-memory tile (32 rows x 64 B, SWIZZLE_64B so that the row-per-thread accesses are
        // conflict free).  The tile first holds the fp16 outlier product (the "stash"), is overwritten
        // in place with the fp16 result, and leaves through one TMA store per chunk: no per-thread
        // global stores, and the tensor map clips rows >= M / columns >= N.
        const int quarter = warp_idx & 3;
        const int half = (warp_idx - kEpilogueWarp0) >> 2;
        const int ew = warp_idx - kEpilogueWarp0;           // 0..7
        const int et = threadIdx.x - kEpilogueWarp0 * 32;   // 0..255
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        constexpr int kCols = BLOCK_N / 2;
        constexpr int kChunks = kCols / 32;                 // 4
        const int col0 = half * kCols;
        uint8_t* warp_tiles = reinterpret_cast<uint8_t*>(stash) + ew * (kChunks * 2048);
        const uint32_t swz = static_cast<uint32_t>((lane >> 1) & 3);
        auto vec_ptr = [&](int c, int v) {                  // 16-byte vector v (8 columns) of this thread's row in chunk c
            return reinterpret_cast<uint4*>(warp_tiles + c * 2048 + lane * 64 + ((static_cast<uint32_t>(v) ^ swz) << 4));
        };
        auto arrive = [&](uint64_t* bar) {
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CTA == 2) ptx::mbar_arrive_cluster(bar, 0); else ptx::mbar_arrive(bar);
            }
        };
        // slot of worker w, CTA rank r: 32 vectors x 256 threads x 16 B = 128 KB, thread-major (coalesced)
        auto slot_of = [&](int w) { return sk_slots + (static_cast<size_t>(w) * CTA + cta_rank) * (48 * kStashEpiThreads) + et; };  // 32 >= kCols / 4
        // ... followed by 16 vectors x 256 threads x 16 B = 64 KB for the fp16 outlier product of a tail segment
        auto fslot_of = [&](int w) { return slot_of(w) + 32 * kStashEpiThreads; };
        auto flag_of = [&](int w) { return sk_flags + static_cast<size_t>(w) * CTA + cta_rank; };
        uint32_t n_int[2] = {0, 0}, n_f[2] = {0, 0};
        int s = 0;
        while (seg.next(tile, kb0, kb1)) {
            const int b = s & 1;
            const bool finisher = kb0 == 0;
            const bool has_f = has_outlier && kb1 == num_kb;   // this segment computed the outlier product
            const TileCoord tc = tile_coord(tile, m_tiles, n_tiles, group_m);
            const int m0 = tc.m_blk * T::kTileM + static_cast<int>(cta_rank) * kBlockM;
            const int n0 = tc.n_blk * BLOCK_N;
            const int gm = m0 + row;
            float* sbt = sb_s + b * BLOCK_N;
            float sa_f = 0.0f;
            if (finisher) {
                if (et < BLOCK_N) {
                    sbt[et] = (n0 + et < N) ? __half2float(scale_b[n0 + et]) : 0.0f;
                    bias_sm[b * BLOCK_N + et] = (epi.bias && n0 + et < N) ? __half2float(epi.bias[n0 + et]) : 0.0f;
                }
                sa_f = gm < M ? __half2float(scale_a[gm]) : 0.0f;
                // the previous tile's TMA stores must have finished READING this warp's tiles
                if (lane == 0) ptx::tma_store_wait_read<0>();
            }
            int ar_owner = 0, ar_slot = 0;
            if constexpr (AR) {
                ar_owner = tile % ar.world;
                ar_slot = tile / ar.world;
            }
            ptx::named_bar_sync(1, kStashEpiThreads);

            if (has_f) {
                // ---- drain the outlier accumulator (parked in the other buffer) into the tiles as fp16
                ptx::mbar_wait(&f_full_bar[b ^ 1], n_f[b ^ 1] & 1);
                ++n_f[b ^ 1];
                ptx::tc_fence_after_sync();
                const uint32_t t_f = tmem_base + lane_base + (b ^ 1) * BLOCK_N + col0;
                uint32_t vf[2][32];
                ptx::tmem_ld_32x32(t_f, vf[0]);
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    ptx::tmem_ld_wait();
                    if (c + 1 < kChunks) ptx::tmem_ld_32x32(t_f + (c + 1) * 32, vf[(c + 1) & 1]);
                    const uint32_t* va = vf[c & 1];
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        uint32_t h[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const __half2 o = __floats2half2_rn(__uint_as_float(va[v * 8 + q * 2]), __uint_as_float(va[v * 8 + q * 2 + 1]));
                            h[q] = *reinterpret_cast<const uint32_t*>(&o);
                        }
                        if (finisher) *vec_ptr(c, v) = make_uint4(h[0], h[1], h[2], h[3]);
                        else fslot_of(group_id)[(c * 4 + v) * kStashEpiThreads] = make_uint4(h[0], h[1], h[2], h[3]);  // ship it
                    }
                }
                arrive(&f_drained_bar[b ^ 1]);
            }

            ptx::mbar_wait(&tmem_full_bar[b], n_int[b] & 1);
            ++n_int[b];
            ptx::tc_fence_after_sync();
            if (et == 0) trace_stamp(s == 0 ? 5 : 6);
            const uint32_t t_i = tmem_base + lane_base + b * BLOCK_N + col0;

            if (!finisher) {
                // ---- PEER: dump the raw int32 partial sums, then raise this worker's flag
                uint4* slot = slot_of(group_id);
                uint32_t v[32];
#pragma unroll 1
                for (int c = 0; c < kChunks; ++c) {
                    ptx::tmem_ld_32x32(t_i + c * 32, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        slot[(c * 8 + g) * kStashEpiThreads] = make_uint4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
                }
                arrive(&tmem_empty_bar[b]);
                ptx::named_bar_sync(2, kStashEpiThreads);  // every thread's slot stores precede the flag
                if (et == 0) {
                    __threadfence();
                    st_release_gpu(flag_of(group_id), 1u);
                }
            } else {
                // ---- FINISHER: [collect the peers' partial sums,] dequantise, TMA-store
                int n_peers = 0;
                if (kb1 < num_kb) {
                    // the rest of this tile [kb1, num_kb) was computed by the following workers
                    const long long total = static_cast<long long>(num_tiles) * num_kb;
                    const long long tile_end = static_cast<long long>(tile + 1) * num_kb;
                    int p = group_id + 1;
                    while (p < num_groups && total * p / num_groups < tile_end) {
                        ++n_peers;
                        ++p;
                    }
                    if (et == 0) {
                        for (int q = 1; q <= n_peers; ++q) {
                            uint32_t spins = 0;
                            while (ld_acquire_gpu(flag_of(group_id + q)) == 0u) {
                                if (++spins > (1u << 23)) __trap();
                            }
                        }
                    }
                    ptx::named_bar_sync(2, kStashEpiThreads);
                }
                if (et == 0) trace_stamp(8);
                const float4* sb4 = reinterpret_cast<const float4*>(sbt + col0);
                uint32_t vi[2][32];
                ptx::tmem_ld_32x32(t_i, vi[0]);
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    uint4 pv[8];
                    if (n_peers > 0) {  // issue the first peer's loads before waiting for TMEM
                        const uint4* slot = slot_of(group_id + 1);
#pragma unroll
                        for (int g = 0; g < 8; ++g) pv[g] = __ldcg(slot + (c * 8 + g) * kStashEpiThreads);
                    }
                    ptx::tmem_ld_wait();
                    if (c + 1 < kChunks) ptx::tmem_ld_32x32(t_i + (c + 1) * 32, vi[(c + 1) & 1]);
                    uint32_t* v = vi[c & 1];
                    for (int q = 1; q <= n_peers; ++q) {
                        if (q > 1) {
                            const uint4* slot = slot_of(group_id + q);
#pragma unroll
                            for (int g = 0; g < 8; ++g) pv[g] = __ldcg(slot + (c * 8 + g) * kStashEpiThreads);
                        }
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            v[g * 4] += pv[g].x;
                            v[g * 4 + 1] += pv[g].y;
                            v[g * 4 + 2] += pv[g].z;
                            v[g * 4 + 3] += pv[g].w;
                        }
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4* tp = vec_ptr(c, g);
                        uint4 f = make_uint4(0u, 0u, 0u, 0u);
                        if (has_f) f = *tp;                                            // own outlier product (stash)
                        else if (has_outlier) f = __ldcg(fslot_of(group_id + n_peers) + (c * 4 + g) * kStashEpiThreads);  // the tail segment's
                        const uint32_t fw[4] = {f.x, f.y, f.z, f.w};
                        const float4 s0 = sb4[c * 8 + g * 2], s1 = sb4[c * 8 + g * 2 + 1];
                        const float sbv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                        uint32_t packed[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int j = g * 8 + q * 2;
                            const float p0 = __fmul_rn(sbv[q * 2], sa_f);
                            const float p1 = __fmul_rn(sbv[q * 2 + 1], sa_f);
                            const float2 of = __half22float2(*reinterpret_cast<const __half2*>(&fw[q]));
                            const float r0 = __fmaf_rn(__int2float_rn(static_cast<int>(v[j])), p0, of.x);
                            const float r1 = __fmaf_rn(__int2float_rn(static_cast<int>(v[j + 1])), p1, of.y);
                            const __half2 r = epi_finish(r0, r1, bias_sm + b * BLOCK_N + col0 + c * 32 + j, epi);
                            packed[q] = *reinterpret_cast<const uint32_t*>(&r);
                        }
                        *tp = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    }
                    // hand the finished 32x32 tile to the TMA engine
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (AR)   // the partial goes to the tile owner's staging area (peer memory), 2 KB contiguous
                            ptx::bulk_store_1d(ar.stage_push[ar_owner] +
                                                   (static_cast<size_t>((ar_slot * 8 + static_cast<int>(cta_rank) * 4 + quarter) *
                                                                        (BLOCK_N / 32) + half * kChunks + c) << 11),
                                               warp_tiles + c * 2048, 2048);
                        else
                            ptx::tma_store_2d(&tm_out, warp_tiles + c * 2048, n0 + col0 + c * 32, m0 + quarter * 32);
                        ptx::tma_store_commit();
                    }
                }
                if (et == 0) trace_stamp(9);
                arrive(&tmem_empty_bar[b]);
                if (n_peers > 0) {
                    // re-arm the flags for the next launch once every thread of this CTA has read the slots
                    ptx::named_bar_sync(2, kStashEpiThreads);
                    if (et == 0)
                        for (int q = 1; q <= n_peers; ++q) *flag_of(group_id + q) = 0u;
                }
            }
            ++s;
        }
        if (et == 0) trace_stamp(10);
        if (lane == 0) ptx::tma_store_wait_all<0>();  // outstanding output tiles fully written before the CTA retires
        if (et == 0) trace_stamp(11);

        if constexpr (AR) {
            // ===================== phase B: reduce the row groups this rank owns, broadcast the result ===========
            const int world = ar.world;
            uint32_t* my_cnt = ar.counters[ar.rank];
            const uint32_t parity = ld_acquire_sys(my_cnt + kArWordEpoch) & 1u;   // stable until CTA 0 exits this launch
            ptx::named_bar_sync(2, kStashEpiThreads);          // every warp's bulk copies have completed (wait_all above)
            if (et == 0) {
                // this CTA's partial tiles are on their way: one fence for all of them, then tell every owner
                fence_sys();
                for (int pr = 0; pr < world; ++pr)
                    signal_add_relaxed_sys(ar.counters[(ar.rank + 1 + pr) % world] + kArWordPushed + parity, 1u);
                trace_stamp(12);
            }
            const int owned = num_tiles > ar.rank ? (num_tiles - ar.rank + world - 1) / world : 0;
            const int units = owned * 8;                       // (owned slot, 32-row group)
            constexpr int kSlabBytes = 32 * BLOCK_N * 2;       // one unit, row-major [32][BLOCK_N] fp16
            constexpr int kNumSlabs = T::kStashBytes / kSlabBytes;                 // 4
            constexpr int kUnitVecs = 32 * BLOCK_N / 8;                            // 16-byte vectors per unit and source
            constexpr int kVecPerThread = kUnitVecs / kStashEpiThreads;            // 4 (BLOCK_N = 256), 3 (192)
            static_assert(kUnitVecs % kStashEpiThreads == 0 && kNumSlabs >= 2, "phase B work split");
            uint8_t* slabs = reinterpret_cast<uint8_t*>(stash);
            const size_t region_vecs = static_cast<size_t>(ar.slots) * T::kTileM * BLOCK_N / 8;   // vectors per source rank
            uint32_t n_done = 0;
            int k = 0;
            if (ar.out_mc != nullptr) {
                // ---- broadcast through the switch: each warp reduces whole 512-byte rows (lane = chunk * 4 + 16-byte slot of the
                // chunk-major staging image) and issues ONE multicast store per row; the egress of this phase drops from
                // (world - 1) copies to one.
                if (et == 0 && static_cast<int>(blockIdx.x) < units) {   // a CTA without units must not poll a word CTA 0 may already have re-armed
                    wait_counter_sys(my_cnt + kArWordPushed + parity, static_cast<uint32_t>(world) * gridDim.x, ar.timeout_ns,
                                     my_cnt + kArWordError);
                    trace_stamp(13);
                }
                ptx::named_bar_sync(2, kStashEpiThreads);
                const int chunk = lane >> 2, slot = lane & 3;
                const bool lane_ok = chunk < BLOCK_N / 32;
                for (int u = blockIdx.x; u < units; u += gridDim.x) {
                    const int j = u >> 3, rg = u & 7;
                    const TileCoord tc = tile_coord(ar.rank + j * world, m_tiles, n_tiles, group_m);
                    const uint4* unit = reinterpret_cast<const uint4*>(ar.stage_local) + static_cast<size_t>(u) * kUnitVecs;
                    uint4 v[4][MIXQ_MAX_RANKS];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = ew * 4 + i;                           // row of the 32-row group
#pragma unroll
                        for (int sr = 0; sr < MIXQ_MAX_RANKS; ++sr)
                            if (sr < world && lane_ok) v[i][sr] = __ldcg(unit + sr * region_vecs + chunk * 128 + r * 4 + slot);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = ew * 4 + i;
                        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int sr = 0; sr < MIXQ_MAX_RANKS; ++sr) {
                            if (sr < world) {
                                const uint32_t w[4] = {v[i][sr].x, v[i][sr].y, v[i][sr].z, v[i][sr].w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                                    acc[2 * e] = sr == 0 ? f.x : acc[2 * e] + f.x;       // rank order; world = 1 is the identity
                                    acc[2 * e + 1] = sr == 0 ? f.y : acc[2 * e + 1] + f.y;
                                }
                            }
                        }
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __half2 h = __floats2half2_rn(acc[2 * e], acc[2 * e + 1]);
                            pk[e] = *reinterpret_cast<const uint32_t*>(&h);
                        }
                        const int gm = tc.m_blk * T::kTileM + rg * 32 + r;
                        const int col = tc.n_blk * BLOCK_N + chunk * 32 + ((slot ^ ((r >> 1) & 3)) << 3);   // un-swizzle the slot
                        if (lane_ok && gm < M && col < N)
                            multimem_st_v4(ar.out_mc + static_cast<size_t>(gm) * N + col, pk[0], pk[1], pk[2], pk[3]);
                    }
                    ++n_done;
                }
                ptx::named_bar_sync(2, kStashEpiThreads);      // every thread's multicast stores precede the fence below
            } else
            for (int u = blockIdx.x; u < units; u += gridDim.x, ++k) {
                const int j = u >> 3, rg = u & 7;
                const TileCoord tc = tile_coord(ar.rank + j * world, m_tiles, n_tiles, group_m);
                uint8_t* slab = slabs + (k % kNumSlabs) * kSlabBytes;
                if (et == 0) {
                    ptx::tma_store_wait_read<kNumSlabs - 1>();   // the store that last used this slab has read it
                    if (k == 0) {
                        // every CTA of every rank has pushed (and fenced) its partial tiles
                        wait_counter_sys(my_cnt + kArWordPushed + parity, static_cast<uint32_t>(world) * gridDim.x, ar.timeout_ns,
                                         my_cnt + kArWordError);
                        trace_stamp(13);
                    }
                }
                ptx::named_bar_sync(2, kStashEpiThreads);
                // unit u of every source rank is one contiguous run of kUnitVecs vectors (chunk-major)
                const uint4* src = reinterpret_cast<const uint4*>(ar.stage_local) + static_cast<size_t>(u) * kUnitVecs + et;
#pragma unroll
                for (int i0 = 0; i0 < kVecPerThread; i0 += 2) {
                    uint4 v[2][MIXQ_MAX_RANKS];
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                        for (int sr = 0; sr < MIXQ_MAX_RANKS; ++sr)
                            if (i0 + q < kVecPerThread && sr < world)
                                v[q][sr] = __ldcg(src + sr * region_vecs + (i0 + q) * kStashEpiThreads);
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        if (i0 + q >= kVecPerThread) continue;
                        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int sr = 0; sr < MIXQ_MAX_RANKS; ++sr) {
                            if (sr < world) {
                                const uint32_t w[4] = {v[q][sr].x, v[q][sr].y, v[q][sr].z, v[q][sr].w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                                    acc[2 * e] = sr == 0 ? f.x : acc[2 * e] + f.x;       // rank order; world = 1 is the identity
                                    acc[2 * e + 1] = sr == 0 ? f.y : acc[2 * e + 1] + f.y;
                                }
                            }
                        }
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __half2 h = __floats2half2_rn(acc[2 * e], acc[2 * e + 1]);
                            pk[e] = *reinterpret_cast<const uint32_t*>(&h);
                        }
                        // vector vv of the unit: chunk vv / 128, row (vv % 128) / 4, 16-byte slot vv % 4 of the swizzled row
                        // image (slot = 8-column group ^ ((row >> 1) & 3), as vec_ptr stored it)
                        const int vv = et + (i0 + q) * kStashEpiThreads;
                        const int vrow = (vv & 127) >> 2;
                        *reinterpret_cast<uint4*>(slab + vrow * (BLOCK_N * 2) + (vv >> 7) * 64 + (((vv & 3) ^ ((vrow >> 1) & 3)) << 4)) =
                            make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
                ptx::fence_proxy_async_smem();
                ptx::named_bar_sync(2, kStashEpiThreads);
                if (et == 0) {
                    // one 32-row x BLOCK_N box into EVERY rank's Out (the tensor maps clip the M / N edges)
                    for (int pr = 0; pr < world; ++pr)
                        ptx::tma_store_2d(&ar.tm_out[pr], slab, tc.n_blk * BLOCK_N, tc.m_blk * T::kTileM + rg * 32);
                    ptx::tma_store_commit();
                    ++n_done;
                }
            }
            if (et == 0) {
                trace_stamp(14);
                ptx::tma_store_wait_all<0>();
                if (n_done) {
                    fence_sys();
                    for (int pr = 0; pr < world; ++pr)   // row groups of Out complete
                        signal_add_relaxed_sys(ar.counters[(ar.rank + 1 + pr) % world] + kArWordDone, n_done);
                }
                if (blockIdx.x == 0) {
                    // this rank's Out is complete when every row group of every tile has been delivered
                    wait_counter_sys(my_cnt + kArWordDone, static_cast<uint32_t>(num_tiles) * 8u, ar.timeout_ns, my_cnt + kArWordError);
                    // A rank that owns no tile (fewer tiles than ranks) never waited for `pushed` itself: make sure every
                    // peer's bump of this launch has landed before the word is re-armed, or a late one would leak into the
                    // launch after next.
                    wait_counter_sys(my_cnt + kArWordPushed + parity, static_cast<uint32_t>(world) * gridDim.x, ar.timeout_ns,
                                     my_cnt + kArWordError);
                    // every CTA of this rank that had row groups to reduce is past its `pushed` wait (its groups are
                    // counted in `done`): re-arm for the launch after next and flip the epoch
                    my_cnt[kArWordDone] = 0u;
                    my_cnt[kArWordPushed + parity] = 0u;
                    my_cnt[kArWordEpoch] = parity ^ 1u;
                }
                trace_stamp(15);
            }
        }
    }

    if (threadIdx.x == kEpilogueWarp0 * 32) trace_stamp(7);
    ptx::tc_fence_before_sync();
    if constexpr (CTA == 2) ptx::cluster_sync(); else __syncthreads();
    if (warp_idx == 2) {
        if constexpr (CTA == 2) ptx::tmem_dealloc_2cta(tmem_base, T::kTmemCols);
        else ptx::tmem_dealloc(tmem_base, T::kTmemCols);
    }
}

// ---------------------------------------------------------------- host side
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// 2D row-major [rows, cols] tensor of `elem_bytes` elements, box = `box_bytes` x box_rows with the matching
// swizzle (128 B for the operand tiles, 64 B for the output tiles).
int encode_tmap(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t rows, uint64_t cols,
                uint32_t box_rows, int box_bytes) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return set_error(MIXQ_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * static_cast<uint64_t>(elem_bytes)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_bytes / elem_bytes), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     box_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : box_bytes > 128 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        static thread_local char buf[160];
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed: CUresult %d (rows=%llu cols=%llu elem=%d)", (int)r,
                 (unsigned long long)rows, (unsigned long long)cols, elem_bytes);
        return set_error(MIXQ_ERR_CUDA, buf);
    }
    return MIXQ_OK;
}

// Tensor maps are pure functions of (base, shape, box): a small cache keeps the per-enqueue host cost at a few
// compares once a layer's pointers have been seen (weights are engine constants, the workspace is reused).
struct TmapKey {
    const void* base;
    uint64_t rows, cols;
    uint32_t box_rows;
    int dt, elem_bytes, box_bytes;
    bool operator==(const TmapKey& o) const {
        return base == o.base && rows == o.rows && cols == o.cols && box_rows == o.box_rows && dt == o.dt &&
               elem_bytes == o.elem_bytes && box_bytes == o.box_bytes;
    }
};
constexpr int kTmapCacheSize = 256;
struct TmapCache {
    std::mutex lock;
    TmapKey keys[kTmapCacheSize];
    CUtensorMap maps[kTmapCacheSize];
    bool valid[kTmapCacheSize] = {};
    unsigned next = 0;
};
TmapCache& tmap_cache() {
    static TmapCache c;
    return c;
}

// 2D row-major [rows, cols] tensor of `elem_bytes` elements, box = `box_bytes` x box_rows with the matching
// swizzle (128 B for the operand tiles, 64 B for the output tiles).
int make_tmap(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t rows, uint64_t cols,
              uint32_t box_rows, int box_bytes = kBlockKBytes) {
    const TmapKey key{base, rows, cols, box_rows, static_cast<int>(dt), elem_bytes, box_bytes};
    TmapCache& c = tmap_cache();
    const unsigned h = static_cast<unsigned>((reinterpret_cast<uintptr_t>(base) >> 8) * 2654435761u + rows * 40503u + box_rows) % kTmapCacheSize;
    {
        std::lock_guard<std::mutex> g(c.lock);
        for (int probe = 0; probe < 4; ++probe) {
            const unsigned i = (h + probe) % kTmapCacheSize;
            if (c.valid[i] && c.keys[i] == key) {
                *map = c.maps[i];
                return MIXQ_OK;
            }
        }
    }
    const int rc = encode_tmap(map, dt, elem_bytes, base, rows, cols, box_rows, box_bytes);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(c.lock);
    unsigned slot = h;
    for (int probe = 0; probe < 4; ++probe) {
        const unsigned i = (h + probe) % kTmapCacheSize;
        if (!c.valid[i]) {
            slot = i;
            break;
        }
        if (probe == 3) slot = (h + (c.next++ & 3)) % kTmapCacheSize;
    }
    c.keys[slot] = key;
    c.maps[slot] = *map;
    c.valid[slot] = true;
    return MIXQ_OK;
}

template <class T>
struct KernelOf {
    static constexpr int kThreads = kGemmThreads;
    static auto get() { return mixq_gemm_dequant_kernel<T>; }
};
template <int CTA, int STAGES>
struct KernelOf<StashTraits<CTA, STAGES, 256>> {
    static constexpr int kThreads = kStashThreads;
    static auto get() { return mixq_gemm_dequant_stash_kernel<StashTraits<CTA, STAGES, 256>>; }
};

template <int CTA, int STAGES, int BLOCK_N = 256>
struct StreamKTraits : StashTraits<CTA, STAGES, BLOCK_N> {};
template <int CTA, int STAGES, int BLOCK_N>
struct KernelOf<StreamKTraits<CTA, STAGES, BLOCK_N>> {
    static constexpr int kThreads = kStashThreads;
    static auto get() { return mixq_gemm_dequant_streamk_kernel<StashTraits<CTA, STAGES, BLOCK_N>, false>; }
    static auto get_ar() { return mixq_gemm_dequant_streamk_kernel<StashTraits<CTA, STAGES, BLOCK_N>, true>; }
};
template <class T>
struct ARowsOf : std::integral_constant<int, kBlockM> {
    static constexpr bool plain = false;
};
template <int CTA, int BLOCK_N, int ACC, int STAGES, int A_ROWS>
struct ARowsOf<GemmTraits<CTA, BLOCK_N, ACC, STAGES, A_ROWS>> : std::integral_constant<int, A_ROWS> {
    static constexpr bool plain = true;   // mixq_gemm_dequant_kernel
};
template <class T>
struct IsStreamK : std::false_type {};
template <int CTA, int STAGES, int BLOCK_N>
struct IsStreamK<StreamKTraits<CTA, STAGES, BLOCK_N>> : std::true_type {};

template <class T>
int launch_cfg(const void* A8, const void* W8, const void* scale_a, const void* scale_b, const void* fp_A,
               const void* fp_weight, void* Out, int64_t M, int64_t N, int64_t K, cudaStream_t stream, bool pdl,
               void* sk_ws = nullptr, int stream_k = 0, const mixq_peer_group* pg = nullptr, EpiArgs epi = EpiArgs{nullptr, 0},
               LaunchOpts opts = LaunchOpts{}, int ksplit = 1) {
    const DeviceInfo& dev = device_info();
    CUtensorMap tm_a8, tm_w8, tm_fa, tm_fw;
    int rc;
    constexpr int a_rows = ARowsOf<T>::value;
    if ((rc = make_tmap(&tm_a8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, A8, M, K, a_rows))) return rc;
    if ((rc = make_tmap(&tm_w8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, W8, N, K, T::kLoadN))) return rc;
    const int has_outlier = (fp_A && fp_weight) ? 1 : 0;
    if (has_outlier) {
        if ((rc = make_tmap(&tm_fa, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, fp_A, M, MIXQ_NUM_OUTLIERS, a_rows))) return rc;
        if ((rc = make_tmap(&tm_fw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, fp_weight, N, MIXQ_NUM_OUTLIERS, T::kLoadN)))
            return rc;
    } else {
        tm_fa = tm_a8;
        tm_fw = tm_w8;
    }
    auto kern = KernelOf<T>::get();
    cudaError_t e = cudaSuccess;
    static std::atomic<uint64_t> attr_set_mask{0};   // per template instantiation: one bit per device ordinal
    if (!(attr_set_mask.load(std::memory_order_acquire) >> dev.device & 1)) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(T::kSmemBytes));
        if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(gemm_dequant)");
        attr_set_mask.fetch_or(uint64_t(1) << dev.device, std::memory_order_release);
    }

    const int m_tiles = static_cast<int>((M + T::kTileM - 1) / T::kTileM);
    const int n_tiles = static_cast<int>((N + T::kBlockN - 1) / T::kBlockN);
    const int64_t num_tiles = static_cast<int64_t>(m_tiles) * n_tiles;
    if (num_tiles > (1ll << 30)) return set_error(MIXQ_ERR_UNSUPPORTED, "gemm_dequant: too many tiles");
    const int64_t max_groups = usable_sms(opts) / T::kCta;
    int grid = static_cast<int>(num_tiles < max_groups ? num_tiles : max_groups) * T::kCta;
    if (ksplit > 1) {
        // K-split (plain one-CTA tiles, one row-block of tokens): a cluster of `ksplit` CTAs per tile, every tile its own cluster
        constexpr bool kPlain = !IsStreamK<T>::value && ARowsOf<T>::plain && T::kCta == 1;
        const int64_t row_warps = (M + 31) / 32 < 4 ? (M + 31) / 32 : 4;
        if (!kPlain || m_tiles != 1 || ksplit > 8 || (ksplit & (ksplit - 1)) ||
            static_cast<int64_t>(ksplit - 1) * row_warps * T::kBlockN * 128 > static_cast<int64_t>(T::kStages) * T::kStageBytes ||
            (K + kBlockKBytes - 1) / kBlockKBytes < 2 * ksplit)
            return set_error(MIXQ_ERR_UNSUPPORTED, "gemm_dequant: K-split needs a one-CTA tile configuration (ids 1, 3, 15), M <= 128, a power-of-two split <= 8 and enough K");
        grid = static_cast<int>(num_tiles) * ksplit;
    }
    if (IsStreamK<T>::value && sk_ws && stream_k) {
        // one equal span of (tile, K-block) units per CTA group; never more groups than units
        const int64_t units = num_tiles * ((K + kBlockKBytes - 1) / kBlockKBytes);
        grid = static_cast<int>(units < max_groups ? units : max_groups) * T::kCta;
    }
    // band height of the grouped rasterisation in tiles (4096 rows by default: measured best of 512..4096 at
    // M = 65536, W is then re-streamed from L2/HBM 16 instead of 64 times); MIXQ_GROUP_M overrides for tuning
    static const int env_group_m = [] {
        const char* e = getenv("MIXQ_GROUP_M");
        return e ? atoi(e) : 0;
    }();
    const int group_m = env_group_m > 0 ? env_group_m : 4096 / T::kTileM;

    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(KernelOf<T>::kThreads);
    cfg.dynamicSmemBytes = T::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (T::kCta == 2 || ksplit > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = T::kCta == 2 ? 2 : ksplit;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    if constexpr (IsStreamK<T>::value) {
        if (pg) {
            // fused all-reduce: the epilogue stores partial tiles into the owners' staging areas (see ArParams)
            ArParams ar{};
            ar.world = pg->world;
            ar.rank = pg->rank;
            ar.slots = static_cast<int>((num_tiles + pg->world - 1) / pg->world);
            const size_t region = static_cast<size_t>(ar.slots) * T::kTileM * T::kBlockN * 2;   // bytes per source rank
            if (region * pg->world > pg->staging_bytes)
                return set_error(MIXQ_ERR_WORKSPACE, "gemm_dequant_allreduce: staging area too small (mixq_allreduce_staging_size)");
            if (kArCounterBytes > pg->counter_bytes)
                return set_error(MIXQ_ERR_WORKSPACE, "gemm_dequant_allreduce: counter block too small (mixq_allreduce_counter_size)");
            for (int i = 0; i < pg->world; ++i) {
                if (!pg->out[i] || !pg->staging[i] || !pg->counters[i])
                    return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant_allreduce: null peer pointer");
                ar.counters[i] = static_cast<uint32_t*>(pg->counters[i]);
                ar.stage_push[i] = static_cast<uint8_t*>(pg->staging[i]) + region * pg->rank;
                if ((rc = make_tmap(&ar.tm_out[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, pg->out[i], M, N, 32, T::kBlockN * 2)))
                    return rc;
            }
            {   // peer-wait timeout: MIXQ_AR_TIMEOUT_MS (default 60 s, 0 = wait for ever)
                static const long long timeout_ms = [] {
                    const char* e = std::getenv("MIXQ_AR_TIMEOUT_MS");
                    return e ? std::atoll(e) : 60000ll;
                }();
                ar.timeout_ns = timeout_ms > 0 ? static_cast<unsigned long long>(timeout_ms) * 1000000ull : 0ull;
            }
            ar.stage_local = static_cast<const __half*>(pg->staging[pg->rank]);
            ar.out_mc = static_cast<__half*>(pg->out_multicast);
            auto kern_ar = KernelOf<T>::get_ar();
            static std::atomic<uint64_t> ar_attr_set_mask{0};
            if (!(ar_attr_set_mask.load(std::memory_order_acquire) >> dev.device & 1)) {
                e = cudaFuncSetAttribute(kern_ar, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(T::kSmemBytes));
                if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(gemm_dequant_allreduce)");
                ar_attr_set_mask.fetch_or(uint64_t(1) << dev.device, std::memory_order_release);
            }
            e = cudaLaunchKernelEx(&cfg, kern_ar, tm_a8, tm_w8, tm_fa, tm_fw, tm_a8 /* tm_out unused */,
                                   static_cast<const __half*>(scale_a), static_cast<const __half*>(scale_b),
                                   static_cast<__half*>(Out), static_cast<int>(M), static_cast<int>(N),
                                   static_cast<int>(K), has_outlier, m_tiles, n_tiles, group_m, 0,
                                   static_cast<uint4*>(nullptr), static_cast<uint32_t*>(nullptr), ar, EpiArgs{nullptr, 0});
            if (e != cudaSuccess) return set_cuda_error(e, "launch gemm_dequant_allreduce");
            count_launch();
            return MIXQ_OK;
        }
        CUtensorMap tm_out;
        if ((rc = make_tmap(&tm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, Out, M, N, 32, 64))) return rc;
        uint32_t* flags = static_cast<uint32_t*>(sk_ws);
        uint4* slots = reinterpret_cast<uint4*>(static_cast<uint8_t*>(sk_ws) + kStreamKFlagBytes);
        // a worker is a peer at most once; spans shorter than a tile would need more than the reserved slots
        e = cudaLaunchKernelEx(&cfg, kern, tm_a8, tm_w8, tm_fa, tm_fw, tm_out, static_cast<const __half*>(scale_a),
                               static_cast<const __half*>(scale_b), static_cast<__half*>(Out), static_cast<int>(M),
                               static_cast<int>(N), static_cast<int>(K), has_outlier, m_tiles, n_tiles, group_m,
                               (sk_ws && stream_k) ? 1 : 0, slots, flags, ArNone{0}, epi);
    } else {
        e = cudaLaunchKernelEx(&cfg, kern, tm_a8, tm_w8, tm_fa, tm_fw, static_cast<const __half*>(scale_a),
                               static_cast<const __half*>(scale_b), static_cast<__half*>(Out), static_cast<int>(M),
                               static_cast<int>(N), static_cast<int>(K), has_outlier, m_tiles, n_tiles, group_m, ksplit, epi);
    }
    if (e != cudaSuccess) return set_cuda_error(e, "launch gemm_dequant");
    count_launch();
    return MIXQ_OK;
}

// ---- decode kernel (gemm_decode.cuh) ---------------------------------------------------------------------------
using DecodeT = DecodeTraits<6>;
constexpr size_t kDecodeSlotRegion = static_cast<size_t>(kStreamKMaxWorkers) * kStreamKSlotBytes;   // shared with config 8
static_assert(static_cast<size_t>(kDecodeMaxWorkers) * 2 * kDecodeSlotBytesPerCta <= kDecodeSlotRegion, "decode slots fit");
static_assert(kDecodeFlagBytes == kStreamKFlagBytes && kDecodeMaxWorkers * 2 * 4 <= kDecodeFlagBytes, "decode flags fit");

size_t decode_out0_bytes(int64_t M, int64_t N) {
    if (M <= 0 || N <= 0) return 0;
    const int64_t rows = M < kDecodeMaxM ? M : kDecodeMaxM;
    const int64_t tiles = ((rows + DecodeT::kTileM - 1) / DecodeT::kTileM) * ((N + DecodeT::kBlockN - 1) / DecodeT::kBlockN);
    return static_cast<size_t>(tiles) * 2 * kDecodeOut0BytesPerCta;
}

// Split decision of the two-phase schedule: the `r = tiles mod pairs` remainder tiles are cut along K when that
// leaves every pair a span worth having; r == tiles (fewer tiles than pairs) splits everything.
struct DecodePlan {
    int groups, sk_tiles, gran;
};
DecodePlan plan_decode(int64_t num_tiles, int64_t num_kb, int64_t max_groups, bool allow_split) {
    DecodePlan p{static_cast<int>(num_tiles < max_groups ? num_tiles : max_groups), 0, 1};
    if (!allow_split || max_groups < 2 || max_groups > kDecodeMaxWorkers) return p;
    const int64_t r = num_tiles % max_groups;
    if (r == 0) return p;
    const int64_t span = r * num_kb / max_groups;   // K-blocks per pair in the split region
    if (span < 8) return p;                          // not worth a fix-up
    p.groups = static_cast<int>(max_groups);
    p.sk_tiles = static_cast<int>(r);
    p.gran = (num_kb % 4 == 0 && span >= 16) ? 4 : 1;
    return p;
}

int launch_decode(const void* A8, const void* W8, const void* scale_a, const void* scale_b, const void* fp_A,
                  const void* fp_weight, void* Out, int64_t M, int64_t N, int64_t K, cudaStream_t stream, bool pdl,
                  void* sk_ws, bool allow_split, EpiArgs epi, LaunchOpts opts) {
    using T = DecodeT;
    const DeviceInfo& dev = device_info();
    CUtensorMap tm_a8, tm_w8, tm_fa, tm_fw;
    int rc;
    if ((rc = make_tmap(&tm_a8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, A8, M, K, kBlockM))) return rc;
    if ((rc = make_tmap(&tm_w8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, W8, N, K, T::kLoadN))) return rc;
    const int has_outlier = (fp_A && fp_weight) ? 1 : 0;
    if (has_outlier) {
        if ((rc = make_tmap(&tm_fa, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, fp_A, M, MIXQ_NUM_OUTLIERS, kBlockM))) return rc;
        if ((rc = make_tmap(&tm_fw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, fp_weight, N, MIXQ_NUM_OUTLIERS, T::kLoadN))) return rc;
    } else {
        tm_fa = tm_a8;
        tm_fw = tm_w8;
    }
    auto kern = mixq_gemm_dequant_decode_kernel<T>;
    cudaError_t e = cudaSuccess;
    static std::atomic<uint64_t> attr_set_mask{0};
    if (!(attr_set_mask.load(std::memory_order_acquire) >> dev.device & 1)) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(T::kSmemBytes));
        if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(gemm_dequant_decode)");
        attr_set_mask.fetch_or(uint64_t(1) << dev.device, std::memory_order_release);
    }
    const int m_tiles = static_cast<int>((M + T::kTileM - 1) / T::kTileM);
    const int n_tiles = static_cast<int>((N + T::kBlockN - 1) / T::kBlockN);
    const int64_t num_tiles = static_cast<int64_t>(m_tiles) * n_tiles;
    const int64_t num_kb = (K + kBlockKBytes - 1) / kBlockKBytes;
    const DecodePlan plan = plan_decode(num_tiles, num_kb, usable_sms(opts) / 2, allow_split);

    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(plan.groups * 2);
    cfg.blockDim = dim3(kStashThreads);
    cfg.dynamicSmemBytes = T::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    uint8_t* ws = static_cast<uint8_t*>(sk_ws);
    e = cudaLaunchKernelEx(&cfg, kern, tm_a8, tm_w8, tm_fa, tm_fw, static_cast<const __half*>(scale_a),
                           static_cast<const __half*>(scale_b), static_cast<__half*>(Out), static_cast<int>(M),
                           static_cast<int>(N), static_cast<int>(K), has_outlier, m_tiles, n_tiles, m_tiles /* one band */,
                           plan.sk_tiles, plan.gran, reinterpret_cast<uint32_t*>(ws),
                           reinterpret_cast<uint4*>(ws + kDecodeFlagBytes),
                           reinterpret_cast<uint4*>(ws + kDecodeFlagBytes + kDecodeSlotRegion), epi);
    if (e != cudaSuccess) return set_cuda_error(e, "launch gemm_dequant_decode");
    count_launch();
    return MIXQ_OK;
}

// ---- fat-tile decode kernel (gemm_fat.cuh) ---------------------------------------------------------------------
// One tile of 256 x Nt per CTA pair and wave: Nt is the widest multiple of 16 (<= 336) that still spreads the N range
// of a row-block over all the pairs serving it.
struct FatPlan {
    int Nt, n_tiles, m_tiles, stages, stage_bytes, waves;
};
FatPlan plan_fat(int64_t M, int64_t N, int pairs, bool gated = false, int epi_warps = 8) {
    // N: accumulator columns of the whole problem (gated: gate + up = 2 x the output channels; a tile then holds Nt / 2
    // channels of each, so Nt is a multiple of 32)
    FatPlan p{};
    const int gran = gated ? 32 : 16;
    const int max_nt = kFatMaxN / gran * gran;
    p.m_tiles = static_cast<int>((M + 255) / 256);
    const int per_m = pairs / p.m_tiles > 0 ? pairs / p.m_tiles : 1;
    const int64_t waves = (N + static_cast<int64_t>(max_nt) * per_m - 1) / (static_cast<int64_t>(max_nt) * per_m);
    const int64_t want = waves * per_m;                      // tiles per row-block
    int64_t nt = ((N + want - 1) / want + gran - 1) / gran * gran;
    if (nt > max_nt) nt = max_nt;
    if (nt < gran) nt = gran;
    p.Nt = static_cast<int>(nt);
    p.n_tiles = static_cast<int>((N + nt - 1) / nt);
    p.stage_bytes = kBlockM * kBlockKBytes + (p.Nt / 2) * kBlockKBytes;
    int st = static_cast<int>((227 * 1024 - 1024 - fat_fixed_bytes(epi_warps)) / p.stage_bytes);
    p.stages = st > kFatMaxStages ? kFatMaxStages : st;
    const int64_t tiles = static_cast<int64_t>(p.m_tiles) * p.n_tiles;
    p.waves = static_cast<int>((tiles + pairs - 1) / pairs);
    return p;
}

// the second projection of the gated mode (gemm_fat.cuh): up-projection weights of the same shape as the gate's
struct FatGated {
    const void* W8_up;
    const void* scale_b_up;
    const void* fp_weight_up;
};

// CTA clusters of 4 with the fat kernel's shared-memory footprint that can be resident at once (split-K mode); cached per device
int fat_max_clusters4() {
    static std::atomic<int> cached[kMaxDevices];
    const DeviceInfo& dev = device_info();
    if (dev.device < 0) return 0;
    int v = cached[dev.device].load(std::memory_order_acquire);
    if (v != 0) return v > 0 ? v : 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(dev.num_sms / 4 * 4));
    cfg.blockDim = dim3(kStashThreads);
    cfg.dynamicSmemBytes = 227 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e = cudaFuncSetAttribute(mixq_gemm_dequant_fat_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, mixq_gemm_dequant_fat_kernel<8>, &cfg);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    cached[dev.device].store(n > 0 ? n : -1, std::memory_order_release);
    return n;
}

// Split-K plan (two CTA pairs per tile in a cluster of 4, gemm_fat.cuh): only when every tile gets its own cluster (one wave) and
// the partial-sum landing zone fits the ring
bool plan_fat_split(int64_t M, int64_t N, int64_t K, const LaunchOpts& opts, FatPlan* out) {
    int clusters = fat_max_clusters4();
    const int by_sms = usable_sms(opts) / 4;
    if (by_sms < clusters) clusters = by_sms;
    if (clusters < 1 || (K + kBlockKBytes - 1) / kBlockKBytes < 8) return false;
    const FatPlan pl = plan_fat(M, N, clusters);
    const int n_chunks = pl.Nt / 16;
    int c_split = ((n_chunks + 1) / 2 + 1) & ~1;
    if (c_split > n_chunks) c_split = n_chunks;
    if (pl.waves != 1 || pl.stages < 4 || (c_split + 1) / 2 > kFatMaxPartPairs) return false;
    if (static_cast<size_t>(8) * c_split * 2048 > static_cast<size_t>(pl.stages) * pl.stage_bytes) return false;
    *out = pl;
    return true;
}

int launch_fat(const void* A8, const void* W8, const void* scale_a, const void* scale_b, const void* fp_A, const void* fp_weight,
               void* Out, int64_t M, int64_t N, int64_t K, cudaStream_t stream, bool pdl, EpiArgs epi, LaunchOpts opts,
               const FatGated* gated = nullptr, int ksplit = 1, int epi_warps = 12) {
    const DeviceInfo& dev = device_info();
    const int pairs = usable_sms(opts) / 2;
    if (pairs < 1) return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant: sm_limit leaves no CTA pair");
    if (ksplit == 2) epi_warps = 8;   // the split-K landing zone is laid out for two warps per lane quarter
    FatPlan pl = plan_fat(M, gated ? 2 * N : N, pairs, gated != nullptr, epi_warps);
    if (ksplit == 2 && (gated || !plan_fat_split(M, N, K, opts, &pl)))
        return set_error(MIXQ_ERR_UNSUPPORTED, "gemm_dequant: the split-K fat-tile schedule needs one cluster of 4 per tile");
    if (pl.stages < 3) return set_error(MIXQ_ERR_UNSUPPORTED, "gemm_dequant: fat tile does not fit shared memory");
    const int N1 = gated ? pl.Nt / 2 : (pl.Nt > 256 ? 256 : pl.Nt), N2 = pl.Nt - N1;
    const void* W2 = gated ? gated->W8_up : W8;
    const void* fw2 = gated ? gated->fp_weight_up : fp_weight;
    CUtensorMap tm_a8, tm_w1, tm_w2, tm_fa, tm_fw1, tm_fw2;
    int rc;
    if ((rc = make_tmap(&tm_a8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, A8, M, K, kBlockM))) return rc;
    if ((rc = make_tmap(&tm_w1, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, W8, N, K, N1 / 2))) return rc;
    tm_w2 = tm_w1;
    if (N2 && (rc = make_tmap(&tm_w2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, W2, N, K, N2 / 2))) return rc;
    const int has_outlier = (fp_A && fp_weight) ? 1 : 0;
    tm_fa = tm_a8;
    tm_fw1 = tm_fw2 = tm_w1;
    if (has_outlier) {
        if ((rc = make_tmap(&tm_fa, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, fp_A, M, MIXQ_NUM_OUTLIERS, kBlockM))) return rc;
        if ((rc = make_tmap(&tm_fw1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, fp_weight, N, MIXQ_NUM_OUTLIERS, N1 / 2))) return rc;
        tm_fw2 = tm_fw1;
        if (N2 && (rc = make_tmap(&tm_fw2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, fw2, N, MIXQ_NUM_OUTLIERS, N2 / 2))) return rc;
    }
    CUtensorMap tm_out;
    if ((rc = make_tmap(&tm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, Out, M, N, 32, 64))) return rc;
    auto kern = epi_warps == 8 ? mixq_gemm_dequant_fat_kernel<8> : mixq_gemm_dequant_fat_kernel<12>;
    constexpr int kMaxSmem = 227 * 1024;
    cudaError_t e = cudaSuccess;
    static std::atomic<uint64_t> attr_set_mask{0};
    if (!(attr_set_mask.load(std::memory_order_acquire) >> dev.device & 1)) {
        e = cudaFuncSetAttribute(mixq_gemm_dequant_fat_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mixq_gemm_dequant_fat_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(gemm_dequant_fat)");
        attr_set_mask.fetch_or(uint64_t(1) << dev.device, std::memory_order_release);
    }
    const int64_t tiles = static_cast<int64_t>(pl.m_tiles) * pl.n_tiles;
    const int workers = ksplit == 2 ? static_cast<int>(tiles) : pairs;     // split-K: exactly one cluster of 4 per tile
    const int groups = static_cast<int>(tiles < workers ? tiles : workers);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(groups * 2 * ksplit);
    cfg.blockDim = dim3((kEpilogueWarp0 + (epi_warps == 8 ? 8 : 12)) * 32);
    cfg.dynamicSmemBytes = 1024 + static_cast<size_t>(pl.stages) * pl.stage_bytes + fat_fixed_bytes(epi_warps == 8 ? 8 : 12);
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2 * ksplit;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    e = cudaLaunchKernelEx(&cfg, kern, tm_a8, tm_w1, tm_w2, tm_fa, tm_fw1, tm_fw2, tm_out, static_cast<const __half*>(scale_a),
                           static_cast<const __half*>(scale_b), static_cast<const __half*>(gated ? gated->scale_b_up : scale_b),
                           static_cast<__half*>(Out), static_cast<int>(M), static_cast<int>(N), static_cast<int>(K), has_outlier,
                           pl.m_tiles, pl.n_tiles, pl.Nt, pl.stages, gated ? 1 : 0, ksplit, epi);
    if (e != cudaSuccess) return set_cuda_error(e, "launch gemm_dequant_fat");
    count_launch();
    return MIXQ_OK;
}

// out[i] *= other[i] (fp16): the last step of the gated linear where the two projections ran as separate GEMMs (M > 1024)
__global__ void mixq_mul_inplace_kernel(__half2* __restrict__ out, const __half2* __restrict__ other, size_t n2) {
    ptx::pdl_wait_prior_grid();
    ptx::pdl_launch_dependents();   // dependents may be scheduled as our CTAs retire; they wait for this grid's completion themselves
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n2; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        out[i] = __hmul2(out[i], other[i]);
}

}  // namespace

size_t decode_workspace_bytes(int64_t M, int64_t N) { return streamk_workspace_bytes() + decode_out0_bytes(M, N); }

void fat_plan_for(int64_t M, int64_t N, int pairs, int gated, int epi_warps, int* out5) {
    const FatPlan p = plan_fat(M, gated ? 2 * N : N, pairs, gated != 0, epi_warps);
    out5[0] = p.Nt; out5[1] = p.n_tiles; out5[2] = p.m_tiles; out5[3] = p.stages; out5[4] = p.waves;
}

int set_trace_buffer(void* dev_buf) {
    unsigned long long* p = static_cast<unsigned long long*>(dev_buf);
    cudaError_t e = cudaMemcpyToSymbol(g_trace, &p, sizeof(p));
    return e == cudaSuccess ? MIXQ_OK : set_cuda_error(e, "cudaMemcpyToSymbol(g_trace)");
}

size_t streamk_workspace_bytes() { return kStreamKFlagBytes + static_cast<size_t>(kStreamKMaxWorkers) * kStreamKSlotBytes; }

int launch_gemm_dequant(const void* A8, const void* W8, const void* scale_a, const void* scale_b, const void* fp_A,
                        const void* fp_weight, void* Out, int64_t M, int64_t N, int64_t K, cudaStream_t stream,
                        bool pdl, void* sk_ws, size_t sk_ws_bytes, bool sk_flags_clean, const void* bias, int act,
                        LaunchOpts opts) {
    if (M == 0 || N == 0) return MIXQ_OK;
    // gemm_config = tile id + 100 x K-split (one-CTA tiles 1 / 3 / 15 with M <= 128: 2, 4 or 8 CTAs of a cluster share a tile along K)
    int ksplit = opts.cfg >= 200 ? opts.cfg / 100 : 1;
    if (opts.cfg >= 200) opts.cfg %= 100;
    if (opts.cfg < 0 || opts.cfg >= kCfgCount) return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant: unknown config id");
    if (ksplit > 1 && opts.cfg != kCfgN128x2 && opts.cfg != kCfgN64x2 && opts.cfg != kCfgN32x2)
        return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant: K-split goes with the one-CTA tile ids 1, 3, 15");
    if (opts.sm_limit < 0) return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant: negative sm_limit");
    if (act != MIXQ_ACT_NONE && act != MIXQ_ACT_SILU) return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant: unknown activation");
    const EpiArgs epi{static_cast<const __half*>(bias), act};
    if (!A8 || !W8 || !scale_a || !scale_b || !Out) return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant: null pointer");
    if ((fp_A == nullptr) != (fp_weight == nullptr))
        return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant: fp_A and fp_weight must both be given or both be null");
    if (M < 0 || N < 0 || K <= 0 || M > INT32_MAX || N > INT32_MAX || K > INT32_MAX)
        return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant: bad dimensions");
    if ((K & 15) != 0) return set_error(MIXQ_ERR_UNSUPPORTED, "gemm_dequant: K must be a multiple of 16 (TMA row pitch)");
    if ((N & 7) != 0) return set_error(MIXQ_ERR_UNSUPPORTED, "gemm_dequant: N must be a multiple of 8 (16-byte stores)");
    const uintptr_t al = reinterpret_cast<uintptr_t>(A8) | reinterpret_cast<uintptr_t>(W8) |
                         reinterpret_cast<uintptr_t>(Out) | reinterpret_cast<uintptr_t>(fp_A) |
                         reinterpret_cast<uintptr_t>(fp_weight);
    if (al & 15) return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant: tensors must be 16-byte aligned");
    if (!device_info().ok) return set_error(MIXQ_ERR_CUDA, "no usable sm_100 device");

    const bool sk_ok = sk_ws && sk_ws_bytes >= streamk_workspace_bytes() && (reinterpret_cast<uintptr_t>(sk_ws) & 15) == 0;
    const bool decode_ok = sk_ws && M <= kDecodeMaxM && sk_ws_bytes >= decode_workspace_bytes(M, N) &&
                           (reinterpret_cast<uintptr_t>(sk_ws) & 15) == 0;
    int cfg = opts.cfg;
    if (cfg == kCfgAuto) {
        // Pick the tile shape with the lowest estimated time = K-blocks per worker x (cycles per K-block) + exposed
        // tail.  Cycles per K-block are the measured steady-state figures (profiles/): the 128-wide
        // tiles are shared-memory bound, only the 256-wide pair tile runs near the tensor peak, and wider tiles
        // leave a longer un-overlapped final epilogue.  A single row-block of tokens cannot use a CTA pair.
        const int64_t nkb = (K + kBlockKBytes - 1) / kBlockKBytes + kOutlierKBlocks;
        struct Cand {
            int id, tile_m, tile_n, cta;
            int64_t kb_cycles, tail_cycles;
        };
        // one row-block of tokens (M <= 128): a CTA is bound by the bytes that land in its shared memory per K-block -- the A box
        // (32 / 64 / 128 rows, see launch_cfg) plus its W rows -- at ~48 B/clk; the 64-wide tile halves the W part and doubles the
        // CTAs that stream weights, which pays when N / 128 tiles leave SMs idle
        const int64_t a_rows = M <= 32 ? 32 : M <= 64 ? 64 : 128;
        const Cand cands[] = {{kCfgN128x2, 128, 128, 1, M <= 128 ? (a_rows + 128) * 128 / 48 : 512, 3000},
                              {kCfgN64x2, 128, 64, 1, M <= 128 ? (a_rows + 64) * 128 / 48 : 100000, 3000},
                              {kCfgN32x2, 128, 32, 1, M <= 128 ? (a_rows + 32) * 128 / 48 : 100000, 3000},
                              {kCfg2CtaN128x2, 256, 128, 2, M <= kDecodeMaxM ? 500 : 384, 3000},
                              {kCfg2CtaN192Tma, 256, 192, 2, 436, 4500},
                              {kCfg2CtaN256Tma, 256, 256, 2, 556, 6000}};
        int64_t best = INT64_MAX;
        if (M <= 128) {
            // One row-block of tokens.  Measured: the K loop costs ~370 clk per K-block whatever the tile width (an MMA below N = 128
            // takes a fixed ~93 clk) unless the landed bytes cost more, so the levers are CTAs and K-blocks per CTA: every
            // (tile width, K-split) pair is priced as waves x K-blocks per CTA x cycles + tail (+ the exchange of a split).
            const int64_t rw = (M + 31) / 32;
            struct Small { int id, tile_n, stages; };
            const Small smalls[] = {{kCfgN128x2, 128, M <= 32 ? 10 : M <= 64 ? 8 : 6}, {kCfgN64x2, 64, M <= 32 ? 16 : M <= 64 ? 12 : 8},
                                    {kCfgN32x2, 32, M <= 32 ? 24 : M <= 64 ? 16 : 10}};
            for (const Small& c : smalls)
                for (int sp = 1; sp <= 4; sp *= 2) {      // clusters of 8 schedule badly (measured 3x slower): opt-in only
                    const int64_t stage_bytes = (a_rows + c.tile_n) * 128;
                    // a split pays its cluster launch and exchange (~3 us) only on a long K loop (measured: K = 11008 -3 us, K = 4096 +3 us)
                    if (sp > 1 && ((sp - 1) * rw * c.tile_n * 128 > c.stages * stage_bytes || (nkb - kOutlierKBlocks) < 64)) continue;
                    const int64_t ctas = ((N + c.tile_n - 1) / c.tile_n) * sp;
                    const int64_t waves = (ctas + usable_sms(opts) - 1) / usable_sms(opts);
                    const int64_t kb_cyc = std::max<int64_t>(370, stage_bytes / 48);
                    const int64_t est = waves * ((nkb - kOutlierKBlocks + sp - 1) / sp + kOutlierKBlocks) * kb_cyc + 3000 + (sp > 1 ? 6000 : 0);
                    if (est < best) {
                        best = est;
                        cfg = c.id;
                        ksplit = sp;
                    }
                }
        }
        for (const Cand& c : cands) {
            if (M <= 128) break;
            // decode batches (weights streamed from HBM once): every configuration is bound by the rate at which TMA lands
            // operand bytes in an SM (85-100 GB/s per SM = ~48 B/clk at the 1.97 GHz these short kernels run at,
            // tools/microbench_ingest.cu), so cycles per K-block are bytes per K-block and CTA / 48; the wide TMA-store tiles'
            // lower bytes/MAC is eaten by their 1.3-wave schedule (profiles/r2_decode_ab.txt)
            if (M <= kDecodeMaxM && (c.id == kCfg2CtaN192Tma || c.id == kCfg2CtaN256Tma)) continue;
            const int64_t tiles = ((M + c.tile_m - 1) / c.tile_m) * ((N + c.tile_n - 1) / c.tile_n);
            const int64_t workers = usable_sms(opts) / c.cta;
            const int64_t est = ((tiles + workers - 1) / workers) * c.kb_cycles * nkb + c.tail_cycles;
            if (est < best) {
                best = est;
                cfg = c.id;
            }
        }
        if (M > 128 && M <= kDecodeMaxM && usable_sms(opts) >= 2) {
            // one fat tile per CTA pair and wave (gemm_fat.cuh): longer exposed head (outlier product first) and tail
            const FatPlan pl = plan_fat(M, N, usable_sms(opts) / 2, false, 12);
            const int64_t est = static_cast<int64_t>(pl.waves) * (pl.stage_bytes / 48) * nkb + 7000;
            if (pl.stages >= 4 && est < best) {
                best = est;
                cfg = kCfg2CtaFat;
            }
            // (config 14, two pairs per tile each reducing half of K inside a cluster of 4, is not a candidate: measured on the
            // B200 the main loop runs at the same rate per MAC whatever the tile width -- the tensor pipe at these clocks, not the
            // operand bytes, sets it -- so the split only adds its exchange: 512x4096x11008 28.7 us against 22.6 us for id 5)
        }
    }
    if (cfg == kCfg2CtaFat || cfg == kCfg2CtaFatSplitK || cfg == kCfg2CtaFatEpi8) {
        if (M > kDecodeMaxM) return set_error(MIXQ_ERR_UNSUPPORTED, "gemm_dequant: the fat-tile kernel serves M <= 1024");
        return launch_fat(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, epi, opts, nullptr,
                          cfg == kCfg2CtaFatSplitK ? 2 : 1, cfg == kCfg2CtaFatEpi8 ? 8 : 12);
    }
    if (cfg == kCfg2CtaN256Decode || cfg == kCfg2CtaN256DecodeNoSplit) {
        if (!decode_ok)
            return set_error(MIXQ_ERR_WORKSPACE, "gemm_dequant: the decode kernel needs M <= 1024 and mixq_decode_workspace_size(M, N) bytes of workspace");
        if (!sk_flags_clean) {
            cudaError_t e = cudaMemsetAsync(sk_ws, 0, kDecodeFlagBytes, stream);
            if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(split-K flags)");
        }
        return launch_decode(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, sk_ws,
                             cfg == kCfg2CtaN256Decode, epi, opts);
    }
    if (cfg == kCfg2CtaN192Tma)  // 256x192 pair tiles: more tiles per wave for decode-sized M
        return launch_cfg<StreamKTraits<2, 5, 192>>(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, nullptr, 0, nullptr, epi, opts);
    if (cfg == kCfg2CtaN256Tma)  // the stream-K kernel with whole tiles: TMA-store epilogue, no scratch needed
        return launch_cfg<StreamKTraits<2, 4>>(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, nullptr, 0, nullptr, epi, opts);
    if (cfg == kCfg2CtaN256StreamK) {
        if (!sk_ok) return set_error(MIXQ_ERR_WORKSPACE, "gemm_dequant: stream-K needs mixq_gemm_workspace_size() bytes of workspace");
        if (!sk_flags_clean) {
            cudaError_t e = cudaMemsetAsync(sk_ws, 0, kStreamKFlagBytes, stream);
            if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(stream-K flags)");
        }
        return launch_cfg<StreamKTraits<2, 4>>(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, sk_ws, 1, nullptr, epi, opts);
    }
    switch (cfg) {
        case kCfgN128x2:
            if (M <= 32) return launch_cfg<GemmTraits<1, 128, 2, 10, 32>>(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, nullptr, 0, nullptr, epi, opts, ksplit);
            if (M <= 64) return launch_cfg<GemmTraits<1, 128, 2, 8, 64>>(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, nullptr, 0, nullptr, epi, opts, ksplit);
            return launch_cfg<GemmTraits<1, 128, 2, 6>>(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, nullptr, 0, nullptr, epi, opts, ksplit);
        case kCfgN256x1:
            return launch_cfg<GemmTraits<1, 256, 1, 4>>(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, nullptr, 0, nullptr, epi, opts);
        case kCfgN64x2:
            if (M <= 32) return launch_cfg<GemmTraits<1, 64, 2, 16, 32>>(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, nullptr, 0, nullptr, epi, opts, ksplit);
            if (M <= 64) return launch_cfg<GemmTraits<1, 64, 2, 12, 64>>(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, nullptr, 0, nullptr, epi, opts, ksplit);
            return launch_cfg<GemmTraits<1, 64, 2, 8>>(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, nullptr, 0, nullptr, epi, opts, ksplit);
        case kCfgN32x2:   // 128 x 32 tiles: four times the CTAs of id 1 stream weights (few output channels, one row-block of tokens)
            if (M <= 32) return launch_cfg<GemmTraits<1, 32, 2, 24, 32>>(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, nullptr, 0, nullptr, epi, opts, ksplit);
            if (M <= 64) return launch_cfg<GemmTraits<1, 32, 2, 16, 64>>(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, nullptr, 0, nullptr, epi, opts, ksplit);
            return launch_cfg<GemmTraits<1, 32, 2, 10>>(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, nullptr, 0, nullptr, epi, opts, ksplit);
        case kCfg2CtaN256x1:
            return launch_cfg<GemmTraits<2, 256, 1, 6>>(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, nullptr, 0, nullptr, epi, opts);
        case kCfg2CtaN128x2:
            return launch_cfg<GemmTraits<2, 128, 2, 8>>(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, nullptr, 0, nullptr, epi, opts);
        case kCfg2CtaN256Stash:
            return launch_cfg<StashTraits<2, 4>>(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, nullptr, 0, nullptr, epi, opts);
        case kCfgN256Stash:
            return launch_cfg<StashTraits<1, 3>>(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, M, N, K, stream, pdl, nullptr, 0, nullptr, epi, opts);
        default:
            return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant: unknown config id");
    }
}


int launch_gemm_dequant_gated(const void* A8, const void* scale_a, const void* fp_A, const void* W8_gate, const void* sb_gate,
                              const void* fpw_gate, const void* W8_up, const void* sb_up, const void* fpw_up, void* Out, int64_t M,
                              int64_t N, int64_t K, cudaStream_t stream, bool pdl, void* scratch, size_t scratch_bytes,
                              LaunchOpts opts) {
    if (M == 0 || N == 0) return MIXQ_OK;
    if (opts.sm_limit < 0) return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant_gated: negative sm_limit");
    if (!A8 || !scale_a || !W8_gate || !sb_gate || !W8_up || !sb_up || !Out) return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant_gated: null pointer");
    if ((fp_A == nullptr) != (fpw_gate == nullptr) || (fp_A == nullptr) != (fpw_up == nullptr))
        return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant_gated: fp_A and both fp_weight tensors must all be given or all be null");
    if (M < 0 || N < 0 || K <= 0 || M > INT32_MAX || N > INT32_MAX || K > INT32_MAX)
        return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant_gated: bad dimensions");
    if ((K & 15) != 0 || (N & 7) != 0) return set_error(MIXQ_ERR_UNSUPPORTED, "gemm_dequant_gated: K must be a multiple of 16 and N of 8");
    const uintptr_t al = reinterpret_cast<uintptr_t>(A8) | reinterpret_cast<uintptr_t>(W8_gate) | reinterpret_cast<uintptr_t>(W8_up) |
                         reinterpret_cast<uintptr_t>(Out) | reinterpret_cast<uintptr_t>(fp_A) | reinterpret_cast<uintptr_t>(fpw_gate) |
                         reinterpret_cast<uintptr_t>(fpw_up);
    if (al & 15) return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant_gated: tensors must be 16-byte aligned");
    if (!device_info().ok) return set_error(MIXQ_ERR_CUDA, "no usable sm_100 device");
    if (M <= kDecodeMaxM && usable_sms(opts) >= 2 && opts.cfg != kCfgGatedUnfused) {
        const FatGated g{W8_up, sb_up, fpw_up};
        // 8 epilogue warps: the gated epilogue (two accumulators and a SiLU per output) is issue-bound, a third warp per lane
        // quarter does not shorten it (45.5 vs 45.2 us at 512 x 11008 x 4096); id 13 asks for 12 all the same (A/B)
        return launch_fat(A8, W8_gate, scale_a, sb_gate, fp_A, fpw_gate, Out, M, N, K, stream, pdl, EpiArgs{nullptr, 0}, opts, &g, 1,
                          opts.cfg == kCfg2CtaFat ? 12 : 8);
    }
    // prefill-sized batches are tensor-bound and their tiles fill the machine: two GEMMs over the shared quantised A
    // (the gate's with the fused SiLU) and one elementwise pass, the same roundings in the same order
    const size_t need = static_cast<size_t>(M) * N * 2;
    if (!scratch || scratch_bytes < need || (reinterpret_cast<uintptr_t>(scratch) & 15))
        return set_error(MIXQ_ERR_WORKSPACE, "gemm_dequant_gated: M > 1024 needs M*N*2 bytes of 16-byte aligned scratch");
    LaunchOpts o2 = opts;
    o2.cfg = kCfgAuto;
    int rc = launch_gemm_dequant(A8, W8_gate, scale_a, sb_gate, fp_A, fpw_gate, Out, M, N, K, stream, pdl, nullptr, 0, false, nullptr,
                                 MIXQ_ACT_SILU, o2);
    if (rc) return rc;
    rc = launch_gemm_dequant(A8, W8_up, scale_a, sb_up, fp_A, fpw_up, scratch, M, N, K, stream, /*pdl=*/false, nullptr, 0, false, nullptr,
                             MIXQ_ACT_NONE, o2);
    if (rc) return rc;
    const size_t n2 = static_cast<size_t>(M) * N / 2;
    const int blocks = static_cast<int>(std::min<size_t>((n2 + 255) / 256, static_cast<size_t>(usable_sms(opts)) * 8));
    mixq_mul_inplace_kernel<<<blocks, 256, 0, stream>>>(static_cast<__half2*>(Out), static_cast<const __half2*>(scratch), n2);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "launch mul_inplace");
    count_launch();
    return MIXQ_OK;
}

// Tile of the one-kernel exchange: 256 x 256, or 256 x 128 for decode batches (twice the tiles: a 512 x 4096 result is 64 tiles
// for the 74 CTA pairs instead of 32, and a tile's main loop is half as long).  MIXQ_AR_DECODE_TILE_N=256 restores the wide tile.
using ArBulkT = StreamKTraits<2, 4>;
using ArDecodeT = StreamKTraits<2, 7, 128>;
bool ar_decode_tile(int64_t M) {
    static const int wide = [] {
        const char* e = std::getenv("MIXQ_AR_DECODE_TILE_N");
        return (e && std::atoi(e) == 256) ? 1 : 0;
    }();
    return M <= kDecodeMaxM && !wide;
}
size_t allreduce_staging_bytes(int64_t M, int64_t N, int world) {
    if (M <= 0 || N <= 0 || world <= 0) return 0;
    size_t need = 0;
    for (int bn : {ArBulkT::kBlockN, ArDecodeT::kBlockN}) {     // the larger of the two tilings: one allocation serves every batch
        const int64_t tiles = ((M + ArBulkT::kTileM - 1) / ArBulkT::kTileM) * ((N + bn - 1) / bn);
        const int64_t slots = (tiles + world - 1) / world;
        const size_t b = static_cast<size_t>(slots) * world * ArBulkT::kTileM * bn * 2;
        need = b > need ? b : need;
    }
    return need;
}
size_t allreduce_counter_bytes(int64_t M, int64_t N, int world) {
    (void)M; (void)N; (void)world;
    return kArCounterBytes;
}

int launch_gemm_dequant_allreduce(const void* A8, const void* W8, const void* scale_a, const void* scale_b,
                                  const void* fp_A, const void* fp_weight, int64_t M, int64_t N, int64_t K,
                                  const mixq_peer_group* pg, cudaStream_t stream, bool pdl, LaunchOpts opts) {
    if (!pg) return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant_allreduce: null peer group");
    if (pg->world < 1 || pg->world > MIXQ_MAX_RANKS || pg->rank < 0 || pg->rank >= pg->world)
        return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant_allreduce: bad world/rank");
    if (M == 0 || N == 0) return MIXQ_OK;
    if (!A8 || !W8 || !scale_a || !scale_b) return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant_allreduce: null pointer");
    if ((fp_A == nullptr) != (fp_weight == nullptr))
        return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant_allreduce: fp_A and fp_weight must both be given or both be null");
    if (M < 0 || N < 0 || K <= 0 || M > INT32_MAX || N > INT32_MAX || K > INT32_MAX)
        return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant_allreduce: bad dimensions");
    if ((K & 15) != 0 || (N & 7) != 0)
        return set_error(MIXQ_ERR_UNSUPPORTED, "gemm_dequant_allreduce: K must be a multiple of 16 and N of 8");
    const uintptr_t al = reinterpret_cast<uintptr_t>(A8) | reinterpret_cast<uintptr_t>(W8) | reinterpret_cast<uintptr_t>(fp_A) |
                         reinterpret_cast<uintptr_t>(fp_weight);
    if (al & 15) return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant_allreduce: tensors must be 16-byte aligned");
    for (int i = 0; i < pg->world; ++i)
        if ((reinterpret_cast<uintptr_t>(pg->out[i]) | reinterpret_cast<uintptr_t>(pg->staging[i]) |
             reinterpret_cast<uintptr_t>(pg->counters[i])) & 15)
            return set_error(MIXQ_ERR_BAD_ARG, "gemm_dequant_allreduce: peer buffers must be 16-byte aligned");
    if (!device_info().ok) return set_error(MIXQ_ERR_CUDA, "no usable sm_100 device");
    if (ar_decode_tile(M))
        return launch_cfg<ArDecodeT>(A8, W8, scale_a, scale_b, fp_A, fp_weight, pg->out[pg->rank], M, N, K, stream, pdl, nullptr, 0, pg,
                                     EpiArgs{nullptr, 0}, opts);
    return launch_cfg<ArBulkT>(A8, W8, scale_a, scale_b, fp_A, fp_weight, pg->out[pg->rank], M, N, K, stream, pdl, nullptr, 0, pg,
                               EpiArgs{nullptr, 0}, opts);
}

}  // namespace mixq
