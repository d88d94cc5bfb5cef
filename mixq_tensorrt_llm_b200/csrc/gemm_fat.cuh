// gemm_fat.cuh -- stage 2 for decode batches (128 < M <= 1024): ONE fat tile per CTA pair.  Included by
// gemm_i8_tcgen05.cu inside namespace mixq::{anonymous}, after the shared helpers.
//
// Why (measured, tools/microbench_ingest.cu + profiles/r2_decode_ab.txt): TMA lands operand bytes in an SM's shared memory
// at 85-100 GB/s whatever the number of active SMs and whether or not the boxes are multicast, and at M = 512 every
// schedule is bound by that rate -- the tensor pipe would take 0.28 us for a 256x256x128 K-block whose 32 KB per CTA need
// 0.33-0.39 us to land.  So the only lever is MACs per landed byte, i.e. tile area over tile perimeter, under two limits:
// 512 TMEM columns per CTA and one wave (a second, partial wave costs more than a better tile saves).
//     tile = 256 rows x Nt columns per CTA pair, Nt = N / (pairs per row-block) rounded up to 16, at most 336
//     (M = 512, N = 12288: 37 pairs per row-block, Nt = 336: 37 KB per K-block and CTA for 5.5 M MACs = 149 MAC/B,
//      against 85 MAC/B of the 256x128 tiles, and every pair busy for the whole kernel).
// Nt > 256 is two tcgen05.mma per K step (N = 256 and N = Nt - 256) that share the A block in shared memory; Nt and the
// ring depth are run-time values (instruction descriptors and TMA boxes are data).  One tile per pair means nothing
// overlaps the epilogue anyway, so the int32 accumulator is single-buffered: columns [0, Nt).  The fp32 outlier product is
// computed FIRST into the same columns, drained by the epilogue warps, rounded to fp16 (the reference's store of the
// cuBLAS result, TsinghuaMixQPlugin.cpp:521) and parked two-per-column in TMEM columns [Nt, Nt + Nt/2) with tcgen05.st
// (336 + 168 = 504 <= 512) while the ring already fills with the first INT8 K-blocks; the final epilogue reads both from
// TMEM.  No shared-memory stash of the outlier product, no scratch in global memory; the fp16 result is staged per warp as
// 32 x 32 SWIZZLE_64B tiles (32 KB in all, double-buffered) and leaves through TMA stores.
// Arithmetic and rounding points are those of the other stage-2 kernels (reference
// kernel/symmetric/gemm/kernel/gemm_dequant.h:224-292, epilogue/thread/linear_combination_dequant.h:152-157).
//
// Gated mode (`gated` != 0): the gate and up projections of one MLP in ONE launch.  The two tcgen05.mma
// of a K step then read the gate rows and the up rows of the SAME Ng = Nt/2 output channels (two tensor maps over the two
// weight tensors as the checkpoint holds them: no interleaved repacking), the accumulators sit side by side in TMEM
// (gate [0, Ng), up [Ng, 2 Ng), parked outlier products behind them) and the epilogue writes
//     Out[m, n] = fp16( fp16(silu(gate[m, n])) * fp16(up[m, n]) )            [M, N], N = channels of ONE projection
// i.e. MixLlamaMLP.forward's  gate_proj.forward_without_preconditionFusedSilu(x) *= up_proj(x)
// (MixQ/src/mixquant/modules/fused/mlp.py:57-70, modules/linear.py:288-373): SiLU in fp32 before the rounding of the gate
// (epilogue/thread/linear_combination_dequant.h:167-272), the up projection rounded to fp16, an fp16 multiply.  One
// quantised A serves both (the reference's MixGemmCache), the [M, N] intermediate of each projection never exists.

constexpr int kFatMaxN = 336;          // Nt + Nt / 2 <= 512 TMEM columns, Nt % 16 == 0
constexpr int kFatMaxStages = 8;
constexpr int kFatSbFloats = 352;
// per epilogue warp two 32-row x 32-column fp16 tiles (double-buffered TMA-store staging)
constexpr int fat_out_stage_bytes(int epi_warps) { return epi_warps * 2 * 2048; }
constexpr int kFatMaxPartPairs = 6;                // split-K: chunk pairs (32 columns) per epilogue warp, (336 / 16 / 2 + 1) / 2 rounded up
// + scale_b / bias staging (2 x 352 floats, one tile at a time), barriers, TMEM pointer
constexpr int fat_fixed_bytes(int epi_warps) { return 3072 + 512 + fat_out_stage_bytes(epi_warps); }

__device__ __forceinline__ uint32_t fat_idesc_i8(int n) { return ptx::make_idesc_i8(256, n); }
__device__ __forceinline__ uint32_t fat_idesc_f16(int n) { return ptx::make_idesc_f16(256, n); }

// Chunk (16 accumulator columns) range of part p when the `parts` epilogue warps that share a TMEM lane quarter split a tile's
// n chunks: even-sized shares, so every part starts on a 32-column boundary (the TMA-store tiles are 32 columns wide).
__device__ __forceinline__ void fat_chunk_range(int n, int parts, int p, int& b, int& e) {
    const int per = ((n + parts - 1) / parts + 1) & ~1;
    b = min(n, p * per);
    e = min(n, b + per);
}

// EW = epilogue warps per CTA (8 or 12: two or three per TMEM lane quarter).  The exposed epilogue of a one-wave tile is a
// dependent chain per warp (tcgen05.ld -> dequantise -> staging tile -> TMA store), not a bandwidth: more warps shorten it.
// The split-K mode (ksplit == 2) is written for EW = 8.
template <int EW>
__global__ void __launch_bounds__((kEpilogueWarp0 + EW) * 32, 1)
mixq_gemm_dequant_fat_kernel(const __grid_constant__ CUtensorMap tm_a8, const __grid_constant__ CUtensorMap tm_w1,
                             const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_fa,
                             const __grid_constant__ CUtensorMap tm_fw1, const __grid_constant__ CUtensorMap tm_fw2,
                             const __grid_constant__ CUtensorMap tm_out, const __half* __restrict__ scale_a, const __half* __restrict__ scale_b,
                             const __half* __restrict__ scale_b2, __half* __restrict__ Out, int M, int N, int K, int has_outlier,
                             int m_tiles, int n_tiles, int Nt, int stages, int gated, int ksplit, EpiArgs epi) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int N1 = gated ? Nt / 2 : (Nt > 256 ? 256 : Nt);   // columns of the first / second tcgen05.mma of a K step
    const int N2 = Nt - N1;
    const int tile_cols = gated ? N1 : Nt;         // output columns per tile
    const int w1_bytes = (N1 / 2) * kBlockKBytes;  // W rows each CTA stages for MMA 1 (half of them: cta_group::2)
    const int stage_bytes = kBlockM * kBlockKBytes + (Nt / 2) * kBlockKBytes;
    uint8_t* ring = smem;
    uint8_t* out_stage = ring + static_cast<size_t>(stages) * stage_bytes;                       // 2 KB tiles, 1024-aligned
    float* sb_s = reinterpret_cast<float*>(out_stage + fat_out_stage_bytes(EW));                 // [352], rewritten per tile
    float* bias_sm = sb_s + kFatSbFloats;                                                         // [352]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_sm + kFatSbFloats);
    uint64_t* empty_bar = full_bar + kFatMaxStages;
    uint64_t* f_full_bar = empty_bar + kFatMaxStages;    // outlier product complete in TMEM columns [0, Nt)
    uint64_t* f_drained_bar = f_full_bar + 1;            // ... rounded to fp16 and parked in columns [Nt, Nt + Nt/2)
    uint64_t* tmem_full_bar = f_drained_bar + 1;         // int32 accumulators complete
    uint64_t* tmem_empty_bar = tmem_full_bar + 1;        // epilogue has read them
    uint64_t* ring_free_bar = tmem_empty_bar + 1;        // split-K, in the sending CTA: the finishing CTA's ring may be overwritten
    uint64_t* part_bar = ring_free_bar + 1;              // split-K, in the finishing CTA: [8 warps][kFatMaxPartPairs] partial sums landed
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(part_bar + 8 * kFatMaxPartPairs);

    constexpr int kEpiThreads = EW * 32;
    constexpr int kParts = EW / 4;
    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // Split-K (ksplit == 2): a cluster of 4 = two CTA pairs on one tile.  Pair 0 (cluster ranks 0, 1) reduces the outlier
    // product and the first K-blocks and FINISHES the tile; pair 1 (ranks 2, 3) reduces the remaining K-blocks and SENDS its
    // int32 partial sums into the finishing CTAs' shared memory (the ring, idle by then) with st.async, 16 bytes a lane,
    // counted on mbarriers there.  Integer addition is associative: the result is bit-identical to the unsplit kernel.
    const uint32_t crank = ptx::cluster_ctarank();
    const uint32_t cta_rank = crank & 1u;              // rank inside the CTA pair
    const uint32_t kpart = crank >> 1;                 // which K range this pair reduces
    const uint32_t leader_rank = crank & ~1u;
    const uint16_t pair_mask = static_cast<uint16_t>(3u << (kpart * 2));
    const bool is_leader = cta_rank == 0;
    const bool sender = EW == 8 && kpart != 0;            // (split-K exists in the EW = 8 instantiation only)
    const int cluster_ctas = 2 * ksplit;
    const int group_id = blockIdx.x / cluster_ctas;
    const int num_groups = gridDim.x / cluster_ctas;
    if (sender) has_outlier = 0;                        // the outlier product belongs to the finishing pair
    // chunk (16 accumulator columns) ranges of the two epilogue warps that share a TMEM lane quarter
    const int n_chunks_all = Nt >> 4;
    const int c_split_all = min(n_chunks_all, ((n_chunks_all + 1) / 2 + 1) & ~1);   // split-K (EW = 8): even, both halves start on a 32-column boundary
    if (threadIdx.x == 0) trace_stamp(0);

    if (warp_idx == 0 && ptx::elect_one()) {
        ptx::prefetch_tensormap(&tm_a8);
        ptx::prefetch_tensormap(&tm_w1);
        ptx::prefetch_tensormap(&tm_out);
        if (N2) ptx::prefetch_tensormap(&tm_w2);
        if (has_outlier) {
            ptx::prefetch_tensormap(&tm_fa);
            ptx::prefetch_tensormap(&tm_fw1);
            if (N2) ptx::prefetch_tensormap(&tm_fw2);
        }
    }
    if (warp_idx == 1 && ptx::elect_one()) {
        for (int i = 0; i < stages; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        ptx::mbar_init(f_full_bar, 1);
        ptx::mbar_init(tmem_full_bar, 1);
        ptx::mbar_init(f_drained_bar, 2 * kEpiThreads / 32);
        ptx::mbar_init(tmem_empty_bar, 2 * kEpiThreads / 32);
        ptx::mbar_init(ring_free_bar, 1);
        for (int i = 0; i < 8 * kFatMaxPartPairs; ++i) ptx::mbar_init(&part_bar[i], 1);
        if (ksplit == 2 && !sender) {
            // arm the landing barriers: warp w's pair p of chunks brings 2 KB per chunk (32 lanes x 16 columns x 4 bytes)
            for (int w = 0; w < 8; ++w) {
                const int cb = (w >> 2) == 0 ? 0 : c_split_all, ce = (w >> 2) == 0 ? c_split_all : n_chunks_all;
                for (int c = cb, pp = 0; c < ce; c += 2, ++pp)
                    ptx::mbar_arrive_expect_tx(&part_bar[w * kFatMaxPartPairs + pp], static_cast<uint32_t>(min(2, ce - c)) * 2048u);
            }
        }
        ptx::fence_barrier_init();
    }
    if (warp_idx == 2) {
        ptx::tmem_alloc_2cta(tmem_ptr_s, 512);
        ptx::tmem_relinquish_2cta();
    }
    ptx::tc_fence_before_sync();
    ptx::cluster_sync();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (threadIdx.x == 0) trace_stamp(1);
    ptx::pdl_wait_prior_grid();
    ptx::pdl_launch_dependents();   // dependents may be scheduled as our CTAs retire; they wait for this grid's completion themselves

    const int num_tiles = m_tiles * n_tiles;
    const int num_kb_all = (K + kBlockKBytes - 1) / kBlockKBytes;
    // the finishing pair also does the outlier K-blocks and waits for their drain: it gets ~4 K-blocks less
    const int kb_cut = ksplit == 2 ? max(1, min(num_kb_all - 1, (num_kb_all - 4) / 2)) : num_kb_all;
    const int kb_first = sender ? kb_cut : 0;
    const int num_kb = sender ? num_kb_all - kb_cut : kb_cut;      // K-blocks of THIS pair
    const int n_f = has_outlier ? kOutlierKBlocks : 0;

    if (warp_idx == 0) {
        if (ptx::elect_one()) {
            // ===================== TMA producer (every CTA) =====================
            int stage = 0;
            uint32_t phase = 0;
            auto load_block = [&](const CUtensorMap* ma, const CUtensorMap* mb1, const CUtensorMap* mb2, int k0, int m0, int n1, int n2) {
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                if (is_leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(stage_bytes) * 2);
                uint8_t* sA = ring + static_cast<size_t>(stage) * stage_bytes;
                uint8_t* sB1 = sA + kBlockM * kBlockKBytes;
                ptx::tma_load_2d_2cta(sA, ma, &full_bar[stage], k0, m0, ptx::kEvictNormal);
                ptx::tma_load_2d_2cta(sB1, mb1, &full_bar[stage], k0, n1, ptx::kEvictNormal);
                if (N2) ptx::tma_load_2d_2cta(sB1 + w1_bytes, mb2, &full_bar[stage], k0, n2, ptx::kEvictNormal);
                if (++stage == stages) {
                    stage = 0;
                    phase ^= 1;
                }
            };
            trace_stamp(2);
            for (int tile = group_id; tile < num_tiles; tile += num_groups) {
                const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;   // the row-blocks that share a W tile run side by side
                const int m0 = m_blk * 256 + static_cast<int>(cta_rank) * kBlockM;
                const int n1 = n_blk * tile_cols + static_cast<int>(cta_rank) * (N1 / 2);
                const int n2 = n_blk * tile_cols + (gated ? 0 : N1) + static_cast<int>(cta_rank) * (N2 / 2);   // gated: the same channels of the up projection
                for (int it = 0; it < n_f; ++it) load_block(&tm_fa, &tm_fw1, &tm_fw2, it * (kBlockKBytes / 2), m0, n1, n2);
                for (int kb = 0; kb < num_kb; ++kb) load_block(&tm_a8, &tm_w1, &tm_w2, (kb_first + kb) * kBlockKBytes, m0, n1, n2);
            }
        }
        __syncwarp();
    } else if (warp_idx == 1) {
        if (is_leader && ptx::elect_one()) {
            // ===================== MMA issuer (leader CTA only) =====================
            const uint32_t idesc_i8_1 = fat_idesc_i8(N1), idesc_i8_2 = fat_idesc_i8(N2 ? N2 : 16);
            const uint32_t idesc_f16_1 = fat_idesc_f16(N1), idesc_f16_2 = fat_idesc_f16(N2 ? N2 : 16);
            constexpr uint32_t kDescStep = kUmmaKBytes >> 4;
            const uint64_t desc_a0 = ptx::make_smem_desc_sw128(ptx::smem_u32(ring));
            const uint64_t desc_b0 = ptx::make_smem_desc_sw128(ptx::smem_u32(ring) + kBlockM * kBlockKBytes);
            const uint64_t desc_c0 = ptx::make_smem_desc_sw128(ptx::smem_u32(ring) + kBlockM * kBlockKBytes + w1_bytes);
            const uint32_t stage_step = static_cast<uint32_t>(stage_bytes) >> 4;
            const uint32_t tmem_2 = tmem_base + static_cast<uint32_t>(N1);
            int stage = 0;
            uint32_t phase = 0;
            bool ready = false;
            auto issue_block = [&](auto kind_tag, bool first) {
                if (!ready) ptx::mbar_wait(&full_bar[stage], phase);
                const int nstage = (stage + 1 == stages) ? 0 : stage + 1;
                const uint32_t nphase = (stage + 1 == stages) ? phase ^ 1 : phase;
                ready = ptx::mbar_try_wait(&full_bar[nstage], nphase);
                ptx::tc_fence_after_sync();
                const uint64_t da = desc_a0 + static_cast<uint64_t>(stage) * stage_step;
                const uint64_t db = desc_b0 + static_cast<uint64_t>(stage) * stage_step;
                const uint64_t dc = desc_c0 + static_cast<uint64_t>(stage) * stage_step;
#pragma unroll
                for (int k = 0; k < kBlockKBytes / kUmmaKBytes; ++k) {
                    const uint32_t acc = (first && k == 0) ? 0u : 1u;
                    if constexpr (decltype(kind_tag)::value == 0) {
                        ptx::umma_f16_2cta(tmem_base, da + k * kDescStep, db + k * kDescStep, idesc_f16_1, acc);
                        if (N2) ptx::umma_f16_2cta(tmem_2, da + k * kDescStep, dc + k * kDescStep, idesc_f16_2, acc);
                    } else {
                        ptx::umma_i8_2cta(tmem_base, da + k * kDescStep, db + k * kDescStep, idesc_i8_1, acc);
                        if (N2) ptx::umma_i8_2cta(tmem_2, da + k * kDescStep, dc + k * kDescStep, idesc_i8_2, acc);
                    }
                }
                ptx::umma_commit_2cta_mask(&empty_bar[stage], pair_mask);
                stage = nstage;
                phase = nphase;
            };
            int lt = 0;
            for (int tile = group_id; tile < num_tiles; tile += num_groups, ++lt) {
                if (lt > 0) {   // the previous tile's accumulators (and its parked outlier product) have been read
                    ptx::mbar_wait(tmem_empty_bar, (lt - 1) & 1);
                    ptx::tc_fence_after_sync();
                }
                if (has_outlier) {
                    for (int it = 0; it < kOutlierKBlocks; ++it) issue_block(std::integral_constant<int, 0>{}, it == 0);
                    ptx::umma_commit_2cta_mask(f_full_bar, pair_mask);
                    ptx::mbar_wait(f_drained_bar, lt & 1);   // columns [0, Nt) are free again (the ring keeps filling meanwhile)
                    ptx::tc_fence_after_sync();
                }
                for (int kb = 0; kb < num_kb; ++kb) issue_block(std::integral_constant<int, 1>{}, kb == 0);
                ptx::umma_commit_2cta_mask(tmem_full_bar, pair_mask);
                if (lt == 0) trace_stamp(3);
            }
            trace_stamp(4);
        }
        __syncwarp();
    } else if (warp_idx >= kEpilogueWarp0) {
        // ===================== epilogue (EW warps; every CTA: its own 128 accumulator rows) =====================
        // the kParts warps of a TMEM lane quarter split the tile's 16-column chunks between them
        const int quarter = warp_idx & 3;
        const int part = (warp_idx - kEpilogueWarp0) >> 2;
        const int et = threadIdx.x - kEpilogueWarp0 * 32;   // 0 .. kEpiThreads - 1
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        const int n_chunks = n_chunks_all;
        const int c_split = c_split_all;
        int c_begin, c_end;
        fat_chunk_range(n_chunks, kParts, part, c_begin, c_end);
        const uint32_t t_acc = tmem_base + lane_base;                                   // int32 / fp32 accumulators
        const uint32_t t_out0 = tmem_base + lane_base + static_cast<uint32_t>(Nt);      // packed fp16 outlier product
        auto arrive = [&](uint64_t* bar) {
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(bar, leader_rank);
        };
        int lt = 0;
        for (int tile = group_id; tile < num_tiles; tile += num_groups, ++lt) {
            const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
            const int m0 = m_blk * 256 + static_cast<int>(cta_rank) * kBlockM;
            const int n0 = n_blk * tile_cols;
            const int gm = m0 + row;
            const bool row_ok = gm < M;
            const float sa_f = row_ok ? __half2float(scale_a[gm]) : 0.0f;
            // split-K landing zone in the FINISHING CTA's ring: [warp 0..7][chunk slot][16-byte vector 0..3][lane] (2 KB per chunk)
            const int ew = warp_idx - kEpilogueWarp0;
            const uint32_t zone_warp = ptx::smem_u32(ring) + static_cast<uint32_t>(ew * c_split) * 2048u + lane * 16u;

            if (EW == 8 && sender) {
                // ---- split-K sending pair: ship this CTA's int32 partial sums to the CTA of the finishing pair that holds the
                // same rows (cluster rank - 2), a chunk (32 lanes x 16 columns) at a time
                ptx::mbar_wait(tmem_full_bar, lt & 1);
                ptx::tc_fence_after_sync();
                ptx::mbar_wait(ring_free_bar, lt & 1);          // the finishing pair's tensor cores have read their whole ring
                if (et == 0) trace_stamp(lt == 0 ? 6 : 8);
                const uint32_t peer = crank - 2u;
                const uint32_t r_zone = ptx::mapa_shared(zone_warp, peer);
                const uint32_t r_bar0 = ptx::mapa_shared(ptx::smem_u32(&part_bar[ew * kFatMaxPartPairs]), peer);
                uint32_t va[16], vb[16];
                auto ship = [&](const uint32_t (&v)[16], int c) {
                    const uint32_t dst = r_zone + static_cast<uint32_t>(c - c_begin) * 2048u;
                    const uint32_t bar = r_bar0 + static_cast<uint32_t>((c - c_begin) >> 1) * 8u;
#pragma unroll
                    for (int g = 0; g < 4; ++g) ptx::st_async_v4(dst + g * 512u, v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3], bar);
                };
                if (c_begin < c_end) ptx::tmem_ld_32x16(t_acc + c_begin * 16, va);
                for (int c = c_begin; c < c_end; c += 2) {
                    ptx::tmem_ld_wait();
                    if (c + 1 < c_end) ptx::tmem_ld_32x16(t_acc + (c + 1) * 16, vb);
                    ship(va, c);
                    if (c + 1 < c_end) {
                        ptx::tmem_ld_wait();
                        if (c + 2 < c_end) ptx::tmem_ld_32x16(t_acc + (c + 2) * 16, va);
                        ship(vb, c + 1);
                    }
                }
                if (et == 0) trace_stamp(9);
                arrive(tmem_empty_bar);
                continue;
            }

            if (has_outlier) {
                // ---- outlier product: fp32 accumulators -> fp16 (the reference's rounding), two per TMEM column
                ptx::mbar_wait(f_full_bar, lt & 1);
                ptx::tc_fence_after_sync();
                if (et == 0 && lt == 0) trace_stamp(5);
                uint32_t fa[16], fb[16];
                auto park = [&](const uint32_t (&v)[16], int c) {
                    uint32_t pk[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const __half2 o = __floats2half2_rn(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
                        pk[q] = *reinterpret_cast<const uint32_t*>(&o);
                    }
                    ptx::tmem_st_32x8(t_out0 + c * 8, pk);
                };
                if (c_begin < c_end) ptx::tmem_ld_32x16(t_acc + c_begin * 16, fa);
                for (int c = c_begin; c < c_end; c += 2) {          // two chunks per trip: fixed register buffers
                    ptx::tmem_ld_wait();
                    if (c + 1 < c_end) ptx::tmem_ld_32x16(t_acc + (c + 1) * 16, fb);
                    park(fa, c);
                    if (c + 1 < c_end) {
                        ptx::tmem_ld_wait();
                        if (c + 2 < c_end) ptx::tmem_ld_32x16(t_acc + (c + 2) * 16, fa);
                        park(fb, c + 1);
                    }
                }
                ptx::tmem_st_wait();
                arrive(f_drained_bar);
                if (et == 0 && lt == 0) trace_stamp(12);
            }
            // the staging vectors are single-buffered: past this barrier every thread is done with the previous tile
            ptx::named_bar_sync(1, kEpiThreads);
            for (int j = et; j < Nt; j += kEpiThreads) {
                if (gated) {   // [0, Ng): gate scales, [Ng, 2 Ng): up scales of the same channels
                    const int jj = j < N1 ? j : j - N1;
                    sb_s[j] = (n0 + jj < N) ? __half2float((j < N1 ? scale_b : scale_b2)[n0 + jj]) : 0.0f;
                } else {
                    sb_s[j] = (n0 + j < N) ? __half2float(scale_b[n0 + j]) : 0.0f;
                    if (epi.bias) bias_sm[j] = (n0 + j < N) ? __half2float(epi.bias[n0 + j]) : 0.0f;
                }
            }
            ptx::named_bar_sync(1, kEpiThreads);

            // ---- dequantise the int32 accumulators
            ptx::mbar_wait(tmem_full_bar, lt & 1);
            ptx::tc_fence_after_sync();
            if (et == 0) trace_stamp(lt == 0 ? 6 : 8);
            if (EW == 8 && ksplit == 2 && et == 0) ptx::mbar_arrive_cluster(ring_free_bar, crank + 2u);   // the sending CTA may overwrite the ring now
            __half* out_row = Out + static_cast<size_t>(gm) * N + n0;
            // The result leaves through shared memory: row-per-thread 16-byte global stores touch 32 cache lines per instruction;
            // instead each warp writes 32-row x 32-column SWIZZLE_64B tiles (conflict free for row-per-thread writes) that go out
            // as one TMA store each (the tensor map clips rows >= M and columns >= N).  Two tiles per warp: the store of one is
            // in flight while the next is being filled.  An odd trailing 16-column chunk is stored directly.
            uint8_t* my_tiles = out_stage + (warp_idx - kEpilogueWarp0) * (2 * 2048);
            const uint32_t swz = static_cast<uint32_t>((lane >> 1) & 3);
            uint32_t ia[16], ib[16], oa[8], ob[8];
            if (!has_outlier) {
#pragma unroll
                for (int q = 0; q < 8; ++q) oa[q] = ob[q] = 0u;
            }
            // shared-state-space accesses spelled out: through the lambdas the compiler falls back to generic loads, and a
            // branch per element (activation / bias variants) serialises the chunk; the plugin's plain epilogue is branch free
            const uint32_t sb_addr = ptx::smem_u32(sb_s), bias_addr = ptx::smem_u32(bias_sm);
            const bool plain = epi.act == MIXQ_ACT_NONE && epi.bias == nullptr;
            auto finish = [&](const uint32_t (&vi)[16], const uint32_t (&vo)[8], int c, uint32_t (&packed)[8]) {
                float sbv[16];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float4 f = ptx::ld_shared_f4(sb_addr + (c * 16 + g * 4) * 4);
                    sbv[g * 4] = f.x; sbv[g * 4 + 1] = f.y; sbv[g * 4 + 2] = f.z; sbv[g * 4 + 3] = f.w;
                }
                float r[16];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float2 of = __half22float2(*reinterpret_cast<const __half2*>(&vo[q]));
                    r[2 * q] = __fmaf_rn(__int2float_rn(static_cast<int>(vi[2 * q])), __fmul_rn(sbv[2 * q], sa_f), of.x);
                    r[2 * q + 1] = __fmaf_rn(__int2float_rn(static_cast<int>(vi[2 * q + 1])), __fmul_rn(sbv[2 * q + 1], sa_f), of.y);
                }
                if (plain) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const __half2 h = __floats2half2_rn(r[2 * q], r[2 * q + 1]);
                        packed[q] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                } else {
                    float bv[16];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4 f = epi.bias ? ptx::ld_shared_f4(bias_addr + (c * 16 + g * 4) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                        bv[g * 4] = f.x; bv[g * 4 + 1] = f.y; bv[g * 4 + 2] = f.z; bv[g * 4 + 3] = f.w;
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const __half2 h = epi_finish(r[2 * q], r[2 * q + 1], bv + 2 * q, epi);
                        packed[q] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                }
            };
            auto store_direct = [&](const uint32_t (&packed)[8], int c) {
                if (row_ok) {
                    if (n0 + c * 16 + 8 <= N) ptx::st_global_v4(out_row + c * 16, packed[0], packed[1], packed[2], packed[3]);
                    if (n0 + c * 16 + 16 <= N) ptx::st_global_v4(out_row + c * 16 + 8, packed[4], packed[5], packed[6], packed[7]);
                }
            };
            const uint32_t tiles_addr = ptx::smem_u32(my_tiles) + lane * 64;
            auto store_staged = [&](const uint32_t (&packed)[8], uint32_t t, uint32_t v0) {   // vectors v0, v0 + 1 of this thread's 64-byte row
                ptx::st_shared_v4(t + ((v0 ^ swz) << 4), packed[0], packed[1], packed[2], packed[3]);
                ptx::st_shared_v4(t + (((v0 + 1) ^ swz) << 4), packed[4], packed[5], packed[6], packed[7]);
            };
            auto load = [&](uint32_t (&vi)[16], uint32_t (&vo)[8], int c) {
                ptx::tmem_ld_32x16(t_acc + c * 16, vi);
                if (has_outlier) ptx::tmem_ld_32x8(t_out0 + c * 8, vo);
            };
            if (!gated) {
            // split-K: the other pair's partial sums of chunk c wait in the landing zone once the chunk pair's barrier has fired
            auto add_partial = [&](uint32_t (&vi)[16], int c) {
                if (EW != 8 || ksplit != 2) return;
                const uint32_t src = zone_warp + static_cast<uint32_t>(c - c_begin) * 2048u;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint4 pv = ptx::ld_shared_u4(src + g * 512u);
                    vi[4 * g] += pv.x; vi[4 * g + 1] += pv.y; vi[4 * g + 2] += pv.z; vi[4 * g + 3] += pv.w;
                }
            };
            if (c_begin < c_end) load(ia, oa, c_begin);
            for (int c = c_begin; c < c_end; c += 2) {
                uint32_t pa[8], pb[8];
                const bool pair = c + 1 < c_end;
                const int pi = ((c - c_begin) >> 1) & 1;
                const uint32_t t = tiles_addr + pi * 2048;
                if (EW == 8 && ksplit == 2) ptx::mbar_wait(&part_bar[ew * kFatMaxPartPairs + ((c - c_begin) >> 1)], lt & 1);
                ptx::tmem_ld_wait();
                if (pair) load(ib, ob, c + 1);
                add_partial(ia, c);
                finish(ia, oa, c, pa);
                if (pair) {
                    // the TMA store that last used this staging tile (two pairs ago) must have finished READING it
                    if (lane == 0) ptx::tma_store_wait_read<1>();
                    __syncwarp();
                    store_staged(pa, t, 0u);
                    ptx::tmem_ld_wait();
                    if (c + 2 < c_end) load(ia, oa, c + 2);
                    add_partial(ib, c + 1);
                    finish(ib, ob, c + 1, pb);
                    store_staged(pb, t, 2u);
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        ptx::tma_store_2d(&tm_out, my_tiles + pi * 2048, n0 + c * 16, m0 + quarter * 32);
                        ptx::tma_store_commit();
                    }
                } else {
                    store_direct(pa, c);
                }
            }
            } else {
                // ---- gated: out = fp16(silu(gate)) * fp16(up) over the tile's Ng channels; gate accumulators in columns [0, Ng), up
                // in [Ng, 2 Ng), their parked outlier products in [Nt, Nt + Ng/2) and [Nt + Ng/2, Nt + Ng)
                const int g_chunks = N1 >> 4;
                int g_begin, g_end;
                fat_chunk_range(g_chunks, kParts, part, g_begin, g_end);
                const uint32_t t_up = t_acc + static_cast<uint32_t>(N1), t_out0_up = t_out0 + static_cast<uint32_t>(N1 >> 1);
                uint32_t iu[16], ou[8];
                if (!has_outlier) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) ou[q] = 0u;
                }
                auto gload = [&](int c) {
                    ptx::tmem_ld_32x16(t_acc + c * 16, ia);
                    ptx::tmem_ld_32x16(t_up + c * 16, iu);
                    if (has_outlier) {
                        ptx::tmem_ld_32x8(t_out0 + c * 8, oa);
                        ptx::tmem_ld_32x8(t_out0_up + c * 8, ou);
                    }
                };
                auto gfinish = [&](int c, uint32_t (&packed)[8]) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4 sg = ptx::ld_shared_f4(sb_addr + (c * 16 + g * 4) * 4);
                        const float4 su = ptx::ld_shared_f4(sb_addr + (N1 + c * 16 + g * 4) * 4);
                        const float sgv[4] = {sg.x, sg.y, sg.z, sg.w}, suv[4] = {su.x, su.y, su.z, su.w};
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            const int q = g * 2 + h2;
                            const float2 og = __half22float2(*reinterpret_cast<const __half2*>(&oa[q]));
                            const float2 ouf = __half22float2(*reinterpret_cast<const __half2*>(&ou[q]));
                            float g0 = __fmaf_rn(__int2float_rn(static_cast<int>(ia[2 * q])), __fmul_rn(sgv[2 * h2], sa_f), og.x);
                            float g1 = __fmaf_rn(__int2float_rn(static_cast<int>(ia[2 * q + 1])), __fmul_rn(sgv[2 * h2 + 1], sa_f), og.y);
                            const float u0 = __fmaf_rn(__int2float_rn(static_cast<int>(iu[2 * q])), __fmul_rn(suv[2 * h2], sa_f), ouf.x);
                            const float u1 = __fmaf_rn(__int2float_rn(static_cast<int>(iu[2 * q + 1])), __fmul_rn(suv[2 * h2 + 1], sa_f), ouf.y);
                            g0 = __fdividef(g0, 1.0f + __expf(-g0));     // SiLU in fp32 before the gate's rounding (epi_finish)
                            g1 = __fdividef(g1, 1.0f + __expf(-g1));
                            const __half2 o = __hmul2(__floats2half2_rn(g0, g1), __floats2half2_rn(u0, u1));
                            packed[q] = *reinterpret_cast<const uint32_t*>(&o);
                        }
                    }
                };
                for (int c = g_begin; c < g_end; c += 2) {
                    uint32_t pa[8], pb[8];
                    const bool pair = c + 1 < g_end;
                    const int pi = ((c - g_begin) >> 1) & 1;
                    const uint32_t t = tiles_addr + pi * 2048;
                    gload(c);
                    ptx::tmem_ld_wait();
                    gfinish(c, pa);
                    if (pair) {
                        gload(c + 1);
                        if (lane == 0) ptx::tma_store_wait_read<1>();
                        __syncwarp();
                        store_staged(pa, t, 0u);
                        ptx::tmem_ld_wait();
                        gfinish(c + 1, pb);
                        store_staged(pb, t, 2u);
                        ptx::fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            ptx::tma_store_2d(&tm_out, my_tiles + pi * 2048, n0 + c * 16, m0 + quarter * 32);
                            ptx::tma_store_commit();
                        }
                    } else {
                        store_direct(pa, c);
                    }
                }
            }
            if (et == 0) trace_stamp(9);
            arrive(tmem_empty_bar);
        }
        if (lane == 0) ptx::tma_store_wait_all<0>();   // outstanding output tiles fully written before the CTA retires
        if (et == 0) trace_stamp(10);
    }

    if (threadIdx.x == kEpilogueWarp0 * 32) trace_stamp(7);
    ptx::tc_fence_before_sync();
    ptx::cluster_sync();
    if (warp_idx == 2) ptx::tmem_dealloc_2cta(tmem_base, 512);
}
