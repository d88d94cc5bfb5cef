// gemm_decode.cuh -- the decode-batch (M <= 1024) variant of stage 2.  Included by gemm_i8_tcgen05.cu inside
// namespace mixq::{anonymous}, after the shared helpers (tile_coord, EpiArgs, the acquire/release wrappers).
//
// Same arithmetic as the other stage-2 kernels (int32-exact INT8 contraction, fp32 outlier product rounded to fp16,
// one FMA + one RN-even rounding; reference kernel/symmetric/gemm/kernel/gemm_dequant.h:224-292 +
// epilogue/thread/linear_combination_dequant.h:152-157 + the cuBLAS outlier GEMM TsinghuaMixQPlugin.cpp:122-161),
// re-shaped for what bounds a 512-token batch on this part (measured, profiles/r2_decode_*.txt):
//   * every weight byte comes from HBM exactly once, so a ring slot is in flight for ~1 us.  With 256x256 pair tiles
//     (the only tile whose shared-memory traffic fits under the tensor pipe) a slot is 32 KB and is consumed in
//     0.28 us: the ring has to be >= 5 slots deep.  The wide-tile prefill kernel spends 64 KB of shared memory on the
//     fp16 "stash" of the outlier product and gets 4 slots.  Here the stash lives in an L2-resident scratch instead
//     (thread-private vectors, written mid-loop, read back one 32-column chunk ahead in the final epilogue) and the
//     result leaves through 16-byte stores: 6 slots = 192 KB of loads in flight per CTA pair member.
//   * 96 tiles on 74 CTA pairs is 1.3 waves.  Two-phase schedule: the r = tiles mod pairs "remainder" tiles are cut
//     along K into one equal span per pair (stream-K; int32 partial sums meet in the workspace, integer addition is
//     associative so the result is bit-identical), and they run FIRST; the whole tiles run LAST.  The fix-up of a
//     split tile (wait for the peers, add their partial sums) then overlaps the main loop of the pair's next whole
//     tile and the kernel's exposed tail is a plain epilogue.  With fewer tiles than pairs everything is split.
// Roles: warp 0 TMA producer, warp 1 MMA issuer (leader CTA), warp 2 TMEM allocator, warps 4-11 epilogue
// (two per TMEM lane quarter; each thread owns one accumulator row x 128 columns).

template <int STAGES>
struct DecodeTraits {
    static constexpr int kCta = 2;
    static constexpr int kBlockN = 256;
    static constexpr int kLoadN = kBlockN / kCta;
    static constexpr int kTileM = kBlockM * kCta;
    static constexpr int kStages = STAGES;
    static constexpr int kABytes = kBlockM * kBlockKBytes;
    static constexpr int kBBytes = kLoadN * kBlockKBytes;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = 512;
    static constexpr int kNumBarriers = 2 * STAGES + 8;
    static constexpr size_t kSmemBytes = 1024 + static_cast<size_t>(STAGES) * kStageBytes + 4 * kBlockN * sizeof(float) +
                                         kNumBarriers * 8 + 16;
    static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

// Scratch of the decode kernel (carved from the caller's workspace, 128-byte aligned blocks):
//   flags    [workers][2]                      uint32, zero between launches (the finisher re-arms them)
//   slots    [workers][2][32][256]             uint4: raw int32 partial sums of a peer segment (128 KB per CTA)
//   out0     [tiles][2][16][256]               uint4: fp16 outlier product of a tile (64 KB per CTA)
constexpr size_t kDecodeFlagBytes = 4096;
constexpr size_t kDecodeSlotBytesPerCta = 32 * kStashEpiThreads * 16;   // 128 KB
constexpr size_t kDecodeOut0BytesPerCta = 16 * kStashEpiThreads * 16;   // 64 KB
constexpr int kDecodeMaxWorkers = 80;
constexpr int64_t kDecodeMaxM = 1024;

// Work list of one CTA pair: first its span of the split ("stream-K") region, then its whole tiles.
struct DecodeSched {
    int num_kb, num_groups, gid, num_tiles, gran;
    int total;       // units (tile, K-block) in the split region = sk_tiles * num_kb
    int u, u_end;    // this pair's span of the split region
    int dp_tile;     // next whole tile
    __device__ __forceinline__ int bound(int g) const {
        if (g >= num_groups) return total;
        return static_cast<int>(static_cast<long long>(total) * g / num_groups) / gran * gran;
    }
    __device__ void init(int sk_tiles, int nt, int nkb, int g, int ng, int granule) {
        num_kb = nkb; num_groups = ng; gid = g; num_tiles = nt; gran = granule;
        total = sk_tiles * nkb;
        u = bound(g);
        u_end = bound(g + 1);
        dp_tile = sk_tiles + g;
    }
    __device__ bool next(int& tile, int& kb0, int& kb1) {
        if (u < u_end) {
            tile = u / num_kb;
            kb0 = u - tile * num_kb;
            const int rem = u_end - u;
            kb1 = rem < num_kb - kb0 ? kb0 + rem : num_kb;
            u += kb1 - kb0;
            return true;
        }
        if (dp_tile < num_tiles) {
            tile = dp_tile;
            dp_tile += num_groups;
            kb0 = 0;
            kb1 = num_kb;
            return true;
        }
        return false;
    }
};

template <class T>
__global__ void __launch_bounds__(kStashThreads, 1)
mixq_gemm_dequant_decode_kernel(const __grid_constant__ CUtensorMap tm_a8, const __grid_constant__ CUtensorMap tm_w8,
                                const __grid_constant__ CUtensorMap tm_fa, const __grid_constant__ CUtensorMap tm_fw,
                                const __half* __restrict__ scale_a, const __half* __restrict__ scale_b,
                                __half* __restrict__ Out, int M, int N, int K, int has_outlier, int m_tiles, int n_tiles,
                                int group_m, int sk_tiles, int sk_gran, uint32_t* __restrict__ sk_flags,
                                uint4* __restrict__ sk_slots, uint4* __restrict__ out0_scratch, EpiArgs epi) {
    constexpr int BLOCK_N = T::kBlockN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    float* sb_s = reinterpret_cast<float*>(ring + static_cast<size_t>(T::kStages) * T::kStageBytes);
    float* bias_sm = sb_s + 2 * BLOCK_N;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sb_s + 4 * BLOCK_N);
    uint64_t* empty_bar = full_bar + T::kStages;
    uint64_t* tmem_full_bar = empty_bar + T::kStages;   // [2] int32 accumulators of buffer b complete
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;       // [2] epilogue has read buffer b's int32 accumulators
    uint64_t* f_full_bar = tmem_empty_bar + 2;          // [2] outlier accumulators parked in buffer b complete
    uint64_t* f_drained_bar = f_full_bar + 2;           // [2] outlier accumulators of buffer b moved to the scratch
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(f_drained_bar + 2);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = ptx::cluster_ctarank();
    const bool is_leader = cta_rank == 0;
    const int group_id = blockIdx.x >> 1;
    const int num_groups = gridDim.x >> 1;
    if (threadIdx.x == 0) trace_stamp(0);

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
            ptx::mbar_init(&tmem_empty_bar[i], 2 * kStashEpiThreads / 32);
            ptx::mbar_init(&f_drained_bar[i], 2 * kStashEpiThreads / 32);
        }
        ptx::fence_barrier_init();
    }
    if (warp_idx == 2) {
        ptx::tmem_alloc_2cta(tmem_ptr_s, T::kTmemCols);
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
    const int num_kb = (K + kBlockKBytes - 1) / kBlockKBytes;
    DecodeSched sched;
    sched.init(sk_tiles, num_tiles, num_kb, group_id, num_groups, sk_gran);
    int tile, kb0, kb1;

    if (warp_idx == 0) {
        if (ptx::elect_one()) {
            // ===================== TMA producer (every CTA) =====================
            int stage = 0;
            uint32_t phase = 0;
            auto load_block = [&](const CUtensorMap* ma, const CUtensorMap* mb, int k0, int m0, int n0) {
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                if (is_leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], T::kStageBytes * 2);
                uint8_t* sA = ring + static_cast<size_t>(stage) * T::kStageBytes;
                uint8_t* sB = sA + T::kABytes;
                ptx::tma_load_2d_2cta(sA, ma, &full_bar[stage], k0, m0, ptx::kEvictNormal);
                ptx::tma_load_2d_2cta(sB, mb, &full_bar[stage], k0, n0, ptx::kEvictNormal);
                if (++stage == T::kStages) {
                    stage = 0;
                    phase ^= 1;
                }
            };
            trace_stamp(2);
            int s = 0;
            while (sched.next(tile, kb0, kb1)) {
                const TileCoord tc = tile_coord(tile, m_tiles, n_tiles, group_m);
                const int m0 = tc.m_blk * T::kTileM + static_cast<int>(cta_rank) * kBlockM;
                const int n0 = tc.n_blk * BLOCK_N + static_cast<int>(cta_rank) * T::kLoadN;
                // the outlier K-blocks belong to the segment that holds the LAST K-block of the tile: first thing in a
                // pair's first segment (both accumulator buffers are free), half-way through the segment otherwise
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
                    if constexpr (decltype(kind_tag)::value == 0)
                        ptx::umma_f16_2cta(tmem_d, da + k * kDescStep, db + k * kDescStep, idesc_f16, acc);
                    else
                        ptx::umma_i8_2cta(tmem_d, da + k * kDescStep, db + k * kDescStep, idesc_i8, acc);
                }
                ptx::umma_commit_2cta(&empty_bar[stage]);
                stage = nstage;
                phase = nphase;
            };
            // completions this thread has caused / must have seen, per accumulator buffer
            uint32_t n_int[2] = {0, 0};     // segments that used buffer x for int32 accumulators (= tmem_full commits)
            uint32_t n_f[2] = {0, 0};       // outlier accumulators parked in buffer x (= f_full commits)
            bool f_pending[2] = {false, false};
            int s = 0;
            while (sched.next(tile, kb0, kb1)) {
                const int b = s & 1;
                const bool has_f = has_outlier && kb1 == num_kb;
                const uint32_t tmem_i = tmem_base + b * BLOCK_N;
                const uint32_t tmem_f = tmem_base + (b ^ 1) * BLOCK_N;
                // buffer b must be free: its previous int32 tenant read by the epilogue, a parked outlier accumulator drained
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
                    ptx::umma_commit_2cta(&f_full_bar[b ^ 1]);
                    ++n_f[b ^ 1];
                    f_pending[b ^ 1] = true;
                }
                for (int kb = split; kb < kb1; ++kb) issue_block(std::integral_constant<int, 1>{}, tmem_i, kb == kb0);
                ptx::umma_commit_2cta(&tmem_full_bar[b]);
                if (s == 0) trace_stamp(3);
                ++n_int[b];
                ++s;
            }
            trace_stamp(4);
        }
        __syncwarp();
    } else if (warp_idx >= kEpilogueWarp0) {
        // ===================== epilogue (8 warps; every CTA: its own 128 accumulator rows) =====================
        const int quarter = warp_idx & 3;
        const int half = (warp_idx - kEpilogueWarp0) >> 2;
        const int et = threadIdx.x - kEpilogueWarp0 * 32;   // 0..255
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        constexpr int kCols = BLOCK_N / 2;                  // 128 columns per thread
        constexpr int kChunks = kCols / 32;                 // 4
        const int col0 = half * kCols;
        auto arrive = [&](uint64_t* bar) {
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(bar, 0);
        };
        // thread-major scratch: vector v of thread et lives at base[v * 256 + et] (coalesced 512 B per warp access)
        auto slot_of = [&](int w) { return sk_slots + (static_cast<size_t>(w) * 2 + cta_rank) * (32 * kStashEpiThreads) + et; };
        auto out0_of = [&](int t) { return out0_scratch + (static_cast<size_t>(t) * 2 + cta_rank) * (16 * kStashEpiThreads) + et; };
        auto flag_of = [&](int w) { return sk_flags + static_cast<size_t>(w) * 2 + cta_rank; };
        uint32_t n_int[2] = {0, 0}, n_f[2] = {0, 0};
        int s = 0;
        while (sched.next(tile, kb0, kb1)) {
            const int b = s & 1;
            const bool finisher = kb0 == 0;
            const bool has_f = has_outlier && kb1 == num_kb;   // this segment computed the outlier product
            const TileCoord tc = tile_coord(tile, m_tiles, n_tiles, group_m);
            const int m0 = tc.m_blk * T::kTileM + static_cast<int>(cta_rank) * kBlockM;
            const int n0 = tc.n_blk * BLOCK_N;
            const int gm = m0 + row;
            float* sbt = sb_s + b * BLOCK_N;
            float* bt = bias_sm + b * BLOCK_N;
            float sa_f = 0.0f;
            if (finisher) {
                sbt[et] = (n0 + et < N) ? __half2float(scale_b[n0 + et]) : 0.0f;
                bt[et] = (epi.bias && n0 + et < N) ? __half2float(epi.bias[n0 + et]) : 0.0f;
                sa_f = gm < M ? __half2float(scale_a[gm]) : 0.0f;
            }
            ptx::named_bar_sync(1, kStashEpiThreads);
            uint4* o0 = out0_of(tile);

            if (has_f) {
                // ---- drain the outlier accumulator (parked in the other buffer) to the scratch as fp16
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
                        o0[(c * 4 + v) * kStashEpiThreads] = make_uint4(h[0], h[1], h[2], h[3]);
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
                // ---- PEER: dump the raw int32 partial sums, then raise this pair's flag (it also covers the outlier
                // product written above when this is the tile's tail segment)
                uint4* slot = slot_of(group_id);
                uint32_t v[2][32];
                ptx::tmem_ld_32x32(t_i, v[0]);
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    ptx::tmem_ld_wait();
                    if (c + 1 < kChunks) ptx::tmem_ld_32x32(t_i + (c + 1) * 32, v[(c + 1) & 1]);
                    const uint32_t* vc = v[c & 1];
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        slot[(c * 8 + g) * kStashEpiThreads] = make_uint4(vc[g * 4], vc[g * 4 + 1], vc[g * 4 + 2], vc[g * 4 + 3]);
                }
                arrive(&tmem_empty_bar[b]);
                ptx::named_bar_sync(2, kStashEpiThreads);  // every thread's scratch stores precede the flag
                if (et == 0) {
                    __threadfence();
                    st_release_gpu(flag_of(group_id), 1u);
                }
            } else {
                // ---- FINISHER: [collect the peers' partial sums,] dequantise, store
                int n_peers = 0;
                if (kb1 < num_kb) {
                    // the rest of this tile [kb1, num_kb) was computed by the following pairs' first segments
                    const int tile_end = (tile + 1) * num_kb;
                    int p = group_id + 1;
                    while (p < num_groups && sched.bound(p) < tile_end) {
                        ++n_peers;
                        ++p;
                    }
                    if (et == 0) {
                        for (int q = 1; q <= n_peers; ++q) {
                            uint32_t spins = 0;
                            while (ld_acquire_gpu(flag_of(group_id + q)) == 0u) {
                                if (++spins > (1u << 26)) __trap();
                            }
                        }
                    }
                    ptx::named_bar_sync(2, kStashEpiThreads);
                }
                if (et == 0) trace_stamp(8);
                const float4* sb4 = reinterpret_cast<const float4*>(sbt + col0);
                __half* out_row = Out + static_cast<size_t>(gm) * N + n0 + col0;
                const bool row_ok = gm < M;
                // the L2 round trip of the outlier product is what needs hiding (one chunk ahead); TMEM loads are short
                uint32_t v[32];
                uint4 fo[2][4];
                auto load_out0 = [&](int c, uint4 (&dst)[4]) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) dst[g] = has_outlier ? __ldcg(o0 + (c * 4 + g) * kStashEpiThreads) : make_uint4(0u, 0u, 0u, 0u);
                };
                load_out0(0, fo[0]);
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    uint4 pv[8];
                    if (n_peers > 0) {  // issue the first peer's loads before waiting for TMEM
                        const uint4* slot = slot_of(group_id + 1);
#pragma unroll
                        for (int g = 0; g < 8; ++g) pv[g] = __ldcg(slot + (c * 8 + g) * kStashEpiThreads);
                    }
                    ptx::tmem_ld_32x32(t_i + c * 32, v);
                    if (c + 1 < kChunks) load_out0(c + 1, fo[(c + 1) & 1]);
                    ptx::tmem_ld_wait();
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
                        const uint4 f = fo[c & 1][g];
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
    }

    if (threadIdx.x == kEpilogueWarp0 * 32) trace_stamp(7);
    ptx::tc_fence_before_sync();
    ptx::cluster_sync();
    if (warp_idx == 2) ptx::tmem_dealloc_2cta(tmem_base, T::kTmemCols);
}
