// gemm_split.cu — the parity-grade tcgen05 GEMM / implicit-GEMM convolution (default mode of the library).
//
// Operands are fp32-grade numbers carried as two fp16 planes, x * s = hi + lo (s a power of two, hi = fp16(x s),
// lo = fp16(x s - hi)): 11 + 11 mantissa bits plus the sign of lo, i.e. an fp32 significand.  A k-step issues the three
// products whose weight is >= 2^-12, (a_lo b_hi), (a_hi b_lo), (a_hi b_hi), on the fp16 tensor pipe with fp32 accumulation
// in TMEM: three UMMAs where an fp32 FMA pipe would need 2 x 16 x more issue slots (the dropped a_lo b_lo is 2^-24 relative).
//
// What makes the result match an fp32 reference to ~1e-7 instead of ~1e-5 (measurements: tools/trunc_probe.py):
//   * the tensor pipe TRUNCATES when it adds a k-step's products to the running fp32 accumulator.  So a TMEM accumulator only
//     lives for a BLOCK of `lblk` (1 or 2) 64-wide k-blocks; the epilogue warps then move the block sum to fp32 REGISTERS and add
//     it there with round-to-nearest adds, while the tensor pipe fills another of the four TMEM buffers.
//   * inside a block the small cross products (of all its stages) are issued BEFORE the hi x hi products: they are added while
//     the accumulator is still ~2^-11 of its final size, so only the 4 * lblk large adds truncate at full size; one TMEM
//     accumulator serves all three products.
//   * what remains is a data- and K-independent relative shrink of every block sum (8.9e-8 for one k-block, 1.55e-7 for two);
//     the epilogue multiplies it back (GemmParams::trunc_comp), leaving an unbiased error of ~1e-7 rel-rms per GEMM.
//
// Structure (same skeleton as gemm_tc.cu): 640 threads; warp 0 TMA producer (A hi/lo bricks, B hi/lo tiles, 128-byte swizzle,
// four stages), warp 1 UMMA issuer, warp 2 TMEM allocator, warps 4..19 epilogue: warp (quarter, cg) owns TMEM lanes 32 quarter..
// and accumulator columns 32 cg..; per block one tcgen05.ld + 16 packed adds; per tile bias / exact erf-GELU / SiLU / ReLU /
// LayerScale / residual on packed fp32 pairs (epi_split.cuh), then fp16 hi / lo planes (64-column chunks staged in swizzled
// shared memory by the two warps of a chunk, one TMA store per plane) or fp32 (the pair's buffer used in turn).  128-wide tiles
// run as CTA pairs (cta_group::2, M = 256): each SM stages its own A rows and half of the B rows, so a UMMA reads 6 KB of shared
// memory per 64 tensor-pipe cycles instead of 8.  64-wide single-CTA tiles serve N <= 64 and the DFL epilogue.
//
// Reference arithmetic replaced: see include/wedetect_b200.h (WD_OP_GEMM).
#include "gemm_params.h"
#include "epi_split.cuh"

namespace wd {

constexpr int kSplitEpiWarps = 16;
constexpr int kSplitThreads = 128 + 32 * kSplitEpiWarps;   // warps 0..3: TMA / UMMA / TMEM roles; warps 4..19: epilogue

template <int BN, bool kPair>
struct SCfg {
    static_assert(BN == 64 || BN == 128, "split tiles are 64 or 128 columns wide");
    static constexpr int A_BYTES = kTileM * 128;                 // one plane of this CTA's A rows
    static constexpr int B_ROWS = kPair ? BN / 2 : BN;
    static constexpr int B_BYTES = B_ROWS * 128;                 // one plane of this CTA's B rows
    static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);  // [A hi][A lo][B hi][B lo]
    static constexpr int EPI_BYTES = 8 * 4096;                   // 16-bit outputs: 32 rows x 128 B per PAIR of epilogue warps
    static constexpr int BAR_BYTES = 1024;
    static constexpr int kStagesRaw = (232448 - 1024 - BAR_BYTES - EPI_BYTES) / STAGE_BYTES;
    static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
    static constexpr int SMEM_BYTES = kStages * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
    static constexpr int NBUF = 4;                               // TMEM accumulator buffers (block sums in flight)
    static constexpr int TMEM_COLS = NBUF * BN;
    static constexpr int WCOLS = 32;                             // accumulator columns per epilogue warp
    static constexpr int NGRP = BN / WCOLS;                      // active column groups (x 4 lane quarters = active epilogue warps)
    static_assert(kStages >= 3, "pipeline too shallow");
    static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM cols");
};

// UMMA instruction descriptor (kind::f16): fp32 accumulate, fp16 A/B (format 0), both K-major
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) { return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24); }

template <int BN, typename OutT, bool kPair>
__global__ void __launch_bounds__(kSplitThreads, 1) gemm_split_kernel(const __grid_constant__ GemmParams p) {
    constexpr int kClu = kPair ? 2 : 1;
    pdl_launch_dependents();
    using C = SCfg<BN, kPair>;
    constexpr int CH = 128 / (int)sizeof(OutT);   // columns per output chunk (one 128 B swizzle row)
    constexpr bool kOut16 = sizeof(OutT) == 2;
    constexpr int WCOLS = C::WCOLS;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_epi = smem + C::kStages * C::STAGE_BYTES;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_epi + C::EPI_BYTES);
    uint64_t* bar_empty = bar_full + 8;
    uint64_t* bar_tfull = bar_empty + 8;
    uint64_t* bar_tempty = bar_tfull + 8;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_tempty + 8);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA[0]);
        tma_prefetch_desc(&p.tmA[1]);
        tma_prefetch_desc(&p.tmB[0]);
        tma_prefetch_desc(&p.tmB[1]);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 8; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
        for (int a = 0; a < C::NBUF; ++a) {
            mbar_init(&bar_tfull[a], 1);
            mbar_init(&bar_tempty[a], 4 * C::NGRP * kClu);   // one arrival per epilogue warp (of both CTAs in pair mode)
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        if constexpr (kPair) {
            tmem_alloc_2sm(tmem_ptr_smem, C::TMEM_COLS);
            tmem_relinquish_2sm();
        } else {
            tmem_alloc(tmem_ptr_smem, C::TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (kPair) {
        __syncwarp();
        cluster_sync_all();   // peer barriers initialised before any multicast can signal them
    }
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int k_iters = p.kc_iters * p.ntaps;
    const int lblk = p.lblk;
    const int crank = kPair ? (int)(blockIdx.x & 1) : 0;
    const int t_first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int t_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int t_total = kPair ? p.num_pair_tiles : p.num_tiles;

    pdl_wait();   // everything above overlapped the previous kernel's tail

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
        // ================= TMA producer =================
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx_bytes = 2u * (uint32_t)(p.rows_a * 128 + C::B_BYTES);
        for (int tile = t_first; tile < t_total; tile += t_step) {
            const int m_blk = (tile / p.num_n_tiles) * kClu + crank, n_blk = tile % p.num_n_tiles;
            const int t0 = m_blk % p.nt0, t1 = (m_blk / p.nt0) % p.nt1, t2 = m_blk / (p.nt0 * p.nt1);
            const int o0 = t0 * p.E0, o1 = t1 * p.E1, o2 = t2 * p.E2;   // beyond the tensor for the odd pair's ghost tile
            int kc = 0, tx_ = 0, dy = -p.pad;
            for (int kit = 0; kit < k_iters; ++kit) {
                const int dx = tx_ - p.pad;
                mbar_wait(&bar_empty[stage], phase ^ 1);
                __syncwarp();
                if (elect_one()) {
                    uint8_t* sA = smem + stage * C::STAGE_BYTES;
                    uint8_t* sB = sA + 2 * C::A_BYTES;
                    const int a0 = kc * kBlockK, a1 = o0 * p.a_step + dx, a2 = o1 * p.a_step + dy;
                    if constexpr (kPair) {
                        const uint32_t lead_full = mapa_u32(smem_u32(&bar_full[stage]), 0);
                        if (crank == 0) mbar_arrive_expect_tx(&bar_full[stage], 2u * tx_bytes);
                        const int b1 = n_blk * BN + crank * (BN / 2);
                        tma_load_4d_2sm(&p.tmA[0], lead_full, sA, a0, a1, a2, o2);
                        tma_load_4d_2sm(&p.tmA[1], lead_full, sA + C::A_BYTES, a0, a1, a2, o2);
                        tma_load_2d_2sm(&p.tmB[0], lead_full, sB, kit * kBlockK, b1);
                        tma_load_2d_2sm(&p.tmB[1], lead_full, sB + C::B_BYTES, kit * kBlockK, b1);
                    } else {
                        mbar_arrive_expect_tx(&bar_full[stage], tx_bytes);
                        tma_load_4d(&p.tmA[0], &bar_full[stage], sA, a0, a1, a2, o2);
                        tma_load_4d(&p.tmA[1], &bar_full[stage], sA + C::A_BYTES, a0, a1, a2, o2);
                        tma_load_2d(&p.tmB[0], &bar_full[stage], sB, kit * kBlockK, n_blk * BN);
                        tma_load_2d(&p.tmB[1], &bar_full[stage], sB + C::B_BYTES, kit * kBlockK, n_blk * BN);
                    }
                }
                __syncwarp();
                if (++kc == p.kc_iters) {
                    kc = 0;
                    if (++tx_ == p.tap_w) {
                        tx_ = 0;
                        ++dy;
                    }
                }
                if (++stage == C::kStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ================= UMMA issuer =================
        if (crank == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(kPair ? 2 * kTileM : kTileM, BN);
            const uint32_t smem0 = smem_u32(smem);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            // An accumulator block = `lblk` (1 or 2) consecutive pipeline stages.  All cross products of the block are issued before
            // its hi x hi products: the small terms meet a small accumulator, so only the 4 * lblk large adds truncate at full size.
            for (int tile = t_first; tile < t_total; tile += t_step) {
                for (int kit = 0; kit < k_iters; kit += lblk) {
                    // (an odd number of k-blocks ends with a one-stage block; a block may wrap around the ring)
                    const int nst = min(lblk, k_iters - kit);
                    const int st1 = stage + 1 == C::kStages ? 0 : stage + 1;
                    mbar_wait(&bar_tempty[as], aphase ^ 1);   // the epilogue has moved this buffer's previous block to registers
                    mbar_wait(&bar_full[stage], phase);
                    if (nst == 2) mbar_wait(&bar_full[st1], st1 == 0 ? phase ^ 1 : phase);
                    tc_fence_after();
                    __syncwarp();
                    if (elect_one()) {
                        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
                        for (int h = 0; h < nst; ++h) {
                            const uint32_t sbase = smem0 + (uint32_t)((h == 0 ? stage : st1) * C::STAGE_BYTES);
                            const uint32_t ah = (sbase >> 4) & 0x3FFFu, al = ((sbase + C::A_BYTES) >> 4) & 0x3FFFu;
                            const uint32_t bh = ((sbase + 2 * C::A_BYTES) >> 4) & 0x3FFFu, bl = ((sbase + 2 * C::A_BYTES + C::B_BYTES) >> 4) & 0x3FFFu;
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; ++k) {
                                const uint64_t dal = umma_desc_from_lo(al + 2 * k), dbh = umma_desc_from_lo(bh + 2 * k);
                                const uint64_t dah = umma_desc_from_lo(ah + 2 * k), dbl = umma_desc_from_lo(bl + 2 * k);
                                if constexpr (kPair) {
                                    umma_bf16_2sm(tmem_d, dal, dbh, idesc, (h == 0 && k == 0) ? 0u : 1u);
                                    umma_bf16_2sm(tmem_d, dah, dbl, idesc, 1u);
                                } else {
                                    umma_bf16(tmem_d, dal, dbh, idesc, (h == 0 && k == 0) ? 0u : 1u);
                                    umma_bf16(tmem_d, dah, dbl, idesc, 1u);
                                }
                            }
                        }
                        for (int h = 0; h < nst; ++h) {
                            const uint32_t sbase = smem0 + (uint32_t)((h == 0 ? stage : st1) * C::STAGE_BYTES);
                            const uint32_t ah = (sbase >> 4) & 0x3FFFu, bh = ((sbase + 2 * C::A_BYTES) >> 4) & 0x3FFFu;
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; ++k) {
                                const uint64_t dah = umma_desc_from_lo(ah + 2 * k), dbh = umma_desc_from_lo(bh + 2 * k);
                                if constexpr (kPair) umma_bf16_2sm(tmem_d, dah, dbh, idesc, 1u);
                                else umma_bf16(tmem_d, dah, dbh, idesc, 1u);
                            }
                        }
                        for (int h = 0; h < nst; ++h) {
                            if constexpr (kPair) umma_commit_2sm_mc(&bar_empty[h == 0 ? stage : st1], (uint16_t)3);
                            else umma_commit(&bar_empty[h == 0 ? stage : st1]);
                        }
                        if constexpr (kPair) umma_commit_2sm_mc(&bar_tfull[as], (uint16_t)3);
                        else umma_commit(&bar_tfull[as]);
                    }
                    __syncwarp();
                    stage += nst;
                    if (stage >= C::kStages) {
                        stage -= C::kStages;
                        phase ^= 1;
                    }
                    if (++as == C::NBUF) {
                        as = 0;
                        aphase ^= 1;
                    }
                }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ================= epilogue warps =================
        // Warp (quarter, cg) owns TMEM lanes 32 quarter .. +31 (rows of the tile) and accumulator columns 32 cg .. +31.
        const int e = warp - 4;
        const int quarter = warp & 3, cg = e >> 2;
        if (cg < C::NGRP) {
        const int r = quarter * 32 + lane;
        // 16-bit outputs leave through swizzled shared memory in 64-column (128-byte) chunks: the two warps of a chunk
        // (same rows, column groups 2m and 2m+1) share one 32-row staging buffer and a named barrier; the even one stores
        const int pairbuf = (cg >> 1) * 4 + quarter;
        uint8_t* wbuf = smem_epi + pairbuf * 4096;
        const uint32_t srow_s = smem_u32(wbuf) + lane * 128;
        const bool store_issuer = (cg & 1) == 0 && lane == 0;
        const int nblk = (k_iters + lblk - 1) / lblk;   // an odd number of k-blocks ends with a one-stage block ...
        const bool short_last = (k_iters % lblk) != 0;    // ... whose (smaller) truncation shrink is brought to the full blocks' first
        const uint64_t dlast = splat2(p.trunc_comp1 - p.trunc_comp);
        int as = 0;
        uint32_t aphase = 0;
        auto release_acc = [&](int a) {
            if (kPair && crank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_tempty[a]), 0));
            else mbar_arrive(&bar_tempty[a]);
        };
        // Everything a division would give is computed once: the thread's row inside the tile brick, the quarter's sub-brick, and
        // the tile coordinates, which then advance by the decomposed grid step (a runtime integer division is ~20 instructions;
        // ten of them per tile were a fifth of this loop's issue slots on the short-K layers).
        const int i0 = r % p.E0, i1 = (r / p.E0) % p.E1, i2 = r / (p.E0 * p.E1);
        const int qr = quarter * 32;
        const int q0 = qr % p.E0, q1 = (qr / p.E0) % p.E1, q2 = qr / (p.E0 * p.E1);
        const int step_n = t_step % p.num_n_tiles, step_m = (t_step / p.num_n_tiles) * kClu;
        const int a0 = step_m % p.nt0, a1 = (step_m / p.nt0) % p.nt1, a2 = step_m / (p.nt0 * p.nt1);
        int n_blk = t_first % p.num_n_tiles;
        int t0, t1, t2;
        {
            const int m_blk = (t_first / p.num_n_tiles) * kClu + crank;
            t0 = m_blk % p.nt0, t1 = (m_blk / p.nt0) % p.nt1, t2 = m_blk / (p.nt0 * p.nt1);
        }
        for (int tile = t_first; tile < t_total; tile += t_step) {
            const int o0 = t0 * p.E0, o1 = t1 * p.E1, o2 = t2 * p.E2;
            const int nb = n_blk * BN + cg * WCOLS;     // first output column of this warp
            const int n_blk_cur = n_blk;
            {   // coordinates of the next tile of this CTA
                n_blk += step_n;
                t0 += a0;
                if (n_blk >= p.num_n_tiles) {
                    n_blk -= p.num_n_tiles;
                    t0 += kClu;
                }
                while (t0 >= p.nt0) {
                    t0 -= p.nt0;
                    ++t1;
                }
                t1 += a1;
                while (t1 >= p.nt1) {
                    t1 -= p.nt1;
                    ++t2;
                }
                t2 += a2;
            }
            // pull this warp's bias / LayerScale line into L1 behind the tile's UMMAs (the epilogue's loads then hit)
            if (lane == 0) {
                if (p.bias && nb < p.N) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.bias + nb));
                if (p.gamma && nb < p.N) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.gamma + nb));
            }

            // ---- sum the block accumulators in fp32 registers (packed round-to-nearest adds) ----
            uint64_t acc2[WCOLS / 2];
            for (int b = 0; b < nblk; ++b) {
                mbar_wait(&bar_tfull[as], aphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + cg * WCOLS);
                uint32_t t[32];
                tmem_ld_32x32(taddr, t);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release_acc(as);   // the values are in registers: the tensor pipe may refill this buffer
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    uint64_t tv;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(tv) : "r"(t[2 * j]), "r"(t[2 * j + 1]));
                    if (short_last && b == nblk - 1) tv = fma2(tv, dlast, tv);
                    acc2[j] = b == 0 ? tv : add2(acc2[j], tv);
                }
                if (++as == C::NBUF) {
                    as = 0;
                    aphase ^= 1;
                }
            }
            // ---- this thread's output row ----
            const int d0 = o0 + i0, d1 = o1 + i1, d2 = o2 + i2;
            const bool row_ok = (i2 < p.E2) && d0 < p.D0 && d1 < p.D1 && d2 < p.D2;
            const long long pix = ((long long)d2 * p.D1 + d1) * p.D0 + d0;

            if (p.epi_mode == 1) {
                // ---- DFL epilogue (BN == N == 64): softmax over 16 bins x 4 sides, expectation (yolo_world_head.py:283-291);
                //      this warp's 32 columns are sides 2 cg and 2 cg + 1 ----
                if constexpr (BN == 64) {
                    float acc[32];
#pragma unroll
                    for (int j = 0; j < 16; ++j) upk2(acc2[j], acc[2 * j], acc[2 * j + 1]);
                    float out2[2];
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        float mx = -INFINITY;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float a = fmaf(acc[s * 16 + j], p.trunc_comp, acc[s * 16 + j]);
                            acc[s * 16 + j] = fmaf(a, p.acc_scale, __ldg(p.bias + cg * 32 + s * 16 + j));
                            mx = fmaxf(mx, acc[s * 16 + j]);
                        }
                        float den = 0.f, num = 0.f;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float ex = expf(acc[s * 16 + j] - mx);
                            den += ex;
                            num += ex * (float)j;
                        }
                        out2[s] = num / den;
                    }
                    if (row_ok) *reinterpret_cast<float2*>(p.dfl_out + pix * 4 + cg * 2) = make_float2(out2[0], out2[1]);
                }
                continue;
            }

            // ---- acc = osc * gamma * act(acc * (1 + trunc_comp) * acc_scale + bias), two columns per instruction; osc = the plane
            //      scale of a 16-bit output (powers of two commute with every rounding below) ----
            const float osc = kOut16 ? kPlaneScale : 1.f;
            switch (p.act) {
                case WD_ACT_RELU: split_epi_math<WCOLS, WD_ACT_RELU>(acc2, p.trunc_comp, p.acc_scale, osc, p.bias, p.gamma, nb, p.N); break;
                case WD_ACT_SILU: split_epi_math<WCOLS, WD_ACT_SILU>(acc2, p.trunc_comp, p.acc_scale, osc, p.bias, p.gamma, nb, p.N); break;
                case WD_ACT_GELU: split_epi_math<WCOLS, WD_ACT_GELU>(acc2, p.trunc_comp, p.acc_scale, osc, p.bias, p.gamma, nb, p.N); break;
                default: split_epi_math<WCOLS, WD_ACT_NONE>(acc2, p.trunc_comp, p.acc_scale, osc, p.bias, p.gamma, nb, p.N); break;
            }
            const bool cols_ok = nb < p.N;      // warp-uniform: some of this warp's columns exist
            // ---- acc += resid * alpha ----
            if (cols_ok && row_ok) {
                const uint64_t alpha2 = splat2(p.alpha * osc);   // acc already carries the output's plane scale
                if (p.resid_dtype == 2) {
                    const float* rp = reinterpret_cast<const float*>(p.resid) + pix * p.ld_res + nb;
                    if (nb + WCOLS <= p.N && (p.ld_res & 7) == 0 && (reinterpret_cast<uintptr_t>(p.resid) & 31) == 0) {
                        // whole 32-byte sectors per lane: one request per sector instead of two (rows are 128 bytes apart per lane)
#pragma unroll
                        for (int j = 0; j < WCOLS; j += 8) {
                            uint64_t x[4];
                            asm volatile("ld.global.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(x[0]), "=l"(x[1]), "=l"(x[2]), "=l"(x[3]) : "l"(rp + j));
#pragma unroll
                            for (int q = 0; q < 4; ++q) acc2[j / 2 + q] = fma2(alpha2, x[q], acc2[j / 2 + q]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < WCOLS; j += 4) {
                            if (nb + j < p.N) {
                                const float4 x = *reinterpret_cast<const float4*>(rp + j);
                                acc2[j / 2] = fma2(alpha2, pk2(x.x, x.y), acc2[j / 2]);
                                acc2[j / 2 + 1] = fma2(alpha2, pk2(x.z, x.w), acc2[j / 2 + 1]);
                            }
                        }
                    }
                } else if (p.resid_dtype == 1) {
                    // fp16 hi (+ lo) planes; alpha already carries 1 / kPlaneScale
                    const __half* rp = reinterpret_cast<const __half*>(p.resid) + pix * p.ld_res + nb;
                    if (nb + WCOLS <= p.N && (p.ld_res & 15) == 0 && (p.resid_ps & 15) == 0 && (reinterpret_cast<uintptr_t>(p.resid) & 31) == 0) {
                        // whole 32-byte sectors per lane (16 columns per load), as for the fp32 residual above
#pragma unroll
                        for (int j = 0; j < WCOLS; j += 16) {
                            uint32_t x[8];
                            asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                         : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]) : "l"(rp + j));
                            uint64_t xs[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q) xs[q] = h2_to_f2(x[q]);
                            if (p.resid_ps) {
                                uint32_t y[8];
                                asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                             : "=r"(y[0]), "=r"(y[1]), "=r"(y[2]), "=r"(y[3]), "=r"(y[4]), "=r"(y[5]), "=r"(y[6]), "=r"(y[7]) : "l"(rp + p.resid_ps + j));
#pragma unroll
                                for (int q = 0; q < 8; ++q) xs[q] = add2(xs[q], h2_to_f2(y[q]));
                            }
#pragma unroll
                            for (int q = 0; q < 8; ++q) acc2[j / 2 + q] = fma2(alpha2, xs[q], acc2[j / 2 + q]);
                        }
                    } else
#pragma unroll
                    for (int j = 0; j < WCOLS; j += 8) {
                        if (nb + j < p.N) {
                            const uint4 x = *reinterpret_cast<const uint4*>(rp + j);
                            uint64_t xs[4] = {h2_to_f2(x.x), h2_to_f2(x.y), h2_to_f2(x.z), h2_to_f2(x.w)};
                            if (p.resid_ps) {
                                const uint4 y = *reinterpret_cast<const uint4*>(rp + p.resid_ps + j);
                                xs[0] = add2(xs[0], h2_to_f2(y.x)); xs[1] = add2(xs[1], h2_to_f2(y.y));
                                xs[2] = add2(xs[2], h2_to_f2(y.z)); xs[3] = add2(xs[3], h2_to_f2(y.w));
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) acc2[j / 2 + q] = fma2(alpha2, xs[q], acc2[j / 2 + q]);
                        }
                    }
                }
            }
            // ---- output ----
            const int g = p.group_cols > nb ? 0 : nb / p.group_cols, c0 = nb - g * p.group_cols;     // column inside its output group (concat / deconv layouts)
            const long long grow = (long long)d0 * p.sc0 + (long long)d1 * p.sc1 + (long long)d2 * p.sc2 + (long long)g * p.scg + c0;
            if constexpr (!kOut16) {
                // fp32: this warp's 32 columns are one 128-byte chunk.  The two warps of a pair take turns on the pair's staging
                // buffer (even warp first): the odd warp waits for the even warp's TMA store to have read the buffer.
                if (p.warp_store) {
                    const bool odd = (cg & 1) != 0;
                    if (odd && lane == 0) tma_store_wait_read<0>();      // my store of the previous tile has released the buffer
                    named_bar_sync(1 + pairbuf, 64);
                    if (odd) named_bar_sync(1 + pairbuf, 64);            // ... until the even warp's store has read its rows
                    if (cols_ok) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float f0, f1, f2, f3;
                            upk2(acc2[q * 2], f0, f1);
                            upk2(acc2[q * 2 + 1], f2, f3);
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow_s + ((q ^ (lane & 7)) << 4)), "f"(f0), "f"(f1), "f"(f2), "f"(f3) : "memory");
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if (p.red_store) tma_reduce_add_5d(&p.tmCw[0], wbuf, c0, o0 + q0, o1 + q1, o2 + q2, g);
                            else tma_store_5d(&p.tmCw[0], wbuf, c0, o0 + q0, o1 + q1, o2 + q2, g);
                            tma_store_commit();
                            if (!odd) tma_store_wait_read<0>();
                        }
                        __syncwarp();
                    }
                    if (!odd) named_bar_sync(1 + pairbuf, 64);           // hand the buffer to the odd warp
                } else if (cols_ok && row_ok) {
                    // the 32 rows of a lane quarter are not a sub-brick of the tile: each thread writes its own row
                    float* op = reinterpret_cast<float*>(p.out) + grow;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (c0 + q * 4 < p.cols_valid) {
                            float f0, f1, f2, f3;
                            upk2(acc2[q * 2], f0, f1);
                            upk2(acc2[q * 2 + 1], f2, f3);
                            *reinterpret_cast<float4*>(op + q * 4) = make_float4(f0, f1, f2, f3);
                        }
                    }
                }
            } else {
                // fp16 hi / lo planes of acc * kPlaneScale (one pass per plane)
                const int chunk_n = n_blk_cur * BN + (cg & ~1) * WCOLS;      // first column of the 64-column chunk this warp pair writes
                const bool chunk_ok = chunk_n < p.N;                        // uniform over the pair
                const int gc = p.group_cols > chunk_n ? 0 : chunk_n / p.group_cols, cc0 = chunk_n - gc * p.group_cols;
#pragma unroll
                for (int pl = 0; pl < 2; ++pl) {
                    uint4 w[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t ww[4];
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const uint64_t a = acc2[q * 4 + h];
                            ww[h] = cvt_h2_sat(a);
                            if (pl == 0) acc2[q * 4 + h] = add2(a, neg2(h2_to_f2(ww[h])));   // the remainder goes to the low plane
                        }
                        w[q] = make_uint4(ww[0], ww[1], ww[2], ww[3]);
                    }
                    if (p.warp_store) {
                        if (!chunk_ok) continue;
                        if (store_issuer) tma_store_wait_read<0>();   // the pair's staging buffer is no longer being read by the previous store
                        named_bar_sync(1 + pairbuf, 64);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow_s + ((((cg & 1) * 4 + q) ^ (lane & 7)) << 4)), "r"(w[q].x), "r"(w[q].y), "r"(w[q].z), "r"(w[q].w) : "memory");
                        fence_proxy_async_smem();
                        named_bar_sync(1 + pairbuf, 64);
                        if (store_issuer) {
                            tma_store_5d(&p.tmCw[pl], wbuf, cc0, o0 + q0, o1 + q1, o2 + q2, gc);
                            tma_store_commit();
                        }
                    } else if (cols_ok && row_ok) {
                        // the 32 rows of a lane quarter are not a sub-brick of the tile: each thread writes its own row
                        __half* op = reinterpret_cast<__half*>(p.out) + (long long)pl * p.out_ps + grow;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (c0 + q * 8 < p.cols_valid) *reinterpret_cast<uint4*>(op + q * 8) = w[q];
                    }
                }
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (kPair) {
        __syncwarp();
        cluster_sync_all();   // the peer may still be arriving on this CTA's barriers
    }
    if (warp == 2) {
        if constexpr (kPair) tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
        else tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int BN, typename OutT, bool kPair>
static int launch_split_inst(const GemmOp& g, cudaStream_t s) {
    auto kern = gemm_split_kernel<BN, OutT, kPair>;
    using C = SCfg<BN, kPair>;
    WD_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(kern), C::SMEM_BYTES));
    WD_CHECK_CUDA(launch_pdl(kern, dim3(g.grid), dim3(kSplitThreads), (size_t)C::SMEM_BYTES, s, kPair ? 2 : 1, g.prm));
    WD_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_gemm_split(const GemmOp& g, cudaStream_t s) {
    const bool pair = g.prm.clu == 2;
    if (g.block_n == 64 && !pair) return g.out_f32 ? launch_split_inst<64, float, false>(g, s) : launch_split_inst<64, __half, false>(g, s);
    if (g.block_n == 128 && !pair) return g.out_f32 ? launch_split_inst<128, float, false>(g, s) : launch_split_inst<128, __half, false>(g, s);
    if (g.block_n == 128 && pair) return g.out_f32 ? launch_split_inst<128, float, true>(g, s) : launch_split_inst<128, __half, true>(g, s);
    set_last_error("gemm (split): unsupported block_n=%d pair=%d", g.block_n, (int)pair);
    return -1;
}

}  // namespace wd
