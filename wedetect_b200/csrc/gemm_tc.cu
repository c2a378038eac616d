// gemm_tc.cu — persistent, warp-specialised tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   C[m, n] = epilogue( sum_{tap, k} A[pix(m) + off(tap), k] * B[n, tap*Kc + k] )
//
// * A rows live in a (d0, d1, d2) pixel space (NHWC activations: d0 = x, d1 = y, d2 = image; a plain
//   row-major matrix is D0 = M, D1 = D2 = 1).  A tile is an (E0, E1, E2) brick of <= 128 pixels, so a
//   3x3 convolution is 9 shifted TMA brick loads (out-of-bounds pixels are zero-filled by TMA = conv
//   padding) — no im2col buffer.
// * operands are bf16, K-major, 128-byte swizzled in shared memory (written by TMA, read by UMMA);
//   accumulation is fp32 in TMEM, double-buffered (2 x BLOCK_N columns) so the epilogue of tile i
//   overlaps the MMAs of tile i+1.
// * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (single elected lane), warp 2 = TMEM
//   allocator, warps 4..11 = two epilogue warpgroups that split the accumulator in column chunks,
//   apply bias / activation / LayerScale / residual (or DFL), stage the tile in swizzled smem and
//   write it back with TMA stores (which clip partial tiles and scatter into concat / upsample
//   layouts through a rank-5 tensor map).
// * the parity-grade mode (fp16 hi/lo operand pairs, three UMMAs per k-step) is a separate kernel: gemm_split.cu.
//
// Reference arithmetic replaced: see include/wedetect_b200.h (WD_OP_GEMM).
#include "gemm_params.h"
#include "epi_math.cuh"
#include <stdio.h>
#include <string.h>

namespace wd {

template <int BN>
struct Cfg {
    static constexpr int A_BYTES = kTileM * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 5 : 6);
    static constexpr int EPI_BUFS = BN <= 128 ? 2 : 1;
    static constexpr int EPI_BUF_BYTES = 16384;
    static constexpr int EPI_BYTES = kNumEpiWG * EPI_BUFS * EPI_BUF_BYTES;
    static constexpr int BAR_BYTES = 1024;
    static constexpr int SMEM_BYTES = kStages * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
    static constexpr int TMEM_COLS = 2 * BN;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
    static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM cols");
};

template <int BN, typename OutT, bool kPair = false>
__global__ void __launch_bounds__(kNumThreads, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
    // kPair: the kernel is launched in clusters of two CTAs and every tcgen05 instruction is the cta_group::2 form (a kernel
    // may not mix the two forms: ptxas tags it TCGEN05_2CTA_USED and the driver then refuses a launch without clusters)
    constexpr int kClu = kPair ? 2 : 1;
    pdl_launch_dependents();
    using C = Cfg<BN>;
    constexpr int CH = 128 / (int)sizeof(OutT);  // columns per epilogue chunk (one 128 B swizzle row)
    constexpr bool kOutBf16 = sizeof(OutT) == 2;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_epi = smem + C::kStages * C::STAGE_BYTES;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_epi + C::EPI_BYTES);
    uint64_t* bar_empty = bar_full + 8;    // up to 8 stages (pair mode uses 6 half-width stages)
    uint64_t* bar_tfull = bar_empty + 8;
    uint64_t* bar_tempty = bar_tfull + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA[0]);
        tma_prefetch_desc(&p.tmB[0]);
        tma_prefetch_desc(&p.tmC[0]);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 8; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bar_tfull[a], 1);
            mbar_init(&bar_tempty[a], 4 * kNumEpiWG * kClu);   // pair mode: the leader waits for both CTAs' epilogues
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
    // work distribution: plain persistent CTAs, or CTA pairs walking (m-pair, n) tiles in lock step
    // (cluster dims are (2,1,1): rank and cluster id follow from blockIdx, which the compiler knows to be warp-uniform)
    const int crank = kPair ? (int)(blockIdx.x & 1) : 0;
    const int t_first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int t_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int t_total = kPair ? p.num_pair_tiles : p.num_tiles;
    // pair mode (cta_group::2): each CTA stages its own 128 A rows and HALF of the B tile; one UMMA of M = 256 issued by
    // the leader reads both halves, so every SM moves 32 KB per k-step through its shared memory instead of 48 KB
    const int stage_bytes = kPair ? (C::A_BYTES + C::B_BYTES / 2) : C::STAGE_BYTES;
    const int nstages = kPair ? (C::kStages * C::STAGE_BYTES) / (C::A_BYTES + C::B_BYTES / 2) : C::kStages;

    pdl_wait();   // everything above overlapped the previous kernel's tail; from here on global memory is read / written

    // Register re-balancing between the warpgroups (inside the role branches, so the allocator sees which code runs under
    // which budget): the producer / MMA / allocator warps need few registers; the epilogue warps hold a 64-column accumulator
    // slice, a prefetched residual slice and staging addresses each.  128 x 80 + 256 x 208 <= 384 x 168 (the launch allocation).
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    if (warp == 0) {
        // ================= TMA producer =================
        // The whole warp walks the loops (warp-uniform values live in uniform registers and feed UTMALDG directly); one
        // elected lane arms the barrier and issues.  Taps and k-chunks advance by counters: no division per k-step.
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx_bytes = (uint32_t)(p.rows_a * 128 + C::B_BYTES);
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
                    if constexpr (kPair) {
                        uint8_t* sA = smem + stage * stage_bytes;
                        const uint32_t lead_full = mapa_u32(smem_u32(&bar_full[stage]), 0);
                        if (crank == 0) mbar_arrive_expect_tx(&bar_full[stage], 2u * (uint32_t)(p.rows_a * 128 + C::B_BYTES / 2));
                        tma_load_4d_2sm(&p.tmA[0], lead_full, sA, kc * kBlockK, o0 * p.a_step + dx, o1 * p.a_step + dy, o2);
                        tma_load_2d_2sm(&p.tmB[0], lead_full, sA + C::A_BYTES, kit * kBlockK, n_blk * BN + crank * (BN / 2));
                    } else {
                        mbar_arrive_expect_tx(&bar_full[stage], tx_bytes);
                        uint8_t* sA = smem + stage * C::STAGE_BYTES;
                        tma_load_4d(&p.tmA[0], &bar_full[stage], sA, kc * kBlockK, o0 * p.a_step + dx, o1 * p.a_step + dy, o2);
                        tma_load_2d(&p.tmB[0], &bar_full[stage], sA + C::A_BYTES, kit * kBlockK, n_blk * BN);
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
                if (++stage == nstages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // Same scheme: warp-uniform loops, one elected lane issues the UMMAs and their commits.  The issue loop has to stay
        // well under 128 cycles per UMMA (the tensor pipe's time for a 128x256x16 step), or the pipe idles between them.
        if (crank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
            constexpr uint32_t idesc2 = umma_idesc_bf16(2 * kTileM, BN);   // M = 256 across the CTA pair
            const uint32_t smem0 = smem_u32(smem);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = t_first; tile < t_total; tile += t_step) {
                mbar_wait(&bar_tempty[as], aphase ^ 1);
                tc_fence_after();
                for (int kit = 0; kit < k_iters; ++kit) {
                    const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
                    mbar_wait(&bar_full[stage], phase);
                    tc_fence_after();
                    __syncwarp();
                    if (elect_one()) {
                        // descriptor low words: (address >> 4) in 14 bits; +32 bytes along K inside the swizzle atom = +2
                        const uint32_t a_lo = ((smem0 + (uint32_t)(stage * stage_bytes)) >> 4) & 0x3FFFu;
                        const uint32_t b_lo = ((smem0 + (uint32_t)(stage * stage_bytes) + C::A_BYTES) >> 4) & 0x3FFFu;
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint64_t ad = umma_desc_from_lo(a_lo + 2 * k), bd = umma_desc_from_lo(b_lo + 2 * k);
                            if constexpr (kPair) umma_bf16_2sm(tmem_d, ad, bd, idesc2, (kit | k) != 0 ? 1u : 0u);
                            else umma_bf16(tmem_d, ad, bd, idesc, (kit | k) != 0 ? 1u : 0u);
                        }
                        // frees the smem slot when these MMAs retire (in pair mode: tells BOTH CTAs' producers)
                        if constexpr (kPair) umma_commit_2sm_mc(&bar_empty[stage], (uint16_t)3);
                        else umma_commit(&bar_empty[stage]);
                        if (kit == k_iters - 1) {
                            // accumulator complete -> epilogue (of both CTAs in pair mode)
                            if constexpr (kPair) umma_commit_2sm_mc(&bar_tfull[as], (uint16_t)3);
                            else umma_commit(&bar_tfull[as]);
                        }
                    }
                    __syncwarp();
                    if (++stage == nstages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (++as == 2) {
                    as = 0;
                    aphase ^= 1;
                }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        // ================= epilogue warpgroups =================
        const int wg = (warp - 4) >> 2;
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        const int tid_wg = threadIdx.x - 128 - wg * 128;
        const bool issuer = (tid_wg == 0);
        const int r = quarter * 32 + lane;
        uint8_t* wg_bufs = smem_epi + wg * C::EPI_BUFS * C::EPI_BUF_BYTES;
        int as = 0;
        uint32_t aphase = 0;
        int buf = 0;
        constexpr int n_chunks = BN / CH;
        // hand an accumulator buffer back to the MMA issuer (the pair leader's barrier in pair mode)
        auto release_acc = [&](int a) {
            if (kPair && crank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_tempty[a]), 0));
            else mbar_arrive(&bar_tempty[a]);
        };

        for (int tile = t_first; tile < t_total; tile += t_step) {
            const int m_blk = (tile / p.num_n_tiles) * kClu + crank, n_blk = tile % p.num_n_tiles;
            const int t0 = m_blk % p.nt0, t1 = (m_blk / p.nt0) % p.nt1, t2 = m_blk / (p.nt0 * p.nt1);
            const int o0 = t0 * p.E0, o1 = t1 * p.E1, o2 = t2 * p.E2;
            const int i0 = r % p.E0, i1 = (r / p.E0) % p.E1, i2 = r / (p.E0 * p.E1);
            const int d0 = o0 + i0, d1 = o1 + i1, d2 = o2 + i2;
            // first row of this warp's quarter inside the tile brick (per-warp TMA stores)
            const int qr = quarter * 32;
            const int q0 = qr % p.E0, q1 = (qr / p.E0) % p.E1, q2 = qr / (p.E0 * p.E1);
            const bool row_ok = (i2 < p.E2) && d0 < p.D0 && d1 < p.D1 && d2 < p.D2;
            const long long pix = ((long long)d2 * p.D1 + d1) * p.D0 + d0;

            // ---- residual prefetch (single-plane residual of the output's dtype): the 8 x 16 B loads of a chunk are
            //      issued before the accumulator is waited for / read, so their DRAM latency hides behind the MMAs ----
            uint4 rq[8];
            bool rq_valid = false;
            const bool rq_ok = row_ok && p.resid_ps == 0 && p.resid_dtype == (kOutBf16 ? 1 : 2);
            auto prefetch_resid = [&](int c) {
                rq_valid = false;
                const int n_base = n_blk * BN + c * CH;
                if (!rq_ok || c >= n_chunks || n_base >= p.N) return;
                const uint8_t* rp = reinterpret_cast<const uint8_t*>(p.resid) + (pix * p.ld_res + n_base) * (long long)sizeof(OutT);
                constexpr int EPV = 16 / (int)sizeof(OutT);   // elements per 16-byte vector
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    rq[q] = make_uint4(0u, 0u, 0u, 0u);
                    if (n_base + q * EPV < p.N) rq[q] = *reinterpret_cast<const uint4*>(rp + q * 16);
                }
                rq_valid = true;
            };
            // ---- per column chunk: v = resid*alpha + gamma * act(acc + bias) -> swizzled smem -> TMA store ----
            auto finish_chunk = [&](float* v, int c) {
                const int n_base = n_blk * BN + c * CH;
                if (n_base < p.N) {  // warp-uniform: whole chunk beyond N is skipped (nothing to store)
                    // ---- math: v = resid*alpha + gamma * act(acc + bias); the activation is uniform per launch ----
                    constexpr bool kFast = kOutBf16;
                    // warp-uniform: 1 / 2 = whole chunk inside N with bias (and gamma) -> unchecked forms
                    const int mode = (p.bias != nullptr && n_base + CH <= p.N) ? (p.gamma ? 2 : 1) : 0;
#define WD_EPI(ACT, FAST)                                                                        \
    do {                                                                                         \
        if (mode == 1) epi_bias_act<CH, ACT, FAST, 1>(v, p.bias, p.gamma, n_base, p.N);          \
        else if (mode == 2) epi_bias_act<CH, ACT, FAST, 2>(v, p.bias, p.gamma, n_base, p.N);     \
        else epi_bias_act<CH, ACT, FAST, 0>(v, p.bias, p.gamma, n_base, p.N);                    \
    } while (0)
                    if (kFast && !p.exact_act) {
                        switch (p.act) {
                            case WD_ACT_RELU: epi_bias_act<CH, WD_ACT_RELU, kFast, 0>(v, p.bias, p.gamma, n_base, p.N); break;
                            case WD_ACT_SILU: WD_EPI(WD_ACT_SILU, kFast); break;
                            case WD_ACT_GELU: WD_EPI(WD_ACT_GELU, kFast); break;
                            default: WD_EPI(WD_ACT_NONE, kFast); break;
                        }
                    } else {
                        switch (p.act) {
                            case WD_ACT_RELU: epi_bias_act<CH, WD_ACT_RELU, false, 0>(v, p.bias, p.gamma, n_base, p.N); break;
                            case WD_ACT_SILU: epi_bias_act<CH, WD_ACT_SILU, false, 0>(v, p.bias, p.gamma, n_base, p.N); break;
                            case WD_ACT_GELU: epi_bias_act<CH, WD_ACT_GELU, false, 0>(v, p.bias, p.gamma, n_base, p.N); break;
                            default: WD_EPI(WD_ACT_NONE, false); break;
                        }
                    }
#undef WD_EPI
                    if (rq_valid) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if constexpr (kOutBf16) {
                                v[q * 8 + 0] += p.alpha * bf16_lo(rq[q].x); v[q * 8 + 1] += p.alpha * bf16_hi(rq[q].x);
                                v[q * 8 + 2] += p.alpha * bf16_lo(rq[q].y); v[q * 8 + 3] += p.alpha * bf16_hi(rq[q].y);
                                v[q * 8 + 4] += p.alpha * bf16_lo(rq[q].z); v[q * 8 + 5] += p.alpha * bf16_hi(rq[q].z);
                                v[q * 8 + 6] += p.alpha * bf16_lo(rq[q].w); v[q * 8 + 7] += p.alpha * bf16_hi(rq[q].w);
                            } else {
                                v[q * 4 + 0] += p.alpha * __uint_as_float(rq[q].x); v[q * 4 + 1] += p.alpha * __uint_as_float(rq[q].y);
                                v[q * 4 + 2] += p.alpha * __uint_as_float(rq[q].z); v[q * 4 + 3] += p.alpha * __uint_as_float(rq[q].w);
                            }
                        }
                        prefetch_resid(c + kNumEpiWG);   // next chunk of this warpgroup: overlaps staging, store and TMEM read
                    } else if (p.resid_dtype != 0 && row_ok) {
                        if (p.resid_dtype == 2) {
                            const float* rp = reinterpret_cast<const float*>(p.resid) + pix * p.ld_res + n_base;
#pragma unroll
                            for (int j = 0; j < CH; j += 4) {
                                if (n_base + j < p.N) {
                                    const float4 x = *reinterpret_cast<const float4*>(rp + j);
                                    v[j + 0] += p.alpha * x.x;
                                    v[j + 1] += p.alpha * x.y;
                                    v[j + 2] += p.alpha * x.z;
                                    v[j + 3] += p.alpha * x.w;
                                }
                            }
                        } else {
                            const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.resid) + pix * p.ld_res + n_base;
#pragma unroll
                            for (int j = 0; j < CH; j += 8) {
                                if (n_base + j < p.N) {
                                    const uint4 x = *reinterpret_cast<const uint4*>(rp + j);
                                    const float xs[8] = {bf16_lo(x.x), bf16_hi(x.x), bf16_lo(x.y), bf16_hi(x.y),
                                                         bf16_lo(x.z), bf16_hi(x.z), bf16_lo(x.w), bf16_hi(x.w)};
#pragma unroll
                                    for (int q = 0; q < 8; ++q) v[j + q] += p.alpha * xs[q];
                                }
                            }
                        }
                    }
                    // ---- stage into swizzled smem, then TMA store ----
                    uint8_t* sbuf = wg_bufs + buf * C::EPI_BUF_BYTES;
                    const uint32_t srow_s = smem_u32(sbuf) + r * 128;   // explicit shared-space stores (STS), not generic ST
                    const int g = n_base / p.group_cols, c0 = n_base - g * p.group_cols;
                    {
                        constexpr int pl = 0;
                        // buffer `buf` no longer being read by an earlier store.  warp_store: every warp stages and stores its
                        // own 32 rows (4 KB of the buffer) with its own bulk groups, so the four warps never wait for each other
                        if (p.warp_store) {
                            if (lane == 0) tma_store_wait_read<C::EPI_BUFS - 1>();
                            __syncwarp();
                        } else {
                            if (issuer) tma_store_wait_read<C::EPI_BUFS - 1>();
                            named_bar_sync(1 + wg, 128);
                        }
                        if constexpr (kOutBf16) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                uint4 w;
                                w.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
                                w.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
                                w.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
                                w.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow_s + ((q ^ (r & 7)) << 4)), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow_s + ((q ^ (r & 7)) << 4)), "f"(v[q * 4 + 0]), "f"(v[q * 4 + 1]), "f"(v[q * 4 + 2]), "f"(v[q * 4 + 3]) : "memory");
                            }
                        }
                        fence_proxy_async_smem();
                        if (p.warp_store) {
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_5d(&p.tmCw[pl], sbuf + quarter * 4096, c0, o0 + q0, o1 + q1, o2 + q2, g);
                                tma_store_commit();
                            }
                        } else {
                            named_bar_sync(1 + wg, 128);
                            if (issuer) {
                                tma_store_5d(&p.tmC[pl], sbuf, c0, o0, o1, o2, g);
                                tma_store_commit();
                            }
                        }
                    }
                    buf = (buf + 1) % C::EPI_BUFS;
                }
            };
            auto finish_dfl = [&](float* v) {
                float out4[4];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    float mx = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        v[s * 16 + j] += __ldg(p.bias + s * 16 + j);
                        mx = fmaxf(mx, v[s * 16 + j]);
                    }
                    float den = 0.f, num = 0.f;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float e = expf(v[s * 16 + j] - mx);
                        den += e;
                        num += e * (float)j;
                    }
                    out4[s] = num / den;
                }
                if (row_ok)
                    *reinterpret_cast<float4*>(p.dfl_out + pix * 4) = make_float4(out4[0], out4[1], out4[2], out4[3]);
            };
            {
                if (p.epi_mode == 0) prefetch_resid(wg);
                mbar_wait(&bar_tfull[as], aphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN);
                if (p.epi_mode == 1) {
                    // ---- DFL epilogue (BN == 64): softmax over 16 bins x 4 sides, expectation ----
                    if constexpr (BN == 64) if (wg == 0) {
                        float v[64];
                        tmem_ld_32x32(taddr, reinterpret_cast<uint32_t*>(v));
                        tmem_ld_32x32(taddr + 32, reinterpret_cast<uint32_t*>(v + 32));
                        tmem_ld_wait();
                        finish_dfl(v);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) release_acc(as);
                } else {
                    if (wg >= n_chunks) {
                        // this warpgroup owns no column chunk of the tile (BN == CH): still hand the accumulator back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) release_acc(as);
                    }
                    for (int c = wg; c < n_chunks; c += kNumEpiWG) {
                        float v[CH];
#pragma unroll
                        for (int j = 0; j < CH; j += 32) tmem_ld_32x32(taddr + c * CH + j, reinterpret_cast<uint32_t*>(v + j));
                        tmem_ld_wait();
                        if (c + kNumEpiWG >= n_chunks) {
                            // last TMEM read of this warp for this tile: hand the accumulator back
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) release_acc(as);
                        }
                        finish_chunk(v, c);
                    }
                }
                if (++as == 2) {
                    as = 0;
                    aphase ^= 1;
                }
            }
        }
        if (p.warp_store ? (lane == 0) : issuer) tma_store_wait_all<0>();
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
template <int BN, typename OutT, bool kPair = false>
static int launch_inst(const GemmOp& g, cudaStream_t s) {
    auto kern = gemm_tc_kernel<BN, OutT, kPair>;
    WD_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(kern), Cfg<BN>::SMEM_BYTES));
    WD_CHECK_CUDA(launch_pdl(kern, dim3(g.grid), dim3(kNumThreads), (size_t)Cfg<BN>::SMEM_BYTES, s, kPair ? 2 : 1, g.prm));
    WD_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int GemmOp::launch(cudaStream_t s) {
    if (split) return launch_gemm_split(*this, s);
    if (block_n == 64) return out_f32 ? launch_inst<64, float>(*this, s) : launch_inst<64, __nv_bfloat16>(*this, s);
    if (block_n == 128) return out_f32 ? launch_inst<128, float>(*this, s) : launch_inst<128, __nv_bfloat16>(*this, s);
    if (block_n == 256 && prm.clu == 2) return out_f32 ? launch_inst<256, float, true>(*this, s) : launch_inst<256, __nv_bfloat16, true>(*this, s);
    if (block_n == 256) return out_f32 ? launch_inst<256, float>(*this, s) : launch_inst<256, __nv_bfloat16>(*this, s);
    set_last_error("gemm: unsupported block_n=%d", block_n);
    return -1;
}

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

int compile_gemm(const wd_op& op, std::unique_ptr<CompiledOp>& out) {
    const int32_t* I = op.i;
    auto g = std::make_unique<GemmOp>();
    GemmParams& P = g->prm;
    memset(&P, 0, sizeof(P));
    P.D0 = I[0]; P.D1 = I[1]; P.D2 = I[2];
    P.E0 = I[3]; P.E1 = I[4]; P.E2 = I[5];
    const int Kc = I[6];
    P.ntaps = I[7];
    P.N = I[8];
    const long long sa0 = I[9], sa1 = I[10], sa2 = I[11];
    const int ldb = I[12];
    g->block_n = I[13];
    g->out_f32 = I[14];
    P.act = I[15];
    P.resid_dtype = I[16];
    P.ld_res = I[17];
    P.group_cols = I[18];
    const int n_groups = I[19];
    const long long sc0 = I[20], sc1 = I[21], sc2 = I[22], scg = I[23];
    P.epi_mode = I[24];
    P.exact_act = I[35];
    P.tap_w = I[25] > 0 ? I[25] : 1;
    P.pad = I[26];
    const int group_valid = I[27] > 0 ? I[27] : I[18];   // columns of a group that exist in memory
    const int k_valid = I[28] > 0 ? I[28] : Kc;          // A channels that exist (rest zero-filled by TMA)
    const long long bk_valid = I[29] > 0 ? I[29] : (long long)Kc * I[7];
    P.alpha = op.f[0];
    P.bias = (const float*)op.p[3];
    P.gamma = (const float*)op.p[4];
    P.resid = op.p[5];
    g->split = I[30] == 2 ? 1 : 0;
    WD_REQUIRE(I[30] == 0 || I[30] == 1 || I[30] == 2, "gemm: planes must be 1 (bf16) or 2 (fp16 hi/lo)");
    const long long a_ps = I[31], b_ps = I[32], c_ps = I[33];
    P.resid_ps = g->split ? I[34] : 0;
    WD_REQUIRE(!g->split || (a_ps > 0 && b_ps > 0), "gemm: split mode needs plane strides for A and B");
    P.acc_scale = op.f[1] != 0.f ? op.f[1] : 1.f;
    P.lblk = I[40] == 2 ? 2 : 1;   // resolved below (needs the tile configuration)
    const int no_comp = I[41];
    // a residual given as fp16 hi/lo planes is stored times kPlaneScale
    if (g->split && P.resid_dtype == 1) P.alpha /= kPlaneScale;

    WD_REQUIRE(P.D0 > 0 && P.D1 > 0 && P.D2 > 0, "gemm: bad dims %d %d %d", P.D0, P.D1, P.D2);
    WD_REQUIRE(P.E0 > 0 && P.E1 > 0 && P.E2 > 0 && P.E0 * P.E1 * P.E2 <= kTileM, "gemm: bad tile %d %d %d", P.E0, P.E1, P.E2);
    WD_REQUIRE(Kc > 0 && Kc % kBlockK == 0, "gemm: Kc=%d must be a positive multiple of 64", Kc);
    WD_REQUIRE(P.ntaps == 1 || P.ntaps == 9, "gemm: ntaps=%d", P.ntaps);
    WD_REQUIRE(P.N > 0 && P.N % 8 == 0, "gemm: N=%d must be a positive multiple of 8", P.N);
    WD_REQUIRE(g->block_n == 64 || g->block_n == 128 || g->block_n == 256, "gemm: block_n=%d", g->block_n);
    WD_REQUIRE(op.p[0] && op.p[1], "gemm: null operand");
    WD_REQUIRE(P.group_cols > 0 && n_groups > 0 && P.group_cols * n_groups >= P.N, "gemm: bad groups");
    const int CH = g->out_f32 ? 32 : 64;
    WD_REQUIRE(P.group_cols % CH == 0 || n_groups == 1, "gemm: group_cols must be a multiple of the chunk width");
    WD_REQUIRE(P.resid_dtype == 0 || (P.resid && P.ld_res % 8 == 0), "gemm: residual needs ld_res %% 8 == 0");
    if (P.epi_mode == 1) {
        WD_REQUIRE(g->block_n == 64 && P.N == 64 && P.bias && op.p[2], "gemm: DFL epilogue needs N == block_n == 64 and bias");
        P.dfl_out = (float*)op.p[2];
    } else {
        WD_REQUIRE(op.p[2], "gemm: null output");
    }

    P.kc_iters = Kc / kBlockK;
    P.nt0 = ceil_div(P.D0, P.E0); P.nt1 = ceil_div(P.D1, P.E1); P.nt2 = ceil_div(P.D2, P.E2);
    P.num_m_tiles = P.nt0 * P.nt1 * P.nt2;
    P.num_n_tiles = ceil_div(P.N, g->block_n);
    P.num_tiles = P.num_m_tiles * P.num_n_tiles;
    P.rows_a = P.E0 * P.E1 * P.E2;
    // pair mode: 256-wide tiles of the fast path; halves the B bytes each SM pulls through L2 (the L2->SM fabric, not
    // HBM or the tensor pipe, is what bounds 128x256 tiles).  I[36] = 1 disables it (A/B measurements).
    // split mode: 128-wide tiles pair up (a UMMA then reads 6 KB of operands from this SM's shared memory instead of 8)
    if (g->split) P.clu = (g->block_n == 128 && P.epi_mode == 0 && P.num_m_tiles >= 2 && I[36] == 0) ? 2 : 1;
    else P.clu = (g->block_n == 256 && P.epi_mode == 0 && P.num_m_tiles >= 2 && I[36] == 0) ? 2 : 1;
    WD_REQUIRE(!(g->split && g->block_n == 256), "gemm: split (fp16 hi/lo) mode has 64- and 128-wide tiles");
    if (g->split) {
        // Accumulator blocks of two k-blocks: an odd number of k-blocks ends with a one-block sum, a block may wrap around the ring.
        if (P.lblk == 2 && P.kc_iters * P.ntaps < 2) P.lblk = 1;
        // Measured on B200 (tools/trunc_probe.py, profiles/trunc_probe_r02.json): the tensor pipe truncates when it adds a
        // k-step's products to the fp32 accumulator, so a TMEM block sum comes out too small by a data- and K-independent
        // relative amount: 8.9e-8 (~1.5 * 2^-24) per 64-wide block, kSplitComp2 for a two-stage block.  The epilogue undoes it.
        // I[41] = 1 turns the compensation off (the probe uses it).
        P.trunc_comp = no_comp ? 0.f : (P.lblk == 1 ? 8.9e-8f : kSplitComp2);
        P.trunc_comp1 = no_comp ? 0.f : 8.9e-8f;
    }
    P.num_pair_tiles = ((P.num_m_tiles + 1) / 2) * P.num_n_tiles;

    // --- A: rank-4 (k, d0, d1, d2), bf16, box (64, E0, E1, E2), 128B swizzle
    {
        WD_REQUIRE(k_valid <= Kc && k_valid % 8 == 0, "gemm: K_valid=%d must be <= Kc and a multiple of 8", k_valid);
        // stride-2 convolution (I[38], I[39] = input width / height): the tensor map walks the input with element strides
        // (1, 2, 2, 1), so a box of (2 E0) x (2 E1) source pixels delivers the E0 x E1 pixels one tap needs: no im2col
        const int in_w = I[38], in_h = I[39];
        P.a_step = (in_w > 0 && in_h > 0) ? 2 : 1;
        WD_REQUIRE(P.a_step == 1 || (P.ntaps == 9 && 2 * P.E0 <= 256 && 2 * P.E1 <= 256), "gemm: stride-2 walk needs a 3x3 tap set and E0, E1 <= 128");
        uint64_t dims[4] = {(uint64_t)k_valid, (uint64_t)(P.a_step == 2 ? in_w : P.D0), (uint64_t)(P.a_step == 2 ? in_h : P.D1), (uint64_t)P.D2};
        uint64_t str[3] = {(uint64_t)sa0 * 2, (uint64_t)sa1 * 2, (uint64_t)sa2 * 2};
        uint32_t box[4] = {kBlockK, (uint32_t)(P.E0 * P.a_step), (uint32_t)(P.E1 * P.a_step), (uint32_t)P.E2};
        uint32_t est[4] = {1u, (uint32_t)P.a_step, (uint32_t)P.a_step, 1u};
        for (int pl = 0; pl < (g->split ? 2 : 1); ++pl)
            if (encode_tmap(&P.tmA[pl], (const __nv_bfloat16*)op.p[0] + pl * a_ps, 2, 4, dims, str, box, true, est)) return -1;
    }
    // --- B: rank-2 (k_total, n)
    {
        WD_REQUIRE(bk_valid <= (long long)Kc * P.ntaps && bk_valid % 8 == 0, "gemm: bad B K extent");
        uint64_t dims[2] = {(uint64_t)bk_valid, (uint64_t)P.N};
        uint64_t str[1] = {(uint64_t)ldb * 2};
        uint32_t box[2] = {kBlockK, (uint32_t)(P.clu == 2 ? g->block_n / 2 : g->block_n)};
        for (int pl = 0; pl < (g->split ? 2 : 1); ++pl)
            if (encode_tmap(&P.tmB[pl], (const __nv_bfloat16*)op.p[1] + pl * b_ps, 2, 2, dims, str, box, true)) return -1;
    }
    // --- C: rank-5 (c, d0, d1, d2, group), box (CH, E0, E1, E2, 1)
    if (P.epi_mode == 0) {
        const int eb = g->out_f32 ? 4 : 2;
        const int cols = n_groups == 1 ? P.N : group_valid;
        uint64_t dims[5] = {(uint64_t)cols, (uint64_t)P.D0, (uint64_t)P.D1, (uint64_t)P.D2, (uint64_t)n_groups};
        uint64_t str[4] = {(uint64_t)sc0 * eb, (uint64_t)sc1 * eb, (uint64_t)sc2 * eb, (uint64_t)(n_groups == 1 ? sc2 * P.D2 : scg) * eb};
        uint32_t box[5] = {(uint32_t)CH, (uint32_t)P.E0, (uint32_t)P.E1, (uint32_t)P.E2, 1};
        if (encode_tmap(&P.tmC[0], op.p[2], eb, 5, dims, str, box, true)) return -1;
        // per-warp stores (I[37] = 1 turns them off): the 32 rows of one TMEM lane quarter must form a sub-brick (w0, w1, w2).
        // In-run A/B on B200: -3.8 % on the GELU and residual epilogues of the ConvNeXt stage-2 GEMMs (no warpgroup barriers).
        P.warp_store = 0;
        P.w0 = P.w1 = P.w2 = 1;
        if (I[37] == 0) {
            if (P.E0 % 32 == 0) { P.w0 = 32; P.warp_store = 1; }
            else if (32 % P.E0 == 0) {
                const int r1 = 32 / P.E0;
                if (P.E1 % r1 == 0) { P.w0 = P.E0; P.w1 = r1; P.warp_store = 1; }
                else if (r1 % P.E1 == 0 && P.E2 % (r1 / P.E1) == 0) { P.w0 = P.E0; P.w1 = P.E1; P.w2 = r1 / P.E1; P.warp_store = 1; }
            }
        }
        if (P.warp_store) {
            uint32_t wbox[5] = {(uint32_t)CH, (uint32_t)P.w0, (uint32_t)P.w1, (uint32_t)P.w2, 1};
            if (encode_tmap(&P.tmCw[0], op.p[2], eb, 5, dims, str, wbox, true)) return -1;
            if (g->split && !g->out_f32)
                if (encode_tmap(&P.tmCw[1], (__nv_bfloat16*)op.p[2] + c_ps, eb, 5, dims, str, wbox, true)) return -1;
        }
        if (g->split && !g->out_f32) {
            WD_REQUIRE(c_ps > 0, "gemm: split mode with a 16-bit output needs a C plane stride");
            if (encode_tmap(&P.tmC[1], (__nv_bfloat16*)op.p[2] + c_ps, eb, 5, dims, str, box, true)) return -1;
        }
        // In-place fp32 residual with alpha == 1 (the ConvNeXt block's x += gamma * pw2(...), mm_backbone.py:122-124): the per-lane
        // residual rows are 32 uncoalesced 128-byte requests per load instruction and cost 12-19 % of the launch; the TMA store
        // adds the tile in the L2 instead.  I[42] = 1 keeps the loads (A/B measurements).
        P.red_store = 0;
        if (g->split && g->out_f32 && P.warp_store && P.resid_dtype == 2 && P.resid == op.p[2] && P.alpha == 1.f && n_groups == 1 && I[42] == 0 &&
            P.ld_res == sc0 && (P.D1 == 1 || sc1 == (long long)P.D0 * sc0) && (P.D2 == 1 || sc2 == (long long)P.D1 * P.D0 * sc0)) {
            P.red_store = 1;
            P.resid_dtype = 0;
        }
        // direct-store fallback of the split kernel
        P.out = op.p[2];
        P.out_ps = c_ps;
        P.sc0 = sc0; P.sc1 = sc1; P.sc2 = sc2; P.scg = n_groups == 1 ? 0 : scg;
        P.cols_valid = cols;
        if (n_groups == 1) P.group_cols = 1 << 30;  // never wrap
    } else {
        P.tmC[0] = P.tmA[0];  // unused, keep a valid descriptor for the prefetch
    }

    const int sms = device_sm_count();
    if (sms <= 0) return -2;
    g->grid = P.num_tiles < sms ? P.num_tiles : sms;
    if (P.clu == 2) {
        const int want = 2 * P.num_pair_tiles;
        g->grid = want < (sms & ~1) ? want : (sms & ~1);
    }
    out = std::move(g);
    return 0;
}

}  // namespace wd
