// mlp_fused.cu — ConvNeXt block MLP in ONE kernel for the narrow stage (C = 128, hidden 512):
//
//     x[m, :] += gamma * ( W2 · GELU(W1 · t[m, :] + b1) + b2 )          (mm_backbone.py:117-124, Block.forward)
//
// Why: at C = 128 the two pointwise GEMMs are not tensor bound.  pw1 writes the 4C-wide hidden activation (839 MB per
// layer at bs 32, 160x160) and its GELU epilogue bounds the tile time; pw2 reads it back and runs at the HBM roofline.
// Here the hidden tile never leaves the SM: GEMM1 accumulates in TMEM, the epilogue warps apply bias + GELU and write
// the bf16 tile into shared memory in the K-major 128B-swizzled layout UMMA reads, GEMM2 consumes it from there.
//
// Layout of the work (cta_group::2, the CTA pair owns 256 rows; every tcgen05 instruction is the pair form):
//   * weights are RESIDENT: each CTA keeps its half of W1 (N split: 256 of the 512 hidden units per 256-wide chunk ->
//     128 rows x 128 K per chunk) and of W2 (N split: 64 of the 128 output channels x 512 K) = 128 KB, loaded once;
//   * per 256-row tile: TMA loads the two 128-row halves of t (32 KB per CTA); for each of the two hidden chunks
//     GEMM1 (M 256, N 256, K 128) -> TMEM acc1 -> epilogue (bias, GELU, bf16) -> smem Hs (4 sub-tiles of 64 hidden units
//     = the 4 K-blocks of GEMM2) -> GEMM2 (M 256, N 128, K 256) accumulates into TMEM acc2; final epilogue
//     x + gamma * (acc2 + b2) -> fp32 -> swizzled staging (the Hs bytes) -> per-warp TMA stores;
//   * warp roles: warp 0 TMA producer, warp 1 MMA issuer (pair leader only), warp 2 TMEM allocator, warps 4..19 sixteen
//     epilogue warps = (64-column sub-tile / 32-column output chunk) x (TMEM lane quarter).
// HBM traffic per layer: t 210 MB + x read 419 MB + x write 419 MB (was 2.7 GB for the two GEMMs).
#include "internal.h"
#include "epi_math.cuh"
#include <string.h>

namespace wd {

constexpr int kMC = 128;          // channels
constexpr int kMH = 512;          // hidden units
// Epilogue width: the GELU epilogue is latency bound with two warps per scheduler; sixteen epilogue warps (four per
// scheduler, one 64-column sub-tile / one 32-column output chunk each) hide it.  Register budget: 640 threads are compiled
// for 96 registers; the four service warps drop to 56, the epilogue warps rise to 104.
constexpr int kMEpiWarps = 16;
constexpr int kMThreads = 128 + 32 * kMEpiWarps;
constexpr int kW1Bytes = 4 * 16384;   // [chunk 0..1][k-block 0..1][128 rows x 128 B]
constexpr int kW2Bytes = 8 * 8192;    // [k-block 0..7][64 rows x 128 B]
constexpr int kABytes = 2 * 16384;    // [k-block 0..1][128 rows x 128 B]
constexpr int kHsBytes = 4 * 16384;   // [sub-tile 0..3][128 rows x 128 B]; doubles as the fp32 output staging
constexpr int kMlpSmem = kW1Bytes + kW2Bytes + kABytes + kHsBytes + 1024 + 1024;
static_assert(kMlpSmem <= 232448, "shared memory budget exceeded");

// Biases / LayerScale ride in the kernel parameters (3 KB of the 4 KB parameter space): the epilogue reads them with
// uniform constant-bank loads instead of global loads whose latency sat on every column chunk's critical path.
struct MlpParams {
    CUtensorMap tmA, tmW1, tmW2, tmX;
    float b1[kMH], b2[kMC], gamma[kMC];
    float* x;
    int M, num_pairs;
};
static_assert(sizeof(MlpParams) <= 4096, "kernel parameter space");

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
// arrive on the barrier at `cluster_addr` (any CTA of the cluster) releasing this thread's prior writes cluster-wide
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

__global__ void __launch_bounds__(kMThreads, 1) mlp_fused_kernel(const __grid_constant__ MlpParams p) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW1 = smem;
    uint8_t* sW2 = sW1 + kW1Bytes;
    uint8_t* sA = sW2 + kW2Bytes;
    uint8_t* sHs = sA + kABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sHs + kHsBytes);
    uint64_t* bar_w_full = bars + 0;       // leader: weights of both CTAs landed
    uint64_t* bar_a_full = bars + 1;       // leader: both halves of the A tile landed
    uint64_t* bar_a_empty = bars + 2;      // both: GEMM1 of the tile's last chunk retired -> A may be overwritten
    uint64_t* bar_acc1_full = bars + 3;    // both: a GEMM1 retired
    uint64_t* bar_acc1_empty = bars + 4;   // leader: all epilogue warps of both CTAs have read acc1
    uint64_t* bar_hs_full = bars + 5;      // leader, [4]: sub-tile s of Hs written by both CTAs (8 warps)
    uint64_t* bar_hs_empty = bars + 9;     // both: GEMM2 of chunk 0 retired -> Hs may be overwritten
    uint64_t* bar_acc2_full = bars + 10;   // both: GEMM2 of the tile's last chunk retired
    uint64_t* bar_acc2_empty = bars + 11;  // leader: all epilogue warps of both CTAs have read acc2
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int crank = (int)(blockIdx.x & 1);
    const int t_first = (int)(blockIdx.x >> 1), t_step = (int)(gridDim.x >> 1);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmW1);
        tma_prefetch_desc(&p.tmW2);
        tma_prefetch_desc(&p.tmX);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(bar_w_full, 1);
        mbar_init(bar_a_full, 1);
        mbar_init(bar_a_empty, 1);
        mbar_init(bar_acc1_full, 1);
        mbar_init(bar_acc1_empty, 2 * kMEpiWarps);
        for (int s = 0; s < 4; ++s) mbar_init(&bar_hs_full[s], 8);
        mbar_init(bar_hs_empty, 1);
        mbar_init(bar_acc2_full, 1);
        mbar_init(bar_acc2_empty, 2 * kMEpiWarps);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc_2sm(tmem_ptr_smem, 512);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    __syncthreads();
    __syncwarp();
    cluster_sync_all();   // peer barriers initialised before anything can signal them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t lead = 0;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == 0) {
            // ================= TMA producer =================
            const uint32_t lead_w = mapa_u32(smem_u32(bar_w_full), lead), lead_a = mapa_u32(smem_u32(bar_a_full), lead);
            // weights: static data, loaded before waiting for the previous kernel (they do not depend on it)
            if (elect_one()) {
                if (crank == 0) mbar_arrive_expect_tx(bar_w_full, 2u * (uint32_t)(kW1Bytes + kW2Bytes));
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
                        tma_load_2d_2sm(&p.tmW1, lead_w, sW1 + (j * 2 + kb) * 16384, kb * 64, j * 256 + crank * 128);
#pragma unroll
                for (int kb = 0; kb < 8; ++kb) tma_load_2d_2sm(&p.tmW2, lead_w, sW2 + kb * 8192, kb * 64, crank * 64);
            }
            __syncwarp();
            pdl_wait();
            uint32_t it = 0;
            for (int tile = t_first; tile < p.num_pairs; tile += t_step, ++it) {
                mbar_wait(bar_a_empty, (it & 1) ^ 1);
                __syncwarp();
                if (elect_one()) {
                    if (crank == 0) mbar_arrive_expect_tx(bar_a_full, 2u * (uint32_t)kABytes);
                    const int row0 = (tile * 2 + crank) * 128;
                    tma_load_2d_2sm(&p.tmA, lead_a, sA, 0, row0);
                    tma_load_2d_2sm(&p.tmA, lead_a, sA + 16384, 64, row0);
                }
                __syncwarp();
            }
        } else if (warp == 1) {
            // ================= MMA issuer (pair leader) =================
            pdl_wait();
            if (crank == 0) {
                constexpr uint32_t idesc1 = umma_idesc_bf16(256, 256);   // GEMM1: M 256 (pair), N 256
                constexpr uint32_t idesc2 = umma_idesc_bf16(256, 128);   // GEMM2: M 256 (pair), N 128
                const uint32_t a_lo = (smem_u32(sA) >> 4) & 0x3FFFu, w1_lo = (smem_u32(sW1) >> 4) & 0x3FFFu;
                const uint32_t w2_lo = (smem_u32(sW2) >> 4) & 0x3FFFu, hs_lo = (smem_u32(sHs) >> 4) & 0x3FFFu;
                const uint32_t acc1 = tmem_base, acc2 = tmem_base + 256;
                mbar_wait(bar_w_full, 0);
                uint32_t it = 0, n1 = 0, nh = 0;   // tiles done, GEMM1s issued, Hs fills consumed
                for (int tile = t_first; tile < p.num_pairs; tile += t_step, ++it) {
                    mbar_wait(bar_a_full, it & 1);
                    for (int j = 0; j < 2; ++j, ++n1) {
                        mbar_wait(bar_acc1_empty, (n1 & 1) ^ 1);
                        tc_fence_after();
                        __syncwarp();
                        if (elect_one()) {
#pragma unroll
                            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_bf16_2sm(acc1, umma_desc_from_lo(a_lo + kb * 1024 + 2 * k),
                                                  umma_desc_from_lo(w1_lo + (j * 2 + kb) * 1024 + 2 * k), idesc1, (kb | k) != 0 ? 1u : 0u);
                            umma_commit_2sm_mc(bar_acc1_full, (uint16_t)3);
                            if (j == 1) umma_commit_2sm_mc(bar_a_empty, (uint16_t)3);
                        }
                        __syncwarp();
                        if (j == 0) continue;
                        // both GEMM1s of the tile are queued; now the GEMM2s, chunk by chunk, sub-tile by sub-tile
                        mbar_wait(bar_acc2_empty, (it & 1) ^ 1);
                        tc_fence_after();
                        for (int jj = 0; jj < 2; ++jj, ++nh) {
                            for (int s = 0; s < 4; ++s) {
                                mbar_wait(&bar_hs_full[s], nh & 1);
                                tc_fence_after();
                                __syncwarp();
                                if (elect_one()) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        umma_bf16_2sm(acc2, umma_desc_from_lo(hs_lo + s * 1024 + 2 * k),
                                                      umma_desc_from_lo(w2_lo + (jj * 4 + s) * 512 + 2 * k), idesc2, (jj | s | k) != 0 ? 1u : 0u);
                                    if (s == 3) umma_commit_2sm_mc(jj == 0 ? bar_hs_empty : bar_acc2_full, (uint16_t)3);
                                }
                                __syncwarp();
                            }
                        }
                    }
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ================= epilogue warps: warp = (sub-tile / output chunk `sub`, TMEM lane quarter) =================
        pdl_wait();
        const int sub = (warp - 4) >> 2, quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t acc1_empty_c = mapa_u32(smem_u32(bar_acc1_empty), lead), acc2_empty_c = mapa_u32(smem_u32(bar_acc2_empty), lead);
        const uint32_t hs_full_c = mapa_u32(smem_u32(&bar_hs_full[sub]), lead);
        const uint32_t srow = smem_u32(sHs) + sub * 16384 + r * 128;
        uint32_t it = 0, n1 = 0;
        for (int tile = t_first; tile < p.num_pairs; tile += t_step, ++it) {
            const long long row = (long long)(tile * 2 + crank) * 128 + r;
            const bool row_ok = row < p.M;
            // ---- hidden chunks: acc1 -> bias + GELU -> bf16 -> Hs ----
            for (int j = 0; j < 2; ++j, ++n1) {
                mbar_wait(bar_acc1_full, n1 & 1);
                tc_fence_after();
                // this warp's 64 columns leave TMEM first, so acc1 goes back to the MMA warp before any math
                float v[64];
                tmem_ld_32x32(lane_base + sub * 64, reinterpret_cast<uint32_t*>(v));
                tmem_ld_32x32(lane_base + sub * 64 + 32, reinterpret_cast<uint32_t*>(v + 32));
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc1_empty_c);
                const int n_base = j * 256 + sub * 64;
#pragma unroll
                for (int c = 0; c < 64; c += 2) {
                    const uint64_t b = pk2(p.b1[n_base + c], p.b1[n_base + c + 1]);
                    upk2(act2_fast<WD_ACT_GELU>(add2(pk2(v[c], v[c + 1]), b)), v[c], v[c + 1]);
                }
                if (j == 0) {
                    // Hs is also this warp's output staging of the previous tile: its TMA store must have read it
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
                } else {
                    mbar_wait(bar_hs_empty, it & 1);   // GEMM2 of chunk 0 has consumed Hs (the wait hid behind the math above)
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    uint4 w;
                    w.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
                    w.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
                    w.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
                    w.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((q ^ (r & 7)) << 4)), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
                }
                fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's (async proxy) reads
                __syncwarp();
                if (lane == 0) mbar_arrive_release_cluster(hs_full_c);
            }
            // ---- output chunk `sub` (32 fp32 columns): x + gamma * (acc2 + b2) ----
            uint4 rq[8];
            {
                const float* rp = p.x + row * kMC + sub * 32;
#pragma unroll
                for (int q = 0; q < 8; ++q) rq[q] = row_ok ? *reinterpret_cast<const uint4*>(rp + q * 4) : make_uint4(0u, 0u, 0u, 0u);
            }
            mbar_wait(bar_acc2_full, it & 1);
            tc_fence_after();
            float o[32];
            tmem_ld_32x32(lane_base + 256 + sub * 32, reinterpret_cast<uint32_t*>(o));
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc2_empty_c);
            // acc2_full means GEMM2 of the last chunk has retired: Hs is free and becomes the output staging
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int n = sub * 32 + q * 4;
                const float y0 = fmaf(p.gamma[n + 0], o[q * 4 + 0] + p.b2[n + 0], __uint_as_float(rq[q].x));
                const float y1 = fmaf(p.gamma[n + 1], o[q * 4 + 1] + p.b2[n + 1], __uint_as_float(rq[q].y));
                const float y2 = fmaf(p.gamma[n + 2], o[q * 4 + 2] + p.b2[n + 2], __uint_as_float(rq[q].z));
                const float y3 = fmaf(p.gamma[n + 3], o[q * 4 + 3] + p.b2[n + 3], __uint_as_float(rq[q].w));
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((q ^ (r & 7)) << 4)), "f"(y0), "f"(y1), "f"(y2), "f"(y3) : "memory");
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&p.tmX, sHs + sub * 16384 + quarter * 4096, sub * 32, (tile * 2 + crank) * 128 + quarter * 32);
                tma_store_commit();
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    __syncwarp();
    cluster_sync_all();   // the peer may still be arriving on this CTA's barriers / reading its shared memory
    if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int compile_mlp_fused(const wd_op& op, std::unique_ptr<CompiledOp>& out) {
    const int32_t* I = op.i;
    const int M = I[0], C = I[1], H = I[2];
    WD_REQUIRE(C == kMC && H == kMH, "mlp_fused: only C = %d, hidden = %d is implemented (got %d, %d)", kMC, kMH, C, H);
    WD_REQUIRE(M > 0, "mlp_fused: bad M");
    for (int k = 0; k <= 6; ++k) WD_REQUIRE(op.p[k], "mlp_fused: null pointer %d", k);
    const int sms = device_sm_count();
    if (sms <= 0) return -2;
    struct MlpOp : CompiledOp {
        MlpParams prm;
        int grid;
        int launch(cudaStream_t s) override {
            WD_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(mlp_fused_kernel), kMlpSmem));
            WD_CHECK_CUDA(launch_pdl(mlp_fused_kernel, dim3(grid), dim3(kMThreads), (size_t)kMlpSmem, s, 2, prm));
            WD_CHECK_CUDA(cudaGetLastError());
            count_launch();
            return 0;
        }
    };
    auto m = std::make_unique<MlpOp>();
    MlpParams& P = m->prm;
    memset(&P, 0, sizeof(P));
    P.M = M;
    P.num_pairs = (M + 255) / 256;
    // biases / LayerScale become kernel parameters: read them back once at compile time (synchronous, plan build only)
    WD_CHECK_CUDA(cudaMemcpy(P.b1, op.p[3], sizeof(P.b1), cudaMemcpyDeviceToHost));
    WD_CHECK_CUDA(cudaMemcpy(P.b2, op.p[4], sizeof(P.b2), cudaMemcpyDeviceToHost));
    WD_CHECK_CUDA(cudaMemcpy(P.gamma, op.p[5], sizeof(P.gamma), cudaMemcpyDeviceToHost));
    P.x = (float*)op.p[6];
    {   // t: bf16 [M, C], box (64 k, 128 rows)
        const uint64_t dims[2] = {(uint64_t)C, (uint64_t)M};
        const uint64_t str[1] = {(uint64_t)(I[3] > 0 ? I[3] : C) * 2};
        const uint32_t box[2] = {64u, 128u};
        if (encode_tmap(&P.tmA, op.p[0], 2, 2, dims, str, box, true)) return -1;
    }
    {   // W1: bf16 [H, C], box (64 k, 128 hidden units)
        const uint64_t dims[2] = {(uint64_t)C, (uint64_t)H};
        const uint64_t str[1] = {(uint64_t)C * 2};
        const uint32_t box[2] = {64u, 128u};
        if (encode_tmap(&P.tmW1, op.p[1], 2, 2, dims, str, box, true)) return -1;
    }
    {   // W2: bf16 [C, H], box (64 k, 64 output channels)
        const uint64_t dims[2] = {(uint64_t)H, (uint64_t)C};
        const uint64_t str[1] = {(uint64_t)H * 2};
        const uint32_t box[2] = {64u, 64u};
        if (encode_tmap(&P.tmW2, op.p[2], 2, 2, dims, str, box, true)) return -1;
    }
    {   // x: fp32 [M, C], store box (32 cols, 32 rows)
        const uint64_t dims[2] = {(uint64_t)C, (uint64_t)M};
        const uint64_t str[1] = {(uint64_t)C * 4};
        const uint32_t box[2] = {32u, 32u};
        if (encode_tmap(&P.tmX, op.p[6], 4, 2, dims, str, box, true)) return -1;
    }
    const int want = 2 * P.num_pairs;
    m->grid = want < (sms & ~1) ? want : (sms & ~1);
    out = std::move(m);
    return 0;
}

}  // namespace wd
