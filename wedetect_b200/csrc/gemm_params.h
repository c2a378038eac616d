// gemm_params.h — kernel parameter block and host-side op record shared by the two tcgen05 GEMM kernels:
//   gemm_tc.cu     single-plane bf16 operands (the opt-in "fast" mode)
//   gemm_split.cu  fp16 hi/lo operand pairs, three UMMAs per k-step (the default, parity-grade mode)
#pragma once
#include "internal.h"

namespace wd {

constexpr int kNumEpiWG = 2;
constexpr int kNumThreads = 128 + 128 * kNumEpiWG;
constexpr int kTileM = 128;
constexpr int kBlockK = 64;  // 16-bit elements = 128 bytes = one swizzle atom

// Activations written as fp16 hi/lo planes are stored multiplied by this power of two (value = (hi + lo) / scale): it keeps
// the low plane of O(1) activations out of the fp16 subnormal range while leaving headroom below 65504 for outliers (the
// converters clamp).  Weight matrices carry their own per-matrix power of two, folded into the op's `acc_scale`.
constexpr float kPlaneScale = WD_ACT_PLANE_SCALE;
constexpr float kSplitComp2 = 1.55e-7f;   // truncation shrink of a two-stage accumulator block (cross products of both stages first; tools/trunc_probe.py)

struct GemmParams {
    CUtensorMap tmA[2], tmB[2], tmC[2];   // plane 0 (+ plane 1: the low fp16 plane in split mode)
    CUtensorMap tmCw[2];                  // C with a 32-row box: the quarter of the tile one epilogue warp owns (warp_store)
    int warp_store, w0, w1, w2;           // warp_store: per-warp stores enabled; (w0, w1, w2) = the 32-row sub-brick
    int D0, D1, D2, E0, E1, E2, nt0, nt1, nt2;
    int kc_iters, ntaps, tap_w, pad;
    int N, num_m_tiles, num_n_tiles, num_tiles;
    int act, resid_dtype, ld_res, group_cols, epi_mode, rows_a, exact_act;
    int a_step;          // 1, or 2 for a stride-2 tap walk (input pixel = 2 * output pixel + tap offset)
    int clu, num_pair_tiles;   // clu == 2: clusters of two CTAs (adjacent m-blocks of one n-block) drive cta_group::2 UMMAs
    int lblk;            // split mode: 64-wide k-blocks accumulated in TMEM before the partial sum moves to fp32 registers
    float alpha;
    float acc_scale;     // split mode: accumulator * acc_scale = A . W^T in real units (undoes the operands' power-of-two scales)
    float trunc_comp;    // split mode: relative shrink of a TMEM block sum of `lblk` k-blocks caused by the tensor pipe's truncating adds
    float trunc_comp1;   // ... of a trailing one-k-block sum (odd number of k-blocks with lblk = 2), brought to trunc_comp as it is added
    const float* bias;
    const float* gamma;
    const void* resid;
    long long resid_ps;  // plane stride (elements) of an fp16 hi/lo residual, 0 = single plane
    float* dfl_out;
    // direct-store fallback of the split kernel (tiles whose 32-row warp quarters are not sub-bricks): plain global stores
    void* out;
    long long out_ps, sc0, sc1, sc2, scg;
    int cols_valid;
    // fp32 warp stores of an in-place residual with alpha == 1 (C aliases resid): the TMA store adds the staged tile to memory
    // (cp.reduce.async.bulk.tensor .add, one rounding per element like the fused add), no residual loads in the epilogue
    int red_store;
};

struct GemmOp : CompiledOp {
    GemmParams prm;
    int block_n, out_f32, split, grid, smem;
    int launch(cudaStream_t s) override;
};

int launch_gemm_split(const GemmOp& g, cudaStream_t s);   // gemm_split.cu

}  // namespace wd
