// epi_split.cuh — fp32-grade epilogue arithmetic of gemm_split.cu on packed fp32 pairs (FFMA2 / FMUL2 / FADD2).
//
// The epilogue warps of the parity-grade GEMM are bound by the FMA pipe (one 3-operand FFMA per 2 cycles per scheduler):
// the library erff / expf / division forms cost ~40 issue slots per element, 20 000 cycles per 128 x 256 tile against
// 12 000 cycles of UMMA.  Everything here works on two columns per instruction and keeps the results within ~1 ulp of the
// fp32 forms the reference evaluates (torch erf-GELU / SiLU, mm_backbone.py:120, yolo_world_pafpn.py:40-68).
#pragma once
#include "epi_math.cuh"
#include <cuda_fp16.h>

namespace wd {

__device__ __forceinline__ uint64_t splat2(float c) { return pk2(c, c); }
// -x on both halves; ptxas folds the negations into the consuming FFMA2's operand modifiers
__device__ __forceinline__ uint64_t neg2(uint64_t x) {
    uint64_t y;
    asm("{.reg .f32 lo, hi; mov.b64 {lo, hi}, %1; neg.f32 lo, lo; neg.f32 hi, hi; mov.b64 %0, {lo, hi};}" : "=l"(y) : "l"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// erf(|x| / sqrt 2) on a pair, and |x| in `t`: ONE formula for the whole range,
//     erf(T / sqrt 2) = 1 - 2^(T P(T)),   T = min(|x|, 5.9),   P = degree-7 minimax fit of log2(erfc(T / sqrt 2)) / T
// (weighted for the error of erf, fitted in float64: 1.6e-8; tools/erf_fit.py - float32 Horner rounding dominates from degree 7 on).  It is the large-argument form of N. Juffa's erff with
// the GELU's 1 / sqrt 2 and the exponential's log2(e) folded into the coefficients, extended down to T = 0: there 2^(.) -> 1 and
// the subtraction loses RELATIVE accuracy of erf, which the GELU does not need - erf is added to 1.  Working on |x| keeps the sign
// out of it: 0.5 x (1 + erf(x / sqrt 2)) = h + |h| erf(|x| / sqrt 2) with h = 0.5 x.  Eight packed FMAs / multiplies and one MUFU.EX2
// per element instead of the two-interval form's fifteen plus compare / select (kept below under WD_GELU_TWO_INTERVAL).
// Against a float64 GELU over [-8, 8] (2 M points, fp32 Horner): max abs error 3.4e-7 (at |x| = 4.5, < 1 ulp of the result), rms
// 5.7e-8 - torch's own fp32 GELU on the same points: 1.2e-6 / 1.4e-7.
#ifndef WD_GELU_TWO_INTERVAL
__device__ __forceinline__ uint64_t erf_gelu_abs2(uint64_t x, uint64_t& t) {
    float x0, x1;
    upk2(x, x0, x1);
    const float a0 = fabsf(x0), a1 = fabsf(x1);
    const uint64_t ta = pk2(a0, a1);
    t = pk2(fminf(a0, 5.9f), fminf(a1, 5.9f));      // beyond 5.9 the result is 1 to fp32 precision (and P is only fitted up to there)
    uint64_t p = fma2(splat2(-2.8348981686576735e-06f), t, splat2(3.937751898774877e-05f));
    p = fma2(p, t, splat2(-0.00018617944442667067f));
    p = fma2(p, t, splat2(-0.00013693823711946607f));
    p = fma2(p, t, splat2(0.007063422352075577f));
    p = fma2(p, t, splat2(-0.052496179938316345f));
    p = fma2(p, t, splat2(-0.4592081904411316f));
    p = fma2(p, t, splat2(-1.1511051654815674f));
    float r0, r1;
    upk2(mul2(p, t), r0, r1);
    t = ta;                                          // the caller's |x|
    return fma2(pk2(ex2_approx(r0), ex2_approx(r1)), splat2(-1.f), splat2(1.f));   // 1 - 2^(T P(T))
}
#else
// N. Juffa's two-interval erff (max error < 1 ulp on each interval), both intervals evaluated and selected per element:
//   |x| >  1.3120145 (|z| > 0.927734375): erf = 1 - 2^(T P(T)), T = |x|;   |x| <= 1.3120145: erf = |x| Q(x^2)
__device__ __forceinline__ uint64_t erf_gelu_abs2(uint64_t x, uint64_t& t) {
    float x0, x1;
    upk2(x, x0, x1);
    const float t0 = fabsf(x0), t1 = fabsf(x1);
    t = pk2(t0, t1);
    const uint64_t s = mul2(x, x);
    uint64_t r = fma2(splat2(-2.204183147114236e-06f), t, splat2(6.910457159392536e-05f));
    const uint64_t u = fma2(splat2(-0.0009905463084578514f), t, splat2(0.008748006075620651f));
    r = fma2(r, s, u);
    r = fma2(r, t, splat2(-0.05446416139602661f));
    r = fma2(r, t, splat2(-0.4579450786113739f));
    r = fma2(r, t, splat2(-1.151449203491211f));
    r = mul2(r, t);
    float r0, r1;
    upk2(r, r0, r1);
    const uint64_t big = fma2(pk2(ex2_approx(r0), ex2_approx(r1)), splat2(-1.f), splat2(1.f));   // 1 - 2^(.)
    uint64_t q = fma2(splat2(-1.3186695468903054e-05f), s, splat2(0.0002205816999776289f));
    q = fma2(q, s, splat2(-0.0023659912403672934f));
    q = fma2(q, s, splat2(0.019943933933973312f));
    q = fma2(q, s, splat2(-0.13298039138317108f));
    q = fma2(q, s, splat2(0.7978845834732056f));
    q = mul2(q, t);
    float b0, b1, q0, q1;
    upk2(big, b0, b1);
    upk2(q, q0, q1);
    return pk2(t0 > 1.3120145f ? b0 : q0, t1 > 1.3120145f ? b1 : q1);
}
#endif

// exact-form activations on a pair, times `osc` (the power-of-two plane scale of a 16-bit output, or 1)
template <int ACT>
__device__ __forceinline__ uint64_t act2_exact(uint64_t x, float osc) {
    if constexpr (ACT == WD_ACT_GELU) {
        // 0.5 x (1 + erf(x / sqrt 2))
        const uint64_t hs = splat2(0.5f * osc);
        uint64_t t;
        const uint64_t e = erf_gelu_abs2(x, t);
        return fma2(mul2(t, hs), e, mul2(x, hs));
    } else if constexpr (ACT == WD_ACT_SILU) {
        // x / (1 + 2^(-x log2 e)): MUFU.EX2 + MUFU.RCP.  The exponential's 2^-22 relative error already bounds the quotient, so the
        // reciprocal (2^-23) is not refined; x -> -inf gives 2^(+big) = inf, 1 / inf = 0, x * 0 = -0: the right limit, no clamp needed.
        float m0, m1;
        upk2(mul2(x, splat2(-1.4426950408889634f)), m0, m1);
        const uint64_t d = add2(pk2(ex2_approx(m0), ex2_approx(m1)), splat2(1.f));
        float d0, d1;
        upk2(d, d0, d1);
        return mul2(mul2(x, splat2(osc)), pk2(rcp_approx(d0), rcp_approx(d1)));
    } else if constexpr (ACT == WD_ACT_RELU) {
        float a, b;
        upk2(x, a, b);
        return pk2(fmaxf(a, 0.f) * osc, fmaxf(b, 0.f) * osc);
    } else {
        return mul2(x, splat2(osc));
    }
}

// a2[j] <- osc * gamma[n] * act(a2[j] * (1 + comp) * s + bias[n]) over NC columns starting at n_base (two columns per element of a2).
// comp undoes the tensor pipe's truncating accumulation (a measured, data-independent shrink of each TMEM block sum); columns
// at or beyond N get no bias / gamma (they are never stored).
template <int NC, int ACT, bool kFull>
__device__ __forceinline__ void split_epi_math_cols(uint64_t* a2, float comp, float s, float osc, const float* __restrict__ bias, const float* __restrict__ gamma,
                                                    int n_base, int N) {
    const uint64_t comp2 = splat2(comp), s2 = splat2(s);
#pragma unroll
    for (int j = 0; j < NC; j += 4) {
        const int n = n_base + j;
        const bool in = kFull || n < N;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias && in) b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
        uint64_t x0 = fma2(a2[j / 2], comp2, a2[j / 2]), x1 = fma2(a2[j / 2 + 1], comp2, a2[j / 2 + 1]);
        x0 = act2_exact<ACT>(fma2(x0, s2, pk2(b4.x, b4.y)), osc);
        x1 = act2_exact<ACT>(fma2(x1, s2, pk2(b4.z, b4.w)), osc);
        if (gamma && in) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + n));
            x0 = mul2(x0, pk2(g4.x, g4.y));
            x1 = mul2(x1, pk2(g4.z, g4.w));
        }
        a2[j / 2] = x0;
        a2[j / 2 + 1] = x1;
    }
}
template <int NC, int ACT>
__device__ __forceinline__ void split_epi_math(uint64_t* a2, float comp, float s, float osc, const float* __restrict__ bias, const float* __restrict__ gamma,
                                               int n_base, int N) {
    if (n_base + NC <= N) split_epi_math_cols<NC, ACT, true>(a2, comp, s, osc, bias, gamma, n_base, N);   // (warp-uniform)
    else split_epi_math_cols<NC, ACT, false>(a2, comp, s, osc, bias, gamma, n_base, N);
}

// fp16 hi / lo planes of a pair (already multiplied by the plane scale): hi saturates at +-65504, lo = fp16(v - hi)
__device__ __forceinline__ uint32_t cvt_h2_sat(uint64_t v) {
    float a, b;
    upk2(v, a, b);
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // first source operand -> upper half
    return r;
}
__device__ __forceinline__ uint64_t h2_to_f2(uint32_t h) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
    return pk2(f.x, f.y);
}

}  // namespace wd
