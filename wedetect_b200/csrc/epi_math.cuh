// epi_math.cuh — epilogue arithmetic shared by the tcgen05 kernels (gemm_tc.cu, mlp_fused.cu): activations (exact and
// MUFU.TANH forms), packed fp32x2 helpers, bias / activation / LayerScale over one accumulator column chunk.
#pragma once
#include "common.cuh"
#include "../../include/wedetect_b200.h"

namespace wd {

__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Activations.  kFast (bf16 output of the fast path): MUFU.TANH forms, 3-6 instructions per element, error well
// below the bf16 rounding of the result:  silu(x) = h + h*tanh(h), h = x/2;  gelu(x) ~ h + h*tanh(x*(c0+c1*x^2+c2*x^4))
// (coefficients fitted to the erf form, |err| <= 2.5e-5 before the 2^-11 MUFU error).  Otherwise exact forms.
template <int ACT, bool kFast>
__device__ __forceinline__ float act_fn(float x) {
    if constexpr (ACT == WD_ACT_RELU) return fmaxf(x, 0.f);
    if constexpr (ACT == WD_ACT_SILU) {
        if constexpr (kFast) {
            const float h = 0.5f * x;
            return fmaf(h, tanh_approx(h), h);
        } else {
            return x / (1.f + expf(-x));   // exact path (fp32 outputs / precise mode / exact_act)
        }
    }
    if constexpr (ACT == WD_ACT_GELU) {
        if constexpr (kFast) {
            const float s = x * x;
            const float p = fmaf(s, fmaf(s, -3.51516790e-4f, 3.70056460e-2f), 7.97507884e-1f);
            const float h = 0.5f * x;
            return fmaf(h, tanh_approx(x * p), h);
        } else {
            return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
        }
    }
    return x;
}

// packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2): two lanes of fp32 per instruction
__device__ __forceinline__ uint64_t pk2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// fast activations on a pair (same formulas as act_fn<., true>)
template <int ACT>
__device__ __forceinline__ uint64_t act2_fast(uint64_t x) {
    if constexpr (ACT == WD_ACT_GELU) {
        const uint64_t s = mul2(x, x);
        uint64_t p = fma2(s, pk2(-3.51516790e-4f, -3.51516790e-4f), pk2(3.70056460e-2f, 3.70056460e-2f));
        p = fma2(s, p, pk2(7.97507884e-1f, 7.97507884e-1f));
        float u0, u1;
        upk2(mul2(x, p), u0, u1);
        const uint64_t h = mul2(x, pk2(0.5f, 0.5f));
        return fma2(h, pk2(tanh_approx(u0), tanh_approx(u1)), h);
    } else if constexpr (ACT == WD_ACT_SILU) {
        const uint64_t h = mul2(x, pk2(0.5f, 0.5f));
        float h0, h1;
        upk2(h, h0, h1);
        return fma2(h, pk2(tanh_approx(h0), tanh_approx(h1)), h);
    } else {
        return x;
    }
}

// v[j] = gamma[n] * act(v[j] + bias[n]) over one column chunk.
// kMode 0: generic (null / ragged checks per vector); 1: whole chunk inside N, bias, no gamma; 2: whole chunk, bias and gamma.
// The checked form costs ~2x the instructions of the math itself, and the epilogue of a short-K GEMM is what bounds its tile
// time, so the common shapes get the unchecked forms.
template <int CH, int ACT, bool kFast, int kMode>
__device__ __forceinline__ void epi_bias_act(float* v, const float* __restrict__ bias, const float* __restrict__ gamma, int n_base, int N) {
#pragma unroll
    for (int j = 0; j < CH; j += 4) {
        const int n = n_base + j;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (kMode != 0) b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
        else if (bias && n < N) b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
        if constexpr (kFast && (ACT == WD_ACT_GELU || ACT == WD_ACT_SILU)) {
            upk2(act2_fast<ACT>(add2(pk2(v[j + 0], v[j + 1]), pk2(b4.x, b4.y))), v[j + 0], v[j + 1]);
            upk2(act2_fast<ACT>(add2(pk2(v[j + 2], v[j + 3]), pk2(b4.z, b4.w))), v[j + 2], v[j + 3]);
        } else {
            v[j + 0] = act_fn<ACT, kFast>(v[j + 0] + b4.x);
            v[j + 1] = act_fn<ACT, kFast>(v[j + 1] + b4.y);
            v[j + 2] = act_fn<ACT, kFast>(v[j + 2] + b4.z);
            v[j + 3] = act_fn<ACT, kFast>(v[j + 3] + b4.w);
        }
        if constexpr (kMode == 2) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + n));
            v[j + 0] *= g4.x; v[j + 1] *= g4.y; v[j + 2] *= g4.z; v[j + 3] *= g4.w;
        } else if constexpr (kMode == 0) {
            if (gamma && n < N) {
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + n));
                v[j + 0] *= g4.x; v[j + 1] *= g4.y; v[j + 2] *= g4.z; v[j + 3] *= g4.w;
            }
        }
    }
}


}  // namespace wd
