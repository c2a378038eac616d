// rowops.cu — the HBM-/CUDA-core-bound kernels around the tensor-core GEMMs:
//   LayerNorm rows (+ space-to-depth write), depthwise 7x7 + LayerNorm, stem patch gather,
//   stride-2 im2col, casts, XLM-R embedding / short attention / pooling, contrastive-head folding,
//   kept-proposal embedding gather.
// All activations are NHWC ("rows" = pixels, columns = channels), loads/stores are 8/16-byte vectors,
// coalesced along channels.  `*_lo` outputs are the low bf16 plane of the bf16x3 precise mode.
#include "gemm_params.h"
#include <cuda_fp16.h>
#include <functional>
#include <algorithm>
#include <math.h>
#include <stdlib.h>

namespace wd {

struct FnOp : CompiledOp {
    std::function<int(cudaStream_t)> fn;
    int launch(cudaStream_t s) override { return fn(s); }
};

// Store fp32 values as a GEMM operand.  ps == 0: one bf16 plane (fast mode).  ps > 0: the parity-grade format of
// gemm_split.cu, two fp16 planes hi + lo = value * sc (plane stride ps elements; sc = kPlaneScale for activations).
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float sat_f16(float x) { return fminf(fmaxf(x, -65504.f), 65504.f); }
__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* hi, long long ps, long long idx, float a, float b, float c, float d, float sc = kPlaneScale) {
    uint2 w;
    if (ps) {
        a = sat_f16(a * sc); b = sat_f16(b * sc); c = sat_f16(c * sc); d = sat_f16(d * sc);
        w.x = pack_f16x2(a, b);
        w.y = pack_f16x2(c, d);
        *reinterpret_cast<uint2*>(hi + idx) = w;
        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&w.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
        w.x = pack_f16x2(a - f0.x, b - f0.y);
        w.y = pack_f16x2(c - f1.x, d - f1.y);
        *reinterpret_cast<uint2*>(hi + ps + idx) = w;
    } else {
        w.x = pack_bf16x2(a, b);
        w.y = pack_bf16x2(c, d);
        *reinterpret_cast<uint2*>(hi + idx) = w;
    }
}
__device__ __forceinline__ void store_bf16x1(__nv_bfloat16* hi, long long ps, long long idx, float a, float sc = kPlaneScale) {
    if (ps) {
        a = sat_f16(a * sc);
        const __half h = __float2half_rn(a);
        reinterpret_cast<__half*>(hi)[idx] = h;
        reinterpret_cast<__half*>(hi)[ps + idx] = __float2half_rn(a - __half2float(h));
    } else {
        hi[idx] = __float2bfloat16_rn(a);
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over rows (one warp per row).  mm_backbone.py:145-155 / F.layer_norm semantics:
// u = mean(x); s = mean((x-u)^2); y = (x-u)/sqrt(s+eps)*w + b
// ------------------------------------------------------------------------------------------------
constexpr int kLnMaxVec = 12;  // C <= 1536

__global__ void __launch_bounds__(256) ln_rows_kernel(const float* __restrict__ in, int rows, int C, int ld_in, const float* __restrict__ w,
                                                      const float* __restrict__ b, float eps, __nv_bfloat16* out_hi, long long out_ps,
                                                      float* out_f32, int ld_out, int s2d, int W, int H) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* x = in + (long long)row * ld_in;
    float4 v[kLnMaxVec];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < C) {
            v[i] = *reinterpret_cast<const float4*>(x + c);
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float mean = warp_sum(sum) / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < C) {
            const float a = v[i].x - mean, bb = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
            sq += (a * a + bb * bb) + (cc * cc + d * d);
        }
    }
    const float rstd = 1.f / sqrtf(warp_sum(sq) / (float)C + eps);
    long long orow = row;
    int coff = 0;
    if (s2d) {
        const int xx = row % W, yy = (row / W) % H, bi = row / (W * H);
        orow = ((long long)bi * (H / 2) + yy / 2) * (W / 2) + xx / 2;
        coff = ((yy & 1) * 2 + (xx & 1)) * C;
    }
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < C) {
            const float4 ww = __ldg(reinterpret_cast<const float4*>(w + c));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
            const float y0 = (v[i].x - mean) * rstd * ww.x + bb.x;
            const float y1 = (v[i].y - mean) * rstd * ww.y + bb.y;
            const float y2 = (v[i].z - mean) * rstd * ww.z + bb.z;
            const float y3 = (v[i].w - mean) * rstd * ww.w + bb.w;
            if (out_hi) store_bf16x4(out_hi, out_ps, orow * ld_out + coff + c, y0, y1, y2, y3);
            if (out_f32) *reinterpret_cast<float4*>(out_f32 + (long long)row * C + c) = make_float4(y0, y1, y2, y3);
        }
    }
}

// Narrow rows (C <= 128*NV): R rows per warp in flight so that each warp keeps R*NV 16-byte loads outstanding
// (the one-row kernel above is latency bound at C = 128: 512 B per warp per round trip).
template <int NV, int R>
__global__ void __launch_bounds__(256) ln_rows_multi_kernel(const float* __restrict__ in, int rows, int C, int ld_in, const float* __restrict__ w,
                                                            const float* __restrict__ b, float eps, __nv_bfloat16* out_hi, long long out_ps,
                                                            float* out_f32, int ld_out, int s2d, int W, int H) {
    pdl_launch_dependents();
    pdl_wait();
    const int wrp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const long long row0 = (long long)wrp * R;
    if (row0 >= rows) return;
    float4 v[R][NV];
    float sum[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        sum[r] = 0.f;
        const long long row = row0 + r;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            v[r][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < rows && c < C) v[r][i] = *reinterpret_cast<const float4*>(in + row * ld_in + c);
            sum[r] += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
        }
    }
    float mean[R], rstd[R];
#pragma unroll
    for (int r = 0; r < R; ++r) mean[r] = warp_sum(sum[r]) / (float)C;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < C) {
                const float a = v[r][i].x - mean[r], bb = v[r][i].y - mean[r], cc = v[r][i].z - mean[r], d = v[r][i].w - mean[r];
                sq += (a * a + bb * bb) + (cc * cc + d * d);
            }
        }
        rstd[r] = 1.f / sqrtf(warp_sum(sq) / (float)C + eps);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c >= C) continue;
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w + c));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long row = row0 + r;
            if (row >= rows) continue;
            const float y0 = (v[r][i].x - mean[r]) * rstd[r] * ww.x + bb.x, y1 = (v[r][i].y - mean[r]) * rstd[r] * ww.y + bb.y;
            const float y2 = (v[r][i].z - mean[r]) * rstd[r] * ww.z + bb.z, y3 = (v[r][i].w - mean[r]) * rstd[r] * ww.w + bb.w;
            if (out_hi) {
                long long orow = row;
                int coff = 0;
                if (s2d) {
                    const int xx = (int)(row % W), yy = (int)((row / W) % H), bi = (int)(row / ((long long)W * H));
                    orow = ((long long)bi * (H / 2) + yy / 2) * (W / 2) + xx / 2;
                    coff = ((yy & 1) * 2 + (xx & 1)) * C;
                }
                store_bf16x4(out_hi, out_ps, orow * ld_out + coff + c, y0, y1, y2, y3);
            }
            if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * C + c) = make_float4(y0, y1, y2, y3);
        }
    }
}

// Narrow rows (C <= 256): a row is shared by LPR lanes holding NV float4 each, a warp works on 32 / LPR rows at a time.
// The one-warp-per-row mapping spends ~65 instructions per 16 bytes on the two 5-step warp reductions (ncu: 2.9 IPC, 48 %
// of DRAM bandwidth at C = 128: issue bound); here a reduction is log2(LPR) steps over 4x the data per lane, the weights
// stay in registers across the grid-stride loop over rows, and the kernel goes back to being HBM bound.
template <int LPR, int NV>
__global__ void __launch_bounds__(256) ln_rows_group_kernel(const float* __restrict__ in, int rows, int C, int ld_in, const float* __restrict__ w,
                                                            const float* __restrict__ b, float eps, __nv_bfloat16* out_hi, long long out_ps,
                                                            float* out_f32, int ld_out, int s2d, int W, int H) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane % LPR, rw = lane / LPR;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (long long)gridDim.x * (blockDim.x >> 5);
    float4 ww[NV], bb[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * LPR + sub) * 4;
        ww[i] = bb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < C) {
            ww[i] = __ldg(reinterpret_cast<const float4*>(w + c));
            bb[i] = __ldg(reinterpret_cast<const float4*>(b + c));
        }
    }
    const float inv_c = 1.f / (float)C;
    for (long long base = gw * RPW; base < rows; base += nw * RPW) {   // warp-uniform bound: every lane reaches the shuffles
        const long long row = base + rw;
        const bool ok = row < rows;
        float4 v[NV];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * LPR + sub) * 4;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok && c < C) v[i] = *reinterpret_cast<const float4*>(in + row * ld_in + c);
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum * inv_c;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * LPR + sub) * 4;
            if (c < C) {
                const float a = v[i].x - mean, bq = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
                sq += (a * a + bq * bq) + (cc * cc + d * d);
            }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = 1.f / sqrtf(sq * inv_c + eps);
        if (!ok) continue;
        long long orow = row;
        int coff = 0;
        if (s2d) {
            const int xx = (int)(row % W), yy = (int)((row / W) % H), bi = (int)(row / ((long long)W * H));
            orow = ((long long)bi * (H / 2) + yy / 2) * (W / 2) + xx / 2;
            coff = ((yy & 1) * 2 + (xx & 1)) * C;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * LPR + sub) * 4;
            if (c >= C) continue;
            const float y0 = (v[i].x - mean) * rstd * ww[i].x + bb[i].x, y1 = (v[i].y - mean) * rstd * ww[i].y + bb[i].y;
            const float y2 = (v[i].z - mean) * rstd * ww[i].z + bb[i].z, y3 = (v[i].w - mean) * rstd * ww[i].w + bb[i].w;
            if (out_hi) store_bf16x4(out_hi, out_ps, orow * ld_out + coff + c, y0, y1, y2, y3);
            if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * C + c) = make_float4(y0, y1, y2, y3);
        }
    }
}

static void launch_ln_rows(cudaStream_t s, const float* in, int rows, int C, int ld_in, const float* w, const float* b, float eps, __nv_bfloat16* oh,
                           long long ol, float* of, int ld_out, int s2d, int W, int H) {
    static const bool grouped = getenv("WD_LN_NO_GROUP") == nullptr;   // A/B switch
    if (grouped && C <= 256) {
        // 8 lanes x 4 float4 (C <= 128) or 16 lanes x 4 float4 (C <= 256) per row; grid-stride so the weights load once per warp
        const int rpw = C <= 128 ? 4 : 2;
        const long long groups = ((long long)rows + rpw - 1) / rpw;
        const int blocks = (int)std::min<long long>((groups + 7) / 8, (long long)device_sm_count() * 16);
        if (C <= 128) launch_pdl(ln_rows_group_kernel<8, 4>, dim3(blocks), dim3(256), (size_t)(0), s, 1, in, rows, C, ld_in, w, b, eps, oh, ol, of, ld_out, s2d, W, H);
        else launch_pdl(ln_rows_group_kernel<16, 4>, dim3(blocks), dim3(256), (size_t)(0), s, 1, in, rows, C, ld_in, w, b, eps, oh, ol, of, ld_out, s2d, W, H);
    } else if (C <= 128) {
        const int warps = (rows + 3) / 4;
        launch_pdl(ln_rows_multi_kernel<1, 4>, dim3((warps + 7) / 8), dim3(256), (size_t)(0), s, 1, in, rows, C, ld_in, w, b, eps, oh, ol, of, ld_out, s2d, W, H);
    } else if (C <= 256) {
        const int warps = (rows + 3) / 4;
        launch_pdl(ln_rows_multi_kernel<2, 4>, dim3((warps + 7) / 8), dim3(256), (size_t)(0), s, 1, in, rows, C, ld_in, w, b, eps, oh, ol, of, ld_out, s2d, W, H);
    } else if (C <= 512) {
        const int warps = (rows + 1) / 2;
        launch_pdl(ln_rows_multi_kernel<4, 2>, dim3((warps + 7) / 8), dim3(256), (size_t)(0), s, 1, in, rows, C, ld_in, w, b, eps, oh, ol, of, ld_out, s2d, W, H);
    } else {
        launch_pdl(ln_rows_kernel, dim3((rows + 7) / 8), dim3(256), (size_t)(0), s, 1, in, rows, C, ld_in, w, b, eps, oh, ol, of, ld_out, s2d, W, H);
    }
}

// ------------------------------------------------------------------------------------------------
// depthwise 7x7 (pad 3) + bias + LayerNorm(C) -> bf16.   mm_backbone.py:114-116.
// block = C/4 threads (rounded up to a warp multiple); a block owns a strip of TY x TX output pixels
// and ALL channels of them (needed for the per-pixel LayerNorm); a thread owns 4 channels.
// ------------------------------------------------------------------------------------------------
constexpr int kDwTX = 8, kDwTY = 2;

__global__ void __launch_bounds__(384) dwconv7_ln_kernel(const float* __restrict__ in, int B, int H, int W, int C, const float* __restrict__ wt,
                                                         const float* __restrict__ bias, const float* __restrict__ lnw, const float* __restrict__ lnb,
                                                         float eps, __nv_bfloat16* out_hi, long long out_ps, int ld_out) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[kDwTY * kDwTX][12];
    const int cg = threadIdx.x;
    const int c = cg * 4;
    const bool active = c < C;
    const int x0 = blockIdx.x * kDwTX, y0 = blockIdx.y * kDwTY, bi = blockIdx.z;
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    float4 acc[kDwTY][kDwTX];
#pragma unroll
    for (int oy = 0; oy < kDwTY; ++oy)
#pragma unroll
        for (int ox = 0; ox < kDwTX; ++ox) acc[oy][ox] = make_float4(0.f, 0.f, 0.f, 0.f);

    if (active) {
        const float* inb = in + (long long)bi * H * W * C + c;
#pragma unroll 1
        for (int iy = 0; iy < kDwTY + 6; ++iy) {
            const int yin = y0 - 3 + iy;
            if (yin < 0 || yin >= H) continue;
            float4 wrow[kDwTY][7];
#pragma unroll
            for (int oy = 0; oy < kDwTY; ++oy) {
                const int dy = iy - oy;
#pragma unroll
                for (int dx = 0; dx < 7; ++dx)
                    wrow[oy][dx] = (dy >= 0 && dy < 7) ? __ldg(reinterpret_cast<const float4*>(wt + (long long)(dy * 7 + dx) * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const float* inr = inb + (long long)yin * W * C;
#pragma unroll
            for (int ix = 0; ix < kDwTX + 6; ++ix) {
                const int xin = x0 - 3 + ix;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (xin >= 0 && xin < W) v = __ldg(reinterpret_cast<const float4*>(inr + (long long)xin * C));
#pragma unroll
                for (int oy = 0; oy < kDwTY; ++oy) {
#pragma unroll
                    for (int ox = 0; ox < kDwTX; ++ox) {
                        const int dx = ix - ox;
                        if (dx >= 0 && dx < 7) {
                            acc[oy][ox].x = fmaf(v.x, wrow[oy][dx].x, acc[oy][ox].x);
                            acc[oy][ox].y = fmaf(v.y, wrow[oy][dx].y, acc[oy][ox].y);
                            acc[oy][ox].z = fmaf(v.z, wrow[oy][dx].z, acc[oy][ox].z);
                            acc[oy][ox].w = fmaf(v.w, wrow[oy][dx].w, acc[oy][ox].w);
                        }
                    }
                }
            }
        }
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c));
#pragma unroll
        for (int oy = 0; oy < kDwTY; ++oy)
#pragma unroll
            for (int ox = 0; ox < kDwTX; ++ox) {
                acc[oy][ox].x += b4.x; acc[oy][ox].y += b4.y; acc[oy][ox].z += b4.z; acc[oy][ox].w += b4.w;
            }
    }

    // ---- LayerNorm over channels: two block reductions (mean, then centred second moment) ----
    float mean[kDwTY][kDwTX], rstd[kDwTY][kDwTX];
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int oy = 0; oy < kDwTY; ++oy)
#pragma unroll
            for (int ox = 0; ox < kDwTX; ++ox) {
                float s = 0.f;
                if (active) {
                    const float4 a = acc[oy][ox];
                    if (pass == 0) s = (a.x + a.y) + (a.z + a.w);
                    else {
                        const float m = mean[oy][ox];
                        s = ((a.x - m) * (a.x - m) + (a.y - m) * (a.y - m)) + ((a.z - m) * (a.z - m) + (a.w - m) * (a.w - m));
                    }
                }
                s = warp_sum(s);
                if (lane == 0) red[oy * kDwTX + ox][warp] = s;
            }
        __syncthreads();
#pragma unroll
        for (int oy = 0; oy < kDwTY; ++oy)
#pragma unroll
            for (int ox = 0; ox < kDwTX; ++ox) {
                float t = 0.f;
                for (int w2 = 0; w2 < nwarps; ++w2) t += red[oy * kDwTX + ox][w2];
                if (pass == 0) mean[oy][ox] = t / (float)C;
                else rstd[oy][ox] = 1.f / sqrtf(t / (float)C + eps);
            }
        __syncthreads();
    }
    if (!active) return;
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(lnw + c));
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(lnb + c));
#pragma unroll
    for (int oy = 0; oy < kDwTY; ++oy) {
        const int y = y0 + oy;
        if (y >= H) continue;
#pragma unroll
        for (int ox = 0; ox < kDwTX; ++ox) {
            const int x = x0 + ox;
            if (x >= W) continue;
            const float m = mean[oy][ox], r = rstd[oy][ox];
            const float4 a = acc[oy][ox];
            const long long idx = (((long long)bi * H + y) * W + x) * ld_out + c;
            store_bf16x4(out_hi, out_ps, idx, (a.x - m) * r * w4.x + g4.x, (a.y - m) * r * w4.y + g4.y, (a.z - m) * r * w4.z + g4.z,
                         (a.w - m) * r * w4.w + g4.w);
        }
    }
}

constexpr int kDw2CK = 32;

__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float unpack_lo(uint64_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float unpack_hi(uint64_t v) { return __uint_as_float((uint32_t)(v >> 32)); }

// ------------------------------------------------------------------------------------------------
// stem patch gather: NCHW image -> rows [B*(H/4)*(W/4), 64], k = c*16 + dy*4 + dx (48 valid, 16 zero)
// ------------------------------------------------------------------------------------------------
template <typename InT>
__global__ void stem_patch_kernel(const InT* __restrict__ in, int B, int H, int W, float scale, __nv_bfloat16* out_hi, long long out_ps) {
    pdl_launch_dependents();
    pdl_wait();
    const int Wo = W / 4, Ho = H / 4;
    const long long total = (long long)B * Ho * Wo * 16;  // 16 groups of 4 k-values per row
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int grp = (int)(t & 15);
        const long long m = t >> 4;
        const int px = (int)(m % Wo), py = (int)((m / Wo) % Ho), bi = (int)(m / ((long long)Wo * Ho));
        float a = 0.f, b = 0.f, c = 0.f, d = 0.f;
        if (grp < 12) {
            const int ch = grp >> 2, dy = grp & 3;
            const InT* src = in + (((long long)bi * 3 + ch) * H + (py * 4 + dy)) * W + px * 4;
            a = (float)src[0] * scale; b = (float)src[1] * scale; c = (float)src[2] * scale; d = (float)src[3] * scale;
        }
        store_bf16x4(out_hi, out_ps, m * 64 + grp * 4, a, b, c, d);
    }
}

// ------------------------------------------------------------------------------------------------
// im2col for 3x3 stride-2 pad-1 conv, bf16 NHWC -> rows [B*Ho*Wo, 9*C], k = (ky*3+kx)*C + c
// ------------------------------------------------------------------------------------------------
__global__ void im2col_s2_kernel(const __nv_bfloat16* __restrict__ in, long long in_ps, int B, int H, int W, int C, int ld_in,
                                 __nv_bfloat16* out, long long out_ps) {
    pdl_launch_dependents();
    pdl_wait();
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const int vec = C / 8;
    const long long total = (long long)B * Ho * Wo * 9 * vec;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(t % vec);
        long long r = t / vec;
        const int tap = (int)(r % 9);
        r /= 9;
        const int ox = (int)(r % Wo), oy = (int)((r / Wo) % Ho), bi = (int)(r / ((long long)Wo * Ho));
        const int iy = oy * 2 + tap / 3 - 1, ix = ox * 2 + tap % 3 - 1;
        const bool inb = iy >= 0 && iy < H && ix >= 0 && ix < W;
        const long long src = (((long long)bi * H + iy) * W + ix) * ld_in + cv * 8;
        const long long dst = r * (9LL * C) + (long long)tap * C + cv * 8;
        for (int pl = 0; pl < (out_ps ? 2 : 1); ++pl) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (inb) v = *reinterpret_cast<const uint4*>(in + pl * in_ps + src);
            *reinterpret_cast<uint4*>(out + pl * out_ps + dst) = v;
        }
    }
}

__global__ void cast_bf16_kernel(const float* __restrict__ in, long long rows, int C, int ld_in, int ld_out, __nv_bfloat16* out, long long out_ps) {
    pdl_launch_dependents();
    pdl_wait();
    const int vec = C / 4;
    const long long total = rows * vec;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / vec;
        const int c = (int)(t % vec) * 4;
        const float4 v = *reinterpret_cast<const float4*>(in + r * ld_in + c);
        store_bf16x4(out, out_ps, r * ld_out + c, v.x, v.y, v.z, v.w);
    }
}

// ------------------------------------------------------------------------------------------------
// XLM-R embeddings + LayerNorm (one warp per token)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) text_embed_kernel(const int* __restrict__ ids, int S, int L, int Hd, int pad_idx, const float* __restrict__ word,
                                                         const float* __restrict__ pos, const float* __restrict__ type, const float* __restrict__ lnw,
                                                         const float* __restrict__ lnb, float eps, float* out_f32, __nv_bfloat16* out_hi,
                                                         long long out_ps) {
    const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tok >= S * L) return;
    const int lane = threadIdx.x & 31;
    const int s = tok / L, l = tok % L;
    const int id = ids[tok];
    int pos_id = pad_idx;
    if (id != pad_idx) {
        int cnt = 0;
        for (int j = 0; j <= l; ++j) cnt += (ids[s * L + j] != pad_idx) ? 1 : 0;
        pos_id = cnt + pad_idx;
    }
    const float* wr = word + (long long)id * Hd;
    const float* pr = pos + (long long)pos_id * Hd;
    float4 v[8];  // Hd <= 1024
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < Hd) {
            const float4 a = *reinterpret_cast<const float4*>(wr + c);
            const float4 t = *reinterpret_cast<const float4*>(type + c);
            const float4 p4 = *reinterpret_cast<const float4*>(pr + c);
            // HF order: inputs_embeds + token_type_embeddings, then + position_embeddings
            v[i] = make_float4((a.x + t.x) + p4.x, (a.y + t.y) + p4.y, (a.z + t.z) + p4.z, (a.w + t.w) + p4.w);
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float mean = warp_sum(sum) / (float)Hd;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < Hd) {
            const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
            sq += (a * a + b * b) + (cc * cc + d * d);
        }
    }
    const float rstd = 1.f / sqrtf(warp_sum(sq) / (float)Hd + eps);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < Hd) {
            const float4 ww = *reinterpret_cast<const float4*>(lnw + c);
            const float4 bb = *reinterpret_cast<const float4*>(lnb + c);
            const float y0 = (v[i].x - mean) * rstd * ww.x + bb.x, y1 = (v[i].y - mean) * rstd * ww.y + bb.y;
            const float y2 = (v[i].z - mean) * rstd * ww.z + bb.z, y3 = (v[i].w - mean) * rstd * ww.w + bb.w;
            const long long idx = (long long)tok * Hd + c;
            *reinterpret_cast<float4*>(out_f32 + idx) = make_float4(y0, y1, y2, y3);
            store_bf16x4(out_hi, out_ps, idx, y0, y1, y2, y3);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// short-sequence attention: one warp per (sequence, head); L <= 32, head_dim == 64.
// softmax(q k^T * scale + mask) v, masked keys get -inf (== HF additive float-min mask after softmax)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_small_kernel(const float* __restrict__ qkv, const int* __restrict__ mask, int S, int L, int heads, int ld,
                                                         float scale, __nv_bfloat16* out_hi, long long out_ps) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = blockIdx.x * 4 + warp;
    if (pair >= S * heads) return;
    const int s = pair / heads, h = pair % heads;
    const int Hd = heads * 64;
    float* Q = sm + warp * (3 * L * 65);
    float* K = Q + L * 65;
    float* V = K + L * 65;
    for (int t = lane; t < L * 64; t += 32) {
        const int j = t >> 6, d = t & 63;
        const float* row = qkv + (long long)(s * L + j) * ld + h * 64 + d;
        Q[j * 65 + d] = row[0];
        K[j * 65 + d] = row[Hd];
        V[j * 65 + d] = row[2 * Hd];
    }
    __syncwarp();
    const bool key_ok = lane < L && mask[s * L + (lane < L ? lane : 0)] != 0;
    for (int i = 0; i < L; ++i) {
        float sc = -INFINITY;
        if (lane < L) {
            float a = 0.f;
#pragma unroll 16
            for (int d = 0; d < 64; ++d) a = fmaf(Q[i * 65 + d], K[lane * 65 + d], a);
            sc = key_ok ? a * scale : -INFINITY;
        }
        const float mx = warp_max(sc);
        const float e = (lane < L && key_ok) ? expf(sc - mx) : 0.f;
        const float den = warp_sum(e);
        const float pj = e / den;
        float o0 = 0.f, o1 = 0.f;
        for (int j = 0; j < L; ++j) {
            const float pp = __shfl_sync(0xffffffffu, pj, j);
            o0 = fmaf(pp, V[j * 65 + lane], o0);
            o1 = fmaf(pp, V[j * 65 + lane + 32], o1);
        }
        const long long idx = (long long)(s * L + i) * Hd + h * 64;
        store_bf16x1(out_hi, out_ps, idx + lane, o0);
        store_bf16x1(out_hi, out_ps, idx + lane + 32, o1);
    }
}

// Longer prompts (32 < L <= 128): one block of 4 warps per (sequence, head); K and V of the head stay in shared memory, a warp
// owns queries warp, warp + 4, ...; a lane scores keys lane, lane + 32, ... (same arithmetic as above, key by key).
__global__ void __launch_bounds__(128) attn_mid_kernel(const float* __restrict__ qkv, const int* __restrict__ mask, int S, int L, int heads, int ld,
                                                       float scale, __nv_bfloat16* out_hi, long long out_ps) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x / heads, h = blockIdx.x % heads;
    const int Hd = heads * 64;
    float* K = sm;
    float* V = K + L * 65;
    float* Q = V + L * 65;      // 4 x 64: the query row each warp is working on
    for (int t = threadIdx.x; t < L * 64; t += blockDim.x) {
        const int j = t >> 6, d = t & 63;
        const float* row = qkv + (long long)(s * L + j) * ld + h * 64 + d;
        K[j * 65 + d] = row[Hd];
        V[j * 65 + d] = row[2 * Hd];
    }
    __syncthreads();
    constexpr int kMaxK = 4;    // keys per lane: L <= 128
    bool ok[kMaxK];
#pragma unroll
    for (int u = 0; u < kMaxK; ++u) {
        const int j = lane + 32 * u;
        ok[u] = j < L && mask[s * L + (j < L ? j : 0)] != 0;
    }
    for (int i = warp; i < L; i += 4) {
        const float* qrow = qkv + (long long)(s * L + i) * ld + h * 64;
        Q[warp * 64 + lane] = qrow[lane];
        Q[warp * 64 + lane + 32] = qrow[lane + 32];
        __syncwarp();
        float sc[kMaxK], mx = -INFINITY;
#pragma unroll
        for (int u = 0; u < kMaxK; ++u) {
            const int j = lane + 32 * u;
            sc[u] = -INFINITY;
            if (ok[u]) {
                float a = 0.f;
#pragma unroll 16
                for (int d = 0; d < 64; ++d) a = fmaf(Q[warp * 64 + d], K[j * 65 + d], a);
                sc[u] = a * scale;
            }
            mx = fmaxf(mx, sc[u]);
        }
        mx = warp_max(mx);
        float e[kMaxK], den = 0.f;
#pragma unroll
        for (int u = 0; u < kMaxK; ++u) {
            e[u] = ok[u] ? expf(sc[u] - mx) : 0.f;
            den += e[u];
        }
        den = warp_sum(den);
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int u = 0; u < kMaxK; ++u) {
            const float pj = e[u] / den;
            const int n = min(32, L - 32 * u);
            for (int j = 0; j < n; ++j) {
                const float pp = __shfl_sync(0xffffffffu, pj, j);
                o0 = fmaf(pp, V[(32 * u + j) * 65 + lane], o0);
                o1 = fmaf(pp, V[(32 * u + j) * 65 + lane + 32], o1);
            }
        }
        const long long idx = (long long)(s * L + i) * Hd + h * 64;
        store_bf16x1(out_hi, out_ps, idx + lane, o0);
        store_bf16x1(out_hi, out_ps, idx + lane + 32, o1);
        __syncwarp();
    }
}

__global__ void l2norm_rows_kernel(const float* __restrict__ in, int S, int C, int ld_in, float* out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= S) return;
    const int lane = threadIdx.x & 31;
    float sq = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float v = in[(long long)row * ld_in + c];
        sq += v * v;
    }
    const float nrm = fmaxf(sqrtf(warp_sum(sq)), 1e-12f);
    for (int c = lane; c < C; c += 32) out[(long long)row * C + c] = in[(long long)row * ld_in + c] / nrm;
}

__global__ void gather_rows_kernel(const float* __restrict__ in, int S, int C, int row_stride, int ld_in, __nv_bfloat16* out, long long out_ps) {
    const int vec = C / 4;
    const long long total = (long long)S * vec;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / vec;
        const int c = (int)(t % vec) * 4;
        const float4 v = *reinterpret_cast<const float4*>(in + r * row_stride * ld_in + c);
        store_bf16x4(out, out_ps, r * C + c, v.x, v.y, v.z, v.w);
    }
}

// ------------------------------------------------------------------------------------------------
// fold BNContrastiveHead (+ optional L2-normalised text) into GEMM weights; one block per class k
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fold_text_kernel(const float* __restrict__ text, int K, int C, int normalize, const float* __restrict__ g,
                                                        const float* __restrict__ hh, const float* __restrict__ logit_scale, const float* __restrict__ bias,
                                                        __nv_bfloat16* W, long long W_ps, float w_scale, float* bprime) {
    __shared__ float red[8];
    __shared__ float bc;
    const int k = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (k >= K) {  // zero padding rows
        for (int c = threadIdx.x; c < C; c += blockDim.x) store_bf16x1(W, W_ps, (long long)k * C + c, 0.f, w_scale);
        if (threadIdx.x == 0) bprime[k] = 0.f;
        return;
    }
    const float* t = text + (long long)k * C;
    float inv = 1.f;
    if (normalize) {
        float sq = 0.f;
        for (int c = threadIdx.x; c < C; c += blockDim.x) sq += t[c] * t[c];
        sq = warp_sum(sq);
        if (lane == 0) red[warp] = sq;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = 0.f;
            for (int i = 0; i < (blockDim.x >> 5); ++i) tot += red[i];
            bc = 1.f / fmaxf(sqrtf(tot), 1e-12f);
        }
        __syncthreads();
        inv = bc;
        __syncthreads();
    }
    const float es = expf(logit_scale[0]);
    float dot = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float tn = t[c] * inv;
        const float wv = tn * g[c] * es;
        store_bf16x1(W, W_ps, (long long)k * C + c, wv, w_scale);
        dot += hh[c] * tn;
    }
    dot = warp_sum(dot);
    if (lane == 0) red[warp] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int i = 0; i < (blockDim.x >> 5); ++i) tot += red[i];
        bprime[k] = es * tot + bias[0];
    }
}

// kept proposals -> BN'd embedding rows (generate_proposal.py:1129, 1209-1212)
struct GatherEmbedArgs {
    const __nv_bfloat16* emb[3];
    long long emb_ps[3];
    int lvl_size[3];
    int nlevels, B, C, max_keep;
};
// optional (extract_embedding.py:1181-1190,1253-1260): per kept proposal the logit_scale / bias of its pyramid level
__global__ void gather_embed_kernel(GatherEmbedArgs a, const int* __restrict__ keep_anchor, const int* __restrict__ counts, const float* __restrict__ g,
                                    const float* __restrict__ hh, float* out, const float* __restrict__ lvl_scale, const float* __restrict__ lvl_bias,
                                    float* out_scale, float* out_bias) {
    const int j = blockIdx.x, b = blockIdx.y;
    float* o = out + ((long long)b * a.max_keep + j) * a.C;
    if (j >= counts[b]) {
        for (int c = threadIdx.x; c < a.C; c += blockDim.x) o[c] = 0.f;
        if (threadIdx.x == 0 && out_scale) {
            out_scale[b * a.max_keep + j] = 0.f;
            out_bias[b * a.max_keep + j] = 0.f;
        }
        return;
    }
    int anchor = keep_anchor[b * a.max_keep + j];
    int lvl = 0;
    while (lvl + 1 < a.nlevels && anchor >= a.lvl_size[lvl]) {
        anchor -= a.lvl_size[lvl];
        ++lvl;
    }
    if (threadIdx.x == 0 && out_scale) {
        out_scale[b * a.max_keep + j] = lvl_scale[lvl];
        out_bias[b * a.max_keep + j] = lvl_bias[lvl];
    }
    const long long row = (long long)b * a.lvl_size[lvl] + anchor;
    const __nv_bfloat16* e = a.emb[lvl] + row * a.C;
    const long long eps_ = a.emb_ps[lvl];
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
        float v;
        if (eps_) {   // fp16 hi + lo planes of value * kPlaneScale
            const __half* eh = reinterpret_cast<const __half*>(e);
            v = (__half2float(eh[c]) + __half2float(eh[eps_ + c])) * (1.f / kPlaneScale);
        } else {
            v = __bfloat162float(e[c]);
        }
        o[c] = v * g[lvl * a.C + c] + hh[lvl * a.C + c];
    }
}

// ------------------------------------------------------------------------------------------------
// Retrieval scoring (eval_retrieval/retrieval_metric.py:362-374): per image
//     max_j sigmoid( (emb[j] . text[k]) * exp(scale[j]) + bias[j] )
// = scale_rows (emb * exp(scale) -> bf16 GEMM operand) -> tensor-core GEMM against the text matrix -> retr_reduce.
// ------------------------------------------------------------------------------------------------
__global__ void scale_rows_kernel(const float* __restrict__ in, const float* __restrict__ scale, const int* __restrict__ counts, long long rows, int P,
                                  int C, __nv_bfloat16* out, long long out_ps) {
    pdl_launch_dependents();
    pdl_wait();
    const int vec = C / 4;
    const long long total = rows * vec;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / vec;
        const int c = (int)(t % vec) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool valid = counts == nullptr || (int)(r % P) < counts[r / P];
        if (valid) {
            v = *reinterpret_cast<const float4*>(in + r * C + c);
            const float es = scale ? expf(scale[r]) : 1.f;
            v.x *= es; v.y *= es; v.z *= es; v.w *= es;
        }
        store_bf16x4(out, out_ps, r * C + c, v.x, v.y, v.z, v.w);
    }
}

constexpr int kRetrJ = 4;   // proposal lanes per block (blockDim.y)
__global__ void __launch_bounds__(128 * kRetrJ) retr_reduce_kernel(const float* __restrict__ z, int ldz, const float* __restrict__ bias,
                                                                   const int* __restrict__ counts, int P, int K, float* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[kRetrJ][128];
    const int k = blockIdx.x * 128 + threadIdx.x, b = blockIdx.y;
    const int n = counts ? min(counts[b], P) : P;
    float m = 0.f;   // sigmoid > 0: an image without proposals scores 0 everywhere
    if (k < K) {
        const float* zb = z + (long long)b * P * ldz + k;
        for (int j = threadIdx.y; j < n; j += kRetrJ) {
            const float v = zb[(long long)j * ldz] + (bias ? bias[b * P + j] : 0.f);
            m = fmaxf(m, 1.f / (1.f + expf(-v)));
        }
    }
    red[threadIdx.y][threadIdx.x] = m;
    __syncthreads();
    if (threadIdx.y == 0 && k < K) {
#pragma unroll
        for (int y = 1; y < kRetrJ; ++y) m = fmaxf(m, red[y][threadIdx.x]);
        out[(long long)b * K + k] = m;
    }
}

// ------------------------------------------------------------------------------------------------
// host: compile
// ------------------------------------------------------------------------------------------------
static int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148LL * 32;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ------------------------------------------------------------------------------------------------
// Depthwise 7x7, persistent + TMA-pipelined (the production path).  One CTA per SM walks a list of work items
// (image, 8tx x 4ty output tile, 32-channel chunk).  A producer warp streams each item's halo (a rank-4 TMA box whose
// out-of-bounds part is zero-filled = the conv's zero padding) and its 49x32 weights into a two-slot shared-memory ring;
// eight consumer warps (two per scheduler) run the FFMA2 loop on one slot while the other slot is in flight, so the SM
// never stops issuing FMAs to wait for memory and no instruction is spent on addressing the halo.
//   thread = 2 channels x (8 wide x 4 tall) outputs: 32 packed fp32x2 accumulators + a sliding window of 4 weight rows;
//            every halo value is read from shared memory once per thread and feeds up to 28 packed FMAs.
// Measured (B200, stage 2 of WeDetect-Base, bs 32): 87 us per layer = the 3-register-operand FMA issue rate (one FFMA2
// per 4 cycles per scheduler, i.e. 64 FMA/clk/SM) x the 78 % tile efficiency of a 40x40 map; LayerNorm follows as ln_rows
// on the L2-resident fp32 result.
// ------------------------------------------------------------------------------------------------
struct DwTmaParams {
    CUtensorMap tm_in;   // fp32 [B][H][W][C], box (32, 8tx+6, 4ty+6, 1)
    CUtensorMap tm_w;    // fp32 [49][C],      box (32, 49)
    const float* bias;
    float* yscr;
    int H, W, C, nchunks, tx, ty, ntx, nty, num_items;
};
constexpr int kDwConsumerWarps = 8;

__global__ void __launch_bounds__((kDwConsumerWarps + 1) * 32, 1) dwconv7_tma_kernel(const __grid_constant__ DwTmaParams p) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(128) uint8_t dsm_raw[];
    // align by offsetting the __shared__ array itself: the pointer keeps its address space, so the loads below are LDS
    uint8_t* dsm8 = dsm_raw + ((128u - (smem_u32(dsm_raw) & 127u)) & 127u);
    const int HW_ = 8 * p.tx + 6, HH_ = 4 * p.ty + 6;
    const int halo_bytes = HH_ * HW_ * kDw2CK * 4, slot_bytes = halo_bytes + 49 * kDw2CK * 4 + 128 - (49 * kDw2CK * 4) % 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(dsm8 + 2 * slot_bytes);   // full[2], empty[2]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars[i], 1);
            mbar_init(&bars[2 + i], kDwConsumerWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int per_img = p.nty * p.ntx * p.nchunks;
    // each CTA walks a contiguous range of items (chunk fastest, then x tile, y tile, image): the coordinates advance by
    // counters, the divisions of the decode happen once per CTA instead of once per item and warp
    const int ipc = (p.num_items + (int)gridDim.x - 1) / (int)gridDim.x;
    const int it0 = min((int)blockIdx.x * ipc, p.num_items), it1 = min(it0 + ipc, p.num_items);
    if (warp == kDwConsumerWarps) {
        if (lane == 0) {
            int slot = 0, phase = 0;
            int bi = it0 / per_img, ck, xt, yt;
            {
                const int r0 = it0 - bi * per_img, r1 = r0 / p.nchunks;
                ck = r0 % p.nchunks; xt = r1 % p.ntx; yt = r1 / p.ntx;
            }
            for (int item = it0; item < it1; ++item) {
                mbar_wait(&bars[2 + slot], phase ^ 1);
                mbar_arrive_expect_tx(&bars[slot], (uint32_t)(halo_bytes + 49 * kDw2CK * 4));
                uint8_t* dst = dsm8 + slot * slot_bytes;
                tma_load_4d(&p.tm_in, &bars[slot], dst, ck * kDw2CK, xt * 8 * p.tx - 3, yt * 4 * p.ty - 3, bi);
                tma_load_2d(&p.tm_w, &bars[slot], dst + halo_bytes, ck * kDw2CK, 0);
                if (++slot == 2) {
                    slot = 0;
                    phase ^= 1;
                }
                if (++ck == p.nchunks) { ck = 0; if (++xt == p.ntx) { xt = 0; if (++yt == p.nty) { yt = 0; ++bi; } } }
            }
        }
        return;
    }
    const int pair = tid & 15, tile = tid >> 4;
    const bool active = tile < p.tx * p.ty;
    const int tix = tile % p.tx, tiy = tile / p.tx;
    int slot = 0, phase = 0;
    int bi = it0 / per_img, ck, xt, yt;
    {
        const int r0 = it0 - bi * per_img, r1 = r0 / p.nchunks;
        ck = r0 % p.nchunks; xt = r1 % p.ntx; yt = r1 / p.ntx;
    }
    for (int item = it0; item < it1; ++item) {
        const int c0 = ck * kDw2CK, x0 = xt * 8 * p.tx, y0 = yt * 4 * p.ty;
        mbar_wait(&bars[slot], phase);
        if (active && c0 + pair * 2 < p.C && y0 + tiy * 4 < p.H && x0 + tix * 8 < p.W) {
            const float* halo = reinterpret_cast<const float*>(dsm8 + slot * slot_bytes);
            const float* wsm = reinterpret_cast<const float*>(dsm8 + slot * slot_bytes + halo_bytes);
            uint64_t acc[4][8];
#pragma unroll
            for (int oy = 0; oy < 4; ++oy)
#pragma unroll
                for (int ox = 0; ox < 8; ++ox) acc[oy][ox] = 0ull;
            uint64_t wrow[4][7];
#pragma unroll
            for (int oy = 0; oy < 4; ++oy)
#pragma unroll
                for (int dx = 0; dx < 7; ++dx) wrow[oy][dx] = 0ull;
            const float* hbase = halo + ((tiy * 4) * HW_ + tix * 8) * kDw2CK + pair * 2;
#pragma unroll
            for (int iy = 0; iy < 10; ++iy) {
                // wrow[oy] holds the weight row dy = iy - oy
#pragma unroll
                for (int oy = 3; oy > 0; --oy)
#pragma unroll
                    for (int dx = 0; dx < 7; ++dx) wrow[oy][dx] = wrow[oy - 1][dx];
                if (iy < 7) {
#pragma unroll
                    for (int dx = 0; dx < 7; ++dx) wrow[0][dx] = *reinterpret_cast<const uint64_t*>(wsm + (iy * 7 + dx) * kDw2CK + pair * 2);
                }
                const float* hrow = hbase + iy * HW_ * kDw2CK;
#pragma unroll
                for (int ix = 0; ix < 14; ++ix) {
                    const uint64_t v = *reinterpret_cast<const uint64_t*>(hrow + ix * kDw2CK);
#pragma unroll
                    for (int oy = 0; oy < 4; ++oy) {
                        if (iy - oy >= 0 && iy - oy < 7) {
#pragma unroll
                            for (int ox = 0; ox < 8; ++ox) {
                                if (ix - ox >= 0 && ix - ox < 7) acc[oy][ox] = ffma2(v, wrow[oy][ix - ox], acc[oy][ox]);
                            }
                        }
                    }
                }
            }
            const float2 b2 = __ldg(reinterpret_cast<const float2*>(p.bias + c0 + pair * 2));
            float* obase = p.yscr + (((long long)bi * p.H + y0 + tiy * 4) * p.W + x0 + tix * 8) * p.C + c0 + pair * 2;
            if (y0 + tiy * 4 + 3 < p.H && x0 + tix * 8 + 7 < p.W) {
                // interior thread tile (every tile of the 160 / 80 / 40 maps): unchecked 8-byte stores off two running pointers
                uint64_t b2p;
                asm("mov.b64 %0, {%1, %2};" : "=l"(b2p) : "f"(b2.x), "f"(b2.y));
                const long long cs = p.C, rs = (long long)p.W * p.C;
                float* prow = obase;
#pragma unroll
                for (int oy = 0; oy < 4; ++oy) {
                    float* pp = prow;
#pragma unroll
                    for (int ox = 0; ox < 8; ++ox) {
                        uint64_t v;
                        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(acc[oy][ox]), "l"(b2p));
                        *reinterpret_cast<uint64_t*>(pp) = v;
                        pp += cs;
                    }
                    prow += rs;
                }
            } else {
#pragma unroll
                for (int oy = 0; oy < 4; ++oy) {
                    if (y0 + tiy * 4 + oy < p.H) {
#pragma unroll
                        for (int ox = 0; ox < 8; ++ox) {
                            if (x0 + tix * 8 + ox < p.W)
                                *reinterpret_cast<float2*>(obase + ((long long)oy * p.W + ox) * p.C) =
                                    make_float2(unpack_lo(acc[oy][ox]) + b2.x, unpack_hi(acc[oy][ox]) + b2.y);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[2 + slot]);   // this warp is done reading the slot
        if (++slot == 2) {
            slot = 0;
            phase ^= 1;
        }
        if (++ck == p.nchunks) { ck = 0; if (++xt == p.ntx) { xt = 0; if (++yt == p.nty) { yt = 0; ++bi; } } }
    }
}

int compile_rowops(const wd_op& op, std::unique_ptr<CompiledOp>& out) {
    auto f = std::make_unique<FnOp>();
    const int32_t* I = op.i;
    void* const* P = op.p;
    const float* F = op.f;
    if (device_sm_count() <= 0) return -2;
    switch (op.kind) {
        case WD_OP_LN_ROWS: {
            const int rows = I[0], C = I[1], s2d = I[4], W = I[5], H = I[6], ld_in = I[7], ld_out = I[8];
            const float eps = F[0];
            WD_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && C <= kLnMaxVec * 128, "ln_rows: bad shape rows=%d C=%d", rows, C);
            WD_REQUIRE(P[0] && P[2] && P[3] && (P[1] || P[6]), "ln_rows: null pointer");
            WD_REQUIRE(!s2d || (W % 2 == 0 && H % 2 == 0 && rows % (W * H) == 0 && P[1]), "ln_rows: bad s2d geometry");
            WD_REQUIRE(ld_in % 4 == 0 && ld_out % 4 == 0, "ln_rows: ld must be a multiple of 4");
            const float* in = (const float*)P[0];
            const float *w = (const float*)P[2], *b = (const float*)P[3];
            __nv_bfloat16* oh = (__nv_bfloat16*)P[1];
            const long long ol = I[30];
            float* of = (float*)P[6];
            f->fn = [=](cudaStream_t s) {
                launch_ln_rows(s, in, rows, C, ld_in, w, b, eps, oh, ol, of, ld_out, s2d, W, H);
                WD_CHECK_CUDA(cudaGetLastError());
                count_launch();
                return 0;
            };
            break;
        }
        case WD_OP_DWCONV_LN: {
            const int B = I[0], H = I[1], W = I[2], C = I[3];
            const float eps = F[0];
            WD_REQUIRE(B > 0 && H > 0 && W > 0 && C % 4 == 0 && C / 4 <= 384, "dwconv_ln: bad shape");
            WD_REQUIRE(P[0] && P[1] && P[2] && P[3] && P[4] && P[5], "dwconv_ln: null pointer");
            const float* in = (const float*)P[0];
            __nv_bfloat16* oh = (__nv_bfloat16*)P[1];
            const long long ol = I[30];
            const float *wt = (const float*)P[2], *bs = (const float*)P[3], *lw = (const float*)P[4], *lb = (const float*)P[5];
            const int threads = ((C / 4 + 31) / 32) * 32;
            const int ld_out = I[4] > 0 ? I[4] : C;
            WD_REQUIRE(ld_out >= C && ld_out % 4 == 0, "dwconv_ln: bad ld_out");
            float* yscr = (float*)P[7];
            if (yscr) {  // persistent TMA-pipelined conv -> fp32 scratch (L2 resident), then LayerNorm rows -> bf16
                const int tx = I[5], ty = I[6];
                WD_REQUIRE(tx >= 1 && ty >= 1 && tx * ty <= 2 * kDwConsumerWarps && C % 4 == 0 && C <= kLnMaxVec * 128, "dwconv_ln: bad tile %d x %d", tx, ty);
                const int HW_ = 8 * tx + 6, HH_ = 4 * ty + 6;
                const int halo_bytes = HH_ * HW_ * kDw2CK * 4, slot_bytes = halo_bytes + 49 * kDw2CK * 4 + 128 - (49 * kDw2CK * 4) % 128;
                const int smem = 2 * slot_bytes + 64 + 128;
                WD_REQUIRE(HW_ <= 256 && HH_ <= 256 && smem <= 227 * 1024, "dwconv_ln: tile %d x %d needs %d bytes of shared memory", tx, ty, smem);
                struct DwOp : CompiledOp {
                    DwTmaParams prm;
                    int grid, smem, rows, ld_out;
                    const float *lw, *lb;
                    float eps;
                    __nv_bfloat16* oh;
                    long long ol;
                    int launch(cudaStream_t s) override {
                        WD_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(dwconv7_tma_kernel), 227 * 1024));
                        launch_pdl(dwconv7_tma_kernel, dim3(grid), dim3((kDwConsumerWarps + 1) * 32), (size_t)(smem), s, 1, prm);
                        launch_ln_rows(s, prm.yscr, rows, prm.C, prm.C, lw, lb, eps, oh, ol, nullptr, ld_out, 0, 0, 0);
                        WD_CHECK_CUDA(cudaGetLastError());
                        count_launch(2);
                        return 0;
                    }
                    int num_kernels() const override { return 2; }
                };
                auto d = std::make_unique<DwOp>();
                DwTmaParams& q = d->prm;
                q.bias = bs; q.yscr = yscr; q.H = H; q.W = W; q.C = C;
                q.nchunks = (C + kDw2CK - 1) / kDw2CK;
                q.tx = tx; q.ty = ty;
                q.ntx = (W + 8 * tx - 1) / (8 * tx);
                q.nty = (H + 4 * ty - 1) / (4 * ty);
                const long long items = (long long)B * q.nty * q.ntx * q.nchunks;
                WD_REQUIRE(items < (1ll << 31), "dwconv_ln: too many work items");
                q.num_items = (int)items;
                {
                    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
                    const uint64_t str[3] = {(uint64_t)C * 4, (uint64_t)W * C * 4, (uint64_t)H * W * C * 4};
                    const uint32_t box[4] = {(uint32_t)kDw2CK, (uint32_t)HW_, (uint32_t)HH_, 1u};
                    if (encode_tmap(&q.tm_in, in, 4, 4, dims, str, box, false)) return -1;
                    const uint64_t wd[2] = {(uint64_t)C, 49};
                    const uint64_t ws[1] = {(uint64_t)C * 4};
                    const uint32_t wb[2] = {(uint32_t)kDw2CK, 49u};
                    if (encode_tmap(&q.tm_w, wt, 4, 2, wd, ws, wb, false)) return -1;
                }
                d->grid = (int)std::min<long long>(items, device_sm_count());
                d->smem = smem; d->rows = B * H * W; d->ld_out = ld_out;
                d->lw = lw; d->lb = lb; d->eps = eps; d->oh = oh; d->ol = ol;
                out = std::move(d);
                return 0;
            }
            f->fn = [=](cudaStream_t s) {
                dim3 grid((W + kDwTX - 1) / kDwTX, (H + kDwTY - 1) / kDwTY, B);
                launch_pdl(dwconv7_ln_kernel, dim3(grid), dim3(threads), (size_t)(0), s, 1, in, B, H, W, C, wt, bs, lw, lb, eps, oh, ol, ld_out);
                WD_CHECK_CUDA(cudaGetLastError());
                count_launch();
                return 0;
            };
            break;
        }
        case WD_OP_STEM_PATCH: {
            const int B = I[0], H = I[1], W = I[2], dt = I[3], layout = I[4];
            const float scale = F[0];
            WD_REQUIRE(B > 0 && H % 4 == 0 && W % 4 == 0 && layout == 0 && (dt == 0 || dt == 2), "stem_patch: bad arguments");
            WD_REQUIRE(P[0] && P[1], "stem_patch: null pointer");
            const void* in = P[0];
            __nv_bfloat16* oh = (__nv_bfloat16*)P[1];
            const long long ol = I[30];
            const long long total = (long long)B * (H / 4) * (W / 4) * 16;
            f->fn = [=](cudaStream_t s) {
                if (dt == 0) launch_pdl(stem_patch_kernel<uint8_t>, dim3(grid_for(total, 256)), dim3(256), (size_t)0, s, 1, (const uint8_t*)in, B, H, W, scale, oh, ol);
                else launch_pdl(stem_patch_kernel<float>, dim3(grid_for(total, 256)), dim3(256), (size_t)0, s, 1, (const float*)in, B, H, W, scale, oh, ol);
                WD_CHECK_CUDA(cudaGetLastError());
                count_launch();
                return 0;
            };
            break;
        }
        case WD_OP_IM2COL_S2: {
            const int B = I[0], H = I[1], W = I[2], C = I[3], ld_in = I[4];
            WD_REQUIRE(B > 0 && H > 0 && W > 0 && C % 8 == 0 && ld_in % 8 == 0 && P[0] && P[1], "im2col_s2: bad arguments");
            const __nv_bfloat16* in = (const __nv_bfloat16*)P[0];
            __nv_bfloat16* o = (__nv_bfloat16*)P[1];
            const long long inl = I[31], ol = I[30];
            WD_REQUIRE((inl == 0) == (ol == 0), "im2col_s2: both or neither plane stride");
            const long long total = (long long)B * ((H - 1) / 2 + 1) * ((W - 1) / 2 + 1) * 9 * (C / 8);
            f->fn = [=](cudaStream_t s) {
                launch_pdl(im2col_s2_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)0, s, 1, in, inl, B, H, W, C, ld_in, o, ol);
                WD_CHECK_CUDA(cudaGetLastError());
                count_launch();
                return 0;
            };
            break;
        }
        case WD_OP_CAST_BF16: {
            const int rows = I[0], C = I[1], ld_in = I[2], ld_out = I[3];
            WD_REQUIRE(rows > 0 && C % 4 == 0 && ld_in % 4 == 0 && ld_out % 4 == 0 && P[0] && P[1], "cast_bf16: bad arguments");
            const float* in = (const float*)P[0];
            __nv_bfloat16* o = (__nv_bfloat16*)P[1];
            const long long ol = I[30];
            f->fn = [=](cudaStream_t s) {
                launch_pdl(cast_bf16_kernel, dim3(grid_for((long long)rows * (C / 4), 256)), dim3(256), (size_t)0, s, 1, in, rows, C, ld_in, ld_out, o, ol);
                WD_CHECK_CUDA(cudaGetLastError());
                count_launch();
                return 0;
            };
            break;
        }
        case WD_OP_TEXT_EMBED: {
            const int S = I[0], L = I[1], Hd = I[2], pad_idx = I[3];
            const float eps = F[0];
            WD_REQUIRE(S > 0 && L > 0 && Hd % 4 == 0 && Hd <= 1024, "text_embed: bad shape");
            for (int k = 0; k <= 8; ++k) WD_REQUIRE(k == 1 || P[k], "text_embed: null pointer %d", k);
            const int* ids = (const int*)P[0];
            const float *word = (const float*)P[2], *pos = (const float*)P[3], *type = (const float*)P[4], *lw = (const float*)P[5], *lb = (const float*)P[6];
            float* of = (float*)P[7];
            __nv_bfloat16* oh = (__nv_bfloat16*)P[8];
            const long long ol = I[30];
            f->fn = [=](cudaStream_t s) {
                text_embed_kernel<<<(S * L + 7) / 8, 256, 0, s>>>(ids, S, L, Hd, pad_idx, word, pos, type, lw, lb, eps, of, oh, ol);
                WD_CHECK_CUDA(cudaGetLastError());
                count_launch();
                return 0;
            };
            break;
        }
        case WD_OP_ATTN_SMALL: {
            const int S = I[0], L = I[1], heads = I[2], hd = I[3], ld = I[4];
            const float scale = F[0];
            WD_REQUIRE(S > 0 && L > 0 && L <= 128 && hd == 64 && heads > 0 && P[0] && P[1] && P[2], "attn_small: needs L <= 128 tokens and head_dim == 64");
            const float* qkv = (const float*)P[0];
            const int* mask = (const int*)P[1];
            __nv_bfloat16* oh = (__nv_bfloat16*)P[2];
            const long long ol = I[30];
            const int smem = 4 * 3 * L * 65 * (int)sizeof(float);
            const int smem_mid = (2 * L * 65 + 4 * 64) * (int)sizeof(float);
            f->fn = [=](cudaStream_t s) {
                if (L > 32) {   // longer prompts: one block per (sequence, head)
                    WD_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(attn_mid_kernel), (2 * 128 * 65 + 4 * 64) * 4));
                    attn_mid_kernel<<<S * heads, 128, smem_mid, s>>>(qkv, mask, S, L, heads, ld, scale, oh, ol);
                    WD_CHECK_CUDA(cudaGetLastError());
                    count_launch();
                    return 0;
                }
                WD_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(attn_small_kernel), 4 * 3 * 32 * 65 * 4));
                attn_small_kernel<<<(S * heads + 3) / 4, 128, smem, s>>>(qkv, mask, S, L, heads, ld, scale, oh, ol);
                WD_CHECK_CUDA(cudaGetLastError());
                count_launch();
                return 0;
            };
            break;
        }
        case WD_OP_L2NORM_ROWS: {
            const int S = I[0], C = I[1], ld_in = I[2];
            WD_REQUIRE(S > 0 && C > 0 && P[0] && P[1], "l2norm_rows: bad arguments");
            const float* in = (const float*)P[0];
            float* o = (float*)P[1];
            f->fn = [=](cudaStream_t s) {
                l2norm_rows_kernel<<<(S + 7) / 8, 256, 0, s>>>(in, S, C, ld_in, o);
                WD_CHECK_CUDA(cudaGetLastError());
                count_launch();
                return 0;
            };
            break;
        }
        case WD_OP_GATHER_ROWS: {
            const int S = I[0], C = I[1], rs = I[2], ld_in = I[3];
            WD_REQUIRE(S > 0 && C % 4 == 0 && ld_in % 4 == 0 && P[0] && P[1], "gather_rows: bad arguments");
            const float* in = (const float*)P[0];
            __nv_bfloat16* o = (__nv_bfloat16*)P[1];
            const long long ol = I[30];
            f->fn = [=](cudaStream_t s) {
                gather_rows_kernel<<<grid_for((long long)S * (C / 4), 256), 256, 0, s>>>(in, S, C, rs, ld_in, o, ol);
                WD_CHECK_CUDA(cudaGetLastError());
                count_launch();
                return 0;
            };
            break;
        }
        case WD_OP_FOLD_TEXT: {
            const int K = I[0], C = I[1], normalize = I[2], Kpad = I[3];
            WD_REQUIRE(K > 0 && C > 0 && Kpad >= K, "fold_text: bad shape");
            for (int k = 0; k <= 6; ++k) WD_REQUIRE(P[k], "fold_text: null pointer %d", k);
            const float *t = (const float*)P[0], *g = (const float*)P[1], *hh = (const float*)P[2], *ls = (const float*)P[3], *bi = (const float*)P[4];
            __nv_bfloat16* Wd = (__nv_bfloat16*)P[5];
            const long long Wl = I[30];
            float* bp = (float*)P[6];
            const float wsc = F[0] != 0.f ? F[0] : 1.f;   // power of two the fp16 hi/lo planes of W' are stored at
            f->fn = [=](cudaStream_t s) {
                fold_text_kernel<<<Kpad, 256, 0, s>>>(t, K, C, normalize, g, hh, ls, bi, Wd, Wl, wsc, bp);
                WD_CHECK_CUDA(cudaGetLastError());
                count_launch();
                return 0;
            };
            break;
        }
        case WD_OP_GATHER_EMBED: {
            GatherEmbedArgs a;
            a.B = I[0]; a.C = I[2]; a.max_keep = I[3]; a.nlevels = I[4];
            WD_REQUIRE(a.B > 0 && a.C > 0 && a.max_keep > 0 && a.nlevels >= 1 && a.nlevels <= 3, "gather_embed: bad shape");
            for (int l = 0; l < 3; ++l) {
                a.lvl_size[l] = I[5 + l];
                a.emb[l] = (const __nv_bfloat16*)P[l];
                a.emb_ps[l] = I[30 + l];
            }
            WD_REQUIRE(P[0] && P[3] && P[4] && P[5] && P[6] && P[7], "gather_embed: null pointer");
            const int* ka = (const int*)P[3];
            const int* cnt = (const int*)P[4];
            const float *g = (const float*)P[5], *hh = (const float*)P[6];
            float* o = (float*)P[7];
            const float *ls = (const float*)P[8], *lb = (const float*)P[9];
            float *os = (float*)P[10], *ob = (float*)P[11];
            WD_REQUIRE((os == nullptr) == (ob == nullptr) && (os == nullptr || (ls && lb)), "gather_embed: scale / bias outputs need both outputs and both level tables");
            f->fn = [=](cudaStream_t s) {
                gather_embed_kernel<<<dim3(a.max_keep, a.B), 128, 0, s>>>(a, ka, cnt, g, hh, o, ls, lb, os, ob);
                WD_CHECK_CUDA(cudaGetLastError());
                count_launch();
                return 0;
            };
            break;
        }
        case WD_OP_SCALE_ROWS: {
            const int B = I[0], Pn = I[1], C = I[2];
            WD_REQUIRE(B > 0 && Pn > 0 && C > 0 && C % 4 == 0 && P[0] && P[3], "scale_rows: bad arguments");
            const float *in = (const float*)P[0], *sc = (const float*)P[1];
            const int* cnt = (const int*)P[2];
            __nv_bfloat16* o = (__nv_bfloat16*)P[3];
            const long long ol = I[30];
            const long long rows = (long long)B * Pn;
            f->fn = [=](cudaStream_t s) {
                launch_pdl(scale_rows_kernel, dim3(grid_for(rows * (C / 4), 256)), dim3(256), (size_t)0, s, 1, in, sc, cnt, rows, Pn, C, o, ol);
                WD_CHECK_CUDA(cudaGetLastError());
                count_launch();
                return 0;
            };
            break;
        }
        case WD_OP_RETR_REDUCE: {
            const int B = I[0], Pn = I[1], K = I[2], ldz = I[3];
            WD_REQUIRE(B > 0 && B <= 65535 && Pn > 0 && K > 0 && ldz >= K && P[0] && P[3], "retr_reduce: bad arguments");
            const float *z = (const float*)P[0], *bias = (const float*)P[1];
            const int* cnt = (const int*)P[2];
            float* o = (float*)P[3];
            f->fn = [=](cudaStream_t s) {
                launch_pdl(retr_reduce_kernel, dim3((K + 127) / 128, B), dim3(128, kRetrJ), (size_t)0, s, 1, z, ldz, bias, cnt, Pn, K, o);
                WD_CHECK_CUDA(cudaGetLastError());
                count_launch();
                return 0;
            };
            break;
        }
        default: set_last_error("rowops: unknown kind %d", op.kind); return -1;
    }
    out = std::move(f);
    return 0;
}

}  // namespace wd
