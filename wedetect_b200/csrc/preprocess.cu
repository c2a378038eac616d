// preprocess.cu — Uni-path image preparation on the device (generate_proposal.py:17-82 letterbox, :1087-1101):
// PIL-exact BILINEAR resize + centred paste on a grey canvas, written straight into the detector's planar uint8 input.
//
// PIL's 8-bit resampler (Pillow src/libImaging/Resample.c) is integer arithmetic: per output coordinate a window
// (first source index, count) and 22-bit fixed-point triangle-filter weights (the support widens with the down-scale
// factor), accumulators start at 2^21, result = clip8(acc >> 22); the horizontal pass runs first into an 8-bit
// intermediate image that holds only the source rows the vertical pass will read.  The weight tables are built by the
// host in double precision exactly as precompute_coeffs does (wedetect_b200/preprocess.py); the two kernels below do the
// byte work and are bit-exact with PIL.  A pass that PIL skips (size unchanged) is expressed as identity tables
// (window of one pixel, weight 2^22), which reproduces the copy exactly.  Pillow >= 12 resizes very tall images (h > 100 w,
// height shrinking) vertically first (Image.resize); the host flags those and both kernels swap their roles.
//
// HBM-bound byte work: each source byte is read once per pass (neighbouring threads read neighbouring pixels, reuse
// across the filter window comes from L1), each output byte is written once, fully coalesced per colour plane.
#include "internal.h"
#include <algorithm>

namespace wd {

constexpr int kLbPrecision = 32 - 8 - 2;
constexpr int kLbDesc = 16;   // int32 words per image in the descriptor table (layout in wedetect_b200.h)

__device__ __forceinline__ uint8_t lb_clip8(int v) {
    v >>= kLbPrecision;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

__device__ __forceinline__ long long lb_off64(const int* d, int lo) {
    return (long long)(((unsigned long long)(unsigned)d[lo + 1] << 32) | (unsigned long long)(unsigned)d[lo]);
}

// One resampled pixel: sum_j in[window start + j * tap] * weight[j] for the three channels.
//   along_x: the window runs along a row (tables indexed by the output column), else down a column (indexed by the row).
__device__ __forceinline__ void lb_resample_px(const uint8_t* __restrict__ in, long long row_stride, const int* __restrict__ bounds,
                                               const int* __restrict__ kk, int ksize, bool along_x, int row, int col, uint8_t& v0, uint8_t& v1,
                                               uint8_t& v2) {
    const int o = along_x ? col : row;
    const int lo = bounds[2 * o], cnt = bounds[2 * o + 1];
    const int* k = kk + (long long)o * ksize;
    const uint8_t* p = along_x ? in + row * row_stride + (long long)lo * 3 : in + lo * row_stride + (long long)col * 3;
    const long long tap = along_x ? 3 : row_stride;
    int s0 = 1 << (kLbPrecision - 1), s1 = s0, s2 = s0;
    for (int j = 0; j < cnt; ++j) {
        const int kj = k[j];
        const uint8_t* q = p + j * tap;
        s0 += (int)q[0] * kj;
        s1 += (int)q[1] * kj;
        s2 += (int)q[2] * kj;
    }
    v0 = lb_clip8(s0);
    v1 = lb_clip8(s1);
    v2 = lb_clip8(s2);
}

struct LbGeom {
    int src_w, new_w, new_h, left, top, first_row, tmp_rows, ksize_h, ksize_v, vfirst;
    const int *bh, *kh, *bv, *kv;
    long long src_off, tmp_off;
};
__device__ __forceinline__ LbGeom lb_geom(const int* __restrict__ desc, const int* __restrict__ coef, int b) {
    const int* d = desc + b * kLbDesc;
    LbGeom g;
    g.src_off = lb_off64(d, 0); g.src_w = d[2]; g.new_w = d[4]; g.new_h = d[5]; g.left = d[6]; g.top = d[7];
    g.first_row = d[8]; g.tmp_rows = d[9]; g.tmp_off = lb_off64(d, 10); g.ksize_h = d[13]; g.ksize_v = d[14]; g.vfirst = d[15];
    g.bh = coef + d[12];
    g.kh = g.bh + 2 * g.new_w;
    g.bv = g.kh + (long long)g.new_w * g.ksize_h;
    g.kv = g.bv + 2 * g.new_h;
    return g;
}

// first pass: source -> 8-bit intermediate.  Normal order (PIL's ImagingResample): horizontal, over the source rows
// [first_row, first_row + tmp_rows) the vertical pass needs -> tmp [tmp_rows, new_w, 3].  vfirst (PIL >= 12 resizes very
// tall images, h > 100 w, vertically first): vertical over full-width rows -> tmp [new_h, src_w, 3].
__global__ void __launch_bounds__(256) lb_pass1_kernel(const uint8_t* __restrict__ src, const int* __restrict__ desc, const int* __restrict__ coef,
                                                       uint8_t* __restrict__ tmp) {
    const LbGeom g = lb_geom(desc, coef, blockIdx.y);
    const int cols = g.vfirst ? g.src_w : g.new_w;
    const long long total = (long long)g.tmp_rows * cols;
    if (total == 0) return;
    const long long stride = (long long)g.src_w * 3;
    const uint8_t* sb = src + g.src_off + (g.vfirst ? 0 : g.first_row * stride);
    uint8_t* tb = tmp + g.tmp_off;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(t / cols), x = (int)(t - (long long)y * cols);
        uint8_t v0, v1, v2;
        if (g.vfirst) lb_resample_px(sb, stride, g.bv, g.kv, g.ksize_v, false, y, x, v0, v1, v2);
        else lb_resample_px(sb, stride, g.bh, g.kh, g.ksize_h, true, y, x, v0, v1, v2);
        uint8_t* o = tb + t * 3;
        o[0] = v0;
        o[1] = v1;
        o[2] = v2;
    }
}

// second pass + paste: tmp -> out [B, 3, H, W] planar; everything outside the pasted rectangle is the pad colour
__global__ void __launch_bounds__(256) lb_pass2_paste_kernel(const uint8_t* __restrict__ tmp, const int* __restrict__ desc, const int* __restrict__ coef,
                                                             uint8_t* __restrict__ out, int H, int W, int pad) {
    const int b = blockIdx.y;
    const LbGeom g = lb_geom(desc, coef, b);
    const uint8_t* tb = tmp + g.tmp_off;
    const long long plane = (long long)H * W;
    uint8_t* ob = out + (long long)b * 3 * plane;
    const long long rstride = (long long)(g.vfirst ? g.src_w : g.new_w) * 3;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < plane; t += (long long)gridDim.x * blockDim.x) {
        const int Y = (int)(t / W), X = (int)(t - (long long)Y * W);
        const int yy = Y - g.top, xx = X - g.left;
        uint8_t v0 = (uint8_t)pad, v1 = (uint8_t)pad, v2 = (uint8_t)pad;
        if (yy >= 0 && yy < g.new_h && xx >= 0 && xx < g.new_w) {
            if (g.vfirst) lb_resample_px(tb, rstride, g.bh, g.kh, g.ksize_h, true, yy, xx, v0, v1, v2);
            else lb_resample_px(tb, rstride, g.bv, g.kv, g.ksize_v, false, yy, xx, v0, v1, v2);
        }
        ob[t] = v0;
        ob[plane + t] = v1;
        ob[2 * plane + t] = v2;
    }
}

// ------------------------------------------------------------------------------------------------
// mmcv / cv2 test pipeline (WeDetectKeepRatioResize + WeDetectLetterResize, transforms.py:94-123,180-272): cv2.resize of a
// decoded uint8 BGR image (INTER_AREA when shrinking, INTER_LINEAR when growing) pasted at (left, top) on a grey canvas.
// Bit-exact with OpenCV 4.x (modules/imgproc/src/resize.cpp); the host builds the tables exactly as OpenCV does
// (wedetect_b200/preprocess.py), this kernel does the per-pixel arithmetic in OpenCV's operation order:
//   mode 1  INTER_AREA, fractional scale (ResizeArea_<uchar, float>): per source row a float sum of S * alpha over the x-table
//           entries in order (separate multiply and add, no FMA), rows combined as beta * rowsum, first row assigns;
//           saturate_cast<uchar> = round-half-even + clamp.
//   mode 2  INTER_AREA, integer scale (ResizeAreaFast_): integer box sum; 2x2 boxes (sum + 2) >> 2, else round(sum * (1.f / area)).
//   mode 3  INTER_LINEAR (8u): 11-bit fixed-point horizontal taps (int), vertical ((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2
//           on rows clipped to the image; columns >= xmax replicate the last source column.
//   mode 0  no resize (copy).
// One thread per canvas pixel, three channels; output planes are written fully coalesced.
__device__ __forceinline__ uint8_t cv_sat8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

__global__ void __launch_bounds__(256) cv_resize_pad_kernel(const uint8_t* __restrict__ src, const int* __restrict__ desc, const int* __restrict__ coef,
                                                            uint8_t* __restrict__ out, int H, int W, int pad) {
    const int b = blockIdx.y;
    const int* d = desc + b * kLbDesc;
    const long long src_off = lb_off64(d, 0);
    const int sw = d[2], sh = d[3], nw = d[4], nh = d[5], left = d[6], top = d[7], mode = d[8];
    const int* tab = coef + d[9];
    const uint8_t* sb = src + src_off;
    const long long rs = (long long)sw * 3;
    const long long plane = (long long)H * W;
    uint8_t* ob = out + (long long)b * 3 * plane;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < plane; t += (long long)gridDim.x * blockDim.x) {
        const int Y = (int)(t / W), X = (int)(t - (long long)Y * W);
        const int dy = Y - top, dx = X - left;
        uint8_t v[3] = {(uint8_t)pad, (uint8_t)pad, (uint8_t)pad};
        if (dy >= 0 && dy < nh && dx >= 0 && dx < nw) {
            if (mode == 0) {
                const uint8_t* q = sb + dy * rs + (long long)dx * 3;
                v[0] = q[0]; v[1] = q[1]; v[2] = q[2];
            } else if (mode == 1) {
                const int* xidx = tab;
                const int* yidx = tab + nw + 1;
                const int nx = xidx[nw], ny = yidx[nh];
                const int* xs = yidx + nh + 1;
                const float* xa = reinterpret_cast<const float*>(xs + nx);
                const int* ys = xs + 2 * nx;
                const float* yb = reinterpret_cast<const float*>(ys + ny);
                const int k0 = xidx[dx], k1 = xidx[dx + 1];
                float sum[3] = {0.f, 0.f, 0.f};
                for (int j = yidx[dy]; j < yidx[dy + 1]; ++j) {
                    const uint8_t* row = sb + ys[j] * rs;
                    const float beta = yb[j];
                    float h0 = 0.f, h1 = 0.f, h2 = 0.f;
                    for (int k = k0; k < k1; ++k) {
                        const uint8_t* q = row + (long long)xs[k] * 3;
                        const float a = xa[k];
                        h0 = __fadd_rn(h0, __fmul_rn((float)q[0], a));
                        h1 = __fadd_rn(h1, __fmul_rn((float)q[1], a));
                        h2 = __fadd_rn(h2, __fmul_rn((float)q[2], a));
                    }
                    if (j == yidx[dy]) {
                        sum[0] = __fmul_rn(beta, h0); sum[1] = __fmul_rn(beta, h1); sum[2] = __fmul_rn(beta, h2);
                    } else {
                        sum[0] = __fadd_rn(sum[0], __fmul_rn(beta, h0));
                        sum[1] = __fadd_rn(sum[1], __fmul_rn(beta, h1));
                        sum[2] = __fadd_rn(sum[2], __fmul_rn(beta, h2));
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = cv_sat8(__float2int_rn(sum[c]));
            } else if (mode == 2) {
                const int kx = d[10], ky = d[11];
                const float scale = __int_as_float(d[12]);
                int s0 = 0, s1 = 0, s2 = 0;
                for (int yy = 0; yy < ky; ++yy) {
                    const uint8_t* q = sb + (long long)(dy * ky + yy) * rs + (long long)dx * kx * 3;
                    for (int xx = 0; xx < kx; ++xx, q += 3) {
                        s0 += q[0]; s1 += q[1]; s2 += q[2];
                    }
                }
                if (kx == 2 && ky == 2) {
                    v[0] = (uint8_t)((s0 + 2) >> 2); v[1] = (uint8_t)((s1 + 2) >> 2); v[2] = (uint8_t)((s2 + 2) >> 2);
                } else {
                    v[0] = cv_sat8(__float2int_rn(__fmul_rn((float)s0, scale)));
                    v[1] = cv_sat8(__float2int_rn(__fmul_rn((float)s1, scale)));
                    v[2] = cv_sat8(__float2int_rn(__fmul_rn((float)s2, scale)));
                }
            } else {
                const int* xofs = tab;
                const int* xab = tab + nw;
                const int* yofs = xab + nw;
                const int* yab = yofs + nh;
                const int xmax = d[13];
                const int sx = xofs[dx], a0 = (int)(short)(xab[dx] & 0xffff), a1 = xab[dx] >> 16;
                const int sy = yofs[dy], b0 = (int)(short)(yab[dy] & 0xffff), b1 = yab[dy] >> 16;
                const int r0 = min(max(sy, 0), sh - 1), r1 = min(max(sy + 1, 0), sh - 1);
                const uint8_t* q0 = sb + r0 * rs + (long long)sx * 3;
                const uint8_t* q1 = sb + r1 * rs + (long long)sx * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    int h0, h1;
                    if (dx < xmax) {
                        h0 = (int)q0[c] * a0 + (int)q0[c + 3] * a1;
                        h1 = (int)q1[c] * a0 + (int)q1[c + 3] * a1;
                    } else {
                        h0 = (int)q0[c] * 2048;
                        h1 = (int)q1[c] * 2048;
                    }
                    v[c] = cv_sat8((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
                }
            }
        }
        ob[t] = v[0];
        ob[plane + t] = v[1];
        ob[2 * plane + t] = v[2];
    }
}

int compile_preprocess(const wd_op& op, std::unique_ptr<CompiledOp>& out) {
    const int32_t* I = op.i;
    void* const* P = op.p;
    if (device_sm_count() <= 0) return -2;
    WD_REQUIRE(op.kind == WD_OP_LETTERBOX || op.kind == WD_OP_CV_RESIZE_PAD, "preprocess: unknown kind %d", op.kind);
    const int B = I[0], H = I[1], W = I[2], pad = I[3];
    WD_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && pad >= 0 && pad <= 255, "letterbox: bad shape B=%d H=%d W=%d pad=%d", B, H, W, pad);
    const bool cv = op.kind == WD_OP_CV_RESIZE_PAD;
    for (int k = 0; k <= 4; ++k) WD_REQUIRE(P[k] || (cv && k == 3), "letterbox: null pointer %d", k);
    struct LbOp : CompiledOp {
        const uint8_t* src;
        const int *desc, *coef;
        uint8_t *tmp, *o;
        int B, H, W, pad, gx, cv;
        int launch(cudaStream_t s) override {
            if (cv) {
                cv_resize_pad_kernel<<<dim3(gx, B), 256, 0, s>>>(src, desc, coef, o, H, W, pad);
            } else {
                lb_pass1_kernel<<<dim3(gx, B), 256, 0, s>>>(src, desc, coef, tmp);
                lb_pass2_paste_kernel<<<dim3(gx, B), 256, 0, s>>>(tmp, desc, coef, o, H, W, pad);
            }
            WD_CHECK_CUDA(cudaGetLastError());
            count_launch(cv ? 1 : 2);
            return 0;
        }
        int num_kernels() const override { return cv ? 1 : 2; }
    };
    auto d = std::make_unique<LbOp>();
    d->src = (const uint8_t*)P[0];
    d->desc = (const int*)P[1];
    d->coef = (const int*)P[2];
    d->tmp = (uint8_t*)P[3];
    d->o = (uint8_t*)P[4];
    d->B = B; d->H = H; d->W = W; d->pad = pad; d->cv = cv ? 1 : 0;
    d->gx = std::max(8, (device_sm_count() * 16 + B - 1) / B);   // ~16 resident blocks per SM over the whole batch
    out = std::move(d);
    return 0;
}

}  // namespace wd
