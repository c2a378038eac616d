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

int compile_preprocess(const wd_op& op, std::unique_ptr<CompiledOp>& out) {
    const int32_t* I = op.i;
    void* const* P = op.p;
    if (device_sm_count() <= 0) return -2;
    WD_REQUIRE(op.kind == WD_OP_LETTERBOX, "preprocess: unknown kind %d", op.kind);
    const int B = I[0], H = I[1], W = I[2], pad = I[3];
    WD_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && pad >= 0 && pad <= 255, "letterbox: bad shape B=%d H=%d W=%d pad=%d", B, H, W, pad);
    for (int k = 0; k <= 4; ++k) WD_REQUIRE(P[k], "letterbox: null pointer %d", k);
    struct LbOp : CompiledOp {
        const uint8_t* src;
        const int *desc, *coef;
        uint8_t *tmp, *o;
        int B, H, W, pad, gx;
        int launch(cudaStream_t s) override {
            lb_pass1_kernel<<<dim3(gx, B), 256, 0, s>>>(src, desc, coef, tmp);
            lb_pass2_paste_kernel<<<dim3(gx, B), 256, 0, s>>>(tmp, desc, coef, o, H, W, pad);
            WD_CHECK_CUDA(cudaGetLastError());
            count_launch(2);
            return 0;
        }
        int num_kernels() const override { return 2; }
    };
    auto d = std::make_unique<LbOp>();
    d->src = (const uint8_t*)P[0];
    d->desc = (const int*)P[1];
    d->coef = (const int*)P[2];
    d->tmp = (uint8_t*)P[3];
    d->o = (uint8_t*)P[4];
    d->B = B; d->H = H; d->W = W; d->pad = pad;
    d->gx = std::max(8, (device_sm_count() * 16 + B - 1) / B);   // ~16 resident blocks per SM over the whole batch
    out = std::move(d);
    return 0;
}

}  // namespace wd
