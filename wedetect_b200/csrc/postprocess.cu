// postprocess.cu — detection post-process on device, no host synchronisation, bit-exact index logic.
//
//   scores = sigmoid(logits); candidates = {(anchor, class) : score > thr}            (multi-label)
//   order candidates by (score desc, flat index asc)  == stable sort of the reference  -> top nms_pre
//   decode ltrb*stride around the prior, optional rescale, class-aware greedy NMS (IoU > thr, strict),
//   keep the first max_per_img survivors in score order, clamp.
//
// Reference: wedetect/models/dense_heads/yolo_world_head.py:619-749, generate_proposal.py:85-131,
// 1000-1048, 1150-1218; mmdet filter_scores_and_topk; mmcv.ops.batched_nms / torchvision.ops.batched_nms
// (SURVEY.md §8c).  Every floating-point step that feeds a comparison uses explicit round-to-nearest
// intrinsics (no FMA contraction) so the CPU oracle (oracle/postprocess_ref.c) reproduces it bit for bit.
//
// Pipeline (all sizes live on the device):
//   1 pp_candidates   : one pass over the logits, warp-aggregated append of 64-bit keys
//                       key = image | (0x3FFFFFFF - score_bits) | flat_index
//   1b top-k select   : images with more than nms_pre candidates (score_thr 0 of the Uni path, dense score maps, 1203
//                       classes) keep only the keys whose 16 leading score bits reach the nms_pre-th best: two 8-bit
//                       radix-select histograms over the keys + one compaction.  Everything dropped scores strictly below
//                       everything kept and at least nms_pre keys are kept, so the sorted top nms_pre is unchanged - but the
//                       sort now sees ~nms_pre keys per image instead of anchors x classes (2.15 M per image for Uni).
//   2 radix sort #1   : stable LSD radix sort (8-bit digits, warp match_any multisplit) of the kept keys
//   3 pp_segments     : per-image segment starts, n_sel = min(count, nms_pre)
//   4 pp_decode       : box decode / rescale of the selected candidates, max coordinate (mmcv offsets),
//                       second key = image | class | rank
//   5 radix sort #2   : groups candidates by (image, class) keeping score order inside a class
//   6 pp_nms          : one block per (class, image) segment: chunked greedy NMS with 128-bit masks
//   7 pp_finalize     : first max_per_img kept candidates in rank order -> outputs
#include "internal.h"
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

namespace wd {

constexpr int RS_THREADS = 256;
constexpr int RS_ROWS = 16;                       // rows of 32 keys per warp
constexpr int RS_TILE = RS_THREADS * RS_ROWS;     // 4096 keys per block-tile
constexpr int RS_GRID = 148 * 4;

struct PPCtrl {  // lives in device memory
    unsigned int total;      // candidates appended
    unsigned int total2;     // selected candidates (sum of n_sel)
    unsigned int overflow;
    unsigned int totalc;     // candidates left after the top-k select
};

struct PPDev {
    wd_pp_params p;
    int A;                  // anchors per image
    int lvl_off[5];         // anchor offset of each level
    unsigned long long cap; // key capacity
    int idx_bits, b_bits, cls_bits;
    float logit_lo;
    // workspace pointers
    unsigned long long *keys0, *keys1, *keys2a, *keys2b;
    unsigned int* hist;
    unsigned int* totals;   // [32][256] per-pass digit totals (zeroed by pp_reset)
    PPCtrl* ctrl;
    unsigned int *counts, *seg, *nsel, *seg2;
    unsigned int* shist;    // [2][B][256] radix-select histograms (leading / second score digit)
    unsigned int* sel;      // [B][4]: keep_all, leading digit of the cut, candidates still needed inside it, 16-bit cut prefix
    unsigned int* counts2;  // [B] candidates per image after the select
    unsigned int* kept_hist;  // [B][256] confirmed NMS survivors per bucket of 256 ranks (cross-class early exit)
    unsigned int* maxc;     // per image max coordinate (order-preserving uint encoding of float)
    float4* cand_box;
    float* cand_score;
    int *cand_label, *cand_anchor;
    unsigned char* keep;
};

__device__ __forceinline__ float sigmoid_dr(float x) {
    // correctly-rounded-in-practice sigmoid: evaluate in double, round once to float
    return (float)(1.0 / (1.0 + exp(-(double)x)));
}
__device__ __forceinline__ unsigned int float_ord(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_float(unsigned int o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

__global__ void pp_reset_kernel(PPDev d) {
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) {
        d.ctrl->total = 0;
        d.ctrl->total2 = 0;
        d.ctrl->overflow = 0;
        d.ctrl->totalc = 0;
    }
    if (t < d.p.B) {
        d.counts[t] = 0;
        d.counts2[t] = 0;
        d.maxc[t] = 0;  // smaller than the encoding of any float
    }
    if (t < 32 * 256) d.totals[t] = 0;
    for (int i = t; i < 2 * d.p.B * 256; i += gridDim.x * blockDim.x) d.shist[i] = 0;
    for (int i = t; i < d.p.B * 256; i += gridDim.x * blockDim.x) d.kept_hist[i] = 0;
}

// One warp per contiguous range of anchor rows (image, anchor-in-level): lanes stride over the classes (coalesced, no
// per-element division).  Passing keys are compacted into a per-warp shared-memory buffer and flushed with ONE atomic per
// ~160 keys (the global append counter is a single address: an atomic per row serialised the whole kernel); the per-image
// counts are accumulated in a register and flushed when the image changes.
constexpr int PP_WBUF = 192;
__global__ void __launch_bounds__(256) pp_candidates_kernel(PPDev d, int lvl) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ unsigned long long wbuf[8][PP_WBUF];
    const int K = d.p.K;
    const int hw = d.p.lvl_h[lvl] * d.p.lvl_w[lvl];
    const int rows = d.p.B * hw;
    const float* logits = d.p.logits[lvl];
    const int ld = d.p.ld_logit[lvl];
    const float thr = d.p.score_thr;
    const float logit_lo = d.logit_lo;   // conservative float bound: x < logit_lo  =>  sigmoid(x) <= thr, skip the double evaluation
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const int rpw = (rows + nwarps - 1) / nwarps;
    const int r0 = (blockIdx.x * (blockDim.x >> 5) + w) * rpw;
    const int r1 = r0 + rpw < rows ? r0 + rpw : rows;
    unsigned int cnt = 0, img_cnt = 0;   // warp-uniform
    int cur_b = -1;
    auto flush = [&]() {
        if (cnt) {
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&d.ctrl->total, cnt);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (unsigned int i = lane; i < cnt; i += 32) {
                const unsigned long long pos = (unsigned long long)base + i;
                if (pos < d.cap) d.keys0[pos] = wbuf[w][i];
                else d.ctrl->overflow = 1;
            }
            cnt = 0;
            __syncwarp();
        }
    };
    for (int row = r0; row < r1; ++row) {
        const int b = row / hw, a = row - b * hw;
        if (b != cur_b) {
            if (lane == 0 && img_cnt) atomicAdd(&d.counts[cur_b], img_cnt);
            img_cnt = 0;
            cur_b = b;
        }
        const float* lr = logits + (long long)row * ld;
        const unsigned long long hi = (unsigned long long)b << (d.idx_bits + 30);
        const unsigned int idx0 = (unsigned int)(d.lvl_off[lvl] + a) * (unsigned int)K;
        for (int k0 = 0; k0 < K; k0 += 32) {
            const int k = k0 + lane;
            bool pass = false;
            unsigned long long key = 0;
            if (k < K) {
                const float x = lr[k];
                const float sc = x < logit_lo ? 0.f : sigmoid_dr(x);
                if (sc > thr) {
                    pass = true;
                    const unsigned int inv = 0x3FFFFFFFu - (__float_as_uint(sc) & 0x3FFFFFFFu);
                    key = hi | ((unsigned long long)inv << d.idx_bits) | (unsigned long long)(idx0 + k);
                }
            }
            const unsigned int m = __ballot_sync(0xffffffffu, pass);
            if (m) {
                if (pass) wbuf[w][cnt + __popc(m & ((1u << lane) - 1))] = key;
                cnt += __popc(m);
                img_cnt += __popc(m);
                __syncwarp();
                if (cnt > PP_WBUF - 32) flush();
            }
        }
    }
    flush();
    if (lane == 0 && img_cnt) atomicAdd(&d.counts[cur_b], img_cnt);
}

// ---------------- stable LSD radix sort (keys only, 8-bit digit) ----------------
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ n_ptr, int shift,
                                                             unsigned int* __restrict__ hist, unsigned int* __restrict__ totals) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ unsigned int h[256];
    const unsigned int n = *n_ptr;
    const unsigned int num_tiles = (n + RS_TILE - 1) / RS_TILE;
    for (unsigned int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        h[threadIdx.x] = 0;
        __syncthreads();
        const unsigned long long base = (unsigned long long)tile * RS_TILE;
#pragma unroll
        for (int i = 0; i < RS_ROWS; ++i) {
            const unsigned long long idx = base + (unsigned long long)i * RS_THREADS + threadIdx.x;
            if (idx < n) atomicAdd(&h[(unsigned int)(keys[idx] >> shift) & 255u], 1u);
        }
        __syncthreads();
        hist[(unsigned long long)threadIdx.x * num_tiles + tile] = h[threadIdx.x];
        if (h[threadIdx.x]) atomicAdd(&totals[threadIdx.x], h[threadIdx.x]);   // 256 addresses: the digit totals of this pass
        __syncthreads();
    }
}

// Exclusive scan of the (digit-major) tile histograms, one warp per digit: the digit's base is the sum of the totals of all
// smaller digits (accumulated by the histogram kernel), then a coalesced shuffle scan along the digit's row of tiles.
// 256 independent warps instead of one block sweeping the whole table.
__global__ void __launch_bounds__(256) rs_scan_kernel(const unsigned int* __restrict__ n_ptr, unsigned int* __restrict__ hist,
                                                      const unsigned int* __restrict__ totals) {
    pdl_launch_dependents();
    pdl_wait();
    const unsigned int n = *n_ptr;
    const unsigned int num_tiles = (n + RS_TILE - 1) / RS_TILE;
    const int lane = threadIdx.x & 31;
    const unsigned int digit = blockIdx.x * 8 + (threadIdx.x >> 5);
    unsigned int part = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const unsigned int i = lane * 8 + j;
        if (i < digit) part += totals[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    unsigned int run = part;
    unsigned int* row = hist + (unsigned long long)digit * num_tiles;
    for (unsigned int t0 = 0; t0 < num_tiles; t0 += 32) {
        const unsigned int t = t0 + lane;
        const unsigned int v = t < num_tiles ? row[t] : 0u;
        unsigned int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (t < num_tiles) row[t] = run + incl - v;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const unsigned long long* __restrict__ keys, unsigned long long* __restrict__ out,
                                                                const unsigned int* __restrict__ n_ptr, int shift, const unsigned int* __restrict__ hist) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ unsigned int wh[RS_THREADS / 32][256];
    const unsigned int n = *n_ptr;
    const unsigned int num_tiles = (n + RS_TILE - 1) / RS_TILE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int lt = (1u << lane) - 1;
    for (unsigned int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int w = 0; w < RS_THREADS / 32; ++w) wh[w][threadIdx.x] = 0;
        __syncthreads();
        unsigned long long k[RS_ROWS];
        unsigned short rank[RS_ROWS];
        const unsigned long long base = (unsigned long long)tile * RS_TILE + (unsigned long long)warp * (32 * RS_ROWS);
#pragma unroll
        for (int i = 0; i < RS_ROWS; ++i) {
            const unsigned long long idx = base + (unsigned long long)i * 32 + lane;
            const bool valid = idx < n;
            k[i] = valid ? keys[idx] : 0ull;
            const unsigned int dgt = valid ? ((unsigned int)(k[i] >> shift) & 255u) : 256u;
            const unsigned int peers = __match_any_sync(0xffffffffu, dgt);
            unsigned int prev = 0;
            if (valid) prev = wh[warp][dgt];
            __syncwarp();
            if (valid && (peers & lt) == 0) wh[warp][dgt] = prev + __popc(peers);
            __syncwarp();
            rank[i] = (unsigned short)(prev + __popc(peers & lt));
        }
        __syncthreads();
        {
            const unsigned int dgt = threadIdx.x;
            unsigned int run = hist[(unsigned long long)dgt * num_tiles + tile];
#pragma unroll
            for (int w = 0; w < RS_THREADS / 32; ++w) {
                const unsigned int t = wh[w][dgt];
                wh[w][dgt] = run;
                run += t;
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RS_ROWS; ++i) {
            const unsigned long long idx = base + (unsigned long long)i * 32 + lane;
            if (idx < n) {
                const unsigned int dgt = (unsigned int)(k[i] >> shift) & 255u;
                out[wh[warp][dgt] + rank[i]] = k[i];
            }
        }
        __syncthreads();
    }
}

// ---------------- top-k select: drop keys that cannot be among an image's nms_pre best ----------------
// inv = 0x3FFFFFFF - score_bits (30 bits, smaller = better).  Pass 0 histograms inv >> 22 per image, the scan finds the digit
// that holds the nms_pre-th best key; pass 1 histograms (inv >> 14) & 255 inside that digit; the compaction keeps
// inv >> 14 <= cut.  Images with at most nms_pre candidates keep everything.  Histograms are block-private in shared memory
// for a window of 32 images (warp-aggregated with match_any), flushed with one global atomic per non-empty bin.
constexpr int SEL_WIN = 32;
__global__ void __launch_bounds__(256) pp_sel_hist_kernel(PPDev d, int pass) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ unsigned int h[SEL_WIN * 256];
    const unsigned int n = d.ctrl->total < d.cap ? d.ctrl->total : (unsigned int)d.cap;
    const int lane = threadIdx.x & 31;
    unsigned int* gh = d.shist + (size_t)pass * d.p.B * 256;
    const unsigned int tile_keys = 256 * 16;
    const unsigned int num_tiles = (n + tile_keys - 1) / tile_keys;
    for (unsigned int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int i = threadIdx.x; i < SEL_WIN * 256; i += 256) h[i] = 0;
        __syncthreads();
        const unsigned int base = tile * tile_keys;
        const int win0 = d.p.B <= SEL_WIN ? 0 : (int)(d.keys0[base] >> (d.idx_bits + 30));
#pragma unroll 4
        for (int it = 0; it < 16; ++it) {
            const unsigned int i = base + it * 256 + threadIdx.x;
            unsigned int id = 0xFFFFFFFFu;
            if (i < n) {
                const unsigned long long key = d.keys0[i];
                const int b = (int)(key >> (d.idx_bits + 30));
                const unsigned int inv = (unsigned int)(key >> d.idx_bits) & 0x3FFFFFFFu;
                const unsigned int* sl = d.sel + b * 4;
                if (pass == 0) {
                    if (d.counts[b] > (unsigned int)d.p.nms_pre) id = (unsigned int)b * 256u + (inv >> 22);
                } else if (!sl[0] && (inv >> 22) == sl[1]) {
                    id = (unsigned int)b * 256u + ((inv >> 14) & 255u);
                }
            }
            const unsigned int peers = __match_any_sync(0xffffffffu, id);
            if (id != 0xFFFFFFFFu && (peers & ((1u << lane) - 1)) == 0) {
                const int bw = (int)(id >> 8) - win0;
                if (bw >= 0 && bw < SEL_WIN) atomicAdd(&h[bw * 256 + (id & 255u)], (unsigned int)__popc(peers));
                else atomicAdd(&gh[id], (unsigned int)__popc(peers));
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < SEL_WIN * 256; i += 256) {
            const int b = win0 + (i >> 8);
            if (h[i] && b < d.p.B) atomicAdd(&gh[(size_t)b * 256 + (i & 255)], h[i]);
        }
        __syncthreads();
    }
}

// one warp per image: smallest digit whose cumulative count reaches the number still needed
__global__ void __launch_bounds__(256) pp_sel_scan_kernel(PPDev d, int pass) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= d.p.B) return;
    unsigned int* sl = d.sel + b * 4;
    if (pass == 0) {
        if (d.counts[b] <= (unsigned int)d.p.nms_pre) {
            if (lane == 0) { sl[0] = 1; sl[1] = 0; sl[2] = 0; sl[3] = 0xFFFFu; }
            return;
        }
    } else if (sl[0]) {
        return;
    }
    const unsigned int need = pass == 0 ? (unsigned int)d.p.nms_pre : sl[2];
    const unsigned int* hrow = d.shist + ((size_t)pass * d.p.B + b) * 256;
    unsigned int v[8], part = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        v[j] = hrow[lane * 8 + j];
        part += v[j];
    }
    unsigned int incl = part;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    const unsigned int before = incl - part;            // keys in digits below this lane's 8 digits
    const bool mine = before < need && incl >= need;    // exactly one lane (the total is >= need by construction)
    if (mine) {
        unsigned int run = before;
        int dsel = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (run < need && run + v[j] >= need) { dsel = lane * 8 + j; break; }
            run += v[j];
        }
        if (pass == 0) { sl[0] = 0; sl[1] = (unsigned int)dsel; sl[2] = need - run; sl[3] = 0; }
        else sl[3] = (sl[1] << 8) | (unsigned int)dsel;
    }
}

// keys0 -> keys1: keep inv >> 14 <= cut (per image); same per-warp buffered append as pp_candidates
__global__ void __launch_bounds__(256) pp_sel_compact_kernel(PPDev d) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ unsigned long long wbuf[8][PP_WBUF];
    const unsigned int n = d.ctrl->total < d.cap ? d.ctrl->total : (unsigned int)d.cap;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned int nwarps = gridDim.x * 8, gw = blockIdx.x * 8 + w;
    const unsigned int per = ((n + nwarps - 1) / nwarps + 31) / 32 * 32;   // contiguous range per warp: long runs of one image
    const unsigned int i0 = gw * per, i1 = i0 + per < n ? i0 + per : n;
    unsigned int cnt = 0, img_cnt = 0;
    int cur_b = -1;
    auto flush = [&]() {
        if (cnt) {
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&d.ctrl->totalc, cnt);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (unsigned int i = lane; i < cnt; i += 32) d.keys1[base + i] = wbuf[w][i];   // totalc <= total <= cap
            cnt = 0;
            __syncwarp();
        }
    };
    for (unsigned int i = i0; i < i1; i += 32) {
        const unsigned int k = i + lane;
        bool pass = false;
        unsigned long long key = 0;
        int b = -1;
        if (k < i1) {
            key = d.keys0[k];
            b = (int)(key >> (d.idx_bits + 30));
            const unsigned int inv = (unsigned int)(key >> d.idx_bits) & 0x3FFFFFFFu;
            pass = (inv >> 14) <= d.sel[b * 4 + 3];
        }
        // per-image counts: the common case is one image per 32 keys; otherwise count image by image
        const int b0 = __shfl_sync(0xffffffffu, b, 0);
        const bool uniform = __all_sync(0xffffffffu, b == b0 || b < 0);
        const unsigned int m = __ballot_sync(0xffffffffu, pass);
        if (uniform) {
            if (b0 != cur_b) {
                if (lane == 0 && img_cnt) atomicAdd(&d.counts2[cur_b], img_cnt);
                img_cnt = 0;
                cur_b = b0;
            }
            img_cnt += __popc(m);
        } else {
            // a chunk that straddles images: one atomic per image present (an atomic per key serialises on B addresses)
            const unsigned int peers = __match_any_sync(0xffffffffu, pass ? b : -1);
            if (pass && (peers & ((1u << lane) - 1)) == 0) atomicAdd(&d.counts2[b], (unsigned int)__popc(peers));
        }
        if (m) {
            if (pass) wbuf[w][cnt + __popc(m & ((1u << lane) - 1))] = key;
            cnt += __popc(m);
            __syncwarp();
            if (cnt > PP_WBUF - 32) flush();
        }
    }
    flush();
    if (lane == 0 && img_cnt && cur_b >= 0) atomicAdd(&d.counts2[cur_b], img_cnt);
}

// ---------------- segments, decode ----------------
__global__ void pp_segments_kernel(PPDev d) {
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned int run = 0, run2 = 0;
        for (int b = 0; b < d.p.B; ++b) {
            d.seg[b] = run;
            d.seg2[b] = run2;
            const unsigned int c = d.counts2[b];   // after the top-k select: min(c, nms_pre) equals min(all candidates, nms_pre)
            const unsigned int ns = c < (unsigned int)d.p.nms_pre ? c : (unsigned int)d.p.nms_pre;
            d.nsel[b] = ns;
            run += c;
            run2 += ns;
        }
        d.seg[d.p.B] = run;
        d.seg2[d.p.B] = run2;
        d.ctrl->total2 = run2;
    }
}

__global__ void __launch_bounds__(256) pp_decode_kernel(PPDev d, const unsigned long long* __restrict__ sorted) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y;
    const unsigned int ns = d.nsel[b];
    const int K = d.p.K;
    const float* meta = d.p.img_meta + b * 8;
    const float psx = meta[0], psy = meta[1], pdx = meta[2], pdy = meta[3];
    float local_max = -INFINITY;
    for (unsigned int r = blockIdx.x * blockDim.x + threadIdx.x; r < ns; r += gridDim.x * blockDim.x) {
        const unsigned long long key = sorted[d.seg[b] + r];
        const unsigned int idx = (unsigned int)(key & ((1ull << d.idx_bits) - 1));
        const unsigned int inv = (unsigned int)(key >> d.idx_bits) & 0x3FFFFFFFu;
        const float score = __uint_as_float(0x3FFFFFFFu - inv);
        const int anchor = idx / K, cls = idx % K;
        int lvl = 0;
        while (lvl + 1 < d.p.nlevels && anchor >= d.lvl_off[lvl + 1]) ++lvl;
        const int a = anchor - d.lvl_off[lvl];
        const int w = d.p.lvl_w[lvl], hw = d.p.lvl_h[lvl] * w;
        const float stride = (float)d.p.lvl_stride[lvl];
        // MlvlPointGenerator: (i + 0.5) * stride (generate_proposal.py:880-892)
        const float px = __fmul_rn((float)(a % w) + 0.5f, stride);
        const float py = __fmul_rn((float)(a / w) + 0.5f, stride);
        const float4 lt = *reinterpret_cast<const float4*>(d.p.dist[lvl] + ((long long)b * hw + a) * 4);
        // flatten_bbox_preds * stride, then distance2bbox (yolo_world_head.py:654-667, coder :51-53)
        float x1 = __fsub_rn(px, __fmul_rn(lt.x, stride));
        float y1 = __fsub_rn(py, __fmul_rn(lt.y, stride));
        float x2 = __fadd_rn(px, __fmul_rn(lt.z, stride));
        float y2 = __fadd_rn(py, __fmul_rn(lt.w, stride));
        // mmdet rescale before NMS (yolo_world_head.py:728-734); identity when sub = 0, div = 1
        x1 = __fdiv_rn(__fsub_rn(x1, psx), pdx);
        y1 = __fdiv_rn(__fsub_rn(y1, psy), pdy);
        x2 = __fdiv_rn(__fsub_rn(x2, psx), pdx);
        y2 = __fdiv_rn(__fsub_rn(y2, psy), pdy);
        const unsigned int o = d.seg2[b] + r;
        d.cand_box[o] = make_float4(x1, y1, x2, y2);
        d.cand_score[o] = score;
        d.cand_label[o] = cls;
        d.cand_anchor[o] = anchor;
        d.keep[o] = 0;
        d.keys2a[o] = ((unsigned long long)b << (16 + d.cls_bits)) | ((unsigned long long)cls << 16) | r;
        local_max = fmaxf(local_max, fmaxf(fmaxf(x1, y1), fmaxf(x2, y2)));
    }
    local_max = warp_max(local_max);
    if ((threadIdx.x & 31) == 0 && local_max > -INFINITY) atomicMax(&d.maxc[b], float_ord(local_max));
}

// ---------------- class-aware greedy NMS: one block per (class, image) ----------------
// IoU > thr with the reference's arithmetic (inter / union in fp32, round-to-nearest, strict compare).  The division is
// only evaluated when the outcome is not already decided: no overlap -> the quotient is 0 (or NaN): false; otherwise
// inter vs thr * union with a 1e-6 relative margin (>> the 2^-24 rounding of either side) decides all but razor-edge pairs.
__device__ __forceinline__ bool iou_gt(const float4 a, float area_a, const float4 b, float area_b, float thr) {
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y), xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(w, h);
    if (thr > 0.f && !(inter > 0.f)) return false;
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    const float q = __fmul_rn(thr, uni);
    if (q > 1e-30f && q < 1e30f) {
        if (inter > __fmul_rn(q, 1.000001f)) return true;
        if (inter < __fmul_rn(q, 0.999999f)) return false;
    }
    return __fdiv_rn(inter, uni) > thr;
}

__global__ void __launch_bounds__(128) pp_nms_kernel(PPDev d, const unsigned long long* __restrict__ sorted2, float4* __restrict__ kept_scratch) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float4 cbox[128];
    __shared__ float carea[128];
    __shared__ unsigned int cmask[128][4];
    __shared__ unsigned char calive[128];
    __shared__ int s_nk;
    __shared__ unsigned int s_sum;
    const int cls = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const unsigned int lo_b = d.seg2[b], hi_b = d.seg2[b + 1];
    if (lo_b == hi_b) return;
    // segment of this class inside the image's (class, rank)-sorted key range
    const unsigned long long k_lo = ((unsigned long long)b << (16 + d.cls_bits)) | ((unsigned long long)cls << 16);
    const unsigned long long k_hi = k_lo + (1ull << 16);
    unsigned int lo = lo_b, hi = hi_b;
    while (lo < hi) {
        const unsigned int mid = (lo + hi) >> 1;
        if (sorted2[mid] < k_lo) lo = mid + 1; else hi = mid;
    }
    const unsigned int s0 = lo;
    hi = hi_b;
    while (lo < hi) {
        const unsigned int mid = (lo + hi) >> 1;
        if (sorted2[mid] < k_hi) lo = mid + 1; else hi = mid;
    }
    const unsigned int s1 = lo;
    const int n = (int)(s1 - s0);
    if (n == 0) return;

    // coordinate offsets (mmcv: always; torchvision: only when 4 * n_img <= tv_numel_thr)
    const unsigned int n_img = hi_b - lo_b;
    bool use_off = d.p.nms_mode == 0 || (4u * n_img <= (unsigned int)d.p.tv_numel_thr);
    float off = 0.f;
    if (use_off) off = __fmul_rn((float)cls, __fadd_rn(ord_float(d.maxc[b]), 1.0f));
    const float thr = d.p.iou_thr;
    float4* kept = kept_scratch + s0;  // at most n kept boxes, private to this segment
    if (tid == 0) s_nk = 0;
    __syncthreads();

    for (int base = 0; base < n; base += 128) {
        // a class can contribute at most max_per_img detections to the image's final top-max_per_img (its kept
        // candidates are in score order), so the rest of the segment cannot matter: exact early exit
        if (s_nk >= d.p.max_per_img) break;
        // Exact cross-class early exit: the image's output is its first max_per_img survivors in rank (= score) order.  Every
        // block publishes its confirmed survivors per bucket of 256 ranks; once max_per_img survivors are known in buckets that
        // lie entirely before this chunk's first rank, nothing this block could still keep can reach the output.  (Which
        // blocks get to skip work depends on timing; the output does not.)
        {
            const unsigned int lim = (unsigned int)(sorted2[s0 + base] & 0xFFFFu) >> 8;
            if (tid == 0) s_sum = 0;
            __syncthreads();
            const volatile unsigned int* kh = d.kept_hist + b * 256;
            unsigned int v = 0;
            if ((unsigned int)tid < lim) v += kh[tid];
            if ((unsigned int)tid + 128u < lim) v += kh[tid + 128];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((tid & 31) == 0 && v) atomicAdd(&s_sum, v);
            __syncthreads();
            if (s_sum >= (unsigned int)d.p.max_per_img) break;
        }
        const int t = base + tid;
        const bool valid = t < n;
        float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned int slot = 0, r = 0;
        if (valid) {
            r = (unsigned int)(sorted2[s0 + t] & 0xFFFFu);
            slot = lo_b + r;
            const float4 bx = d.cand_box[slot];
            mine = make_float4(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off), __fadd_rn(bx.z, off), __fadd_rn(bx.w, off));
        }
        const float area = __fmul_rn(__fsub_rn(mine.z, mine.x), __fsub_rn(mine.w, mine.y));
        bool alive = valid;
        const int nk = s_nk;
        for (int i = 0; i < nk && alive; ++i) {
            const float4 kb = kept[i];
            const float ka = __fmul_rn(__fsub_rn(kb.z, kb.x), __fsub_rn(kb.w, kb.y));
            if (iou_gt(kb, ka, mine, area, thr)) alive = false;
        }
        cbox[tid] = mine;
        carea[tid] = area;
        calive[tid] = alive ? 1 : 0;
        __syncthreads();
        unsigned int m[4] = {0, 0, 0, 0};
        if (alive) {
            for (int j = tid + 1; j < 128 && base + j < n; ++j)
                if (calive[j] && iou_gt(mine, area, cbox[j], carea[j], thr)) m[j >> 5] |= 1u << (j & 31);
        }
        cmask[tid][0] = m[0]; cmask[tid][1] = m[1]; cmask[tid][2] = m[2]; cmask[tid][3] = m[3];
        __syncthreads();
        if (tid == 0) {
            unsigned int rem[4] = {0, 0, 0, 0};
            int nk2 = s_nk;
            const int lim = (n - base) < 128 ? (n - base) : 128;
            for (int j = 0; j < lim; ++j) {
                if (calive[j] && !((rem[j >> 5] >> (j & 31)) & 1u)) {
                    kept[nk2++] = cbox[j];
                    calive[j] = 2;  // kept
                    rem[0] |= cmask[j][0]; rem[1] |= cmask[j][1]; rem[2] |= cmask[j][2]; rem[3] |= cmask[j][3];
                }
            }
            s_nk = nk2;
        }
        __syncthreads();
        if (valid && calive[tid] == 2) {
            d.keep[slot] = 1;
            atomicAdd(&d.kept_hist[b * 256 + (r >> 8)], 1u);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) pp_finalize_kernel(PPDev d) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int ns = d.nsel[b], lo_b = d.seg2[b];
    const int maxk = d.p.max_per_img;
    const float* meta = d.p.img_meta + b * 8;
    const float qsx = meta[4], qsy = meta[5], qd = meta[6];
    const float cw = d.p.clamp_wh[b * 2 + 0], chh = d.p.clamp_wh[b * 2 + 1];
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (unsigned int r0 = 0; r0 < ns; r0 += 256) {
        const unsigned int r = r0 + tid;
        const bool k = r < ns && d.keep[lo_b + r] != 0;
        const unsigned int m = __ballot_sync(0xffffffffu, k);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        const int pos = off + __popc(m & ((1u << lane) - 1));
        if (k && pos < maxk) {
            const float4 bx = d.cand_box[lo_b + r];
            float x1 = __fdiv_rn(__fsub_rn(bx.x, qsx), qd), y1 = __fdiv_rn(__fsub_rn(bx.y, qsy), qd);
            float x2 = __fdiv_rn(__fsub_rn(bx.z, qsx), qd), y2 = __fdiv_rn(__fsub_rn(bx.w, qsy), qd);
            x1 = fminf(fmaxf(x1, 0.f), cw); x2 = fminf(fmaxf(x2, 0.f), cw);
            y1 = fminf(fmaxf(y1, 0.f), chh); y2 = fminf(fmaxf(y2, 0.f), chh);
            const long long o = (long long)b * maxk + pos;
            *reinterpret_cast<float4*>(d.p.out_boxes + o * 4) = make_float4(x1, y1, x2, y2);
            d.p.out_scores[o] = d.cand_score[lo_b + r];
            d.p.out_labels[o] = d.cand_label[lo_b + r];
            d.p.out_anchor[o] = d.cand_anchor[lo_b + r];
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < 8; ++w) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
        if (s_base >= maxk) break;
    }
    const int cnt = s_base < maxk ? s_base : maxk;
    if (tid == 0) d.p.out_counts[b] = cnt;
    for (int j = cnt + tid; j < maxk; j += 256) {
        const long long o = (long long)b * maxk + j;
        *reinterpret_cast<float4*>(d.p.out_boxes + o * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        d.p.out_scores[o] = 0.f;
        d.p.out_labels[o] = -1;
        d.p.out_anchor[o] = -1;
    }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static int bits_for(unsigned long long n) {  // bits needed to represent values < n
    int b = 1;
    while ((1ull << b) < n) ++b;
    return b;
}
static unsigned long long align256(unsigned long long x) { return (x + 255) & ~255ull; }

struct PPLayout {
    unsigned long long off_keys0, off_keys1, off_hist, off_ctrl, off_counts, off_seg, off_nsel, off_seg2, off_maxc, off_box, off_score, off_label,
        off_anchor, off_keys2a, off_keys2b, off_keep, off_kept, off_shist, off_sel, off_counts2, off_kept_hist, total;
};
static PPLayout pp_layout(int B, int A, int K, int nms_pre) {
    PPLayout L;
    const unsigned long long cap = (unsigned long long)B * A * K;
    const unsigned long long sel = (unsigned long long)B * nms_pre < cap ? (unsigned long long)B * nms_pre : cap;
    const unsigned long long tiles = (cap + RS_TILE - 1) / RS_TILE + 1;
    unsigned long long o = 0;
    L.off_keys0 = o; o = align256(o + cap * 8);
    L.off_keys1 = o; o = align256(o + cap * 8);
    L.off_hist = o; o = align256(o + (tiles + 32) * 256 * 4);   // tile histograms + 32 passes of digit totals
    L.off_ctrl = o; o = align256(o + sizeof(PPCtrl));
    L.off_counts = o; o = align256(o + (B + 1) * 4ull);
    L.off_seg = o; o = align256(o + (B + 1) * 4ull);
    L.off_nsel = o; o = align256(o + (B + 1) * 4ull);
    L.off_seg2 = o; o = align256(o + (B + 1) * 4ull);
    L.off_maxc = o; o = align256(o + (B + 1) * 4ull);
    L.off_box = o; o = align256(o + sel * 16);
    L.off_score = o; o = align256(o + sel * 4);
    L.off_label = o; o = align256(o + sel * 4);
    L.off_anchor = o; o = align256(o + sel * 4);
    L.off_keys2a = o; o = align256(o + sel * 8);
    L.off_keys2b = o; o = align256(o + sel * 8);
    L.off_keep = o; o = align256(o + sel);
    L.off_kept = o; o = align256(o + sel * 16);
    L.off_shist = o; o = align256(o + 2ull * B * 256 * 4);
    L.off_sel = o; o = align256(o + (B + 1) * 16ull);
    L.off_counts2 = o; o = align256(o + (B + 1) * 4ull);
    L.off_kept_hist = o; o = align256(o + (unsigned long long)B * 256 * 4);
    L.total = o;
    return L;
}

struct PostOp : CompiledOp {
    PPDev d;
    int passes1, passes2, kernels;
    int num_kernels() const override { return kernels; }
    int sort(unsigned long long* a, unsigned long long* b, const unsigned int* n_ptr, int passes, int slot0, cudaStream_t s) {
        for (int p = 0; p < passes; ++p) {
            const unsigned long long* src = (p & 1) ? b : a;
            unsigned long long* dst = (p & 1) ? a : b;
            unsigned int* tot = d.totals + (size_t)(slot0 + p) * 256;
            launch_pdl(rs_hist_kernel, dim3(RS_GRID), dim3(RS_THREADS), (size_t)0, s, 1, src, n_ptr, p * 8, d.hist, tot);
            launch_pdl(rs_scan_kernel, dim3(32), dim3(256), (size_t)0, s, 1, n_ptr, d.hist, tot);
            launch_pdl(rs_scatter_kernel, dim3(RS_GRID), dim3(RS_THREADS), (size_t)0, s, 1, src, dst, n_ptr, p * 8, d.hist);
            count_launch(3);
        }
        WD_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    int launch(cudaStream_t s) override {
        // WD_PP_PROFILE=1: synchronous per-phase timing printed to stderr (debug / tuning aid, never in benchmarks)
        static const bool prof = getenv("WD_PP_PROFILE") != nullptr;
        cudaEvent_t ev[10];
        int ne = 0;
        auto mark = [&]() {
            if (prof) {
                cudaEventCreate(&ev[ne]);
                cudaEventRecord(ev[ne++], s);
            }
        };
        mark();
        launch_pdl(pp_reset_kernel, dim3((d.p.B + 255) / 256 > 32 ? (d.p.B + 255) / 256 : 32), dim3(256), (size_t)0, s, 1, d);
        count_launch();
        for (int l = 0; l < d.p.nlevels; ++l) {
            const long long rows_l = (long long)d.p.B * d.p.lvl_h[l] * d.p.lvl_w[l];
            long long g = (rows_l + 63) / 64;   // >= 8 rows per warp, 8 warps per block
            if (g > 148 * 8) g = 148 * 8;
            if (g < 1) g = 1;
            launch_pdl(pp_candidates_kernel, dim3((int)g), dim3(256), (size_t)0, s, 1, d, l);
            count_launch();
        }
        mark();
        // top-k select (keys0 -> keys1), then the sort ping-pongs starting from keys1
        for (int pass = 0; pass < 2; ++pass) {
            launch_pdl(pp_sel_hist_kernel, dim3(RS_GRID), dim3(256), (size_t)0, s, 1, d, pass);
            launch_pdl(pp_sel_scan_kernel, dim3((d.p.B + 7) / 8), dim3(256), (size_t)0, s, 1, d, pass);
        }
        launch_pdl(pp_sel_compact_kernel, dim3(RS_GRID), dim3(256), (size_t)0, s, 1, d);
        count_launch(5);
        mark();
        if (sort(d.keys1, d.keys0, &d.ctrl->totalc, passes1, 0, s)) return -2;
        mark();
        const unsigned long long* sorted1 = (passes1 & 1) ? d.keys0 : d.keys1;
        launch_pdl(pp_segments_kernel, dim3(1), dim3(32), (size_t)0, s, 1, d);
        launch_pdl(pp_decode_kernel, dim3(dim3(32, d.p.B)), dim3(256), (size_t)0, s, 1, d, sorted1);
        count_launch(2);
        mark();
        if (sort(d.keys2a, d.keys2b, &d.ctrl->total2, passes2, 16, s)) return -2;
        mark();
        const unsigned long long* sorted2 = (passes2 & 1) ? d.keys2b : d.keys2a;
        launch_pdl(pp_nms_kernel, dim3(dim3(d.p.K, d.p.B)), dim3(128), (size_t)0, s, 1, d, sorted2, kept);
        mark();
        launch_pdl(pp_finalize_kernel, dim3(d.p.B), dim3(256), (size_t)0, s, 1, d);
        count_launch(2);
        mark();
        WD_CHECK_CUDA(cudaGetLastError());
        if (prof) {
            cudaStreamSynchronize(s);
            static const char* names[] = {"candidates", "select", "sort1", "segments+decode", "sort2", "nms", "finalize"};
            float tot = 0.f;
            for (int i = 0; i + 1 < ne; ++i) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
                fprintf(stderr, "[pp] %-16s %.3f ms\n", names[i], ms);
                tot += ms;
            }
            PPCtrl hc;
            cudaMemcpy(&hc, d.ctrl, sizeof(hc), cudaMemcpyDeviceToHost);
            fprintf(stderr, "[pp] total %.3f ms (passes %d + %d), candidates %u, after top-k select %u, selected %u\n", tot, passes1, passes2, hc.total,
                    hc.totalc, hc.total2);
            for (int i = 0; i < ne; ++i) cudaEventDestroy(ev[i]);
        }
        return 0;
    }
    float4* kept;
};

int compile_postprocess(const wd_op& op, std::unique_ptr<CompiledOp>& out) {
    WD_REQUIRE(op.p[0], "postprocess: p[0] must point to a host wd_pp_params");
    if (device_sm_count() <= 0) return -2;
    auto o = std::make_unique<PostOp>();
    memset(&o->d, 0, sizeof(o->d));
    o->d.p = *reinterpret_cast<const wd_pp_params*>(op.p[0]);
    const wd_pp_params& p = o->d.p;
    WD_REQUIRE(p.B > 0 && p.B <= 1024 && p.K > 0 && p.nlevels >= 1 && p.nlevels <= 4, "postprocess: bad B/K/nlevels");
    WD_REQUIRE(p.multi_label == 1, "postprocess: only multi_label=True is implemented (all shipped configs)");
    WD_REQUIRE(p.nms_pre > 0 && p.nms_pre <= 65535 && p.max_per_img > 0, "postprocess: nms_pre must be in [1, 65535]");
    WD_REQUIRE(p.nms_mode == 0 || p.nms_mode == 1, "postprocess: bad nms_mode");
    int A = 0;
    for (int l = 0; l < p.nlevels; ++l) {
        WD_REQUIRE(p.logits[l] && p.dist[l] && p.lvl_h[l] > 0 && p.lvl_w[l] > 0 && p.ld_logit[l] >= p.K, "postprocess: bad level %d", l);
        o->d.lvl_off[l] = A;
        A += p.lvl_h[l] * p.lvl_w[l];
    }
    o->d.lvl_off[p.nlevels] = A;
    o->d.A = A;
    WD_REQUIRE(p.img_meta && p.clamp_wh && p.out_boxes && p.out_scores && p.out_labels && p.out_anchor && p.out_counts, "postprocess: null pointer");
    const unsigned long long cap = (unsigned long long)p.B * A * p.K;
    WD_REQUIRE(cap < (1ull << 32), "postprocess: B*A*K too large");
    o->d.cap = cap;
    {
        // sigmoid is monotone: scores > thr need logits > log(thr/(1-thr)); keep a safety margin far above any rounding
        const double t = (double)p.score_thr;
        double lo = -120.0;                                   // thr <= 0: every representable positive score passes
        if (t >= 1.0) lo = 120.0;
        else if (t > 0.0) { lo = log(t / (1.0 - t)); lo -= 1e-3 * (fabs(lo) > 1.0 ? fabs(lo) : 1.0); }
        o->d.logit_lo = (float)lo;
    }
    o->d.idx_bits = bits_for((unsigned long long)A * p.K);
    o->d.b_bits = bits_for(p.B);
    o->d.cls_bits = bits_for(p.K);
    WD_REQUIRE(o->d.idx_bits + 30 + o->d.b_bits <= 64, "postprocess: key overflow");
    const PPLayout L = pp_layout(p.B, A, p.K, p.nms_pre);
    WD_REQUIRE(p.workspace && p.workspace_bytes >= L.total, "postprocess: workspace too small (%llu < %llu)", (unsigned long long)p.workspace_bytes,
               (unsigned long long)L.total);
    WD_REQUIRE((reinterpret_cast<uintptr_t>(p.workspace) & 255) == 0, "postprocess: workspace must be 256-byte aligned");
    uint8_t* w = reinterpret_cast<uint8_t*>(p.workspace);
    o->d.keys0 = (unsigned long long*)(w + L.off_keys0);
    o->d.keys1 = (unsigned long long*)(w + L.off_keys1);
    o->d.hist = (unsigned int*)(w + L.off_hist);
    o->d.totals = o->d.hist + (((unsigned long long)p.B * A * p.K + RS_TILE - 1) / RS_TILE + 1) * 256;
    o->d.ctrl = (PPCtrl*)(w + L.off_ctrl);
    o->d.counts = (unsigned int*)(w + L.off_counts);
    o->d.seg = (unsigned int*)(w + L.off_seg);
    o->d.nsel = (unsigned int*)(w + L.off_nsel);
    o->d.seg2 = (unsigned int*)(w + L.off_seg2);
    o->d.maxc = (unsigned int*)(w + L.off_maxc);
    o->d.cand_box = (float4*)(w + L.off_box);
    o->d.cand_score = (float*)(w + L.off_score);
    o->d.cand_label = (int*)(w + L.off_label);
    o->d.cand_anchor = (int*)(w + L.off_anchor);
    o->d.keys2a = (unsigned long long*)(w + L.off_keys2a);
    o->d.keys2b = (unsigned long long*)(w + L.off_keys2b);
    o->d.keep = (unsigned char*)(w + L.off_keep);
    o->kept = (float4*)(w + L.off_kept);
    o->d.shist = (unsigned int*)(w + L.off_shist);
    o->d.sel = (unsigned int*)(w + L.off_sel);
    o->d.counts2 = (unsigned int*)(w + L.off_counts2);
    o->d.kept_hist = (unsigned int*)(w + L.off_kept_hist);
    o->passes1 = (o->d.idx_bits + 30 + o->d.b_bits + 7) / 8;
    o->passes2 = (16 + o->d.cls_bits + o->d.b_bits + 7) / 8;
    o->kernels = 1 + p.nlevels + 5 + 3 * (o->passes1 + o->passes2) + 4;
    out = std::move(o);
    return 0;
}

}  // namespace wd

extern "C" uint64_t wd_pp_workspace_bytes(int B, int anchors, int K, int nms_pre) {
    if (B <= 0 || anchors <= 0 || K <= 0 || nms_pre <= 0) return 0;
    return wd::pp_layout(B, anchors, K, nms_pre).total;
}
