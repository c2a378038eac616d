/* postprocess_ref.c — CPU ORACLE (test infrastructure, NOT part of the product path).
 *
 * Plain-C restatement of the reference's detection post-process.  Only tests/, bench.py's
 * cpu_baseline leg and __graft_entry__.smoke() may load this; the product never does.
 *
 * Follows, step by step:
 *   - flatten + sigmoid:            wedetect/models/dense_heads/yolo_world_head.py:654-667
 *                                   generate_proposal.py:1180-1195
 *   - filter_scores_and_topk:       generate_proposal.py:85-131 (copy of mmdet's), called from
 *                                   yolo_world_head.py:721-722 and generate_proposal.py:1201-1202
 *   - priors / decode:              generate_proposal.py:849-905 (MlvlPointGenerator), :1000-1048
 *                                   (distance2bbox), task_modules/coders/distance_point_bbox_coder.py:51-53
 *   - rescale before NMS:           yolo_world_head.py:728-734
 *   - batched NMS:                  mmcv.ops.batched_nms (mmcv 2.1.0, via mmdet _bbox_post_process,
 *                                   yolo_world_head.py:740-744) = coordinate-offset trick + greedy NMS,
 *                                   torchvision.ops.batched_nms (generate_proposal.py:1210) = the same
 *                                   trick when boxes.numel() <= threshold, per-class NMS otherwise
 *   - [:max_per_img], clamp:        yolo_world_head.py:745-746 ; Uni un-letterbox generate_proposal.py:1106-1116
 *
 * Tie-break: the reference sorts with torch.sort(descending=True) which is not documented as
 * stable; this oracle (and the CUDA kernels) define the order as (score desc, flat index asc), i.e.
 * what a stable sort yields.  Scores are sigmoid evaluated in double and rounded once to float.
 * All comparisons that decide an index use single-precision IEEE operations in the reference's
 * operation order; compile with -ffp-contract=off.
 *
 * Uses the same parameter struct as the CUDA library (include/wedetect_b200.h), with HOST pointers.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/wedetect_b200.h"

typedef struct {
    float score;
    uint32_t idx; /* anchor * K + class */
} cand_t;

static int cmp_cand(const void* a, const void* b) {
    const cand_t* x = (const cand_t*)a;
    const cand_t* y = (const cand_t*)b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

static float sigmoid_dr(float x) { return (float)(1.0 / (1.0 + exp(-(double)x))); }

static int iou_gt(const float* a, float area_a, const float* b, float area_b, float thr) {
    const float xx1 = a[0] > b[0] ? a[0] : b[0], yy1 = a[1] > b[1] ? a[1] : b[1];
    const float xx2 = a[2] < b[2] ? a[2] : b[2], yy2 = a[3] < b[3] ? a[3] : b[3];
    float w = xx2 - xx1, h = yy2 - yy1;
    if (!(w > 0.f)) w = 0.f;
    if (!(h > 0.f)) h = 0.f;
    const float inter = w * h;
    const float uni = (area_a + area_b) - inter;
    const float ovr = inter / uni;
    return ovr > thr;
}

int wd_ref_postprocess(const wd_pp_params* p) {
    const int B = p->B, K = p->K;
    int lvl_off[5];
    int A = 0;
    for (int l = 0; l < p->nlevels; ++l) {
        lvl_off[l] = A;
        A += p->lvl_h[l] * p->lvl_w[l];
    }
    lvl_off[p->nlevels] = A;
    if (!p->multi_label) return -1;
    cand_t* cand = (cand_t*)malloc(sizeof(cand_t) * (size_t)A * K);
    const int cap = p->nms_pre;
    float* box = (float*)malloc(sizeof(float) * 4 * cap);
    float* nbox = (float*)malloc(sizeof(float) * 4 * cap);
    float* area = (float*)malloc(sizeof(float) * cap);
    int* label = (int*)malloc(sizeof(int) * cap);
    int* anchor = (int*)malloc(sizeof(int) * cap);
    int* next_kept = (int*)malloc(sizeof(int) * cap); /* per-class singly linked list of kept candidates */
    int* head = (int*)malloc(sizeof(int) * K);
    int* tail = (int*)malloc(sizeof(int) * K);
    if (!cand || !box || !nbox || !area || !label || !anchor || !next_kept || !head || !tail) return -2;

    for (int b = 0; b < B; ++b) {
        /* ---- scores > thr, in flat index order (torch.nonzero order) ---- */
        size_t n = 0;
        for (int l = 0; l < p->nlevels; ++l) {
            const int hw = p->lvl_h[l] * p->lvl_w[l];
            for (int a = 0; a < hw; ++a) {
                const float* row = p->logits[l] + ((size_t)b * hw + a) * p->ld_logit[l];
                for (int k = 0; k < K; ++k) {
                    const float s = sigmoid_dr(row[k]);
                    if (s > p->score_thr) {
                        cand[n].score = s;
                        cand[n].idx = (uint32_t)((lvl_off[l] + a) * K + k);
                        ++n;
                    }
                }
            }
        }
        /* ---- sort descending (stable), keep nms_pre ---- */
        qsort(cand, n, sizeof(cand_t), cmp_cand);
        const int ns = n < (size_t)cap ? (int)n : cap;
        const float* meta = p->img_meta + b * 8;
        float maxc = -INFINITY;
        for (int r = 0; r < ns; ++r) {
            const int an = (int)(cand[r].idx / (uint32_t)K), cls = (int)(cand[r].idx % (uint32_t)K);
            int l = 0;
            while (l + 1 < p->nlevels && an >= lvl_off[l + 1]) ++l;
            const int a = an - lvl_off[l], w = p->lvl_w[l], hw = p->lvl_h[l] * w;
            const float stride = (float)p->lvl_stride[l];
            const float px = ((float)(a % w) + 0.5f) * stride, py = ((float)(a / w) + 0.5f) * stride;
            const float* d = p->dist[l] + ((size_t)b * hw + a) * 4;
            const float dl = d[0] * stride, dt = d[1] * stride, dr = d[2] * stride, db = d[3] * stride;
            float x1 = px - dl, y1 = py - dt, x2 = px + dr, y2 = py + db;
            x1 = (x1 - meta[0]) / meta[2];
            y1 = (y1 - meta[1]) / meta[3];
            x2 = (x2 - meta[0]) / meta[2];
            y2 = (y2 - meta[1]) / meta[3];
            box[r * 4 + 0] = x1; box[r * 4 + 1] = y1; box[r * 4 + 2] = x2; box[r * 4 + 3] = y2;
            label[r] = cls;
            anchor[r] = an;
            const float m1 = x1 > y1 ? x1 : y1, m2 = x2 > y2 ? x2 : y2;
            const float m = m1 > m2 ? m1 : m2;
            if (m > maxc) maxc = m;
        }
        /* ---- batched NMS ---- */
        const int use_off = p->nms_mode == 0 || (4 * (long long)ns <= (long long)p->tv_numel_thr);
        const float step = maxc + 1.0f;
        for (int r = 0; r < ns; ++r) {
            float off = 0.f;
            if (use_off) off = (float)label[r] * step;
            for (int c = 0; c < 4; ++c) nbox[r * 4 + c] = box[r * 4 + c] + off;
            area[r] = (nbox[r * 4 + 2] - nbox[r * 4 + 0]) * (nbox[r * 4 + 3] - nbox[r * 4 + 1]);
        }
        for (int k = 0; k < K; ++k) head[k] = tail[k] = -1;
        int nkeep = 0;
        float* ob = p->out_boxes + (size_t)b * p->max_per_img * 4;
        float* os = p->out_scores + (size_t)b * p->max_per_img;
        int32_t* ol = p->out_labels + (size_t)b * p->max_per_img;
        int32_t* oa = p->out_anchor + (size_t)b * p->max_per_img;
        for (int r = 0; r < ns && nkeep < p->max_per_img; ++r) {
            const int cls = label[r];
            int dead = 0;
            for (int j = head[cls]; j >= 0; j = next_kept[j]) {
                if (iou_gt(nbox + j * 4, area[j], nbox + r * 4, area[r], p->iou_thr)) {
                    dead = 1;
                    break;
                }
            }
            if (dead) continue;
            next_kept[r] = -1;
            if (tail[cls] >= 0) next_kept[tail[cls]] = r; else head[cls] = r;
            tail[cls] = r;
            /* ---- output: un-letterbox (Uni), clamp ---- */
            float x1 = (box[r * 4 + 0] - meta[4]) / meta[6], y1 = (box[r * 4 + 1] - meta[5]) / meta[6];
            float x2 = (box[r * 4 + 2] - meta[4]) / meta[6], y2 = (box[r * 4 + 3] - meta[5]) / meta[6];
            const float cw = p->clamp_wh[b * 2], ch = p->clamp_wh[b * 2 + 1];
            x1 = x1 < 0.f ? 0.f : (x1 > cw ? cw : x1); x2 = x2 < 0.f ? 0.f : (x2 > cw ? cw : x2);
            y1 = y1 < 0.f ? 0.f : (y1 > ch ? ch : y1); y2 = y2 < 0.f ? 0.f : (y2 > ch ? ch : y2);
            ob[nkeep * 4 + 0] = x1; ob[nkeep * 4 + 1] = y1; ob[nkeep * 4 + 2] = x2; ob[nkeep * 4 + 3] = y2;
            os[nkeep] = cand[r].score;
            ol[nkeep] = cls;
            oa[nkeep] = anchor[r];
            ++nkeep;
        }
        p->out_counts[b] = nkeep;
        for (int j = nkeep; j < p->max_per_img; ++j) {
            ob[j * 4 + 0] = ob[j * 4 + 1] = ob[j * 4 + 2] = ob[j * 4 + 3] = 0.f;
            os[j] = 0.f;
            ol[j] = -1;
            oa[j] = -1;
        }
    }
    free(cand); free(box); free(nbox); free(area); free(label); free(anchor); free(next_kept); free(head); free(tail);
    return 0;
}
