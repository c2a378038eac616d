"""Lowering of the WeDetect forward into a flat `wd_op` program (the host-side "graph builder").

One `VisionPlan` = one (model size, batch, resolution, class count) instance: it owns every activation
buffer in HBM (NHWC; bf16 GEMM operands, fp32 residual stream / logits / boxes), the op list and the
compiled `Program` (prebuilt TMA descriptors, optionally a CUDA graph).  Concats are never materialised:
producers TMA-store straight into channel slices of the consumer's input buffer.

Stage map (reference file:line in brackets):
  stem      patch gather -> GEMM(48->C0) -> LN                      [mm_backbone.py:188-191]
  block     dw7x7+LN -> GEMM(C->4C)+GELU -> GEMM(4C->C)*gamma+x     [mm_backbone.py:112-125]
  downsamp  LN + space-to-depth -> GEMM(4C->2C)                      [mm_backbone.py:193-198]
  neck      1x1 / 3x3 implicit GEMMs with folded BN + ReLU/SiLU, BottleRep alpha-residual in the
            epilogue, ConvTranspose as scatter-store GEMMs            [yolo_world_pafpn.py:1114-1137]
  head      cls stack -> 768-d region embeddings -> similarity GEMM against the folded text matrix;
            reg stack -> DFL in the epilogue                         [yolo_world_head.py:263-294]
  post      sigmoid / threshold / top-k / decode / class-aware NMS   [yolo_world_head.py:619-749]
"""
import torch

from . import _lib as L
from . import ops, schema
from .ops import P3

F32 = torch.float32


class Act:
    """A 16-bit activation [rows, C] with NHWC geometry; `ps` > 0 = fp16 hi/lo planes at `scale` (ops.P3), else bf16."""

    def __init__(self, t, ps, B, H, W, scale=1.0):
        self.t, self.ps, self.B, self.H, self.W, self.scale = t, ps, B, H, W, scale

    @property
    def C(self):
        return self.t.shape[1]

    def cols(self, c0, c1):
        return Act(self.t[:, c0:c1], self.ps, self.B, self.H, self.W, self.scale)

    @property
    def p3(self):
        return P3(self.t, self.ps, self.scale)

    @property
    def p3_4d(self):
        t = self.t
        ld = t.stride(0)
        v = t.as_strided((self.B, self.H, self.W, t.shape[1]), (self.H * self.W * ld, self.W * ld, ld, 1), t.storage_offset())
        return P3(v, self.ps, self.scale)

    def value(self):
        return self.p3.value()


class VisionPlan:
    def __init__(self, weights, size, B, H, W, *, K, uni=False, input_dtype=torch.float32, score_thr=0.001,
                 nms_pre=30000, iou_thr=0.7, max_per_img=300, nms_mode=0, tv_numel_thr=20000, extract=False, device="cuda:0"):
        assert H % 32 == 0 and W % 32 == 0
        self.Wt, self.size, self.B, self.H, self.W, self.K, self.uni = weights, size, B, H, W, K, uni
        self.dev = torch.device(device)
        self.precise = weights.precise
        self.extract = bool(extract)
        assert uni or not extract, "extract (labels / scales / bias per proposal) is a Uni-path output"
        self.cfg = schema.SIZES[size]
        self.ops = []
        self.keep = []
        self.bufs = {}
        self.K_pad = (K + 7) // 8 * 8
        self.level_hw = schema.level_hw(H, W)
        self.image = torch.zeros(B, 3, H, W, dtype=input_dtype, device=self.dev)
        self._build_backbone()
        self._build_neck()
        self._build_head()
        self._build_post(score_thr, nms_pre, iou_thr, max_per_img, nms_mode, tv_numel_thr)
        self.program = L.Program(self.ops, keepalive=self.keep)
        self._fold_program = None
        self._graph = False
        if uni:
            self.set_text(self.Wt["prompts"], normalize=False)

    # ------------------------------------------------------------------ buffers
    def _f32(self, name, rows, C):
        t = torch.zeros(rows, C, dtype=F32, device=self.dev)
        self.bufs[name] = t
        return t

    def _act(self, name, B, H, W, C):
        q = P3.zeros((B * H * W, C), self.dev, self.precise)
        a = Act(q.t, q.ps, B, H, W, q.scale)
        self.bufs[name] = a
        return a

    def _view_act(self, scratch, rows, C, B, H, W):
        """A [rows, C] activation carved out of a flat scratch P3 (all planes keep the scratch's plane stride)."""
        return Act(scratch.t[: rows * C].view(rows, C), scratch.ps, B, H, W, scratch.scale)

    def _mat(self, name):
        return self.Wt.mat(name)

    # ------------------------------------------------------------------ op helpers
    def _linear(self, a, wname, out, *, bias=None, act=L.ACT_NONE, gamma=None, resid=None, alpha=1.0, dfl=False):
        C = out.p3 if isinstance(out, Act) else out
        r = resid.p3 if isinstance(resid, Act) else resid
        self.ops.append(ops.linear(a.p3, self._mat(wname), C, bias=bias, gamma=gamma, resid=r, alpha=alpha, act=act, dfl=dfl))

    def _conv3x3(self, a, wname, out, *, bias, act, resid=None, alpha=1.0):
        self.ops.append(ops.conv3x3(a.p3_4d, self._mat(wname), out.p3_4d, bias=bias, act=act,
                                    resid=None if resid is None else resid.p3_4d, alpha=alpha))

    def _conv3x3_s2(self, a, wname, out, *, bias, act):
        """3x3 stride-2 (4 sites in the neck): implicit GEMM, the A tensor map walks the input with element strides 2."""
        Ho, Wo = (a.H - 1) // 2 + 1, (a.W - 1) // 2 + 1
        o4 = out.p3_4d
        assert tuple(o4.t.shape[:3]) == (a.B, Ho, Wo), (tuple(o4.t.shape), a.B, Ho, Wo)
        self.ops.append(ops.conv3x3(a.p3_4d, self._mat(wname), o4, bias=bias, act=act, stride=2))

    def _cba1x1(self, a, name, out, act):
        self._linear(a, f"neck.{name}.w", out, bias=self.Wt[f"neck.{name}.b"], act=act)

    # ------------------------------------------------------------------ backbone
    def _build_backbone(self):
        W_, cfg, B = self.Wt, self.cfg, self.B
        dims, depths = cfg["dims"], cfg["depths"]
        hs = [self.H // 4, self.H // 8, self.H // 16, self.H // 32]
        ws = [self.W // 4, self.W // 8, self.W // 16, self.W // 32]
        rows = [B * h * w for h, w in zip(hs, ws)]
        # scratch shared by all stages (stage 0 is the largest)
        scratch_ln = P3.zeros((max(r * d for r, d in zip(rows, dims)),), self.dev, self.precise)
        scratch_hid = P3.zeros((max(r * 4 * d for r, d in zip(rows, dims)),), self.dev, self.precise)
        scratch_dw = torch.zeros(max(r * d for r, d in zip(rows, dims)), dtype=F32, device=self.dev)   # depthwise conv output, pre-LN
        self.keep += [scratch_ln, scratch_hid, scratch_dw]
        # stem
        patch = self._act("stem.patch", B, hs[0], ws[0], 64)
        self.ops.append(ops.stem_patch(self.image, patch.p3, 1.0))
        x = self._f32("x0", rows[0], dims[0])
        self._linear(patch, "stem.w", x, bias=W_["stem.b"])
        self.ops.append(ops.ln_rows(x, W_["stem.ln_w"], W_["stem.ln_b"], schema.LN_EPS, out_f32=x))
        self.c_feats = []
        for s in range(4):
            C, M = dims[s], rows[s]
            if s > 0:
                s2d = self._view_act(scratch_hid, M, 4 * dims[s - 1], B, hs[s], ws[s])
                self.ops.append(ops.ln_rows(x, W_[f"down{s}.ln_w"], W_[f"down{s}.ln_b"], schema.LN_EPS, out_bf16=s2d.p3,
                                            s2d_hw=(hs[s - 1], ws[s - 1])))
                x = self._f32(f"x{s}", M, C)
                self._linear(s2d, f"down{s}.w", x, bias=W_[f"down{s}.b"])
            t_ln = self._view_act(scratch_ln, M, C, B, hs[s], ws[s])
            t_hid = self._view_act(scratch_hid, M, 4 * C, B, hs[s], ws[s])
            x4 = x.view(B, hs[s], ws[s], C)
            for j in range(depths[s]):
                q = f"s{s}.b{j}."
                self.ops.append(ops.dwconv_ln(x4, t_ln.p3, W_[q + "dw_w"], W_[q + "dw_b"], W_[q + "ln_w"], W_[q + "ln_b"], schema.LN_EPS, scratch=scratch_dw))
                if ops.mlp_fused_ok(t_ln.p3, self._mat(q + "w1"), self._mat(q + "w2"), x):
                    # C = 128: pw1 -> GELU -> pw2 -> LayerScale -> residual in one kernel, the hidden tile stays on chip
                    self.ops.append(ops.mlp_fused(t_ln.p3, self._mat(q + "w1"), self._mat(q + "w2"), W_[q + "b1"], W_[q + "b2"], W_[q + "gamma"], x))
                    continue
                self._linear(t_ln, q + "w1", t_hid, bias=W_[q + "b1"], act=L.ACT_GELU)
                self._linear(t_hid, q + "w2", x, bias=W_[q + "b2"], gamma=W_[q + "gamma"], resid=x, alpha=1.0)
            c = self._act(f"c{s + 1}", B, hs[s], ws[s], C)
            self.ops.append(ops.cast_bf16(x, c.p3))
            self.c_feats.append(c)
        self.stage_x = [self.bufs[f"x{s}"] for s in range(4)]

    # ------------------------------------------------------------------ neck
    def _bepc3(self, name, x, out, n):
        """CSPStackRep (yolo_world_pafpn.py:631-647): cv3(cat(m(cv1 x), cv2 x)); out is the destination Act."""
        W_ = self.Wt
        c_ = W_.mat(f"neck.{name}.cv1.w").t.shape[0]
        cat = self._act(f"{name}.cat", x.B, x.H, x.W, 2 * c_)
        a = self._act(f"{name}.a0", x.B, x.H, x.W, c_)
        b = self._act(f"{name}.a1", x.B, x.H, x.W, c_)
        t = self._act(f"{name}.t", x.B, x.H, x.W, c_)
        self._cba1x1(x, f"{name}.cv1", a, L.ACT_SILU)
        self._cba1x1(x, f"{name}.cv2", cat.cols(c_, 2 * c_), L.ACT_SILU)
        cur, nxt = a, b
        for i in range(n):
            blk = f"{name}.m.conv1" if i == 0 else f"{name}.m.block.{i - 1}"
            dst = cat.cols(0, c_) if i == n - 1 else nxt
            self._conv3x3(cur, f"neck.{blk}.conv1.w", t, bias=W_[f"neck.{blk}.conv1.b"], act=L.ACT_SILU)
            self._conv3x3(t, f"neck.{blk}.conv2.w", dst, bias=W_[f"neck.{blk}.conv2.b"], act=L.ACT_SILU, resid=cur, alpha=W_[f"neck.{blk}.alpha"])
            cur, nxt = dst, cur
        self._cba1x1(cat, f"{name}.cv3", out, L.ACT_SILU)

    def _bifusion(self, name, top, mid, low, out):
        """BiFusion (yolo_world_pafpn.py:692-715): cv3(cat(up(top), cv1(mid), down(cv2(low))))."""
        W_ = self.Wt
        co = W_.mat(f"neck.{name}.cv1.w").t.shape[0]
        cat = self._act(f"{name}.cat", mid.B, mid.H, mid.W, 3 * co)
        up = cat.cols(0, co)
        self.ops += ops.deconv2x2(top.p3_4d, self._mat(f"neck.{name}.upsample.w"), up.p3_4d, W_[f"neck.{name}.upsample.b"])
        self._cba1x1(mid, f"{name}.cv1", cat.cols(co, 2 * co), L.ACT_RELU)
        t = self._act(f"{name}.t", low.B, low.H, low.W, co)
        self._cba1x1(low, f"{name}.cv2", t, L.ACT_RELU)
        self._conv3x3_s2(t, f"neck.{name}.downsample.w", cat.cols(2 * co, 3 * co), bias=W_[f"neck.{name}.downsample.b"], act=L.ACT_RELU)
        self._cba1x1(cat, f"{name}.cv3", out, L.ACT_RELU)

    def _build_neck(self):
        W_, B = self.Wt, self.B
        c1, c2, c3, c4 = self.c_feats
        ch = schema.neck_channels(self.size)
        n = self.cfg["neck_repeats"] // 2
        cat_n4 = self._act("cat_n4", B, c4.H, c4.W, ch[9] + ch[5])
        cat_n3 = self._act("cat_n3", B, c3.H, c3.W, ch[7] + ch[6])
        fpn0 = cat_n4.cols(ch[9], ch[9] + ch[5])
        self._cba1x1(c4, "reduce_layer0", fpn0, L.ACT_RELU)
        bif0 = self._act("bif0", B, c3.H, c3.W, ch[5])
        self._bifusion("Bifusion0", fpn0, c3, c2, bif0)
        f0 = self._act("f0", B, c3.H, c3.W, ch[5])
        self._bepc3("Rep_p4", bif0, f0, n)
        fpn1 = cat_n3.cols(ch[7], ch[7] + ch[6])
        self._cba1x1(f0, "reduce_layer1", fpn1, L.ACT_RELU)
        bif1 = self._act("bif1", B, c2.H, c2.W, ch[6])
        self._bifusion("Bifusion1", fpn1, c2, c1, bif1)
        p3 = self._act("p3", B, c2.H, c2.W, ch[6])
        self._bepc3("Rep_p3", bif1, p3, n)
        self._conv3x3_s2(p3, "neck.downsample2.w", cat_n3.cols(0, ch[7]), bias=W_["neck.downsample2.b"], act=L.ACT_RELU)
        p4 = self._act("p4", B, c3.H, c3.W, ch[8])
        self._bepc3("Rep_n3", cat_n3, p4, n)
        self._conv3x3_s2(p4, "neck.downsample1.w", cat_n4.cols(0, ch[9]), bias=W_["neck.downsample1.b"], act=L.ACT_RELU)
        p5 = self._act("p5", B, c4.H, c4.W, ch[10])
        self._bepc3("Rep_n4", cat_n4, p5, n)
        self.pyramid = [p3, p4, p5]

    # ------------------------------------------------------------------ head
    def _build_head(self):
        W_, B = self.Wt, self.B
        self.embeds, self.logits, self.dists, self.sim_w, self.sim_b = [], [], [], [], []
        for l, p in enumerate(self.pyramid):
            M = B * p.H * p.W
            h1 = self._act(f"head{l}.c1", B, p.H, p.W, schema.HEAD_CLS_CH)
            h2 = self._act(f"head{l}.c2", B, p.H, p.W, schema.HEAD_CLS_CH)
            emb = self._act(f"head{l}.embed", B, p.H, p.W, schema.EMBED_DIM)
            q = f"head.cls_preds.{l}."
            self._conv3x3(p, q + "0.w", h1, bias=W_[q + "0.b"], act=L.ACT_SILU)
            self._conv3x3(h1, q + "1.w", h2, bias=W_[q + "1.b"], act=L.ACT_SILU)
            self._linear(h2, q + "2.w", emb, bias=W_[q + "2.b"])
            # similarity GEMM against the folded (BN * normalised text * exp(logit_scale)) matrix
            # folded similarity matrix: |w| <= max|g| * exp(logit_scale) * max|text element| (<= 1 for normalised class texts)
            tmax = float(W_["prompts"].abs().max()) if self.uni else 1.0
            sw = P3.zeros((self.K_pad, schema.EMBED_DIM), self.dev, self.precise,
                          scale=ops.weight_scale(torch.tensor(W_[f"head.contrast.{l}.wmax"] * max(tmax, 1e-30))))
            sb = torch.zeros(self.K_pad, dtype=F32, device=self.dev)
            lg = self._f32(f"head{l}.logits", M, self.K_pad)
            self.ops.append(ops.linear(emb.p3, sw, lg, bias=sb))
            r1 = self._act(f"head{l}.r1", B, p.H, p.W, schema.HEAD_REG_CH)
            r2 = self._act(f"head{l}.r2", B, p.H, p.W, schema.HEAD_REG_CH)
            q = f"head.reg_preds.{l}."
            self._conv3x3(p, q + "0.w", r1, bias=W_[q + "0.b"], act=L.ACT_SILU)
            self._conv3x3(r1, q + "1.w", r2, bias=W_[q + "1.b"], act=L.ACT_SILU)
            dist = self._f32(f"head{l}.dist", M, 4)
            self._linear(r2, q + "2.w", dist, bias=W_[q + "2.b"], dfl=True)
            self.embeds.append(emb)
            self.logits.append(lg)
            self.dists.append(dist)
            self.sim_w.append(sw)
            self.sim_b.append(sb)

    def _build_post(self, score_thr, nms_pre, iou_thr, max_per_img, nms_mode, tv_numel_thr):
        B = self.B
        self.img_meta = torch.zeros(B, 8, dtype=F32, device=self.dev)
        self.img_meta[:, 2] = 1.0
        self.img_meta[:, 3] = 1.0
        self.img_meta[:, 6] = 1.0
        self.clamp_wh = torch.tensor([[float(self.W), float(self.H)]] * B, dtype=F32, device=self.dev)
        self.post = ops.PostProcess(logits=self.logits, dists=self.dists, level_hw=self.level_hw, strides=list(schema.STRIDES), K=self.K, B=B,
                                    score_thr=score_thr, nms_pre=nms_pre, iou_thr=iou_thr, max_per_img=max_per_img, nms_mode=nms_mode,
                                    tv_numel_thr=tv_numel_thr, img_meta=self.img_meta, clamp_wh=self.clamp_wh)
        self.ops.append(self.post.op)
        self.keep.append(self.post)
        self.max_per_img = max_per_img
        if self.uni:
            self.kept_embed = torch.zeros(B, max_per_img, schema.EMBED_DIM, dtype=F32, device=self.dev)
            kw = {}
            if self.extract:   # eval_retrieval/extract_embedding.py:1181-1190: per-proposal logit_scale / bias of its level
                self.kept_scale = torch.zeros(B, max_per_img, dtype=F32, device=self.dev)
                self.kept_bias = torch.zeros(B, max_per_img, dtype=F32, device=self.dev)
                self.lvl_scale = torch.cat([self.Wt[f"head.contrast.{l}.logit_scale"] for l in range(3)]).contiguous()
                self.lvl_bias = torch.cat([self.Wt[f"head.contrast.{l}.bias"] for l in range(3)]).contiguous()
                kw = dict(lvl_scale=self.lvl_scale, lvl_bias=self.lvl_bias, out_scale=self.kept_scale, out_bias=self.kept_bias)
            self.ops.append(ops.gather_embed([e.p3 for e in self.embeds], self.post.anchors, self.post.counts, self.Wt["head.contrast.g_all"],
                                             self.Wt["head.contrast.h_all"], self.kept_embed, **kw))

    # ------------------------------------------------------------------ run-time API
    def set_text(self, text_feats, normalize=True):
        """Fold class embeddings [K, 768] (fp32, device) into the three per-level similarity matrices."""
        assert text_feats.shape == (self.K, schema.EMBED_DIM) and text_feats.dtype == F32
        self._text = text_feats.contiguous()
        fold = []
        for l in range(3):
            c = f"head.contrast.{l}."
            fold.append(ops.fold_text(self._text, self.Wt[c + "g"], self.Wt[c + "h"], self.Wt[c + "logit_scale"], self.Wt[c + "bias"], self.sim_w[l],
                                      self.sim_b[l], normalize))
        self._fold_program = L.Program(fold)
        self._fold_program.run(torch.cuda.current_stream().cuda_stream)

    def set_meta(self, img_meta, clamp_wh):
        self.img_meta.copy_(img_meta)
        self.clamp_wh.copy_(clamp_wh)

    def run(self, stream=None):
        s = torch.cuda.current_stream().cuda_stream if stream is None else stream
        if self._graph:
            self.program.replay(s)
        else:
            self.program.run(s)

    def capture(self):
        """Capture the whole forward into a CUDA graph (removes ~300 launch latencies per batch)."""
        st = torch.cuda.Stream()
        torch.cuda.synchronize()
        with torch.cuda.stream(st):
            self.program.run(st.cuda_stream)       # warm-up outside capture (attribute setup, lazy init)
            st.synchronize()
            self.program.capture(st.cuda_stream)
        torch.cuda.synchronize()
        self._graph = True

    @property
    def num_launches(self):
        return self.program.num_launches

    def results(self):
        p = self.post
        out = dict(boxes=p.boxes, scores=p.scores, labels=p.labels, anchors=p.anchors, counts=p.counts)
        if self.uni:
            out["embeddings"] = self.kept_embed
        if self.extract:
            out["scales"], out["bias"] = self.kept_scale, self.kept_bias
        return out


class TextPlan:
    """XLM-RoBERTa text tower -> L2-normalised class embeddings [S, 768] (mm_backbone.py:376-390).

    embeddings+LN -> per layer: fused QKV GEMM -> short-sequence attention (one warp per (sequence, head))
    -> O-proj GEMM (+residual in the epilogue) -> LN -> FFN GEMMs (GELU / +residual epilogues) -> LN;
    CLS rows -> Linear -> L2 norm.  The result is cached by the caller per text set.
    """

    def __init__(self, weights, size, S, Lt, device="cuda:0"):
        t = schema.TEXT[schema.SIZES[size]["text"]]
        H, I, nh = t["hidden"], t["inter"], t["heads"]
        if Lt > 128:
            raise ValueError(f"text sequences of {Lt} tokens exceed the 128-token attention kernels")
        self.S, self.Lt, self.dev = S, Lt, torch.device(device)
        W_, dev, T = weights, self.dev, S * Lt
        pr = weights.precise

        def bf(rows, cols):
            return P3.zeros((rows, cols), dev, pr)

        self.ids = torch.zeros(S, Lt, dtype=torch.int32, device=dev)
        self.mask = torch.zeros(S, Lt, dtype=torch.int32, device=dev)
        x = torch.zeros(T, H, dtype=F32, device=dev)
        y = torch.zeros(T, H, dtype=F32, device=dev)
        qkv = torch.zeros(T, 3 * H, dtype=F32, device=dev)
        xb, att, hid, cls = bf(T, H), bf(T, H), bf(T, I), bf(S, H)
        ho = torch.zeros(S, schema.EMBED_DIM, dtype=F32, device=dev)
        self.head_out = ho     # CLS -> Linear output before the L2 norm (what extract_embedding.py's standalone text tower returns)
        self.feats = torch.zeros(S, schema.EMBED_DIM, dtype=F32, device=dev)
        self._keep = [x, y, qkv, xb, att, hid, cls, ho, weights]
        o = []

        def lin(a, wname, out, **kw):
            o.append(ops.linear(a, W_.mat(wname), out, **kw))

        o.append(ops.text_embed(self.ids, W_["emb.word"], W_["emb.pos"], W_["emb.type"], W_["emb.ln_w"], W_["emb.ln_b"], schema.TEXT_EPS,
                                schema.TEXT_PAD, x, xb))
        for i in range(t["layers"]):
            q = f"l{i}."
            lin(xb, q + "qkv.w", qkv, bias=W_[q + "qkv.b"])
            o.append(ops.attn_small(qkv, self.mask, att, nh, 0.125))
            lin(att, q + "o.w", y, bias=W_[q + "o.b"], resid=x, alpha=1.0)
            o.append(ops.ln_rows(y, W_[q + "ln1_w"], W_[q + "ln1_b"], schema.TEXT_EPS, out_bf16=xb, out_f32=x))
            lin(xb, q + "f1.w", hid, bias=W_[q + "f1.b"], act=L.ACT_GELU)
            lin(hid, q + "f2.w", y, bias=W_[q + "f2.b"], resid=x, alpha=1.0)
            o.append(ops.ln_rows(y, W_[q + "ln2_w"], W_[q + "ln2_b"], schema.TEXT_EPS, out_bf16=xb, out_f32=x))
        o.append(ops.gather_rows(x, cls, S, Lt))
        lin(cls, "head.w", ho, bias=W_["head.b"])
        o.append(ops.l2norm_rows(ho, self.feats))
        self.program = L.Program(o)

    def run(self, ids, mask):
        self.ids.copy_(ids)
        self.mask.copy_(mask)
        self.program.run(torch.cuda.current_stream().cuda_stream)
        return self.feats
