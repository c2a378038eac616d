"""Device-side image preparation of the WeDetect-Uni entry points.

    letterbox_params   generate_proposal.py:44-78  ratio, rounded size, paste offset, (dw/2, dh/2)
    Letterbox          generate_proposal.py:17-82 + :1087-1101: PIL BILINEAR resize + centred paste on a 114 canvas for a
                       batch of decoded RGB images, executed by WD_OP_LETTERBOX (csrc/preprocess.cu) straight into the
                       detector's planar uint8 input.  Bit-exact with PIL (tests/test_gpu_letterbox.py).

Host work per image is only what PIL's precompute_coeffs / normalize_coeffs_8bpc do (Pillow src/libImaging/Resample.c):
the window and 22-bit fixed-point weights of every output column and row, computed in double precision in the same
operation order.  Pixels go pageable -> pinned -> HBM untouched.  There is no CPU resize path.
"""
import functools
import math
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _lib as L
from ._lib import WdOp

PRECISION_BITS = 32 - 8 - 2
DESC_WORDS = 16


def letterbox_params(w, h, new_shape):
    """Scale / offsets of generate_proposal.py:17-82 (scale_up=True) without touching pixels."""
    nw, nh = new_shape[1], new_shape[0]
    r = min(nw / w, nh / h)
    unpad = (int(round(w * r)), int(round(h * r)))
    dw, dh = nw - unpad[0], nh - unpad[1]
    return r, unpad, (dw // 2, dh // 2), (dw / 2, dh / 2)


@functools.lru_cache(maxsize=256)
def resample_tables(in_size, out_size):
    """(ksize, bounds int32 [out, 2] = (first, count), weights int32 [out, ksize]) of PIL's 8-bit BILINEAR resampler for
    the full source range; in_size == out_size gives identity tables (the pass PIL skips).  Cached per (in, out): datasets
    repeat a handful of sizes; callers must not modify the returned arrays."""
    if in_size == out_size:
        b = np.stack([np.arange(out_size, dtype=np.int32), np.ones(out_size, dtype=np.int32)], 1)
        return 1, b, np.full((out_size, 1), 1 << PRECISION_BITS, dtype=np.int32)
    scale = float(np.float32(in_size)) / out_size
    filterscale = max(scale, 1.0)
    support = filterscale                       # bilinear: filter support 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    center = 0.0 + (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    inv = 1.0 / filterscale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)           # C (int) cast: truncation toward zero
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    j = np.arange(ksize, dtype=np.int64)
    arg = np.abs(((j[None, :] + xmin[:, None]) - center[:, None] + 0.5) * inv)
    w = np.where(arg < 1.0, 1.0 - arg, 0.0)
    w[j[None, :] >= xmax[:, None]] = 0.0
    ww = np.zeros(out_size, dtype=np.float64)
    for c in range(ksize):                      # the C loop's left-to-right sum (np.sum would reassociate)
        ww = ww + w[:, c]
    k = np.divide(w, ww[:, None], out=np.zeros_like(w), where=ww[:, None] != 0.0)
    kk = (0.5 + k * float(1 << PRECISION_BITS)).astype(np.int32)              # weights are >= 0 for this filter
    return ksize, np.stack([xmin, xmax], 1).astype(np.int32), kk


_POOL = None


def _pool():
    global _POOL
    if _POOL is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        _POOL = ThreadPoolExecutor(max_workers=max(1, min(8, n)))
    return _POOL


def load_rgb(item):
    """One input of SimpleYOLOWorldDetector.forward -> uint8 [h, w, 3] RGB: a file name is opened and converted to RGB
    (generate_proposal.py:1089-1090), a PIL image is taken as is (converted if it is not RGB), an array is passed through."""
    from PIL import Image
    if isinstance(item, (str, bytes)) or hasattr(item, "__fspath__"):
        item = Image.open(item).convert("RGB")
    if isinstance(item, Image.Image):
        item = np.asarray(item if item.mode == "RGB" else item.convert("RGB"))
    item = np.asarray(item)
    if item.dtype != np.uint8 or item.ndim != 3 or item.shape[2] != 3:
        raise TypeError(f"expected a file name, a PIL image or a uint8 [h, w, 3] RGB array, got {item.dtype} {item.shape}")
    return item


def decode_images(items):
    """Decode a batch on the host thread pool (PIL releases the GIL while decoding): at >1000 images/s per GPU a serial
    decode loop would be the bottleneck of the Uni entry points."""
    items = list(items)
    if len(items) <= 1:
        return [load_rgb(it) for it in items]
    return list(_pool().map(load_rgb, items))


def pack_batch(images, H, W, with_src=True):
    """Host side of WD_OP_LETTERBOX for one batch: geometry + PIL tables per image, packed the way the kernels read them.
    Returns dict(desc int32 [n,16], coef int32 [...], src_parts [(byte offset, flat uint8 view)], src_bytes, tmp_bytes,
    ratios, offsets, shapes) and, with_src, `src`: the concatenated source bytes (tests; the device path copies the parts
    straight into pinned memory instead)."""
    desc = np.zeros((len(images), DESC_WORDS), dtype=np.int32)
    coef_parts, src_parts = [], []
    src_bytes = coef_words = tmp_bytes = 0
    ratios, offsets, shapes = [], [], []
    for b, im in enumerate(images):
        im = np.ascontiguousarray(im)
        if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
            raise TypeError(f"expected uint8 [h, w, 3] RGB, got {im.dtype} {im.shape}")
        h, w = im.shape[:2]
        r, (nw, nh), (left, top), off = letterbox_params(w, h, (H, W))
        if nw < 1 or nh < 1:
            raise ValueError(f"image of {w}x{h} letterboxes to an empty {nw}x{nh} area")
        ksh, bh, kh = resample_tables(w, nw)
        ksv, bv, kv = resample_tables(h, nh)
        # Pillow >= 12 (Image.resize): very tall images that shrink vertically are resampled vertically first
        vfirst = int(h > w * 100 and nh < h)
        if vfirst:
            first, rows, tmp_cols = 0, nh, w
        else:
            first = int(bv[0, 0])
            rows, tmp_cols = int(bv[-1, 0] + bv[-1, 1]) - first, nw
            bv = bv.copy()
            bv[:, 0] -= first
        tables = np.concatenate([bh.reshape(-1), kh.reshape(-1), bv.reshape(-1), kv.reshape(-1)])
        # byte offsets are (lo, hi) int32 pairs in the ABI; one batch of decoded images stays below 2 GiB (checked below)
        desc[b] = (src_bytes, 0, w, h, nw, nh, left, top, first, rows, tmp_bytes, 0, coef_words, ksh, ksv, vfirst)
        src_parts.append((src_bytes, im.reshape(-1)))
        coef_parts.append(tables)
        src_bytes += (im.size + 15) // 16 * 16
        coef_words += tables.size
        tmp_bytes += (rows * tmp_cols * 3 + 15) // 16 * 16
        if src_bytes >= 2 ** 31 or tmp_bytes >= 2 ** 31:
            raise ValueError("batch of source images exceeds 2 GiB")
        ratios.append(r); offsets.append(off); shapes.append((h, w))
    out = dict(desc=desc, coef=np.concatenate(coef_parts), src_parts=src_parts, src_bytes=src_bytes, tmp_bytes=tmp_bytes, ratios=ratios,
               offsets=offsets, shapes=shapes)
    if with_src:
        src = np.zeros(src_bytes, dtype=np.uint8)
        for off, flat in src_parts:
            src[off: off + flat.size] = flat
        out["src"] = src
    return out


class Letterbox:
    """Letterboxes up to B decoded RGB images into `out` (uint8 [B, 3, H, W], device) on the current stream."""

    def __init__(self, out, pad=114):
        L.load(require_gpu=True)
        assert out.dtype == torch.uint8 and out.dim() == 4 and out.shape[1] == 3 and out.is_contiguous() and out.is_cuda
        self.out, self.pad = out, int(pad)
        self.B, _, self.H, self.W = out.shape
        self.dev = out.device
        self._cap = dict(src=0, coef=0, tmp=0)
        self._host, self._devb = {}, {}
        self._desc_host = torch.zeros(self.B, DESC_WORDS, dtype=torch.int32).pin_memory()
        self._desc_dev = torch.zeros(self.B, DESC_WORDS, dtype=torch.int32, device=self.dev)
        self._program = None
        self._copied = None

    def _ensure(self, name, need, dtype, host=True):
        if need <= self._cap[name]:
            return False
        cap = max(need, int(self._cap[name] * 1.5), 1 << 16)
        if host:
            self._host[name] = torch.empty(cap, dtype=dtype).pin_memory()
        self._devb[name] = torch.empty(cap, dtype=dtype, device=self.dev)
        self._cap[name] = cap
        return True

    def run(self, images):
        """images: list (1..B) of uint8 arrays [h, w, 3] RGB; unused batch slots become plain padding.
        Returns (ratios, offsets (dw/2, dh/2), ori_shapes (h, w))."""
        if not 1 <= len(images) <= self.B:
            raise ValueError(f"{len(images)} images for a batch of {self.B}")
        if self._copied is not None:
            self._copied.synchronize()           # the previous batch's H2D must have left the pinned buffers
        pk = pack_batch(images, self.H, self.W, with_src=False)
        n_src, n_coef = pk["src_bytes"], pk["coef"].size
        grew = self._ensure("src", n_src, torch.uint8)
        grew |= self._ensure("coef", n_coef, torch.int32)
        grew |= self._ensure("tmp", pk["tmp_bytes"], torch.uint8, host=False)
        if grew or self._program is None:
            op = WdOp()
            op.kind = L.OP_LETTERBOX
            op.i[0], op.i[1], op.i[2], op.i[3] = self.B, self.H, self.W, self.pad
            for k, t in enumerate((self._devb["src"], self._desc_dev, self._devb["coef"], self._devb["tmp"], self.out)):
                op.p[k] = t.data_ptr()
            self._program = L.Program([op])
        desc = self._desc_host.numpy()
        desc[:] = 0
        desc[: len(images)] = pk["desc"]
        src_np = self._host["src"].numpy()
        # pageable -> pinned: one memcpy per image, spread over a few threads (numpy releases the GIL for the copy)
        list(_pool().map(lambda part: np.copyto(src_np[part[0]: part[0] + part[1].size], part[1]), pk["src_parts"]))
        self._host["coef"].numpy()[:n_coef] = pk["coef"]
        self._devb["src"][:n_src].copy_(self._host["src"][:n_src], non_blocking=True)
        self._devb["coef"][:n_coef].copy_(self._host["coef"][:n_coef], non_blocking=True)
        self._desc_dev.copy_(self._desc_host, non_blocking=True)
        self._copied = torch.cuda.Event()
        self._copied.record()
        self._program.run(torch.cuda.current_stream().cuda_stream)
        self.h2d_bytes = n_src + 4 * n_coef + 4 * self.B * DESC_WORDS
        return pk["ratios"], pk["offsets"], pk["shapes"]
