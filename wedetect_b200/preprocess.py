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


class EncodedImage:
    """A JPEG still in its compressed form, to be decoded on the device by nvJPEG (wd_jpeg_*) straight into the source buffer of the
    resize kernels; (h, w) come from the header."""

    def __init__(self, blob, h, w):
        self.blob, self.h, self.w = blob, int(h), int(w)
        self.shape, self.size = (self.h, self.w, 3), self.h * self.w * 3


_JPEG_SOI = b"\xff\xd8\xff"


def exif_orientation(blob):
    """EXIF orientation tag (0x0112) of a JPEG byte string, 1 when absent: cv2.imdecode(IMREAD_COLOR) rotates by it, nvJPEG and PIL's
    plain open() do not."""
    i, n = 2, len(blob)
    while i + 4 <= n and blob[i] == 0xFF:
        marker, seglen = blob[i + 1], int.from_bytes(blob[i + 2: i + 4], "big")
        if marker == 0xDA or seglen < 2:          # start of scan: no more headers
            break
        if marker == 0xE1 and blob[i + 4: i + 10] == b"Exif\x00\x00":
            t = i + 10
            order = "little" if blob[t: t + 2] == b"II" else "big"
            ifd = t + int.from_bytes(blob[t + 4: t + 8], order)
            if ifd + 2 > n:
                return 1
            for k in range(int.from_bytes(blob[ifd: ifd + 2], order)):
                e = ifd + 2 + 12 * k
                if e + 12 > n:
                    break
                if int.from_bytes(blob[e: e + 2], order) == 0x0112:
                    return int.from_bytes(blob[e + 8: e + 10], order)
            return 1
        i += 2 + seglen
    return 1


def read_encoded(item, decoder, honour_exif=False):
    """File name / bytes of a 3-component (or grey) JPEG -> EncodedImage; anything else -> None (decoded on the host instead).
    honour_exif: files whose EXIF orientation is not 1 also go to the host decoder (cv2 rotates them, nvJPEG would not)."""
    blob = item if isinstance(item, (bytes, bytearray, memoryview)) else None
    if blob is None and (isinstance(item, str) or hasattr(item, "__fspath__")):
        with open(os.fsdecode(item), "rb") as f:
            blob = f.read()
    if blob is None:
        return None
    blob = bytes(blob)
    if blob[:3] != _JPEG_SOI or (honour_exif and exif_orientation(blob) != 1):
        return None
    try:
        w, h, nc, _ = decoder.info(blob)
    except L.WdError:
        return None
    if w < 1 or h < 1 or nc not in (1, 3):
        return None
    return EncodedImage(blob, h, w)


def load_bgr(item):
    """One input of the mmdet test pipeline -> uint8 [h, w, 3] BGR: a file name is decoded as LoadImageFromFile does (mmcv.imfrombytes,
    cv2 backend, flag 'color' = cv2.IMREAD_COLOR: infer_wedetect.py:111, config/wedetect_base.py:112); an array is taken as decoded."""
    if isinstance(item, (str, bytes)) or hasattr(item, "__fspath__"):
        try:
            import cv2
        except ImportError as e:          # decoding is the one host step left; without cv2 pass decoded arrays
            raise RuntimeError("decoding image files for the mmdet pipeline needs cv2 (pass decoded uint8 BGR arrays instead)") from e
        name = os.fsdecode(item)
        arr = cv2.imread(name, cv2.IMREAD_COLOR)
        if arr is None:
            raise FileNotFoundError(name)
        return arr
    item = np.asarray(item)
    if item.dtype != np.uint8 or item.ndim != 3 or item.shape[2] != 3:
        raise TypeError(f"expected a file name or a uint8 [h, w, 3] BGR array, got {item.dtype} {item.shape}")
    return item


def decode_images_bgr(items):
    items = list(items)
    if len(items) <= 1:
        return [load_bgr(it) for it in items]
    return list(_pool().map(load_bgr, items))


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
        if not isinstance(im, EncodedImage):
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
        src_parts.append((src_bytes, im if isinstance(im, EncodedImage) else im.reshape(-1)))
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
            if not isinstance(flat, EncodedImage):
                src[off: off + flat.size] = flat
        out["src"] = src
    return out


class Letterbox:
    """Letterboxes up to B decoded RGB images into `out` (uint8 [B, 3, H, W], device) on the current stream."""
    _KIND = L.OP_LETTERBOX
    _BGR = False          # channel order nvJPEG writes for this entry point (the Uni path works on RGB, the mmdet path on BGR)

    def _pack(self, images):
        return pack_batch(images, self.H, self.W, with_src=False)

    def decoder(self):
        """nvJPEG state of the calling thread on this device (created on first use): the Huffman stage of nvjpegDecode runs on the
        host, so a batch is decoded from the thread pool, one decoder state per worker."""
        import threading
        if getattr(self, "_tls", None) is None:
            self._tls = threading.local()
        if getattr(self._tls, "jpeg", None) is None:
            with torch.cuda.device(self.dev):
                self._tls.jpeg = L.JpegDecoder()
        return self._tls.jpeg

    def encoded(self, items, host_decode):
        """items -> list for run(): JPEG files / byte strings stay compressed (EncodedImage, decoded on the device), everything
        else goes through `host_decode`."""
        dec = self.decoder()
        out = [it if isinstance(it, np.ndarray) else read_encoded(it, dec, honour_exif=self._BGR) for it in items]
        rest = [i for i, o in enumerate(out) if o is None]
        if rest:
            for i, arr in zip(rest, host_decode([items[i] for i in rest])):
                out[i] = arr
        return out

    def _result(self, pk):
        return pk["ratios"], pk["offsets"], pk["shapes"]

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
        if need <= self._cap[name] and name in self._devb:
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
        pk = self._pack(images)
        n_src, n_coef = pk["src_bytes"], pk["coef"].size
        grew = self._ensure("src", n_src, torch.uint8)
        grew |= self._ensure("coef", n_coef, torch.int32)
        grew |= self._ensure("tmp", pk["tmp_bytes"], torch.uint8, host=False)
        if grew or self._program is None:
            op = WdOp()
            op.kind = self._KIND
            op.i[0], op.i[1], op.i[2], op.i[3] = self.B, self.H, self.W, self.pad
            for k, t in enumerate((self._devb["src"], self._desc_dev, self._devb["coef"], self._devb["tmp"], self.out)):
                op.p[k] = t.data_ptr()
            self._program = L.Program([op])
        desc = self._desc_host.numpy()
        desc[:] = 0
        desc[: len(images)] = pk["desc"]
        src_np = self._host["src"].numpy()
        # pageable -> pinned: one memcpy per image, spread over a few threads (numpy releases the GIL for the copy)
        host_parts = [p_ for p_ in pk["src_parts"] if not isinstance(p_[1], EncodedImage)]
        list(_pool().map(lambda part: np.copyto(src_np[part[0]: part[0] + part[1].size], part[1]), host_parts))
        self._host["coef"].numpy()[:n_coef] = pk["coef"].view(np.int32)
        if host_parts:           # (a batch of compressed images only sends its JPEG bytes: nothing to copy here)
            self._devb["src"][:n_src].copy_(self._host["src"][:n_src], non_blocking=True)
        self._devb["coef"][:n_coef].copy_(self._host["coef"][:n_coef], non_blocking=True)
        self._desc_dev.copy_(self._desc_host, non_blocking=True)
        self._copied = torch.cuda.Event()
        self._copied.record()
        stream = torch.cuda.current_stream().cuda_stream
        # compressed images: nvJPEG writes the pixels where the H2D copy would have put them (all on this stream, ahead of the kernel)
        enc = [(off, part) for off, part in pk["src_parts"] if isinstance(part, EncodedImage)]
        base = self._devb["src"].data_ptr()

        def dec_one(job):
            with torch.cuda.device(self.dev):
                self.decoder().decode(job[1].blob, base + job[0], 3 * job[1].w, bgr=self._BGR, stream=stream)
        if len(enc) > 1:
            list(_pool().map(dec_one, enc))
        elif enc:
            dec_one(enc[0])
        enc_bytes = sum(len(part.blob) for _, part in enc)
        self._program.run(stream)
        self.h2d_bytes = (n_src if host_parts else 0) + enc_bytes + 4 * n_coef + 4 * self.B * DESC_WORDS
        return self._result(pk)


# ------------------------------------------------------------------------------------------------
# The mmcv test pipeline of infer_wedetect.py / test.py (config/wedetect_base.py:111-133) on the device:
#   WeDetectKeepRatioResize   transforms.py:62-123   ratio = min(long / max(h, w), short / min(h, w)); mmcv.imresize to
#                             (int(w ratio), int(h ratio)) with 'area' when ratio < 1, else 'bilinear'; scale_factor = new / old
#   WeDetectLetterResize      transforms.py:180-272, 318-328   ratio (<= 1 unless allow_scale_up), no_pad_shape = round(shape ratio),
#                             optional second mmcv.imresize ('bilinear'), pad to `scale` with top / left = int(round(pad // 2 - 0.1)),
#                             scale_factor multiplied with the first transform's, pad_param float32 [top, bottom, left, right]
# mmcv.imresize(backend='cv2') is cv2.resize and mmcv.impad is cv2.copyMakeBorder (mmcv 2.1.0, un-vendored): the pixel
# arithmetic below is OpenCV 4.x's (modules/imgproc/src/resize.cpp), executed by WD_OP_CV_RESIZE_PAD.  Host work per image is
# the geometry and OpenCV's coefficient tables (double precision, same operation order, cached per (source, target) size).
# ------------------------------------------------------------------------------------------------
_DBL_EPS = 2.220446049250313e-16
CV_COPY, CV_AREA, CV_AREA_INT, CV_LINEAR = 0, 1, 2, 3


def keep_ratio_resize_geometry(h, w, scale):
    """WeDetectKeepRatioResize._resize_img (transforms.py:62-123) without pixels: ((new_h, new_w), interpolation or None,
    scale_factor (w, h))."""
    if isinstance(scale, (int, float)):
        if scale <= 0:
            raise ValueError(f"Invalid scale {scale}, must be positive.")
        ratio = scale
    else:
        ratio = min(max(scale) / max(h, w), min(scale) / min(h, w))
    nh, nw, interp = h, w, None
    if ratio != 1:
        nw, nh = int(w * ratio), int(h * ratio)
        interp = "area" if ratio < 1 else "bilinear"
    return (nh, nw), interp, (nw / w, nh / h)


def letter_resize_geometry(h, w, scale_hw, allow_scale_up=True, use_mini_pad=False, stretch_only=False, half_pad_param=False):
    """WeDetectLetterResize._resize_img (transforms.py:180-272) without pixels: ((no_pad_h, no_pad_w), resize needed,
    scale_factor (w, h), (top, bottom, left, right), pad_param float32[4])."""
    ratio = min(scale_hw[0] / h, scale_hw[1] / w)
    if not allow_scale_up:
        ratio = min(ratio, 1.0)
    no_pad = (int(round(h * ratio)), int(round(w * ratio)))
    padding_h, padding_w = scale_hw[0] - no_pad[0], scale_hw[1] - no_pad[1]
    if use_mini_pad:
        padding_w, padding_h = int(np.mod(padding_w, 32)), int(np.mod(padding_h, 32))
    elif stretch_only:
        padding_h, padding_w = 0.0, 0.0
        no_pad = (scale_hw[0], scale_hw[1])
    top, left = int(round(padding_h // 2 - 0.1)), int(round(padding_w // 2 - 0.1))
    pads = (top, padding_h - top, left, padding_w - left)
    if half_pad_param:
        pad_param = np.array([padding_h / 2, padding_h / 2, padding_w / 2, padding_w / 2], dtype=np.float32)
    else:
        pad_param = np.array(pads, dtype=np.float32)
    return no_pad, (h, w) != no_pad, (no_pad[1] / w, no_pad[0] / h), pads, pad_param


def _cv_scale(ssize, dsize):
    return 1.0 / (dsize / ssize)          # cv::resize: inv_scale = (double)dsize / ssize; scale = 1. / inv_scale


@functools.lru_cache(maxsize=512)
def cv_area_table(ssize, dsize):
    """OpenCV computeResizeAreaTab for one axis (cn = 1): (idx int32 [dsize + 1], si int32 [n], alpha float32 [n]); the entries of
    output d are idx[d] .. idx[d + 1] - 1, in OpenCV's order."""
    scale = _cv_scale(ssize, dsize)
    d = np.arange(dsize, dtype=np.float64)
    f1 = d * scale
    f2 = f1 + scale
    cell = np.minimum(scale, ssize - f1)
    s1 = np.ceil(f1).astype(np.int64)
    s2 = np.minimum(np.floor(f2).astype(np.int64), ssize - 1)
    s1 = np.minimum(s1, s2)
    hl = (s1 - f1) > 1e-3
    nm = s2 - s1
    hr = (f2 - s2) > 1e-3
    cnt = hl.astype(np.int64) + nm + hr.astype(np.int64)
    idx = np.concatenate([[0], np.cumsum(cnt)])
    n = int(idx[-1])
    si = np.empty(n, dtype=np.int64)
    al = np.empty(n, dtype=np.float32)
    pl = idx[:-1][hl]
    si[pl] = s1[hl] - 1
    al[pl] = ((s1 - f1) / cell).astype(np.float32)[hl]
    rep = np.repeat(np.arange(dsize), nm)
    within = np.arange(int(nm.sum())) - np.repeat(np.cumsum(nm) - nm, nm)
    pm = idx[:-1][rep] + hl[rep] + within
    si[pm] = s1[rep] + within
    al[pm] = (1.0 / cell).astype(np.float32)[rep]
    pr = idx[1:][hr] - 1
    si[pr] = s2[hr]
    al[pr] = (np.minimum(np.minimum(f2 - s2, 1.0), cell) / cell).astype(np.float32)[hr]
    return idx.astype(np.int32), si.astype(np.int32), al


@functools.lru_cache(maxsize=512)
def cv_linear_table(ssize, dsize, clamp):
    """OpenCV's INTER_LINEAR tables for one axis of an 8-bit image (resize.cpp, cv::resize -> resizeGeneric_): (ofs int32 [dsize],
    packed int32 [dsize] = alpha0 | alpha1 << 16 in 11-bit fixed point, xmax).  clamp=True is the x axis (taps clamped into the row,
    columns >= xmax replicate the last source pixel); the y axis keeps its fractions and clips the ROWS when they are read."""
    scale = _cv_scale(ssize, dsize)
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    xmax = dsize
    if clamp:
        neg = s < 0
        f[neg], s[neg] = 0.0, 0
        hi = s + 1 >= ssize
        if hi.any():
            xmax = int(np.argmax(hi))
        last = s >= ssize - 1
        f[last], s[last] = 0.0, ssize - 1
    a0 = np.clip(np.rint((np.float32(1.0) - f) * np.float32(2048.0)), -32768, 32767).astype(np.int64)
    a1 = np.clip(np.rint(f * np.float32(2048.0)), -32768, 32767).astype(np.int64)
    packed = ((a0 & 0xffff) | ((a1 & 0xffff) << 16)).astype(np.uint32).view(np.int32)
    return s.astype(np.int32), packed, xmax


def cv_resize_plan(h, w, nh, nw, interp):
    """(mode, tables int32, kx, ky, float bits of 1 / (kx ky), xmax) of cv2.resize((h, w) -> (nh, nw), interp): OpenCV picks the
    integer-box INTER_AREA when both scales are whole numbers (to DBL_EPSILON), the table INTER_AREA when both are >= 1, and treats
    INTER_AREA as INTER_LINEAR otherwise (its bilinear weights then differ from INTER_LINEAR's and are not built here)."""
    if (h, w) == (nh, nw):
        return CV_COPY, np.zeros(0, np.int32), 0, 0, 0, 0
    if interp == "area":
        sx, sy = _cv_scale(w, nw), _cv_scale(h, nh)
        if sx < 1 or sy < 1:
            raise NotImplementedError("INTER_AREA with an up-scaled axis")
        kx, ky = int(np.rint(sx)), int(np.rint(sy))
        if abs(sx - kx) < _DBL_EPS and abs(sy - ky) < _DBL_EPS:
            return CV_AREA_INT, np.zeros(0, np.int32), kx, ky, int(np.array(np.float32(1.0) / np.float32(kx * ky), np.float32).view(np.int32)), 0
        xi, xs, xa = cv_area_table(w, nw)
        yi, ys, ya = cv_area_table(h, nh)
        return CV_AREA, np.concatenate([xi, yi, xs, xa.view(np.int32), ys, ya.view(np.int32)]), 0, 0, 0, 0
    if interp == "bilinear":
        xo, xp, xmax = cv_linear_table(w, nw, True)
        yo, yp, _ = cv_linear_table(h, nh, False)
        return CV_LINEAR, np.concatenate([xo, xp, yo, yp]), 0, 0, 0, xmax
    raise NotImplementedError(f"interpolation {interp!r}")


def mm_test_geometry(h, w, scale=(640, 640), allow_scale_up=False, keep_ratio_first=True, use_mini_pad=False, stretch_only=False,
                     half_pad_param=False):
    """Both transforms for one image of (h, w): dict(resize=(nh, nw), interp, pads=(top, bottom, left, right), scale_factor (w, h),
    pad_param float32[4], img_shape (H, W, 3), ori_shape (h, w)).  `scale` is (w, h) as in the config."""
    (h1, w1), interp, sf = (keep_ratio_resize_geometry(h, w, tuple(scale)) if keep_ratio_first else ((h, w), None, None))
    no_pad, again, sf2, pads, pad_param = letter_resize_geometry(h1, w1, tuple(scale)[::-1], allow_scale_up, use_mini_pad, stretch_only, half_pad_param)
    if again:
        if interp is not None:
            raise NotImplementedError("two chained resizes (WeDetectLetterResize resizing the output of WeDetectKeepRatioResize)")
        interp = "bilinear"                                # MMDET_Resize's default interpolation
    scale_factor = sf2 if sf is None else (sf2[0] * sf[0], sf2[1] * sf[1])        # transforms.py:318-325
    H, W = no_pad[0] + pads[0] + pads[1], no_pad[1] + pads[2] + pads[3]
    return dict(resize=no_pad, interp=interp, pads=pads, scale_factor=scale_factor, pad_param=pad_param, img_shape=(int(H), int(W), 3), ori_shape=(h, w))


def pack_mm_batch(images, H, W, **pipe):
    """Host side of WD_OP_CV_RESIZE_PAD for one batch (same contract as pack_batch)."""
    desc = np.zeros((len(images), DESC_WORDS), dtype=np.int32)
    coef_parts, src_parts, metas = [], [], []
    src_bytes = coef_words = 0
    for b, im in enumerate(images):
        if not isinstance(im, EncodedImage):
            im = np.ascontiguousarray(im)
            if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
                raise TypeError(f"expected uint8 [h, w, 3] BGR, got {im.dtype} {im.shape}")
        h, w = im.shape[:2]
        g = mm_test_geometry(h, w, **pipe)
        if g["img_shape"][:2] != (H, W):
            raise ValueError(f"image of {w}x{h} pads to {g['img_shape'][:2]}, the batch canvas is {(H, W)}")
        nh, nw = g["resize"]
        if nw < 1 or nh < 1:
            raise ValueError(f"image of {w}x{h} resizes to an empty {nw}x{nh} area")
        mode, tab, kx, ky, sbits, xmax = cv_resize_plan(h, w, nh, nw, g["interp"])
        desc[b] = (src_bytes, 0, w, h, nw, nh, g["pads"][2], g["pads"][0], mode, coef_words, kx, ky, sbits, xmax, 0, 0)
        src_parts.append((src_bytes, im if isinstance(im, EncodedImage) else im.reshape(-1)))
        coef_parts.append(tab)
        src_bytes += (im.size + 15) // 16 * 16
        coef_words += tab.size
        if src_bytes >= 2 ** 31:
            raise ValueError("batch of source images exceeds 2 GiB")
        metas.append(g)
    coef = np.concatenate(coef_parts) if coef_words else np.zeros(1, np.int32)
    return dict(desc=desc, coef=coef, src_parts=src_parts, src_bytes=src_bytes, tmp_bytes=0, metas=metas)


class MMTestPipeline(Letterbox):
    """The resize / pad half of the reference's test pipeline for up to B decoded uint8 BGR images, on the device, into `out`
    (uint8 [B, 3, H, W], channel order passed through).  run(images) returns one metainfo dict per image with the keys
    PackDetInputs forwards (ori_shape, img_shape, scale_factor, pad_param): config/wedetect_base.py:111-133."""
    _KIND = L.OP_CV_RESIZE_PAD
    _BGR = True

    def __init__(self, out, scale=(640, 640), pad=114, allow_scale_up=False, **pipe):
        super().__init__(out, pad=pad)
        self.pipe = dict(scale=tuple(scale), allow_scale_up=allow_scale_up, **pipe)

    def _pack(self, images):
        return pack_mm_batch(images, self.H, self.W, **self.pipe)

    def _result(self, pk):
        return [dict(ori_shape=g["ori_shape"], img_shape=g["img_shape"], scale_factor=g["scale_factor"], pad_param=g["pad_param"]) for g in pk["metas"]]
