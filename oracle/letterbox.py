"""CPU ORACLE of the Uni-path image preparation.  TEST INFRASTRUCTURE ONLY.

Restates, as integer arithmetic on numpy arrays:
  letterbox_ref            generate_proposal.py:17-82  (ratio, rounded size, PIL BILINEAR resize, centred paste on a
                           114-grey canvas) as called from SimpleYOLOWorldDetector.forward (:1087-1101)
  resize_bilinear_ref      PIL.Image.resize(size, Image.Resampling.BILINEAR) for 8-bit RGB.  The arithmetic lives in the
                           third-party dependency Pillow (the reference's requirements leave it unpinned; installed here:
                           see PIL.__version__), file src/libImaging/Resample.c: precompute_coeffs (triangle filter whose
                           support grows with the down-scale factor), normalize_coeffs_8bpc (22-bit fixed point),
                           ImagingResampleHorizontal_8bpc then ImagingResampleVertical_8bpc with an 8-bit intermediate
                           image, clip8 after each pass.
Pinned bit-exactly by tests/test_letterbox_cpu.py against PIL itself (installed in this image, so the pin also runs on the
GPU box), against the reference's own letterbox() when /root/reference is mounted, and against committed digests
(tests/golden/letterbox_digests.json, generator tests/golden/make_golden_letterbox.py).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def coeffs_ref(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc for the full box (0, in_size) and the bilinear filter.
    Returns (ksize, bounds [out,2] = (xmin, count), kk int32 [out, ksize])."""
    scale = float(np.float32(in_size - 0.0)) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ww = 0.0
        k = [0.0] * ksize
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            w = 1.0 - a if a < 1.0 else 0.0
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):
            v = k[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _pass(img, bounds, kk, axis):
    """One resampling pass along `axis` (1 = horizontal, 0 = vertical) of an HxWx3 uint8 image."""
    src = img.astype(np.int64)
    n = bounds.shape[0]
    shape = list(img.shape)
    shape[axis] = n
    out = np.zeros(shape, dtype=np.uint8)
    for o in range(n):
        lo, cnt = int(bounds[o, 0]), int(bounds[o, 1])
        acc = np.full(out.take(0, axis=axis).shape, 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for j in range(cnt):
            acc = acc + src.take(lo + j, axis=axis) * int(kk[o, j])
        if axis == 1:
            out[:, o] = _clip8(acc)
        else:
            out[o] = _clip8(acc)
    return out


def resize_bilinear_ref(img, out_w, out_h):
    """img uint8 [H, W, 3] -> uint8 [out_h, out_w, 3], PIL's two-pass algorithm (horizontal first, 8-bit intermediate,
    only the source rows the vertical pass needs; a pass whose size does not change is skipped; both skipped = copy)."""
    H, W = img.shape[:2]
    need_h, need_v = out_w != W, out_h != H
    if not need_h and not need_v:
        return img.copy()
    if H > W * 100 and out_h < H:
        # PIL/Image.py (Pillow >= 12) Image.resize: very tall images are first resized vertically to (W, out_h), then
        # horizontally - two separate ImagingResample calls, i.e. the pass order is swapped
        _, bv, kv = coeffs_ref(H, out_h)
        cur = _pass(img, bv, kv, axis=0)
        if need_h:
            _, bh, kh = coeffs_ref(W, out_w)
            cur = _pass(cur, bh, kh, axis=1)
        return cur
    cur = img
    if need_v:
        _, bv, kv = coeffs_ref(H, out_h)
        first, last = int(bv[0, 0]), int(bv[-1, 0] + bv[-1, 1])
    if need_h:
        _, bh, kh = coeffs_ref(W, out_w)
        if need_v:
            cur = cur[first:last]
            bv = bv.copy()
            bv[:, 0] -= first
        cur = _pass(cur, bh, kh, axis=1)
    if need_v:
        cur = _pass(cur, bv, kv, axis=0)
    return cur


def letterbox_params_ref(w, h, new_shape=(640, 640)):
    """generate_proposal.py:44-78: ratio, resized size, paste offsets and the (dw/2, dh/2) floats used to map boxes back."""
    nw, nh = new_shape[1], new_shape[0]
    r = min(nw / w, nh / h)
    new_unpad = (int(round(w * r)), int(round(h * r)))
    dw, dh = nw - new_unpad[0], nh - new_unpad[1]
    return r, new_unpad, (dw // 2, dh // 2), (dw / 2, dh / 2)


def letterbox_ref(img, new_shape=(640, 640), color=(114, 114, 114)):
    """img uint8 [H, W, 3] RGB -> (canvas uint8 [new_h, new_w, 3], ratio, (dw/2, dh/2))."""
    h, w = img.shape[:2]
    r, new_unpad, (left, top), off = letterbox_params_ref(w, h, new_shape)
    res = resize_bilinear_ref(img, new_unpad[0], new_unpad[1])
    canvas = np.empty((new_shape[0], new_shape[1], 3), dtype=np.uint8)
    canvas[:] = np.array(color, dtype=np.uint8)
    # PIL paste clips to the canvas (a resized side can exceed it by rounding only when it equals the canvas side, so no clip here)
    canvas[top:top + new_unpad[1], left:left + new_unpad[0]] = res
    return canvas, r, off
