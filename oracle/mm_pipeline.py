"""CPU restatement of the reference's test-time image transforms (TEST INFRASTRUCTURE: only tests/ and the smoke / bench
checkers may import this; the product never does).

    keep_ratio_resize   WeDetectKeepRatioResize._resize_img        wedetect/datasets/transformers/transforms.py:62-123
    letter_resize       WeDetectLetterResize._resize_img + transform transforms.py:180-272, 318-328
    test_pipeline       the two composed as config/wedetect_base.py:111-118 does
    cv_resize           cv2.resize for uint8 3-channel images, INTER_AREA (down) and INTER_LINEAR, restated from OpenCV 4.x
                        modules/imgproc/src/resize.cpp (computeResizeAreaTab, ResizeArea_, ResizeAreaFast_, resizeGeneric_ with
                        HResizeLinear / VResizeLinear<uchar, int, short>): mmcv.imresize(backend='cv2') is cv2.resize, and neither mmcv
                        (2.1.0) nor OpenCV is vendored in the reference.

Pinned two ways (tests/test_mm_pipeline_cpu.py): cv_resize against the installed cv2 on random sizes, and test_pipeline against
tests/golden/mm_pipeline.npz, produced by the reference's unmodified transform classes (tests/golden/make_golden_mm_pipeline.py).
Written with plain loops over table entries, independently of the product's vectorised table builders.
"""
import math

import numpy as np

DBL_EPSILON = 2.220446049250313e-16
f32 = np.float32


def _scale(ssize, dsize):
    return 1.0 / (dsize / ssize)


def area_tab(ssize, dsize):
    """computeResizeAreaTab: list of (di, si, alpha float32)."""
    scale = _scale(ssize, dsize)
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, f32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, f32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, f32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def _sat8(v):
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)        # saturate_cast<uchar>(float): cvRound (half to even) + clamp


def resize_area(img, dw, dh):
    """ResizeArea_<uchar, float>: per source row buf[dx] += S * alpha in table order; rows: sum = beta * buf, then sum += beta * buf."""
    h, w, _ = img.shape
    S = img.astype(np.float32)
    xt, yt = area_tab(w, dw), area_tab(h, dh)
    rows = {}
    out = np.zeros((dh, dw, 3), np.uint8)
    sums = {}
    for dy, sy, beta in yt:
        if sy not in rows:
            buf = np.zeros((dw, 3), np.float32)
            for dx, sx, alpha in xt:
                buf[dx] = buf[dx] + S[sy, sx] * alpha
            rows[sy] = buf
        sums[dy] = beta * rows[sy] if dy not in sums else sums[dy] + beta * rows[sy]
    for dy in range(dh):
        out[dy] = _sat8(sums[dy])
    return out


def resize_area_int(img, kx, ky):
    """ResizeAreaFast_: integer box sums; 2x2 boxes go through the SIMD (sum + 2) >> 2, everything else saturate_cast(sum * (1.f / area))."""
    h, w, _ = img.shape
    dh, dw = h // ky, w // kx
    s = img[: dh * ky, : dw * kx].astype(np.int64).reshape(dh, ky, dw, kx, 3).sum(axis=(1, 3))
    if (kx, ky) == (2, 2):
        return ((s + 2) >> 2).astype(np.uint8)
    return _sat8(s.astype(np.float32) * (f32(1.0) / f32(kx * ky)))


def linear_tab(ssize, dsize, x_axis):
    scale = _scale(ssize, dsize)
    ofs, coef, xmax = [], [], dsize
    for d in range(dsize):
        fx = f32((d + 0.5) * scale - 0.5)
        sx = int(math.floor(fx))
        fx = f32(fx - f32(sx))
        if x_axis:
            if sx < 0:
                fx, sx = f32(0), 0
            if sx + 1 >= ssize:
                xmax = min(xmax, d)
                if sx >= ssize - 1:
                    fx, sx = f32(0), ssize - 1
        ofs.append(sx)
        coef.append((int(np.clip(np.rint(f32((f32(1) - fx) * f32(2048))), -32768, 32767)), int(np.clip(np.rint(f32(fx * f32(2048))), -32768, 32767))))
    return ofs, coef, xmax


def resize_linear(img, dw, dh):
    h, w, _ = img.shape
    S = img.astype(np.int64)
    xo, xc, xmax = linear_tab(w, dw, True)
    yo, yc, _ = linear_tab(h, dh, False)
    H = np.zeros((h, dw, 3), np.int64)
    for dx in range(dw):
        H[:, dx] = S[:, xo[dx]] * xc[dx][0] + S[:, xo[dx] + 1] * xc[dx][1] if dx < xmax else S[:, xo[dx]] * 2048
    out = np.zeros((dh, dw, 3), np.uint8)
    for dy in range(dh):
        r0, r1 = min(max(yo[dy], 0), h - 1), min(max(yo[dy] + 1, 0), h - 1)
        b0, b1 = yc[dy]
        out[dy] = np.clip((((b0 * (H[r0] >> 4)) >> 16) + ((b1 * (H[r1] >> 4)) >> 16) + 2) >> 2, 0, 255)
    return out


def cv_resize(img, size, interpolation):
    """cv2.resize(img, size=(w, h), interpolation='area' | 'bilinear') for uint8 [h, w, 3]."""
    dw, dh = size
    h, w, _ = img.shape
    if (dh, dw) == (h, w):
        return img.copy()
    if interpolation == "area":
        sx, sy = _scale(w, dw), _scale(h, dh)
        assert sx >= 1 and sy >= 1, "INTER_AREA up-scaling is a different (bilinear-like) code path in OpenCV"
        kx, ky = int(np.rint(sx)), int(np.rint(sy))
        if abs(sx - kx) < DBL_EPSILON and abs(sy - ky) < DBL_EPSILON:
            return resize_area_int(img, kx, ky)
        return resize_area(img, dw, dh)
    assert interpolation == "bilinear"
    return resize_linear(img, dw, dh)


def keep_ratio_resize(results, scale):
    """transforms.py:94-123 (scale a (w, h) tuple)."""
    image = results["img"]
    oh, ow = image.shape[:2]
    ratio = min(max(scale) / max(oh, ow), min(scale) / min(oh, ow))
    if ratio != 1:
        image = cv_resize(image, (int(ow * ratio), int(oh * ratio)), "area" if ratio < 1 else "bilinear")
    rh, rw = image.shape[:2]
    results.update(img=image, img_shape=image.shape[:2], scale_factor=(rw / ow, rh / oh))
    return results


def letter_resize(results, scale, allow_scale_up=True, pad_val=114):
    """transforms.py:180-272 and the scale_factor product of :318-325 (use_mini_pad / stretch_only / half_pad_param off)."""
    image = results["img"]
    sc = scale[::-1]
    shape = image.shape[:2]
    ratio = min(sc[0] / shape[0], sc[1] / shape[1])
    if not allow_scale_up:
        ratio = min(ratio, 1.0)
    no_pad = (int(round(shape[0] * ratio)), int(round(shape[1] * ratio)))
    ph, pw = sc[0] - no_pad[0], sc[1] - no_pad[1]
    if shape != no_pad:
        image = cv_resize(image, (no_pad[1], no_pad[0]), "bilinear")
    sf = (no_pad[1] / shape[1], no_pad[0] / shape[0])
    if "scale_factor" in results:
        sf = (sf[0] * results["scale_factor"][0], sf[1] * results["scale_factor"][1])
    top, left = int(round(ph // 2 - 0.1)), int(round(pw // 2 - 0.1))
    pads = [top, ph - top, left, pw - left]
    if any(pads):
        canvas = np.full((image.shape[0] + ph, image.shape[1] + pw, 3), pad_val, np.uint8)
        canvas[top: top + image.shape[0], left: left + image.shape[1]] = image
        image = canvas
    results.update(img=image, img_shape=image.shape, scale_factor=sf, pad_param=np.array(pads, dtype=np.float32))
    return results


def test_pipeline(img, scale=(640, 640), allow_scale_up=False, pad_val=114):
    """config/wedetect_base.py:111-118 for one decoded uint8 BGR image -> dict(img, img_shape, scale_factor, pad_param, ori_shape)."""
    res = dict(img=img, ori_shape=img.shape[:2])
    return letter_resize(keep_ratio_resize(res, tuple(scale)), tuple(scale), allow_scale_up, pad_val)


test_pipeline.__test__ = False      # not a pytest test
