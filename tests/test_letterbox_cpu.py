"""Pins oracle/letterbox.py (the CPU restatement of generate_proposal.py's letterbox, i.e. PIL's 8-bit BILINEAR resampler)
bit-exactly: against PIL itself, against the reference's own letterbox() when /root/reference is mounted, against the
committed digests the reference produced, and checks that the product's host-side table builder agrees with the oracle."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def test_resize_oracle_is_bit_exact_with_pil():
    from PIL import Image
    from oracle.letterbox import resize_bilinear_ref
    rng = np.random.default_rng(1)
    cases = [(64, 48, 32, 24), (64, 48, 100, 75), (57, 33, 64, 37), (33, 57, 20, 64), (80, 60, 80, 60), (100, 100, 100, 50), (100, 100, 37, 100),
             (500, 375, 640, 480), (1000, 333, 640, 213), (7, 5, 64, 46), (300, 200, 64, 43), (13, 900, 9, 640), (1, 1, 5, 5), (5, 5, 1, 1),
             (2, 300, 1, 150), (1023, 17, 640, 11)]
    for _ in range(12):
        cases.append(tuple(int(v) for v in rng.integers(1, 400, 4)))
    for (w, h, ow, oh) in cases:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.Resampling.BILINEAR))
        got = resize_bilinear_ref(img, ow, oh)
        assert np.array_equal(got, want), (w, h, ow, oh, int(np.abs(got.astype(int) - want.astype(int)).max()))


def test_letterbox_oracle_matches_reference_digests():
    from make_golden_letterbox import source
    from oracle.letterbox import letterbox_ref
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "letterbox_digests.json")))
    for c in gold["cases"]:
        canvas, ratio, (dw, dh) = letterbox_ref(source(c["w"], c["h"], c["seed"]), (640, 640))
        assert hashlib.sha256(canvas.tobytes()).hexdigest() == c["sha256"], (c["w"], c["h"])
        assert ratio == c["ratio"] and dw == c["dw"] and dh == c["dh"]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not mounted")
def test_letterbox_oracle_matches_reference_live():
    sys.path.insert(0, REF)
    from PIL import Image
    import generate_proposal as gp
    from oracle.letterbox import letterbox_ref
    rng = np.random.default_rng(7)
    for (w, h) in [(417, 233), (233, 417), (640, 427), (1500, 1000), (50, 600), (640, 640)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want, ratio, off = gp.letterbox(Image.fromarray(img), (640, 640))
        got, r2, off2 = letterbox_ref(img, (640, 640))
        assert np.array_equal(got, np.asarray(want)) and ratio == r2 and tuple(off) == tuple(off2)


def test_product_resample_tables_match_oracle():
    """wedetect_b200.preprocess builds the kernel's window / weight tables on the host; they must be PIL's."""
    from oracle.letterbox import coeffs_ref, letterbox_params_ref
    from wedetect_b200.preprocess import letterbox_params, resample_tables
    rng = np.random.default_rng(3)
    pairs = [(640, 480), (480, 640), (1920, 640), (7, 640), (641, 640), (639, 640), (3000, 427), (2, 1), (1, 2)]
    pairs += [tuple(int(v) for v in rng.integers(1, 2500, 2)) for _ in range(60)]
    for a, b in pairs:
        if a == b:
            continue
        ks, bo, kk = resample_tables(a, b)
        ks2, bo2, kk2 = coeffs_ref(a, b)
        assert ks == ks2 and np.array_equal(bo, bo2) and np.array_equal(kk, kk2), (a, b)
    ks, bo, kk = resample_tables(9, 9)      # the pass PIL skips: identity window, weight 2^22 reproduces the byte
    assert ks == 1 and bo.tolist() == [[i, 1] for i in range(9)] and int(kk[0, 0]) == 1 << 22
    for _ in range(500):
        w, h = (int(v) for v in rng.integers(1, 5000, 2))
        assert letterbox_params(w, h, (640, 640)) == letterbox_params_ref(w, h, (640, 640))


def _kernel_model(pk, B, H, W, pad=114):
    """numpy transcription of csrc/preprocess.cu (lb_pass1_kernel / lb_pass2_paste_kernel) reading the packed buffers exactly
    as the device does: validates the host-side packing (descriptor words, table layout, offsets) without a GPU."""
    desc, coef, src = pk["desc"], pk["coef"].astype(np.int64), pk["src"].astype(np.int64)
    out = np.full((B, 3, H, W), pad, dtype=np.uint8)
    tmp = np.zeros(max(pk["tmp_bytes"], 1), dtype=np.int64)

    def px(buf, base, row_stride, bounds, kk, ks, along_x, row, col):
        o = col if along_x else row
        lo, cnt = int(coef[bounds + 2 * o]), int(coef[bounds + 2 * o + 1])
        p = base + (row * row_stride + lo * 3 if along_x else lo * row_stride + col * 3)
        tap = 3 if along_x else row_stride
        acc = np.full(3, 1 << 21, dtype=np.int64)
        for j in range(cnt):
            acc += buf[p + j * tap: p + j * tap + 3] * coef[kk + o * ks + j]
        return np.clip(acc >> 22, 0, 255)

    for b in range(desc.shape[0]):
        d = [int(v) for v in desc[b]]
        src_off, src_w, new_w, new_h, left, top, first, rows, tmp_off, coff, ksh, ksv, vfirst = d[0], d[2], d[4], d[5], d[6], d[7], d[8], d[9], d[10], d[12], d[13], d[14], d[15]
        bh = coff; kh = bh + 2 * new_w; bv = kh + new_w * ksh; kv = bv + 2 * new_h
        stride = src_w * 3
        cols = src_w if vfirst else new_w
        for y in range(rows):
            for x in range(cols):
                if vfirst:
                    v = px(src, src_off, stride, bv, kv, ksv, False, y, x)
                else:
                    v = px(src, src_off + first * stride, stride, bh, kh, ksh, True, y, x)
                tmp[tmp_off + (y * cols + x) * 3: tmp_off + (y * cols + x) * 3 + 3] = v
        rstride = cols * 3
        for yy in range(new_h):
            for xx in range(new_w):
                if vfirst:
                    v = px(tmp, tmp_off, rstride, bh, kh, ksh, True, yy, xx)
                else:
                    v = px(tmp, tmp_off, rstride, bv, kv, ksv, False, yy, xx)
                out[b, :, top + yy, left + xx] = v
    return out


def test_packed_batch_drives_kernel_model_to_oracle_result():
    from oracle.letterbox import letterbox_ref
    from wedetect_b200.preprocess import pack_batch
    rng = np.random.default_rng(5)
    for (H, W), sizes, vf in (((32, 32), [(50, 37), (20, 31), (32, 32), (9, 64), (64, 5), (32, 20)], None),    # (w, h): down, up, identity, tall, wide, h-only
                              ((640, 640), [(12, 1300), (31, 17)], 0)):                                     # Pillow >= 12 vertical-first case + an upscale
        imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (w, h) in sizes]
        pk = pack_batch(imgs, H, W)
        assert [int(v) for v in pk["desc"][:, 15]] == [1 if i == vf else 0 for i in range(len(imgs))]
        got = _kernel_model(pk, len(imgs) + 1, H, W)
        for b, im in enumerate(imgs):
            want, r, off = letterbox_ref(im, (H, W))
            assert np.array_equal(got[b].transpose(1, 2, 0), want), sizes[b]
            assert pk["ratios"][b] == r and tuple(pk["offsets"][b]) == tuple(off)
        assert (got[len(imgs)] == 114).all()            # unused slot: padding only
    with pytest.raises(ValueError):                     # the reference fails too (PIL: height and width must be > 0)
        pack_batch([np.zeros((400, 3, 3), dtype=np.uint8)], 32, 32)


def test_decode_images_accepts_paths_pil_and_arrays(tmp_path):
    """Host-side input handling of SimpleYOLOWorldDetector.forward (generate_proposal.py:1087-1092): paths are opened and
    converted to RGB, PIL images of any mode become RGB, arrays pass through; order is kept by the thread pool."""
    from PIL import Image
    from wedetect_b200.preprocess import decode_images
    rng = np.random.default_rng(2)
    a = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    p_png, p_jpg = str(tmp_path / "a.png"), str(tmp_path / "b.jpg")
    Image.fromarray(a).save(p_png)
    Image.fromarray(a).save(p_jpg, quality=90)
    gray = Image.fromarray(a[..., 0], mode="L")
    rgba = Image.fromarray(np.dstack([a, a[..., :1]]), mode="RGBA")
    out = decode_images([p_png, p_jpg, gray, rgba, a, tmp_path / "a.png"])
    assert [o.shape for o in out] == [(37, 53, 3)] * 6 and all(o.dtype == np.uint8 for o in out)
    assert np.array_equal(out[0], a) and np.array_equal(out[4], a) and np.array_equal(out[5], a)
    assert np.array_equal(out[1], np.asarray(Image.open(p_jpg).convert("RGB")))
    assert np.array_equal(out[2], np.asarray(gray.convert("RGB"))) and np.array_equal(out[3], np.asarray(rgba.convert("RGB")))
    with pytest.raises(TypeError):
        decode_images([np.zeros((4, 4), dtype=np.uint8), a])
