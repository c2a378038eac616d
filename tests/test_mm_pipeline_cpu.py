"""The mmcv test pipeline (WeDetectKeepRatioResize + WeDetectLetterResize, transforms.py:28-328) on the CPU side:
  * oracle/mm_pipeline.py is pinned against the reference's unmodified transform classes (tests/golden/mm_pipeline.npz) and
    against the installed cv2 on random sizes;
  * the product's geometry and OpenCV coefficient tables (wedetect_b200/preprocess.py) reproduce the oracle when the table format
    of WD_OP_CV_RESIZE_PAD is executed by a few lines of numpy that follow the kernel's loops.
"""
import os

import numpy as np
import pytest

from oracle import mm_pipeline as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mm_pipeline.npz")


def seeded_image(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def golden_cases():
    z = np.load(GOLD)
    scale = tuple(int(v) for v in z["scale"])
    for i, (h, w) in enumerate(z["cases"]):
        yield i, int(h), int(w), scale, z


def test_oracle_matches_reference_transforms():
    n = 0
    for i, h, w, scale, z in golden_cases():
        res = O.test_pipeline(seeded_image(1000 + i, h, w), scale=scale)
        assert np.array_equal(res["img"], z[f"img_{i}"]), (h, w)
        assert tuple(res["img_shape"]) == tuple(z[f"img_shape_{i}"])
        assert tuple(res["scale_factor"]) == tuple(z[f"scale_factor_{i}"])            # identical float64 arithmetic
        assert np.array_equal(res["pad_param"], z[f"pad_param_{i}"]) and res["pad_param"].dtype == np.float32
        n += 1
    assert n >= 15


def test_oracle_resize_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    codes = dict(area=cv2.INTER_AREA, bilinear=cv2.INTER_LINEAR)
    for t in range(60):
        h, w = int(rng.integers(6, 80)), int(rng.integers(6, 80))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if t % 3 == 0:       # integer boxes
            kx, ky = int(rng.integers(1, 5)), int(rng.integers(1, 5))
            img = rng.integers(0, 256, (h * ky, w * kx, 3), dtype=np.uint8)
            size, interp = (w, h), "area"
        elif t % 3 == 1:
            size, interp = (int(rng.integers(3, w + 1)), int(rng.integers(3, h + 1))), "area"
        else:
            size, interp = (int(rng.integers(3, 3 * w)), int(rng.integers(3, 3 * h))), "bilinear"
        want = cv2.resize(img, size, interpolation=codes[interp])
        assert np.array_equal(O.cv_resize(img, size, interp), want), (img.shape, size, interp)


def run_tables(img, desc, coef, H, W, pad):
    """numpy executor of one image of WD_OP_CV_RESIZE_PAD (csrc/preprocess.cu cv_resize_pad_kernel), loop for loop."""
    sw, sh, nw, nh, left, top, mode, off, kx, ky, sbits, xmax = (int(v) for v in desc[2:14])
    tab = coef[off:]
    out = np.full((H, W, 3), pad, np.uint8)
    S = img
    f32 = np.float32
    if mode == 0:
        res = S.copy()
    elif mode == 1:
        xidx, yidx = tab[: nw + 1], tab[nw + 1: nw + nh + 2]
        nx, ny = int(xidx[nw]), int(yidx[nh])
        p = nw + nh + 2
        xs, xa = tab[p: p + nx], tab[p + nx: p + 2 * nx].view(np.float32)
        ys, yb = tab[p + 2 * nx: p + 2 * nx + ny], tab[p + 2 * nx + ny: p + 2 * nx + 2 * ny].view(np.float32)
        res = np.zeros((nh, nw, 3), np.uint8)
        Sf = S.astype(np.float32)
        rows = {}
        for dy in range(nh):
            total = None
            for j in range(int(yidx[dy]), int(yidx[dy + 1])):
                sy = int(ys[j])
                if sy not in rows:
                    hb = np.zeros((nw, 3), np.float32)
                    for dx in range(nw):
                        acc = np.zeros(3, np.float32)
                        for k in range(int(xidx[dx]), int(xidx[dx + 1])):
                            acc = acc + Sf[sy, int(xs[k])] * xa[k]
                        hb[dx] = acc
                    rows[sy] = hb
                total = yb[j] * rows[sy] if total is None else total + yb[j] * rows[sy]
            res[dy] = np.clip(np.rint(total), 0, 255).astype(np.uint8)
    elif mode == 2:
        s = S[: nh * ky, : nw * kx].astype(np.int64).reshape(nh, ky, nw, kx, 3).sum(axis=(1, 3))
        scale = np.array(sbits, np.int32).view(np.float32)
        res = ((s + 2) >> 2).astype(np.uint8) if (kx, ky) == (2, 2) else np.clip(np.rint(s.astype(np.float32) * scale), 0, 255).astype(np.uint8)
    else:
        xofs, xab, yofs, yab = tab[:nw], tab[nw: 2 * nw], tab[2 * nw: 2 * nw + nh], tab[2 * nw + nh: 2 * nw + 2 * nh]
        Si = S.astype(np.int64)
        res = np.zeros((nh, nw, 3), np.uint8)
        for dy in range(nh):
            sy = int(yofs[dy])
            b0, b1 = int(np.int16(yab[dy] & 0xffff)), int(yab[dy] >> 16)
            r0, r1 = min(max(sy, 0), sh - 1), min(max(sy + 1, 0), sh - 1)
            for dx in range(nw):
                sx = int(xofs[dx])
                a0, a1 = int(np.int16(xab[dx] & 0xffff)), int(xab[dx] >> 16)
                if dx < xmax:
                    h0, h1 = Si[r0, sx] * a0 + Si[r0, sx + 1] * a1, Si[r1, sx] * a0 + Si[r1, sx + 1] * a1
                else:
                    h0, h1 = Si[r0, sx] * 2048, Si[r1, sx] * 2048
                res[dy, dx] = np.clip((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2, 0, 255)
    out[top: top + nh, left: left + nw] = res
    return out


def test_product_tables_reproduce_the_golden_pipeline():
    from wedetect_b200 import preprocess as P
    for i, h, w, scale, z in golden_cases():
        if h * w > 60000:
            continue                                     # the loop executor is slow; the big case runs on the GPU
        img = seeded_image(1000 + i, h, w)
        pk = P.pack_mm_batch([img], scale[1], scale[0], scale=scale, allow_scale_up=False)
        got = run_tables(img, pk["desc"][0], pk["coef"].view(np.int32), scale[1], scale[0], 114)
        assert np.array_equal(got, z[f"img_{i}"]), (h, w, int(pk["desc"][0][8]))
        g = pk["metas"][0]
        assert tuple(g["scale_factor"]) == tuple(z[f"scale_factor_{i}"]) and np.array_equal(g["pad_param"], z[f"pad_param_{i}"])
        assert g["img_shape"] == tuple(z[f"img_shape_{i}"]) and g["ori_shape"] == (h, w)


def test_tables_match_oracle_tables_on_many_sizes():
    """Vectorised OpenCV tables (product) == the loop restatement (oracle), including the sizes of the shipped 640 canvas."""
    from wedetect_b200 import preprocess as P
    rng = np.random.default_rng(11)
    sizes = [(1280, 640), (1920, 640), (1333, 640), (4032, 640), (641, 640), (500, 640), (333, 426), (7, 96)]
    sizes += [(int(a), int(b)) for a, b in zip(rng.integers(2, 3000, 40), rng.integers(2, 700, 40))]
    for s, d in sizes:
        if s >= d:
            idx, si, al = P.cv_area_table(s, d)
            want = O.area_tab(s, d)
            assert len(want) == len(si) and [t[1] for t in want] == si.tolist()
            assert np.array_equal(np.array([t[2] for t in want], np.float32), al)
            assert np.array_equal(np.repeat(np.arange(d), np.diff(idx)), np.array([t[0] for t in want]))
        for clamp in (True, False):
            ofs, packed, xmax = P.cv_linear_table(s, d, clamp)
            wo, wc, wx = O.linear_tab(s, d, clamp)
            assert ofs.tolist() == wo and xmax == wx
            assert [(int(np.int16(p & 0xffff)), int(p >> 16)) for p in packed] == wc


def test_geometry_branches():
    from wedetect_b200 import preprocess as P
    g = P.mm_test_geometry(480, 640)                      # already fits: pad only
    assert g["interp"] is None and g["resize"] == (480, 640) and g["pads"] == (80, 80, 0, 0) and g["scale_factor"] == (1.0, 1.0)
    g = P.mm_test_geometry(427, 640, scale=(640, 640))    # odd padding: top = round(213 // 2 - 0.1) = 106, bottom = 107
    assert g["pads"] == (106, 107, 0, 0) and g["pad_param"].tolist() == [106.0, 107.0, 0.0, 0.0]
    g = P.mm_test_geometry(1080, 1920)
    assert g["interp"] == "area" and g["resize"] == (360, 640) and P.cv_resize_plan(1080, 1920, 360, 640, "area")[0] == P.CV_AREA_INT
    g = P.mm_test_geometry(375, 500)
    assert g["interp"] == "bilinear" and g["resize"] == (480, 640) and g["scale_factor"] == (640 / 500, 480 / 375)
    # WeDetectLetterResize alone (no keep-ratio stage): bilinear shrink to the rounded size
    g = P.mm_test_geometry(800, 1333, keep_ratio_first=False)
    assert g["interp"] == "bilinear" and g["resize"] == (int(round(800 * 640 / 1333)), 640)
    # 1077 * (640 / 1077) = 639.9999999999999: the first stage truncates to 639, the second pads one more pixel (no up-scaling)
    g = P.mm_test_geometry(1077, 500)
    assert g["resize"][0] == 639 and g["pads"][:2] == (0, 1)
    with pytest.raises(NotImplementedError):             # ... and with allow_scale_up the second stage would resize again
        P.mm_test_geometry(1077, 500, allow_scale_up=True)


def test_exif_orientation_parser():
    """cv2.imdecode rotates by the EXIF orientation, nvJPEG does not: such files must be routed to the host decoder."""
    import io
    from PIL import Image
    from wedetect_b200.preprocess import exif_orientation
    im = Image.fromarray(np.zeros((8, 8, 3), np.uint8))
    buf = io.BytesIO()
    im.save(buf, "JPEG")
    assert exif_orientation(buf.getvalue()) == 1
    for val in (1, 3, 6, 8):
        ex = Image.Exif()
        ex[0x0112] = val
        buf = io.BytesIO()
        im.save(buf, "JPEG", exif=ex.tobytes())
        assert exif_orientation(buf.getvalue()) == val
    assert exif_orientation(b"\xff\xd8\xff\xd9") == 1 and exif_orientation(b"") == 1


def test_read_encoded_routes_files_to_the_right_decoder(tmp_path):
    """Which inputs stay compressed for the device decoder: 1- / 3-component JPEG bytes or files; PNGs, arrays, CMYK JPEGs and (on the
    cv2-compatible path) EXIF-rotated JPEGs go to the host decoder.  The nvJPEG header parse is stubbed (no GPU here)."""
    import io
    from PIL import Image
    from wedetect_b200.preprocess import EncodedImage, pack_mm_batch, read_encoded

    class StubDecoder:
        def info(self, blob):
            im = Image.open(io.BytesIO(blob))
            return im.size[0], im.size[1], len(im.getbands()), 0

    def jpeg(mode="RGB", size=(40, 30), **kw):
        b = io.BytesIO()
        Image.new(mode, size).save(b, "JPEG", **kw)
        return b.getvalue()

    dec = StubDecoder()
    e = read_encoded(jpeg(), dec)
    assert isinstance(e, EncodedImage) and (e.h, e.w) == (30, 40) and e.shape == (30, 40, 3) and e.size == 3600
    assert isinstance(read_encoded(jpeg("L"), dec), EncodedImage)
    assert read_encoded(jpeg("CMYK"), dec) is None
    png = io.BytesIO()
    Image.new("RGB", (8, 8)).save(png, "PNG")
    assert read_encoded(png.getvalue(), dec) is None and read_encoded(np.zeros((4, 4, 3), np.uint8), dec) is None
    ex = Image.Exif()
    ex[0x0112] = 6
    rotated = jpeg(exif=ex.tobytes())
    assert read_encoded(rotated, dec, honour_exif=True) is None and isinstance(read_encoded(rotated, dec), EncodedImage)
    f = tmp_path / "a.jpg"
    f.write_bytes(jpeg(size=(1280, 720)))
    e = read_encoded(str(f), dec)
    assert (e.h, e.w) == (720, 1280)
    # a compressed image packs like a decoded one of the same shape: same descriptor, no host bytes to copy
    pk_e = pack_mm_batch([e], 640, 640)
    pk_d = pack_mm_batch([np.zeros((720, 1280, 3), np.uint8)], 640, 640)
    assert np.array_equal(pk_e["desc"], pk_d["desc"]) and np.array_equal(pk_e["coef"], pk_d["coef"]) and pk_e["src_parts"][0][1] is e
