"""The reference's test-time image transforms on the device (WD_OP_CV_RESIZE_PAD): WeDetectKeepRatioResize + WeDetectLetterResize
(transforms.py:94-123,180-272) with cv2's INTER_AREA / INTER_LINEAR arithmetic, bit-exact, and the image entry point built on it."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
D = "cuda:0"
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "mm_pipeline.npz")
CFG = os.path.join(HERE, "configs", "wedetect_base_min.py")


def seeded_image(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def _run(images, H, W, **pipe):
    from wedetect_b200.preprocess import MMTestPipeline
    out = torch.zeros(len(images), 3, H, W, dtype=torch.uint8, device=D)
    metas = MMTestPipeline(out, **pipe).run(images)
    torch.cuda.synchronize()
    return out.permute(0, 2, 3, 1).cpu().numpy(), metas


def test_device_pipeline_matches_reference_goldens():
    """Every fixture case (produced by the reference's unmodified transform classes) in ONE batch: pixels, scale_factor, pad_param."""
    z = np.load(GOLD)
    scale = tuple(int(v) for v in z["scale"])
    cases = [(int(h), int(w)) for h, w in z["cases"]]
    imgs = [seeded_image(1000 + i, h, w) for i, (h, w) in enumerate(cases)]
    got, metas = _run(imgs, scale[1], scale[0], scale=scale)
    for i, (h, w) in enumerate(cases):
        assert np.array_equal(got[i], z[f"img_{i}"]), (h, w)
        assert tuple(metas[i]["scale_factor"]) == tuple(z[f"scale_factor_{i}"]) and np.array_equal(metas[i]["pad_param"], z[f"pad_param_{i}"])
        assert metas[i]["img_shape"] == tuple(z[f"img_shape_{i}"]) and metas[i]["ori_shape"] == (h, w)


def test_device_pipeline_matches_cv2_at_640():
    """The shipped 640 x 640 canvas on photo-sized inputs, against the installed cv2 driven exactly as the transforms drive it
    (and against the oracle restatement where cv2 is missing): integer-box, fractional INTER_AREA, INTER_LINEAR, pad-only."""
    from oracle import mm_pipeline as O
    try:
        import cv2
    except ImportError:
        cv2 = None
    shapes = [(720, 1280), (1080, 1920), (800, 1333), (1333, 800), (427, 640), (640, 480), (375, 500), (1077, 500), (2000, 3008), (333, 333), (640, 640), (120, 90)]
    imgs = [seeded_image(7 + i, h, w) for i, (h, w) in enumerate(shapes)]
    got, metas = _run(imgs, 640, 640)
    for i, (h, w) in enumerate(shapes):
        if cv2 is not None:
            ratio = min(640 / max(h, w), 640 / min(h, w))
            im = imgs[i]
            if ratio != 1:
                im = cv2.resize(im, (int(w * ratio), int(h * ratio)), interpolation=cv2.INTER_AREA if ratio < 1 else cv2.INTER_LINEAR)
            t, b, l, r = (int(v) for v in metas[i]["pad_param"])
            want = cv2.copyMakeBorder(im, t, b, l, r, cv2.BORDER_CONSTANT, value=(114, 114, 114))
            assert want.shape == (640, 640, 3), (h, w, want.shape)
        else:
            want = O.test_pipeline(imgs[i])["img"]
        assert np.array_equal(got[i], want), (h, w, int(np.abs(got[i].astype(int) - want.astype(int)).max()))
        ref = O.test_pipeline(imgs[i]) if h * w <= 640 * 640 else None          # the loop oracle is slow on the big ones
        if ref is not None:
            assert np.array_equal(got[i], ref["img"]) and tuple(ref["scale_factor"]) == tuple(metas[i]["scale_factor"])
            assert np.array_equal(ref["pad_param"], metas[i]["pad_param"])


def test_pipeline_buffers_are_reused_and_partial_batches_pad():
    from wedetect_b200.preprocess import MMTestPipeline
    out = torch.zeros(4, 3, 96, 96, dtype=torch.uint8, device=D)
    pipe = MMTestPipeline(out, scale=(96, 96))
    a = [seeded_image(1, 144, 192), seeded_image(2, 50, 37)]
    pipe.run(a)
    first = out.clone()
    pipe.run([seeded_image(3, 300, 20)] * 4)
    pipe.run(a)                                           # two images in a batch of four: the other slots are plain padding
    torch.cuda.synchronize()
    assert torch.equal(out[:2], first[:2]) and bool((out[2:] == 114).all())


def test_predict_images_equals_test_step_on_the_oracle_pipeline():
    """model.predict_images(decoded BGR images) == model.test_step on what the reference's CPU pipeline would hand over: same input
    bytes, same metainfo, so the detections are identical bit for bit; the inference_detector tail on top."""
    from oracle import mm_pipeline as O, synth
    from wedetect_b200.api import DetDataSample, inference_detector, init_detector
    sd = synth.synth_state_dict("base", seed=0, with_text=False, regime="sparse")
    model = init_detector(CFG, checkpoint=dict(state_dict=sd), device=D)
    g = torch.Generator().manual_seed(3)
    model.set_text_features(torch.nn.functional.normalize(torch.randn(1, 12, 768, generator=g), dim=-1))
    imgs = [seeded_image(21, 480, 600), seeded_image(22, 1080, 1920), seeded_image(23, 300, 231)]
    out = model.predict_images(imgs)
    refs = [O.test_pipeline(im) for im in imgs]
    x = torch.from_numpy(np.stack([r["img"] for r in refs])).permute(0, 3, 1, 2).contiguous()
    samples = [DetDataSample(dict(ori_shape=r["ori_shape"], img_shape=r["img_shape"], scale_factor=r["scale_factor"], pad_param=r["pad_param"])) for r in refs]
    want = model.test_step(dict(inputs=x.to(D), data_samples=samples))
    for o, w_, r in zip(out, want, refs):
        assert len(o.pred_instances) == len(w_.pred_instances) > 0
        assert torch.equal(o.pred_instances.bboxes, w_.pred_instances.bboxes) and torch.equal(o.pred_instances.scores, w_.pred_instances.scores)
        assert torch.equal(o.pred_instances.labels, w_.pred_instances.labels)
        assert o.metainfo["ori_shape"] == r["ori_shape"] and tuple(o.metainfo["scale_factor"]) == tuple(r["scale_factor"])
        oh, ow = r["ori_shape"]
        b = o.pred_instances.bboxes
        assert float(b.min()) >= 0 and float(b[:, [0, 2]].max()) <= ow and float(b[:, [1, 3]].max()) <= oh     # rescaled into the ORIGINAL image
    thr = float(out[0].pred_instances.scores[min(20, len(out[0].pred_instances) - 1)])
    det = inference_detector(model, imgs[0], None, max_dets=10, score_thr=thr)
    keep = out[0].pred_instances.scores > thr
    assert len(det["confidence"]) == min(10, int(keep.sum())) and det["xyxy"].shape == (len(det["confidence"]), 4) and det["class_id"].dtype == np.int64
    assert np.array_equal(det["confidence"], out[0].pred_instances.scores[keep][:10].cpu().numpy())


def _photo_like(seed, h, w):
    """Smooth structure + mild noise: what a JPEG codec is built for (pure noise would measure the quantiser, not the decoder)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.stack([128 + 100 * np.sin(xx / (17 + 5 * c) + c) * np.cos(yy / (23 + 3 * c)) for c in range(3)], -1)
    img += rng.normal(0, 6, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def test_device_jpeg_decode_feeds_the_pipeline():
    """decode='nvjpeg': compressed bytes in, nvJPEG writes BGR pixels into the resize kernel's source buffer.  The geometry and
    metainfo are exactly those of the host-decoded path; pixels are compared with cv2.imdecode (libjpeg-turbo): the two
    decoders differ in IDCT rounding, colour conversion and chroma up-sampling: a few grey levels at most, < 1 on average."""
    cv2 = pytest.importorskip("cv2")
    from wedetect_b200.preprocess import MMTestPipeline, Letterbox, EncodedImage
    shapes = [(480, 640), (720, 1280), (375, 500), (333, 500)]
    blobs, decoded = [], []
    for i, (h, w) in enumerate(shapes):
        img = _photo_like(40 + i, h, w)
        sub = cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444 if i == 0 else cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420
        ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 92, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sub])
        assert ok
        blobs.append(enc.tobytes())
        decoded.append(cv2.imdecode(enc, cv2.IMREAD_COLOR))
    grey = cv2.imencode(".jpg", _photo_like(50, 200, 300)[:, :, 0])[1]
    blobs.append(grey.tobytes())
    decoded.append(cv2.imdecode(grey, cv2.IMREAD_COLOR))
    out_h = torch.zeros(len(blobs), 3, 640, 640, dtype=torch.uint8, device=D)
    out_d = torch.zeros_like(out_h)
    metas_h = MMTestPipeline(out_h).run(decoded)
    pipe = MMTestPipeline(out_d)
    items = pipe.encoded(blobs, lambda rest: [cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR) for b in rest])
    assert all(isinstance(it, EncodedImage) for it in items)
    metas_d = pipe.run(items)
    torch.cuda.synchronize()
    assert pipe.h2d_bytes < sum(d.size for d in decoded) // 4                       # only compressed bytes + tables crossed PCIe
    for a, b in zip(metas_h, metas_d):
        assert a["ori_shape"] == b["ori_shape"] and tuple(a["scale_factor"]) == tuple(b["scale_factor"]) and np.array_equal(a["pad_param"], b["pad_param"])
    diff = (out_h.int() - out_d.int()).abs()
    per_img = [int(diff[i].max()) for i in range(len(blobs))]
    mean = [float(diff[i].float().mean()) for i in range(len(blobs))]
    print("nvJPEG vs cv2 after resize/pad: max abs", per_img, "mean abs", [round(m, 3) for m in mean])
    # measured on B200 / nvJPEG 12.4 vs OpenCV 4.13 (libjpeg-turbo): max abs [4, 6, 9, 8, 1], mean abs 0.34-0.63 grey levels
    assert per_img[0] <= 6 and max(per_img) <= 12 and max(mean) <= 1.0
    # a PNG (not a JPEG) and an array go through the host decoder inside the same call
    png = cv2.imencode(".png", decoded[0])[1].tobytes()
    items = pipe.encoded([png, decoded[1]], lambda rest: [cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR) for b in rest])
    assert isinstance(items[0], np.ndarray) and np.array_equal(items[0], decoded[0]) and items[1] is decoded[1]
    # the Uni entry point's RGB order
    out_u = torch.zeros(1, 3, 640, 640, dtype=torch.uint8, device=D)
    lb = Letterbox(out_u)
    lb.run(lb.encoded([blobs[0]], None))
    ref_u = torch.zeros_like(out_u)
    Letterbox(ref_u).run([decoded[0][:, :, ::-1]])
    torch.cuda.synchronize()
    assert int((out_u.int() - ref_u.int()).abs().max()) <= 6


def test_predict_images_with_device_decode(tmp_path):
    """File names in, detections out, decode on the device: same boxes as the host-decoded run up to the decoder's pixel noise."""
    cv2 = pytest.importorskip("cv2")
    from oracle import synth
    from wedetect_b200.api import init_detector
    sd = synth.synth_state_dict("base", seed=0, with_text=False, regime="sparse")
    model = init_detector(CFG, checkpoint=dict(state_dict=sd), device=D)
    g = torch.Generator().manual_seed(3)
    model.set_text_features(torch.nn.functional.normalize(torch.randn(1, 12, 768, generator=g), dim=-1))
    files = []
    for i, (h, w) in enumerate([(480, 640), (600, 400)]):
        f = str(tmp_path / f"img{i}.jpg")
        cv2.imwrite(f, _photo_like(60 + i, h, w), [cv2.IMWRITE_JPEG_QUALITY, 95, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444])
        files.append(f)
    host = model.predict_images(files)
    dev = model.predict_images(files, decode="nvjpeg")
    for a, b in zip(host, dev):
        assert a.metainfo["ori_shape"] == b.metainfo["ori_shape"] and a.metainfo["img_path"] == b.metainfo["img_path"]
        n = min(len(a.pred_instances), len(b.pred_instances), 20)
        assert n > 0
        # the strongest detections survive +-1 grey level of decoder noise: same labels, boxes within a pixel
        la, lb_ = a.pred_instances.labels[:5].tolist(), b.pred_instances.labels[:5].tolist()
        assert len(set(la) & set(lb_)) >= 3
