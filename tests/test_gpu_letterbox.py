"""GPU parity of the device-side letterbox (WD_OP_LETTERBOX): bit-exact against oracle/letterbox.py, which
tests/test_letterbox_cpu.py pins to PIL and to the reference's own letterbox()."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
D = "cuda:0"


def _check(lb, imgs, H, W):
    from oracle.letterbox import letterbox_ref
    ratios, offsets, shapes = lb.run(imgs)
    torch.cuda.synchronize()
    got = lb.out.cpu().numpy()
    for b, im in enumerate(imgs):
        want, r, off = letterbox_ref(im, (H, W))
        g = got[b].transpose(1, 2, 0)
        assert np.array_equal(g, want), (b, im.shape, int((g != want).sum()))
        assert ratios[b] == r and tuple(offsets[b]) == tuple(off) and shapes[b] == im.shape[:2]
    for b in range(len(imgs), lb.B):
        assert (got[b] == 114).all()


def test_letterbox_kernel_bit_exact_with_pil_oracle():
    from wedetect_b200.preprocess import Letterbox
    rng = np.random.default_rng(0)
    H = W = 640
    out = torch.zeros(8, 3, H, W, dtype=torch.uint8, device=D)
    lb = Letterbox(out)
    # (w, h): 4:3 down, 3:4 down, identity (PIL copies), up-scale, extreme wide / tall, Pillow>=12 vertical-first, 1-pixel-off sizes
    sizes = [(1280, 960), (480, 640), (640, 640), (97, 31), (1500, 40), (12, 1300), (641, 639), (7, 5)]
    _check(lb, [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (w, h) in sizes], H, W)
    # second batch: fewer images (unused slots are padding), larger sources (buffers grow, program is rebuilt), smooth content
    yy, xx = np.mgrid[0:1365, 0:2048]
    smooth = np.stack([xx * 255 // 2047, yy * 255 // 1364, (xx + yy) % 256], -1).astype(np.uint8)
    _check(lb, [smooth, rng.integers(0, 256, (427, 640, 3), dtype=np.uint8), smooth[:700, :333].copy()], H, W)
    # and a small batch again on the grown buffers
    _check(lb, [rng.integers(0, 256, (33, 57, 3), dtype=np.uint8)], H, W)
    with pytest.raises(ValueError):
        lb.run([])
    with pytest.raises(TypeError):
        lb.run([np.zeros((4, 4), dtype=np.uint8)])


def test_uni_forward_with_device_letterbox_matches_tensor_path():
    """SimpleYOLOWorldDetector.forward(images) = device letterbox + uint8 stem; forward_tensor(oracle-letterboxed fp32) is the
    path the other parity tests pin.  Same kept proposals (up to razor-edge ties), boxes mapped back to the source image."""
    from oracle import synth
    from oracle.letterbox import letterbox_ref
    from wedetect_b200.detector import SimpleYOLOWorldDetector
    sd = synth.synth_state_dict("base", seed=0, uni=True, regime="sparse")
    m = SimpleYOLOWorldDetector("base", 768, 256, 300, device=D, precise=True)
    m.load_state_dict(sd)
    m.img_size = (320, 320)                      # keep the test small; the facade's default is 640 (1280 for large)
    rng = np.random.default_rng(3)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (w, h) in [(500, 375), (200, 320)]]
    out = m.forward(imgs)
    kept = m.last_batch_result["anchors"].cpu().clone()
    res = [{k: v.cpu().clone() for k, v in o.items()} for o in out]
    canv, ratios, offs = zip(*[letterbox_ref(im, (320, 320)) for im in imgs])
    x = torch.from_numpy(np.stack(canv)).permute(0, 3, 1, 2).float() / 255.0
    ref = m.forward_tensor(x.to(D), ratios, offs, [im.shape[:2] for im in imgs])
    kept_ref = m.last_batch_result["anchors"].cpu()
    for b in range(2):
        n = len(ref[b]["scores"])
        assert len(res[b]["scores"]) == n
        a, r = set(kept[b, :n].tolist()), set(kept_ref[b, :n].tolist())
        assert len(a & r) >= 0.98 * len(r)
        if torch.equal(kept[b, :n], kept_ref[b, :n]):
            assert float((res[b]["scores"] - ref[b]["scores"].cpu()).abs().max()) <= 1e-4
            assert float((res[b]["bboxes"] - ref[b]["bboxes"].cpu()).abs().max()) <= 1e-2
        h, w = imgs[b].shape[:2]
        bb = res[b]["bboxes"]
        assert float(bb[:, 0::2].max()) <= w and float(bb[:, 1::2].max()) <= h and float(bb.min()) >= 0.0


def test_letterbox_full_size_sources():
    """Camera-sized sources (the shapes tools/bench_aux.py times): 1280x960 and 1920x1080 -> 640x640, bit-exact."""
    from wedetect_b200.preprocess import Letterbox
    rng = np.random.default_rng(11)
    out = torch.zeros(4, 3, 640, 640, dtype=torch.uint8, device=D)
    lb = Letterbox(out)
    imgs = [rng.integers(0, 256, (960, 1280, 3), dtype=np.uint8), rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8),
            rng.integers(0, 256, (1280, 960, 3), dtype=np.uint8)]
    _check(lb, imgs, 640, 640)
