"""CPU models of the two exact shortcuts of csrc/postprocess.cu: they pin the *argument* that the shortcuts
cannot change the output, independently of the GPU runs that compare the kernels with the C oracle.

  * top-k select: keeping the keys whose 16 leading score bits reach the nms_pre-th best key (two 8-bit radix-select passes)
    leaves the sorted top nms_pre unchanged;
  * NMS cross-class early exit: a (class, image) block may stop once max_per_img confirmed survivors lie in rank buckets
    entirely before its next chunk, under ANY interleaving of the blocks; the first max_per_img survivors are unchanged.
"""
import random

import numpy as np


def _select_model(inv, nms_pre):
    """inv: uint32 array (30-bit inverted scores of one image, smaller = better).  Returns the boolean keep mask of
    pp_sel_hist / pp_sel_scan / pp_sel_compact."""
    n = inv.size
    if n <= nms_pre:
        return np.ones(n, dtype=bool)
    d1 = inv >> 22
    h1 = np.bincount(d1, minlength=256)
    cum = np.cumsum(h1)
    d1s = int(np.searchsorted(cum, nms_pre))                 # smallest digit with cumulative count >= nms_pre
    rem = nms_pre - (int(cum[d1s - 1]) if d1s > 0 else 0)
    d2 = (inv >> 14) & 255
    h2 = np.bincount(d2[d1 == d1s], minlength=256)
    cum2 = np.cumsum(h2)
    d2s = int(np.searchsorted(cum2, rem))
    cut16 = (d1s << 8) | d2s
    return (inv >> 14) <= cut16


def test_topk_select_model_keeps_the_sorted_top():
    rng = np.random.default_rng(0)
    cases = []
    for n, nms_pre in [(50, 100), (100, 100), (101, 100), (5000, 300), (20000, 1000), (70000, 30000)]:
        scores = rng.random(n).astype(np.float32)
        cases.append((scores, nms_pre))
        cases.append((np.round(scores * 8) / 8 + np.float32(1e-3), nms_pre))                     # tie-heavy: 9 distinct values
        cases.append((np.full(n, 0.5, dtype=np.float32), nms_pre))                                # all equal
        cases.append(((scores * 1e-6 + 0.731).astype(np.float32), nms_pre))                       # packed into a few prefixes
        cases.append((np.sort(scores)[::-1].copy(), nms_pre))
    for scores, nms_pre in cases:
        bits = scores.view(np.uint32) & 0x3FFFFFFF
        inv = (0x3FFFFFFF - bits).astype(np.uint32)
        idx = np.arange(scores.size, dtype=np.uint64)
        key = (inv.astype(np.uint64) << np.uint64(32)) | idx                                      # (score desc, flat index asc)
        keep = _select_model(inv, nms_pre)
        k = min(nms_pre, scores.size)
        assert int(keep.sum()) >= k
        want = np.sort(key)[:k]
        got = np.sort(key[keep])[:k]
        assert np.array_equal(want, got)
        if scores.size > nms_pre:                       # everything dropped ranks strictly after everything kept
            assert key[~keep].size == 0 or key[~keep].min() > key[keep].max()


def _iou(a, b):
    xx1, yy1, xx2, yy2 = max(a[0], b[0]), max(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3])
    w, h = max(0.0, xx2 - xx1), max(0.0, yy2 - yy1)
    inter = w * h
    return inter / ((a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter + 1e-12)


def _nms_full(boxes, labels, thr):
    """class-aware greedy NMS over candidates already in rank (score) order -> keep flags."""
    keep = np.zeros(len(boxes), dtype=bool)
    kept = {}
    for r in range(len(boxes)):
        c = int(labels[r])
        if all(_iou(boxes[k], boxes[r]) <= thr for k in kept.get(c, [])):
            keep[r] = True
            kept.setdefault(c, []).append(r)
    return keep


def _nms_blocks_model(boxes, labels, thr, max_per_img, chunk, bucket, rnd):
    """One block per class walks its candidates in chunks; blocks are interleaved in a random order chunk by chunk; the
    early-exit rules of pp_nms_kernel apply (own class has max_per_img survivors; or max_per_img confirmed survivors lie in
    buckets entirely before the next chunk's first rank)."""
    n = len(boxes)
    keep = np.zeros(n, dtype=bool)
    nb = (n + bucket - 1) // bucket + 1
    kept_hist = np.zeros(nb, dtype=np.int64)
    blocks = {}
    for r in range(n):
        blocks.setdefault(int(labels[r]), []).append(r)
    state = {c: dict(pos=0, kept=[]) for c in blocks}
    live = list(blocks)
    while live:
        c = rnd.choice(live)
        st, ranks = state[c], blocks[c]
        if st["pos"] >= len(ranks) or len(st["kept"]) >= max_per_img:
            live.remove(c)
            continue
        lim = ranks[st["pos"]] // bucket
        if kept_hist[:lim].sum() >= max_per_img:
            live.remove(c)
            continue
        for r in ranks[st["pos"]: st["pos"] + chunk]:
            if all(_iou(boxes[k], boxes[r]) <= thr for k in st["kept"]):
                keep[r] = True
                st["kept"].append(r)
                kept_hist[r // bucket] += 1
        st["pos"] += chunk
    return keep


def test_nms_early_exit_model_is_schedule_independent():
    rnd = random.Random(3)
    rng = np.random.default_rng(4)
    for trial in range(12):
        n, K = int(rng.integers(50, 600)), int(rng.integers(1, 9))
        max_per_img = int(rng.integers(5, 60))
        ctr = rng.random((n, 2)) * 100
        wh = rng.random((n, 2)) * 30 + 2
        boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1)
        labels = rng.integers(0, K, n)
        if trial % 3 == 0:
            labels = np.where(rng.random(n) < 0.8, 0, labels)      # one dominant class: the long-segment case the exit is for
        full = _nms_full(boxes, labels, 0.5)
        want = np.flatnonzero(full)[:max_per_img]
        for _ in range(4):                                          # different interleavings of the same blocks
            got_flags = _nms_blocks_model(boxes, labels, 0.5, max_per_img, chunk=8, bucket=16, rnd=rnd)
            got = np.flatnonzero(got_flags)[:max_per_img]
            assert np.array_equal(got, want), (trial, n, K, max_per_img)
            assert not (got_flags & ~full).any()                   # a block never keeps something full NMS suppresses
