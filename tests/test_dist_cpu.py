"""world_size-2 gloo test of the multi-GPU host logic (sharding + single all-gather of padded detections)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wedetect_b200 import dist as wd
    idx = list(wd.shard_indices(total, world, rank))
    B, M = len(idx), 5
    g = torch.Generator().manual_seed(100)
    allb = torch.rand(total, M, 4, generator=g)
    alls = torch.rand(total, M, generator=g)
    alll = torch.randint(0, 80, (total, M), generator=g)
    allc = torch.randint(0, M + 1, (total,), generator=g)
    blk = wd.pack_detections(allb[idx], alls[idx], alll[idx], allc[idx])
    out = wd.gather_detections(blk)
    b, s, l, c = wd.unpack_detections(out)
    ok = torch.equal(b, allb) and torch.equal(s, alls) and torch.equal(l, alll) and torch.equal(c, allc)
    q.put((rank, idx, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_indices_match_reference_sampler():
    from wedetect_b200.dist import shard_indices
    for total in (0, 1, 7, 64, 100000):
        for world in (1, 2, 8):
            parts = [list(shard_indices(total, world, r)) for r in range(world)]
            assert sum(parts, []) == list(range(total))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_all_gather_detections_world2():
    world, total = 2, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
    assert [r[1] for r in res] == [[0, 1, 2, 3], [4, 5, 6, 7]]
    assert all(r[2] for r in res)


def _gather_rows_worker(rank, world, port, total, q):
    import torch
    import torch.distributed as dist
    from wedetect_b200 import dist as wdist
    from wedetect_b200.retrieval import gather_rows
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = wdist.shard_indices(total, world, rank)
    local = torch.stack([torch.full((3,), float(i)) for i in mine]) if len(mine) else torch.zeros(0, 3)
    full = gather_rows(local, total)
    q.put((rank, full[:, 0].tolist()))
    dist.destroy_process_group()


def test_retrieval_gather_rows_unequal_shards_gloo():
    """C5's exchange step: ONE fixed-shape all-gather of per-image rows from contiguous, possibly unequal shards."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    total, world, port = 5, 2, 29653
    ps = [ctx.Process(target=_gather_rows_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in ps:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
    for r in range(world):
        assert got[r] == [0.0, 1.0, 2.0, 3.0, 4.0]


def test_retrieval_metric_restatement():
    """evaluate_retrieval_per_class / predictions_from_scores follow retrieval_metric.py:14-47,374-377."""
    import torch
    from wedetect_b200.retrieval import evaluate_retrieval_per_class, predictions_from_scores
    scores = torch.tensor([[0.9, 0.1, 0.31], [0.2, 0.8, 0.3], [0.5, 0.5, 0.0]])
    pred = predictions_from_scores(scores, [7, 8, 9], ["a", "b", "c"], thre=0.3)
    assert pred == {"a": [7, 9], "b": [8, 9], "c": [7]}          # strict >, ids in image order
    res = evaluate_retrieval_per_class(pred, {"a": {7}, "b": {8, 9, 10}, "c": set(), "d": {1}})
    assert res["a"] == dict(precision=0.5, recall=1.0, f1=0.6667, support=1, n_pred=2)
    assert res["b"] == dict(precision=1.0, recall=0.6667, f1=0.8, support=3, n_pred=2)
    assert "c" not in res and res["d"]["recall"] == 0.0


class _FakeExtractModel:
    """Stands in for SimpleYOLOWorldDetector(extract=True) in the corpus loop: proposals / scores derived from the image's
    first pixel so that order, padding of the last batch and per-image counts can be checked on the CPU."""
    extract, num_proposals, device = True, 4, torch.device("cpu")

    def forward_tensor(self, batch):
        B = batch.shape[0]
        v = batch[:, 0, 0, 0]                                        # image "id"
        cnt = (v.long() % 5).clamp(max=4).int()                      # 0..4 proposals
        emb = v[:, None, None].expand(B, 4, 768).clone()
        self.last_batch_result = dict(embeddings=emb, scales=torch.full((B, 4), -1.0), bias=v[:, None].expand(B, 4).clone(), counts=cnt)
        self.calls = getattr(self, "calls", 0) + 1

    def score_text(self, text):
        r = self.last_batch_result
        return r["embeddings"][:, 0, :1] * torch.ones(1, text.shape[0])


def test_extract_corpus_host_loop_single_process():
    """extract_embedding.py:1718-1774 loop: batches of 3 over 7 images (last batch padded), .pth payload layout, scores rows."""
    from wedetect_b200.retrieval import extract_corpus, predictions_from_scores
    N = 7
    imgs = torch.zeros(N, 3, 8, 8)
    imgs[:, 0, 0, 0] = torch.arange(10, 10 + N).float()
    text = torch.randn(5, 768)
    m = _FakeExtractModel()
    out = extract_corpus(m, imgs, list(range(100, 100 + N)), batch_size=3, text_embedding=text)
    assert m.calls == 3 and out["image_ids"].tolist() == list(range(100, 107)) and out["scores"].shape == (N, 5)
    assert out["scores"][:, 0].tolist() == [float(v) for v in range(10, 17)]
    for i, it in enumerate(out["image_embedding"]):
        n = (10 + i) % 5
        assert it["image_id"] == 100 + i and it["embedding"].shape == (n, 768) and it["scale"].shape == (n,) and it["bias"].shape == (n,)
        if n:
            assert float(it["embedding"][0, 0]) == 10 + i and float(it["bias"][0]) == 10 + i and float(it["scale"][0]) == -1.0
    assert out["text_embedding"] is text
    light = extract_corpus(m, imgs, list(range(N)), batch_size=4, text_embedding=text, keep_embeddings=False)
    assert "image_embedding" not in light and light["scores"].shape == (N, 5)
    pred = predictions_from_scores(out["scores"], out["image_ids"].tolist(), [f"c{k}" for k in range(5)], thre=12.5)
    assert pred["c0"] == [103, 104, 105, 106]


def _extract_worker(rank, world, port, q):
    import torch.distributed as dist
    from wedetect_b200.retrieval import extract_corpus
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    N = 7
    imgs = torch.zeros(N, 3, 8, 8)
    imgs[:, 0, 0, 0] = torch.arange(10, 10 + N).float()
    out = extract_corpus(_FakeExtractModel(), imgs, list(range(100, 100 + N)), batch_size=2, text_embedding=torch.ones(3, 768))
    q.put((rank, out["scores"][:, 0].tolist(), [int(it["embedding"].shape[0]) for it in out["image_embedding"]]))
    dist.destroy_process_group()


def test_extract_corpus_sharded_gloo():
    """Config 5's multi-GPU shape on CPU: 7 images over 2 ranks (shards of 4 and 3), every rank ends with the full payload in
    image order after ONE fixed-shape all-gather per array."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_extract_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = dict((r, (s, n)) for r, s, n in (q.get(timeout=180) for _ in range(2)))
    for p in ps:
        p.join(timeout=60)
    for r in range(2):
        assert got[r][0] == [float(v) for v in range(10, 17)] and got[r][1] == [(10 + i) % 5 for i in range(7)]
