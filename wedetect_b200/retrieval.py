"""Object-retrieval path: corpus embedding extraction + image x class scoring (BASELINE config 5).

Host-side mirror of the reference's two scripts, backed by the sm_100a library (no CPU / PyTorch fallback):

    eval_retrieval/extract_embedding.py:1653-1774   per image: <= 300 proposals -> {embedding [n,768], scale [n], bias [n]},
                                                    sharded with InferenceSampler, merged with all_gather_object, saved as
                                                    {"image_embedding": [...], "text_embedding": [K,768]}
    eval_retrieval/retrieval_metric.py:362-378      per image: sigmoid(emb @ text^T * exp(scale) + bias).max(0) > thre
    eval_retrieval/retrieval_metric.py:14-47        per-class precision / recall / F1

What changes underneath: the score of an image for every class is computed on the device that produced the proposals
(`RetrievalScorer`: row-scaled cast -> tcgen05 GEMM against the text matrix -> sigmoid/max reduce), so the only thing
ranks need to exchange is ONE all-gather of fixed-shape `[N_local, K]` score rows (or, to keep the reference's .pth
format, of the padded `[N_local, P, 770]` embedding blocks) instead of pickled Python lists.
"""
import torch

from . import _lib as L
from . import dist as wdist
from . import ops, schema
from .ops import P3


class RetrievalScorer:
    """scores[b, k] = max_{j < counts[b]} sigmoid((emb[b,j] . text[k]) * exp(scale[b,j]) + bias[b,j]).

    emb / scale / bias / counts may be the live result buffers of a `plan.VisionPlan(extract=True)` (zero-copy: the scorer
    then runs right behind the detector on the same stream) or buffers owned here and filled by `load()`.
    """

    def __init__(self, text_embedding, B, P, *, device="cuda:0", precise=True, emb=None, scale=None, bias=None, counts=None):
        L.load(require_gpu=True)
        self.dev = torch.device(device)
        K, C = text_embedding.shape
        assert C == schema.EMBED_DIM
        self.B, self.P, self.K, self.C = B, P, K, C
        self.K_pad = (K + 7) // 8 * 8
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.emb = emb if emb is not None else torch.zeros(B, P, C, **f32)
        self.scale = scale if scale is not None else torch.zeros(B, P, **f32)
        self.bias = bias if bias is not None else torch.zeros(B, P, **f32)
        self.counts = counts if counts is not None else torch.zeros(B, dtype=torch.int32, device=self.dev)
        assert self.emb.shape == (B, P, C) and self.scale.shape == (B, P) and self.bias.shape == (B, P) and self.counts.shape == (B,)
        self.text_f32 = torch.zeros(self.K_pad, C, **f32)
        self.text_f32[:K] = text_embedding.to(self.dev, torch.float32)
        if precise:   # fp16 hi/lo planes at the matrix's own power-of-two scale (split on the host, once)
            self.text = P3.from_f32(self.text_f32.cpu(), self.dev)
        else:
            self.text = P3.zeros((self.K_pad, C), self.dev, False)
            L.Program([ops.cast_bf16(self.text_f32, self.text)]).run(torch.cuda.current_stream().cuda_stream)
        self.rows = P3.zeros((B * P, C), self.dev, precise)
        self.z = torch.zeros(B * P, self.K_pad, **f32)
        self.scores = torch.zeros(B, K, **f32)
        self.program = L.Program([
            ops.scale_rows(self.emb, self.rows, scale=self.scale, counts=self.counts),
            ops.linear(self.rows, self.text, self.z),
            ops.retr_reduce(self.z, self.scores, P=P, bias=self.bias, counts=self.counts),
        ])

    def load(self, embeddings, scales, biases):
        """Fill the scorer's own buffers from per-image tensors (the reference's saved `image_embedding` entries)."""
        n_img = len(embeddings)
        assert n_img <= self.B
        emb = torch.zeros(self.B, self.P, self.C)
        sc = torch.zeros(self.B, self.P)
        bi = torch.zeros(self.B, self.P)
        cnt = torch.zeros(self.B, dtype=torch.int32)
        for b, (e, s, t) in enumerate(zip(embeddings, scales, biases)):
            n = e.shape[0]
            if n > self.P:
                raise ValueError(f"image {b} has {n} proposals, scorer was built for {self.P}")
            emb[b, :n], sc[b, :n], bi[b, :n], cnt[b] = e.float(), s.float().reshape(-1), t.float().reshape(-1), n
        self.emb.copy_(emb, non_blocking=True)
        self.scale.copy_(sc, non_blocking=True)
        self.bias.copy_(bi, non_blocking=True)
        self.counts.copy_(cnt, non_blocking=True)

    def run(self, stream=None):
        self.program.run(torch.cuda.current_stream().cuda_stream if stream is None else stream)
        return self.scores


# ------------------------------------------------------------------------------------------------------------------
# corpus level
# ------------------------------------------------------------------------------------------------------------------
def gather_rows(local, total, group=None):
    """All-gather per-image rows from contiguous shards (dist.shard_indices) into `[total, ...]` in image order.
    Shards may differ by one row: every rank pads to the largest shard so that ONE fixed-shape all-gather suffices."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert local.shape[0] == total
        return local
    world = dist.get_world_size(group)
    sizes = [len(wdist.shard_indices(total, world, r)) for r in range(world)]
    pad = max(sizes)
    blk = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    blk[: local.shape[0]] = local
    out = torch.empty((world * pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, blk, group=group)
    return torch.cat([out[r * pad: r * pad + sizes[r]] for r in range(world)], 0)


def extract_corpus(model, images, image_ids, *, batch_size=32, text_embedding=None, keep_embeddings=True, group=None):
    """The loop of extract_embedding.py:1718-1774 on this rank's shard of `images` (a sequence of PIL images / paths, or a
    float tensor [N,3,H,W] of already letterboxed inputs).  Returns the reference's .pth payload (on every rank):
        {"image_embedding": [{image_id, embedding [n,768], scale [n], bias [n]}, ...], "text_embedding": text_embedding}
    plus, when `text_embedding` is given, "scores" [N, K] = the image x class retrieval scores computed on the device.
    `keep_embeddings=False` skips the (large) embedding gather and returns only ids + scores."""
    import torch.distributed as dist
    if not getattr(model, "extract", False):
        raise ValueError("extract_corpus needs SimpleYOLOWorldDetector(..., extract=True)")
    dist_on = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if dist_on else 1
    rank = dist.get_rank(group) if dist_on else 0
    total = len(image_ids)
    mine = wdist.shard_indices(total, world, rank)
    P, C, dev = model.num_proposals, schema.EMBED_DIM, model.device
    blocks, score_rows = [], []
    for s in range(0, len(mine), batch_size):
        idx = list(mine[s: s + batch_size])
        n = len(idx)
        if torch.is_tensor(images):
            batch = images[idx]
            if n < batch_size:     # fixed-shape plans: pad the last batch with copies of its first image
                batch = torch.cat([batch, batch[:1].expand(batch_size - n, -1, -1, -1)], 0)
            model.forward_tensor(batch)
        else:
            batch = [images[i] for i in idx]
            model.forward(batch + [batch[0]] * (batch_size - n))
        r = model.last_batch_result
        if text_embedding is not None:
            score_rows.append(model.score_text(text_embedding)[:n].clone())
        if keep_embeddings:
            blk = torch.zeros(n, P + 1, C + 2, device=dev)
            blk[:, :P, :C] = r["embeddings"][:n]
            blk[:, :P, C] = r["scales"][:n]
            blk[:, :P, C + 1] = r["bias"][:n]
            blk[:, P, 0] = r["counts"][:n].float()
            blocks.append(blk)
    out = {"text_embedding": text_embedding}
    ids = torch.as_tensor(list(image_ids), dtype=torch.int64)
    if text_embedding is not None:
        K = text_embedding.shape[0]
        local = torch.cat(score_rows, 0) if score_rows else torch.zeros(0, K, device=dev)
        out["scores"] = gather_rows(local, total, group).cpu()
    out["image_ids"] = ids
    if keep_embeddings:
        local = torch.cat(blocks, 0) if blocks else torch.zeros(0, P + 1, C + 2, device=dev)
        full = gather_rows(local, total, group).cpu()
        res = []
        for i in range(total):
            n = int(full[i, P, 0])
            res.append({"image_id": int(ids[i]), "embedding": full[i, :n, :C].clone(), "scale": full[i, :n, C].clone(), "bias": full[i, :n, C + 1].clone()})
        out["image_embedding"] = res
    return out


def save_corpus(path, corpus):
    """Same file layout as extract_embedding.py:1774 (readable by the reference's retrieval_metric.py)."""
    torch.save({"image_embedding": corpus["image_embedding"], "text_embedding": corpus["text_embedding"]}, path)


def score_saved(pred, *, device="cuda:0", batch_size=64, model="wedetect", precise=True):
    """Image x class scores [N, K] for a saved corpus (the loop body of retrieval_metric.py:365-373, batched on the GPU).
    model == 'hqclip': plain sigmoid(logits) (no scale / bias), as the reference's switch at :369-370."""
    items = pred["image_embedding"]
    text = pred["text_embedding"].float()
    P = max([1] + [int(it["embedding"].shape[0]) for it in items])
    sc = RetrievalScorer(text, batch_size, P, device=device, precise=precise)
    out = torch.zeros(len(items), text.shape[0])
    for s in range(0, len(items), batch_size):
        chunk = items[s: s + batch_size]
        embs = [it["embedding"] for it in chunk]
        if model == "hqclip":
            zeros = [torch.zeros(e.shape[0]) for e in embs]
            sc.load(embs, zeros, zeros)
        else:
            sc.load(embs, [it["scale"] for it in chunk], [it["bias"] for it in chunk])
        out[s: s + len(chunk)] = sc.run()[: len(chunk)].cpu()
    return out


def predictions_from_scores(scores, image_ids, classnames, thre=0.3):
    """PREDICTIONS[classname] = [image ids whose score for that class exceeds thre] (retrieval_metric.py:374-377)."""
    pred = {name: [] for name in classnames}
    hit = (scores > thre).nonzero().tolist()
    for i, k in hit:
        pred[classnames[k]].append(int(image_ids[i]))
    return pred


def evaluate_retrieval_per_class(predictions, gt):
    """Per-class precision / recall / F1 over image-id sets (retrieval_metric.py:14-47; classes without GT are skipped)."""
    results = {}
    for cat_name, gt_set in gt.items():
        if len(gt_set) == 0:
            continue
        pred_set = set(map(int, predictions.get(cat_name, [])))
        tp, fp, fn = len(pred_set & gt_set), len(pred_set - gt_set), len(gt_set - pred_set)
        precision = tp / (tp + fp) if (tp + fp) > 0 else 0.0
        recall = tp / (tp + fn) if (tp + fn) > 0 else 0.0
        f1 = 2 * precision * recall / (precision + recall) if (precision + recall) > 0 else 0.0
        results[cat_name] = {"precision": round(precision, 4), "recall": round(recall, 4), "f1": round(f1, 4), "support": len(gt_set),
                             "n_pred": len(pred_set)}
    return results
