"""Image-sharded data parallelism: the only parallelism on this path.

    shard_indices   contiguous shard of an image list, identical to the reference's InferenceSampler
                    (eval_retrieval/extract_embedding.py:1630-1638): rank r gets size//W + (r < size%W) items
    gather_detections  ONE all-gather of the fixed-shape padded result block per step, replacing the reference's
                    all_gather_object of pickled python lists (extract_embedding.py:1746-1756)
Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_indices(total, world, rank):
    base, rem = divmod(total, world)
    sizes = [base + (1 if r < rem else 0) for r in range(world)]
    begin = sum(sizes[:rank])
    return range(begin, begin + sizes[rank])


def pack_detections(boxes, scores, labels, counts):
    """[B,M,4], [B,M], [B,M] int, [B] int -> one fp32 block [B, M+1, 6]; row M carries the count."""
    B, M = scores.shape
    blk = torch.zeros(B, M + 1, 6, dtype=torch.float32, device=boxes.device)
    blk[:, :M, :4] = boxes
    blk[:, :M, 4] = scores
    blk[:, :M, 5] = labels.to(torch.float32)
    blk[:, M, 0] = counts.to(torch.float32)
    return blk


def unpack_detections(blk):
    M = blk.shape[1] - 1
    return blk[:, :M, :4], blk[:, :M, 4], blk[:, :M, 5].to(torch.int64), blk[:, M, 0].to(torch.int64)


def gather_detections(blk, group=None):
    """All ranks contribute the same-shaped block; returns [world*B, M+1, 6] ordered by rank (= image order
    for contiguous shards of equal size)."""
    world = dist.get_world_size(group)
    out = torch.empty((world * blk.shape[0],) + tuple(blk.shape[1:]), dtype=blk.dtype, device=blk.device)
    dist.all_gather_into_tensor(out, blk.contiguous(), group=group)
    return out
