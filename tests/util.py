import torch


def report_close(name, got, ref, rtol, atol):
    """Assert closeness with a diagnostic dump that localises layout bugs (rows / column chunks)."""
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = (err > tol) | torch.isnan(got)
    nbad = int(bad.sum())
    if nbad:
        g2 = got.reshape(-1, got.shape[-1])
        r2 = ref.reshape(-1, ref.shape[-1])
        b2 = bad.reshape(-1, bad.shape[-1])
        rows = b2.any(1).nonzero().flatten()
        cols = b2.any(0).nonzero().flatten()
        msg = [f"{name}: {nbad}/{bad.numel()} mismatches, max abs err {float(err[~torch.isnan(err)].max()) if (~torch.isnan(err)).any() else float('nan'):.4g}, "
               f"ref absmax {float(ref.abs().max()):.4g}, nan {int(torch.isnan(got).sum())}",
               f"  bad rows: {len(rows)} first {rows[:12].tolist()} last {rows[-4:].tolist()}",
               f"  bad cols: {len(cols)} first {cols[:12].tolist()} last {cols[-4:].tolist()}"]
        r0 = int(rows[0])
        c0 = int(b2[r0].nonzero()[0])
        msg.append(f"  got[{r0},{c0}:{c0+8}] = {g2[r0, c0:c0+8].tolist()}")
        msg.append(f"  ref[{r0},{c0}:{c0+8}] = {r2[r0, c0:c0+8].tolist()}")
        raise AssertionError("\n".join(msg))
    return float(err.max())


class FixtureTokenizer:
    """Replays tests/golden/tokens_*.json: what the reference's tokenizer (xlm-roberta-base/, absent on the GPU box) returned for
    the reference's own class-prompt lists under `tokenizer(text=..., return_tensors="pt", padding=True)` (mm_backbone.py:382-383)."""

    def __init__(self, name):
        import json
        import os
        d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"tokens_{name}.json"), encoding="utf-8"))
        self.texts, self.ids, self.mask = d["texts"], torch.tensor(d["input_ids"]), torch.tensor(d["attention_mask"])
        self.rows = {t: [i for i, m in zip(r, k) if m] for t, r, k in zip(d["texts"], d["input_ids"], d["attention_mask"])}

    def __call__(self, text, return_tensors="pt", padding=True):
        """Any sub-list of the fixture's prompts, padded to its longest member with <pad> = 1 like the real tokenizer does."""
        assert return_tensors == "pt" and padding is True and all(t in self.rows for t in text), "FixtureTokenizer only knows its own prompts"
        rows = [self.rows[t] for t in text]
        n = max(len(r) for r in rows)
        ids = torch.tensor([r + [1] * (n - len(r)) for r in rows])
        return dict(input_ids=ids, attention_mask=(torch.tensor([[1] * len(r) + [0] * (n - len(r)) for r in rows])))
