"""XLM-R text tower on the GPU (TextPlan through the C ABI) vs the CPU fp32 oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu
D = "cuda:0"


def _ids(S, L, vocab, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, vocab, (S, L), generator=g, dtype=torch.int32)
    ids[:, 0] = 0
    for s in range(S):
        n = 2 + (s % (L - 1))
        ids[s, n - 1] = 2
        ids[s, n:] = 1
    return ids, (ids != 1).int()


@pytest.mark.parametrize("size,precise,tol", [("base", False, 3e-2), ("base", True, 2e-4), ("large", True, 2e-4)])
def test_text_tower(size, precise, tol):
    from oracle import functional as Fn, synth
    from wedetect_b200 import plan, schema, weights
    S, L, vocab = 21, 9, 2000
    sd = synth.synth_state_dict(size, seed=1, with_text=True, text_vocab=vocab, calibrate=False)
    ids, mask = _ids(S, L, vocab, 3)
    with torch.no_grad():
        ref = Fn.text_tower(sd, size, ids, mask)
    Wt = weights.prepare_text(sd, size, D, precise=precise)
    tp = plan.TextPlan(Wt, size, S, L)
    got = tp.run(ids.to(D), mask.to(D))
    torch.cuda.synchronize()
    err = float((got.cpu() - ref).abs().max())
    cos = float((got.cpu() * ref).sum(-1).min())
    print(f"text tower {size} precise={precise}: max abs err {err:.3e}, min cosine {cos:.6f}")
    assert err <= tol, err
    assert abs(float(got.norm(dim=-1).mean()) - 1.0) < 1e-4
