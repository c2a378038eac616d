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
