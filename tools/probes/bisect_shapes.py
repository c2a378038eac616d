"""Debug probe: logit / distance errors of the default mode vs the fp32 oracle for a few (B, K) at 320 x 320."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import test_gpu_e2e as T
for B, K, u8 in ((1, 12, True), (4, 12, True), (1, 80, False), (4, 80, False), (3, 12, True)):
    errs, det, det_ref, p, ref = T.run_case("base", B, 320, 320, K, uni=False, precise=True, regime="sparse", input_u8=u8)
    print(B, K, u8, {k: round(v, 6) if isinstance(v, float) else v for k, v in errs.items()}, flush=True)
