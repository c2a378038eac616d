"""Where the end-to-end step time goes (GPU box): H2D copy, device program, result read-back, Python packaging."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import synth
from wedetect_b200 import _lib as L, schema
from wedetect_b200.detector import YOLOWorldDetector
from wedetect_b200.structures import DetDataSample


def main():
    torch.set_num_threads(bench.host_threads())
    dev = torch.device("cuda", 0)
    L.load(require_gpu=True)
    B, H, W, K = 32, 640, 640, 80
    sd = synth.synth_state_dict("base", seed=0, with_text=False, regime="sparse")
    model = YOLOWorldDetector(size="base", device=dev)
    model.load_state_dict(sd)
    model.set_text_features(torch.randn(K, schema.EMBED_DIM, generator=torch.Generator().manual_seed(5)))
    host = (synth.synth_images(B, H, W, seed=2) * 255).to(torch.uint8).flip(1).contiguous().pin_memory()
    samples = [DetDataSample(dict(ori_shape=(H, W), img_shape=(H, W), scale_factor=(1.0, 1.0), pad_param=(0.0, 0.0, 0.0, 0.0))) for _ in range(B)]
    data = dict(inputs=host, data_samples=samples)
    for _ in range(3):
        model.test_step(data)
    plan = model._plan(B, H, W, K, torch.uint8)
    plan.capture()
    for _ in range(3):
        model.test_step(data)
    torch.cuda.synchronize()

    def t(fn, n=10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return 1000 * (time.perf_counter() - t0) / n

    print(f"h2d copy ({host.numel() / 1e6:.1f} MB pinned)   {t(lambda: plan.image.copy_(host, non_blocking=True)):.3f} ms")
    print(f"device program (graph)         {t(plan.run):.3f} ms")
    print(f"test_step (copy+run+package)   {t(lambda: model.test_step(data)):.3f} ms")
    print(f"test_step + bulk D2H           {t(lambda: (model.test_step(data), {k: v.cpu() for k, v in model.last_batch_result.items()})):.3f} ms")
    r = plan.results()
    print(f"counts.cpu().tolist()          {t(lambda: r['counts'].cpu().tolist()):.3f} ms")


if __name__ == "__main__":
    main()
