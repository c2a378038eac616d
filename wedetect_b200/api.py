"""Entry-point shims with the names the reference's scripts import.

    infer_wedetect.py:      from mmdet.apis import init_detector        ->  from wedetect_b200.api import init_detector
                            from mmengine.config import Config          ->  from wedetect_b200.api import Config
    generate_proposal.py:   SimpleYOLOWorldDetector(...)                ->  from wedetect_b200.api import SimpleYOLOWorldDetector
    eval_retrieval/extract_embedding.py:  SimpleYOLOWorldDetector(...)  ->  SimpleYOLOWorldDetector(..., extract=True) + extract_corpus
    eval_retrieval/retrieval_metric.py:   the per-image scoring loop    ->  score_saved / predictions_from_scores / evaluate_retrieval_per_class
    test.py / dist_test.sh:               runner.test() (mmengine TestLoop)  ->  TestLoop(model, dataloader, evaluator).run()
    infer_wedetect_ref.py:                proposals -> numpy -> list -> cuda  ->  proposals_for_ref(outputs, model.dtype)
    infer_wedetect.py:102-128:            inference_detector(model, image, texts, test_pipeline, ...)  ->  inference_detector(model, image, texts)
                                          (decode on the host, the resize / pad transforms and the detector on the device)
See INTEGRATION.md for the exact diffs.
"""
import torch

from .config import Config, parse_cfg_options  # noqa: F401
from .detector import SimpleYOLOWorldDetector, XLMRobertaLanguageBackbone, YOLOWorldDetector  # noqa: F401
from .retrieval import (RetrievalScorer, evaluate_retrieval_per_class, extract_corpus, predictions_from_scores,  # noqa: F401
                        save_corpus, score_saved)
from .loop import TestLoop, proposals_for_ref, sample_to_dict  # noqa: F401
from .structures import DetDataSample, InstanceData  # noqa: F401


def init_detector(config, checkpoint=None, palette="none", device="cuda:0", cfg_options=None, precise=True):
    """mmdet.apis.init_detector look-alike (infer_wedetect.py:156): config path or Config, checkpoint path."""
    cfg = Config.fromfile(config) if isinstance(config, str) else config
    if cfg_options:
        cfg.merge_from_dict(cfg_options)
    model_cfg = cfg["model"]
    if model_cfg["type"] != "YOLOWorldDetector":
        raise NotImplementedError(model_cfg["type"])
    model = YOLOWorldDetector(model_cfg, device=device, precise=precise)
    if checkpoint is not None:
        # mmengine checkpoints carry meta / message_hub objects besides tensors: a trusted local file, so weights_only=False
        sd = torch.load(checkpoint, map_location="cpu", weights_only=False) if isinstance(checkpoint, str) else checkpoint
        model.load_state_dict(sd)
    model.cfg = cfg
    return model.eval()


def inference_detector(model, image, texts, test_pipeline=None, max_dets=100, score_thr=0.3):
    """infer_wedetect.py:102-128 without the drawing: `image` (file name, decoded uint8 BGR array, or a list of either) goes
    through the config's test pipeline on the device (model.predict_images), then the script's tail: scores > score_thr, top
    max_dets by score, numpy.  Returns dict(xyxy, class_id, confidence) (a list of them for a list of images).  `test_pipeline`
    is accepted for signature compatibility and ignored: the transforms come from model.cfg."""
    many = isinstance(image, (list, tuple))
    outs = model.predict_images(list(image) if many else [image], texts=texts)
    res = []
    for o in outs:
        inst = o.pred_instances
        inst = inst[inst.scores.float() > score_thr]
        if len(inst.scores) > max_dets:
            inst = inst[inst.scores.float().topk(max_dets)[1]]
        inst = inst.cpu().numpy()
        res.append(dict(xyxy=inst["bboxes"], class_id=inst["labels"], confidence=inst["scores"]))
    return res if many else res[0]
