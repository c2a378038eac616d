"""The mmengine-style config surface the drop-in keeps: `_base_` inheritance, `_delete_`, attribute access,
`--cfg-options` overrides (infer_wedetect.py:88-97,149-151), and - where the reference checkout is mounted - its own three
config files read unmodified."""
import os
import textwrap

import pytest

REF = "/root/reference"


def _write(tmp_path, name, body):
    p = tmp_path / name
    p.write_text(textwrap.dedent(body))
    return str(p)


def test_config_inheritance_delete_and_overrides(tmp_path):
    from wedetect_b200.config import Config, parse_cfg_options
    _write(tmp_path, "runtime.py", """
        default_scope = 'mmdet'
        env_cfg = dict(cudnn_benchmark=False, dist_cfg=dict(backend='nccl'))
        """)
    _write(tmp_path, "base.py", """
        _base_ = ['runtime.py']
        num_classes = 80
        model = dict(type='YOLOWorldDetector', mm_neck=False, num_test_classes=num_classes,
                     backbone=dict(type='MultiModalYOLOBackbone', image_model=dict(type='ConvNextVisionBackbone', model_name='base'),
                                   text_model=dict(type='XLMRobertaLanguageBackbone', model_name='./xlm-roberta-base/')),
                     neck=dict(type='CSPRepBiFPANNeck', scale_factor=1.0),
                     test_cfg=dict(multi_label=True, nms_pre=30000, score_thr=0.001, nms=dict(type='nms', iou_threshold=0.7), max_per_img=300))
        """)
    child = _write(tmp_path, "child.py", """
        _base_ = 'base.py'
        model = dict(neck=dict(_delete_=True, type='OtherNeck', width=3), test_cfg=dict(max_per_img=100))
        """)
    cfg = Config.fromfile(child)
    assert cfg.default_scope == "mmdet" and cfg["env_cfg"]["dist_cfg"]["backend"] == "nccl"        # two levels of _base_
    assert cfg.model.type == "YOLOWorldDetector" and cfg.model.num_test_classes == 80
    assert cfg.model.neck == dict(type="OtherNeck", width=3)                                          # _delete_ replaces, does not merge
    assert cfg.model.test_cfg.max_per_img == 100 and cfg.model.test_cfg.nms.iou_threshold == 0.7    # dicts merge key by key
    opts = parse_cfg_options(["model.test_cfg.score_thr=0.05", "model.test_cfg.nms.iou_threshold=0.6", "work_dir=out", "a.b=[1,2]"])
    assert opts == {"model.test_cfg.score_thr": 0.05, "model.test_cfg.nms.iou_threshold": 0.6, "work_dir": "out", "a.b": [1, 2]}
    cfg.merge_from_dict(opts)
    assert cfg.model.test_cfg.score_thr == 0.05 and cfg.model.test_cfg.nms.iou_threshold == 0.6 and cfg.work_dir == "out" and cfg.a.b == [1, 2]
    with pytest.raises(AttributeError):
        cfg.model.no_such_key


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not mounted")
@pytest.mark.parametrize("size,text,scale", [("tiny", "./xlm-roberta-base/", None), ("base", "./xlm-roberta-base/", 1.0), ("large", "./xlm-roberta-large/", 1.5)])
def test_reference_config_files_load_unmodified(size, text, scale):
    from wedetect_b200 import schema
    from wedetect_b200.config import Config
    cfg = Config.fromfile(os.path.join(REF, "config", f"wedetect_{size}.py"))
    m = cfg.model
    assert m.type == "YOLOWorldDetector" and m.backbone.image_model.model_name == size and size in schema.SIZES
    assert m.backbone.text_model.model_name == text and m.neck.get("scale_factor") == scale
    assert m.test_cfg == dict(multi_label=True, nms_pre=30000, score_thr=0.001, nms=dict(type="nms", iou_threshold=0.7), max_per_img=300)
    assert cfg.custom_imports["imports"] == ["wedetect"]


def test_init_detector_needs_the_gpu_library_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from wedetect_b200 import _lib
    from wedetect_b200.api import init_detector
    cfg = _write(tmp_path, "c.py", """
        model = dict(type='YOLOWorldDetector', backbone=dict(image_model=dict(model_name='tiny'), text_model=dict(model_name='x')), test_cfg=dict(
            multi_label=True, nms_pre=30000, score_thr=0.001, nms=dict(type='nms', iou_threshold=0.7), max_per_img=300))
        """)
    with pytest.raises(_lib.WdError):          # no CPU fallback behind the mmdet-shaped entry point either
        init_detector(cfg, checkpoint=None, device="cuda:0")
    bad = _write(tmp_path, "d.py", "model = dict(type='SomethingElse')\n")
    with pytest.raises(NotImplementedError):
        init_detector(bad)


def test_text_state_dict_slicing():
    """Checkpoint handling of the standalone text tower facade (extract_embedding.py:1293-1303): mmengine checkpoint, full
    state dict and the already sliced `model.* / head.*` layout give the same `backbone.text_model.*` slice."""
    import torch
    from oracle import synth
    from wedetect_b200 import schema
    sd = synth.synth_state_dict("tiny", seed=0, with_text=True, text_vocab=64, calibrate=False)
    a = schema.text_state_dict({"state_dict": sd, "meta": {}})
    b = schema.text_state_dict(sd)
    c = schema.text_state_dict({k[len("backbone.text_model."):]: v for k, v in sd.items() if k.startswith("backbone.text_model.")})
    assert a.keys() == b.keys() == c.keys() and all(k.startswith("backbone.text_model.") for k in a)
    assert all(torch.equal(a[k], c[k]) for k in a) and len(a) > 100
    assert schema.text_size_of(a) == "base"
    with pytest.raises(KeyError):
        schema.text_state_dict({"neck.x": torch.zeros(1)})
