# Minimal mmengine-style config for the facade tests: the `model=` surface of a WeDetect-Base detector, written against the
# registry names / keys the drop-in keeps (the shipped file is config/wedetect_base.py in the reference checkout,
# which loads unmodified through wedetect_b200.config.Config where it is mounted: tests/test_config_cpu.py).
num_classes = 80
text_channels = 768
model = dict(
    type='YOLOWorldDetector',
    mm_neck=False,
    num_train_classes=num_classes,
    num_test_classes=num_classes,
    data_preprocessor=dict(type='YOLOWDetDataPreprocessor', mean=[0., 0., 0.], std=[255., 255., 255.], bgr_to_rgb=True),
    backbone=dict(type='MultiModalYOLOBackbone',
                  image_model=dict(type='ConvNextVisionBackbone', model_name='base'),
                  text_model=dict(type='XLMRobertaLanguageBackbone', model_name='./xlm-roberta-base/', model_size='base')),
    neck=dict(type='CSPRepBiFPANNeck', scale_factor=1.0, model_size='base'),
    bbox_head=dict(type='YOLOWorldHead',
                   head_module=dict(type='YOLOWorldHeadModule', use_bn_head=True, embed_dims=text_channels, num_classes=num_classes, model_size='base'),
                   prior_generator=dict(type='MlvlPointGenerator', offset=0.5, strides=[8, 16, 32]),
                   bbox_coder=dict(type='WeDetectDistancePointBBoxCoder')),
    test_cfg=dict(multi_label=True, nms_pre=30000, score_thr=0.001, nms=dict(type='nms', iou_threshold=0.7), max_per_img=300))
img_scale = (640, 640)
test_pipeline = [
    dict(type='LoadImageFromFile', backend_args=None),
    dict(type='WeDetectKeepRatioResize', scale=img_scale),
    dict(type='WeDetectLetterResize', scale=img_scale, allow_scale_up=False, pad_val=dict(img=114)),
    dict(type='LoadAnnotations', with_bbox=True, _scope_='mmdet'),
    dict(type='LoadText'),
    dict(type='PackDetInputs', meta_keys=('img_id', 'img_path', 'ori_shape', 'img_shape', 'scale_factor', 'pad_param', 'texts')),
]
