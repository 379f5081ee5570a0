"""`type`-string construction of the hot-path modules from the reference's config dicts
(the role of mmdet/models/builder.py:8-56 + registry.py for the two module kinds this package provides)."""
from .anchor_heads import RPNProposal3D
from .roi_extractors import SingleRoIExtractor

_ROI_EXTRACTORS = {'SingleRoIExtractor': SingleRoIExtractor}
_RPN_KEYS = ('anchor_scales', 'anchor_depth_scales', 'anchor_ratios', 'anchor_strides', 'anchor_strides_depth',
             'anchor_base_sizes', 'anchor_base_depths', 'target_means', 'target_stds', 'use_sigmoid_cls')


def build_roi_extractor(cfg):
    """cfg e.g. configs/3d-multi-resolution-rcnn.py:38-45 (`bbox_roi_extractor`) or :66-73 (`mask_roi_extractor`)."""
    args = dict(cfg)
    kind = args.pop('type')
    if kind not in _ROI_EXTRACTORS:
        raise KeyError("roi extractor type %r is not part of the 3D RoI hot path" % (kind,))
    return _ROI_EXTRACTORS[kind](**args)


def build_rpn_proposal(cfg):
    """cfg e.g. configs/3d-multi-resolution-rcnn.py:16-27 (`rpn_head`, type 'RPNHead3D'); the convolution keys
    (in_channels, feat_channels) belong to the part of the head that stays in the reference and are ignored here."""
    args = dict(cfg)
    kind = args.pop('type', 'RPNHead3D')
    if kind != 'RPNHead3D':
        raise KeyError("anchor head type %r is not part of the 3D RoI hot path" % (kind,))
    return RPNProposal3D(**{k: v for k, v in args.items() if k in _RPN_KEYS})
