"""The reference's config file drives the drop-in modules unchanged.

CPU test: where the reference tree is present (the build container) the REAL configs/3d-multi-resolution-rcnn.py is
executed with roi3d_b200.load_config and (a) must equal the committed fixture the GPU tests use, (b) must build the
hot-path modules.  On a box without /root/reference only the fixture half runs."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CFG_PATH = os.path.join(os.environ.get("ROI3D_REFERENCE", "/root/reference"), "configs", "3d-multi-resolution-rcnn.py")


def _fixture():
    with open(os.path.join(HERE, "golden", "reference_config_hot_path.json")) as f:
        d = json.load(f)
    d.pop("_source")
    return d


def _check_modules(rpn, bbox_ex, mask_ex):
    assert bbox_ex.num_inputs == 4 and len(bbox_ex.roi_layers) == 4
    assert [l.out_size for l in bbox_ex.roi_layers] == [7] * 4 and [l.out_size_depth for l in bbox_ex.roi_layers] == [3] * 4
    assert [l.spatial_scale for l in bbox_ex.roi_layers] == [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    assert [l.spatial_scale_depth for l in bbox_ex.roi_layers] == [1 / 2, 1 / 4, 1 / 8, 1 / 16]
    assert all(l.sample_num == 2 for l in bbox_ex.roi_layers)
    assert [l.out_size for l in mask_ex.roi_layers] == [14] * 4 and [l.out_size_depth for l in mask_ex.roi_layers] == [10] * 4
    assert bbox_ex.out_channels == 64 and mask_ex.out_channels == 64 and bbox_ex.finest_scale == 56
    assert rpn.anchor_strides == [4, 8, 16, 32, 64] and rpn.anchor_strides_depth == [2, 4, 8, 16, 32]
    assert len(rpn.anchor_generators) == 5 and rpn.num_anchors == 1


def test_fixture_builds_the_hot_path_modules():
    import roi3d_b200
    rpn, bbox_ex, mask_ex = roi3d_b200.build_from_config(_fixture())
    _check_modules(rpn, bbox_ex, mask_ex)


@pytest.mark.skipif(not os.path.exists(REF_CFG_PATH), reason="reference tree not present on this box")
def test_real_reference_config_loads_unchanged_and_matches_the_fixture():
    import roi3d_b200
    from roi3d_b200.models.config import hot_path_config
    cfg = roi3d_b200.load_config(REF_CFG_PATH)
    assert cfg['model']['type'] == 'MaskRCNN3D2Scales'           # the whole file executed, not just our keys
    assert json.loads(json.dumps(hot_path_config(cfg))) == _fixture()
    rpn, bbox_ex, mask_ex = roi3d_b200.build_from_config(cfg)    # straight from the reference's dicts
    _check_modules(rpn, bbox_ex, mask_ex)
    # the proposal path takes the reference's test_cfg.rpn / train_cfg.rpn_proposal objects as they are
    from roi3d_b200.models.anchor_heads.rpn_head_3d import _cfg_get
    for sub in (cfg['test_cfg']['rpn'], cfg['train_cfg']['rpn_proposal']):
        assert _cfg_get(sub, 'nms_pre') == 2000 and _cfg_get(sub, 'nms_thr') == 0.7
