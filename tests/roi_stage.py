"""Harness for BASELINE.json configs[4] (C5): the RoI stage of configs/3d-multi-resolution-rcnn.py on synthetic
FPN pyramids with random-init heads.  The hot path (proposals, RoI extractors, NMS) is the product; the heads
(FC / conv3d, out of scope: cuBLAS / cuDNN) are plain torch layers with the reference's shapes
(SharedFCBBoxHead3D: convfc_bbox_head_3d.py:130-166; FCNMaskHead3D: fcn_mask_head_3d.py:89-97).

The config dicts come from tests/golden/reference_config_hot_path.json, which tests/golden/make_config_golden.py
extracts from the reference's own configs/3d-multi-resolution-rcnn.py (:16-27, :38-45, :66-73, :132-143) with
roi3d_b200.models.config.load_config; tests/test_config.py re-executes the real file (where /root/reference exists)
and checks the fixture against it, so the harness builds the drop-in modules from the reference's unchanged values."""
import json
import os

import torch
import torch.nn as nn

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_config_hot_path.json")) as _f:
    REF_CFG = json.load(_f)
RPN_HEAD = REF_CFG['rpn_head']
BBOX_ROI_EXTRACTOR = REF_CFG['bbox_roi_extractor']
MASK_ROI_EXTRACTOR = REF_CFG['mask_roi_extractor']
TEST_CFG_RPN = REF_CFG['test_cfg']['rpn']
TEST_CFG_RCNN = REF_CFG['test_cfg']['rcnn']
BBOX_TARGET_STDS = REF_CFG['bbox_head_target_stds']


class RoIStage(nn.Module):
    def __init__(self, channels=64, max_masks=100, seed=7):
        super().__init__()
        import roi3d_b200
        torch.manual_seed(seed)
        self.rpn = roi3d_b200.build_rpn_proposal(RPN_HEAD)
        self.bbox_ex = roi3d_b200.build_roi_extractor(BBOX_ROI_EXTRACTOR)
        self.mask_ex = roi3d_b200.build_roi_extractor(MASK_ROI_EXTRACTOR)
        self.fc = nn.Sequential(nn.Linear(channels * 3 * 7 * 7, 1024), nn.ReLU(), nn.Linear(1024, 1024), nn.ReLU())
        self.fc_cls, self.fc_reg = nn.Linear(1024, 2), nn.Linear(1024, 12)
        convs = []
        for _ in range(4):
            convs += [nn.Conv3d(channels, channels, 3, padding=1), nn.ReLU()]
        self.mask_head = nn.Sequential(*convs, nn.ConvTranspose3d(channels, channels, 2, stride=2), nn.ReLU(),
                                       nn.Conv3d(channels, 2, 1))
        self.max_masks = max_masks

    @torch.no_grad()
    def forward(self, feats, cls_scores, bbox_preds, img_metas):
        """feats: 4+ FPN levels [B,C,D,H,W]; returns per-image (det_bboxes [k,7], det_labels, mask logits)."""
        import roi3d_b200
        proposals = self.rpn.get_proposals(cls_scores, bbox_preds, img_metas, TEST_CFG_RPN)
        rois = roi3d_b200.bbox2roi3D(proposals)
        x = self.bbox_ex(feats[:4], rois)
        h = self.fc(x.flatten(1))
        scores = self.fc_cls(h).softmax(dim=1)
        deltas = self.fc_reg(h)
        out, start = [], 0
        for b, props in enumerate(proposals):
            n = props.shape[0]
            boxes = roi3d_b200.delta2bbox3D(rois[start:start + n, 1:], deltas[start:start + n], [0.0] * 6,
                                            BBOX_TARGET_STDS, img_metas[b]['img_shape'])
            det, lab = roi3d_b200.multiclass_nms_3d(boxes, scores[start:start + n], TEST_CFG_RCNN['score_thr'],
                                                    TEST_CFG_RCNN['nms'], TEST_CFG_RCNN['max_per_img'])
            start += n
            det_m = det[:self.max_masks]
            mrois = torch.cat([det_m.new_full((det_m.shape[0], 1), b), det_m[:, :6]], dim=1)
            masks = self.mask_head(self.mask_ex(feats[:4], mrois)) if det_m.shape[0] else det_m.new_zeros((0,))
            out.append((det, lab, masks))
        return out


def synthetic_inputs(batch, vol_dhw=(160, 512, 512), channels=64, device='cuda', seed=7, channels_last=True):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    D, H, W = vol_dhw
    strides, dstrides = [4, 8, 16, 32, 64], [2, 4, 8, 16, 32]
    feats, cls, reg = [], [], []
    for s, ds in zip(strides, dstrides):
        dims = (max(D // ds, 1), max(H // s, 1), max(W // s, 1))
        f = torch.randn((batch, channels) + dims, device=device, generator=g)
        feats.append(f.contiguous(memory_format=torch.channels_last_3d) if channels_last else f)
        cls.append(2 * torch.randn((batch, 1) + dims, device=device, generator=g))
        reg.append(0.1 * torch.randn((batch, 6) + dims, device=device, generator=g))
    metas = [dict(img_shape=(H, W, 3, D), scale_factor=1.0)] * batch
    return feats, cls, reg, metas
