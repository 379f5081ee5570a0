"""Harness for BASELINE.json configs[4] (C5): the RoI stage of configs/3d-multi-resolution-rcnn.py on synthetic
FPN pyramids with random-init heads.  The hot path (proposals, RoI extractors, NMS) is the product; the heads
(FC / conv3d, out of scope: cuBLAS / cuDNN) are plain torch layers with the reference's shapes
(SharedFCBBoxHead3D: convfc_bbox_head_3d.py:130-166; FCNMaskHead3D: fcn_mask_head_3d.py:89-97).

The config dicts below are the reference's own values (configs/3d-multi-resolution-rcnn.py:16-27,38-45,66-73,
132-143), copied as data so the harness also checks that they construct the drop-in modules unchanged."""
import torch
import torch.nn as nn

RPN_HEAD = dict(type='RPNHead3D', in_channels=64, feat_channels=64, anchor_scales=[2], anchor_depth_scales=[2],
                anchor_ratios=[1.0], anchor_strides=[4, 8, 16, 32, 64], anchor_strides_depth=[2, 4, 8, 16, 32],
                target_means=[.0, .0, .0, .0, .0, .0], target_stds=[1.0, 1.0, 1.0, 1.0, 1.0, 1.0],
                use_sigmoid_cls=True)
BBOX_ROI_EXTRACTOR = dict(type='SingleRoIExtractor',
                          roi_layer=dict(type='RoIAlign3D', out_size=7, out_size_depth=3, sample_num=2),
                          out_channels=64, featmap_strides=[4, 8, 16, 32], featmap_strides_depth=[2, 4, 8, 16])
MASK_ROI_EXTRACTOR = dict(type='SingleRoIExtractor',
                          roi_layer=dict(type='RoIAlign3D', out_size=14, out_size_depth=10, sample_num=2),
                          out_channels=64, featmap_strides=[4, 8, 16, 32], featmap_strides_depth=[2, 4, 8, 16])
TEST_CFG_RPN = dict(nms_across_levels=False, nms_pre=2000, nms_post=2000, max_num=2000, nms_thr=0.7, min_bbox_size=0)
TEST_CFG_RCNN = dict(score_thr=0.2, nms=dict(type='nms', iou_thr=0.5), max_per_img=2000, mask_thr_binary=0.25)
BBOX_TARGET_STDS = [0.1, 0.1, 0.2, 0.2, 0.1, 0.1]


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
        proposals = self.rpn.get_bboxes(cls_scores, bbox_preds, img_metas, TEST_CFG_RPN)
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
