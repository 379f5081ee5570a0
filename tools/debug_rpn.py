import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-multi-resolution-rcnn_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from roi3d_b200 import RPNProposal3D
from roi3d_b200.models.anchor_heads import topk_segmented, decode_proposals
from roi3d_b200.ops import nms3d_batched
dev = torch.device("cuda:0")
B = 2
dims = [(8, 16, 16), (4, 8, 8), (2, 4, 4)]
strides, dstrides = [4, 8, 16], [2, 4, 8]
rng = np.random.default_rng(50)
cls = [(2 * rng.standard_normal((B, 1) + d)).astype(np.float32) for d in dims]
reg = [(0.1 * rng.standard_normal((B, 6) + d)).astype(np.float32) for d in dims]
head = RPNProposal3D(anchor_scales=[2], anchor_depth_scales=[2], anchor_ratios=[1.0], anchor_strides=strides, anchor_strides_depth=dstrides)
b = 0
for l in range(3):
    c = torch.from_numpy(cls[l][b]).to(dev)
    idx, val = topk_segmented([c], 300, apply_sigmoid=True, permute_adhw=True)
    flat = oracle.sigmoid(np.transpose(cls[l][b], (2, 3, 1, 0)).reshape(-1))
    want = oracle.topk(flat, 300)
    n = len(want)
    print("level", l, "topk idx equal:", np.array_equal(idx[0, :n].cpu().numpy(), want), "val maxdiff", np.abs(val[0, :n].cpu().numpy() - flat[want]).max())
    anchors = oracle.grid_anchors(oracle.gen_base_anchors(strides[l], [2], [2], [1.0], dstrides[l]), dims[l], strides[l], dstrides[l])
    deltas = np.transpose(reg[l][b], (2, 3, 1, 0)).reshape(-1, 6)
    wantp = oracle.delta2bbox3d(anchors[want], deltas[want], max_shape=(64, 64, 3, 16))
    got = decode_proposals(torch.from_numpy(reg[l][b]).to(dev), head.anchor_generators[l].base_anchors, strides[l], dstrides[l], idx[0], val[0], head.target_means, head.target_stds, (64, 64, 3, 16)).cpu().numpy()
    print("  decode maxdiff", np.abs(got[:n, :6] - wantp).max())
    d = np.concatenate([wantp, flat[want][:, None]], 1)
    keep_o, so = oracle.nms3d(d, 0.7, return_score_order=True)
    dg = torch.from_numpy(got).to(dev)[None].contiguous()
    keep, ks, num = nms3d_batched(dg, torch.tensor([n], dtype=torch.int32, device=dev), 0.7)
    m = int(num[0])
    print("  nms kept", m, len(keep_o), np.array_equal(ks[0, :m].cpu().numpy(), so))
cfg = dict(nms_pre=300, nms_post=100, max_num=150, nms_thr=0.7, min_bbox_size=0, nms_across_levels=False)
metas = [dict(img_shape=(64, 64, 3, 16), scale_factor=1.0)] * B
got = head.get_proposals([torch.from_numpy(c).to(dev) for c in cls], [torch.from_numpy(r).to(dev) for r in reg], metas, cfg)
anchors = [oracle.grid_anchors(oracle.gen_base_anchors(s, [2], [2], [1.0], ds), d, s, ds) for d, s, ds in zip(dims, strides, dstrides)]
want = oracle.get_bboxes_single([c[b] for c in cls], [r[b] for r in reg], anchors, (64, 64, 3, 16), 300, 100, 150, 0.7)
g = got[b].cpu().numpy()
bad = np.where(np.abs(g - want).max(1) > 1e-3)[0]
print("final mismatch rows", bad[:10], len(bad), "of", len(want))
if len(bad):
    i = bad[0]
    print(g[i], want[i])
    # is want[i] anywhere in g?
    dd = np.abs(g[:, None, :] - want[None, i:i+1, :]).max(2)
    print("want row found at", np.where(dd < 1e-3)[0])
# where does the extra row come from?
extra = g[bad[0]]
for l in range(3):
    flat = oracle.sigmoid(np.transpose(cls[l][b], (2, 3, 1, 0)).reshape(-1))
    want_i = oracle.topk(flat, 300)
    anchors_l = anchors[l]
    deltas = np.transpose(reg[l][b], (2, 3, 1, 0)).reshape(-1, 6)
    wp = oracle.delta2bbox3d(anchors_l[want_i], deltas[want_i], max_shape=(64, 64, 3, 16))
    d = np.concatenate([wp, flat[want_i][:, None]], 1)
    keep_o, so = oracle.nms3d(d, 0.7, return_score_order=True)
    hit = np.where(np.abs(d - extra[None]).max(1) < 1e-3)[0]
    print("level", l, "extra row is sorted position", hit, "kept positions rank:", [int(np.where(so == h)[0][0]) if h in so else -1 for h in hit], "n kept", len(so))
