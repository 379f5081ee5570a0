"""2+-GPU check of the only collective on the path: per-volume detections all-gathered over NCCL.
torchrun --nproc-per-node N tools/dist_gather_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-multi-resolution-rcnn_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import synth  # noqa: E402
from roi3d_b200 import multiclass_nms_3d  # noqa: E402
from roi3d_b200.parallel import gather_detections, shard_indices  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
n_vol = 2 * world + 1
mine = shard_indices(n_vol)
dets, labels = [], []
for v in mine:  # every rank runs the final-detection NMS of its own volumes
    d = synth.c1_boxes(300 + 10 * v, seed=v)
    boxes = torch.from_numpy(d[:, :6]).to(dev)
    sc = torch.from_numpy(np.stack([1 - d[:, 6], d[:, 6]], 1)).to(dev)
    b, l = multiclass_nms_3d(boxes, sc, 0.2, dict(type='nms', iou_thr=0.5), 2000)
    dets.append(b)
    labels.append(l)
out = gather_detections(dets, labels, mine)
assert sorted(out.keys()) == list(range(n_vol)), sorted(out.keys())
# every rank recomputes volume 0 and compares with what it received
d = synth.c1_boxes(300, seed=0)
b0, l0 = multiclass_nms_3d(torch.from_numpy(d[:, :6]).to(dev), torch.from_numpy(np.stack([1 - d[:, 6], d[:, 6]], 1)).to(dev),
                           0.2, dict(type='nms', iou_thr=0.5), 2000)
assert torch.equal(out[0][0].to(dev), b0) and torch.equal(out[0][1].to(dev), l0)
dist.barrier()
if rank == 0:
    print("gather_detections over NCCL ok: %d volumes on %d ranks, %s detections" % (n_vol, world, [int(out[i][0].shape[0]) for i in range(n_vol)]))
dist.destroy_process_group()
