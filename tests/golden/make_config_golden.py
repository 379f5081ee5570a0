"""Writes tests/golden/reference_config_hot_path.json from the reference's own config file (run in the build
container, where /root/reference exists; the GPU box only sees the JSON)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "3d-multi-resolution-rcnn_b200"))
from roi3d_b200.models.config import hot_path_config, load_config  # noqa: E402

REF = os.environ.get("ROI3D_REFERENCE", "/root/reference")
cfg = load_config(os.path.join(REF, "configs", "3d-multi-resolution-rcnn.py"))
out = hot_path_config(cfg)
out["_source"] = "configs/3d-multi-resolution-rcnn.py of arthur801031/3d-multi-resolution-rcnn (hot-path keys only)"
with open(os.path.join(HERE, "reference_config_hot_path.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print("wrote", os.path.join(HERE, "reference_config_hot_path.json"))
