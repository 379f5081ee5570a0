"""Module path kept for `from mmdet.ops.roi_align.modules.roi_align_3d import RoIAlign3D`-style imports."""
