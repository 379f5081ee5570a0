"""The RoIAlign3D layer object.

Interface contract (what callers of the reference rely on, mmdet/ops/roi_align/modules/roi_align_3d.py:5-18 and
mmdet/models/roi_extractors/single_level.py:54-55,88,92):
  * constructed by keyword from the config: ``RoIAlign3D(spatial_scale=1/s, spatial_scale_depth=1/d, out_size=..,
    out_size_depth=.., sample_num=..)``;
  * public attributes ``out_size``, ``out_size_depth``, ``spatial_scale``, ``spatial_scale_depth``, ``sample_num``;
  * ``layer(features[B,C,D,H,W], rois[K,7]) -> [K, C, out_size_depth, out_size, out_size]``, differentiable w.r.t.
    ``features``.
The layer has no parameters or buffers, so checkpoints are unaffected by swapping it in.
"""
import torch.nn as nn

from .functions.roi_align_3d import RoIAlignFunction3D, _out_dims


class RoIAlign3D(nn.Module):
    """Trilinear RoI pooling of a 3D feature map into ``out_size_depth x out_size x out_size`` bins."""

    def __init__(self, out_size, out_size_depth, spatial_scale, spatial_scale_depth, sample_num=0):
        super().__init__()
        _out_dims(out_size, out_size_depth)  # validates the types early instead of at the first forward
        if int(sample_num) < 0:
            raise ValueError("sample_num must be >= 0 (0 = adaptive: ceil(bin size) samples per axis)")
        self.out_size, self.out_size_depth = out_size, out_size_depth
        self.spatial_scale, self.spatial_scale_depth = float(spatial_scale), float(spatial_scale_depth)
        self.sample_num = int(sample_num)

    @property
    def output_shape(self):
        """(depth, height, width) of the pooled bins."""
        return _out_dims(self.out_size, self.out_size_depth)

    def forward(self, features, rois):
        args = (self.out_size, self.out_size_depth, self.spatial_scale, self.spatial_scale_depth, self.sample_num)
        return RoIAlignFunction3D.apply(features, rois, *args)

    def extra_repr(self):
        return "out=%sx%sx%s, spatial_scale=%g, spatial_scale_depth=%g, sample_num=%d" % (
            self.output_shape + (self.spatial_scale, self.spatial_scale_depth, self.sample_num))
