"""Import-path shim: the reference exposes the layer as mmdet.ops.roi_align.modules.roi_align_3d.RoIAlign3D."""
from ..layer import RoIAlign3D

__all__ = ['RoIAlign3D']
