from .functions.roi_align_3d import RoIAlignFunction3D, roi_align_3d, set_bug_compat
from .host import roi_align_3d_host
from .modules.roi_align_3d import RoIAlign3D

__all__ = ['roi_align_3d', 'roi_align_3d_host', 'RoIAlign3D', 'RoIAlignFunction3D', 'set_bug_compat']
