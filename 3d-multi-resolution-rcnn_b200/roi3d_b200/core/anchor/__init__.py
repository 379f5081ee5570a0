from .anchor_generator_3d import AnchorGenerator3D
from .anchor_target import anchor_inside_flags

__all__ = ['AnchorGenerator3D', 'anchor_inside_flags']
