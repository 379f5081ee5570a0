from .nms_wrapper import nms, nms3d_batched, soft_nms

__all__ = ['nms', 'soft_nms', 'nms3d_batched']
