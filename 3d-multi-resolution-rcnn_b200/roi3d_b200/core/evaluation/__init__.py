from .coco_utils import apply_nms, nms_3d_eval_batched, nms_3d_python

__all__ = ['apply_nms', 'nms_3d_eval_batched', 'nms_3d_python']
