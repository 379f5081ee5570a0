from .mask_target import mask_target, mask_target_single

__all__ = ['mask_target', 'mask_target_single']
