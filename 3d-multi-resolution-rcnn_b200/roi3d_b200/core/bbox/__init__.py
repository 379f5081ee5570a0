from .transforms import bbox2roi3D, delta2bbox3D

__all__ = ['bbox2roi3D', 'delta2bbox3D']
