"""Autograd function of RoIAlign3D over the C ABI (see roi_align_3d.py)."""
