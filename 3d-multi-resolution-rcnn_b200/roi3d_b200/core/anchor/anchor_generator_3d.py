"""AnchorGenerator3D (reference: mmdet/core/anchor/anchor_generator_3d.py:6-71).

Base anchors are the same closed form (rounded corner offsets around the cell centre).  grid_anchors keeps the
reference's enumeration order -- np.meshgrid(x, y, z) 'xy' indexing flattened, i.e. flat index
((y*W + x)*D + z)*A + a -- but builds the grid on the device from aranges instead of numpy + a 31 MB H2D copy
per call (anchor_head_3d.py:248-252).  The fused proposal path never materialises the grid at all: the decode
kernel recomputes the anchor of each selected index (csrc/proposal.cu).
"""
import torch


class AnchorGenerator3D(object):

    def __init__(self, base_size, scales, depth_scales, ratios, anchor_depth_base, scale_major=True, ctr=None):
        self.base_size = base_size
        self.anchor_depth_base = anchor_depth_base
        self.scales = torch.Tensor(scales)
        self.anchor_depth_scales = torch.Tensor(depth_scales)
        self.ratios = torch.Tensor(ratios)
        self.scale_major = scale_major
        self.ctr = ctr
        self.base_anchors = self.gen_base_anchors()

    @property
    def num_base_anchors(self):
        return self.base_anchors.size(0)

    def gen_base_anchors(self):
        """[A,6] base anchors (x1,y1,x2,y2,z1,z2) centred on cell 0: in-plane size base_size*scale with aspect
        sqrt(ratio), depth anchor_depth_base*depth_scale*sqrt(ratio); corners at centre -+ (size-1)/2, rounded
        half-to-even (reference anchor_generator_3d.py:22-53)."""
        side, depth = float(self.base_size), float(self.anchor_depth_base)
        centre = self.ctr if self.ctr is not None else tuple(0.5 * (v - 1) for v in (side, side, depth))
        rh = torch.sqrt(self.ratios)
        per_axis = ((side / rh, self.scales), (side * rh, self.scales), (depth * rh, self.anchor_depth_scales))
        sizes = []
        for by_ratio, by_scale in per_axis:
            grid = by_ratio[:, None] * by_scale[None, :] if self.scale_major else by_scale[:, None] * by_ratio[None, :]
            sizes.append(grid.reshape(-1))
        half = [0.5 * (s - 1) for s in sizes]
        cx, cy, cz = centre
        corners = [cx - half[0], cy - half[1], cx + half[0], cy + half[1], cz - half[2], cz + half[2]]
        return torch.stack(corners, dim=-1).round()

    def grid_anchors(self, featmap_size, stride=16, depth_stride=2, device='cuda'):
        """All anchors of one level, [D*H*W*A, 6], in the reference's order (y, x, z, a)."""
        base = self.base_anchors.to(device)
        feat_z, feat_h, feat_w = featmap_size
        sx = torch.arange(0, feat_w, device=device, dtype=torch.float32) * stride
        sy = torch.arange(0, feat_h, device=device, dtype=torch.float32) * stride
        sz = torch.arange(0, feat_z, device=device, dtype=torch.float32) * depth_stride
        yy, xx, zz = torch.meshgrid(sy, sx, sz, indexing='ij')
        shifts = torch.stack([xx, yy, xx, yy, zz, zz], dim=-1).reshape(-1, 6)
        return (base[None, :, :] + shifts[:, None, :]).view(-1, 6)
