"""AnchorGenerator3D (reference: mmdet/core/anchor/anchor_generator_3d.py:6-71).

Base anchors are the same closed form (rounded corner offsets around the cell centre).  grid_anchors keeps the
reference's enumeration order -- np.meshgrid(x, y, z) 'xy' indexing flattened, i.e. flat index
((y*W + x)*D + z)*A + a -- but writes the grid on the device with one kernel (`roi3d_grid_anchors`, which can also emit
the valid / inside flags of the training path) instead of numpy + a 31 MB H2D copy per call
(anchor_head_3d.py:248-252).  The fused proposal path never materialises the grid at all: the decode
kernel recomputes the anchor of each selected index (csrc/proposal.cu).
"""
import torch

from ... import _lib
from ..._util import stream_ptr


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

    def _launch(self, featmap_size, stride, depth_stride, device, valid_size, img_shape, allowed_border, want_anchors,
                want_flags):
        feat_z, feat_h, feat_w = (int(v) for v in featmap_size)
        device = torch.device(device)
        if device.type != 'cuda':
            raise NotImplementedError("AnchorGenerator3D builds anchors on the GPU: the B200 path has no CPU "
                                      "implementation")
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        A = self.num_base_anchors
        n = feat_z * feat_h * feat_w * A
        base = self.base_anchors.to(dtype=torch.float32, device='cpu').contiguous()
        anchors = torch.empty((n, 6), dtype=torch.float32, device=device) if want_anchors else None
        flags = torch.empty((n,), dtype=torch.uint8, device=device) if want_flags else None
        vd, vh, vw = (int(v) for v in valid_size) if valid_size is not None else (feat_z, feat_h, feat_w)
        assert vh <= feat_h and vw <= feat_w and vd <= feat_z
        ih, iw, idp = (float(img_shape[0]), float(img_shape[1]), float(img_shape[3])) if img_shape is not None \
            else (0.0, 0.0, 0.0)
        if n:
            with torch.cuda.device(device):
                _lib.check(_lib.lib.roi3d_grid_anchors(
                    A, feat_z, feat_h, feat_w, float(stride), float(depth_stride), base.data_ptr(), vd, vh, vw, ih, iw,
                    idp, int(allowed_border), None if anchors is None else anchors.data_ptr(),
                    None if flags is None else flags.data_ptr(), stream_ptr()))
        return anchors, flags

    def grid_anchors(self, featmap_size, stride=16, depth_stride=2, device='cuda'):
        """All anchors of one level, [D*H*W*A, 6], in the reference's order (y, x, z, a); one kernel, no numpy grid,
        no H2D copy (reference: anchor_generator_3d.py:56-71)."""
        return self._launch(featmap_size, stride, depth_stride, device, None, None, -1, True, False)[0]

    def valid_flags(self, featmap_size, valid_size, device='cuda'):
        """uint8 [D*H*W*A]: 1 where the cell lies inside the valid (unpadded) extent (reference: :73-92)."""
        return self._launch(featmap_size, 1, 1, device, valid_size, None, -1, False, True)[1]

    def grid_anchors_and_inside_flags(self, featmap_size, stride, depth_stride, valid_size, img_shape, allowed_border=0,
                                      device='cuda'):
        """Anchors plus `anchor_inside_flags(anchors, valid_flags, img_shape, allowed_border)`
        (mmdet/core/anchor/anchor_target.py:203-217) in the same launch.  img_shape = (H, W, 3, D)."""
        return self._launch(featmap_size, stride, depth_stride, device, valid_size, img_shape, allowed_border, True, True)
