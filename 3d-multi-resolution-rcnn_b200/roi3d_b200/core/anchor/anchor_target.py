"""anchor_inside_flags for already materialised anchors (reference: mmdet/core/anchor/anchor_target.py:203-217).
The fused form that never reads the anchors back is AnchorGenerator3D.grid_anchors_and_inside_flags."""


def anchor_inside_flags(flat_anchors, valid_flags, img_shape, allowed_border=0):
    """flat_anchors [N,6] (x1,y1,x2,y2,z1,z2), valid_flags uint8/bool [N], img_shape = (H, W, 3, D)."""
    if flat_anchors.shape[1] != 6:
        raise NotImplementedError("anchor_inside_flags: only 3D anchors (6 columns) are on this path")
    if allowed_border < 0:
        return valid_flags
    img_h, img_w, img_d = img_shape[0], img_shape[1], img_shape[3]
    a = flat_anchors
    inside = (a[:, 0] >= -allowed_border) & (a[:, 1] >= -allowed_border) & (a[:, 4] >= -allowed_border) & \
        (a[:, 2] < img_w + allowed_border) & (a[:, 3] < img_h + allowed_border) & (a[:, 5] < img_d + allowed_border)
    return valid_flags & inside.to(valid_flags.dtype)
