"""Small torch-side helpers shared by the host mirror: stream handle, layout handling, scratch buffers."""
import os

import torch

from . import _lib


# --- NVTX ranges (the reference has no tracing; the build adds named ranges around the hot-path calls)
NVTX = os.environ.get("ROI3D_NVTX", "0") not in ("", "0")


class device_guard(object):
    """`with device_guard(dev):` -- torch.cuda.device(dev), but free when `dev` already is the current device (the
    common case: one process per GPU)."""

    def __init__(self, dev):
        idx = dev.index if isinstance(dev, torch.device) else int(dev)
        self.ctx = None if idx is None or idx == torch.cuda.current_device() else torch.cuda.device(idx)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False


class nvtx_range(object):
    """`with nvtx_range("roi3d.nms"):` -- a named range in Nsight timelines when ROI3D_NVTX=1, free otherwise."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if NVTX:
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def stream_ptr():
    """cudaStream_t of torch's current stream, as an int for ctypes."""
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def is_channels_last_3d(x):
    """True when a [B,C,D,H,W] tensor is stored NDHWC (torch.channels_last_3d) densely."""
    return x.dim() == 5 and x.is_contiguous(memory_format=torch.channels_last_3d)


# --- NCDHW -> channels-last conversion -----------------------------------------------------------
# The reference API takes NCDHW-contiguous feature maps (roi_align_cuda.cpp:35-39); the streamed / ring kernels read
# channels-last.  A detector calls the extractor several times per step on the SAME FPN outputs (bbox, refinement and
# mask extractors: two_stage_3d_2scales.py:234-237,290-291,305-306), so a converted copy can be reused -- but only
# when the caller says the buffer has not been rewritten: a tensor's `_version` does not see writes through
# `data_ptr` (a backbone replayed as a CUDA graph, a custom kernel), so reuse is OPT-IN and scoped:
#
#     with roi3d_b200.reuse_layout_conversions():      # e.g. around the RoI stage of one forward pass
#         bbox_feats = bbox_extractor(feats, rois)
#         mask_feats = mask_extractor(feats, mask_rois)
#
# Outside such a scope every call converts afresh (the conversion runs at HBM speed: 242 us for the 671 MB C2 level).
# Inside, entries are keyed by (data_ptr, shape, stride, stream) and dropped when the scope ends, so nothing outlives
# the pass or pins memory, and a copy made on one stream is never handed to another.
_scopes = []


class reuse_layout_conversions(object):
    """Context manager: NCDHW feature maps converted inside the scope are converted once (see above).  The caller
    promises not to rewrite those feature maps while the scope is open."""

    def __enter__(self):
        _scopes.append({})
        return self

    def __exit__(self, *exc):
        _scopes.pop().clear()
        return False


def clear_layout_cache():
    for sc in _scopes:
        sc.clear()


def to_channels_last_3d(x):
    """Return (tensor stored NDHWC with the same logical shape, was_converted)."""
    if is_channels_last_3d(x):
        return x, False
    if not x.is_contiguous():
        x = x.contiguous()
    cache = _scopes[-1] if _scopes else None
    key = None
    if cache is not None:
        key = (x.data_ptr(), tuple(x.shape), x.device.index, torch.cuda.current_stream(x.device).cuda_stream)
        hit = cache.get(key)
        if hit is not None:
            return hit[1], True
    B, C, D, H, W = x.shape
    conv = torch.empty_like(x, memory_format=torch.channels_last_3d)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib.roi3d_ncdhw_to_ndhwc(x.data_ptr(), conv.data_ptr(), B, C, D, H, W, stream_ptr()))
    if cache is not None:
        cache[key] = (x, conv)  # holding x keeps its address from being reused inside the scope
    return conv, True


FORCE_NATIVE_NCDHW = [False]   # developer switch: always hand NCDHW levels to the native kernels when they can read them


def forward_inputs(feats, out_h, out_w):
    """Pick the memory the forward kernels read for a list of [B,C,D,H,W] levels: (tensors, layout flag).
    Channels-last levels go in as they are.  NCDHW-contiguous levels -- what the reference's callers hold
    (roi_align_cuda.cpp:35-39) -- also go in as they are when a native kernel can read them (square 7- or 14-wide
    output, rows that start on 16-byte boundaries: the streamed kernel's NCDHW twin for 7-wide outputs on multiples of 64
    channels, the planar kernel otherwise); anything else is converted to channels-last first."""
    if all(is_channels_last_3d(f) for f in feats):
        return list(feats), _lib.NDHWC
    native = out_h == out_w and out_h in (7, 14) and all(
        f.is_contiguous() and f.shape[-1] % 4 == 0 and f.data_ptr() % 16 == 0 for f in feats)
    # Inside a reuse_layout_conversions() scope a 7-wide extractor on 64-channel multiples converts once and runs the
    # streamed channels-last kernel: the converted copy serves every extractor call of the pass (C2: 137 us per call
    # after a 240 us conversion).  A lone call reads the NCDHW tensor in place with the streamed kernel's NCDHW twin
    # (planar kernel 357 us, conversion + streamed kernel 377 us).
    if native and out_h == 7 and feats[0].shape[1] % 64 == 0 and _scopes and not FORCE_NATIVE_NCDHW[0]:
        native = False
    if native:
        return list(feats), _lib.NCDHW
    return [to_channels_last_3d(f)[0] for f in feats], _lib.NDHWC


def channels_last_to_contiguous(g):
    """[B,C,D,H,W] stored NDHWC -> NCDHW-contiguous copy (used for grad_input when the forward input was NCDHW)."""
    B, C, D, H, W = g.shape
    out = torch.empty((B, C, D, H, W), dtype=g.dtype, device=g.device)
    with torch.cuda.device(g.device):
        _lib.check(_lib.lib.roi3d_ndhwc_to_ncdhw(g.data_ptr(), out.data_ptr(), B, C, D, H, W, stream_ptr()))
    return out


# --- kernel workspaces ----------------------------------------------------------------------------
def workspace(device, nbytes):
    """A fresh 256-byte aligned byte buffer for ONE call (NMS / top-k / assigner workspaces).  torch's caching
    allocator makes this cheap, orders reuse by stream, and -- under CUDA-graph capture -- serves it from the graph's
    own pool, so two streams never share a workspace and a replayed graph never writes into memory that was handed
    to someone else.  Returns (tensor to keep alive until the call is enqueued, aligned address)."""
    buf = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
    return buf, buf.data_ptr() + ((-buf.data_ptr()) % 256)


def check_cuda_f32(t, name, ndim=None, last=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor, got %s" % (name, type(t)))
    if not t.is_cuda:
        raise NotImplementedError("%s must be a CUDA tensor: the B200 path has no CPU implementation" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32 (got %s); the B200 path computes in fp32" % (name, t.dtype))
    if ndim is not None and t.dim() != ndim:
        raise ValueError("%s must have %d dims, got shape %s" % (name, ndim, tuple(t.shape)))
    if last is not None and t.shape[-1] != last:
        raise ValueError("%s must have last dim %d, got shape %s" % (name, last, tuple(t.shape)))
