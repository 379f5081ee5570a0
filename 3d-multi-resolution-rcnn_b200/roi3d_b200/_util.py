"""Small torch-side helpers shared by the host mirror: stream handle, layout handling, scratch buffers."""
import collections
import os
import weakref

import torch

from . import _lib


def stream_ptr():
    """cudaStream_t of torch's current stream, as an int for ctypes."""
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def is_channels_last_3d(x):
    """True when a [B,C,D,H,W] tensor is stored NDHWC (torch.channels_last_3d) densely."""
    return x.dim() == 5 and x.is_contiguous(memory_format=torch.channels_last_3d)


# --- NCDHW -> channels-last conversion cache ----------------------------------------------------
# The reference API takes NCDHW-contiguous feature maps (roi_align_cuda.cpp:35-39); the kernels read
# channels-last.  A detector calls the extractor several times per step on the SAME FPN outputs
# (bbox, refinement and mask extractors: two_stage_3d_2scales.py:234-237,290-291,305-306), so the
# converted copy is cached per source tensor (identity + version counter), bounded LRU.
_CACHE_SIZE = int(os.environ.get("ROI3D_LAYOUT_CACHE", "8"))
_cache = collections.OrderedDict()


def clear_layout_cache():
    _cache.clear()


def to_channels_last_3d(x):
    """Return (tensor stored NDHWC with the same logical shape, was_converted)."""
    if is_channels_last_3d(x):
        return x, False
    if not x.is_contiguous():
        x = x.contiguous()
    key = (x.data_ptr(), tuple(x.shape), x.device.index)
    hit = _cache.get(key)
    if hit is not None:
        src_ref, version, conv = hit
        if src_ref() is x and version == x._version:
            _cache.move_to_end(key)
            return conv, True
        del _cache[key]
    B, C, D, H, W = x.shape
    conv = torch.empty_like(x, memory_format=torch.channels_last_3d)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib.roi3d_ncdhw_to_ndhwc(x.data_ptr(), conv.data_ptr(), B, C, D, H, W, stream_ptr()))
    if _CACHE_SIZE > 0:
        _cache[key] = (weakref.ref(x), x._version, conv)
        while len(_cache) > _CACHE_SIZE:
            _cache.popitem(last=False)
    return conv, True


def channels_last_to_contiguous(g):
    """[B,C,D,H,W] stored NDHWC -> NCDHW-contiguous copy (used for grad_input when the forward input was NCDHW)."""
    B, C, D, H, W = g.shape
    out = torch.empty((B, C, D, H, W), dtype=g.dtype, device=g.device)
    with torch.cuda.device(g.device):
        _lib.check(_lib.lib.roi3d_ndhwc_to_ncdhw(g.data_ptr(), out.data_ptr(), B, C, D, H, W, stream_ptr()))
    return out


# --- grow-only per-device byte scratch (NMS / top-k workspaces) ---------------------------------
_scratch = {}


def scratch(device, nbytes, tag="ws"):
    key = (device.index, tag)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
        _scratch[key] = buf
    off = (-buf.data_ptr()) % 256
    return buf, buf.data_ptr() + off


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
