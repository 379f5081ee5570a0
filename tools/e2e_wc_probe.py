"""e2e leg of bench.py with the host feature buffer in (a) torch pinned memory, (b) write-combined pinned memory
(cudaHostAlloc WriteCombined): does the host side of concurrent uploads limit N > 1?  Run under torchrun."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-multi-resolution-rcnn_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import synth
from roi3d_b200 import _lib
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", torch.cuda.current_device())
if world > 1:
    dist.init_process_group("nccl")
shape = (1, 256, 40, 128, 128)
g = torch.Generator().manual_seed(1)
rois_np = synth.c2_rois(512, seed=2)
rois_h = torch.from_numpy(rois_np).pin_memory()
out_h = torch.empty((512, 256, 7, 7, 7), dtype=torch.float32).pin_memory()
n = int(np.prod(shape))
src = torch.randn(n, generator=g)
cudart = ctypes.CDLL("libcudart.so.12")
def alloc(flags):
    p = ctypes.c_void_p()
    rc = cudart.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n * 4), ctypes.c_uint(flags))
    assert rc == 0, rc
    arr = np.ctypeslib.as_array((ctypes.c_float * n).from_address(p.value))
    arr[:] = src.numpy()
    return p.value
for name, ptr in (("torch pinned", None), ("cudaHostAlloc default", alloc(0)), ("cudaHostAlloc write-combined", alloc(4))):
    if ptr is None:
        keep = torch.empty(n, dtype=torch.float32).pin_memory(); keep.copy_(src); ptr = keep.data_ptr()
    def step():
        _lib.check(_lib.lib.roi3d_roi_align3d_forward_host(ptr, _lib.NCDHW, 1, 256, 40, 128, 128, rois_h.data_ptr(), 512,
                                                           7, 7, 7, 0.25, 0.5, 2, out_h.data_ptr()))
    for _ in range(2): step()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    for _ in range(8): step()
    dt = (time.perf_counter() - t0) / 8
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print("%-30s e2e ms/step (max over %d ranks) %.2f   checksum %.3f" % (name, world, t.item() * 1e3, float(out_h[:4].sum())), flush=True)
