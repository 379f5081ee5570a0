"""Multi-GPU plumbing of the RoI path: volumes are independent, so they are sharded round-robin over ranks
(the reference's `range(rank, len(dataset), world_size)`, mmdet/core/evaluation/eval_hooks.py:118) with NO
collective on the data path.  The only exchange is assembling per-volume detections for evaluation, which the
reference does through pickle files on disk plus barriers (eval_hooks.py:134-149); here it is one padded
all_gather over NCCL (gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank=None, world_size=None):
    """Indices of the volumes this rank owns."""
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    return list(range(rank, n_items, world_size))


def gather_detections(local_dets, local_labels, local_ids, group=None):
    """All-gather variable-length detections.

    local_dets: list of [n_i, 7] fp32 tensors (one per local volume), local_labels: list of [n_i] int64,
    local_ids: list of global volume ids.  Returns {volume_id: (dets, labels)} on every rank.
    Two small collectives: counts, then one padded payload (<= 64 KB per volume -> latency bound).
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return {i: (d, l) for i, d, l in zip(local_ids, local_dets, local_labels)}
    world = dist.get_world_size(group)
    dev = local_dets[0].device if local_dets else torch.device('cpu')
    if dist.get_backend(group) == 'nccl' and dev.type != 'cuda':
        dev = torch.device('cuda', torch.cuda.current_device())
    nloc = len(local_dets)
    meta = torch.tensor([nloc], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    max_vol = int(max(int(m.item()) for m in metas))
    if max_vol == 0:
        return {}
    counts = torch.zeros((max_vol, 2), dtype=torch.int64, device=dev)  # (volume id, n)
    counts[:, 0] = -1
    for j, (i, d) in enumerate(zip(local_ids, local_dets)):
        counts[j, 0], counts[j, 1] = i, d.shape[0]
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    max_n = int(max(int(c[:, 1].max().item()) for c in all_counts))
    payload = torch.zeros((max_vol, max(max_n, 1), 8), dtype=torch.float32, device=dev)
    for j, (d, l) in enumerate(zip(local_dets, local_labels)):
        n = d.shape[0]
        if n:
            payload[j, :n, :7] = d.to(dev)
            payload[j, :n, 7] = l.to(dev).to(torch.float32)
    all_payload = [torch.zeros_like(payload) for _ in range(world)]
    dist.all_gather(all_payload, payload, group=group)
    out = {}
    for c, p in zip(all_counts, all_payload):
        for j in range(max_vol):
            vid, n = int(c[j, 0].item()), int(c[j, 1].item())
            if vid >= 0:
                out[vid] = (p[j, :n, :7].clone(), p[j, :n, 7].to(torch.int64))
    return out
