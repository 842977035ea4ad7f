"""Multi-GPU plumbing of the forward path: clips are independent in eval mode, so the batch is sharded across one
process per GPU with NO collective inside the forward; the only exchange is the all-gather of the per-rank outputs
that the reference's test loop performs (`du.all_gather([preds, labels, video_idx])`,
SlowFast/tools/test_net.py:95-98 -> slowfast/utils/distributed.py:15-34).  Works on any torch.distributed backend
(NCCL over NVLink on the B200 box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_batch(tensors, rank, world_size):
    """Contiguous batch shard of every tensor for `rank` (what DistributedSampler gives each process of the reference,
    datasets/loader.py:87-104).  The batch must divide evenly (config/defaults.py:624 asserts the same)."""
    out = []
    for t in tensors:
        assert t.shape[0] % world_size == 0, "batch %d not divisible by world size %d" % (t.shape[0], world_size)
        n = t.shape[0] // world_size
        out.append(t[rank * n:(rank + 1) * n])
    return out


def all_gather(tensors):
    """utils/distributed.py:15-34 of the reference: all-gather every tensor of the list along dim 0 (rank order).
    One `all_gather_into_tensor` per tensor on the current stream; no host synchronisation."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(tensors)
    world = dist.get_world_size()
    outs = []
    for t in tensors:
        t = t.contiguous()
        g = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(g, t)
        outs.append(g)
    return outs
