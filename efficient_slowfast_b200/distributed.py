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


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index, sysfs="/sys"):
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, BEFORE it allocates pinned host memory.

    Why: with one process per GPU, every rank streams its clips (1.5 GB of FP32 per 64-clip batch) out of pinned host
    memory.  Pages are placed on the node of the thread that first touches them; ranks left on the default node make
    half the GPUs of a two-socket box pull their input across the inter-socket link and through one socket's memory
    controllers (round-1 SCALE: 8 x 1.54 GB per step reached only 179 GB/s aggregate).  The reference leaves this to
    the DataLoader workers' placement (datasets/loader.py pin_memory=True).

    Returns {"node", "cpus"} on success or {"skipped": reason}; never raises (containers may hide sysfs / forbid it)."""
    import os

    try:
        bus = torch.cuda.get_device_properties(device_index)
        pci = "%04x:%02x:%02x.0" % (bus.pci_domain_id, bus.pci_bus_id, bus.pci_device_id)
        node_dir = os.path.join(sysfs, "devices/system/node")
        nodes = sorted(int(n[4:]) for n in os.listdir(node_dir) if n.startswith("node") and n[4:].isdigit())
        how = "sysfs"
        try:
            with open(os.path.join(sysfs, "bus/pci/devices", pci, "numa_node")) as fh:
                node = int(fh.read().strip())
        except OSError:
            node = -1
        if node < 0:
            # virtualised PCI topology (the device reports no node): on HGX boards GPUs are split evenly and in order
            # over the sockets, so the device ordinal picks the node
            if len(nodes) < 2:
                return {"skipped": "device %s reports no NUMA node and the host shows %d node(s)" % (pci, len(nodes))}
            node = nodes[min(len(nodes) - 1, device_index * len(nodes) // max(1, torch.cuda.device_count()))]
            how = "ordinal"
        with open(os.path.join(node_dir, "node%d/cpulist" % node)) as fh:
            cpus = _parse_cpulist(fh.read())
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return {"skipped": "no allowed CPU on node %d" % node}
        os.sched_setaffinity(0, use)
        torch.set_num_threads(max(1, min(torch.get_num_threads(), len(use))))
        return {"node": node, "cpus": len(use), "nodes": len(nodes), "how": how}
    except Exception as e:  # noqa: BLE001 -- best effort by design
        return {"skipped": "%s: %s" % (type(e).__name__, e)}
