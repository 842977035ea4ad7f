"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch, run the (CPU oracle) forward on their shard and
all-gather the outputs; the gathered result must equal the single-process forward on the whole batch."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import helpers
    import recipe
    from efficient_slowfast_b200 import distributed as D
    from oracle import slowfast_oracle as O

    torch.set_num_threads(2)
    cfg, model, gold = helpers.case_model_and_weights("shufflenetv2_w05")
    x = recipe.seeded_clip(4, 16, 64, seed=3)
    slow, fast = recipe.pack_pathway_output(x, cfg.SLOWFAST.ALPHA)
    labels = torch.arange(4)
    s, f, lab = D.shard_batch([slow, fast, labels], rank, world)
    y = O.forward(cfg, model.state_dict(), [s, f])
    preds, labs = D.all_gather([y, lab])
    if rank == 0:
        full = O.forward(cfg, model.state_dict(), [slow, fast])
        q.put((preds, labs, full))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_forward_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    preds, labs, full = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(labs, torch.arange(4))
    assert torch.allclose(preds, full, atol=1e-6, rtol=1e-5)   # clips are independent: sharding changes nothing


def test_shard_batch_requires_divisible_batch():
    from efficient_slowfast_b200 import distributed as D

    with pytest.raises(AssertionError):
        D.shard_batch([torch.zeros(3, 2)], 0, 2)
    a, = D.shard_batch([torch.arange(8).reshape(4, 2)], 1, 2)
    assert a.tolist() == [[4, 5], [6, 7]]
    assert D.all_gather([torch.ones(2)])[0].tolist() == [1.0, 1.0]   # no process group: identity


def test_numa_bind_is_best_effort_and_never_raises(tmp_path):
    """bind_to_gpu_numa_node on a box without a GPU / with a hidden topology reports why it did nothing."""
    import os

    from efficient_slowfast_b200 import distributed as esf_dist

    before = os.sched_getaffinity(0)
    out = esf_dist.bind_to_gpu_numa_node(0, sysfs=str(tmp_path))
    assert "skipped" in out or "node" in out
    assert os.sched_getaffinity(0) == before or "node" in out
    assert esf_dist._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
