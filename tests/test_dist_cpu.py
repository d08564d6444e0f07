"""World-size-2 `gloo` tests (CPU) of the one-process-per-GPU data-parallel plumbing in byo-gan_b200/dist.py,
the replacement for the reference's nn.DataParallel (train.py:71,79): parameters broadcast once, one bucketed
all-reduce(avg) of the ACTIVE parameters' gradients per optimizer step, inactive parameters skipped.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "byo-gan_b200")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, bucket_bytes, out):
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import dist as bdist

    r, w, _ = bdist.init_from_env()
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)                      # different initial values per rank on purpose
    # the 1100 x 1000 weight (4.4 MB) takes the large-gradient path: averaged in place, outside the buckets
    model = torch.nn.ModuleList([torch.nn.Linear(37, 53), torch.nn.Linear(53, 11), torch.nn.Linear(1100, 1000),
                                 torch.nn.Linear(11, 5)])
    bdist.broadcast_parameters(model)
    params = list(model.parameters())
    flat0 = torch.cat([p.detach().reshape(-1) for p in params])
    # gradients: rank-dependent values; the LAST layer is "inactive" (grad None) like an unused progressive stage
    for i, p in enumerate(params[:6]):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    sync = bdist.GradSync(bucket_bytes=bucket_bytes)
    sync.begin()
    for p in params:
        sync.ready(p)                                  # inactive ones (grad None) must be skipped, not waited on
    sync.ready_all(params)                             # the end-of-backward sweep must not reduce anything twice
    sync.finish()
    got = [None if p.grad is None else p.grad.clone() for p in params]
    gathered = [torch.zeros_like(flat0) for _ in range(world)]
    dist.all_gather(gathered, flat0)
    if rank == 0:
        torch.save({"grads": got, "params_equal": all(torch.equal(g, gathered[0]) for g in gathered),
                    "bytes": sync.bytes_reduced}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes", [1, 4096, 25 * 1024 * 1024])
def test_gradsync_world2_gloo(tmp_path, bucket_bytes):
    world, out = 2, str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, _free_port(), bucket_bytes, out), nprocs=world, join=True)
    res = torch.load(out)
    assert res["params_equal"], "broadcast_parameters must leave identical replicas"
    mean_scale = (1 + 2) / 2.0                          # average of rank+1 over the two ranks
    for i, g in enumerate(res["grads"]):
        if i < 6:
            assert torch.allclose(g, torch.full_like(g, mean_scale * (i + 1))), (i, g.flatten()[:4])
        else:
            assert g is None, "inactive parameter gradients stay None on every rank"
    assert res["bytes"] == 4 * (37 * 53 + 53 + 53 * 11 + 11 + 1100 * 1000 + 1000)


def test_single_process_is_a_noop():
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import dist as bdist

    sync = bdist.GradSync()
    assert not sync.enabled
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    sync.begin()
    sync.ready(p)
    sync.finish()
    assert torch.equal(p.grad, torch.full((3,), 2.0))
