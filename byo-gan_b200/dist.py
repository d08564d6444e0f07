"""Data-parallel plumbing for the one-process-per-GPU launch (replaces the reference's nn.DataParallel,
train.py:71,79): parameters are broadcast once from rank 0, then every optimizer step is preceded by ONE
bucketed NCCL all-reduce (average) of the gradients of the parameters that are active at the current
progressive stage.  Inactive parameters have .grad None on every rank and are skipped, so
no rank waits on a bucket that never fills.

Buckets are flat fp32 buffers (~25 MB) filled in the order gradients become ready and reduced on NCCL's own
stream with async_op=True, so the reduction of early buckets overlaps the rest of the backward; finish()
joins them and re-points each small .grad at its slice of the reduced bucket before optimizer.step().
The batch is sharded by construction (each rank draws its own B_local samples); the minibatch-stddev layer
sees per-rank statistics exactly like a DataParallel replica does (gan.py:273-298 under train.py:79).
"""
from __future__ import annotations

import os
from typing import List

import torch
import torch.distributed as dist

BUCKET_BYTES = 25 * 1024 * 1024
LARGE_BYTES = 4 * 1024 * 1024


def init_from_env():
    """torchrun-style rendezvous; returns (rank, world, local_rank).  world == 1 -> no process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def broadcast_parameters(module: torch.nn.Module, src: int = 0):
    """Once, at start-up: replicas stay in sync afterwards because they apply identical averaged gradients
    (the reference re-broadcasts ~82 MB of parameters on every forward instead)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src)


class GradSync:
    """Bucketed, overlapped gradient averaging.  Usage per optimizer step:

        sync.begin()
        ... backward; call sync.ready(p) as each p.grad becomes final (or sync.ready_all(params) at the end)
        sync.finish()      # before optimizer.step()
    """

    def __init__(self, bucket_bytes: int = BUCKET_BYTES):
        self.bucket_bytes = bucket_bytes
        self.enabled = dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if self.enabled else 1
        self._pending: List[torch.nn.Parameter] = []
        self._pending_bytes = 0
        self._inflight = []
        self._scale_later: List[torch.nn.Parameter] = []
        self._seen = set()
        self.bytes_reduced = 0
        # NCCL averages inside the collective; gloo (CPU tests) has no AVG: sum, then scale in finish()
        self._avg_op = dist.ReduceOp.SUM
        if self.enabled and dist.get_backend() == "nccl":
            self._avg_op = dist.ReduceOp.AVG

    def begin(self):
        self._pending, self._pending_bytes, self._inflight, self._scale_later = [], 0, [], []
        self._seen = set()

    def ready(self, p: torch.nn.Parameter):
        if not self.enabled or p.grad is None:
            return
        if id(p) in self._seen:        # already queued in this step (layer hook first, ready_all() sweep afterwards)
            return
        self._seen.add(id(p))
        if p.grad.numel() * p.grad.element_size() >= LARGE_BYTES and p.grad.is_contiguous():
            # big conv / FC gradients (>= 4 MB; they are ~90 % of the payload) are averaged IN PLACE, each as its own
            # collective: no flatten copy, no scatter-back copy, no separate division pass
            work = dist.all_reduce(p.grad, op=self._avg_op, async_op=True)
            if self._avg_op is dist.ReduceOp.SUM:
                self._scale_later.append(p)
            self._inflight.append((work, None, [p]))
            self.bytes_reduced += p.grad.numel() * 4
            return
        self._pending.append(p)
        self._pending_bytes += p.grad.numel() * p.grad.element_size()
        if self._pending_bytes >= self.bucket_bytes:
            self._flush()

    def ready_all(self, params):
        for p in params:
            self.ready(p)

    def _flush(self):
        if not self._pending:
            return
        params, self._pending, self._pending_bytes = self._pending, [], 0
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        if self._avg_op is dist.ReduceOp.SUM:
            flat.div_(self.world)
        work = dist.all_reduce(flat, op=self._avg_op, async_op=True)
        self._inflight.append((work, flat, params))
        self.bytes_reduced += flat.numel() * 4

    def finish(self):
        if not self.enabled:
            return
        self._flush()
        for work, flat, params in self._inflight:
            work.wait()
            if flat is None:
                continue
            # the averaged values stay where they are: every small gradient becomes a VIEW of the reduced bucket (no
            # scatter copies — ~130 tiny launches per iteration on the critical path between the all-reduce and Adam)
            off = 0
            for p in params:
                n = p.grad.numel()
                p.grad = flat[off:off + n].view(p.grad.shape)
                off += n
        for p in self._scale_later:
            p.grad.div_(self.world)
        self._inflight, self._scale_later = [], []
