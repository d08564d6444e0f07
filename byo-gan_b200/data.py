"""Data feed for the training loop (train.py:43-50 transforms, train.py:109-117 ImageFolder + DataLoader,
train.py:150-158 resize + .to(device)).

The reference decodes, flips, converts to float and normalises every image on the host (PIL + torchvision transforms in
`dataloader_threads` workers), collates float64->float32 batches and copies 12 bytes per pixel to the device inside the
training loop.  Here the host only produces uint8 HWC batches (3 bytes per pixel) in pinned memory from a background
thread; the copy runs on its own stream one batch ahead; flip, float conversion, Normalize((.5,.5,.5),(.5,.5,.5)) and the
HWC->CHW transpose happen in ONE kernel on the device (bg_image_feed_u8).  Sharding across processes follows
torch.utils.data.DistributedSampler: one seeded permutation per epoch, wrap-padded so that every rank draws the same
number of batches (the gradient all-reduce needs equal iteration counts), rank r takes indices r, r + world, ...
"""
from __future__ import annotations

import math
import queue
import threading
from typing import Optional

import torch

import bg_native as bgn


def shard_indices(n: int, epoch: int, rank: int, world: int, seed: int = 0, shuffle: bool = True):
    """Indices of this rank for `epoch` (DistributedSampler semantics: shuffle=True of train.py:115, then pad + stride)."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        order = torch.randperm(n, generator=g).tolist()
    else:
        order = list(range(n))
    per_rank = math.ceil(n / world)
    total = per_rank * world
    if total > n:
        order += order[: total - n] if order else []
    return order[rank:total:world]


class TensorSource:
    """Images already in memory: uint8 (N, H, W, 3)."""

    def __init__(self, images_u8: torch.Tensor):
        assert images_u8.dtype == torch.uint8 and images_u8.dim() == 4 and images_u8.shape[3] == 3
        self.images = images_u8

    def __len__(self):
        return self.images.shape[0]

    def shape(self):
        return tuple(self.images.shape[1:3])

    def fetch(self, indices, out: torch.Tensor):
        torch.index_select(self.images, 0, torch.as_tensor(indices), out=out[: len(indices)])


class FolderSource:
    """prepared/set_k laid out for torchvision.datasets.ImageFolder (train.py:110-112); decodes with PIL on the host."""

    def __init__(self, root: str):
        from torchvision import datasets

        self.ds = datasets.ImageFolder(root)
        import numpy as np

        self._np = np
        im, _ = self.ds[0]
        self._hw = (im.size[1], im.size[0])

    def __len__(self):
        return len(self.ds)

    def shape(self):
        return self._hw

    def fetch(self, indices, out: torch.Tensor):
        for j, i in enumerate(indices):
            im, _ = self.ds[i]
            out[j].copy_(torch.from_numpy(self._np.asarray(im.convert("RGB"))))


def device_transform(u8: torch.Tensor, flip: Optional[torch.Tensor] = None) -> torch.Tensor:
    """uint8 (B,H,W,3) on the device -> float32 (B,3,H,W) = ToTensor + Normalize(.5,.5) (x / 127.5 - 1), with
    RandomHorizontalFlip applied to the samples whose `flip` entry (uint8 (B,), device) is non-zero (train.py:43-50)."""
    b, h, w, _ = u8.shape
    out = torch.empty(b, 3, h, w, dtype=torch.float32, device=u8.device)
    bgn.call("bg_image_feed_u8", u8, flip, out, b, h, w)
    return out


class ImageFeed:
    """Re-iterable over one stage's batches: `for real in feed:` yields float32 (B,3,R,R) CUDA tensors in [-1, 1].
    len(feed) = batches per epoch on this rank (what train.py:121 multiplies into the fade-in length)."""

    def __init__(self, source, batch: int, device, rank: int = 0, world: int = 1, seed: int = 0, flip: bool = True,
                 resolution: Optional[int] = None, prefetch: int = 2, shuffle: bool = True):
        self.source, self.batch, self.device = source, batch, torch.device(device)
        self.rank, self.world, self.seed, self.flip, self.shuffle = rank, world, seed, flip, shuffle
        self.resolution, self.prefetch = resolution, prefetch
        self.epoch = 0
        self._n_rank = math.ceil(len(source) / world)

    def __len__(self):
        return math.ceil(self._n_rank / self.batch)

    def _producer(self, idx, q, bufs, flips):
        try:
            for k in range(0, len(idx), self.batch):
                chunk = idx[k:k + self.batch]
                slot = bufs.get()                                   # a free pinned buffer (back-pressure)
                self.source.fetch(chunk, slot)
                q.put((slot, len(chunk), flips[k:k + len(chunk)]))
            q.put(None)
        except BaseException as e:  # noqa: BLE001 - surfaced in the consumer
            q.put(e)

    def __iter__(self):
        idx = shard_indices(len(self.source), self.epoch, self.rank, self.world, self.seed, self.shuffle)
        g = torch.Generator()
        g.manual_seed(1_000_003 * (self.seed + self.epoch) + self.rank)
        flips = (torch.rand(len(idx), generator=g) < 0.5).to(torch.uint8) if self.flip else torch.zeros(len(idx), dtype=torch.uint8)
        self.epoch += 1
        h, w = self.source.shape()
        bufs: "queue.Queue" = queue.Queue()
        for _ in range(self.prefetch + 1):
            bufs.put(torch.empty(self.batch, h, w, 3, dtype=torch.uint8).pin_memory())
        q: "queue.Queue" = queue.Queue(maxsize=self.prefetch)
        threading.Thread(target=self._producer, args=(idx, q, bufs, flips), daemon=True).start()
        copy_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)

        def upload():
            item = q.get()
            if item is None:
                return None
            if isinstance(item, BaseException):
                raise item
            slot, n, fl = item
            with torch.cuda.stream(copy_stream):
                dev_u8 = slot[:n].to(self.device, non_blocking=True)
                dev_fl = fl.pin_memory().to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return slot, dev_u8, dev_fl, ev

        nxt = upload()
        while nxt is not None:
            slot, dev_u8, dev_fl, ev = nxt
            main.wait_event(ev)
            dev_u8.record_stream(main)
            dev_fl.record_stream(main)
            ev.synchronize()                                        # the pinned buffer may be refilled now
            bufs.put(slot)
            nxt = upload()                                          # next batch's copy runs under this batch's step
            real = device_transform(dev_u8, dev_fl)
            if self.resolution is not None and real.shape[2] != self.resolution:       # train.py:150-156
                real = torch.nn.functional.interpolate(real, size=(self.resolution, self.resolution), mode="bilinear")
            yield real
