"""Checkpoint I/O either side of the hot path (train.py:90-100 load, train.py:247-259 / 263-274 save).

Layout: exactly the reference's dict — {"gen", "critic", "iter", "im_count", "step", "epoch", "alpha"} with
DataParallel-style `module.`-prefixed fp32 state_dicts — so generate_samples.py / interpolate.py / train.py of an
unmodified checkout load these files, and files written by the reference load here.  Two OPTIONAL keys are added,
which the reference's loaders never read: "gen_opt" / "critic_opt" (torch.optim.Adam.state_dict()).  Without them a
resumed run restarts Adam's second-moment estimate from zero: with betas (0, 0.99) the first ~100 updates after a
resume are then ~10x too large.  `im_count` is restored into the fade-in schedule as well (the reference stores it,
train.py:254, but resets it to 0 at the stage start, train.py:109, so a resumed stage fades in again from alpha = 0).

Saving is asynchronous: snapshot() copies parameters and optimizer state device->pinned-host on a side stream (the
training stream only waits for an event), AsyncSaver.save() serialises them in a background thread.
"""
from __future__ import annotations

import os
import threading
from typing import Optional

import torch

PREFIX = "module."          # train.py:250-251 saves nn.DataParallel state_dicts


def _strip(sd):
    return {(k[len(PREFIX):] if k.startswith(PREFIX) else k): v for k, v in sd.items()}


def _to_host(obj, stream):
    """Deep copy with every CUDA tensor moved to pinned host memory on `stream` (non-blocking)."""
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            host = torch.empty(obj.shape, dtype=obj.dtype, device="cpu", pin_memory=True)
            with torch.cuda.stream(stream):
                host.copy_(obj, non_blocking=True)
            obj.record_stream(stream)
            return host
        return obj.detach().clone()
    if isinstance(obj, dict):
        return {k: _to_host(v, stream) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_host(v, stream) for v in obj)
    return obj


def snapshot(gen, critic, iters, im_count, step, epoch, alpha, gen_opt=None, critic_opt=None):
    """The checkpoint dict of train.py:247-259 (+ optional optimizer state), every tensor in pinned host memory.
    The copies run on a side stream ordered after the work already queued on the current stream; the returned dict
    carries the event that marks their completion under the private key "_ready" (dropped before writing)."""
    gen_m = gen.module if hasattr(gen, "module") else gen
    critic_m = critic.module if hasattr(critic, "module") else critic
    cuda = next(gen_m.parameters()).is_cuda
    stream = None
    if cuda:
        stream = torch.cuda.Stream()
        stream.wait_stream(torch.cuda.current_stream())
    state = {"gen": {PREFIX + k: v for k, v in gen_m.state_dict().items()},
             "critic": {PREFIX + k: v for k, v in critic_m.state_dict().items()},
             "iter": int(iters), "im_count": int(im_count), "step": int(step), "epoch": int(epoch), "alpha": alpha}
    if gen_opt is not None:
        state["gen_opt"] = gen_opt.state_dict()
    if critic_opt is not None:
        state["critic_opt"] = critic_opt.state_dict()
    state = _to_host(state, stream) if cuda else state
    if cuda:
        ev = torch.cuda.Event()
        ev.record(stream)
        state["_ready"] = ev
    return state


def write(state: dict, path: str):
    """Blocking write of a snapshot (waits for its device->host copies first).  Atomic: temp file + rename."""
    ready = state.pop("_ready", None)
    if ready is not None:
        ready.synchronize()
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = path + ".tmp"
    torch.save(state, tmp)
    os.replace(tmp, path)


class AsyncSaver:
    """save() returns at once; the file is written by a background thread (one write in flight: a second save() waits
    for the first, so checkpoints land in order).  wait() joins the writer and re-raises its exception, if any."""

    def __init__(self):
        self._thread: Optional[threading.Thread] = None
        self._error: Optional[BaseException] = None

    def _run(self, state, path):
        try:
            write(state, path)
        except BaseException as e:  # noqa: BLE001 - reported by wait()
            self._error = e

    def save(self, state: dict, path: str):
        self.wait()
        self._thread = threading.Thread(target=self._run, args=(state, path), daemon=True)
        self._thread.start()

    def wait(self):
        if self._thread is not None:
            self._thread.join()
            self._thread = None
        if self._error is not None:
            err, self._error = self._error, None
            raise err


def load(path: str, gen, critic, gen_opt=None, critic_opt=None, map_location=None) -> dict:
    """train.py:90-100: weights into (possibly DataParallel-wrapped) modules; optimizer state when the file has it and
    an optimizer is given.  Accepts reference files (no optimizer keys) and prefixed or unprefixed state_dicts.
    Returns {"iter", "im_count", "step", "epoch", "alpha", "has_optimizer_state"}."""
    save = torch.load(path, map_location=map_location)
    for module, key in ((gen, "gen"), (critic, "critic")):
        if module is None:
            continue
        wrapped = hasattr(module, "module")
        sd = save[key]
        has_prefix = all(k.startswith(PREFIX) for k in sd)
        if wrapped and not has_prefix:
            sd = {PREFIX + k: v for k, v in sd.items()}
        elif not wrapped and has_prefix:
            sd = _strip(sd)
        module.load_state_dict(sd)                       # strict: a layout mismatch must not pass silently
    had = False
    for opt, key in ((gen_opt, "gen_opt"), (critic_opt, "critic_opt")):
        if opt is not None and key in save:
            opt.load_state_dict(save[key])
            had = True
    return {"iter": int(save.get("iter", 0)), "im_count": int(save.get("im_count", 0)), "step": int(save.get("step", 1)),
            "epoch": int(save.get("epoch", 0)), "alpha": save.get("alpha"), "has_optimizer_state": had}
