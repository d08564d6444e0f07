"""Host-side overhead shims for an UNMODIFIED train.py.  Both are opt-in and change no arithmetic.

train.py synchronises the host with the device twice per iteration (`c_loss.item()`, `g_loss.item()`, train.py:191,219)
and runs a 25-image preview forward every iteration although it only looks at the result every `display_step`
iterations (train.py:236-237).  With the step at ~18 ms on a B200 those dominate the loop's wall clock.

* LazyScalar: what `.item()` returns when `gan.DEFER_LOSS_ITEMS` (env BG_DEFER_ITEMS=1) is on.  The value is copied to
  pinned host memory without blocking; the object behaves like a float and waits for the copy only when it is first USED
  as a number — for train.py that is the `sum(history[-n:]) / n` of the progress bar every `refresh_stat_step` iterations.
* LazyImages: what `Generator.forward` returns under `torch.no_grad()` when `gan.LAZY_NO_GRAD_FORWARD` (env
  BG_LAZY_PREVIEW=1) is on.  The per-layer noise is drawn immediately (so torch's RNG stream is consumed exactly as in
  the reference), the forward itself runs when the images are first touched by any torch function (`torch.clamp` at
  train.py:239) — i.e. never, on the iterations that do not display them.
"""
from __future__ import annotations

import operator

import torch

_PINNED = None
_NEXT = 0
_RING = 4096


def _pinned_slot():
    global _PINNED, _NEXT
    if _PINNED is None:
        _PINNED = torch.zeros(_RING, dtype=torch.float32).pin_memory()
    i = _NEXT
    _NEXT = (_NEXT + 1) % _RING
    return _PINNED[i:i + 1]


class LazyScalar:
    """float-like result of a deferred `.item()`."""

    __slots__ = ("_buf", "_event", "_value")

    def __init__(self, tensor: torch.Tensor):
        self._value = None
        self._buf = _pinned_slot()
        self._buf.copy_(tensor.detach().reshape(1).float(), non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record()

    def resolve(self) -> float:
        if self._value is None:
            self._event.synchronize()
            self._value = float(self._buf)
            self._buf = self._event = None
        return self._value

    __float__ = resolve

    def __int__(self):
        return int(self.resolve())

    def __bool__(self):
        return bool(self.resolve())

    def __repr__(self):
        return repr(self.resolve())

    __str__ = __repr__

    def __format__(self, spec):
        return format(self.resolve(), spec)

    def __hash__(self):
        return hash(self.resolve())

    def __neg__(self):
        return -self.resolve()

    def __abs__(self):
        return abs(self.resolve())

    def __round__(self, n=None):
        return round(self.resolve(), n)


def _binary(op, reflected=False):
    def fn(self, other):
        a = self.resolve()
        b = other.resolve() if isinstance(other, LazyScalar) else other
        return op(b, a) if reflected else op(a, b)

    return fn


for _name, _op in (("add", operator.add), ("sub", operator.sub), ("mul", operator.mul), ("truediv", operator.truediv),
                   ("floordiv", operator.floordiv), ("mod", operator.mod), ("pow", operator.pow)):
    setattr(LazyScalar, f"__{_name}__", _binary(_op))
    setattr(LazyScalar, f"__r{_name}__", _binary(_op, reflected=True))
for _name, _op in (("lt", operator.lt), ("le", operator.le), ("gt", operator.gt), ("ge", operator.ge), ("eq", operator.eq),
                   ("ne", operator.ne)):
    setattr(LazyScalar, f"__{_name}__", _binary(_op))


class DeferredItemTensor(torch.Tensor):
    """A loss tensor whose `.item()` does not synchronise (returns a LazyScalar); everything else is a plain tensor."""

    def item(self):
        return LazyScalar(self)


def defer_item(loss: torch.Tensor) -> torch.Tensor:
    return loss.as_subclass(DeferredItemTensor)


class LazyImages(torch.Tensor):
    """Images that are computed on first use.  Shape / dtype / device are known up front (the object wraps a storage-less
    meta tensor of the right shape; any torch function applied to it first runs the deferred forward)."""

    @staticmethod
    def __new__(cls, thunk, shape, device):
        t = torch.Tensor._make_subclass(cls, torch.empty(tuple(shape), dtype=torch.float32, device="meta"), False)
        t._thunk = thunk
        t._value = None
        t._real_device = torch.device(device)
        return t

    def materialize(self) -> torch.Tensor:
        if self._value is None:
            self._value = self._thunk()
            self._thunk = None
        return self._value

    @property
    def computed(self) -> bool:
        return self._value is not None

    def __repr__(self):
        return f"LazyImages(shape={tuple(self.shape)}, device={self._real_device}, computed={self.computed})"

    _PASSIVE = None

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if cls._PASSIVE is None:
            T = torch.Tensor
            cls._PASSIVE = {T.shape.__get__, T.dtype.__get__, T.size, T.dim, T.__len__, T.requires_grad.__get__, T.ndim.__get__,
                            T.numel}
        if func in cls._PASSIVE:                       # metadata only: answered by the meta tensor, nothing runs
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        if func == torch.Tensor.device.__get__:
            return args[0]._real_device
        if func == torch.Tensor.is_cuda.__get__:
            return args[0]._real_device.type == "cuda"

        def unwrap(a):
            if isinstance(a, LazyImages):
                return a.materialize()
            if isinstance(a, (list, tuple)):
                return type(a)(unwrap(x) for x in a)
            return a

        return func(*[unwrap(a) for a in args], **{k: unwrap(v) for k, v in kwargs.items()})
