"""``upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0))`` — same signature and semantics as the reference wrapper
(model/stylegan/op/upfirdn2d.py:149-165): ``up``/``down`` int or (x, y); ``pad`` (p0, p1) -> (p0, p1, p0, p1) or
(x0, x1, y0, y1).  Forward only (inference); CUDA tensors only — the reference routes CPU tensors to
``upfirdn2d_native``; this library has no CPU path and raises instead."""
from collections import abc

from .. import ops


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    if not isinstance(up, abc.Iterable):
        up = (up, up)
    if not isinstance(down, abc.Iterable):
        down = (down, down)
    if len(pad) == 2:
        pad = (pad[0], pad[1], pad[0], pad[1])
    return ops.upfirdn2d_planar(input, kernel, tuple(int(u) for u in up), tuple(int(d) for d in down),
                                tuple(int(p) for p in pad))
