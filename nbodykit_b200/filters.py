"""
Mesh filters (nbodykit/filters.py:5-60): `TopHat(r)` and `Gaussian(r)` Fourier-space windows for `mesh.apply(...)`.

`filter(k, v)` is written against the array type it is handed: `Field.apply` calls it with DEVICE tensors (torch,
batches of planes of the field and broadcastable wavenumber tensors), so the window is evaluated and applied on the GPU
without the field ever leaving HBM; called by hand with NumPy arrays it behaves like the reference's.  Any user callback
written with arithmetic operators and the helpers of `xp(v)` runs on the device the same way; callbacks that insist on
NumPy fall back to the host plane loop.
"""
import numpy

from .base.mesh import MeshFilter


def xp(a):
    """the array namespace of `a`: torch for device tensors, numpy otherwise"""
    try:
        import torch
        if isinstance(a, torch.Tensor):
            return torch
    except ImportError:      # pragma: no cover
        pass
    return numpy


def _where(m, a, b):
    ns = xp(b)
    if ns is numpy:
        return numpy.where(m, a, b)
    return ns.where(m, ns.as_tensor(a, dtype=b.dtype, device=b.device), b)


class TopHat(MeshFilter):
    """spherical top hat of radius r, as a Fourier-space window W(kr) = 3 (sin kr - kr cos kr) / (kr)^3"""
    kind = 'wavenumber'
    mode = 'complex'

    def __init__(self, r):
        self.r = r

    def filter(self, k, v):
        ns = xp(v)
        kr = sum(ki ** 2 for ki in k) ** 0.5 * self.r
        safe = _where(kr == 0, 1.0, kr)
        w = 3 * (ns.sin(safe) / safe ** 3 - ns.cos(safe) / safe ** 2)
        w = _where(kr == 0, 1.0, w)
        return w * v


class Gaussian(MeshFilter):
    """Gaussian window exp(-k^2 r^2 / 2)"""
    kind = 'wavenumber'
    mode = 'complex'

    def __init__(self, r):
        self.r = r

    def filter(self, k, v):
        ns = xp(v)
        k2 = sum(ki ** 2 for ki in k)
        return ns.exp(-0.5 * k2 * self.r ** 2) * v
