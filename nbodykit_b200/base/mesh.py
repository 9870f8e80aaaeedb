"""
MeshSource -- "something that can be put on a mesh" (the API of nbodykit/base/mesh.py:8-412), organised around the
device fields it produces:

  a source knows how to make ONE of its two representations (`to_real_field` / `to_complex_field`, the other one
  returns NotImplemented); `compute(mode, Nmesh)` turns that into a device RealField / ComplexField, runs the queued
  `actions` (transfer functions in configuration or Fourier space) on the device representation they ask for,
  converts to the requested mode and, if another mesh size is wanted, resamples in Fourier space
  (nbk_resample_complex).  Views made by `apply()` share the parent's data and extend its action queue.
"""
import logging
import warnings

import numpy

from ..pmesh.pm import BaseComplexField, ParticleMesh, RealField

_MODES = ('real', 'complex')


def _is_mode(field, mode):
    return isinstance(field, BaseComplexField) if mode == 'complex' else isinstance(field, RealField)


def _as_mode(field, mode):
    """the field in the wanted representation (one FFT if it is in the other one)"""
    if _is_mode(field, mode):
        return field
    out = field.r2c(out=Ellipsis) if mode == 'complex' else field.c2r(out=Ellipsis)
    out.attrs = getattr(field, 'attrs', {})
    return out


class MeshSource(object):
    """
    Parameters
    ----------
    comm : communicator
    Nmesh : int or 3-vector
    BoxSize : float or 3-vector
    dtype : str
        type of the numbers on the mesh: 'f4' / 'f8', or 'c8' / 'c16' for a mesh that keeps all N^3 Fourier modes
    """
    logger = logging.getLogger('MeshSource')

    def __init__(self, comm, Nmesh, BoxSize, dtype):
        if Nmesh is None or BoxSize is None:
            raise ValueError("both Nmesh and BoxSize must not be None to initialize ParticleMesh")
        self.comm = comm
        self.dtype = dtype
        side = numpy.array(Nmesh)
        cells = numpy.empty(3 if side.ndim == 0 else len(side), dtype='i8')
        cells[:] = side
        self.pm = ParticleMesh(BoxSize=BoxSize, Nmesh=cells, dtype=dtype, comm=comm)
        self.attrs['BoxSize'] = self.pm.BoxSize.copy()
        self.attrs['Nmesh'] = self.pm.Nmesh.copy()
        self._actions = []
        self.base = None

    # ---- attrs / actions / views
    @property
    def attrs(self):
        if not hasattr(self, '_attrs'):
            self._attrs = {}
        return self._attrs

    @property
    def actions(self):
        """the queue of ``(mode, func, kind)`` run when the mesh is computed"""
        return self._actions

    def __finalize__(self, other):
        if isinstance(other, MeshSource):
            self.comm, self.dtype, self.pm = other.comm, other.dtype, other.pm
            self.attrs.update(other.attrs)
            self._actions = list(other.actions)
        return self

    def view(self):
        """a new MeshSource sharing this one's data (base/mesh.py:82-93)"""
        v = object.__new__(MeshSource)
        v.base = self
        return v.__finalize__(self)

    def apply(self, func, kind='wavenumber', mode='complex'):
        """a view whose queue ends with ``func(x, v)``, evaluated in configuration ('real': kind 'relative' | 'index')
        or Fourier space ('complex': kind 'wavenumber' | 'circular' | 'index'); a MeshFilter brings its own kind and
        mode (contract in base/mesh.py:118-176)"""
        if isinstance(func, type) and issubclass(func, MeshFilter):
            func = func()
        if isinstance(func, MeshFilter):
            mode, kind, func = func.mode, func.kind, func.filter
        assert mode in _MODES, "``mode`` should be 'complex' or 'real'"
        allowed = ('relative', 'index') if mode == 'real' else ('wavenumber', 'circular', 'index')
        assert kind in allowed
        v = self.view()
        v._actions.append((mode, func, kind))
        return v

    def __len__(self):
        return 0

    # ---- the two representations; a concrete source overrides one of them, a view defers to its parent
    def to_real_field(self, out=None, normalize=True):
        return self.base.to_real_field() if isinstance(self.base, MeshSource) else NotImplemented

    def to_complex_field(self, out=None):
        return self.base.to_complex_field() if isinstance(self.base, MeshSource) else NotImplemented

    def to_field(self, mode='real', out=None):
        """the mesh as a device Field in configuration ('real') or Fourier ('complex') space, before any action"""
        if mode not in _MODES:
            raise ValueError("mode is either real or complex, %s given" % mode)
        makers = {'real': (self.to_real_field, self.to_complex_field), 'complex': (self.to_complex_field, self.to_real_field)}
        own, other = makers[mode]
        field = own()
        if field is NotImplemented:
            field = other()
            assert field is not NotImplemented
            field = _as_mode(field, mode)
        return field

    # ---- compute
    def compute(self, mode='real', Nmesh=None):
        """the mesh in HBM as a RealField or ComplexField with all :attr:`actions` applied; `Nmesh` other than the
        source's resamples the result in Fourier space"""
        return self._paint_XXX(mode=mode, Nmesh=Nmesh)

    def paint(self, mode="real", Nmesh=None):
        warnings.warn("the paint method is deprecated from the Public API. Use .compute() instead.", DeprecationWarning)
        return self._paint_XXX(mode=mode, Nmesh=Nmesh)

    def _paint_XXX(self, mode="real", Nmesh=None):
        if mode not in _MODES:
            raise ValueError('mode must be "real" or "complex"')
        queue = list(self.actions)
        field = self.to_field(mode=queue[0][0] if queue else mode)
        attrs = getattr(field, 'attrs', {})
        for amode, func, kind in queue:
            field = _as_mode(field, amode)
            kw = {'func': func, 'out': Ellipsis}
            if kind is not None:
                kw['kind'] = kind
            field.apply(**kw)
        field = _as_mode(field, mode)
        target = self.pm.reshape(Nmesh=Nmesh)
        if any(target.Nmesh != self.pm.Nmesh):
            field = field.resample(out=target.create(type=mode))
        field.attrs = attrs
        field.attrs.update(self.attrs)
        if self.comm.rank == 0:
            self.logger.info('field: %s painting done' % str(self))
        return field

    def preview(self, axes=None, Nmesh=None, root=0):
        """the real field as a numpy array on the host, optionally summed over the axes not in `axes`
        (base/mesh.py:340-365); the projection runs on the device"""
        return self.to_field(mode='real').preview(Nmesh=Nmesh, axes=axes)

    def save(self, output, dataset='Field', mode='real'):
        """write the computed mesh and its attrs to a directory of .npy blocks + attrs.json (the role of the
        reference's bigfile output, base/mesh.py:367-412; bigfile itself is not available here).  Read back with
        `nbodykit_b200.source.mesh.file.FileMesh`."""
        from ..source.mesh.file import save_mesh
        return save_mesh(self, output, dataset=dataset, mode=mode)


class MeshFilter(object):
    """a filter applied to a mesh; subclasses set `kind`, `mode` and implement `filter(x, v)`"""
    kind = None
    mode = None

    def filter(self, x, v):
        raise NotImplementedError

    def __call__(self, x, v):
        return self.filter(x, v)
