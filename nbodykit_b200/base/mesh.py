"""
MeshSource -- a source of data on a mesh (API of nbodykit/base/mesh.py): holds the ParticleMesh,
the list of `actions` (transfer functions applied in real / Fourier space), and turns itself into a
device RealField / ComplexField through `compute(mode)`.
"""
import logging
import warnings

import numpy

from ..pmesh.pm import BaseComplexField, ParticleMesh, RealField, _typestr_to_type


class MeshSource(object):
    """
    Parameters
    ----------
    comm : communicator
    Nmesh : int or 3-vector
    BoxSize : float or 3-vector
    dtype : str
        type of the real numbers on the mesh, 'f4' or 'f8'
    """
    logger = logging.getLogger('MeshSource')

    def __init__(self, comm, Nmesh, BoxSize, dtype):
        self.comm = comm
        self.dtype = dtype
        if Nmesh is None or BoxSize is None:
            raise ValueError("both Nmesh and BoxSize must not be None to initialize ParticleMesh")
        Nmesh = numpy.array(Nmesh)
        ndim = 3 if Nmesh.ndim == 0 else len(Nmesh)
        _Nmesh = numpy.empty(ndim, dtype='i8')
        _Nmesh[:] = Nmesh
        self.pm = ParticleMesh(BoxSize=BoxSize, Nmesh=_Nmesh, dtype=self.dtype, comm=self.comm)
        self.attrs['BoxSize'] = self.pm.BoxSize.copy()
        self.attrs['Nmesh'] = self.pm.Nmesh.copy()
        self._actions = []
        self.base = None

    def __finalize__(self, other):
        if isinstance(other, MeshSource):
            self.comm = other.comm
            self.dtype = other.dtype
            self.pm = other.pm
            self.attrs.update(other.attrs)
            self._actions = []
            self._actions.extend(other.actions)
        return self

    def view(self):
        """a new MeshSource sharing this one's data (base/mesh.py:82-93)"""
        view = object.__new__(MeshSource)
        view.base = self
        return view.__finalize__(self)

    @property
    def attrs(self):
        try:
            return self._attrs
        except AttributeError:
            self._attrs = {}
            return self._attrs

    @property
    def actions(self):
        """list of ``(mode, func, kind)`` applied when the mesh is computed"""
        return self._actions

    def apply(self, func, kind='wavenumber', mode='complex'):
        """a view with one more action: ``func(x, v)`` evaluated in real or Fourier space
        (contract in base/mesh.py:118-176)"""
        if isinstance(func, type) and issubclass(func, MeshFilter):
            func = func()
        if isinstance(func, MeshFilter):
            mode = func.mode
            kind = func.kind
            func = func.filter
        assert mode in ['complex', 'real'], "``mode`` should be 'complex' or 'real'"
        if mode == 'real':
            assert kind in ['relative', 'index']
        else:
            assert kind in ['wavenumber', 'circular', 'index']
        view = self.view()
        view._actions.append((mode, func, kind))
        return view

    def __len__(self):
        return 0

    def to_real_field(self, out=None, normalize=True):
        if isinstance(self.base, MeshSource):
            return self.base.to_real_field()
        return NotImplemented

    def to_complex_field(self, out=None):
        if isinstance(self.base, MeshSource):
            return self.base.to_complex_field()
        return NotImplemented

    def to_field(self, mode='real', out=None):
        """the mesh as a Field in configuration ('real') or Fourier ('complex') space"""
        if mode == 'real':
            real = self.to_real_field()
            if real is NotImplemented:
                cplx = self.to_complex_field()
                assert cplx is not NotImplemented
                real = cplx.c2r(out=Ellipsis)
                if hasattr(cplx, 'attrs'):
                    real.attrs = cplx.attrs
            var = real
        elif mode == 'complex':
            cplx = self.to_complex_field()
            if cplx is NotImplemented:
                real = self.to_real_field()
                assert real is not NotImplemented
                cplx = real.r2c(out=Ellipsis)
                if hasattr(real, 'attrs'):
                    cplx.attrs = real.attrs
            var = cplx
        else:
            raise ValueError("mode is either real or complex, %s given" % mode)
        return var

    def compute(self, mode='real', Nmesh=None):
        """compute the mesh into HBM as a RealField or ComplexField, applying :attr:`actions`"""
        return self._paint_XXX(mode=mode, Nmesh=Nmesh)

    def paint(self, mode="real", Nmesh=None):
        warnings.warn("the paint method is deprecated from the Public API. Use .compute() instead.", DeprecationWarning)
        return self._paint_XXX(mode=mode, Nmesh=Nmesh)

    def _paint_XXX(self, mode="real", Nmesh=None):
        if mode not in ['real', 'complex']:
            raise ValueError('mode must be "real" or "complex"')
        actions = self.actions + [(mode, )]
        var = self.to_field(mode=actions[0][0])
        attrs = var.attrs if hasattr(var, 'attrs') else {}
        for action in actions:
            if action[0] == 'complex':
                if not isinstance(var, BaseComplexField):
                    var = var.r2c(out=Ellipsis)
            if action[0] == 'real':
                if not isinstance(var, RealField):
                    var = var.c2r(out=Ellipsis)
            if len(action) > 1:
                kwargs = {'func': action[1]}
                if action[2] is not None:
                    kwargs['kind'] = action[2]
                kwargs['out'] = Ellipsis
                var.apply(**kwargs)
        var = var.cast(type=_typestr_to_type(mode), out=var)
        pm = self.pm.reshape(Nmesh=Nmesh)
        if any(pm.Nmesh != self.pm.Nmesh):
            raise NotImplementedError("Fourier-space resampling to a different Nmesh is not on the B200 FFTPower "
                                      "path; pass the desired Nmesh to to_mesh()")
        var.attrs = attrs
        var.attrs.update(self.attrs)
        if self.comm.rank == 0:
            self.logger.info('field: %s painting done' % str(self))
        return var

    def preview(self, axes=None, Nmesh=None, root=0):
        """the real field as a numpy array on the host, optionally summed over the axes not in `axes`
        (base/mesh.py:340-365); the projection runs on the device"""
        field = self.to_field(mode='real')
        return field.preview(Nmesh=Nmesh, axes=axes)

    def save(self, output, dataset='Field', mode='real'):
        raise NotImplementedError("bigfile output is outside the B200 FFTPower path (SURVEY.md §8f); "
                                  "use mesh.preview() / field.numpy() and numpy.save")


class MeshFilter(object):
    """a filter applied to a mesh; subclasses set `kind`, `mode` and implement `filter(x, v)`"""
    kind = None
    mode = None

    def filter(self, x, v):
        raise NotImplementedError

    def __call__(self, x, v):
        return self.filter(x, v)
