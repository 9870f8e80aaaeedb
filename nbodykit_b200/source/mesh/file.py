"""
On-disk meshes: `MeshSource.save(output, dataset, mode)` and `FileMesh(path, dataset)` -- the roles of
nbodykit/base/mesh.py:367-412 (`save`) and nbodykit/source/mesh/bigfile.py (`BigFileMesh`).  The reference writes
bigfile columns; bigfile is not available here, so the container is a directory:

    <output>/<dataset>/attrs.json            Nmesh, BoxSize, dtype, mode ('real' | 'complex'), the mesh attrs,
                                             and one entry per written block: file, first plane, number of planes
    <output>/<dataset>/block-<rank>.npy      this rank's planes of the field in its NATURAL axis order:
                                             real    [x planes][Ny][Nz]
                                             complex [x planes][Ny][Nz/2+1]   (written from the untransposed view)

Every rank writes the planes it owns; a reader with any number of ranks picks the planes it needs out of the blocks
(memory-mapped), so files written on P GPUs load on Q.  The field crosses PCIe once in each direction.
"""
import json
import os

import numpy
import torch

from ... import CurrentMPIComm
from ...base.mesh import MeshSource
from ...pmesh.pm import ComplexField, RealField
from ...utils import JSONDecoder, JSONEncoder


def _natural_complex_planes(field):
    """(first x plane, planes) of this rank's share of a ComplexField in [x][y][z] order.  With x-slab real fields the
    complex field of P > 1 ranks is stored transposed ([y_n][Nx][Nzc]); its share is written as y planes instead."""
    pm = field.pm
    if not pm.transposed:
        return 'x', 0, field.value
    return 'y', pm.y_start, field.value          # [y_n][Nx][Nzc]: axis 0 is y


def save_mesh(source, output, dataset='Field', mode='real'):
    if mode not in ('real', 'complex'):
        raise ValueError('mode must be "real" or "complex"')
    comm = source.comm
    field = source.compute(mode=mode)
    pm = field.pm
    root = os.path.join(output, dataset)
    if comm.rank == 0:
        os.makedirs(root, exist_ok=True)
    comm.barrier()
    if mode == 'real':
        axis, start, data = 'x', pm.x_start, field.value
    else:
        axis, start, data = _natural_complex_planes(field)
    fn = "block-%d.npy" % comm.rank
    numpy.save(os.path.join(root, fn), data.cpu().numpy())
    entries = comm.allgather({'file': fn, 'axis': axis, 'start': int(start), 'count': int(data.shape[0])})
    if comm.rank == 0:
        meta = {'Nmesh': pm.Nmesh, 'BoxSize': pm.BoxSize, 'dtype': numpy.dtype(pm.dtype).str, 'mode': mode,
                'compressed': bool(not pm.cplx), 'blocks': entries, 'attrs': dict(field.attrs)}
        with open(os.path.join(root, 'attrs.json'), 'w') as ff:
            json.dump(meta, ff, cls=JSONEncoder)
    comm.barrier()
    return root


class FileMesh(MeshSource):
    """a mesh read from a directory written by `MeshSource.save` (the reference's BigFileMesh)"""

    def __repr__(self):
        return "FileMesh(path=%s, dataset=%s)" % (self.path, self.dataset)

    @CurrentMPIComm.enable
    def __init__(self, path, dataset, comm=None, **kwargs):
        self.path, self.dataset = path, dataset
        root = os.path.join(path, dataset)
        if comm.rank == 0:
            with open(os.path.join(root, 'attrs.json')) as ff:
                meta = json.load(ff, cls=JSONDecoder)
        else:
            meta = None
        meta = comm.bcast(meta)
        self._meta = meta
        self._root = root
        self.attrs.update(meta.get('attrs', {}))
        self.attrs.update(kwargs)
        MeshSource.__init__(self, comm, numpy.asarray(meta['Nmesh'], dtype='i8'), numpy.asarray(meta['BoxSize'], dtype='f8'),
                            numpy.dtype(meta['dtype']))

    def _read(self, axis, start, count, shape_tail, dtype):
        """planes [start, start+count) along `axis` ('x' | 'y') assembled from the blocks that hold them"""
        out = numpy.empty((count,) + tuple(shape_tail), dtype=dtype)
        filled = 0
        for b in self._meta['blocks']:
            lo, hi = max(start, b['start']), min(start + count, b['start'] + b['count'])
            if b['axis'] != axis or hi <= lo:
                continue
            arr = numpy.load(os.path.join(self._root, b['file']), mmap_mode='r')
            out[lo - start:hi - start] = arr[lo - b['start']:hi - b['start']]
            filled += hi - lo
        return out if filled == count else None

    def to_real_field(self):
        if self._meta['mode'] != 'real':
            return NotImplemented
        pm = self.pm
        N = [int(v) for v in pm.Nmesh]
        data = self._read('x', pm.x_start, pm.x_n, (N[1], N[2]), numpy.dtype('f%d' % (4 if pm.typestr == 'f4' else 8)))
        if data is None:
            raise IOError("%s does not hold x planes [%d, %d)" % (self._root, pm.x_start, pm.x_start + pm.x_n))
        f = RealField(pm)
        f.value.copy_(torch.from_numpy(data))
        f.attrs = dict(self.attrs)
        return f

    def to_complex_field(self):
        if self._meta['mode'] != 'complex':
            return NotImplemented
        pm = self.pm
        N = [int(v) for v in pm.Nmesh]
        cdt = numpy.dtype('c8' if pm.typestr == 'f4' else 'c16')
        f = ComplexField(pm)
        if not pm.transposed:
            data = self._read('x', 0, N[0], (N[1], pm.Nzc), cdt)
            if data is None:       # written transposed by several ranks: y planes of [y][x][z]
                yx = self._read('y', 0, N[1], (N[0], pm.Nzc), cdt)
                data = None if yx is None else numpy.ascontiguousarray(yx.transpose(1, 0, 2))
        else:
            data = self._read('y', pm.y_start, pm.y_n, (N[0], pm.Nzc), cdt)
            if data is None:
                full = self._read('x', 0, N[0], (N[1], pm.Nzc), cdt)
                data = None if full is None else numpy.ascontiguousarray(full[:, pm.y_start:pm.y_start + pm.y_n].transpose(1, 0, 2))
        if data is None:
            raise IOError("%s does not hold the complex planes this rank needs" % self._root)
        f.value.copy_(torch.from_numpy(data))
        f.attrs = dict(self.attrs)
        return f
