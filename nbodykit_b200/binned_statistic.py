"""
BinnedStatistic -- result container of FFTPower / ConvolvedFFTPower (API of
nbodykit/binned_statistic.py:60-955): named dimensions with bin edges / centres, a structured
array of variables defined on the bin grid, a mask of non-finite bins, metadata.

An independent implementation: the state layout (`__getstate__` / `from_state`, JSON files) and the
behaviour of indexing, `sel`, `take`, `squeeze`, `average`, `reindex` follow the reference so that
saved results and downstream scripts are interchangeable.
"""
import copy as _copy

import numpy


def _block_reduce(arr, axis, factor, how, weights=None):
    """reduce groups of `factor` consecutive entries along `axis` (mean / sum, optionally weighted)"""
    shape = list(arr.shape)
    n = shape[axis] // factor
    new = shape[:axis] + [n, factor] + shape[axis + 1:]
    a = arr.reshape(new)
    if weights is not None:
        w = weights.reshape(new)
        with numpy.errstate(invalid="ignore", divide="ignore"):
            return numpy.nansum(a * w, axis=axis + 1) / numpy.sum(w, axis=axis + 1)
    if how == "sum":
        return numpy.nansum(a, axis=axis + 1)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)   # all-NaN groups -> NaN, as nanmean does
        return numpy.nanmean(a, axis=axis + 1)


class BinnedStatistic(object):
    """
    Parameters
    ----------
    dims : list of str
        names of the binning dimensions
    edges : list of arrays
        bin edges per dimension
    data : structured ndarray, shape = tuple(len(e)-1 for e in edges)
        the variables
    fields_to_sum : list of str
        variables that are summed (not averaged) when re-binning
    coords : list of arrays or None
        explicit bin centres (default: mid-points of the edges)
    **kwargs : stored in :attr:`attrs`
    """

    def __init__(self, dims, edges, data, fields_to_sum=[], coords=None, **kwargs):
        if len(dims) != len(edges):
            raise ValueError("size mismatch between specified `dims` and `edges`")
        if not isinstance(data, numpy.ndarray) or data.dtype.names is None:
            raise TypeError("'data' should be a structured numpy array")
        shape = tuple(len(e) - 1 for e in edges)
        if data.shape != shape:
            raise ValueError("`edges` imply data shape of %s, but data has shape %s" % (shape, data.shape))
        self.dims = list(dims)
        self.edges = {d: numpy.asarray(e) for d, e in zip(self.dims, edges)}
        self.coords = {}
        for i, d in enumerate(self.dims):
            if coords is not None and coords[i] is not None:
                self.coords[d] = numpy.copy(coords[i])
            else:
                e = numpy.asarray(edges[i])
                self.coords[d] = 0.5 * (e[1:] + e[:-1])
        self.data = data.copy()
        self.mask = self._nonfinite(self.data)
        self._fields_to_sum = fields_to_sum
        self.attrs = dict(kwargs)

    # ------------------------------------------------------------------ state / construction helpers
    @staticmethod
    def _nonfinite(data):
        mask = numpy.zeros(data.shape, dtype=bool)
        for name in data.dtype.names:
            mask |= ~numpy.isfinite(data[name])
        return mask

    @classmethod
    def from_state(kls, state):
        obj = kls(dims=state['dims'], edges=state['edges'], coords=state['coords'], data=state['data'])
        obj.attrs.update(state['attrs'])
        return obj

    def __getstate__(self):
        return dict(dims=self.dims,
                    edges=[self.edges[d] for d in self.dims],
                    coords=[self.coords[d] for d in self.dims],
                    data=self.data,
                    attrs=self.attrs)

    def __setstate__(self, state):
        other = self.from_state(state)
        self.__dict__.update(other.__dict__)

    @classmethod
    def _rebuild(cls, data, mask, dims, edges, coords, attrs, fields_to_sum):
        """new instance without re-deriving the mask (slices must carry their parent's mask)"""
        obj = object.__new__(cls)
        obj.dims = list(dims)
        obj.edges = dict(edges)
        obj.coords = dict(coords)
        obj.attrs = dict(attrs)
        obj._fields_to_sum = list(fields_to_sum)
        shape = tuple(len(obj.coords[d]) for d in obj.dims)
        if data.shape != shape:
            try:
                data = data.reshape(shape)
                mask = mask.reshape(shape)
            except Exception:
                raise ValueError("shape mismatch between data and coordinates")
        obj.data = data
        obj.mask = mask
        return obj

    def _subset(self, data, mask, indices):
        """instance holding bins `indices[i]` (sorted, contiguous runs keep exact edges) of each dimension"""
        edges, coords = {}, {}
        for i, d in enumerate(self.dims):
            idx = list(indices[i])
            sel = idx + [idx[-1] + 1] if len(idx) else [0]
            edges[d] = self.edges[d][sel]
            coords[d] = 0.5 * (edges[d][1:] + edges[d][:-1])
        return self._rebuild(data, mask, self.dims, edges, coords, self.attrs, self._fields_to_sum)

    # ------------------------------------------------------------------ basic protocol
    @property
    def shape(self):
        return tuple(len(self.coords[d]) for d in self.dims)

    @property
    def variables(self):
        return list(self.data.dtype.names)

    def __str__(self):
        dims = ", ".join("%s: %d" % (d, n) for d, n in zip(self.dims, self.shape))
        return "<%s: dims: (%s), variables: %s>" % (self.__class__.__name__, dims,
                                                     str(tuple(self.variables)) if len(self.variables) < 5
                                                     else "%d total" % len(self.variables))

    __repr__ = __str__

    def __iter__(self):
        return iter(self.variables)

    def __contains__(self, key):
        return key in self.variables

    def __setitem__(self, key, data):
        """add (or overwrite) a variable"""
        data = numpy.asarray(data)
        if data.shape != self.data.shape:
            raise ValueError("data to be added must have shape %s" % str(self.data.shape))
        keep = [n for n in self.data.dtype.names if n != key]
        descr = [(n, self.data.dtype[n].str) for n in keep] + [(key, data.dtype.str)]
        new = numpy.zeros(self.data.shape, dtype=numpy.dtype(descr))
        for n in keep:
            new[n] = self.data[n]
        new[key] = data
        self.mask = self.mask | ~numpy.isfinite(new[key])
        self.data = new

    def __getitem__(self, key):
        """string -> variable array; list of strings -> sub-statistic; ints/slices/lists -> sliced statistic
        (integer indices squeeze their dimension)"""
        if isinstance(key, str):
            if key in self.variables:
                return self.data[key]
            raise KeyError("`%s` is not a valid variable name" % key)
        nd = len(self.dims)
        indices = [list(range(n)) for n in self.shape]
        if isinstance(key, (list, tuple)) and len(key) and all(isinstance(x, str) for x in key):
            bad = [k for k in key if k not in self.variables]
            if bad:
                raise KeyError("cannot slice variables -- invalid names: (%s)" % ", ".join("'%s'" % k for k in bad))
            return self._subset(self.data[list(key)], self.mask.copy(), indices)
        key_ = key
        if isinstance(key, (slice, int, numpy.integer)) or (isinstance(key, list) and all(isinstance(x, (int, numpy.integer)) for x in key)):
            key_ = [key]
        squeezed = []
        for i, sub in enumerate(key_):
            if i >= nd:
                raise IndexError("too many indices for BinnedStatistic; note that ndim = %d" % nd)
            if isinstance(sub, (int, numpy.integer)):
                indices[i] = [int(sub) % self.shape[i] if sub < 0 else int(sub)]
                squeezed.append(self.dims[i])
            elif isinstance(sub, list):
                indices[i] = sub
            elif isinstance(sub, slice):
                indices[i] = list(range(*sub.indices(self.shape[i])))
        if len(squeezed) == nd:
            raise IndexError("cannot return object with all remaining dimensions squeezed")
        try:
            out = self._subset(self.data[key], self.mask[key], indices)
            for d in squeezed:
                out = out.squeeze(d)
            return out
        except ValueError:
            raise IndexError("this type of slicing not implemented")

    def _get_index(self, dim, val, method=None):
        index = self.coords[dim]
        if method == 'nearest':
            return int(numpy.abs(index - val).argmin())
        try:
            return list(index).index(val)
        except Exception as e:
            raise IndexError("error converting '%s' index; try setting `method = 'nearest'`: %s" % (dim, str(e)))

    # ------------------------------------------------------------------ IO
    def to_json(self, filename):
        import json
        from .utils import JSONEncoder
        with open(filename, 'w') as ff:
            json.dump(self.__getstate__(), ff, cls=JSONEncoder)

    @classmethod
    def from_json(cls, filename, key='data', dims=None, edges=None, **kwargs):
        """load from a JSON file written by `to_json` (or by FFTPower.save with key='power'/'poles')"""
        import json
        from .utils import JSONDecoder
        with open(filename, 'r') as ff:
            state = json.load(ff, cls=JSONDecoder)
        if key not in state:
            raise ValueError("JSON file does not contain a key '%s'" % key)
        data = state[key]
        if isinstance(data, dict) and 'dims' in data and 'edges' in data and 'data' in data:
            obj = cls.from_state(data)          # nested state (FFTPower.save)
            obj.attrs.update(kwargs)
            return obj
        if dims is None:
            dims = state.get('dims', None)
        if dims is None:
            raise ValueError("no `dims` in JSON file; please specify as keyword argument")
        if edges is None:
            edges = state.get('edges', None)
        if edges is None:
            raise ValueError("no `edges` in JSON file; please specify as keyword argument")
        attrs = dict(state.get('attrs', {}))
        attrs.update(kwargs)
        coords = state.get('coords', None)
        return cls(dims, edges, data, coords=coords, **attrs)

    # ------------------------------------------------------------------ copies / renames
    def copy(self, cls=None):
        cls = cls or self.__class__
        if not issubclass(cls, BinnedStatistic):
            raise TypeError("The cls argument must be a subclass of BinnedStatistic")
        return cls._rebuild(self.data.copy(), self.mask.copy(), self.dims,
                            {d: e.copy() for d, e in self.edges.items()},
                            {d: c.copy() for d, c in self.coords.items()}, self.attrs, self._fields_to_sum)

    def rename_variable(self, old_name, new_name):
        if old_name not in self.variables:
            raise ValueError("`%s` is not an existing variable name" % old_name)
        dt = _copy.deepcopy(self.data.dtype)
        names = list(dt.names)
        names[names.index(old_name)] = new_name
        dt.names = names
        self.data.dtype = dt

    # ------------------------------------------------------------------ coordinate-based selection
    def sel(self, method=None, **indexers):
        """select by coordinate value: scalar (squeezes the dimension), list, or slice(start, stop)"""
        indices, squeezed = {}, []
        for dim, key in indexers.items():
            if isinstance(key, list):
                indices[dim] = [self._get_index(dim, k, method=method) for k in key]
            elif isinstance(key, slice):
                lo = self._get_index(dim, key.start, method=method)
                hi = self._get_index(dim, key.stop, method=method)
                n = self.shape[self.dims.index(dim)]
                indices[dim] = list(range(*slice(lo, hi).indices(n)))
            elif not numpy.isscalar(key):
                raise IndexError("please index using a list, slice, or scalar value")
            else:
                indices[dim] = [self._get_index(dim, key, method=method)]
                squeezed.append(dim)
        if len(squeezed) == len(self.dims):
            raise IndexError("cannot return object with all remaining dimensions squeezed")
        out = self.take(**indices)
        for dim in squeezed:
            out = out.squeeze(dim)
        return out

    def take(self, *masks, **indices):
        """keep the bins selected by boolean `masks` (true everywhere along the other axes) and by
        per-dimension index lists / boolean vectors"""
        full = numpy.ones(self.shape, dtype='?')
        for m in masks:
            full = full & m
        keep = []
        for i in range(len(self.dims)):
            other = tuple(a for a in range(len(self.dims)) if a != i)
            keep.append(full.all(axis=other) if other else full.copy())
        for dim, index in indices.items():
            i = self.dims.index(dim)
            if isinstance(index, numpy.ndarray) and index.dtype == numpy.dtype('?'):
                assert index.ndim == 1
                keep[i] &= index
            else:
                m = numpy.zeros(self.shape[i], dtype='?')
                m.put(index, True)
                keep[i] &= m
        idx = [k.nonzero()[0] for k in keep]
        data, mask = self.data.copy(), self.mask.copy()
        for i, ii in enumerate(idx):
            data = numpy.take(data, ii, axis=i)
            mask = numpy.take(mask, ii, axis=i)
        return self._subset(data, mask, idx)

    def squeeze(self, dim=None):
        """drop a length-one dimension"""
        if dim is None:
            cand = [k for k in self.dims if len(self.coords[k]) == 1]
            if not cand:
                raise ValueError("no available dimensions with length one to squeeze")
            if len(cand) > 1:
                raise ValueError("multiple dimensions available to squeeze -- please specify")
            dim = cand[0]
        else:
            if dim not in self.dims:
                raise ValueError("`%s` is not a valid dimension name" % dim)
            if len(self.coords[dim]) != 1:
                raise ValueError("the `%s` dimension must have length one to squeeze" % dim)
        i = self.dims.index(dim)
        dims = [d for d in self.dims if d != dim]
        if not dims:
            raise ValueError("cannot squeeze the only remaining axis")
        edges = {d: self.edges[d].copy() for d in dims}
        coords = {d: self.coords[d].copy() for d in dims}
        return self._rebuild(self.data.squeeze(axis=i).copy(), self.mask.squeeze(axis=i).copy(), dims, edges, coords,
                             self.attrs, self._fields_to_sum)

    def average(self, dim, **kwargs):
        """average every variable over `dim` (removes the dimension)"""
        spacing = (self.edges[dim][-1] - self.edges[dim][0])
        out = self.reindex(dim, spacing, **kwargs)
        return out.sel(**{dim: out.coords[dim][0]})

    def reindex(self, dim, spacing, weights=None, force=True, return_spacing=False, fields_to_sum=[]):
        """re-bin `dim` to bins an integer factor wider; variables are averaged (NaN-aware), optionally
        weighted; those in `fields_to_sum` are summed"""
        i = self.dims.index(dim)
        sum_fields = list(fields_to_sum) + list(self._fields_to_sum)
        old = numpy.diff(self.coords[dim])
        old_spacing = old[0]
        factor = int(numpy.round(spacing / old_spacing))
        if not factor:
            raise ValueError("new spacing must be smaller than original spacing of %.2e" % old_spacing)
        if factor == 1:
            raise ValueError("closest binning size to input spacing is the same as current binning")
        if not numpy.allclose(old_spacing * factor, spacing) and not force:
            raise ValueError("if `force = False`, new bin spacing must be an integral factor smaller than original")
        data = self.data.copy()
        if isinstance(weights, str):
            if weights not in self.variables:
                raise ValueError("cannot weight by `%s`; no such column" % weights)
            weights = self.data[weights]
        edges = self.edges[dim]
        leftover = self.shape[i] % factor
        if leftover and not force:
            raise ValueError("cannot re-bin because they are %d extra bins, using spacing = %.2e"
                             % (leftover, old_spacing * factor))
        if leftover:
            sl = [slice(None)] * len(self.dims)
            sl[i] = slice(None, -leftover)
            data = data[tuple(sl)]
            if weights is not None:
                weights = weights[tuple(sl)]
            edges = edges[:-leftover]
        n_new = data.shape[i] // factor
        new_edges = numpy.linspace(edges[0], edges[-1], n_new + 1)
        new_shape = list(data.shape)
        new_shape[i] = n_new
        new_data = numpy.empty(new_shape, dtype=self.data.dtype)
        for name in self.variables:
            if name in sum_fields:
                new_data[name] = _block_reduce(data[name], i, factor, "sum")
            elif weights is not None:
                new_data[name] = _block_reduce(data[name], i, factor, "mean", weights=weights)
            else:
                new_data[name] = _block_reduce(data[name], i, factor, "mean")
        edges_d = {d: e.copy() for d, e in self.edges.items()}
        coords_d = {d: c.copy() for d, c in self.coords.items()}
        edges_d[dim] = new_edges
        coords_d[dim] = 0.5 * (new_edges[1:] + new_edges[:-1])
        out = self._rebuild(new_data, self._nonfinite(new_data), self.dims, edges_d, coords_d, self.attrs,
                            self._fields_to_sum)
        return (out, spacing) if return_spacing else out
