"""Minimal labelled-array containers used when xarray is not installed.

The reference keeps its state in an ``xarray.Dataset`` (contrack/contrack.py:56-88).  xarray is not part of this image,
so the ``contrack`` class works on anything that offers the small slice of the Dataset interface the tracking path
touches: ``ds[name]`` -> object with ``.data .dims .attrs``, ``ds[name] = variable``, ``ds.dims`` (name -> size),
``ds.data_vars``, ``len(ds)``.  ``Dataset`` / ``DataArray`` below implement exactly that slice; a real xarray Dataset is
accepted as well.  ``.data`` may be a numpy array (host) or a torch CUDA tensor (device resident).
"""
from __future__ import annotations

import numpy as np


def _shape(data):
    return tuple(int(s) for s in data.shape)


class DataArray(object):
    """data + dimension names + attributes (+ 1-D coordinates when it lives in a Dataset)."""

    def __init__(self, data, dims=None, coords=None, attrs=None, name=None):
        if dims is None:
            dims = tuple('dim_%d' % i for i in range(len(_shape(data))))
        dims = (dims,) if isinstance(dims, str) else tuple(dims)
        if len(dims) != len(_shape(data)):
            raise ValueError('dimensions {} must have the same length as the number of data dimensions, ndim={}'
                             .format(dims, len(_shape(data))))
        self.data = data
        self.dims = dims
        self.attrs = dict(attrs or {})
        self.encoding = {}
        self.name = name
        self.coords = dict(coords or {})

    @property
    def values(self):
        d = self.data
        if isinstance(d, np.ndarray):
            return d
        if hasattr(d, 'detach'):
            return d.detach().cpu().numpy()
        return np.asarray(d)

    @property
    def shape(self):
        return _shape(self.data)

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def ndim(self):
        return len(self.dims)

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        v = self.values
        return v if dtype is None else v.astype(dtype)

    def __getitem__(self, key):
        if isinstance(key, str):                      # coordinate by name, as xarray does
            c = self.coords[key]
            return c if isinstance(c, DataArray) else DataArray(np.asarray(c), (key,), name=key)
        return self.values[key]

    def transpose(self, *dims):
        if not dims:
            dims = self.dims[::-1]
        perm = [self.dims.index(d) for d in dims]
        d = self.data
        data = d.permute(*perm) if hasattr(d, 'permute') else np.transpose(d, perm)
        return DataArray(data, dims, self.coords, self.attrs, self.name)

    def __repr__(self):
        return '<contrack_b200.DataArray {} {} {}>'.format(self.name or '', dict(zip(self.dims, self.shape)), self.dtype)


def Variable(dims, data, attrs=None):
    """Signature of xarray.Variable as the reference uses it (contrack/contrack.py:568, 776)."""
    return DataArray(data, dims, attrs=attrs)


class _Dims(dict):
    """name -> size; iterates in sorted order like the Dataset.dims of the xarray versions the reference targets
    (tests/test_contrack.py:57-58 expects ['latitude', 'longitude', 'time'])."""

    def __iter__(self):
        return iter(sorted(dict.keys(self)))

    def keys(self):
        return sorted(dict.keys(self))


class Dataset(object):
    def __init__(self, data_vars=None, coords=None, attrs=None):
        self._coords = {}
        self._vars = {}
        self.attrs = dict(attrs or {})
        for name, c in (coords or {}).items():
            self._coords[name] = self._as_array(name, c, default_dims=(name,))
        for name, v in (data_vars or {}).items():
            self[name] = v

    @staticmethod
    def _as_array(name, v, default_dims=None):
        if isinstance(v, DataArray):
            out = DataArray(v.data, v.dims, None, v.attrs, name)
            out.encoding = dict(getattr(v, 'encoding', {}))
            return out
        if hasattr(v, 'dims') and hasattr(v, 'data'):          # xarray object
            return DataArray(v.data, v.dims, None, getattr(v, 'attrs', {}), name)
        if isinstance(v, tuple):
            dims, data = v[0], v[1]
            attrs = v[2] if len(v) > 2 else None
            return DataArray(data, dims, None, attrs, name)
        return DataArray(np.asarray(v), default_dims, None, None, name)

    # --- the slice of the xarray interface the class uses ---
    @property
    def dims(self):
        d = _Dims()
        for a in list(self._coords.values()) + list(self._vars.values()):
            for n, s in zip(a.dims, a.shape):
                d[n] = s
        return d

    @property
    def sizes(self):
        return self.dims

    @property
    def data_vars(self):
        return dict(self._vars)

    @property
    def coords(self):
        return dict(self._coords)

    @property
    def variables(self):
        d = dict(self._coords)
        d.update(self._vars)
        return d

    def __len__(self):
        return len(self._vars)

    def __contains__(self, name):
        return name in self._vars or name in self._coords

    def __iter__(self):
        return iter(self._vars)

    def __getitem__(self, name):
        if name in self._vars:
            a = self._vars[name]
        elif name in self._coords:
            a = self._coords[name]
        else:
            raise KeyError(name)
        a.coords = {d: self._coords[d] for d in a.dims if d in self._coords and d != name}
        return a

    def __setitem__(self, name, value):
        a = self._as_array(name, value)
        sizes = self.dims
        for n, s in zip(a.dims, a.shape):
            if n in sizes and sizes[n] != s:
                raise ValueError('conflicting sizes for dimension {!r}: length {} on {!r} and length {} on the dataset'
                                 .format(n, s, name, sizes[n]))
        self._vars[name] = a

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __repr__(self):
        return '<contrack_b200.Dataset dims={} data_vars={}>'.format(dict(self.dims), list(self._vars))


def load_npz(filename):
    """A Dataset from an .npz archive: 1-D arrays named like one of their own... every 1-D array is a coordinate, every
    N-D array a data variable whose dimensions are matched to coordinates by length, preferring the conventional order
    (time, latitude, longitude)."""
    z = np.load(filename, allow_pickle=False)
    coords = {k: z[k] for k in z.files if z[k].ndim == 1}
    ds = Dataset(coords={k: v for k, v in coords.items()})
    for k in z.files:
        a = z[k]
        if a.ndim <= 1:
            continue
        names = []
        pool = [n for n in ('time', 'latitude', 'lat', 'longitude', 'lon') if n in coords] + \
               [n for n in coords if n not in ('time', 'latitude', 'lat', 'longitude', 'lon')]
        for ax, s in enumerate(a.shape):
            pick = next((n for n in pool if len(coords[n]) == s and n not in names), None)
            names.append(pick or '%s_dim_%d' % (k, ax))
        ds[k] = DataArray(a, names)
    return ds
