"""TEST INFRASTRUCTURE ONLY -- a stand-in for the slice of xarray that steidani/ConTrack's contrack.py touches.

The build container has no xarray wheel, so `import contrack` of the UNMODIFIED reference fails at contrack.py:19.
tests/golden/make_reference_golden.py puts this directory on sys.path *in front of* /root/reference and then imports and
runs the reference's own source: every arithmetic line of run_contrack / run_lifecycle (scipy.ndimage.label, find_objects,
the date-line loops, the overlap loop, the persistence loop, center_of_mass) is the reference's code executing; only the
labelled-array plumbing below (indexing by name, transpose by name, isel, roll, where) is provided here, each method
documented with the xarray behaviour it stands for.  The product (contrack_b200/) never imports this module.

groupby / rolling / mean (used by calc_clim / calc_anom, contrack.py:483-489, 568-570) delegate to pandas, which is what
xarray's rolling falls back on semantically (centre=True, min_periods=window, NaN-skipping group means); that part is an
INTERPRETATION of xarray, not xarray itself: calc_clim / calc_anom vectors made through it are marked "pandas-backed".
"""
import numpy as np
import pandas as pd

__version__ = '0.0-shim'


def _is_da(x):
    return isinstance(x, DataArray)


class _DtAccessor(object):
    def __init__(self, da):
        self._da = da

    def strftime(self, fmt):
        v = np.asarray(self._da.data)
        idx = pd.DatetimeIndex(v.reshape(-1))
        out = np.array(idx.strftime(fmt), dtype=object).reshape(v.shape)
        return DataArray(out, self._da.dims, self._da.coords, name=self._da.name)

    @property
    def dayofyear(self):
        v = np.asarray(self._da.data)
        return DataArray(np.asarray(pd.DatetimeIndex(v.reshape(-1)).dayofyear).reshape(v.shape), self._da.dims,
                         name='dayofyear')


class DataArray(object):
    """xarray.DataArray: n-d data + dimension names + 1-D coordinates keyed by dimension name."""

    def __init__(self, data, dims=None, coords=None, attrs=None, name=None):
        data = np.asarray(data)
        if dims is None:
            dims = tuple('dim_%d' % i for i in range(data.ndim))
        dims = (dims,) if isinstance(dims, str) else tuple(dims)
        assert len(dims) == data.ndim, (dims, data.shape)
        self.data = data
        self.dims = dims
        self.coords = dict(coords or {})
        self.attrs = dict(attrs or {})
        self.encoding = {}
        self.name = name

    values = property(lambda self: self.data)
    shape = property(lambda self: self.data.shape)
    dtype = property(lambda self: self.data.dtype)
    ndim = property(lambda self: self.data.ndim)

    def __len__(self):
        return self.data.shape[0]

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)

    def _coord(self, name):
        c = self.coords[name]
        return c if _is_da(c) else DataArray(np.asarray(c), (name,), name=name)

    def __getitem__(self, key):
        if isinstance(key, str):                       # da['latitude'] -> coordinate
            return self._coord(key)
        if not isinstance(key, tuple):
            key = (key,)
        out = self.data[key]
        dims, coords = [], {}
        for d, k in zip(self.dims, key + (slice(None),) * (self.ndim - len(key))):
            if isinstance(k, (int, np.integer)):
                continue
            dims.append(d)
            if d in self.coords:
                coords[d] = DataArray(np.asarray(self._coord(d).data)[k], (d,), name=d)
        return DataArray(out, dims, coords, self.attrs, self.name)

    def __contains__(self, v):
        return bool((self.data == v).any())

    def _cmp(self, other, op):
        o = other.data if _is_da(other) else other
        return DataArray(op(self.data, o), self.dims, self.coords, name=self.name)

    def __ge__(self, o): return self._cmp(o, np.greater_equal)
    def __le__(self, o): return self._cmp(o, np.less_equal)
    def __gt__(self, o): return self._cmp(o, np.greater)
    def __lt__(self, o): return self._cmp(o, np.less)

    def _arith(self, other, op):
        if _is_da(other):
            assert other.dims == self.dims or other.ndim == 0, 'shim: only aligned arithmetic'
            other = other.data
        return DataArray(op(self.data, other), self.dims, self.coords, self.attrs, self.name)

    def __sub__(self, o): return self._arith(o, np.subtract)
    def __add__(self, o): return self._arith(o, np.add)
    def __mul__(self, o): return self._arith(o, np.multiply)
    def __truediv__(self, o): return self._arith(o, np.divide)

    def transpose(self, *dims):
        """DataArray.transpose(*dims): reorder by dimension NAME."""
        perm = [self.dims.index(d) for d in dims]
        return DataArray(np.transpose(self.data, perm), dims, self.coords, self.attrs, self.name)

    def isel(self, **idx):
        """positional indexing by dimension name; integer indexers drop the dimension."""
        key = tuple(idx.get(d, slice(None)) for d in self.dims)
        return self[key]

    def roll(self, shifts=None, roll_coords=False, **kw):
        """np.roll along a named dimension, coordinates rolled too when roll_coords=True."""
        shifts = dict(shifts or {}, **kw)
        data, coords = self.data, dict(self.coords)
        for d, s in shifts.items():
            data = np.roll(data, int(s), axis=self.dims.index(d))
            if roll_coords and d in coords:
                coords[d] = DataArray(np.roll(np.asarray(self._coord(d).data), int(s)), (d,), name=d)
        return DataArray(data, self.dims, coords, self.attrs, self.name)

    def to_index(self):
        """pandas Index in xarray; here a datetime64 ndarray, whose `[1:] - [:-1]` `.astype('timedelta64[h]')` is what
        contrack.py:334-339 needs (pandas 3 refuses that astype on a TimedeltaIndex; the value only feeds `_dtime`,
        which the tracking path never reads)."""
        return np.asarray(self.data)

    @property
    def dt(self):
        return _DtAccessor(self)

    def reset_coords(self, names=None, drop=False):
        return self

    # ---- pandas-backed reductions for calc_clim / calc_anom (interpretation of xarray, see module docstring) ----------
    def groupby(self, spec):
        dim, field = spec.split('.')
        t = pd.DatetimeIndex(np.asarray(self._coord(dim).data))
        return _GroupBy(self, dim, field, np.asarray(getattr(t, field)))

    def rolling(self, center=False, min_periods=None, **win):
        (dim, w), = win.items()
        return _Rolling(self, dim, int(w), center, min_periods)

    def mean(self, dim=None, skipna=True):
        ax = self.dims.index(dim)
        with np.errstate(invalid='ignore'):
            out = np.nanmean(self.data, axis=ax) if skipna else self.data.mean(axis=ax)
        dims = tuple(d for d in self.dims if d != dim)
        return DataArray(out.astype(self.data.dtype), dims, {d: c for d, c in self.coords.items() if d != dim},
                         self.attrs, self.name)

    def reindex(self, method=None, **indexers):
        """DataArray.reindex(dim=new_coord, method='nearest'): pandas Index.get_indexer per dimension (what xarray does)."""
        data, coords = self.data, dict(self.coords)
        for d, new in indexers.items():
            new = np.asarray(new.data if _is_da(new) else new)
            idx = pd.Index(np.asarray(self._coord(d).data)).get_indexer(pd.Index(new), method=method)
            assert (idx >= 0).all(), 'shim: reindex without fill values only'
            data = np.take(data, idx, axis=self.dims.index(d))
            coords[d] = DataArray(new, (d,), name=d)
        return DataArray(data, self.dims, coords, self.attrs, self.name)

    def fillna(self, other):
        o = other.data if _is_da(other) else other
        return DataArray(np.where(np.isnan(self.data), o, self.data).astype(self.data.dtype), self.dims, self.coords,
                         self.attrs, self.name)

    def __repr__(self):
        return '<xr_shim.DataArray %s %s %s>' % (self.name, dict(zip(self.dims, self.shape)), self.dtype)


class _GroupBy(object):
    def __init__(self, da, dim, field, keys):
        self.da, self.dim, self.field, self.keys = da, dim, field, keys

    def mean(self, dim=None):
        """group mean over the grouped dimension, NaN-skipping, result dimension named after the field, sorted keys."""
        ax = self.da.dims.index(self.dim)
        uniq = np.unique(self.keys)
        x = np.moveaxis(self.da.data, ax, 0)
        out = np.empty((len(uniq),) + x.shape[1:], x.dtype)
        for i, k in enumerate(uniq):
            with np.errstate(invalid='ignore'):
                out[i] = np.nanmean(x[self.keys == k], axis=0)
        out = np.moveaxis(out, 0, ax)
        dims = tuple(self.field if d == self.dim else d for d in self.da.dims)
        coords = {d: c for d, c in self.da.coords.items() if d != self.dim}
        coords[self.field] = DataArray(uniq, (self.field,), name=self.field)
        return DataArray(out, dims, coords, self.da.attrs, self.da.name)

    def __sub__(self, other):
        """grouped - per-group array: every element minus the entry of its group."""
        ax_o = other.dims.index(self.field)
        pos = np.searchsorted(np.asarray(other._coord(self.field).data), self.keys)
        o = np.moveaxis(other.data, ax_o, 0)[pos]
        ax = self.da.dims.index(self.dim)
        x = np.moveaxis(self.da.data, ax, 0)
        out = np.moveaxis(x - o, 0, ax)
        coords = dict(self.da.coords)
        return DataArray(out, self.da.dims, coords, self.da.attrs, self.da.name)


class _Rolling(object):
    def __init__(self, da, dim, w, center, min_periods):
        self.da, self.dim, self.w, self.center, self.min_periods = da, dim, w, center, min_periods

    def mean(self):
        ax = self.da.dims.index(self.dim)
        x = np.moveaxis(self.da.data, ax, 0)
        flat = x.reshape(x.shape[0], -1)
        r = pd.DataFrame(flat.astype(np.float64)).rolling(self.w, center=self.center,
                                                          min_periods=self.min_periods).mean().to_numpy()
        out = np.moveaxis(r.reshape(x.shape).astype(self.da.data.dtype), 0, ax)
        return DataArray(out, self.da.dims, self.da.coords, self.da.attrs, self.da.name)


def Variable(dims, data, attrs=None):
    """xarray.Variable(dims, data, attrs) as used at contrack.py:417, 568, 776."""
    return DataArray(data.data if _is_da(data) else data, dims, attrs=attrs)


def where(cond, x, y):
    """xr.where: DataArray in -> DataArray out (contrack.py:650-671); ndarray in -> ndarray out (contrack.py:747)."""
    if _is_da(cond):
        return DataArray(np.where(cond.data, x, y), cond.dims, cond.coords, name=cond.name)
    return np.where(cond, x, y)


class _Dims(dict):
    """Dataset.dims: name -> size, iterating in sorted order (the reference's tests/test_contrack.py:57-58)."""

    def __iter__(self):
        return iter(sorted(dict.keys(self)))


class Dataset(object):
    def __init__(self, data_vars=None, coords=None, attrs=None):
        self._coords = {k: (v if _is_da(v) else DataArray(np.asarray(v), (k,), name=k)) for k, v in (coords or {}).items()}
        self._vars = {}
        self.attrs = dict(attrs or {})
        for k, v in (data_vars or {}).items():
            self[k] = v

    @property
    def dims(self):
        d = _Dims()
        for a in list(self._coords.values()) + list(self._vars.values()):
            for n, s in zip(a.dims, a.shape):
                d[n] = s
        return d

    @property
    def data_vars(self):
        return dict(self._vars)

    @property
    def variables(self):
        return dict(self._coords, **self._vars)

    def __len__(self):
        return len(self._vars)

    def __getitem__(self, name):
        a = self._vars[name] if name in self._vars else self._coords[name]
        out = DataArray(a.data, a.dims, {d: self._coords[d] for d in a.dims if d in self._coords}, a.attrs, name)
        out.encoding = a.encoding
        return out

    def __setitem__(self, name, v):
        if isinstance(v, tuple):
            v = DataArray(v[1], v[0], attrs=v[2] if len(v) > 2 else None)
        assert _is_da(v)
        for n, s in zip(v.dims, v.shape):
            if n in self.dims and self.dims[n] != s:
                raise ValueError('conflicting sizes for dimension %r' % n)
        self._vars[name] = DataArray(v.data, v.dims, None, v.attrs, name)

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)


class _Core(object):
    class dataset(object):
        pass


core = _Core()
core.dataset.Dataset = Dataset


def open_dataset(*a, **k):
    raise IOError('xr_shim: no netCDF reader in this container')


def open_dataarray(*a, **k):
    raise IOError('xr_shim: no netCDF reader in this container')
