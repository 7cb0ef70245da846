"""``contrack`` class: host-side mirror of the reference's interface for the tracking path.

Mirrors steidani/ConTrack ``contrack/contrack.py`` for: the container (lines 49-161), ``read`` / ``read_xarray``
(166-199), ``set_up`` and dimension discovery (204-380), ``calc_clim`` (458-491), ``calc_anom`` (494-581) and
``run_contrack`` (583-796) -- same names, arguments, exceptions and result variable (``ds['flag']``).  The arithmetic
runs in the C-ABI library (CUDA kernels for sm_100a + a native ordered table phase); this file only prepares arguments
the way the reference does (dimension names, resolution, area weights, thresholds) and stores the result.

Differences kept deliberately small and documented:
  * xarray is optional: without it ``ds`` is a ``contrack_b200.dataset.Dataset`` (same indexing interface).
  * ``ds[variable].data`` may be a torch CUDA tensor; then the cube never leaves the GPU and ``flag`` is a CUDA tensor.
  * ``flag`` is int32 (the reference's scipy label dtype is int32 below 2**31-2 cells and int64 above; pass
    ``reference_dtype=True`` to widen on the host exactly where scipy would).
  * calc_clim / calc_anom / quantile keep the precision of their input like xarray does: float32 in -> float32 out,
    float64 in -> float64 out (every mean is accumulated in float64 either way); other dtypes are computed as float64.
"""
from __future__ import annotations

import logging

import numpy as np

from . import dataset as _ds
from ._lib import GORL_TO_OP, ContrackLibError
from .engine import Engine

try:                                        # xarray is optional (absent from this image)
    import xarray as xr
except Exception:                           # pragma: no cover
    xr = None

logging.basicConfig(level=logging.INFO)
logger = logging.getLogger(__name__)

GORL_ERRMSG = ' Please select from [>, >=, <, >=] for gorl'       # contrack.py:658, 673


def _is_dataset(ds):
    if isinstance(ds, _ds.Dataset):
        return True
    return xr is not None and isinstance(ds, xr.Dataset)


def _is_dataarray(x):
    if isinstance(x, _ds.DataArray):
        return True
    return xr is not None and isinstance(x, xr.DataArray)


def _host(x):
    """numpy view of coordinate-like data."""
    if isinstance(x, np.ndarray):
        return x
    if hasattr(x, 'detach'):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def time_group_keys(time_values, groupby):
    """Integer key of every time step for ``groupby`` in the sense of xarray's ``time.<groupby>`` accessor."""
    t = _host(time_values)
    if not np.issubdtype(t.dtype, np.datetime64):
        raise ValueError("time coordinate must be datetime64 to group by '{}'".format(groupby))
    day = t.astype('datetime64[D]')
    if groupby == 'dayofyear':
        return (day - t.astype('datetime64[Y]').astype('datetime64[D]')).astype(np.int64) + 1
    if groupby == 'month':
        return t.astype('datetime64[M]').astype(np.int64) % 12 + 1
    if groupby == 'year':
        return t.astype('datetime64[Y]').astype(np.int64) + 1970
    if groupby == 'day':
        return (day - t.astype('datetime64[M]').astype('datetime64[D]')).astype(np.int64) + 1
    if groupby == 'hour':
        return (t.astype('datetime64[h]') - day.astype('datetime64[h]')).astype(np.int64)
    raise ValueError("unsupported groupby '{}'".format(groupby))


class contrack(object):
    """contrack class (interface of contrack/contrack.py:49)."""

    num_of_contrack = 0

    def __init__(self, filename="", ds=None, device=0, **kwargs):
        self._device = device
        if not filename:
            self.ds = None if ds is None else ds
            return
        try:
            self.ds = None
            self.read(filename, **kwargs)
        except (OSError, IOError, RuntimeError):
            try:
                self.read(filename, **kwargs)
            except Exception:
                raise IOError("Unkown fileformat. Known formats are netcdf.")
        contrack.num_of_contrack += 1

    def __repr__(self):
        try:
            string = "\
            Xarray dataset with {} time steps. \n\
            Available fields: {}".format(self.ntime, ", ".join(self.variables))
        except AttributeError:
            string = "\
            Empty contrack container.\n\
            Hint: use read() to load data."
        return string

    def __str__(self):
        return 'Class {}: \n{}'.format(self.__class__.__name__, self.ds)

    def __len__(self):
        return len(self.ds)

    def __getattr__(self, attr):
        if attr in self.__dict__:
            return getattr(self, attr)
        if attr == 'ds':
            raise AttributeError(attr)
        return getattr(self.ds, attr)

    def __getitem__(self, key):
        return self.ds[key]

    @property
    def ntime(self):
        if len(self.ds.dims) != 3:
            logger.warning("\nBe careful with the dimensions, you want dims = 3 and shape:\n(latitude, longitude, time)")
        return self.ds.dims[self._get_name_time()]

    @property
    def variables(self):
        return list(self.ds.data_vars)

    @property
    def dimensions(self):
        return list(self.ds.dims)

    @property
    def grid(self):
        if len(self.ds.dims) != 3:
            logger.warning("\nBe careful with the dimensions, you want dims = 3 and shape:\n(latitude, longitude, time)")
            return None
        print("        latitude: {} \n        longitude: {}".format(
            self.ds.dims[self._get_name_latitude()], self.ds.dims[self._get_name_longitude()]))

    @property
    def dataset(self):
        return self.ds

    # ---- read (contrack.py:166-199) -------------------------------------------------------------------------------
    def read(self, filename, **kwargs):
        """netCDF through xarray when it is installed; ``.npz`` archives always."""
        if self.ds is not None:
            raise ValueError('contrack() is already set!')
        if str(filename).endswith('.npz'):
            self.ds = _ds.load_npz(filename)
        elif xr is not None:
            self.ds = xr.open_dataset(filename, **kwargs)
        else:
            raise IOError('xarray is not installed: cannot open {}'.format(filename))

    def read_xarray(self, ds):
        if self.ds is None:
            if not _is_dataset(ds):
                raise ValueError('ds has to be a xarray data set!')
            self.ds = ds
        else:
            raise ValueError('contrack() is already set!')

    # ---- set_up (contrack.py:204-380) -----------------------------------------------------------------------------
    def set_up(self, time_name=None, longitude_name=None, latitude_name=None, force=False, write=True):
        self._time_name = self._get_name_time() if time_name is None else time_name
        self._longitude_name = self._get_name_longitude() if longitude_name is None else longitude_name
        self._latitude_name = self._get_name_latitude() if latitude_name is None else latitude_name
        if (self._longitude_name and self._latitude_name) is not None:
            self._dlon = self._get_resolution(self._longitude_name, force=force)
            self._dlat = self._get_resolution(self._latitude_name, force=force)
        if self._time_name is not None:
            self._dtime = self._get_resolution(self._time_name, force=force)
        if write:
            logger.info("\n time: '{}'\n longitude: '{}'\n latitude: '{}'\n".format(
                self._time_name, self._longitude_name, self._latitude_name))

    def _get_name_time(self):
        for dim in self.ds.dims:
            if dim not in self.ds.variables:
                if dim in ['time']:
                    return dim
                continue
            a = self.ds[dim]
            if (('units' in a.attrs and 'since' in a.attrs['units']) or
                    ('units' in getattr(a, 'encoding', {}) and 'since' in a.encoding['units']) or dim in ['time']):
                return dim
        for dim in self.ds.variables:
            data = _host(self.ds[dim].data) if self.ds[dim].ndim <= 1 else None
            if data is None:
                continue
            try:
                var = data[0]
            except IndexError:
                var = data
            if isinstance(var, np.datetime64):
                return dim
        logger.warning("\n 'time' dimension (dtype='datetime64[ns]') not found.")
        return None

    def _get_name_longitude(self):
        for dim in self.ds.dims:
            attrs = self.ds[dim].attrs if dim in self.ds.variables else {}
            if (('units' in attrs and attrs['units'] in ['degree_east', 'degrees_east']) or
                    dim in ['lon', 'longitude', 'x']):
                return dim
        logger.warning("\n 'longitude' dimension (unit='degrees_east') not found.")
        return None

    def _get_name_latitude(self):
        for dim in self.ds.dims:
            attrs = self.ds[dim].attrs if dim in self.ds.variables else {}
            if (('units' in attrs and attrs['units'] in ['degree_north', 'degrees_north']) or
                    dim in ['lat', 'latitude', 'y']):
                return dim
        logger.warning("\n 'latitude' dimension (unit='degrees_north') not found.")
        return None

    def _get_resolution(self, dim, force=False):
        """contrack.py:327-380."""
        data = _host(self.ds[dim].data)
        if dim == self._time_name:
            if np.issubdtype(data.dtype, np.datetime64):
                delta = np.unique((data[1:] - data[:-1]).astype('timedelta64[h]'))
            elif 'units' in self.ds[dim].attrs and 'days' in self.ds[dim].attrs['units']:
                delta = np.unique(data[1:] - data[:-1])
            elif np.issubdtype(data.dtype, np.number):
                delta = np.unique(data[1:] - data[:-1])
            else:
                raise ValueError('Can not decode time with unit {}'.format(self.ds[dim].attrs.get('units')))
        else:
            delta = abs(np.unique((data[1:] - data[:-1])))
        if len(delta) > 1:
            errmsg = 'No regular grid found for dimension {}.\n\
            Hint: use set_up(force=True).'.format(dim)
            if force and dim != self._time_name:
                logging.warning(errmsg)
                logging.warning(' '.join(['force=True: using mean of non-equidistant', 'grid {}'.format(delta)]))
                delta = round(delta.mean(), 2)
            else:
                if dim == self._time_name:
                    logging.warning(errmsg)
                else:
                    raise ValueError(errmsg)
        elif len(delta) == 0:
            # a dimension with a single entry: the reference evaluates delta[0] here (contrack.py:371)
            raise IndexError('index 0 is out of bounds for axis 0 with size 0')
        elif len(delta) == 1 and delta[0] == 0:
            raise ValueError('Two equivalent values found for dimension {}.'.format(dim))
        elif len(delta) == 1 and delta[0] < np.zeros((), delta.dtype):
            raise ValueError(' '.join(['{} not increasing. This should', 'not happen?!']).format(dim))
        return delta

    def _ensure_set_up(self):
        if hasattr(self, '_time_name'):
            logger.info("\n time: '{}'\n longitude: '{}'\n latitude: '{}'\n".format(
                self._time_name, self._longitude_name, self._latitude_name))
        else:
            self.set_up()

    # ---- helpers ---------------------------------------------------------------------------------------------------
    def _engine(self):
        return Engine.get(self._device)

    def _cube_tlatlon(self, variable):
        """ds[variable] as a C-contiguous (time, lat, lon) array (numpy or torch CUDA) + the reference's `sort` list
        (contrack.py:677-681)."""
        a = self.ds[variable]
        dims = tuple(a.dims)
        sort = [dims.index(d) for d in [self._time_name, self._latitude_name, self._longitude_name]]
        data = a.data
        if sort != [0, 1, 2]:
            data = data.permute(*sort).contiguous() if hasattr(data, 'permute') else np.ascontiguousarray(
                np.transpose(np.asarray(data), sort))
        return data, dims, sort

    def area_weights(self):
        """Area weight per latitude row, the reference's expression verbatim (contrack.py:703-704): float32 values."""
        lat = _host(self.ds[self._latitude_name].data)
        weight_lat = np.cos(lat * np.pi / 180)
        w = np.array((111 * self._dlat * 111 * self._dlon * weight_lat)).astype(np.float32)
        return np.ascontiguousarray(np.broadcast_to(w, lat.shape), np.float64)

    # ---- calc_clim / calc_anom (contrack.py:458-581) ---------------------------------------------------------------
    def calculate_gph_from_gp(self, gp_name='z', gp_unit='m**2 s**-2', gph_name='z_height'):
        """contrack.py:386-425: geopotential height = geopotential / 9.80665 (float32 division on the device)."""
        g = 9.80665
        if self.ds[gp_name].attrs['units'] != gp_unit:
            raise ValueError('Geopotential unit should be {} not {}'.format(gp_unit, self.ds[gp_name].attrs['units']))
        a = self.ds[gp_name]
        data = a.data
        if str(data.dtype).endswith('float32'):
            out = self._engine().divide(data if hasattr(data, 'is_cuda') else np.ascontiguousarray(data), g)
        else:                                          # float64 input: numpy semantics are a plain float64 division
            out = data / g
        self.ds[gph_name] = self._variable(tuple(a.dims), out, {
            'units': 'm', 'long_name': 'Geopotential Height', 'standard_name': 'geopotential height',
            'history': 'Calculated from {} with g={}'.format(gp_name, g)})
        logger.info('Calculating GPH from GP... DONE')

    def calc_mean(self, variable):
        """contrack.py:428-455: mean along time (NaN skipped); float64 accumulation, float32 result."""
        if not variable:
            variable = 'z'
        elif variable not in self.variables:
            logger.warning("\n Variable '{}' not found. Select from {}.".format(variable, self.variables))
            return None
        if not hasattr(self, '_time_name'):
            self.set_up(write=False)
        data, dims, sort = self._cube_tlatlon(variable)
        T = int(data.shape[0])
        m = self._engine().calc_clim(data, np.zeros(T, np.int32), 1, 1)[0]
        rest = [d for d in dims if d != self._time_name]
        if rest != [self._latitude_name, self._longitude_name]:
            m = m.permute(1, 0) if hasattr(m, 'permute') else m.T
        return self._dataarray(m, rest, {d: self.ds[d] for d in rest if d in self.ds.variables},
                               attrs=dict(self.ds[variable].attrs), name=variable)

    def _groups(self, groupby):
        keys = time_group_keys(self.ds[self._time_name].data, groupby)
        uniq, idx = np.unique(keys, return_inverse=True)
        return uniq, idx.astype(np.int32)

    def calc_clim(self, variable, window=1, groupby='dayofyear'):
        self._ensure_set_up()
        data, dims, sort = self._cube_tlatlon(variable)
        uniq, gidx = self._groups(groupby)
        clim = self._engine().calc_clim(data, gidx, len(uniq), int(window))
        coords = {groupby: uniq}
        for d in (self._latitude_name, self._longitude_name):
            coords[d] = self.ds[d]
        return self._dataarray(clim, (groupby, self._latitude_name, self._longitude_name), coords,
                               attrs=dict(self.ds[variable].attrs), name=variable)

    def calc_anom(self, variable, window=1, smooth=1, groupby='dayofyear', clim=None):
        logger.info("Set up dimensions...")
        self._ensure_set_up()
        data, dims, sort = self._cube_tlatlon(variable)
        eng = self._engine()
        if clim is None:
            logger.info('Calculating climatological mean from {}...'.format(variable))
            uniq, gidx = self._groups(groupby)
            clim_mean = eng.calc_clim(data, gidx, len(uniq), int(window))
            clim_txt = 'from {} with running window time steps {}'.format(variable, window)
            ngroups = len(uniq)
        else:
            # contrack.py:551-565: a climatology that already has the `groupby` dimension, regridded to the grid of the
            # input by nearest neighbour (xarray's reindex(method='nearest') = pandas Index.get_indexer)
            logger.info('Reading climatological mean from {}...'.format(clim))
            if isinstance(clim, str):
                if xr is None:
                    raise IOError('xarray is not installed: cannot open {}'.format(clim))
                clim_da = xr.open_dataarray(clim)
            else:
                clim_da = clim
            clim_txt = clim
            if groupby not in clim_da.dims:
                raise ValueError("the climatology needs the dimension '{}' (contrack.py:562 groups it by time, which the "
                                 "reindex at :565 cannot handle)".format(groupby))
            clim_mean, ngroups, gidx = self._regrid_climatology(clim_da, groupby, eng)
        anom = eng.calc_anom(data, gidx, ngroups, clim_mean, int(smooth))
        if sort != [0, 1, 2]:
            inv = list(np.argsort(sort))
            anom = anom.permute(*inv).contiguous() if hasattr(anom, 'permute') else np.ascontiguousarray(
                np.transpose(anom, inv))
        attrs = self.ds[variable].attrs
        self.ds['anom'] = self._variable(dims, anom, {
            'units': attrs['units'],
            'long_name': attrs['long_name'] + ' Anomaly',
            'standard_name': attrs['long_name'] + ' anomaly',
            'history': ' '.join(['Calculated from {} with input attributes:', 'smoothing time steps = {},',
                                 'climatology = {}.']).format(variable, smooth, clim_txt)})
        logger.info('Calculating Anomaly... DONE')

    def _regrid_climatology(self, clim_da, groupby, eng):
        """(clim [G, H, W] on this dataset's grid, G, group index of every time step) for an external climatology."""
        import pandas as pd
        cdims = tuple(clim_da.dims)
        order = [cdims.index(d) for d in (groupby, self._latitude_name, self._longitude_name)]
        cdata = clim_da.data
        if order != [0, 1, 2]:
            cdata = cdata.permute(*order).contiguous() if hasattr(cdata, 'permute') else np.ascontiguousarray(
                np.transpose(np.asarray(cdata), order))
        lat_c = _host(clim_da[self._latitude_name].data)
        lon_c = _host(clim_da[self._longitude_name].data)
        iy = pd.Index(lat_c).get_indexer(pd.Index(_host(self.ds[self._latitude_name].data)), method='nearest')
        ix = pd.Index(lon_c).get_indexer(pd.Index(_host(self.ds[self._longitude_name].data)), method='nearest')
        regridded = eng.gather_planes(cdata, iy, ix)
        keys = _host(clim_da[groupby].data)
        tkeys = time_group_keys(self.ds[self._time_name].data, groupby)
        srt = np.argsort(keys, kind='stable')
        pos = np.searchsorted(keys[srt], tkeys)
        bad = (pos >= len(keys)) | (keys[srt][np.minimum(pos, len(keys) - 1)] != tkeys)
        if bad.any():
            raise KeyError('the climatology has no entry for {} {}'.format(groupby, sorted(set(tkeys[bad].tolist()))[:5]))
        return regridded, len(keys), srt[pos].astype(np.int32)

    def quantile(self, variable, q, latitude=None):
        """ds[variable].sel(latitude=slice(a, b)).quantile(q, dim=time) of README.rst:150-151 on the device: float64
        DataArray ('quantile', lat, lon).  `latitude` = slice(a, b) in coordinate values, label based and inclusive like
        xarray's .sel (order as the coordinate runs: slice(80, 50) on a north-to-south axis)."""
        self._ensure_set_up()
        data, dims, sort = self._cube_tlatlon(variable)
        lat = _host(self.ds[self._latitude_name].data)
        y0, y1 = 0, len(lat)
        if latitude is not None:
            import pandas as pd
            sl = pd.Index(lat).slice_indexer(latitude.start, latitude.stop)
            y0, y1 = sl.start, sl.stop
        qa = np.atleast_1d(np.asarray(q, np.float64))
        out = self._engine().quantile_time(data, qa, y0, y1)
        coords = {'quantile': qa, self._latitude_name: lat[y0:y1], self._longitude_name: self.ds[self._longitude_name]}
        return self._dataarray(out, ('quantile', self._latitude_name, self._longitude_name), coords, name=variable)

    def quantile_threshold(self, variable, q, latitude=None):
        """README.rst:150-151: float(ds[variable].sel(latitude=...).quantile([q], dim='time').mean()) -- the objective
        threshold recommended for run_contrack.  The per-grid-point quantiles come from the device (numpy's values bit for
        bit); the final mean over the band is numpy's own nanmean on the host."""
        qf = self.quantile(variable, [q], latitude)
        return float(np.nanmean(qf.values))

    def blocking_frequency(self, flag='flag', greater_than=1):
        """README.rst:161: xr.where(ds[flag] > 1, 1, 0).sum(dim='time') / ntime * 100 -> DataArray (lat, lon) float64."""
        self._ensure_set_up()
        data, dims, sort = self._cube_tlatlon(flag)
        cnt = self._engine().flag_count(data, greater_than)
        ntime = int(data.shape[0])
        cnt64 = cnt.double() if hasattr(cnt, 'is_cuda') else cnt.astype(np.int64)      # numpy: int64 / int -> float64
        freq = cnt64 / ntime * 100
        return self._dataarray(freq, (self._latitude_name, self._longitude_name),
                               {d: self.ds[d] for d in (self._latitude_name, self._longitude_name)}, name=flag)

    def _variable(self, dims, data, attrs):
        if xr is not None and isinstance(self.ds, xr.Dataset):
            return xr.Variable(dims, _host(data), attrs=attrs)
        return _ds.Variable(dims, data, attrs)

    def _dataarray(self, data, dims, coords, attrs=None, name=None):
        """A DataArray of the dataset's own kind: xarray.DataArray (host data) when ``ds`` is an xarray.Dataset -- what
        calc_clim / calc_mean return in the reference (tests/test_contrack.py:79-80) -- else the stand-in."""
        if xr is not None and isinstance(self.ds, xr.Dataset):
            return xr.DataArray(_host(data), dims=tuple(dims), coords={k: _host(getattr(v, 'data', v)) for k, v in coords.items()},
                                attrs=dict(attrs or {}), name=name)
        return _ds.DataArray(data, tuple(dims), coords={k: (v if isinstance(v, _ds.DataArray) else
                                                           _ds.DataArray(_host(getattr(v, 'data', v)), (k,)))
                                                        for k, v in coords.items()}, attrs=attrs, name=name)

    # ---- run_contrack (contrack.py:583-796) ------------------------------------------------------------------------
    def run_contrack(self, variable, threshold, gorl, overlap, persistence, twosided=True, reference_dtype=False,
                     time_shard=None, comm=None):
        """The reference's call (contrack.py:583-590).  Two additions, both optional:
        reference_dtype -- widen the int32 result to int64 where scipy would (cubes of 2^31 - 2 cells and more);
        time_shard=(t_begin, T_total) -- this object holds planes [t_begin, t_begin + ntime) of a cube of T_total time steps
        that is spread over several processes / GPUs in time order (one contrack object per rank, each with its own slice of
        the time axis): the call is then collective over ``comm`` (a contrack_b200.sharded.Comm; default: an NCCL
        communicator over the torch.distributed world) and every rank receives its own planes of the flag variable with
        the ids of the unsharded run.  A dayofyear threshold is looked up on this rank's own time axis."""
        logger.info("\nRun ConTrack \n########### \n    threshold:    {} {} \n    overlap:      {} \n"
                    "    persistence:  {} time steps".format(gorl, threshold, overlap, persistence))
        logger.info("Set up dimensions...")
        self._ensure_set_up()

        # step 1 (contrack.py:646-674): threshold per time step + the precision numpy would compare in
        logger.info("Find individual contours...")
        if gorl not in GORL_TO_OP:
            raise ValueError(GORL_ERRMSG)
        data, dims, sort = self._cube_tlatlon(variable)
        is_f32 = str(data.dtype).endswith('float32')
        if _is_dataarray(threshold):
            thr_keys = _host(threshold['dayofyear'].data)
            thr_vals = _host(threshold.data)
            doy = time_group_keys(self.ds[self._time_name].data, 'dayofyear')
            srt = np.argsort(thr_keys, kind='stable')      # xarray aligns the groups by LABEL: the coordinate may be unsorted
            pos = np.searchsorted(thr_keys[srt], doy)
            if (pos >= len(thr_keys)).any() or (thr_keys[srt][np.minimum(pos, len(thr_keys) - 1)] != doy).any():
                raise KeyError('threshold has no value for some dayofyear of the time axis')
            thr = thr_vals[srt[pos]].astype(np.float64)
            thr_is_f32 = is_f32 and thr_vals.dtype == np.float32
        else:
            thr = np.array([threshold], np.float64)
            # numpy 2 promotion: python scalars are weak (compare in the array's float32), numpy float64 scalars are not
            weak = isinstance(threshold, (int, float)) and not isinstance(threshold, np.floating)
            # (numpy integer scalars of 4+ bytes promote a float32 array to float64; the small ones keep float32)
            small_int = isinstance(threshold, np.integer) and np.dtype(type(threshold)).itemsize <= 2
            thr_is_f32 = is_f32 and (weak or small_int or isinstance(threshold, (np.float32, np.float16, np.bool_)))

        # steps 2-4 on the GPU
        logger.info("Apply overlap...")
        logger.info("Apply persistence...")
        if time_shard is None:
            flag, num_features = self._engine().run_contrack(
                data, self.area_weights(), thr, thr_is_f32, GORL_TO_OP[gorl], overlap, persistence, twosided)
        else:
            from . import sharded
            t_begin, T_total = (int(v) for v in time_shard)
            if t_begin < 0 or t_begin + int(data.shape[0]) > T_total:
                raise ValueError('time_shard: planes [%d, %d) are outside a cube of %d time steps'
                                 % (t_begin, t_begin + int(data.shape[0]), T_total))
            flag, num_features, _ = sharded.run_contrack_sharded(
                self._engine(), data, t_begin, T_total, self.area_weights(), thr, thr_is_f32, GORL_TO_OP[gorl], overlap,
                persistence, twosided, comm=comm)
        ncells = int(np.prod([int(n) for n in flag.shape]))       # (a torch tensor's .size is a method)
        if reference_dtype and ncells >= 2 ** 31 - 2:
            flag = flag.long() if hasattr(flag, 'long') else flag.astype(np.int64)

        # step 5 (contrack.py:775-791): the reference applies `transpose(sort)` to the (time, lat, lon) result
        logger.info("Create new variable 'flag'...")
        if sort != [0, 1, 2]:
            flag = flag.permute(*sort) if hasattr(flag, 'permute') else np.transpose(flag, sort)
        self.ds['flag'] = self._variable(dims, flag, {
            'units': 'flag',
            'long_name': 'contrack flag',
            'standard_name': 'contrack flag',
            'history': ' '.join(['Calculated from {} with input attributes:', 'threshold = {} {},',
                                 'overlap fraction = {},', 'persistence time steps = {}.', 'twosided = {}'])
            .format(variable, gorl, threshold, overlap, persistence, twosided),
            'reference': 'https://github.com/steidani/ConTrack'})
        logger.info("Running contrack... DONE\n{} contours tracked".format(num_features))

    # ---- run_lifecycle (contrack.py:799-907) -----------------------------------------------------------------------
    def run_lifecycle(self, flag, variable):
        """Life cycle analysis: intensity, spatial extent and centre of mass of every flagged contour per time step.
        Returns a pandas DataFrame with the reference's columns ['Flag', 'Date', 'Longitude', 'Latitude', 'Intensity',
        'Size'], sorted by (Flag, Date) (contrack.py:907).  The sums come from the CUDA kernels in the reference's own
        summation orders (ct_run_lifecycle); this method does what the reference does with them on the host: the
        divisions, the int() truncations, the coordinate look-ups, round(, 2) and the date strings."""
        import pandas as pd
        logger.info("\nRun Lifecycle \n########### \n    flag:    {}\n    variable:    {}".format(flag, variable))
        logger.info("Set up dimensions...")
        self._ensure_set_up()
        fdata, _, _ = self._cube_tlatlon(flag)
        vdata, _, _ = self._cube_tlatlon(variable)
        lat = _host(self.ds[self._latitude_name].data)
        lon = _host(self.ds[self._longitude_name].data)
        nlon = len(lon)
        times = _host(self.ds[self._time_name].data)
        if not np.issubdtype(times.dtype, np.datetime64):
            raise AttributeError("Can only use .dt accessor with datetimelike values")        # what xarray raises at 862
        res = self._engine().run_lifecycle(fdata, vdata, self.area_weights())
        flag_dtype = np.dtype(str(fdata.dtype).replace('torch.', '')) if not isinstance(fdata, np.ndarray) else fdata.dtype
        rows = []
        date_of = {}
        for i in range(len(res['t'])):
            t = int(res['t'][i])
            if t not in date_of:
                s = str(times[t].astype('datetime64[h]'))
                date_of[t] = s[0:4] + s[5:7] + s[8:10] + '_' + s[11:13]                      # strftime('%Y%m%d_%H')
            areacon = res['area'][i]
            intensitycon = res['wsum'][i] / areacon                                            # contrack.py:876
            com_y = res['sy'][i] / res['norm'][i]                                              # ndimage.center_of_mass
            com_x = res['sx'][i] / res['norm'][i]
            comlatcon = int(lat[int(com_y)])                                                   # contrack.py:888 / 893
            ix = int(com_x)
            roll = int(res['roll'][i])
            if roll >= 0:                                                                      # longitude rolled by -roll
                if ix >= nlon or ix < -nlon:
                    raise IndexError('index {} is out of bounds for axis 0 with size {}'.format(ix, nlon))
                comloncon = int(lon[(ix + roll) % nlon])
            else:
                comloncon = int(lon[ix])
            rows.append((flag_dtype.type(res['label'][i]), date_of[t], comloncon, comlatcon, round(intensitycon, 2),
                         round(areacon, 2)))
        rows.sort(key=lambda x: (x[0], x[1]))
        return pd.DataFrame(rows, columns=['Flag', 'Date', 'Longitude', 'Latitude', 'Intensity', 'Size'])


__all__ = ['contrack', 'ContrackLibError', 'time_group_keys']
