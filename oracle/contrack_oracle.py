"""CPU oracle for ConTrack's ``run_contrack`` / ``calc_clim`` / ``calc_anom`` hot path.

TEST INFRASTRUCTURE ONLY.  This module is the parity checker; nothing in the product package
(``contrack_b200/``) may import it.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it.

It is an xarray-free restatement of the reference (``/root/reference/contrack/contrack.py``).  The reference
cannot be imported in this image (``import xarray`` fails at contrack.py:19), so the restatement replaces only
the xarray wrapper (``xr.where`` -> ``np.where``, ``ds[...]`` -> plain arrays) and calls the *same* third-party
compiled code the reference calls: ``scipy.ndimage.label`` / ``scipy.ndimage.find_objects`` (scipy is unpinned in
the reference's requirements.txt:2; this image has scipy 1.18.1) and numpy.

Parity status: PINNED for run_contrack / run_lifecycle.  The reference's own tests hold only counts on its fixture
(tests/test_contrack.py:83-103: 3 features, 28 lifecycle rows); beyond those, tests/golden/make_reference_golden.py runs
the UNMODIFIED reference source in the build container (under tests/golden/xr_shim, a stand-in for the labelled-array
plumbing of xarray only) and records the sha256 of every flag cube and every lifecycle table it produces
(tests/golden/reference_run.json: the fixture with 4 parameter sets, 9 synthetic runs, 14 stale-box quirk cubes, a
non-default dimension order).  ``tests/test_oracle.py`` holds this restatement to all of them, bit for bit.
calc_clim / calc_anom: pinned through pandas (the shim implements groupby / rolling with pandas, an interpretation of
xarray's semantics, not xarray itself) within rtol 1e-5 / atol 4e-3 -- treat as "partially pinned".
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage

# contrack.py:684-686 -- structure with only the middle (same-time) plane set: per-plane 2-D 8-connectivity
STRUCT_2D_IN_3D = np.array([[[0, 0, 0], [0, 0, 0], [0, 0, 0]],
                            [[1, 1, 1], [1, 1, 1], [1, 1, 1]],
                            [[0, 0, 0], [0, 0, 0], [0, 0, 0]]])
# contrack.py:748-750 -- 8-connectivity in plane, same-pixel link to t-1 / t+1
STRUCT_3D = np.array([[[0, 0, 0], [0, 1, 0], [0, 0, 0]],
                      [[1, 1, 1], [1, 1, 1], [1, 1, 1]],
                      [[0, 0, 0], [0, 1, 0], [0, 0, 0]]])

GORL_ERRMSG = ' Please select from [>, >=, <, >=] for gorl'   # contrack.py:658, 673 (text kept verbatim)


def resolution(coord, force=False, name='dim'):
    """contrack.py:352-380 (lat/lon branch of _get_resolution) -- ``abs(np.unique(diff))``: a length-1 ndarray
    for a regular grid; irregular grids raise unless ``force`` (then the rounded mean is used)."""
    coord = np.asarray(coord)
    delta = abs(np.unique(coord[1:] - coord[:-1]))
    if len(delta) > 1:
        if not force:
            raise ValueError('No regular grid found for dimension {}.\n\
            Hint: use set_up(force=True).'.format(name))
        delta = round(delta.mean(), 2)
    elif delta[0] == 0:
        raise ValueError('Two equivalent values found for dimension {}.'.format(name))
    return delta


def weight_grid(lat, dlat, dlon, nlon):
    """contrack.py:703-704 -- area weights, float32 values widened to a float64 [H, W] grid."""
    weight_lat = np.cos(np.asarray(lat) * np.pi / 180)
    return np.ones((len(lat), nlon)) * np.array((111 * dlat * 111 * dlon * weight_lat)).astype(np.float32)[:, None]


def threshold_mask(anom, threshold, gorl):
    """contrack.py:663-674 -- numeric threshold branch.  ``threshold`` may also be a length-T array
    (the per-time values the DataArray/dayofyear branch, contrack.py:648-661, would broadcast)."""
    thr = threshold
    if isinstance(thr, np.ndarray) and thr.ndim == 1:
        thr = thr[:, None, None]
    if gorl == '>=' or gorl == 'ge':
        return np.where(anom >= thr, 1, 0)
    elif gorl == '<=' or gorl == 'le':
        return np.where(anom <= thr, 1, 0)
    elif gorl == '>' or gorl == 'gt':
        return np.where(anom > thr, 1, 0)
    elif gorl == '<' or gorl == 'lt':
        return np.where(anom < thr, 1, 0)
    raise ValueError(GORL_ERRMSG)


def run_contrack(anom, lat, lon, threshold, gorl, overlap, persistence, twosided=True, stages=None, force=False):
    """Restatement of contrack.py:646-772 on plain arrays.

    anom: [T, H, W] (time, lat, lon) float array; lat [H], lon [W] coordinate arrays.
    Returns the integer ``flag`` array [T, H, W] (dtype as scipy.ndimage.label gives it).
    ``stages``: optional dict that receives copies of intermediate arrays (for staged parity tests).
    """
    anom = np.asarray(anom)
    T, H, W = anom.shape
    dlon = resolution(lon, force, 'longitude')     # contrack.py:251-252
    dlat = resolution(lat, force, 'latitude')

    # step 1 (contrack.py:646-674)
    flag = threshold_mask(anom, threshold, gorl)

    # step 2 (contrack.py:684-687)
    flag, num_features = ndimage.label(flag, structure=STRUCT_2D_IN_3D)
    if stages is not None:
        stages['label2d'] = flag.copy()

    # periodic boundary (contrack.py:691-698)
    for tt in range(T):
        for yy in range(H):
            if flag[tt, yy, 0] > 0 and flag[tt, yy, -1] > 0 and (flag[tt, yy, 0] > flag[tt, yy, -1]):
                flag[tt][flag[tt] == flag[tt, yy, 0]] = flag[tt, yy, -1]
            if flag[tt, yy, 0] > 0 and flag[tt, yy, -1] > 0 and (flag[tt, yy, 0] < flag[tt, yy, -1]):
                flag[tt][flag[tt] == flag[tt, yy, -1]] = flag[tt, yy, 0]
    if stages is not None:
        stages['label2d_seam'] = flag.copy()

    # step 3 (contrack.py:700-742)
    wgrid = weight_grid(lat, dlat, dlon, W)
    for tt in range(1, T - 1):
        slices = ndimage.find_objects(flag[tt])
        label = 0
        for slice_ in slices:
            label = label + 1
            if slice_ is None:
                continue
            areacon = np.sum(wgrid[slice_][flag[tt][slice_] == label])
            areaover_forward = np.sum(wgrid[slice_][(flag[tt][slice_] == label) & (flag[tt + 1][slice_] >= 1)])
            areaover_backward = np.sum(wgrid[slice_][(flag[tt][slice_] == label) & (flag[tt - 1][slice_] >= 1)])

            with np.errstate(divide='ignore', invalid='ignore'):
                fraction_backward = (1 / areacon) * areaover_backward
                fraction_forward = (1 / areacon) * areaover_forward

            if twosided:
                if fraction_backward != 0 and fraction_forward != 0:
                    if (fraction_backward < overlap) or (fraction_forward < overlap):
                        flag[tt][slice_][(flag[tt][slice_] == label)] = 0.
                if fraction_backward != 0 and fraction_forward == 0:
                    if (fraction_backward < overlap):
                        flag[tt][slice_][(flag[tt][slice_] == label)] = 0.
                if fraction_backward == 0 and fraction_forward != 0:
                    if (fraction_forward < overlap):
                        flag[tt][slice_][(flag[tt][slice_] == label)] = 0.
            else:
                if (fraction_forward < overlap):
                    flag[tt][slice_][(flag[tt][slice_] == label)] = 0.
    if stages is not None:
        stages['filtered'] = flag.copy()

    return track_persistence(flag, persistence, stages)


def track_persistence(flag, persistence, stages=None):
    """Step 4 of run_contrack alone (contrack.py:744-772): re-binarise, 3-D label, seam merge through the
    bounding boxes taken BEFORE merging (contrack.py:753), persistence filter."""
    T, H, W = flag.shape
    flag = np.where(flag >= 1, 1, 0)
    flag, num_features = ndimage.label(flag, structure=STRUCT_3D)
    if stages is not None:
        stages['label3d'] = flag.copy()
    slices = ndimage.find_objects(flag)
    for tt in range(T):
        for yy in range(H):
            if flag[tt, yy, 0] > 0 and flag[tt, yy, -1] > 0 and (flag[tt, yy, 0] > flag[tt, yy, -1]):
                slice_ = slices[flag[tt, yy, 0] - 1]
                flag[slice_][(flag[slice_] == flag[tt, yy, 0])] = flag[tt, yy, -1]
            if flag[tt, yy, 0] > 0 and flag[tt, yy, -1] > 0 and (flag[tt, yy, 0] < flag[tt, yy, -1]):
                slice_ = slices[flag[tt, yy, -1] - 1]
                flag[slice_][(flag[slice_] == flag[tt, yy, -1])] = flag[tt, yy, 0]
    if stages is not None:
        stages['label3d_seam'] = flag.copy()
    label = 0
    for slice_ in ndimage.find_objects(flag):
        label = label + 1
        if slice_ is None:
            continue
        if (slice_[0].stop - slice_[0].start) < persistence:
            flag[slice_][(flag[slice_] == label)] = 0.
    return flag


def num_features(flag):
    """contrack.py:793 -- count of distinct non-zero ids."""
    return len(np.unique(flag)) - 1


# ---------------------------------------------------------------------------------------------------------------
# calc_clim / calc_anom (contrack.py:458-581) restated with numpy.  xarray semantics used:
#   groupby(...).mean(time): skip-NaN mean per group; rolling(center=True).mean(): window w, min_periods = w,
#   window for output i covers [i - w//2, i - w//2 + w) (pandas/xarray convention), NaN where incomplete.
# ---------------------------------------------------------------------------------------------------------------

def _rolling_mean_centered(x, window):
    """Centred rolling mean along axis 0 with min_periods = window (NaN where the window is incomplete or
    contains NaN).  Accumulates in float64 and returns x's dtype (tolerance parity only)."""
    n = x.shape[0]
    out = np.full(x.shape, np.nan, dtype=x.dtype)
    if window <= 0 or window > n:
        return out
    left = window // 2
    m = n - window + 1                                     # complete windows: outputs left .. left + m - 1
    # direct window sums (a NaN makes only the windows that contain it NaN, as xarray's rolling mean does; a cumulative
    # sum would poison every later window)
    acc = np.zeros((m,) + x.shape[1:], np.float64)
    for k in range(window):
        acc += x[k:k + m]
    out[left:left + m] = (acc / window).astype(x.dtype)
    return out


def calc_clim(z, groups, window=1):
    """contrack.py:483-489.  ``groups``: int array [T] of group keys (e.g. dayofyear 1..366).
    Returns (unique_keys [G], clim [G, H, W]) sorted by key (as xarray groupby does)."""
    z = np.asarray(z)
    keys = np.unique(groups)
    clim = np.empty((len(keys),) + z.shape[1:], dtype=z.dtype)
    for gi, k in enumerate(keys):
        with np.errstate(invalid='ignore'):
            clim[gi] = np.nanmean(z[groups == k], axis=0)
    smooth = _rolling_mean_centered(clim, window)
    fill = np.nanmean(clim[-window:], axis=0)          # contrack.py:488 -- last `window` UNSMOOTHED entries
    smooth = np.where(np.isnan(smooth), fill[None], smooth)
    return keys, smooth.astype(z.dtype)


def calc_anom(z, groups, window=1, smooth=1):
    """contrack.py:549, 568-570: (x.groupby(g) - clim).rolling(time=smooth, center=True).mean()."""
    z = np.asarray(z)
    keys, clim = calc_clim(z, groups, window)
    idx = np.searchsorted(keys, groups)
    dev = z - clim[idx]
    return _rolling_mean_centered(dev, smooth)


def strftime_ymd_h(time_values):
    """``ds[time].dt.strftime('%Y%m%d_%H')`` (contrack.py:862) for a datetime64 vector."""
    t = np.asarray(time_values).astype('datetime64[h]')
    out = []
    for v in t:
        s = str(v)                                   # 'YYYY-MM-DDTHH'
        out.append(s[0:4] + s[5:7] + s[8:10] + '_' + s[11:13])
    return out


def run_lifecycle(flag, var, lat, lon, time_values, force=False):
    """Restatement of contrack.py:799-907 on plain arrays: per (time step, flag id) size, area-weighted intensity and
    centre of mass (with the longitude roll for contours that touch both the first and the last column).  Same numpy /
    scipy calls as the reference (np.unique, np.sum over boolean-masked arrays, ndimage.center_of_mass, np.roll for
    xarray's .roll(roll_coords=True)).  Returns the sorted list of tuples that the reference feeds to pd.DataFrame
    (columns Flag, Date, Longitude, Latitude, Intensity, Size)."""
    flag = np.asarray(flag)
    var = np.asarray(var)
    lat = np.asarray(lat)
    lon = np.asarray(lon)
    T, H, W = flag.shape
    dlat, dlon = resolution(lat, force, 'latitude'), resolution(lon, force, 'longitude')
    weight_lat = np.cos(lat * np.pi / 180)                                                      # contrack.py:847
    wgrid = np.ones((H, W)) * np.array((111 * dlat * 111 * dlon * weight_lat)).astype(np.float32)[:, None]
    dates = strftime_ymd_h(time_values)
    block_id, time, intensity, size, com_lon, com_lat = [], [], [], [], [], []
    for i_time in range(T):                                                                     # contrack.py:860
        currentstep = dates[i_time]
        f = flag[i_time]
        v = var[i_time]
        labels = np.unique(f)
        labels = labels[labels != 0]
        if len(labels) == 0:
            continue
        for label in labels:
            areacon = np.sum(wgrid[f == label])                                                 # contrack.py:874
            intensitycon = np.sum(wgrid[f == label] * v[f == label])
            intensitycon = intensitycon / areacon
            if label in f[:, 0] and label in f[:, -1]:                                          # contrack.py:880
                yloc, xloc = np.where(f == label)
                lon_roll = np.unique(xloc)[np.argmax(np.diff(np.unique(xloc))) + 1]
                flag_roll = np.roll(f, -lon_roll, axis=1)
                variable_roll = np.roll(v, -lon_roll, axis=1)
                lon_rolled = np.roll(lon, -lon_roll)
                center_of_mass = ndimage.center_of_mass(variable_roll * wgrid, flag_roll, [label])
                comlatcon = int(lat[int(center_of_mass[0][0])])
                comloncon = int(lon_rolled[int(center_of_mass[0][1])])
            else:
                center_of_mass = ndimage.center_of_mass(v * wgrid, f, [label])
                comlatcon = int(lat[int(center_of_mass[0][0])])
                comloncon = int(lon[int(center_of_mass[0][1])])
            block_id.append(label)
            time.append(str(currentstep))
            intensity.append(round(intensitycon, 2))
            size.append(round(areacon, 2))
            com_lon.append(comloncon)
            com_lat.append(comlatcon)
    return sorted(list(zip(block_id, time, com_lon, com_lat, intensity, size)), key=lambda x: (x[0], x[1]))


# ---------------------------------------------------------------------------------------------------------------
# Callers either side of the path: the README's recipes, restated with the same numpy / pandas calls
# xarray makes.
# ---------------------------------------------------------------------------------------------------------------

def quantile_time(x, q):
    """README.rst:150-151 ``ds[var].quantile(q, dim='time')``: xarray (Variable.quantile) calls
    ``np.nanquantile(data, q float64 array, axis=time, method='linear')`` for float data."""
    return np.nanquantile(np.asarray(x), np.atleast_1d(np.asarray(q, dtype=np.float64)), axis=0)


def label_slice(coord, start, stop):
    """positions selected by ``.sel(dim=slice(start, stop))`` (label based, both ends inclusive): pandas slice_indexer."""
    import pandas as pd
    return pd.Index(np.asarray(coord)).slice_indexer(start, stop)


def quantile_threshold(x, lat, q, lat_start, lat_stop):
    """README.rst:151 ``ds[var].sel(latitude=slice(a, b)).quantile([q], dim='time').mean()`` as a Python float."""
    sl = label_slice(lat, lat_start, lat_stop)
    return float(np.nanmean(quantile_time(np.asarray(x)[:, sl, :], [q])))


def blocking_frequency(flag, greater_than=1):
    """README.rst:161 ``xr.where(flag > 1, 1, 0).sum(dim='time') / ntime * 100``."""
    flag = np.asarray(flag)
    return np.where(flag > greater_than, 1, 0).sum(axis=0) / flag.shape[0] * 100


def gph_from_gp(gp):
    """contrack.py:417-419 ``data / g`` with g = 9.80665 (float32 data stay float32)."""
    return np.asarray(gp) / 9.80665


def nearest_index(src_coord, dst_coord):
    """``reindex(..., method='nearest')`` (contrack.py:565): pandas Index.get_indexer, which xarray delegates to."""
    import pandas as pd
    return pd.Index(np.asarray(src_coord)).get_indexer(pd.Index(np.asarray(dst_coord)), method='nearest')


def calc_anom_external(z, time_keys, clim, clim_keys, clim_lat, clim_lon, lat, lon, smooth=1):
    """contrack.py:551-570 with a supplied climatology [G, Hc, Wc] that has the group dimension: nearest-neighbour
    regrid, ``z.groupby(key) - clim``, centred rolling mean over time."""
    clim = np.asarray(clim)[:, nearest_index(clim_lat, lat)][:, :, nearest_index(clim_lon, lon)]
    clim_keys = np.asarray(clim_keys)
    order = np.argsort(clim_keys, kind='stable')
    pos = order[np.searchsorted(clim_keys[order], np.asarray(time_keys))]
    dev = np.asarray(z) - clim[pos]
    return _rolling_mean_centered(dev, smooth)
