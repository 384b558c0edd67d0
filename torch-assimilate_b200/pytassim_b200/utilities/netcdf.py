"""The on-disk format of exported ensemble weights (``weight_save_path``).

Reference: pytassim/utilities/xarray.py:36-173 (``save_netcdf``, ``load_netcdf``, ``encode_multidim``, ``decode_multidim``) as
used by ``BaseAssimilation.store_weights`` / ``load_weights`` (pytassim/interface/base.py:280-324) between
``estimate_weights`` and ``_apply_weights`` (interface/filter.py:157-163).

The reference writes with ``xarray.DataArray.to_netcdf``; xarray and netCDF4 are not part of this image, so the file is
written directly as netCDF-3 (64-bit offset) with ``scipy.io.netcdf_file`` — the layout xarray's own scipy backend
produces for an unnamed DataArray: one data variable ``__xarray_dataarray_variable__``, one coordinate variable per indexed
dimension, and a ``MultiIndex`` dimension encoded as the reference does (xarray.py:78-90): an integer range as the
dimension coordinate carrying the attribute ``multidim_levels = "name_1;name_2"`` plus one coordinate variable per level
(listed in the data variable's ``coordinates`` attribute).  Files written here open with ``xarray.open_dataarray`` +
``decode_multidim`` where xarray exists, and files written by the reference through its scipy backend load here.

Size.  A fixed-size netCDF-3 variable is limited to 2 GiB in scipy's writer (the ``vsize`` field is packed as a signed
32-bit integer); per-grid-point weights ``(N, k, k)`` float64 pass that at N ~ 168 000 for k = 40.  Arrays of 2 GiB or more
are therefore written with their FIRST dimension (``grid``) as the record (unlimited) dimension — one record per grid point,
no limit on the number of records; xarray reads record variables like any other.  String labels are stored the way xarray's
netCDF-3 encoder stores them: character arrays with a trailing ``string<n>`` dimension.
"""
import numpy as np
import pandas as pd
from scipy.io import netcdf_file

from ..xrlite import DataArray

__all__ = ['save_netcdf', 'load_netcdf', 'encode_multidim', 'decode_multidim', 'DATAARRAY_VARIABLE']

DATAARRAY_VARIABLE = '__xarray_dataarray_variable__'        # xarray.backends.api.DATAARRAY_VARIABLE
_TIME_UNITS = 'seconds since 1970-01-01 00:00:00'
FIXED_VARIABLE_LIMIT = 2 ** 31 - 4                           # bytes; larger arrays go to the record layout (see above)


def encode_multidim(array):
    """xarray.py:58-90 on the (values, dims, indexes) protocol: -> (dims coordinates {dim: (values, attrs)}, level coordinates
    {name: (dim, values)}).  A MultiIndex dimension becomes ``arange(n)`` with ``multidim_levels``; its levels become plain
    coordinates on that dimension."""
    dim_coords, level_coords = {}, {}
    for dim in array.dims:
        index = array.indexes[dim]
        if isinstance(index, pd.MultiIndex):
            names = list(index.names)
            if any(n is None for n in names):
                raise ValueError("the levels of the MultiIndex on '{0}' need names to be stored".format(dim))
            for name in names:
                level_coords[name] = (dim, np.asarray(index.get_level_values(name)))
            dim_coords[dim] = (np.arange(len(index)), {'multidim_levels': ';'.join(names)})          # xarray.py:84-89
        else:
            dim_coords[dim] = (np.asarray(index), {})
    return dim_coords, level_coords


def _to_nc3(values):
    """netCDF-3 has no 64-bit integers / booleans / datetimes: the casts xarray's netCDF-3 encoder applies."""
    values = np.asarray(values)
    attrs = {}
    if np.issubdtype(values.dtype, np.datetime64):
        values = (values.astype('datetime64[ns]') - np.datetime64('1970-01-01T00:00:00', 'ns')) / np.timedelta64(1, 's')
        values = np.asarray(values, dtype=np.float64)
        attrs = {'units': _TIME_UNITS, 'calendar': 'proleptic_gregorian'}
    elif values.dtype == np.bool_:
        values = values.astype(np.int8)
        attrs = {'dtype': 'bool'}
    elif np.issubdtype(values.dtype, np.integer) and values.dtype.itemsize > 4:
        if values.size and (values.max() > np.iinfo(np.int32).max or values.min() < np.iinfo(np.int32).min):
            raise ValueError("integer coordinate does not fit netCDF-3's 32-bit integers")
        values = values.astype(np.int32)
    elif values.dtype.kind in 'OUS':
        enc = [v if isinstance(v, bytes) else str(v).encode('utf-8') for v in values.ravel()]
        width = max([len(v) for v in enc] + [1])
        chars = np.zeros((len(enc), width), dtype='S1')
        for i, v in enumerate(enc):
            chars[i, :len(v)] = np.frombuffer(v, dtype='S1')
        values = chars.reshape(values.shape + (width,))
        attrs = {'_Encoding': 'utf-8'}
    return values, attrs


def _create(nc, name, values, dims):
    """createVariable for plain and for character data (trailing ``string<n>`` dimension, shared between variables)."""
    if values.dtype.kind == 'S':
        sdim = 'string{0}'.format(values.shape[-1])
        if sdim not in nc.dimensions:
            nc.createDimension(sdim, int(values.shape[-1]))
        var = nc.createVariable(name, 'c', tuple(dims) + (sdim,))
    else:
        var = nc.createVariable(name, values.dtype, tuple(dims))
    var[:] = values
    return var


def save_netcdf(dataset_to_save, save_path=None, *args, **kwargs):
    """xarray.py:36-55 for a DataArray-like (``values``, ``dims``, ``indexes``): encode MultiIndex dimensions, write netCDF-3."""
    dim_coords, level_coords = encode_multidim(dataset_to_save)
    values = np.asarray(dataset_to_save.values)
    data, _ = _to_nc3(values)
    record_dim = None
    if data.nbytes > FIXED_VARIABLE_LIMIT:
        if data.ndim < 2 or data[0].nbytes > FIXED_VARIABLE_LIMIT:
            raise ValueError("the netCDF-3 weight store holds arrays whose slices along the first dimension stay below 2 GiB; "
                             "got shape {0} of {1}".format(data.shape, data.dtype))
        record_dim = dataset_to_save.dims[0]
    nc = netcdf_file(save_path, 'w', version=2)
    try:
        for dim, n in zip(dataset_to_save.dims, values.shape):
            nc.createDimension(dim, None if dim == record_dim else int(n))
        for dim, (coord, attrs) in dim_coords.items():
            coord, extra = _to_nc3(coord)
            var = _create(nc, dim, coord, (dim,))
            for key, val in dict(attrs, **extra).items():
                setattr(var, key, val)
        for name, (dim, coord) in level_coords.items():
            coord, extra = _to_nc3(coord)
            var = _create(nc, name, coord, (dim,))
            for key, val in extra.items():
                setattr(var, key, val)
        var = _create(nc, DATAARRAY_VARIABLE, data, tuple(dataset_to_save.dims))
        if level_coords:
            var.coordinates = ' '.join(level_coords.keys())
    finally:
        nc.close()
    return None


def decode_multidim(dims, coords, attrs):
    """xarray.py:137-173: dimensions whose coordinate carries ``multidim_levels`` get their MultiIndex back."""
    indexes = {}
    for dim in dims:
        if dim in attrs and 'multidim_levels' in attrs[dim]:
            names = attrs[dim]['multidim_levels'].split(';')
            indexes[dim] = pd.MultiIndex.from_arrays([coords[name] for name in names], names=names)      # xarray.py:163-166
        elif dim in coords:
            indexes[dim] = pd.Index(coords[dim], name=dim)
    return indexes


def load_netcdf(load_path, array=False, *args, **kwargs):
    """xarray.py:93-134 with ``array=True`` (the only form the weight store uses): -> DataArray with decoded MultiIndexes."""
    if not array:
        raise NotImplementedError("the weight store holds a single DataArray: call load_netcdf(path, array=True)")
    nc = netcdf_file(load_path, 'r', mmap=False)
    try:
        referenced = set()
        for var in nc.variables.values():
            referenced.update(str(getattr(var, 'coordinates', b'').decode()).split())
        data_names = [name for name in nc.variables if name not in nc.dimensions and name not in referenced
                      and not name.startswith('string')]
        if DATAARRAY_VARIABLE in nc.variables:
            data_name = DATAARRAY_VARIABLE
        elif len(data_names) == 1:
            data_name = data_names[0]
        else:
            raise ValueError('Given file dataset contains more than one data variable. Please read with '
                             'xarray.open_dataset and then select the variable or variables you want.')
        var = nc.variables[data_name]
        dims = tuple(var.dimensions)
        values = np.array(var[:], dtype=var[:].dtype.newbyteorder('='))
        coords, attrs = {}, {}
        for name, cvar in nc.variables.items():
            if name == data_name:
                continue
            cvals = np.array(cvar[:], dtype=cvar[:].dtype.newbyteorder('='))
            if cvals.dtype.kind == 'S' and cvals.ndim == 2:                  # netCDF-3 strings: (n, string_length) characters
                cvals = np.array([b''.join(row).decode('utf-8').rstrip('\x00') for row in cvals])
            cattrs = {k: (v.decode() if isinstance(v, bytes) else v) for k, v in cvar._attributes.items()}
            if str(cattrs.get('units', '')).startswith('seconds since 1970-01-01'):
                cvals = (np.datetime64('1970-01-01T00:00:00', 'ns') +
                         (cvals * 1e9).round().astype('int64').astype('timedelta64[ns]'))
            coords[name], attrs[name] = cvals, cattrs
    finally:
        nc.close()
    return DataArray(values, decode_multidim(dims, coords, attrs), dims)
