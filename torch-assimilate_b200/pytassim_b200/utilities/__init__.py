from .netcdf import save_netcdf, load_netcdf, encode_multidim, decode_multidim, DATAARRAY_VARIABLE

__all__ = ['save_netcdf', 'load_netcdf', 'encode_multidim', 'decode_multidim', 'DATAARRAY_VARIABLE']
