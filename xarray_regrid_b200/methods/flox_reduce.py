"""Mirror of src/xarray_regrid/methods/flox_reduce.py:40-187 (``statistic_reduce``, ``compute_mode``)."""

from ..regrid import Regridder


def statistic_reduce(data, target_ds, time_dim, method, skipna=False, fill_value=None):
    return Regridder(data).stat(target_ds, method, time_dim=time_dim, skipna=skipna, fill_value=fill_value)


def compute_mode(data, target_ds, values, time_dim, fill_value=None, anti_mode=False):
    r = Regridder(data)
    fn = r.least_common if anti_mode else r.most_common
    return fn(target_ds, values, time_dim=time_dim, fill_value=fill_value)
