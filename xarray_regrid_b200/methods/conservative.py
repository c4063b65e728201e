"""Mirror of src/xarray_regrid/methods/conservative.py:39-101 (``conservative_regrid``)."""

from ..regrid import Regridder


def conservative_regrid(data, target_ds, latitude_coord, skipna=True, nan_threshold=1.0, output_chunks=None):
    return Regridder(data).conservative(target_ds, latitude_coord=latitude_coord, time_dim=None,
                                        skipna=skipna, nan_threshold=nan_threshold,
                                        output_chunks=output_chunks)
