"""Target-grid helpers with the reference's names and behaviour (src/xarray_regrid/utils.py:10-114)."""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np


class InvalidBoundsError(Exception):
    """utils.py:10."""


@dataclass
class Grid:
    """Bounds and resolution of a rectilinear lat/lon grid (utils.py:18-63)."""

    north: float
    east: float
    south: float
    west: float
    resolution_lat: float
    resolution_lon: float

    def __post_init__(self) -> None:
        msg = None
        if self.south > self.north:
            msg = "Value of north bound is greater than south bound.\nPlease check the bounds input."
        if self.west > self.east:
            msg = "Value of west bound is greater than east bound.\nPlease check the bounds input."
        if msg is not None:
            raise InvalidBoundsError(msg)

    def create_regridding_dataset(self, lat_name: str = "latitude", lon_name: str = "longitude"):
        return create_regridding_dataset(self, lat_name, lon_name)


def create_lat_lon_coords(grid: Grid) -> tuple[np.ndarray, np.ndarray]:
    """utils.py:66-89.  An axis whose span is a multiple of the resolution includes its end point; the
    reference tests the LATITUDE resolution for the longitude axis too (utils.py:83) -- kept."""

    def axis(lo, hi, step, test_step):
        if np.remainder(hi - lo, test_step) > 0:
            return np.arange(lo, hi, step)
        return np.arange(lo, hi + step, step)

    return (axis(grid.south, grid.north, grid.resolution_lat, grid.resolution_lat),
            axis(grid.west, grid.east, grid.resolution_lon, grid.resolution_lat))


def create_regridding_dataset(grid: Grid, lat_name: str = "latitude", lon_name: str = "longitude"):
    """utils.py:92-114: a coordinate-only dataset (``labelled.Dataset``; converted to ``xarray.Dataset`` by
    ``labelled.to_xarray`` when xarray is installed)."""
    from .labelled import Dataset

    lat, lon = create_lat_lon_coords(grid)
    return Dataset(
        data_vars={},
        coords={lat_name: lat, lon_name: lon},
        coord_attrs={lat_name: {"units": "degrees_north"}, lon_name: {"units": "degrees_east"}},
    )
