"""Oracle: target-grid construction (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates reference ``src/xarray_regrid/utils.py:18-44`` (``Grid`` validation) and
``utils.py:66-89`` (``create_lat_lon_coords``).
"""

from __future__ import annotations

import numpy as np


class InvalidBoundsError(Exception):
    """Mirrors reference utils.py:10."""


def check_bounds(north, east, south, west):
    """utils.py:29-44 -- south must not exceed north, west must not exceed east."""
    if south > north or west > east:
        raise InvalidBoundsError("grid bounds are inverted")


def lat_lon_coords(north, east, south, west, resolution_lat, resolution_lon):
    """utils.py:66-89.

    An axis whose span is an exact multiple of the resolution includes its upper
    end point.  Quirk kept on purpose: the reference tests the LATITUDE
    resolution when deciding whether the LONGITUDE axis is end-inclusive
    (utils.py:83).
    """
    check_bounds(north, east, south, west)

    def axis(lo, hi, step, test_step):
        inclusive = not (np.remainder(hi - lo, test_step) > 0)
        stop = hi + step if inclusive else hi
        return np.arange(lo, stop, step)

    lat = axis(south, north, resolution_lat, resolution_lat)
    lon = axis(west, east, resolution_lon, resolution_lat)
    return lat, lon
