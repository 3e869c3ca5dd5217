"""Minimal AOI (query-region) layer: inputs of the delay path (reference: tools/RAiDER/llreader.py).

The reference's AOI classes read bounding boxes, station CSVs, radar rasters and geocubes through pandas/rasterio
(out of scope).  ``tropo_delay`` only uses a small duck type -- ``xpts``/``ypts``,
``set_output_spacing``/``set_output_xygrid``, ``readLL``/``readZ`` and "is it a cube AOI?" -- which these two classes
provide without any I/O dependency.  Reference AOI objects work too (they are recognised by class name).
"""
from __future__ import annotations

import numpy as np


class AOI:
    def __init__(self) -> None:
        self._output_spacing = None
        self._bounding_box = None
        self._type = None

    def type(self):
        return self._type

    def bounds(self):
        return list(self._bounding_box).copy()

    def set_output_spacing(self, ll_res=None) -> None:
        """llreader.py:91-104 for a geographic output grid (degrees)."""
        self._output_spacing = ll_res

    def get_output_spacing(self, crs=4326):
        return self._output_spacing

    def set_output_xygrid(self, dst_crs=4326) -> None:
        """llreader.py:173-191: xpts ascending, ypts descending, end points inclusive.  A projected ``dst_crs`` gets the bounding
        box transformed the reference's way (utilFcns.transform_bbox: an 11 x 11 mesh over the box, buffered by 100 m worth of
        degrees, projected; the extremes of the projected mesh)."""
        from .crs import Geographic, parse_crs
        crs = parse_crs(dst_crs)
        S, N, W, E = self.bounds()
        if not isinstance(crs, Geographic):
            buffer = 100.0 / 1.0e5
            X, Y = np.meshgrid(np.linspace(W - buffer, E + buffer, num=11), np.linspace(S - buffer, N + buffer, num=11))
            xx, yy = crs.from_ll(X, Y)
            S, N, W, E = np.nanmin(yy), np.nanmax(yy), np.nanmin(xx), np.nanmax(xx)
        sp = self.get_output_spacing(dst_crs)
        self.xpts = np.arange(W, E + sp, sp)
        self.ypts = np.arange(N, S - sp, -sp)


class BoundingBox(AOI):
    """Parse a bounding box AOI [S, N, W, E] (llreader.py BoundingBox)."""

    def __init__(self, bbox, spacing=None) -> None:
        super().__init__()
        self._bounding_box = [float(b) for b in bbox]
        self._type = 'bounding_box'
        if spacing is not None:
            self.set_output_spacing(spacing)
            self.set_output_xygrid()


class Points(AOI):
    """Explicit query points (the role of StationFile / RasterRDR / GeocodedFile in the reference)."""

    def __init__(self, lats, lons, hgts, pad=0.5) -> None:
        super().__init__()
        self._lats = np.asarray(lats, dtype=np.float64)
        self._lons = np.asarray(lons, dtype=np.float64)
        self._hgts = np.asarray(hgts, dtype=np.float64)
        self._bounding_box = [np.nanmin(self._lats) - pad, np.nanmax(self._lats) + pad, np.nanmin(self._lons) - pad, np.nanmax(self._lons) + pad]
        self._type = 'points'

    def readLL(self):
        return self._lats, self._lons

    def readZ(self):
        return self._hgts


def is_cube_aoi(aoi) -> bool:
    """``isinstance(aoi, (BoundingBox, Geocube))`` of delay.py:98, by duck type so reference AOIs qualify too."""
    return isinstance(aoi, BoundingBox) or type(aoi).__name__ in ('BoundingBox', 'Geocube')
