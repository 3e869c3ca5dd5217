"""Host-side logic added in round 2 (CPU): per-voxel temporal weights, projected output grids, CRS strictness."""
import datetime as dt

import numpy as np
import pytest

from oracle import refpy


def test_inverse_weights_match_the_reference_function():
    """raider_b200.s1_azimuth_timing.get_inverse_weights_for_dates against the reference's own (s1_azimuth_timing.py:337-399),
    imported unmodified (oracle/refpy.py), bit for bit; plus the reference's known answers (test/test_s1_time_grid.py:273-308)."""
    from raider_b200 import s1_azimuth_timing as mine
    dates = [dt.datetime(2021, 1, 1, 0), dt.datetime(2021, 1, 1, 6), dt.datetime(2020, 12, 31, 18)]
    grid = np.full((3, 4, 5), np.datetime64('2021-01-01T00:00:00'), dtype='datetime64[ms]') + (np.arange(60).reshape(3, 4, 5) * 337 - 9000).astype('timedelta64[s]')
    w = mine.get_inverse_weights_for_dates(grid, dates, temporal_window_hours=6)
    assert np.allclose(np.sum(w, axis=0), 1.0) and all(x.shape == grid.shape for x in w)
    # test_triple_date_usage of the reference
    tg = np.full(3, np.datetime64('2021-01-01T00:00:00'), dtype='datetime64[ms]') + np.timedelta64(1, 's') * np.array([-10_000, 0, 10_000])
    w0, w1, w2 = mine.get_inverse_weights_for_dates(tg, dates, temporal_window_hours=6, inverse_regularizer=1e-10)
    assert all(w0 > 0) and w1[0] <= 0 and w1[2] > 0 and w2[0] > 0 and w2[2] <= 0
    with pytest.raises(ValueError):
        mine.get_inverse_weights_for_dates(np.zeros((3, 3)), [dt.datetime(2023, 1, 1)] * 2)
    with pytest.raises(ValueError):
        mine.get_inverse_weights_for_dates(grid, [])
    with pytest.raises(ValueError, match='temporal window'):
        mine.get_inverse_weights_for_dates(grid, [dt.datetime(2019, 1, 1), dt.datetime(2019, 1, 2)], temporal_window_hours=1)
    # scalar weights of the center_time method (cli/raider.py:877-888; SURVEY 8d: 13:30 between 12:00 / 15:00, and 12:05)
    t0, t1 = dt.datetime(2020, 1, 30, 12), dt.datetime(2020, 1, 30, 15)
    assert mine.get_weights_time_interp([t0, t1], dt.datetime(2020, 1, 30, 13, 30)) == [0.5, 0.5]
    a, b = mine.get_weights_time_interp([t0, t1], dt.datetime(2020, 1, 30, 12, 5))
    assert abs(a - 175 / 180) < 1e-15 and abs(b - 5 / 180) < 1e-15
    if refpy.available():
        ref = refpy.load().s1_azimuth_timing
        if ref is None:
            pytest.skip('RAiDER.s1_azimuth_timing not importable under the stand-ins')
        for kw in (dict(temporal_window_hours=6), dict(temporal_window_hours=6, inverse_regularizer=1e-10), dict()):
            a = ref.get_inverse_weights_for_dates(grid, dates if kw else dates[:2], **kw)
            b = mine.get_inverse_weights_for_dates(grid, dates if kw else dates[:2], **kw)
            assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))


def test_projected_output_grid_and_crs_strictness():
    from raider_b200.crs import LambertConformalSphere, parse_crs
    from raider_b200.llreader import BoundingBox
    aoi = BoundingBox([35.0, 36.0, -99.0, -97.0])
    aoi.set_output_spacing(3000.0)
    lcc = LambertConformalSphere()
    aoi.set_output_xygrid(lcc)
    # the projected grid covers the projected corners of the box (utilFcns.transform_bbox: extremes of an 11 x 11 mesh)
    cx, cy = lcc.from_ll(np.array([-99.0, -97.0, -99.0, -97.0]), np.array([35.0, 35.0, 36.0, 36.0]))
    assert aoi.xpts[0] <= cx.min() and aoi.xpts[-1] >= cx.max() and aoi.ypts[0] >= cy.max() and aoi.ypts[-1] <= cy.min()
    assert np.allclose(np.diff(aoi.xpts), 3000.0) and np.allclose(np.diff(aoi.ypts), -3000.0)
    with pytest.raises(NotImplementedError):   # no explicit sphere: PROJ would use GRS80
        parse_crs('+proj=lcc +lat_1=38.5 +lat_2=38.5 +lat_0=38.5 +lon_0=262.5')
    with pytest.raises(NotImplementedError):   # another datum is not WGS-84
        parse_crs('+proj=longlat +datum=NAD27 +no_defs')
    with pytest.raises(NotImplementedError):
        parse_crs('+proj=longlat +ellps=clrk66 +no_defs')
    assert parse_crs('+proj=longlat +ellps=WGS84 +no_defs') == 4326
