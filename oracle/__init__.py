"""CPU oracle for the RAiDER slant/zenith delay hot path -- TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` restates, on the CPU (NumPy + the installed scipy), the
algorithm of the reference path (``/root/reference`` = dbekaert/RAiDER @ e38c4eb4):

* ``oracle.geodesy``   WGS-84 geodetic<->ECEF  (reference: tools/RAiDER/utilFcns.py:77-137 which
                       delegates to PROJ -- not vendored; PROJ's published ``cart`` conversion
                       algorithm is restated, see the module header)
* ``oracle.raytrace``  getTopOfAtmosphere / build_ray / _build_cube_ray / _build_cube
                       (tools/RAiDER/losreader.py:706-733,772-835; tools/RAiDER/delay.py:196-326;
                       sampling through the *installed* scipy RegularGridInterpolator exactly as
                       tools/RAiDER/delayFcns.py:55-56 configures it)
* ``oracle.interp``    RAiDER.interpolate.{interpolate, interpolate_along_axis} and
                       RAiDER.makePoints.makePoints{0..3}D semantics
                       (tools/bindings/interpolate/src/*, tools/bindings/utils/makePoints.pyx)
* ``oracle.build_ref`` recipe that compiles the reference's own native sources into
                       ``oracle/_ref`` (used to validate the restatement above)

Only ``tests/``, ``__graft_entry__.smoke()`` and the cpu-baseline / ``--impl reference`` legs of
``bench.py`` may import this package, and only as the checker or the timed CPU baseline.  The
product (``raider_b200``) never imports it and has no CPU fallback.

Pin status (details in DESIGN.md section 7): sampling, makePoints and interpolate* are pinned bit for bit against the
compiled reference natives and the reference's own vectors; the ray tracer (``oracle.raytrace``) is pinned BIT FOR BIT
against the reference's own Python -- ``oracle.refpy`` imports RAiDER.delay / losreader / delayFcns / utilFcns unmodified
from /root/reference under stand-ins for pyproj / xarray / rasterio / shapely, and tests/test_oracle_vs_reference_py.py
compares every function; ``oracle.geodesy``'s spherical Lambert + ``build_cube`` reproduce the reference's own golden
(test/test_HRRR_ztd.py:18) from the reference's real HDF5 cube to the 7 decimals the reference asserts; ``oracle.orbit`` (isce3's
Hermite interpolation + zero-Doppler solve) standing in for isce3 under the reference's own Raytracing class reproduces the
reference's end-to-end ray-tracing golden (test/test_slant.py:99) to 5e-8 m.  What stays UNPINNED: PROJ's own rounding of the
WGS-84 ``cart`` inverse below ~1e-9 m and isce3's rounding below ~2e-8 rad of look direction -- neither library is available
offline.
"""
