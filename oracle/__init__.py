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

Pin status (details in DESIGN.md): sampling, makePoints and interpolate* are pinned against the
compiled reference natives and the reference's own vectors; the ray tracer is pinned by the
reference's constant-refractivity identity (test/test_synthetic.py:217-274) and analytic
checks; sub-micrometre parity with PROJ and isce3 look vectors is UNPINNED (neither library is
available offline).
"""
