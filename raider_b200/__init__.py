"""raider_b200 -- B200-native (sm_100a) slant/zenith tropospheric-delay ray tracer behind RAiDER's API.

Module map mirrors the reference package for the hot path only:

    raider_b200.delay         tropo_delay, _get_delays_on_cube, _build_cube, _build_cube_ray   (tools/RAiDER/delay.py)
    raider_b200.delayFcns     getInterpolators                                                 (tools/RAiDER/delayFcns.py)
    raider_b200.losreader     LOS classes, getTopOfAtmosphere, build_ray                       (tools/RAiDER/losreader.py)
    raider_b200.interpolate   interpolate, interpolate_along_axis                              (tools/bindings/interpolate)
    raider_b200.interpolator  RegularGridInterpolator, interp_along_axis                       (tools/RAiDER/interpolator.py)
    raider_b200.makePoints    makePoints0D..3D                                                 (tools/bindings/utils/makePoints.pyx)
    raider_b200.utilFcns      lla2ecef, ecef2lla, enu2ecef, ecef2enu, sind, cosd               (tools/RAiDER/utilFcns.py:67-137)
    raider_b200.dist          raster sharding over the GPUs of one node (torch.distributed / NCCL)

All numerics run in ``libraider_b200.so`` (raider_b200/csrc, C ABI in include/raider_b200.h); importing this package
does not load it, calling anything does, and a missing library or GPU is an error, never a fallback.
"""
__version__ = '0.1.0'
