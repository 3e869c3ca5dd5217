"""ctypes binding of libraider_b200.so (include/raider_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``python -m raider_b200.build``.  There is no
fallback of any kind: if the shared object is missing, ``load()`` raises; if no CUDA device is present,
``rdr_create`` fails and :class:`Handle` raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / 'libraider_b200.so'

# enums of include/raider_b200.h
RDR_OK, RDR_ERR_INVALID, RDR_ERR_CUDA, RDR_ERR_STATE, RDR_ERR_NO_LAYERS, RDR_ERR_ALL_NAN = range(6)
MEM_HOST, MEM_DEVICE = 0, 1
F64, F32 = 0, 1
LAYOUT_ZYX, LAYOUT_YXZ = 0, 1
CRS_GEOGRAPHIC, CRS_LCC_SPHERE = 0, 1
GEOM_GRID, GEOM_POINTS = 0, 1
LOS_ARRAY, LOS_ENU_CONST, LOS_ZENITH, LOS_ORBIT, LOS_ENU_ARRAY = 0, 1, 2, 3, 4
SEM_SCIPY, SEM_RAIDER_FILL, SEM_RAIDER_CLAMP = 0, 1, 2
PLAN_ABSURD, PLAN_ALL_NAN, PLAN_KNIFE_EDGE, PLAN_SPAN_TOO_LONG = 1, 2, 4, 8
TRACE_EXACT_K0, TRACE_NO_KNIFE_GUARD = 1, 2
K3_AUTO, K3_FAST, K3_GENERAL = 0, 1, 2
ABI_VERSION = 3

_i64, _f64, _int, _vp = C.c_int64, C.c_double, C.c_int, C.c_void_p
_pd = C.POINTER(C.c_double)
_pi64 = C.POINTER(C.c_int64)

# name -> (restype, argtypes); the complete export list of include/raider_b200.h
SIGNATURES = {
    'rdr_abi_version': (_int, []),
    'rdr_device_count': (_int, []),
    'rdr_create': (_int, [_int, C.POINTER(_vp)]),
    'rdr_destroy': (_int, [_vp]),
    'rdr_last_error': (C.c_char_p, [_vp]),
    'rdr_set_stream': (_int, [_vp, _vp]),
    'rdr_synchronize': (_int, [_vp]),
    'rdr_host_alloc': (_int, [_i64, C.POINTER(_vp)]),
    'rdr_host_free': (_int, [_vp]),
    'rdr_launch_count': (_i64, [_vp]),
    'rdr_last_fix_count': (_i64, [_vp]),
    'rdr_set_cube': (_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _int, _int, _vp, _int]),
    'rdr_blend_cube': (_int, [_vp, _vp, _vp, _int, _f64, _f64, _int]),
    'rdr_sample': (_int, [_vp, _vp, _i64, _vp, _vp, _int, _int, _int]),
    'rdr_sample_grid': (_int, [_vp, _vp, _i64, _vp, _i64, _f64, _vp, _vp, _int]),
    'rdr_sample_grid_levels': (_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _int]),
    'rdr_ray_plan': (_int, [_vp, _f64, _f64, _pi64, _vp, _vp]),
    'rdr_ray_layers': (_int, [_vp, _int, _vp, _vp, _i64, _i64, _int, _vp, _f64, _f64, _vp, _vp, _int]),
    'rdr_ray_integrate': (_int, [_vp, _vp, _f64, _int, _vp, _vp, _int, _int, _vp, _vp, _int]),
    'rdr_set_peer_outputs': (_int, [_vp, _int, _vp, _vp]),
    'rdr_set_multicast_outputs': (_int, [_vp, _vp, _vp]),
    'rdr_trace_begin': (_int, [_vp, _int, _vp, _vp, _i64, _i64, _int, _vp, _f64, _f64, _int, _int]),
    'rdr_trace_finish': (_int, [_vp, _f64, _int, _int, _vp, _vp, _int, _int, _int]),
    'rdr_trace_result': (_int, [_vp, _vp, _vp, _vp]),
    'rdr_set_exchange': (_int, [_vp, _int, _int, _vp]),
    'rdr_exchange_bytes': (_i64, [_int]),
    'rdr_ray_stations': (_int, [_vp, _vp, _vp, _vp, _i64, _int, _vp, _f64, _f64, _vp, _vp, _vp, _int]),
    'rdr_ray_points': (_int, [_vp, _vp, _f64, _i64, _i64, _vp, _int, _pi64, _int]),
    'rdr_top_of_atmosphere': (_int, [_vp, _vp, _i64, _f64, _vp, _vp, _int]),
    'rdr_build_ray': (_int, [_vp, _i64, _f64, _vp, _vp, _i64, _f64, _pi64, _vp, _vp, _vp, _int]),
    'rdr_lla2ecef': (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _int]),
    'rdr_ecef2lla': (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _int]),
    'rdr_make_points_count': (_int, [_f64, _f64, _pi64]),
    'rdr_make_points': (_int, [_f64, _vp, _vp, _i64, _f64, _vp, _i64, _int, _int]),
    'rdr_interpolate': (_int, [_int, C.POINTER(_vp), _pi64, _vp, _vp, _i64, _int, _f64, _vp, _int, _int]),
    'rdr_prepare_cube': (_int, [_i64, _i64, _vp, _vp, _vp, _vp, _int, _vp, _i64, _f64, _f64, _f64, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                _pi64, _int, _int]),
    'rdr_orbit_los': (_int, [_vp, _i64, _int, _vp, _vp, _vp, _f64, _i64, _i64, _f64, _int, _vp, _vp, _vp, _int]),
    'rdr_selftest_div': (_int, [_i64, C.c_uint64, _pi64, _pi64, _int]),
    'rdr_interp_along_axis': (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _f64, _vp, _int, _int]),
}

_lib = None


def load():
    """Load libraider_b200.so (once).  Raises ImportError when it has not been built -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get('RAIDER_B200_LIB', LIB_PATH))
    if not path.exists():
        raise ImportError(
            f'{path} is missing: build it with `python -m raider_b200.build` (nvcc, sm_100a). '
            'raider_b200 has no CPU fallback.'
        )
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.rdr_abi_version() != ABI_VERSION:
        raise ImportError(f'{path}: ABI version {lib.rdr_abi_version()} != {ABI_VERSION}')
    _lib = lib
    return lib


class RaiderB200Error(RuntimeError):
    pass


class NoLayersError(RaiderB200Error):
    """No model layer contributes (losreader.build_ray returns (None, None, None), losreader.py:832-833)."""


def check(rc: int, handle=None):
    """Map C-ABI status codes to the exception types of the reference boundary."""
    if rc == RDR_OK:
        return
    msg = load().rdr_last_error(handle)
    msg = msg.decode() if msg else f'status {rc}'
    if rc == RDR_ERR_INVALID:
        raise TypeError(msg)
    if rc == RDR_ERR_NO_LAYERS:
        raise NoLayersError(msg)
    if rc == RDR_ERR_ALL_NAN:
        raise ValueError(msg)  # delay.py:279-280
    raise RaiderB200Error(msg)


def ptr(a):
    """Raw address of a C-contiguous numpy array, a torch CUDA/CPU tensor, an int address, or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags['C_CONTIGUOUS']:
            raise ValueError('array must be C-contiguous')
        return a.ctypes.data
    if isinstance(a, int):
        return a
    if hasattr(a, 'data_ptr'):
        if not a.is_contiguous():
            raise ValueError('tensor must be contiguous')
        return a.data_ptr()
    raise TypeError(f'cannot take the address of {type(a)}')


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """An uninitialised ndarray in page-locked host memory from the library's pool (rdr_host_alloc).

    The block goes back to the pool when the last view of the array dies.  Result arrays allocated this way are written by
    the integration kernel directly (no staging copy); to every other consumer they are ordinary NumPy arrays.
    """
    import weakref
    lib = load()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    p = _vp()
    check(lib.rdr_host_alloc(nbytes, C.byref(p)))
    buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
    weakref.finalize(buf, lib.rdr_host_free, p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)


def is_device(a) -> bool:
    return hasattr(a, 'data_ptr') and getattr(a, 'is_cuda', False)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def default_device() -> int:
    """LOCAL_RANK under torchrun, else RAIDER_B200_DEVICE, else 0."""
    for key in ('RAIDER_B200_DEVICE', 'LOCAL_RANK'):
        if key in os.environ:
            return int(os.environ[key])
    return 0


class Handle:
    """RAII wrapper around rdr_handle_t: one device, one stream, one staged cube, one set of rays."""

    def __init__(self, device: int | None = None) -> None:
        self.lib = load()
        self.device = default_device() if device is None else int(device)
        h = _vp()
        check(self.lib.rdr_create(self.device, C.byref(h)))
        self._h = h

    def close(self) -> None:
        if getattr(self, '_h', None):
            self.lib.rdr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- thin, typed entry points -------------------------------------------------------------
    def call(self, name: str, *args):
        check(getattr(self.lib, name)(self._h, *args), self._h)

    def set_stream(self, cuda_stream: int | None) -> None:
        self.call('rdr_set_stream', cuda_stream)

    def synchronize(self) -> None:
        self.call('rdr_synchronize')

    @property
    def launches(self) -> int:
        return int(self.lib.rdr_launch_count(self._h))

    @property
    def last_fix_count(self) -> int:
        return int(self.lib.rdr_last_fix_count(self._h))

    def set_cube(self, ys, xs, zs, wet, hydro, layout=LAYOUT_ZYX, crs_kind=CRS_GEOGRAPHIC, crs_params=None) -> None:
        ys, xs, zs = f64(ys), f64(xs), f64(zs)
        dev = is_device(wet)
        if not dev:
            # float32 at the ABI (include/raider_b200.h); float64 callers go through DeviceCube, which splits hi + lo
            wet = np.ascontiguousarray(wet, dtype=np.float32)
            hydro = np.ascontiguousarray(hydro, dtype=np.float32)
        n = ys.size * xs.size * zs.size
        if int(np.prod(wet.shape)) != n or int(np.prod(hydro.shape)) != n:
            raise TypeError(f'cube fields have {tuple(wet.shape)} values but the axes describe {zs.size}x{ys.size}x{xs.size}')
        cp = None if crs_params is None else f64(crs_params)
        self._keep = (ys, xs, zs, cp)
        self.call('rdr_set_cube', ptr(ys), ys.size, ptr(xs), xs.size, ptr(zs), zs.size, ptr(wet), ptr(hydro), layout, crs_kind,
                  ptr(cp), MEM_DEVICE if dev else MEM_HOST)
        self.grid_yxz = (ys, xs, zs)

    def blend_cube(self, wet1, hydro1, w0: float, w1: float, layout=LAYOUT_ZYX) -> None:
        dev = is_device(wet1)
        if not dev:
            wet1 = np.ascontiguousarray(wet1, dtype=np.float32)
            hydro1 = np.ascontiguousarray(hydro1, dtype=np.float32)
        self.call('rdr_blend_cube', ptr(wet1), ptr(hydro1), layout, float(w0), float(w1), MEM_DEVICE if dev else MEM_HOST)
