"""Build recipe for ``oracle/_ref`` -- TEST INFRASTRUCTURE, not product code.

Compiles the reference's two native extensions *from the sources where they lie*
under ``/root/reference`` (nothing is copied into this repository) into
``oracle/_ref/``:

* ``interpolate``  <- tools/bindings/interpolate/src/{module.cpp,interpolate.cpp}
  (pybind11; the reference builds it in setup.py:20-33)
* ``makePoints``   <- tools/bindings/utils/makePoints.pyx (cython; setup.py:35-40)

The reference's own build system (setuptools) is NOT run; this is the short
recipe ``g++``/``cython`` on those files directly.  The resulting ``.so`` files
are git-ignored but travel to the GPU box, where ``/root/reference`` does not
exist.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
cpu-baseline legs may load them (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import importlib.util
import os
import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_OUT = HERE / '_ref'
REFERENCE = Path(os.environ.get('RAIDER_REFERENCE', '/root/reference'))
EXT = sysconfig.get_config_var('EXT_SUFFIX')


def _includes():
    import numpy
    import pybind11
    return [f'-I{sysconfig.get_paths()["include"]}', f'-I{pybind11.get_include()}', f'-I{numpy.get_include()}']


def ref_paths():
    return {'interpolate': REF_OUT / f'interpolate{EXT}', 'makePoints': REF_OUT / f'makePoints{EXT}'}


def available() -> bool:
    return all(p.exists() for p in ref_paths().values())


def build(force: bool = False, verbose: bool = False) -> bool:
    """Build oracle/_ref if the reference sources are present. Returns availability."""
    if available() and not force:
        return True
    src = REFERENCE / 'tools' / 'bindings'
    if not src.exists():
        return available()
    REF_OUT.mkdir(exist_ok=True)
    out = ref_paths()
    run = lambda cmd: subprocess.run(cmd, check=True, capture_output=not verbose)
    # 1) pybind11 interpolator (C++17, same flags family as the reference: -O3, no -march)
    run(['g++', '-O3', '-std=c++17', '-shared', '-fPIC', '-w', *_includes(),
         str(src / 'interpolate' / 'src' / 'module.cpp'), str(src / 'interpolate' / 'src' / 'interpolate.cpp'),
         '-o', str(out['interpolate']), '-pthread'])
    # 2) cython ray-point generator: generated C goes to _ref/, never into the reference tree
    c_file = REF_OUT / 'makePoints.c'
    run([sys.executable, '-m', 'cython', '-3', str(src / 'utils' / 'makePoints.pyx'), '-o', str(c_file)])
    run(['gcc', '-O3', '-shared', '-fPIC', '-w', '-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION', *_includes(),
         str(c_file), '-o', str(out['makePoints'])])
    c_file.unlink()  # generated from reference source: keep no copy of it around
    return available()


def load(name: str):
    """Import one of the compiled reference modules ('interpolate' | 'makePoints')."""
    path = ref_paths()[name]
    if not path.exists():
        raise ImportError(f'oracle/_ref/{path.name} not built (run python -m oracle.build_ref)')
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    ok = build(force='--force' in sys.argv, verbose=True)
    print('oracle/_ref available:', ok, {k: str(v) for k, v in ref_paths().items()})
