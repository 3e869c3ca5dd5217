"""Build libraider_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m raider_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / 'csrc'
OUT = PKG_DIR / 'libraider_b200.so'
SOURCES = [CSRC / 'raider_b200.cu']
HEADERS = sorted(CSRC.glob('*.cuh')) + [PKG_DIR.parent / 'include' / 'raider_b200.h']   # every fragment of the translation unit is a dependency

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-shared', '-Xcompiler', '-fPIC,-fvisibility=hidden',
    '-Xptxas', '-v',
]


def nvcc() -> str:
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not Path(exe).exists():
        raise RuntimeError('nvcc not found: cannot build libraider_b200.so')
    return exe


def up_to_date() -> bool:
    if not OUT.exists():
        return False
    t = OUT.stat().st_mtime
    return all(p.stat().st_mtime <= t for p in SOURCES + HEADERS + [Path(__file__)])


def build(force: bool = False, verbose: bool = False) -> Path:
    if up_to_date() and not force:
        return OUT
    cmd = [nvcc(), *NVCC_FLAGS, '-o', str(OUT), *map(str, SOURCES)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f'nvcc failed ({res.returncode}): {" ".join(cmd)}')
    (PKG_DIR / 'build_ptxas.log').write_text(res.stdout + res.stderr)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
