"""Builds tpnet_b200/_C/libtpnet_b200.so from csrc/*.cu with nvcc for sm_100a.

In-tree, no JIT cache: the .so travels to the GPU box with the repo snapshot.
``python -m tpnet_b200.build [--force] [--verbose]``.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from typing import List

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, 'csrc')
OUT_DIR = os.path.join(PKG_DIR, '_C')
LIB_NAME = 'libtpnet_b200.so'
LIB_PATH = os.path.join(OUT_DIR, LIB_NAME)

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-std=c++17', '-lineinfo',
    '-Xcompiler', '-fPIC',
    '-Xcompiler', '-fopenmp',         # host side: the staging pass of big batches (tpn_stage.cu)
    '-Xptxas', '-v',
    # IEEE arithmetic everywhere: no fast-math, denormals kept, precise div/sqrt.
    '--ftz=false', '--prec-div=true', '--prec-sqrt=true',
]


def _extra_flags() -> List[str]:
    # profiling builds, e.g. TPN_EXTRA_NVCC_FLAGS=-DTPN_HUB2_TIMELINE
    return os.environ.get('TPN_EXTRA_NVCC_FLAGS', '').split()


def _nvcc() -> str:
    cand = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(cand):
        raise RuntimeError('nvcc not found: cannot build the tpnet_b200 CUDA library')
    return cand


def _sources() -> List[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = _sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh'))
    files.append(os.path.join(ROOT, 'include', 'tpnet_b200.h'))
    for f in files:
        h.update(f.encode())
        with open(f, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS + _extra_flags()).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if sources changed; returns the path of the shared library."""
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, 'build.stamp')
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == fp:
                return LIB_PATH
    nvcc = _nvcc()
    objs = []
    log = []
    for src in _sources():
        obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + '.o')
        cmd = [nvcc, *NVCC_FLAGS, *_extra_flags(), '-I', os.path.join(ROOT, 'include'), '-I', CSRC, '-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{r.stdout}\n{r.stderr}')
        objs.append(obj)
    cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-fopenmp', '-o', LIB_PATH, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    with open(os.path.join(OUT_DIR, 'ptxas.log'), 'w') as fh:
        fh.write('\n'.join(log))
    with open(stamp, 'w') as fh:
        fh.write(fp)
    if verbose:
        print('\n'.join(log))
    return LIB_PATH


if __name__ == '__main__':
    p = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(p)
