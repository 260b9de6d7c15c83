"""Build libogbsampler.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
SRC = os.path.join(PKG_DIR, 'csrc', 'ogb_sampler.cu')
DEPS = [
    SRC,
    os.path.join(PKG_DIR, 'csrc', 'device_common.cuh'),
    os.path.join(PKG_DIR, 'csrc', 'relabel_rows.cuh'),
    os.path.join(PKG_DIR, 'csrc', 'gather_frames.cuh'),
    os.path.join(REPO_ROOT, 'include', 'ogb_sampler.h'),
]
LIB_PATH = os.path.join(PKG_DIR, 'libogbsampler.so')

NVCC_FLAGS = [
    '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
    '-Xcompiler', '-fPIC', '-shared',
]


def find_nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: the sampler has no CPU fallback and cannot be built without the CUDA toolkit')


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if it is missing or older than its sources; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS, '-o', LIB_PATH, SRC]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
        print(' '.join(cmd), file=sys.stderr)
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f'nvcc failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}')
    if verbose:
        print(proc.stderr, file=sys.stderr)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
