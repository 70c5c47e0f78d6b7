"""Builds the CUDA library in-tree: ``safe_mpc_b200/csrc/libsafe_mpc_b200.so`` (sm_100a only).

Plays the role of the reference's ``build_controller(build=True)`` (controller.py:243-248: acados code generation
+ gcc): here there is nothing to generate, the kernels are compiled once for every controller.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'csrc')
LIB = os.path.join(CSRC, 'libsafe_mpc_b200.so')
SOURCES = ['api.cu', 'kernels.cu', 'qp.cu', 'qp_f32.cu', 'mlp_tc.cu', 'mlp_tc2.cu', 'peaks.cu']
HEADERS = ['engine.cuh', 'qp_split.cuh', 'qp_tail.cuh', 'dev_model.cuh', 'mlp_tc_common.cuh', os.path.join('..', '..', 'include', 'safe_mpc_b200.h')]
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-Xcompiler', '-fPIC',
              '-Xptxas', '-v']


def nvcc_path():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isfile(cand) or cand == 'nvcc'):
            return cand
    return 'nvcc'


def is_stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS if os.path.isfile(os.path.join(CSRC, f)))


def build_cuda(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    nvcc = nvcc_path()
    log = []

    def compile_one(src):
        obj = os.path.join(CSRC, src[:-3] + '.o')
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get('SMPC_NVCC_EXTRA', '').split() + ['-c', os.path.join(CSRC, src), '-o', obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log.append((src, res.stderr))
        if res.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{res.stderr}')
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    res = subprocess.run([nvcc, '-shared', '-o', LIB] + objs + ['-lcudart'], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'link failed:\n{res.stderr}')
    with open(os.path.join(CSRC, 'build.log'), 'w') as fh:
        for src, err in log:
            fh.write(f'==== {src}\n{err}\n')
    if verbose:
        for src, err in log:
            print(f'==== {src}\n{err}')
    return LIB


if __name__ == '__main__':
    print(build_cuda(force='--force' in sys.argv, verbose=True))
