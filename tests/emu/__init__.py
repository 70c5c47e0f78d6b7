"""TEST HARNESS: host build of the engine's kernel sources (emu.cpp).  Never imported by the product."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = {}


def load(f32=False):
    """Build (if stale) and dlopen tests/emu/libsmpc_emu.so; f32: the fp32-storage flavour of the QP sources (qp_split.cuh QS_REAL)."""
    if f32 not in _LIB:
        so = os.path.join(HERE, 'libsmpc_emu_f32.so' if f32 else 'libsmpc_emu.so')
        csrc = os.path.join(HERE, '..', '..', 'safe_mpc_b200', 'csrc')
        srcs = [os.path.join(HERE, 'emu.cpp')] + [os.path.join(csrc, f) for f in ('dev_model.cuh', 'qp_split.cuh')]
        if not os.path.isfile(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            flav = ['-DQS_REAL=float', '-DQS_FLAVOUR=f32'] if f32 else []
            subprocess.run(['g++', '-O2', '-std=c++17', '-DEMU_QP', '-fPIC', '-shared'] + flav + ['-o', so, srcs[0]], check=True)
        _LIB[f32] = C.CDLL(so)
    return _LIB[f32]


def kernel_source_oracle(prob, batch, threads=0, f32=False):
    """An Oracle whose QPs are solved by the engine's kernel sources on the host (emu_qp_solve1): the CPU stand-in of the GPU
    engine for whole closed loops.  Linearisation, controller logic and plant stay the oracle's."""
    from oracle.oracle import Oracle
    o = Oracle(prob, batch, threads)
    o.set_qp_hook(load(f32).emu_qp_solve1)
    return o
