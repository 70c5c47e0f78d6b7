"""ORACLE -- test infrastructure only.  ctypes wrapper of oracle/liboracle.so (CPU restatement of the
reference hot path, see oracle.h).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from safe_mpc_b200 import abi
from safe_mpc_b200.binding import EngineBase, SimBase

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_DIR, 'liboracle.so')
    srcs = [os.path.join(_DIR, f) for f in ('oracle.cpp', 'oracle.h', 'model.hpp', 'qp.hpp', 'dual.hpp')]
    srcs.append(os.path.join(_DIR, '..', 'include', 'safe_mpc_b200.h'))
    stale = force or not os.path.isfile(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.isfile(s))
    if stale and os.path.isfile(srcs[0]):
        subprocess.run(['make', '-C', _DIR, '-s'], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


class Oracle(EngineBase):
    prefix = 'orc_'
    has_mem = False

    def __init__(self, prob: abi.Problem, batch: int, threads: int = 0):
        super().__init__(lib(), prob, batch, threads)

    def num_threads(self):
        return int(self.lib.orc_num_threads(self.h))

    def set_mlp_fp32(self, on):
        self._call('set_mlp_fp32', C.c_int32(int(on)))

    def set_qp_hook(self, fn_ptr):
        """tests only: every QP of this handle is solved by an external function (oracle.h orc_qp_hook_t), e.g. tests/emu's
        emu_qp_solve1 = the engine's kernel sources compiled for the host.  fn_ptr: ctypes function object or None."""
        self._hook = fn_ptr
        self._call('set_qp_hook', C.cast(fn_ptr, C.c_void_p) if fn_ptr is not None else C.c_void_p(0))

    def set_probe(self, eps):
        """tests only: repeat every solve on data perturbed by a relative eps and count the status flips per problem"""
        self._call('set_probe', C.c_double(eps))

    def probe_flips(self):
        out = np.zeros(self.B, dtype=np.int32)
        self._call('get_probe_flips', C.c_void_p(out.ctypes.data))
        return out

    def mass_bias(self, b, x, nominal=True):
        M = np.empty((abi.NQ, abi.NQ)); h = np.empty(abi.NQ)
        xx = np.ascontiguousarray(x, dtype=np.float64)
        self._call('mass_bias', C.c_int32(b), C.c_int32(int(nominal)), C.c_void_p(xx.ctypes.data),
                   C.c_void_p(M.ctypes.data), C.c_void_p(h.ctypes.data))
        return M, h

    def controller_step_scripted(self, x, status, x_temp, u_temp):
        """controller.step(x) with the solve replaced by a given outcome (tests/test_ref_golden.py) -> (u [B, nu], abort [B])"""
        x = np.ascontiguousarray(x, dtype=np.float64); st = np.ascontiguousarray(status, dtype=np.int32)
        xt = np.ascontiguousarray(x_temp, dtype=np.float64); ut = np.ascontiguousarray(u_temp, dtype=np.float64)
        u = np.zeros((self.B, abi.NU)); ab = np.zeros(self.B, dtype=np.uint8)
        self._call('controller_step_scripted', C.c_void_p(x.ctypes.data), C.c_void_p(st.ctypes.data), C.c_void_p(xt.ctypes.data),
                   C.c_void_p(ut.ctypes.data), C.c_void_p(u.ctypes.data), C.c_void_p(ab.ctypes.data))
        return u, ab.astype(bool)

    def qp_info(self, b):
        res = (C.c_double * 4)(); mu = C.c_double(); it = C.c_int32(); st = C.c_int32()
        self._call('qp_info', C.c_int32(b), res, C.byref(mu), C.byref(it), C.byref(st))
        return np.array(res[:]), mu.value, it.value, st.value


class OracleSim(SimBase):
    def set_script(self, status, xt, ut, bk_status, bk_xt, bk_ut):
        """tests only: the solves of the following steps are replaced by these outcomes (arrays are kept alive here)"""
        self._script = [np.ascontiguousarray(status, dtype=np.int32), np.ascontiguousarray(xt, dtype=np.float64), np.ascontiguousarray(ut, dtype=np.float64),
                        np.ascontiguousarray(bk_status, dtype=np.int32), np.ascontiguousarray(bk_xt, dtype=np.float64), np.ascontiguousarray(bk_ut, dtype=np.float64)]
        self._call('set_script', *[C.c_void_p(a.ctypes.data) for a in self._script])
