"""Product binding: the CUDA engine behind include/safe_mpc_b200.h.

``Engine`` = one batched OCP solver handle on one GPU (what ``AcadosOcpSolver`` is to the reference's
controllers, controller.py:247, but for B problems at once); ``Sim`` = the closed loop of scripts/mpc.py over the
batch.  Arrays are numpy (host; the library copies) or torch CUDA tensors (device pointers, zero copy).
There is no CPU path: a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import abi
from .binding import EngineBase, SimBase, _is_torch

_LIB = None


def library_path():
    if os.environ.get('SMPC_LIB'):          # development: a variant build of the same library (scripts/build_variant.sh)
        return os.environ['SMPC_LIB']
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), 'csrc', 'libsafe_mpc_b200.so')


def load_library():
    """dlopen the in-tree CUDA library; fail loudly when it has not been built."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.isfile(path):
            raise RuntimeError(f'{path} is missing: build it with `python -m safe_mpc_b200.build` '
                               '(or __graft_entry__.build()); the engine has no CPU fallback')
        lib = C.CDLL(path)
        lib.smpc_version.restype = C.c_char_p
        lib.smpc_last_error.restype = C.c_char_p
        lib.smpc_launch_count.restype = C.c_int64
        lib.smpc_launch_count.argtypes = [C.c_void_p]
        lib.smpc_stream.restype = C.c_void_p
        lib.smpc_stream.argtypes = [C.c_void_p]
        lib.smpc_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        _LIB = lib
    return _LIB


class Engine(EngineBase):
    prefix = 'smpc_'
    has_mem = True

    def __init__(self, prob: abi.Problem, batch: int, device: int = 0):
        super().__init__(load_library(), prob, batch, device)
        self.device = device

    # torch tensors live on torch's current stream, the engine launches on its own: a call that takes or returns device memory
    # (mem == SMPC_DEVICE) is ordered on the GPU, without blocking the host -- the engine's stream waits for torch's before the
    # call, torch's stream waits for the engine's after it (the outputs may then be consumed on torch's stream right away)
    def _ext_stream(self):
        import torch
        if getattr(self, '_ext', None) is None:
            self._ext = torch.cuda.ExternalStream(self.stream(), device=torch.device('cuda', self.device))
        return self._ext

    def _call(self, name, *args, mem=None):
        if mem == abi.DEVICE:
            import torch
            cur = torch.cuda.current_stream(torch.device('cuda', self.device))
            ext = self._ext_stream()
            ext.wait_stream(cur)
            super()._call(name, *args, mem=mem)
            cur.wait_stream(ext)
        else:
            super()._call(name, *args, mem=mem)

    def sync(self):
        self._call('sync')

    @staticmethod
    def measure_peaks(device: int = 0):
        """Dense FMA rate of the FP64 and FP32 pipes of this GPU, TFLOP/s (smpc_measure_peaks: micro-benchmark, CUDA events)."""
        lib = load_library()
        f64, f32 = C.c_double(), C.c_double()
        rc = lib.smpc_measure_peaks(C.c_int32(device), C.byref(f64), C.byref(f32))
        if rc != 0:
            return None, None
        return f64.value, f32.value

    def launch_count(self) -> int:
        return int(self.lib.smpc_launch_count(self.h))

    def stream(self) -> int:
        return int(self.lib.smpc_stream(self.h) or 0)

    def set_stream(self, stream):
        """Launch on the caller's CUDA stream (a torch.cuda.Stream, a raw cudaStream_t as int, or None = private stream)."""
        raw = 0 if stream is None else int(getattr(stream, 'cuda_stream', stream))
        self._user_stream = stream                            # keep a torch stream object alive
        self._ext = None
        self._check(self.lib.smpc_set_stream(self.h, C.c_void_p(raw)), 'set_stream')

    PROF_NAMES = ['qs_init', 'qs_prep', 'qs_ctl', 'qs_ric1', 'qs_step0', 'qs_ric2', 'qs_step1', 'qs_red', 'qs_compact',
                  'qs_step2_centering', 'qs_final', 'qs_solo']

    def set_profiling(self, on: bool):
        self._call('set_profiling', C.c_int32(int(on)))

    def profile(self):
        """Per-kernel durations of the last QP solve (after set_profiling(True)): {name: (total_ms, launches)}, span, iterations."""
        ms = (C.c_double * len(self.PROF_NAMES))(); n = (C.c_int32 * len(self.PROF_NAMES))()
        span = C.c_double(); it = C.c_int32()
        self._call('get_profile', ms, n, C.byref(span), C.byref(it))
        return {k: (ms[i], n[i]) for i, k in enumerate(self.PROF_NAMES)}, span.value, it.value

    def times(self):
        out = (C.c_double * 7)()
        self._call('get_times', out)
        fields = ['time_lin', 'time_sim', 'time_qp', 'time_qp_solver_call', 'time_glob', 'time_reg', 'time_tot']
        return dict(zip(fields, [v * 1e-3 for v in out]))     # seconds, like acados get_stats


class Sim(SimBase):
    def sync(self):
        self.main.sync()

    def _call(self, name, *args, mem=None):
        if mem == abi.DEVICE:
            import torch
            cur = torch.cuda.current_stream(torch.device('cuda', self.main.device))
            ext = self.main._ext_stream()
            ext.wait_stream(cur)
            super()._call(name, *args, mem=mem)
            cur.wait_stream(ext)
        else:
            super()._call(name, *args, mem=mem)

    def step_times(self):
        """controller.getTime() of the last closed-loop step (mpc.py:239): the acados time fields of the main controller's batched
        step [s] and the number of problems that solved in it.  Synchronises with the step."""
        t = self.main.times()
        solves = self.counters()['rti_solves']
        n = solves - getattr(self, '_solves_seen', 0)
        self._solves_seen = solves
        import numpy as np
        return np.array([t[f] for f in ('time_lin', 'time_sim', 'time_qp', 'time_qp_solver_call', 'time_glob', 'time_reg', 'time_tot')]), int(n)
