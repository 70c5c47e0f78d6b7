"""ctypes plumbing shared by every library that implements the C ABI of include/safe_mpc_b200.h.

``EngineBase`` turns the flat ``<prefix>_*`` entry points into an object; the product subclass lives in
``engine.py`` (CUDA library, ``smpc_`` prefix, arrays may be numpy = host or torch CUDA tensors = device).
The class is prefix-agnostic on purpose: the test oracle exports the same functions under ``orc_`` and its
wrapper (oracle/oracle.py, test infrastructure) reuses this plumbing -- the product never loads it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi

_P = C.c_void_p


def _is_torch(a):
    return type(a).__module__.startswith('torch')


class EngineBase:
    prefix = 'smpc_'
    has_mem = True      # entry points take the trailing `mem` argument

    def __init__(self, lib, prob: abi.Problem, batch: int, arg: int):
        self.lib = lib
        self.prob = prob
        self.B = int(batch)
        self.N = int(prob.N)
        self._keep = []
        h = _P()
        rc = self._fn('create')(C.byref(prob), C.c_int32(self.B), C.c_int32(arg), C.byref(h))
        if rc != 0:
            raise RuntimeError(f'{self.prefix}create failed ({rc}): {self.last_error(None)}')
        self.h = h

    # -- helpers ---------------------------------------------------------------------------------
    def _fn(self, name):
        f = getattr(self.lib, self.prefix + name)
        return f

    def last_error(self, h='self'):
        if not hasattr(self.lib, self.prefix + 'last_error'):
            return ''
        f = self._fn('last_error')
        f.restype = C.c_char_p
        f.argtypes = [_P]
        s = f(self.h if h == 'self' else None)
        return s.decode() if s else ''

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f'{self.prefix}{what} failed ({rc}): {self.last_error()}')

    def _in(self, a, dtype, shape=None):
        """-> (pointer, mem, keepalive).  numpy / lists are host memory, torch CUDA tensors device memory."""
        if a is None:
            return _P(0), abi.HOST, None
        if _is_torch(a):
            import torch
            want = {np.float64: torch.float64, np.int32: torch.int32, np.uint8: torch.uint8}[dtype]
            if a.dtype == torch.bool and dtype == np.uint8:
                a = a.to(torch.uint8)
            if a.dtype != want or not a.is_contiguous():
                a = a.to(want).contiguous()
            if shape is not None and tuple(a.shape) != tuple(shape):
                raise ValueError(f'expected shape {shape}, got {tuple(a.shape)}')
            if a.is_cuda:
                if not self.has_mem:
                    raise TypeError('this library takes host arrays only')
                return _P(a.data_ptr()), abi.DEVICE, a
            a = a.numpy()
        arr = np.ascontiguousarray(a, dtype=dtype)
        if shape is not None and tuple(arr.shape) != tuple(shape):
            raise ValueError(f'expected shape {shape}, got {arr.shape}')
        return _P(arr.ctypes.data), abi.HOST, arr

    def _out(self, shape, dtype, like=None):
        if like is not None and _is_torch(like) and like.is_cuda:
            import torch
            want = {np.float64: torch.float64, np.int32: torch.int32, np.uint8: torch.uint8}[dtype]
            t = torch.empty(shape, dtype=want, device=like.device)
            return t, _P(t.data_ptr()), abi.DEVICE
        arr = np.empty(shape, dtype=dtype)
        return arr, _P(arr.ctypes.data), abi.HOST

    def _call(self, name, *args, mem=None):
        f = self._fn(name)
        a = [self.h] + list(args)
        if self.has_mem and mem is not None:
            a.append(C.c_int32(mem))
        self._check(f(*a), name)

    @staticmethod
    def _mem(*mems):
        ms = {m for m in mems if m is not None}
        if len(ms) > 1:
            raise ValueError('host and device arrays cannot be mixed in one call')
        return ms.pop() if ms else abi.HOST

    # -- plant / guess ---------------------------------------------------------------------------
    def set_plant_inertial(self, inertial):
        p, m, k = self._in(inertial, np.float64, (self.B, abi.NQ, 10))
        self._call('set_plant_inertial', p, mem=m)

    def set_torque_noise(self, noise):
        p, m, k = self._in(noise, np.float64, (self.B, abi.NU))
        self._call('set_torque_noise', p, mem=m)

    def set_ee_trajectory(self, traj):
        """cost.traj of the reference (controller.py:153-156): traj [n, 3], one end-effector reference per control step; None or an
        empty array restores the constant ee_ref of the problem struct"""
        if traj is None or len(traj) == 0:
            args = [_P(0), C.c_int32(0)]
            m = abi.HOST
        else:
            n = int(len(traj))
            p_, m, k = self._in(traj, np.float64, (n, 3))
            args = [p_, C.c_int32(n)]
        self._call('set_ee_trajectory', *args, mem=m)

    def set_guess(self, xg, ug):
        px, mx, kx = self._in(xg, np.float64, (self.B, self.N + 1, abi.NX))
        pu, mu, ku = self._in(ug, np.float64, (self.B, self.N, abi.NU))
        self._call('set_guess', px, pu, mem=self._mem(mx, mu))

    def get_guess(self, like=None):
        xg, px, m = self._out((self.B, self.N + 1, abi.NX), np.float64, like)
        ug, pu, _ = self._out((self.B, self.N, abi.NU), np.float64, like)
        self._call('get_guess', px, pu, mem=m)
        return xg, ug

    def get_temp(self, like=None):
        xt, px, m = self._out((self.B, self.N + 1, abi.NX), np.float64, like)
        ut, pu, _ = self._out((self.B, self.N, abi.NU), np.float64, like)
        self._call('get_temp', px, pu, mem=m)
        return xt, ut

    def reset_controller(self):
        self._call('reset_controller')

    # -- solves ----------------------------------------------------------------------------------
    def rti_solve(self, x0, active=None):
        px, mx, kx = self._in(x0, np.float64, (self.B, abi.NX))
        pa, ma, ka = self._in(active, np.uint8, (self.B,)) if active is not None else (_P(0), None, None)
        st, ps, _ = self._out((self.B,), np.int32, x0)
        self._call('rti_solve', px, pa, ps, mem=self._mem(mx, ma))
        return st

    def controller_step(self, x, active=None):
        px, mx, kx = self._in(x, np.float64, (self.B, abi.NX))
        pa, ma, ka = self._in(active, np.uint8, (self.B,)) if active is not None else (_P(0), None, None)
        u, pu, _ = self._out((self.B, abi.NU), np.float64, x)
        ab, pab, _ = self._out((self.B,), np.uint8, x)
        if _is_torch(u):
            u.zero_(); ab.zero_()
        else:
            u[:] = 0.0; ab[:] = 0
        self._call('controller_step', px, pa, pu, pab, mem=self._mem(mx, ma))
        return u, ab

    def plant_step(self, x, u):
        px, mx, kx = self._in(x, np.float64, (self.B, abi.NX))
        pu, mu, ku = self._in(u, np.float64, (self.B, abi.NU))
        xn, pxn, _ = self._out((self.B, abi.NX), np.float64, x)
        a, pa, _ = self._out((self.B, abi.NU), np.float64, x)
        self._call('plant_step', px, pu, pxn, pa, mem=self._mem(mx, mu))
        return xn, a

    # -- model pieces ----------------------------------------------------------------------------
    def tau(self, x, u):
        n = len(x)
        px, mx, kx = self._in(x, np.float64, (n, abi.NX))
        pu, mu, ku = self._in(u, np.float64, (n, abi.NU))
        out, po, _ = self._out((n, abi.NU), np.float64, x)
        self._call('tau', C.c_int32(n), px, pu, po, mem=self._mem(mx, mu))
        return out

    def rk4_sens(self, x, tau, dt, sens=True):
        """torque-input dynamics x' = [v; M(q)^-1 (tau - h(q, v))], one explicit RK4 step -> (x_next [n, 10], A [n, 10, 10], B [n, 10, 5])
        with A = d x_next / d x, B = d x_next / d tau; sens=False: x_next only"""
        n = len(x)
        px, mx, kx = self._in(x, np.float64, (n, abi.NX))
        pt, mt, kt = self._in(tau, np.float64, (n, abi.NU))
        xn, pn, _ = self._out((n, abi.NX), np.float64, x)
        if sens:
            A, pa, _ = self._out((n, abi.NX, abi.NX), np.float64, x)
            Bm, pb, _ = self._out((n, abi.NX, abi.NU), np.float64, x)
        else:
            A = Bm = None; pa = pb = _P(0)
        self._call('rk4_sens', C.c_int32(n), px, pt, C.c_double(dt), pn, pa, pb, mem=self._mem(mx, mt))
        return (xn, A, Bm) if sens else xn

    def kinematics(self, x):
        n = len(x)
        px, mx, kx = self._in(x, np.float64, (n, abi.NX))
        ee, pe, _ = self._out((n, 3), np.float64, x)
        dist, pd, _ = self._out((n, abi.NPAIR), np.float64, x)
        self._call('kinematics', C.c_int32(n), px, pe, pd, mem=mx)
        return ee, dist

    def nn_constraint(self, x, grad=True):
        n = len(x)
        px, mx, kx = self._in(x, np.float64, (n, abi.NX))
        c, pc, _ = self._out((n,), np.float64, x)
        if grad:
            g, pg, _ = self._out((n, abi.NX), np.float64, x)
        else:
            g, pg = None, _P(0)
        self._call('nn_constraint', C.c_int32(n), px, pc, pg, mem=mx)
        return (c, g) if grad else c

    def get_lin(self, like=None):
        lin, p, m = self._out((self.B, self.N + 1, abi.REC), np.float64, like)
        self._call('get_lin', p, mem=m)
        return lin

    def get_qp(self, like=None):
        dz, p1, m = self._out((self.B, self.N + 1, abi.NX + abi.NU), np.float64, like)
        pi, p2, _ = self._out((self.B, self.N, abi.NX), np.float64, like)
        lam, p3, _ = self._out((self.B, self.N + 1, abi.QP_NC), np.float64, like)
        t, p4, _ = self._out((self.B, self.N + 1, abi.QP_NC), np.float64, like)
        self._call('get_qp', p1, p2, p3, p4, mem=m)
        return dz, pi, lam, t

    # -- controller state ------------------------------------------------------------------------
    def get_state(self, field, like=None):
        out, p, m = self._out((self.B,), np.int32, like)
        self._call('get_state_i32', C.c_int32(field), p, mem=m)
        return out

    def set_state(self, field, values):
        p, m, k = self._in(values, np.int32, (self.B,))
        self._call('set_state_i32', C.c_int32(field), p, mem=m)

    def get_qp_residuals(self, like=None):
        """max-norm residuals of the last QP of every problem: [B, 5] = stationarity, dynamics, inequality, complementarity, mu"""
        out, p, m = self._out((self.B, 5), np.float64, like)
        self._call('get_qp_residuals', p, mem=m)
        return out

    def get_x_viable(self, like=None):
        out, p, m = self._out((self.B, abi.NX), np.float64, like)
        self._call('get_x_viable', p, mem=m)
        return out

    def close(self):
        if getattr(self, 'h', None):
            f = self._fn('destroy')
            f.restype = None
            f(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SimBase:
    """Closed loop of scripts/mpc.py:102-291 over the whole batch (``<prefix>sim_*``)."""

    def __init__(self, main: EngineBase, backup: EngineBase, n_steps: int):
        self.main, self.backup = main, backup
        self.lib, self.prefix, self.has_mem = main.lib, main.prefix, main.has_mem
        self.n_steps = int(n_steps)
        self.B = main.B
        s = _P()
        rc = getattr(self.lib, self.prefix + 'sim_create')(main.h, backup.h, C.c_int32(self.n_steps), C.byref(s))
        if rc != 0:
            raise RuntimeError(f'{self.prefix}sim_create failed ({rc}): {main.last_error()}')
        self.s = s

    def _call(self, name, *args, mem=None):
        a = [self.s] + list(args)
        if self.has_mem and mem is not None:
            a.append(C.c_int32(mem))
        rc = getattr(self.lib, self.prefix + 'sim_' + name)(*a)
        if rc != 0:
            raise RuntimeError(f'{self.prefix}sim_{name} failed ({rc}): {self.main.last_error()}')

    def reset(self, x_init):
        p, m, k = self.main._in(x_init, np.float64, (self.B, abi.NX))
        self._call('reset', p, mem=m)

    def step(self):
        self._call('step')

    def run(self, n=None):
        self._call('run', C.c_int32(self.n_steps if n is None else int(n)))

    def outcome(self, like=None):
        out, p, m = self.main._out((self.B,), np.int32, like)
        self._call('get_outcome', p, mem=m)
        return out

    def log(self, like=None):
        x, px, m = self.main._out((self.B, self.n_steps + 1, abi.NX), np.float64, like)
        u, pu, _ = self.main._out((self.B, self.n_steps, abi.NU), np.float64, like)
        self._call('get_log', px, pu, mem=m)
        return x, u

    def x_viable(self, like=None):
        xv, p, m = self.main._out((self.B, abi.NX), np.float64, like)
        self._call('get_x_viable', p, mem=m)
        return xv

    def counters(self):
        out = (C.c_int64 * 4)()
        self._call('get_counters', out)
        return {'rti_solves': out[0], 'backup_solves': out[1], 'plant_steps': out[2], 'ipm_iterations': out[3]}

    def close(self):
        if getattr(self, 's', None):
            f = getattr(self.lib, self.prefix + 'sim_destroy')
            f.restype = None
            f(self.s)
            self.s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
