"""Host simulation of the device-resident L-BFGS loop: the REAL controller (csrc/lbfgs_ctl.h compiled for the
host, liblbfgs_ctl_host.so) driven by numpy stand-ins for the two streaming kernels of csrc/lbfgs.cu.  The
numpy passes restate exactly what the kernels compute (same outputs, same slot bookkeeping), so the scalar
logic, the ring buffer and the coefficient-space recursion are tested on the CPU."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_LIB = os.path.join(ROOT, "text-based-image-style-transfer_b200", "liblbfgs_ctl_host.so")

CTL_BEGIN, CTL_MID = 0, 1


class HostLbfgs:
    def __init__(self, x0, closure, max_iter=20):
        """closure(x) -> (loss, grad) evaluated at float32 vector x AFTER it clamps x in place if it wants to."""
        lib = C.CDLL(HOST_LIB)
        lib.nst_ctl_new.restype = C.c_void_p
        lib.nst_ctl_free.argtypes = [C.c_void_p]
        lib.nst_ctl_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int]
        lib.nst_ctl_begin_step.argtypes = [C.c_void_p]
        lib.nst_ctl_get_int.argtypes = [C.c_void_p, C.c_int]
        lib.nst_ctl_get_double.argtypes = [C.c_void_p, C.c_int]
        lib.nst_ctl_get_double.restype = C.c_double
        lib.nst_ctl_coef.argtypes = [C.c_void_p]
        lib.nst_ctl_coef.restype = C.POINTER(C.c_float)
        self.lib = lib
        self.slots = lib.nst_ctl_slots()
        self.ctl = lib.nst_ctl_new()
        n = x0.size
        self.x = x0.astype(np.float32).copy()
        self.g = np.zeros(n, np.float32)
        self.g_prev = np.zeros(n, np.float32)
        self.d = np.zeros(n, np.float32)
        self.S = np.zeros((self.slots, n), np.float32)
        self.Y = np.zeros((self.slots, n), np.float32)
        self.ndot = lib.nst_ctl_ndot()
        self.work = np.zeros(lib.nst_ctl_work_doubles(), np.float64)
        self.dots = np.zeros(self.slots * self.ndot, np.float64)
        self.scal = np.zeros(8, np.float64)
        self.td_part = np.zeros(1, np.float32)
        self.closure = closure
        self.max_iter = max_iter
        self.loss = 0.0
        self.closure_calls = 0
        self.trace = []

    def __del__(self):
        try:
            self.lib.nst_ctl_free(self.ctl)
        except Exception:
            pass

    def geti(self, k):
        return self.lib.nst_ctl_get_int(self.ctl, k)

    def getd(self, k):
        return self.lib.nst_ctl_get_double(self.ctl, k)

    @property
    def stop(self):
        return self.geti(4)

    def _eval(self):
        loss, g = self.closure(self.x)
        self.loss = np.float32(loss)
        self.g[:] = g
        if self.stop == 0:
            self.closure_calls += 1
            self.trace.append(float(loss))

    def _pass1(self):
        if self.stop != 0:
            return
        n_iter, hist_len, head = self.geti(0), self.geti(2), self.geti(3)
        t = np.float32(self.getd(0))
        first = n_iter == 0
        g = self.g
        if first:
            s = np.zeros_like(g)
            y = np.zeros_like(g)
        else:
            y = g - self.g_prev
            s = self.d * t
            pn = (head + hist_len) % self.slots
            self.S[pn] = s
            self.Y[pn] = y
        f = np.float64
        self.scal[:] = [s.astype(f) @ s.astype(f), s.astype(f) @ y.astype(f), y.astype(f) @ y.astype(f),
                        s.astype(f) @ g.astype(f), y.astype(f) @ g.astype(f), g.astype(f) @ g.astype(f),
                        np.abs(g).max(), np.abs(g).astype(f).sum()]
        for i in range(hist_len):
            p = (head + i) % self.slots
            Sp, Yp = self.S[p].astype(f), self.Y[p].astype(f)
            self.dots[4 * p:4 * p + 4] = [Sp @ y, Sp @ g, Yp @ y, Yp @ g]

    def _ctl(self, mode):
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        self.lib.nst_ctl_run(self.ctl, p(self.work), p(self.dots), p(self.scal), C.c_float(float(self.loss)),
                             p(self.td_part), 1, mode)

    def _pass2(self):
        if self.geti(5) == 0:
            return
        hist_len, head = self.geti(2), self.geti(3)
        coef = np.ctypeslib.as_array(self.lib.nst_ctl_coef(self.ctl), shape=(2 * self.slots + 1,)).copy()
        t = np.float32(self.getd(5))
        acc = coef[2 * self.slots] * self.g
        for i in range(hist_len):
            p = (head + i) % self.slots
            acc = acc + (coef[p] * self.S[p] + coef[self.slots + p] * self.Y[p])
        acc = acc.astype(np.float32)
        self.g_prev[:] = self.g
        self.d[:] = acc
        self.x[:] = np.clip(self.x + t * acc, 0.0, 1.0).astype(np.float32) if self.clamp else (self.x + t * acc)
        self.td_part[0] = np.abs(acc * t).max()

    clamp = True

    def step(self):
        self.lib.nst_ctl_begin_step(self.ctl)
        self._eval()
        for k in range(1, self.max_iter + 1):
            self._pass1()
            self._ctl(CTL_BEGIN if k == 1 else CTL_MID)
            self._pass2()
            if k != self.max_iter:
                self._eval()
