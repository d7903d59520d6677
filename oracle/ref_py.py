"""ctypes wrapper over oracle/_ref/libcilqr_ref.so — the reference's own sources compiled in
place against oracle/shim (TEST INFRASTRUCTURE ONLY; exists only where /root/reference was
available at build time or a prebuilt _ref/ travelled with the snapshot)."""
import ctypes as C
import os

import numpy as np

from .oracle_py import Params, _dp, _f64, _ip

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libcilqr_ref.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        assert _lib.ref_sizeof_params() == C.sizeof(Params)
        _lib.ref_solver_create.restype = C.c_void_p
    return _lib


def _args(td, ref_velo, n_obs, obs, borders):
    wx, wy, wyaw = _f64(td.wx), _f64(td.wy), _f64(td.wyaw)
    obs = _f64(obs[:n_obs]) if n_obs > 0 else np.zeros((1, 1, 3))
    borders = _f64(borders)
    keep = (wx, wy, wyaw, obs, borders)
    return (len(wx), _dp(wx), _dp(wy), _dp(wyaw), C.c_double(float(ref_velo)), int(n_obs), int(obs.shape[1]),
            _dp(obs), _dp(borders)), keep


class RefSolver:
    """One reference CILQRSolver object (stateful, like the original)."""

    def __init__(self, params, N):
        self.N = N
        self.p = Params.from_dict(params)
        self.h = C.c_void_p(lib().ref_solver_create(C.byref(self.p), N))

    def close(self):
        if self.h:
            lib().ref_solver_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def solve(self, td, ref_velo, n_obs, obs, borders, x0):
        a, keep = _args(td, ref_velo, n_obs, obs, borders)
        x0 = _f64(x0)
        u, x, info = np.empty((self.N, 2)), np.empty((self.N + 1, 4)), np.zeros(1, np.int32)
        rc = lib().ref_solver_solve(self.h, *a, _dp(x0), _dp(u), _dp(x), _ip(info))
        if rc == -2:
            raise IndexError("Index out of range")
        return u, x, int(info[0])

    def total_cost(self, td, ref_velo, n_obs, obs, borders, u, x):
        a, keep = _args(td, ref_velo, n_obs, obs, borders)
        u, x = _f64(u), _f64(x)
        J = C.c_double()
        lib().ref_total_cost(self.h, *a, _dp(u), _dp(x), C.byref(J))
        return J.value

    def ref_points(self, td, x):
        wx, wy, wyaw, x = _f64(td.wx), _f64(td.wy), _f64(td.wyaw), _f64(x)
        pts = np.empty((x.shape[0], 3))
        lib().ref_ref_points(self.h, len(wx), _dp(wx), _dp(wy), _dp(wyaw), x.shape[0], _dp(x), _dp(pts))
        return pts

    def backward_pass(self, td, ref_velo, n_obs, obs, borders, u, x, lamb):
        N = self.N
        a, keep = _args(td, ref_velo, n_obs, obs, borders)
        u, x = _f64(u), _f64(x)
        r = dict(lx=np.empty((N + 1, 4)), lu=np.empty((N, 2)), lxx=np.empty((N + 1, 4, 4)), luu=np.empty((N, 2, 2)),
                 A=np.empty((N, 4, 4)), B=np.empty((N, 4, 2)), d=np.empty((N, 2)), K=np.empty((N, 2, 4)),
                 dV=np.empty(2))
        st = C.c_int32()
        lib().ref_backward_pass(self.h, *a, _dp(u), _dp(x), C.c_double(lamb), _dp(r["lx"]), _dp(r["lu"]),
                                _dp(r["lxx"]), _dp(r["luu"]), _dp(r["A"]), _dp(r["B"]), _dp(r["d"]), _dp(r["K"]),
                                _dp(r["dV"]), C.byref(st))
        r["status"] = st.value
        return r

    def forward_pass(self, u, x, d, K, alpha):
        N = self.N
        a = [_f64(v) for v in (u, x, d, K)]
        nu, nx = np.empty((N, 2)), np.empty((N + 1, 4))
        lib().ref_forward_pass(self.h, *[_dp(v) for v in a], C.c_double(alpha), _dp(nu), _dp(nx))
        return nu, nx


def reference_line(x, y, width=0.0, cap=70000):
    x, y = _f64(x), _f64(y)
    wx, wy, wyaw, lon = np.empty(cap), np.empty(cap), np.empty(cap), np.empty(cap)
    M = lib().ref_reference_line(len(x), _dp(x), _dp(y), C.c_double(width), cap, _dp(wx), _dp(wy), _dp(wyaw), _dp(lon))
    return wx[:M].copy(), wy[:M].copy(), wyaw[:M].copy(), lon[:M].copy()
