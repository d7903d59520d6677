"""ctypes wrapper over oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module; the product (toy-example-of-ilqr_b200/) never does.
"""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
LIB_PM_PATH = os.path.join(_HERE, "liboracle_pm.so")
LIB_FMA_PATH = os.path.join(_HERE, "liboracle_fma.so")

PARAM_FIELDS = [
    ("dt", "d"),
    ("w_pos", "d"), ("w_vel", "d"), ("w_yaw", "d"), ("w_acc", "d"), ("w_stl", "d"),
    ("obstacle_exp_q1", "d"), ("obstacle_exp_q2", "d"), ("state_exp_q1", "d"), ("state_exp_q2", "d"),
    ("alm_rho_init", "d"), ("alm_gamma", "d"), ("max_rho", "d"), ("max_mu", "d"),
    ("init_lamb", "d"), ("lamb_decay", "d"), ("lamb_amplify", "d"), ("max_lamb", "d"),
    ("convergence_threshold", "d"), ("accept_step_threshold", "d"),
    ("wheelbase", "d"), ("width", "d"), ("length", "d"),
    ("velo_max", "d"), ("velo_min", "d"), ("yaw_lim", "d"), ("acc_max", "d"), ("acc_min", "d"),
    ("stl_lim", "d"), ("d_safe", "d"),
    ("max_iter", "i"), ("solve_type", "i"), ("reference_point", "i"), ("use_last_solution", "i"),
]


class Params(C.Structure):
    _fields_ = [(n, C.c_double if t == "d" else C.c_int32) for n, t in PARAM_FIELDS]

    @classmethod
    def from_dict(cls, d):
        p = cls()
        for n, t in PARAM_FIELDS:
            setattr(p, n, float(d[n]) if t == "d" else int(d[n]))
        return p


_libs = {}

# dtype strings of this module -> (library flavour, dtype code of the C surface)
#   f64 / f32    the restatement with glibc's transcendentals (pinned to the reference sources, oracle/_ref)
#   f80          the same in x87 long double (64-bit mantissa): the "truth" fp64 implementations are measured against
#   f64pm/f32pm  transcendentals from the portable header shared with the CUDA PARITY build (bit-identical on GPU)
#   f64fma/f32fma  the glibc flavour compiled with FMA contraction (another toolchain's rounding pattern)
_DT = {"f64": ("glibc", 0), "f32": ("glibc", 1), "f80": ("glibc", 2), "f64pm": ("pm", 0), "f32pm": ("pm", 1),
       "f64fma": ("fma", 0), "f32fma": ("fma", 1)}
_PATHS = {"glibc": LIB_PATH, "pm": LIB_PM_PATH, "fma": LIB_FMA_PATH}


def fma_available():
    """The fma flavour executes FMA instructions: only on a host CPU that has them."""
    try:
        return " fma " in open("/proc/cpuinfo").read()
    except OSError:
        return False


def lib(dtype="f64"):
    flavour = _DT[dtype][0]
    if flavour not in _libs:
        path = _PATHS[flavour]
        if flavour == "fma" and not fma_available():
            raise RuntimeError("this CPU has no FMA instructions: the fma flavour of the oracle cannot run")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", _HERE, os.path.basename(path)], check=True)
        l = C.CDLL(path)
        assert l.oracle_sizeof_params() == C.sizeof(Params)
        l.oracle_solver_create.restype = C.c_void_p
        _libs[flavour] = l
    return _libs[flavour]


DT = {k: v[1] for k, v in _DT.items()}


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)




def propagate(params, x, u, dtype="f64"):
    p = Params.from_dict(params)
    x, u, out = _f64(x), _f64(u), np.empty(4)
    lib(dtype).oracle_propagate(C.byref(p), DT[dtype], _dp(x), _dp(u), _dp(out))
    return out


def dyn_derivs(params, N, u, x, dtype="f64"):
    p = Params.from_dict(params)
    u, x = _f64(u), _f64(x)
    A, B = np.empty((N, 4, 4)), np.empty((N, 4, 2))
    lib(dtype).oracle_dyn_derivs(C.byref(p), DT[dtype], N, _dp(u), _dp(x), _dp(A), _dp(B))
    return A, B


def ref_match(wx, wy, x, dtype="f64"):
    wx, wy, x = _f64(wx), _f64(wy), _f64(x)
    idx = np.empty(x.shape[0], np.int32)
    lib(dtype).oracle_ref_match(DT[dtype], len(wx), _dp(wx), _dp(wy), x.shape[0], _dp(x), _ip(idx))
    return idx


def _problem_args(td, ref_velo, n_obs, obs, borders):
    wx, wy, wyaw = _f64(td.wx), _f64(td.wy), _f64(td.wyaw)
    obs = _f64(obs[:n_obs]) if n_obs > 0 else np.zeros((1, 1, 3))
    borders = _f64(borders)
    keep = (wx, wy, wyaw, obs, borders)
    args = (len(wx), _dp(wx), _dp(wy), _dp(wyaw), C.c_double(float(ref_velo)), int(n_obs),
            int(obs.shape[1]), _dp(obs), _dp(borders))
    return args, keep


def total_cost(td, N, ref_velo, n_obs, obs, borders, u, x, dtype="f64", alm_mu=None, alm_rho=0.0):
    p = Params.from_dict(td.params)
    args, keep = _problem_args(td, ref_velo, n_obs, obs, borders)
    u, x, mu = _f64(u), _f64(x), _f64(alm_mu)
    J = C.c_double()
    sc = np.empty(N + 1)
    lib(dtype).oracle_total_cost(C.byref(p), DT[dtype], N, *args, _dp(u), _dp(x), _dp(mu), C.c_double(alm_rho),
                            C.byref(J), _dp(sc))
    return J.value, sc


def cost_derivs(td, N, ref_velo, n_obs, obs, borders, u, x, dtype="f64", alm_mu=None, alm_rho=0.0):
    p = Params.from_dict(td.params)
    args, keep = _problem_args(td, ref_velo, n_obs, obs, borders)
    u, x, mu = _f64(u), _f64(x), _f64(alm_mu)
    lx, lu, lxx, luu = np.empty((N + 1, 4)), np.empty((N, 2)), np.empty((N + 1, 4, 4)), np.empty((N, 2, 2))
    mun = np.empty((N, 8 + 2 * n_obs)) if alm_mu is not None else None
    lib(dtype).oracle_cost_derivs(C.byref(p), DT[dtype], N, *args, _dp(u), _dp(x), _dp(mu), C.c_double(alm_rho),
                             _dp(lx), _dp(lu), _dp(lxx), _dp(luu), _dp(mun))
    return dict(lx=lx, lu=lu, lxx=lxx, luu=luu, mu_next=mun)


def riccati(N, lx, lu, lxx, luu, A, B, lamb, dtype="f64"):
    a = [_f64(v) for v in (lx, lu, lxx, luu, A, B)]
    d, K, dV = np.empty((N, 2)), np.empty((N, 2, 4)), np.empty(2)
    st = C.c_int32()
    lib(dtype).oracle_riccati(DT[dtype], N, *[_dp(v) for v in a], C.c_double(lamb), _dp(d), _dp(K), _dp(dV), C.byref(st))
    return d, K, dV, st.value


def forward(params, N, u, x, d, K, alpha, dtype="f64"):
    p = Params.from_dict(params)
    a = [_f64(v) for v in (u, x, d, K)]
    nu, nx = np.empty((N, 2)), np.empty((N + 1, 4))
    lib(dtype).oracle_forward(C.byref(p), DT[dtype], N, *[_dp(v) for v in a], C.c_double(alpha), _dp(nu), _dp(nx))
    return nu, nx


@dataclass
class OracleResult:
    u: np.ndarray
    x: np.ndarray
    K: np.ndarray
    d: np.ndarray
    J: np.ndarray
    step_cost: np.ndarray
    status: int
    iters: int
    exit_reason: int
    lamb: float
    trace: np.ndarray


class Solver:
    """Stateful oracle solver: one reference CILQRSolver object (warm start, caches, ALM state)."""

    def __init__(self, params, N, dtype="f64"):
        self.N = N
        self.p = Params.from_dict(params)
        self.lib = lib(dtype)
        self.h = C.c_void_p(self.lib.oracle_solver_create(C.byref(self.p), DT[dtype], N))

    def close(self):
        if self.h:
            self.lib.oracle_solver_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def solve(self, td, ref_velo, n_obs, obs, borders, x0, trace_cap=128):
        N = self.N
        args, keep = _problem_args(td, ref_velo, n_obs, obs, borders)
        x0 = _f64(x0)
        u, x, K, d = np.empty((N, 2)), np.empty((N + 1, 4)), np.empty((N, 2, 4)), np.empty((N, 2))
        J, sc, info = np.empty(2), np.empty(N + 1), np.zeros(4, np.int32)
        lamb = C.c_double()
        tr = np.zeros((trace_cap, 6))
        rc = self.lib.oracle_solver_solve(self.h, *args, _dp(x0), _dp(u), _dp(x), _dp(K), _dp(d), _dp(J), _dp(sc),
                                       _ip(info), C.byref(lamb), _dp(tr), trace_cap)
        if rc == -2:
            raise IndexError("Index out of range")
        return OracleResult(u, x, K, d, J, sc, int(info[0]), int(info[1]), int(info[2]), lamb.value, tr[: info[3]])


@dataclass
class BatchResult:
    u: np.ndarray
    x: np.ndarray
    K: np.ndarray
    d: np.ndarray
    J: np.ndarray
    status: np.ndarray
    iters: np.ndarray
    exit_reason: np.ndarray
    tr_status: np.ndarray = None
    tr_alpha: np.ndarray = None
    tr_cost: np.ndarray = None


def solve_batch(pb, dtype="f64", nthreads=None, want_traj=True, trace_cap=0):
    """B independent first solves on `nthreads` host threads (one solve per thread at a time)."""
    if nthreads is None:
        nthreads = os.cpu_count() or 1
    B, N = pb.B, pb.N
    nt = len(pb.templates)
    ParamsArr = Params * nt
    parr = ParamsArr(*[Params.from_dict(td.params) for td in pb.templates])
    off = np.zeros(nt + 1, np.int32)
    for i, td in enumerate(pb.templates):
        off[i + 1] = off[i] + len(td.wx)
    wx = _f64(np.concatenate([td.wx for td in pb.templates]))
    wy = _f64(np.concatenate([td.wy for td in pb.templates]))
    wyaw = _f64(np.concatenate([td.wyaw for td in pb.templates]))
    x0, rv, bd, ob = _f64(pb.x0), _f64(pb.ref_velo), _f64(pb.borders), _f64(pb.obs)
    tm = np.ascontiguousarray(pb.tmpl, dtype=np.int32)
    no = np.ascontiguousarray(pb.n_obs, dtype=np.int32)
    if want_traj:
        u, x = np.empty((B, N, 2)), np.empty((B, N + 1, 4))
        K, d = np.empty((B, N, 2, 4)), np.empty((B, N, 2))
    else:
        u = x = K = d = None
    J = np.empty((B, 2))
    st, it, ex = np.empty(B, np.int32), np.empty(B, np.int32), np.empty(B, np.int32)
    trs = tra = trc = None
    if trace_cap > 0:
        trs, tra = np.zeros((B, trace_cap), np.int32), np.zeros((B, trace_cap), np.int32)
        trc = np.zeros((B, trace_cap))
    rc = lib(dtype).oracle_solve_batch(DT[dtype], N, nt, parr, _ip(off), _dp(wx), _dp(wy), _dp(wyaw), B, _ip(tm),
                                  _dp(x0), _dp(rv), _dp(bd), _ip(no), int(pb.max_obs), int(pb.obs_len), _dp(ob),
                                  _dp(u), _dp(x), _dp(K), _dp(d), _dp(J), _ip(st), _ip(it), _ip(ex),
                                  int(trace_cap), _ip(trs), _ip(tra), _dp(trc), int(nthreads))
    if rc != 0:
        raise RuntimeError("oracle_solve_batch failed: %d" % rc)
    return BatchResult(u, x, K, d, J, st, it, ex, trs, tra, trc)


def pmath_eval(fn, a, b=None):
    """The portable transcendentals (csrc/cilqr_pmath.h, as compiled into liboracle_pm.so) over an array."""
    code = {"sin": 0, "cos": 1, "tan": 2, "atan": 3, "exp": 4, "hypot": 5}[fn]
    a = _f64(a)
    b = _f64(b) if b is not None else a
    out = np.empty_like(a)
    rc = lib("f64pm").oracle_pmath_eval(code, a.size, _dp(a), _dp(b), _dp(out))
    assert rc == 0
    return out
