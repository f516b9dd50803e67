"""ctypes loaders for the checker libraries (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under asif_b200/ does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libasif_ref.so")
# the same reference sources with their filter classes constructible on ASIF::QPWrapperB200 (oracle/ref_build/qp_select_shim.h)
REF_B200_SO = os.path.join(HERE, "_ref", "libasif_ref_b200.so")
ORACLE_SO = os.path.join(HERE, "liboracle.so")

CFG_DI_EXPLICIT, CFG_DI_IMPLICIT_TB, CFG_IP_IMPLICIT, CFG_IP_ROBUST, CFG_IP_REALIZABLE, CFG_SEGWAY_TB = 1, 2, 3, 4, 5, 6
CFG_IP_IMPLICIT_RB, CFG_DI_IMPLICIT_RB = 7, 8  # ASIFimplicitRB on the pendulum (split gradients) / double-integrator (fused) callbacks

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


class RefLib:
    """oracle/_ref/libasif_ref.so: unmodified reference sources + OSQP stand-in."""

    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (build with `make -C oracle/ref_build`; needs /root/reference)")
        L = self.lib = C.CDLL(path)
        L.ref_set_qp_mode.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int]
        L.ref_qp_stats.argtypes = [C.POINTER(C.c_longlong)] * 3
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, _dp, C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_dims.argtypes = [C.c_void_p, _ip]
        L.ref_filter_batch.argtypes = [C.c_void_p, C.c_int64, _dp, _dp, _dp, _dp, _ip, _dp]
        L.ref_filter_batch_ex.argtypes = [C.c_void_p, C.c_int64, _dp, _dp, _dp, _dp, _ip, _dp, _ip, _ip]
        L.ref_rollout.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_double, _dp, _dp, _dp, _ip, C.POINTER(C.c_int64)]

    def set_qp_mode(self, eps=1e-8, polish=1, warm_start=0, max_iter=20000, polish_refine_iter=10):
        """Oracle setting by default; set_qp_mode(-1, -1, -1, -1, -1) restores the reference defaults."""
        self.lib.ref_set_qp_mode(eps, polish, warm_start, max_iter)
        self.lib.ref_set_polish_refine(polish_refine_iter)

    def select_backend(self, backend):
        """libasif_ref_b200.so only: 0 = OSQP stand-in, 1 = ASIF::QPWrapperB200, for filters created afterwards"""
        self.lib.ref_select_backend.argtypes = [C.c_int]
        self.lib.ref_select_backend(backend)

    def qp_stats(self):
        a, b, c = C.c_longlong(), C.c_longlong(), C.c_longlong()
        self.lib.ref_qp_stats(C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def create(self, cfg, opts=None):
        o = None if opts is None else np.ascontiguousarray(opts, dtype=np.float64)
        h = self.lib.ref_create(cfg, _d(o), 0 if o is None else o.size)
        if not h:
            raise RuntimeError("ref_create(%d) failed" % cfg)
        return RefFilter(self, h)


class RefFilter:
    def __init__(self, lib, h):
        self.lib, self.h = lib, h
        dims = np.zeros(6, dtype=np.int32)
        lib.lib.ref_dims(h, dims.ctypes.data_as(_ip))
        self.nx, self.nu, self.n_relax, self.nc, self.nv, self.n_diag = (int(v) for v in dims)

    def update_options(self, opts):
        """the class's updateOptions(options); returns its code (1, or 2 / 3 when satSharpness was clamped)"""
        o = np.ascontiguousarray(opts, dtype=np.float64)
        self.lib.lib.ref_update_options.argtypes = [C.c_void_p, _dp, C.c_int]
        self.lib.lib.ref_update_options.restype = C.c_int32
        return int(self.lib.lib.ref_update_options(self.h, _d(o), o.size))

    def set_learning(self, dims, blob):
        """Options.use_learning = true with these networks (implicit classes only)."""
        d = np.ascontiguousarray(dims, dtype=np.uint32)
        b = np.ascontiguousarray(blob, dtype=np.float64)
        self.lib.lib.ref_set_learning.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), _dp]
        assert self.lib.lib.ref_set_learning(self.h, d.ctypes.data_as(C.POINTER(C.c_uint32)), _d(b)) == 0

    def filter_batch(self, x, u_des, want_diag=False):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, self.nx)
        u_des = np.ascontiguousarray(u_des, dtype=np.float64).reshape(-1, self.nu)
        n = x.shape[0]
        u = np.zeros((n, self.nu))
        relax = np.zeros((n, self.n_relax))
        rc = np.zeros(n, dtype=np.int32)
        diag = np.zeros((n, self.n_diag)) if want_diag else None
        r = self.lib.lib.ref_filter_batch(self.h, n, _d(x), _d(u_des), _d(u), _d(relax), rc.ctypes.data_as(_ip), _d(diag))
        assert r == 0
        return (u, relax, rc, diag) if want_diag else (u, relax, rc)

    def filter_batch_cost(self, x, c, H=None):
        """The reference's filter(x, H, c, uAct, relax) per state -> (u, relax, rc, diag, qp_status)."""
        L = self.lib.lib
        L.ref_filter_batch_cost.argtypes = [C.c_void_p, C.c_int64, _dp, _dp, _dp, _dp, _dp, _ip, _dp, _ip]
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, self.nx)
        c = np.ascontiguousarray(c, dtype=np.float64).reshape(-1, self.nv)
        Hf = None if H is None else np.asfortranarray(np.asarray(H, dtype=np.float64).reshape(self.nu, self.nu))
        n = x.shape[0]
        u = np.zeros((n, self.nu))
        relax = np.zeros((n, self.n_relax))
        rc = np.zeros(n, dtype=np.int32)
        diag = np.zeros((n, self.n_diag))
        st = np.zeros(n, dtype=np.int32)
        r = L.ref_filter_batch_cost(self.h, n, _d(x), None if Hf is None else Hf.ctypes.data_as(_dp), _d(c), _d(u), _d(relax),
                                    rc.ctypes.data_as(_ip), _d(diag), st.ctypes.data_as(_ip))
        assert r == 0
        return u, relax, rc, diag, st

    def filter_batch_lie(self, x, u_des, Lfh, Lgh):
        """ASIF::filter(x, uDes, uAct, Lfh, Lgh, relax) per state (cfg 1) -> (u, relax, rc, diag)."""
        L = self.lib.lib
        L.ref_filter_batch_lie.argtypes = [C.c_void_p, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp]
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, self.nx)
        u_des = np.ascontiguousarray(u_des, dtype=np.float64).reshape(-1, self.nu)
        Lfh = np.ascontiguousarray(Lfh, dtype=np.float64)
        Lgh = np.ascontiguousarray(Lgh, dtype=np.float64)
        n = x.shape[0]
        u, relax, rc, diag = np.zeros((n, self.nu)), np.zeros((n, 1)), np.zeros(n, dtype=np.int32), np.zeros((n, self.n_diag))
        assert L.ref_filter_batch_lie(self.h, n, _d(x), _d(u_des), _d(Lfh), _d(Lgh), _d(u), _d(relax), rc.ctypes.data_as(_ip),
                                      _d(diag)) == 0
        return u, relax, rc, diag

    def filter_batch_ex(self, x, u_des):
        """(u, relax, rc, diag, raw OSQP status, ADMM iterations) - single-threaded use only."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, self.nx)
        u_des = np.ascontiguousarray(u_des, dtype=np.float64).reshape(-1, self.nu)
        n = x.shape[0]
        u = np.zeros((n, self.nu))
        relax = np.zeros((n, self.n_relax))
        rc = np.zeros(n, dtype=np.int32)
        diag = np.zeros((n, self.n_diag))
        st = np.zeros(n, dtype=np.int32)
        it = np.zeros(n, dtype=np.int32)
        r = self.lib.lib.ref_filter_batch_ex(self.h, n, _d(x), _d(u_des), _d(u), _d(relax), rc.ctypes.data_as(_ip), _d(diag),
                                             st.ctypes.data_as(_ip), it.ctypes.data_as(_ip))
        assert r == 0
        return u, relax, rc, diag, st, it

    def rollout(self, x0, u_des, steps, dt):
        x = np.array(x0, dtype=np.float64).reshape(-1, self.nx).copy()
        u_des = np.ascontiguousarray(u_des, dtype=np.float64).reshape(-1, self.nu)
        n = x.shape[0]
        u = np.zeros((n, self.nu))
        rc = np.zeros(n, dtype=np.int32)
        hist = np.zeros(8, dtype=np.int64)
        r = self.lib.lib.ref_rollout(self.h, n, steps, dt, _d(x), _d(u_des), _d(u), rc.ctypes.data_as(_ip),
                                     hist.ctypes.data_as(C.POINTER(C.c_int64)))
        assert r == 0
        return x, u, rc, hist

    def rollout_ex(self, x0, u_des, steps, dt):
        """rollout + per agent the number of control steps whose QP ended in an inexact ADMM exit."""
        x = np.array(x0, dtype=np.float64).reshape(-1, self.nx).copy()
        u_des = np.ascontiguousarray(u_des, dtype=np.float64).reshape(-1, self.nu)
        n = x.shape[0]
        u, rc, hist, bad = np.zeros((n, self.nu)), np.zeros(n, dtype=np.int32), np.zeros(8, dtype=np.int64), np.zeros(n, dtype=np.int32)
        L = self.lib.lib
        L.ref_rollout_ex.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_double, _dp, _dp, _dp, _ip, C.POINTER(C.c_int64), _ip]
        assert L.ref_rollout_ex(self.h, n, steps, dt, _d(x), _d(u_des), _d(u), rc.ctypes.data_as(_ip),
                                hist.ctypes.data_as(C.POINTER(C.c_int64)), bad.ctypes.data_as(_ip)) == 0
        return x, u, rc, hist, bad

    def rollout_log(self, x0, u_des, steps, dt):
        """rollout_ex + the state, output and return code of every filter() call: (x, u, rc, hist, bad, x_log, u_log, rc_log, st_log)."""
        x = np.array(x0, dtype=np.float64).reshape(-1, self.nx).copy()
        u_des = np.ascontiguousarray(u_des, dtype=np.float64).reshape(-1, self.nu)
        n = x.shape[0]
        u, rc, hist, bad = np.zeros((n, self.nu)), np.zeros(n, dtype=np.int32), np.zeros(8, dtype=np.int64), np.zeros(n, dtype=np.int32)
        xl, ul = np.zeros((n, steps, self.nx)), np.zeros((n, steps, self.nu))
        rl, sl = np.zeros((n, steps), dtype=np.int32), np.zeros((n, steps), dtype=np.int32)
        L = self.lib.lib
        L.ref_rollout_log.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_double, _dp, _dp, _dp, _ip, C.POINTER(C.c_int64), _ip,
                                      _dp, _dp, _ip, _ip]
        assert L.ref_rollout_log(self.h, n, steps, dt, _d(x), _d(u_des), _d(u), rc.ctypes.data_as(_ip),
                                 hist.ctypes.data_as(C.POINTER(C.c_int64)), bad.ctypes.data_as(_ip), _d(xl), _d(ul),
                                 rl.ctypes.data_as(_ip), sl.ctypes.data_as(_ip)) == 0
        return x, u, rc, hist, bad, xl, ul, rl, sl

    def close(self):
        if self.h:
            self.lib.lib.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OracleLib:
    """oracle/liboracle.so: the C restatement (asif_oracle.c, oracle_models.c, qp_enum.c)."""

    def __init__(self, path=ORACLE_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (build with `make -C oracle`)")
        L = self.lib = C.CDLL(path)
        L.oracle_dims.argtypes = [C.c_int, _dp, C.c_int, _ip]
        L.oracle_filter_batch.argtypes = [C.c_int, _dp, C.c_int, C.c_int64, _dp, _dp, _dp, _dp, _ip, _dp]
        L.oracle_rollout.argtypes = [C.c_int, _dp, C.c_int, C.c_int64, C.c_int32, C.c_double, _dp, _dp, _dp, _ip,
                                     C.POINTER(C.c_int64)]
        L.oracle_qp_solve.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp,
                                      C.POINTER(C.c_ubyte), _dp]

    def set_learning(self, dims=None, blob=None):
        """Process-wide learned-residual switch of the implicit classes; set_learning() turns it off."""
        self.lib.oracle_set_learning.argtypes = [C.POINTER(C.c_uint32), _dp]
        if dims is None:
            assert self.lib.oracle_set_learning(None, None) == 0
            return
        d = np.ascontiguousarray(dims, dtype=np.uint32)
        b = np.ascontiguousarray(blob, dtype=np.float64)
        assert self.lib.oracle_set_learning(d.ctypes.data_as(C.POINTER(C.c_uint32)), _d(b)) == 0

    def dims(self, cfg, opts=None):
        o = None if opts is None else np.ascontiguousarray(opts, dtype=np.float64)
        dims = np.zeros(6, dtype=np.int32)
        r = self.lib.oracle_dims(cfg, _d(o), 0 if o is None else o.size, dims.ctypes.data_as(_ip))
        if r != 0:
            raise RuntimeError("oracle_dims(%d) failed" % cfg)
        return tuple(int(v) for v in dims)

    def filter_batch(self, cfg, x, u_des, opts=None, want_diag=False):
        o = None if opts is None else np.ascontiguousarray(opts, dtype=np.float64)
        nx, nu, n_relax, nc, nv, n_diag = self.dims(cfg, opts)
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, nx)
        u_des = np.ascontiguousarray(u_des, dtype=np.float64).reshape(-1, nu)
        n = x.shape[0]
        u = np.zeros((n, nu))
        relax = np.zeros((n, n_relax))
        rc = np.zeros(n, dtype=np.int32)
        diag = np.zeros((n, n_diag)) if want_diag else None
        r = self.lib.oracle_filter_batch(cfg, _d(o), 0 if o is None else o.size, n, _d(x), _d(u_des), _d(u), _d(relax),
                                         rc.ctypes.data_as(_ip), _d(diag))
        assert r == 0
        return (u, relax, rc, diag) if want_diag else (u, relax, rc)

    def filter_batch_cost(self, cfg, x, c, H=None, opts=None, want_diag=False):
        """filter(x, H, c, uAct, relax) per state: c (n, nv), H (nu, nu) or None."""
        self.lib.oracle_filter_batch_cost.argtypes = [C.c_int, _dp, C.c_int, C.c_int64, _dp, _dp, _dp, _dp, _dp, _ip, _dp]
        o = None if opts is None else np.ascontiguousarray(opts, dtype=np.float64)
        nx, nu, n_relax, nc, nv, n_diag = self.dims(cfg, opts)
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, nx)
        c = np.ascontiguousarray(c, dtype=np.float64).reshape(-1, nv)
        Hf = None if H is None else np.asfortranarray(np.asarray(H, dtype=np.float64).reshape(nu, nu))
        n = x.shape[0]
        u = np.zeros((n, nu))
        relax = np.zeros((n, n_relax))
        rc = np.zeros(n, dtype=np.int32)
        diag = np.zeros((n, n_diag)) if want_diag else None
        r = self.lib.oracle_filter_batch_cost(cfg, _d(o), 0 if o is None else o.size, n, _d(x),
                                              None if Hf is None else Hf.ctypes.data_as(_dp), _d(c), _d(u), _d(relax),
                                              rc.ctypes.data_as(_ip), _d(diag))
        assert r == 0
        return (u, relax, rc, diag) if want_diag else (u, relax, rc)

    def filter_batch_lie(self, x, u_des, Lfh, Lgh, opts=None):
        """ASIF::filter(x, uDes, uAct, Lfh, Lgh, relax) per state (cfg 1)."""
        self.lib.oracle_filter_batch_lie.argtypes = [_dp, C.c_int, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp]
        o = None if opts is None else np.ascontiguousarray(opts, dtype=np.float64)
        nx, nu, n_relax, nc, nv, n_diag = self.dims(CFG_DI_EXPLICIT, opts)
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, nx)
        u_des = np.ascontiguousarray(u_des, dtype=np.float64).reshape(-1, nu)
        Lfh = np.ascontiguousarray(Lfh, dtype=np.float64)
        Lgh = np.ascontiguousarray(Lgh, dtype=np.float64)
        n = x.shape[0]
        u, relax, rc, diag = np.zeros((n, nu)), np.zeros((n, 1)), np.zeros(n, dtype=np.int32), np.zeros((n, n_diag))
        assert self.lib.oracle_filter_batch_lie(_d(o), 0 if o is None else o.size, n, _d(x), _d(u_des), _d(Lfh), _d(Lgh), _d(u),
                                                _d(relax), rc.ctypes.data_as(_ip), _d(diag)) == 0
        return u, relax, rc, diag

    def rollout(self, cfg, x0, u_des, steps, dt, opts=None):
        o = None if opts is None else np.ascontiguousarray(opts, dtype=np.float64)
        nx, nu = self.dims(cfg, opts)[:2]
        x = np.array(x0, dtype=np.float64).reshape(-1, nx).copy()
        u_des = np.ascontiguousarray(u_des, dtype=np.float64).reshape(-1, nu)
        n = x.shape[0]
        u = np.zeros((n, nu))
        rc = np.zeros(n, dtype=np.int32)
        hist = np.zeros(8, dtype=np.int64)
        r = self.lib.oracle_rollout(cfg, _d(o), 0 if o is None else o.size, n, steps, dt, _d(x), _d(u_des), _d(u),
                                    rc.ctypes.data_as(_ip), hist.ctypes.data_as(C.POINTER(C.c_int64)))
        assert r == 0
        return x, u, rc, hist

    def qp_solve(self, H, c, A, b, lb, ub, be=None, diagonal_cost=True):
        """A is (nc, nv); returns (status, sol)."""
        A = np.asarray(A, dtype=np.float64)
        nc, nv = A.shape
        Af = np.asfortranarray(A)
        Hf = np.asfortranarray(np.asarray(H, dtype=np.float64).reshape(nv, nv))
        c, b, lb, ub = (np.ascontiguousarray(v, dtype=np.float64) for v in (c, b, lb, ub))
        sol = np.zeros(nv)
        bep = None
        if be is not None:
            be = np.ascontiguousarray(be, dtype=np.uint8)
            bep = be.ctypes.data_as(C.POINTER(C.c_ubyte))
        st = self.lib.oracle_qp_solve(nv, nc, int(diagonal_cost), Hf.ctypes.data_as(_dp), _d(c), Af.ctypes.data_as(_dp),
                                      _d(b), _d(lb), _d(ub), bep, _d(sol))
        return st, sol


def realizable_export(ref_filter):
    """Kernel geometry + the x-independent facet interval table from the reference build (config 4)."""
    L = ref_filter.lib.lib
    L.ref_realizable_export.argtypes = [C.c_void_p, _ip, _dp, _dp, _ip, _ip, _dp]
    dims = np.zeros(4, dtype=np.int32)
    assert L.ref_realizable_export(ref_filter.h, dims.ctypes.data_as(_ip), None, None, None, None, None) == 0
    nV, nF, maxCrit, maxAct = (int(v) for v in dims)
    vertices = np.zeros((nV, 2))
    normals = np.zeros((nF, 2))
    fv = np.zeros((nF, 2), dtype=np.int32)
    fa = np.zeros((nF, maxAct), dtype=np.int32)
    lie = np.zeros((nF, maxAct, 4))
    assert L.ref_realizable_export(ref_filter.h, dims.ctypes.data_as(_ip), _d(vertices), _d(normals), fv.ctypes.data_as(_ip),
                                   fa.ctypes.data_as(_ip), _d(lie)) == 0
    return dict(vertices=vertices, normals=normals, facet_vertices=fv, facet_active=fa, facet_lie=lie,
                max_critical_facets=maxCrit, max_active_constraints=maxAct)


def learning_blob(dims, seed=7, scale=0.3):
    """Random networks for the learned residual in the blob layout ref_set_learning / oracle_set_learning /
    asif_engine_set_learning share: drift net (w1 [hidden x in] column-major, b1, w2, b2, w3, b3), then the actuation net."""
    g = np.random.Generator(np.random.Philox(key=seed))
    din, ain, dh1, ah1, dh2, ah2, dout, aout = (int(v) for v in dims)
    parts = []
    for (i, h1, h2, o) in ((din, dh1, dh2, dout), (ain, ah1, ah2, aout)):
        for (r, c) in ((h1, i), (h2, h1), (o, h2)):
            parts.append(g.normal(0, scale, r * c))
            parts.append(g.normal(0, scale, r))
    return np.concatenate(parts)
