"""ctypes binding of include/asif_b200.h.  No numerics here; every call goes to libasif_b200.so."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

FILTER_EXPLICIT, FILTER_IMPLICIT_TB, FILTER_IMPLICIT, FILTER_ROBUST, FILTER_REALIZABLE, FILTER_IMPLICIT_RB = 1, 2, 3, 4, 5, 6
MODEL_DOUBLE_INTEGRATOR, MODEL_DOUBLE_INTEGRATOR_TB, MODEL_INVERTED_PENDULUM, MODEL_INVERTED_PENDULUM_TABLE = 1, 2, 3, 4
MODEL_INVERTED_PENDULUM_KERNEL = 5
MODEL_SEGWAY, MODEL_SEGWAY_SHIPPED = 6, 7
MEM_HOST, MEM_DEVICE = 0, 1
QP_SHARED_H, QP_SHARED_BOUNDS = 1, 2

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class AsifError(RuntimeError):
    pass


class EngineConfig(C.Structure):
    """Mirror of struct asif_engine_config (include/asif_b200.h)."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("filter", C.c_int32), ("model", C.c_int32), ("device", C.c_int32),
        ("npBTSS", C.c_int32), ("npSSmax", C.c_int32),
        ("lb", C.c_double * 2), ("ub", C.c_double * 2),
        ("relaxCost", C.c_double), ("relaxLb", C.c_double), ("relaxReachLb", C.c_double),
        ("relaxTTS", C.c_double), ("relaxMinOrtho", C.c_double),
        ("backTrajHorizon", C.c_double), ("backTrajExtend", C.c_double), ("backTrajDt", C.c_double),
        ("backTrajMinOrtho", C.c_double), ("satSharpness", C.c_double), ("inf", C.c_double),
        ("relaxDes", C.c_double), ("relaxOffset", C.c_double),
        ("uncertaintyBounds", C.c_double * 4), ("dynParam", C.c_double * 4),
        ("halfplanes", _dp), ("n_halfplanes", C.c_int32),
        ("kernel_vertices", _dp), ("n_vertices", C.c_int32),
        ("facet_normals", _dp), ("facet_vertices", _ip), ("facet_active", _ip), ("facet_lie", _dp),
        ("n_facets", C.c_int32), ("max_critical_facets", C.c_int32), ("max_active_constraints", C.c_int32),
        ("backContDt", C.c_double), ("x_unc", C.c_double * 4),
    ]


class LearningData(C.Structure):
    """Mirror of struct asif_learning_data = ASIF::LearningData (include/asif_learning_utils.h:8-32)."""
    _fields_ = [(k, C.c_uint32) for k in ("d_drift_in", "d_act_in", "d_drift_hidden", "d_act_hidden", "d_drift_hidden_2",
                                          "d_act_hidden_2", "d_drift_out", "d_act_out")] + \
               [(k, _dp) for k in ("w_1_drift", "w_2_drift", "w_3_drift", "b_1_drift", "b_2_drift", "b_3_drift",
                                   "w_1_act", "w_2_act", "w_3_act", "b_1_act", "b_2_act", "b_3_act")]


class LoopConfig(C.Structure):
    """Mirror of struct asif_loop_config (include/asif_b200.h)."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("steps", C.c_int32), ("dt", C.c_double), ("steps_per_sample", C.c_int32),
        ("smooth_bounds", C.c_int32), ("smooth_lb", C.c_double), ("smooth_ub", C.c_double), ("smooth_rate", C.c_double),
        ("plant_gain", C.c_double), ("log_stride", C.c_int32), ("log_after_step", C.c_int32), ("log_agents", C.c_int64),
    ]


# field offsets inside one closed-loop log record, by name
def loop_log_fields(nx, nu, n_relax):
    names = ["t"] + ["x%d" % i for i in range(nx)] + ["xEstim%d" % i for i in range(nx)] + ["uDes%d" % i for i in range(nu)] + \
        ["uFilter%d" % i for i in range(nu)] + ["uAct%d" % i for i in range(nu)] + ["relax%d" % i for i in range(n_relax)] + \
        ["rc", "smoothLo", "smoothHi", "TTS", "BTorthoBS", "critIdx0"]
    return {k: i for i, k in enumerate(names)}


_lib = None


def lib_path():
    # ASIF_B200_LIB: an experimental build of the same library (scripts/build_variant.sh A/B runs); default = the in-tree product
    return os.environ.get("ASIF_B200_LIB") or os.path.join(HERE, "libasif_b200.so")


def load_library():
    """Load libasif_b200.so (in-tree).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise AsifError("%s is missing: build it with `python -m asif_b200._build` (nvcc, sm_100a). "
                        "There is no CPU fallback." % p)
    L = C.CDLL(p)
    L.asif_b200_abi_version.restype = C.c_int32
    L.asif_last_error.restype = C.c_char_p
    L.asif_device_count.restype = C.c_int32
    L.asif_engine_config_init.argtypes = [C.POINTER(EngineConfig), C.c_int32, C.c_int32]
    L.asif_engine_create.argtypes = [C.POINTER(EngineConfig), C.POINTER(C.c_void_p)]
    L.asif_engine_destroy.argtypes = [C.c_void_p]
    L.asif_engine_dims.argtypes = [C.c_void_p, _ip]
    L.asif_engine_filter_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.asif_engine_filter_batch_cost.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.asif_engine_set_input_cost.argtypes = [C.c_void_p, C.c_void_p]
    L.asif_engine_filter_batch_lie.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 8 + [C.c_int32, C.c_void_p]
    L.asif_engine_rollout.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int32, C.c_void_p]
    L.asif_engine_last_qp_iterations.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.asif_engine_last_host_io.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    L.asif_engine_host_io_stats.argtypes = [C.c_void_p, _dp, _ip]
    L.asif_engine_latency_server.argtypes = [C.c_void_p, C.c_int32]
    L.asif_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_uint64]
    L.asif_host_free.argtypes = [C.c_void_p]
    L.asif_host_register.argtypes = [C.c_void_p, C.c_uint64]
    L.asif_host_unregister.argtypes = [C.c_void_p]
    L.asif_engine_set_learning.argtypes = [C.c_void_p, C.POINTER(LearningData)]
    L.asif_loop_config_init.argtypes = [C.POINTER(LoopConfig)]
    L.asif_engine_loop_log_dims.argtypes = [C.c_void_p, C.POINTER(LoopConfig), C.POINTER(C.c_int64)]
    L.asif_engine_closed_loop.argtypes = [C.c_void_p, C.c_int64, C.POINTER(LoopConfig)] + [C.c_void_p] * 5 + \
        [C.POINTER(C.c_int64), C.c_void_p, C.c_int32, C.c_void_p]
    L.asif_qp_solve_batch.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32] + [C.c_void_p] * 9 + \
                                     [C.c_int32, C.c_int32, C.c_void_p]
    L.asif_qp_configure.argtypes = [C.c_double, C.c_int32, C.c_int32, C.c_int32]
    L.asif_qp_last_info.argtypes = [_ip]
    L.asif_measure_fp64_peak.argtypes = [C.c_int32, _dp, _dp]
    L.asif_engine_group_create.argtypes = [C.POINTER(EngineConfig), _ip, C.c_int32, C.POINTER(C.c_void_p)]
    L.asif_engine_group_destroy.argtypes = [C.c_void_p]
    L.asif_engine_group_size.argtypes = [C.c_void_p]
    L.asif_engine_group_engine.argtypes = [C.c_void_p, C.c_int32]
    L.asif_engine_group_engine.restype = C.c_void_p
    L.asif_engine_group_slice.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_int64)]
    L.asif_engine_group_filter_batch.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6
    L.asif_engine_group_filter_batch_cost.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 7
    L.asif_engine_group_rollout.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.POINTER(C.c_int64)]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise AsifError("asif_b200 error %d: %s" % (rc, load_library().asif_last_error().decode()))


def device_count():
    n = load_library().asif_device_count()
    if n < 0:
        raise AsifError(load_library().asif_last_error().decode())
    return n


def measure_fp64_peak(device=0):
    tf, mhz = C.c_double(), C.c_double()
    _check(load_library().asif_measure_fp64_peak(device, C.byref(tf), C.byref(mhz)))
    return tf.value, mhz.value


def _ptr(a):
    """Host numpy array or torch tensor (host or device) -> (address, is_device)."""
    if a is None:
        return None, None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "arrays must be C-contiguous"
        return a.ctypes.data, False
    # torch tensor (duck-typed so that torch stays optional)
    assert a.is_contiguous(), "tensors must be contiguous"
    return a.data_ptr(), bool(a.is_cuda)


class Engine:
    """One filter object = constructor + initialize(lb, ub, options) of the reference class."""

    def __init__(self, filter, model, device=0, **options):
        L = load_library()
        cfg = self._config(filter, model, device, options)
        self.cfg = cfg
        h = C.c_void_p()
        _check(L.asif_engine_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._read_dims(h)
        self.device = device

    def _read_dims(self, h):
        dims = (C.c_int32 * 6)()
        _check(load_library().asif_engine_dims(h, dims))
        self.nx, self.nu, self.n_relax, self.nc, self.nv, self.n_diag = (int(v) for v in dims)

    def _config(self, filter, model, device, options):
        L = load_library()
        cfg = EngineConfig()
        _check(L.asif_engine_config_init(C.byref(cfg), filter, model))
        cfg.device = device
        self._keep = []
        for k, v in options.items():
            if k in ("lb", "ub", "dynParam", "uncertaintyBounds", "x_unc"):
                for i, vi in enumerate(np.atleast_1d(v)):
                    getattr(cfg, k)[i] = float(vi)
            elif k == "kernel":
                # polytope kernel of ASIFrealizable: dict with vertices, normals, facet_vertices, facet_active, facet_lie
                kv = np.ascontiguousarray(v["vertices"], dtype=np.float64)
                kn = np.ascontiguousarray(v["normals"], dtype=np.float64)
                kfv = np.ascontiguousarray(v["facet_vertices"], dtype=np.int32)
                kfa = np.ascontiguousarray(v["facet_active"], dtype=np.int32)
                kl = np.ascontiguousarray(v["facet_lie"], dtype=np.float64)
                self._keep += [kv, kn, kfv, kfa, kl]
                cfg.kernel_vertices, cfg.n_vertices = kv.ctypes.data_as(_dp), kv.shape[0]
                cfg.facet_normals, cfg.n_facets = kn.ctypes.data_as(_dp), kn.shape[0]
                cfg.facet_vertices, cfg.facet_active = kfv.ctypes.data_as(_ip), kfa.ctypes.data_as(_ip)
                cfg.facet_lie = kl.ctypes.data_as(_dp)
                cfg.max_critical_facets = int(v["max_critical_facets"])
                cfg.max_active_constraints = int(kfa.shape[1])
            elif k == "halfplanes":
                t = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 2)
                self._keep.append(t)
                cfg.halfplanes = t.ctypes.data_as(_dp)
                cfg.n_halfplanes = t.shape[0]
            elif not hasattr(cfg, k):
                raise AsifError("unknown option %r" % k)
            else:
                setattr(cfg, k, v)
        return cfg

    def close(self):
        if getattr(self, "_h", None):
            load_library().asif_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- numpy convenience: allocates outputs on the host
    def filter_batch(self, x, u_des, want_diag=False):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, self.nx)
        u_des = np.ascontiguousarray(u_des, dtype=np.float64).reshape(-1, self.nu)
        n = x.shape[0]
        u = np.empty((n, self.nu))
        relax = np.empty((n, self.n_relax))
        rc = np.empty(n, dtype=np.int32)
        diag = np.empty((n, self.n_diag)) if want_diag else None
        self.filter_batch_into(n, x, u_des, u, relax, rc, diag)
        return (u, relax, rc, diag) if want_diag else (u, relax, rc)

    # -- raw: caller-owned numpy arrays (host) or torch tensors (host pinned / device)
    def filter_batch_into(self, n, x, u_des, u_act, relax, rc, diag=None, stream=None):
        ptrs = [_ptr(a) for a in (x, u_des, u_act, relax, rc)]
        dev = {d for _, d in ptrs}
        assert len(dev) == 1, "all batch arrays must live in the same memory space"
        mem = MEM_DEVICE if dev.pop() else MEM_HOST
        dptr = _ptr(diag)[0] if diag is not None else None
        _check(load_library().asif_engine_filter_batch(self._h, n, ptrs[0][0], ptrs[1][0], ptrs[2][0], ptrs[3][0],
                                                       ptrs[4][0], dptr, mem, stream))

    def filter_batch_cost(self, x, c, H=None, want_diag=False):
        """filter(x, H, c, uAct, relax) on a batch: c is (n, nv), H (nu, nu) or None (keep the current input Hessian)."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, self.nx)
        c = np.ascontiguousarray(c, dtype=np.float64).reshape(-1, self.nv)
        n = x.shape[0]
        assert c.shape[0] == n
        Hp = None
        if H is not None:
            Hc = np.asfortranarray(np.asarray(H, dtype=np.float64).reshape(self.nu, self.nu))
            Hp = Hc.ctypes.data
        u = np.empty((n, self.nu))
        relax = np.empty((n, self.n_relax))
        rc = np.empty(n, dtype=np.int32)
        diag = np.empty((n, self.n_diag)) if want_diag else None
        _check(load_library().asif_engine_filter_batch_cost(self._h, n, x.ctypes.data, Hp, c.ctypes.data, u.ctypes.data,
                                                            relax.ctypes.data, rc.ctypes.data,
                                                            diag.ctypes.data if want_diag else None, MEM_HOST, None))
        return (u, relax, rc, diag) if want_diag else (u, relax, rc)

    def filter_batch_lie(self, x, u_des, Lfh, Lgh, want_diag=False):
        """ASIF::filter(x, uDes, uAct, Lfh, Lgh, relax) on a batch: Lfh (n, nc), Lgh (n, nc*nu) per state column-major."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, self.nx)
        u_des = np.ascontiguousarray(u_des, dtype=np.float64).reshape(-1, self.nu)
        n = x.shape[0]
        Lfh = np.ascontiguousarray(Lfh, dtype=np.float64).reshape(n, self.nc)
        Lgh = np.ascontiguousarray(Lgh, dtype=np.float64).reshape(n, self.nc * self.nu)
        u, relax, rc = np.empty((n, self.nu)), np.empty((n, self.n_relax)), np.empty(n, dtype=np.int32)
        diag = np.empty((n, self.n_diag)) if want_diag else None
        _check(load_library().asif_engine_filter_batch_lie(self._h, n, x.ctypes.data, u_des.ctypes.data, Lfh.ctypes.data,
                                                           Lgh.ctypes.data, u.ctypes.data, relax.ctypes.data, rc.ctypes.data,
                                                           diag.ctypes.data if want_diag else None, MEM_HOST, None))
        return (u, relax, rc, diag) if want_diag else (u, relax, rc)

    def set_input_cost(self, H):
        Hc = np.asfortranarray(np.asarray(H, dtype=np.float64).reshape(self.nu, self.nu))
        _check(load_library().asif_engine_set_input_cost(self._h, Hc.ctypes.data))

    def rollout(self, x0, u_des, steps, dt):
        x = np.array(x0, dtype=np.float64).reshape(-1, self.nx).copy()
        u_des = np.ascontiguousarray(u_des, dtype=np.float64).reshape(-1, self.nu)
        n = x.shape[0]
        u = np.empty((n, self.nu))
        rc = np.empty(n, dtype=np.int32)
        hist = (C.c_int64 * 8)()
        _check(load_library().asif_engine_rollout(self._h, n, steps, dt, x.ctypes.data, u_des.ctypes.data, u.ctypes.data,
                                                  rc.ctypes.data, hist, MEM_HOST, None))
        return x, u, rc, np.array(list(hist), dtype=np.int64)

    def rollout_into(self, n, steps, dt, x, u_des, u_act_last, rc_last, want_hist=False, stream=None):
        ptrs = [_ptr(a) for a in (x, u_des, u_act_last, rc_last)]
        dev = {d for _, d in ptrs}
        assert len(dev) == 1
        mem = MEM_DEVICE if dev.pop() else MEM_HOST
        hist = (C.c_int64 * 8)() if want_hist else None
        _check(load_library().asif_engine_rollout(self._h, n, steps, dt, ptrs[0][0], ptrs[1][0], ptrs[2][0], ptrs[3][0],
                                                  hist, mem, stream))
        return np.array(list(hist), dtype=np.int64) if want_hist else None

    def closed_loop(self, x0, u_des, steps, dt=1e-3, steps_per_sample=1, smooth=None, plant_gain=1.0, log_stride=0,
                    log_agents=0, log_after_step=True):
        """The example main loop for a fleet (asif_engine_closed_loop).  smooth = (lb, ub, rate) enables the smoothBounds
        rate limiter.  Returns a dict: x, u_act, relax, rc (last filter call), rc_hist, log (agents, records, fields)."""
        L = load_library()
        cfg = LoopConfig()
        _check(L.asif_loop_config_init(C.byref(cfg)))
        cfg.steps, cfg.dt, cfg.steps_per_sample, cfg.plant_gain = int(steps), float(dt), int(steps_per_sample), float(plant_gain)
        cfg.log_stride, cfg.log_agents, cfg.log_after_step = int(log_stride), int(log_agents), int(bool(log_after_step))
        if smooth is not None:
            cfg.smooth_bounds = 1
            cfg.smooth_lb, cfg.smooth_ub, cfg.smooth_rate = (float(v) for v in smooth)
        x = np.array(x0, dtype=np.float64).reshape(-1, self.nx).copy()
        u_des = np.ascontiguousarray(u_des, dtype=np.float64).reshape(-1, self.nu)
        n = x.shape[0]
        u = np.empty((n, self.nu))
        relax = np.empty((n, self.n_relax))
        rc = np.empty(n, dtype=np.int32)
        hist = (C.c_int64 * 8)()
        dims = (C.c_int64 * 2)()
        _check(L.asif_engine_loop_log_dims(self._h, C.byref(cfg), dims))
        nlog = min(int(log_agents), n) if log_stride > 0 else 0
        log = np.zeros((nlog, int(dims[1]), int(dims[0]))) if nlog else None
        _check(L.asif_engine_closed_loop(self._h, n, C.byref(cfg), x.ctypes.data, u_des.ctypes.data, u.ctypes.data,
                                         relax.ctypes.data, rc.ctypes.data, hist, log.ctypes.data if nlog else None, MEM_HOST, None))
        return dict(x=x, u_act=u, relax=relax, rc=rc, rc_hist=np.array(list(hist), dtype=np.int64), log=log,
                    fields=loop_log_fields(self.nx, self.nu, self.n_relax))

    def set_learning(self, dims=None, blob=None):
        """Options.use_learning: dims = (d_drift_in, d_act_in, d_drift_hidden, d_act_hidden, d_drift_hidden_2,
        d_act_hidden_2, d_drift_out, d_act_out); blob = drift net (w1, b1, w2, b2, w3, b3; weights column-major) then the
        actuation net.  set_learning() switches the residual off."""
        L = load_library()
        if dims is None:
            _check(L.asif_engine_set_learning(self._h, None))
            return
        b = np.ascontiguousarray(blob, dtype=np.float64)
        din, ain, dh1, ah1, dh2, ah2, dout, aout = (int(v) for v in dims)
        ld = LearningData(din, ain, dh1, ah1, dh2, ah2, dout, aout)
        pos = 0
        for (i, h1, h2, o, tag) in ((din, dh1, dh2, dout, "drift"), (ain, ah1, ah2, aout, "act")):
            for name, ln in (("w_1", h1 * i), ("b_1", h1), ("w_2", h2 * h1), ("b_2", h2), ("w_3", o * h2), ("b_3", o)):
                setattr(ld, "%s_%s" % (name, tag), C.cast(b.ctypes.data + 8 * pos, _dp))
                pos += ln
        assert pos == b.size, "blob length does not match the layer widths"
        _check(L.asif_engine_set_learning(self._h, C.byref(ld)))

    def last_host_io(self):
        """HOST_IO_STAGED / HOST_IO_OUT / HOST_IO_INOUT of the last host-memory batch, -1 before the first one."""
        v = C.c_int32()
        _check(load_library().asif_engine_last_host_io(self._h, C.byref(v)))
        return int(v.value)

    def latency_server(self, on=True):
        """asif_engine_latency_server: a resident warp serves host batches of up to 32 states without a kernel launch"""
        _check(load_library().asif_engine_latency_server(self._h, 1 if on else 0))

    def host_io_stats(self):
        """{mode name: (ms per 1e6 states, batches measured)} of the "auto" host-IO policy"""
        ms, k = (C.c_double * 3)(), (C.c_int32 * 3)()
        _check(load_library().asif_engine_host_io_stats(self._h, ms, k))
        return {nm: (float(ms[i]), int(k[i])) for i, nm in enumerate(("staged", "out", "inout"))}

    def last_qp_iterations(self):
        v = C.c_uint64()
        _check(load_library().asif_engine_last_qp_iterations(self._h, C.byref(v)))
        return int(v.value)


HOST_IO_STAGED, HOST_IO_OUT, HOST_IO_INOUT = 0, 1, 2


class EngineGroup(Engine):
    """One filter object replicated on several GPUs (asif_engine_group_*): a host batch is cut into contiguous slices,
    one per device, and the results land in the caller's arrays.  devices=None takes every visible device."""

    def __init__(self, filter, model, devices=None, **options):
        L = load_library()
        cfg = self._config(filter, model, 0, options)
        self.cfg = cfg
        g = C.c_void_p()
        if devices is None:
            _check(L.asif_engine_group_create(C.byref(cfg), None, 0, C.byref(g)))
        else:
            d = (C.c_int32 * len(devices))(*devices)
            _check(L.asif_engine_group_create(C.byref(cfg), d, len(devices), C.byref(g)))
        self._g = g
        self._h = None
        self.size = int(L.asif_engine_group_size(g))
        self._read_dims(L.asif_engine_group_engine(g, 0))

    def close(self):
        if getattr(self, "_g", None):
            load_library().asif_engine_group_destroy(self._g)
            self._g = None

    def slice(self, n, i):
        b = (C.c_int64 * 2)()
        _check(load_library().asif_engine_group_slice(self._g, n, i, b))
        return int(b[0]), int(b[1])

    def member_host_io(self, i):
        v = C.c_int32()
        L = load_library()
        _check(L.asif_engine_last_host_io(L.asif_engine_group_engine(self._g, i), C.byref(v)))
        return int(v.value)

    def filter_batch_into(self, n, x, u_des, u_act, relax, rc, diag=None, stream=None):
        ptrs = [_ptr(a) for a in (x, u_des, u_act, relax, rc)]
        assert not any(d for _, d in ptrs), "group batches are host arrays"
        dptr = _ptr(diag)[0] if diag is not None else None
        _check(load_library().asif_engine_group_filter_batch(self._g, n, ptrs[0][0], ptrs[1][0], ptrs[2][0], ptrs[3][0], ptrs[4][0], dptr))

    def filter_batch_cost(self, x, c, H=None, want_diag=False):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, self.nx)
        c = np.ascontiguousarray(c, dtype=np.float64).reshape(-1, self.nv)
        n = x.shape[0]
        Hp = None
        if H is not None:
            Hc = np.asfortranarray(np.asarray(H, dtype=np.float64).reshape(self.nu, self.nu))
            Hp = Hc.ctypes.data
        u, relax, rc = np.empty((n, self.nu)), np.empty((n, self.n_relax)), np.empty(n, dtype=np.int32)
        diag = np.empty((n, self.n_diag)) if want_diag else None
        _check(load_library().asif_engine_group_filter_batch_cost(self._g, n, x.ctypes.data, Hp, c.ctypes.data, u.ctypes.data,
                                                                  relax.ctypes.data, rc.ctypes.data, diag.ctypes.data if want_diag else None))
        return (u, relax, rc, diag) if want_diag else (u, relax, rc)

    def rollout(self, x0, u_des, steps, dt):
        x = np.array(x0, dtype=np.float64).reshape(-1, self.nx).copy()
        u_des = np.ascontiguousarray(u_des, dtype=np.float64).reshape(-1, self.nu)
        n = x.shape[0]
        u, rc = np.empty((n, self.nu)), np.empty(n, dtype=np.int32)
        hist = (C.c_int64 * 8)()
        _check(load_library().asif_engine_group_rollout(self._g, n, steps, dt, x.ctypes.data, u_des.ctypes.data, u.ctypes.data,
                                                        rc.ctypes.data, hist))
        return x, u, rc, np.array(list(hist), dtype=np.int64)


class PinnedArray:
    """numpy view of memory from asif_host_alloc (pinned, device-addressable); free() or the destructor releases it."""

    def __init__(self, shape, dtype=np.float64):
        shape = (shape,) if np.isscalar(shape) else tuple(shape)
        dt = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
        p = C.c_void_p()
        _check(load_library().asif_host_alloc(C.byref(p), nbytes))
        self._p = p.value
        buf = (C.c_char * nbytes).from_address(self._p) if nbytes else b""
        self.array = np.frombuffer(buf, dtype=dt).reshape(shape)

    def free(self):
        if getattr(self, "_p", None):
            self.array = None
            load_library().asif_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def qp_solve_batch(H, c, A, b, lb, ub, be=None, diagonal_cost=True, device=0):
    """n QPs  min v'Hv + c'v, A v >= b, lb <= v <= ub.

    c: (n, nv); A: (n, nc, nv) in natural (row, col) indexing (converted to the reference's
    column-major layout here); b: (n, nc); H: (nv, nv) shared or (n, nv, nv); lb/ub: (nv,) shared or (n, nv).
    Returns (sol (n, nv), status (n,)).
    """
    c = np.ascontiguousarray(c, dtype=np.float64)
    n, nv = c.shape
    A = np.asarray(A, dtype=np.float64).reshape(n, -1, nv)
    nc = A.shape[1]
    Acm = np.ascontiguousarray(np.transpose(A, (0, 2, 1)))  # per problem column-major
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(n, nc)
    H = np.asarray(H, dtype=np.float64)
    share = 0
    if H.ndim == 2:
        share |= QP_SHARED_H
        Hcm = np.ascontiguousarray(H.T)
    else:
        Hcm = np.ascontiguousarray(np.transpose(H, (0, 2, 1)))
    lb = np.ascontiguousarray(lb, dtype=np.float64)
    ub = np.ascontiguousarray(ub, dtype=np.float64)
    if lb.ndim == 1:
        share |= QP_SHARED_BOUNDS
    bep = None
    if be is not None:
        be = np.ascontiguousarray(be, dtype=np.uint8)
        bep = be.ctypes.data
    sol = np.empty((n, nv))
    status = np.empty(n, dtype=np.int32)
    _check(load_library().asif_qp_solve_batch(device, nv, nc, n, int(diagonal_cost), Hcm.ctypes.data, c.ctypes.data,
                                              Acm.ctypes.data, b.ctypes.data, lb.ctypes.data, ub.ctypes.data, bep,
                                              sol.ctypes.data, status.ctypes.data, share, MEM_HOST, None))
    return sol, status


def qp_configure(eps=0.0, max_iter=0, polish=-1, polish_refine_iter=-1):
    """Accuracy settings of the nv > 4 solver (asif_qp_configure); the defaults restore 1e-8 / 20000 / polish / 10."""
    _check(load_library().asif_qp_configure(float(eps), int(max_iter), int(polish), int(polish_refine_iter)))


def qp_last_info():
    """(ADMM iterations, rho updates, polish state, active rows, us equilibration / factorisations / iterations / polish)
    of the first problem of this thread's last host call."""
    info = np.zeros(8, dtype=np.int32)
    _check(load_library().asif_qp_last_info(info.ctypes.data_as(_ip)))
    return tuple(int(v) for v in info)
