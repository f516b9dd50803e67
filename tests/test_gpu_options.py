"""Option / bound variants of the TB filter that exercise every saturation mode of the kernel (identity, power-of-two
scale, general division), wide bevels (satSharpness), horizons, dt, cost weights.  CUDA vs oracle; B200 box."""
import numpy as np
import pytest

import conftest as cf

pytestmark = pytest.mark.gpu

VARIANTS = [
    # (lb, ub, overrides of C2_TB_OPTS by index)
    ((-1.0, 1.0), {}),                               # SAT_IDENTITY
    ((-2.0, 2.0), {}),                               # SAT_POW2 (range 4)
    ((-0.5, 0.5), {8: 0.5}),                         # SAT_POW2 (range 1), wider bevel
    ((-0.7, 1.1), {8: 1.0}),                         # SAT_GENERAL, asymmetric bounds, bevel radius 1
    ((-1.0, 1.0), {8: 1.7}),                         # bevel radius close to the clamp of updateOptions (2)
    ((-1.0, 1.0), {4: 3.0, 6: 0.05, 5: 0.05}),       # horizon 3, dt 0.05, extend 5 %: npBT = 64
    ((-3.0, 3.0), {0: 5.0, 1: 2.0, 2: 30.0, 3: 60.0, 7: 0.001}),  # segway-like weights on the double integrator
]


@pytest.fixture(scope="module")
def ab():
    import asif_b200
    if asif_b200.device_count() < 1:
        pytest.fail("no CUDA device: the engine has no CPU fallback")
    return asif_b200


@pytest.mark.parametrize("case", range(len(VARIANTS)))
def test_tb_option_variants(ab, oracle, case):
    (lb, ub), over = VARIANTS[case]
    opts = list(cf.C2_TB_OPTS)
    for k, v in over.items():
        opts[k] = v
    n = 30_000
    x, ud = cf.c2_inputs(n, seed=500 + case)
    ud = ud * 1.5 * max(abs(lb), abs(ub))
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, lb=[lb], ub=[ub], **cf.tb_engine_kwargs(opts))
    u, relax, rc, diag = eng.filter_batch(x, ud, want_diag=True)
    u0, relax0, rc0, diag0 = oracle.filter_batch(2, x, ud, opts + [0.0, lb, ub], want_diag=True)
    print("variant", case, "rc", dict(zip(*np.unique(rc0, return_counts=True))))
    cf.assert_parity("variant %d" % case, (u, relax, rc), (u0, relax0, rc0))
    m = rc0 == 1
    assert m.sum() > 200
    assert np.array_equal(diag[m][:, :3], diag0[m][:, :3]) and np.array_equal(diag[m][:, 4:], diag0[m][:, 4:])
    assert u.min() >= lb and u.max() <= ub


@pytest.mark.parametrize("npSSmax", [1, 2, 3, 7])
def test_explicit_row_selection(ab, oracle, npSSmax):
    """npSSmax < npSS: the npSSmax smallest-h rows, ascending (src/asif.cpp:250-268)."""
    n = 50_000
    x, ud = cf.c1_inputs(n, seed=90 + npSSmax)
    x[: n // 10, 0] = np.round(x[: n // 10, 0], 1)  # exact ties between safety functions
    x[: n // 10, 1] = x[: n // 10, 0]
    opts = list(cf.C1_OPTS[:2]) + [float(npSSmax)]
    eng = ab.Engine(ab.FILTER_EXPLICIT, ab.MODEL_DOUBLE_INTEGRATOR, relaxLb=opts[0], relaxCost=opts[1], npSSmax=npSSmax)
    nc = min(npSSmax, 4)
    assert eng.nc == nc and eng.n_diag == 3 * nc
    u, relax, rc, diag = eng.filter_batch(x, ud, want_diag=True)
    u0, relax0, rc0, diag0 = oracle.filter_batch(1, x, ud, opts, want_diag=True)
    assert np.array_equal(diag, diag0)
    cf.assert_parity("npSSmax", (u, relax, rc), (u0, relax0, rc0))
    out2 = eng.filter_batch(x, ud)
    assert np.array_equal(out2[0], u) and np.array_equal(out2[2], rc)


@pytest.mark.parametrize("model,npBTSS", [("di", 1), ("di", 3), ("di", 8), ("segway", 2), ("segway", 6), ("ip", 4), ("ip", 16)])
def test_other_critical_point_counts(ab, oracle, model, npBTSS):
    """npBTSS other than the examples' 4 / 10 runs the run-time-count instantiation (capacity 8 / 16)."""
    if model == "di":
        n, cfg = 30_000, 2
        opts = list(cf.C2_TB_OPTS) + [0.0, -1.0, 1.0, float(npBTSS)]
        x, ud = cf.c2_inputs(n, seed=700 + npBTSS)
        eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, npBTSS=npBTSS, **cf.tb_engine_kwargs(cf.C2_TB_OPTS))
        head, nc = 4 + npBTSS, npBTSS * 4 + 2
    elif model == "segway":
        n, cfg = 4_000, 6
        opts = list(cf.SEGWAY_TB_OPTS) + [-20.0, 20.0, float(npBTSS)]
        x, ud = cf.c5_inputs(n, seed=700 + npBTSS)
        eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_SEGWAY, npBTSS=npBTSS, **cf.tb_engine_kwargs(cf.SEGWAY_TB_OPTS))
        head, nc = 4 + npBTSS, npBTSS * 4 + 2
    else:
        n, cfg = 20_000, 3
        opts = list(cf.C3A_SHORT_OPTS) + [float(npBTSS)]
        x, ud = cf.c3a_inputs(n, seed=700 + npBTSS)
        eng = ab.Engine(ab.FILTER_IMPLICIT, ab.MODEL_INVERTED_PENDULUM, npBTSS=npBTSS, **cf.implicit_engine_kwargs(cf.C3A_SHORT_OPTS))
        head, nc = 2 + npBTSS, npBTSS * 4 + 1
    assert eng.nc == nc and eng.n_diag == head + nc * eng.nv + nc
    u, relax, rc, diag = eng.filter_batch(x, ud, want_diag=True)
    u0, relax0, rc0, diag0 = oracle.filter_batch(cfg, x, ud, opts, want_diag=True)
    flips = int((rc != rc0).sum())
    assert flips <= (0 if model == "di" else 2)
    keep = rc == rc0
    cf.assert_parity("npBTSS %s %d" % (model, npBTSS), (u[keep], relax[keep], rc[keep]), (u0[keep], relax0[keep], rc0[keep]))
    lo = 4 if model != "ip" else 2
    m = keep & (rc0 == 1) & np.all(diag[:, lo:lo + npBTSS] == diag0[:, lo:lo + npBTSS], axis=1)
    assert m.sum() > 0.95 * (keep & (rc0 == 1)).sum()
    cols = [c for c in range(diag.shape[1]) if not (model != "ip" and c == 3)]  # hBackupEnd_ is a previous-call diagnostic
    d = np.abs(diag[m][:, cols] - diag0[m][:, cols]) / (1.0 + np.abs(diag0[m][:, cols]))
    assert d.max() <= (0.0 if model == "di" else 1e-9)
    out2 = eng.filter_batch(x, ud)
    assert np.array_equal(out2[0], u) and np.array_equal(out2[2], rc)
    if model == "di":  # the fused rollout uses the same instantiation
        xr, ur, rcr, hist = eng.rollout(x[:2000], ud[:2000], 3, 1e-3)
        xs = x[:2000].copy()
        for _ in range(3):
            ua = eng.filter_batch(xs, ud[:2000])[0]
            xs = xs + 1e-3 * np.stack([xs[:, 1], ua[:, 0]], axis=1)
        assert np.abs(xr - xs).max() < 1e-12


def test_critical_point_count_limits(ab):
    with pytest.raises(ab.AsifError):
        ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, npBTSS=9)
    with pytest.raises(ab.AsifError):
        ab.Engine(ab.FILTER_IMPLICIT, ab.MODEL_INVERTED_PENDULUM, npBTSS=17)


@pytest.mark.parametrize("cfg,filt,model,opts,gen,kw", [
    (1, "FILTER_EXPLICIT", "MODEL_DOUBLE_INTEGRATOR", cf.C1_OPTS, cf.c1_inputs, lambda o: dict(relaxLb=o[0], relaxCost=o[1])),
    (2, "FILTER_IMPLICIT_TB", "MODEL_DOUBLE_INTEGRATOR_TB", cf.C2_TB_OPTS, cf.c2_inputs, cf.tb_engine_kwargs),
    (3, "FILTER_IMPLICIT", "MODEL_INVERTED_PENDULUM", cf.C3A_SHORT_OPTS, cf.c3a_inputs, cf.implicit_engine_kwargs),
    (7, "FILTER_IMPLICIT_RB", "MODEL_INVERTED_PENDULUM", cf.RB_IP_OPTS, cf.c3a_inputs, cf.rb_engine_kwargs),
    (6, "FILTER_IMPLICIT_TB", "MODEL_SEGWAY", cf.SEGWAY_TB_OPTS, cf.c5_inputs, cf.tb_engine_kwargs),
])
def test_custom_cost(oracle, cfg, filt, model, opts, gen, kw):
    """asif_engine_filter_batch_cost = the filter(x, H, c, uAct, relax) overloads, against the oracle; H stays in force
    for later uDes calls (updateH), and H = I with the default c reproduces filter(x, uDes) bit for bit."""
    import asif_b200 as ab
    n = 20_000 if cfg != 6 else 4000
    x, ud = gen(n, seed=cf.SEED + 700 + cfg)
    eng = ab.Engine(getattr(ab, filt), getattr(ab, model), **kw(opts))
    base = eng.filter_batch(x, ud, want_diag=True)
    H, c = cf.custom_cost(ud, eng.nv, seed=cfg)
    u, relax, rc, diag = eng.filter_batch_cost(x, c, H, want_diag=True)
    u0, relax0, rc0, diag0 = oracle.filter_batch_cost(cfg, x, c, H, opts, want_diag=True)
    flips = np.nonzero(rc != rc0)[0]
    print("custom cost cfg", cfg, "rc", dict(zip(*np.unique(rc0, return_counts=True))), "flips", flips.size)
    assert flips.size <= (0 if cfg in (1, 2) else max(2, n // 2000))
    keep = rc == rc0
    cf.assert_parity("custom cost", (u[keep], relax[keep], rc[keep]), (u0[keep], relax0[keep], rc0[keep]))
    if cfg == 1:
        assert np.array_equal(diag, diag0)
    if cfg == 2:  # rows of the states that build rows (rc 1); column 3 is hBackupEnd, which the kernel leaves unset without a hit
        m = rc0 == 1
        assert np.array_equal(diag[m][:, :3], diag0[m][:, :3]) and np.array_equal(diag[m][:, 4:], diag0[m][:, 4:])
    # sticky H: filter(x, uDes) now minimises 2.5 (u - uDes)^2 + ..., i.e. c = [-2 uDes, defaults] with H = 2.5
    u1, relax1, rc1 = eng.filter_batch(x, ud)
    if cfg == 1:      # explicit: [relaxLb, relaxCost]
        relax_defaults = [-2 * opts[1] * opts[0]]
    elif cfg in (2, 6):  # TB: [relaxCost, relaxSafeLb, ...]
        relax_defaults = [-2 * opts[0] * opts[1]]
    else:             # implicit / RB: [relaxCost, relaxReachLb, relaxSafeLb, ...]; v = (u, delta_safe, delta_reach)
        relax_defaults = [-2 * opts[0] * opts[2], -2 * opts[0] * opts[1]]
    cdef = np.concatenate([-2.0 * ud, np.tile(relax_defaults, (n, 1))], axis=1)
    u2, relax2, rc2 = eng.filter_batch_cost(x, cdef, None)
    assert np.array_equal(rc1, rc2) and np.array_equal(u1, u2) and np.array_equal(relax1, relax2)
    # back to the identity block: the very first results return
    eng.set_input_cost(np.eye(1))
    again = eng.filter_batch(x, ud, want_diag=True)
    for p, q in zip(base, again):
        assert np.array_equal(p, q)
    u3, relax3, rc3 = eng.filter_batch_cost(x, cdef, np.eye(1))
    assert np.array_equal(base[2], rc3) and np.array_equal(base[0], u3) and np.array_equal(base[1], relax3)


def test_custom_cost_refused_for_lp_dual_classes():
    import asif_b200 as ab
    o = cf.C3B_OPTS
    eng = ab.Engine(ab.FILTER_ROBUST, ab.MODEL_INVERTED_PENDULUM_TABLE, relaxLb=o[0], relaxCost=o[1], dynParam=[o[2], o[3]],
                    halfplanes=cf.halfplane_table())
    with pytest.raises(ab.AsifError):
        eng.filter_batch_cost(np.zeros((4, 2)), np.zeros((4, 2)), None)


@pytest.mark.parametrize("npSSmax", [0, 2])
def test_explicit_custom_lie_derivatives(oracle, npSSmax):
    """asif_engine_filter_batch_lie: caller-supplied Lfh / Lgh for the explicit filter; this call only (not sticky)."""
    import asif_b200 as ab
    n = 50_000
    opts = list(cf.C1_OPTS) + ([float(npSSmax)] if npSSmax else [])
    x, ud = cf.c1_inputs(n, seed=cf.SEED + 810 + npSSmax)
    eng = ab.Engine(ab.FILTER_EXPLICIT, ab.MODEL_DOUBLE_INTEGRATOR, relaxLb=opts[0], relaxCost=opts[1], npSSmax=npSSmax or -1)
    base = eng.filter_batch(x, ud, want_diag=True)
    g = cf.philox(6)
    Lfh, Lgh = g.normal(0, 1, (n, eng.nc)), g.normal(0, 1, (n, eng.nc))
    u, relax, rc, diag = eng.filter_batch_lie(x, ud, Lfh, Lgh, want_diag=True)
    u0, relax0, rc0, diag0 = oracle.filter_batch_lie(x, ud, Lfh, Lgh, opts)
    assert np.array_equal(diag, diag0)
    cf.assert_parity("lie", (u, relax, rc), (u0, relax0, rc0))
    again = eng.filter_batch(x, ud, want_diag=True)
    for p, q in zip(base, again):
        assert np.array_equal(p, q)


@pytest.mark.parametrize("mem", ["pageable", "pinned", "device"])
def test_custom_cost_first_call_on_fresh_engine(ab, oracle, mem):
    """The call pattern of scripts/parity_report.py's last record: a fresh engine whose FIRST call
    is asif_engine_filter_batch_cost, n = 1e5 (several pipeline chunks), diagnostics on; through pageable numpy arrays,
    pinned arrays and device pointers.  filter(x, H, c, ...) = src/asif_implicit_tb.cpp:252-259, updateH :748-762."""
    n = 100_000
    x, ud = cf.c2_inputs(n, seed=cf.SEED + 93)
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS))
    H, c = cf.custom_cost(ud, eng.nv, seed=2)
    want = oracle.filter_batch_cost(2, x, c, H, cf.C2_TB_OPTS, True)
    if mem == "pageable":
        u, relax, rc, diag = eng.filter_batch_cost(x, c, H, want_diag=True)
    else:
        import ctypes as C
        import torch
        L = ab.load_library()
        Hc = np.asfortranarray(H)
        t = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in dict(x=x, c=c).items()}
        t.update(u=torch.empty((n, eng.nu), dtype=torch.float64), relax=torch.empty((n, eng.n_relax), dtype=torch.float64),
                 rc=torch.empty(n, dtype=torch.int32), diag=torch.empty((n, eng.n_diag), dtype=torch.float64))
        t = {k: (v.pin_memory() if mem == "pinned" else v.cuda()) for k, v in t.items()}
        r = L.asif_engine_filter_batch_cost(eng._h, n, t["x"].data_ptr(), Hc.ctypes.data, t["c"].data_ptr(), t["u"].data_ptr(),
                                            t["relax"].data_ptr(), t["rc"].data_ptr(), t["diag"].data_ptr(),
                                            ab.MEM_DEVICE if mem == "device" else ab.MEM_HOST, None)
        assert r == 0, L.asif_last_error()
        u, relax, rc, diag = (t[k].cpu().numpy() for k in ("u", "relax", "rc", "diag"))
    u0, relax0, rc0, diag0 = want
    print("first-call custom cost", mem, "rc", dict(zip(*np.unique(rc0, return_counts=True))), "flips", int((rc != rc0).sum()))
    cf.assert_parity("first-call custom cost / " + mem, (u, relax, rc), (u0, relax0, rc0))
    m = rc0 == 1
    assert np.array_equal(diag[m][:, :3], diag0[m][:, :3]) and np.array_equal(diag[m][:, 4:], diag0[m][:, 4:])
