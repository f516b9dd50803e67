"""asif_engine_closed_loop (the loop around filter() in the example programs -- sample-and-hold,
smoothBounds, Euler plant, per-step log) against a numpy restatement of the example main loops that calls the oracle
for every filter() evaluation.  B200 box."""
import numpy as np
import pytest

import conftest as cf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ab():
    import asif_b200
    if asif_b200.device_count() < 1:
        pytest.fail("no CUDA device: the engine has no CPU fallback")
    return asif_b200


def example_loop(filter_fn, plant, x0, ud, steps, dt, k=1, smooth=None, hold=False):
    """examples/DoubleIntegrator_RealizableSampled.cpp:110-205 (the other example loops are special cases).
    Returns final x, last (u_act, relax, rc), rc histogram, and per-step records."""
    x = x0.copy()
    n = x.shape[0]
    s_lo = np.full(n, smooth[0]) if smooth else None
    s_hi = np.full(n, smooth[1]) if smooth else None
    hist = np.zeros(8, dtype=np.int64)
    recs = []
    t = 0.0
    skip = 0
    u_f = np.zeros_like(ud)
    for s in range(steps):
        if skip == 0:
            x_est = x.copy()
            u_new, relax, rc = filter_fn(x_est, ud)
            # ASIF / ASIFrobust / ASIFrealizable leave uAct untouched when the QP fails: the loop keeps the previous output
            u_f = np.where((rc < 1)[:, None], u_f, u_new) if hold else u_new
            u_a = u_f.copy()
            if smooth:
                sc = smooth[2]
                ua, uds = u_a[:, 0], ud[:, 0]
                up = (ua > uds) & (ua > s_lo)
                dn = ~up & (ua < uds) & (ua < s_hi)
                other = ~up & ~dn
                s_lo = np.where(up, ua, s_lo)
                s_hi = np.where(up, np.where(s_lo > s_hi, s_lo, s_hi + sc), s_hi)
                s_hi = np.where(dn, ua, s_hi)
                s_lo = np.where(dn, np.where(s_lo > s_hi, s_hi, s_lo - sc), s_lo)
                s_lo = np.where(other, s_lo - sc, s_lo)
                s_hi = np.where(other, s_hi + sc, s_hi)
                s_lo = np.maximum(s_lo, smooth[0])
                s_hi = np.minimum(s_hi, smooth[1])
                u_a[:, 0] = np.maximum(np.minimum(ua, s_hi), s_lo)
            for r in rc:
                hist[r + 3 if -3 <= r <= 2 else 7] += 1
        skip = (skip + 1) % k
        rec = dict(t=t, x=x.copy(), x_est=x_est.copy(), u_f=u_f.copy(), u_a=u_a.copy(), relax=relax.copy(), rc=rc.copy())
        f, g = plant(x)
        x = x + dt * ((0.0 + f) + g * u_a)
        t += dt
        rec["t_after"], rec["x_after"] = t, x.copy()
        recs.append(rec)
    return x, u_a, relax, rc, hist, recs


def di_plant(x):
    return np.stack([x[:, 1], np.zeros(len(x))], 1), np.tile([0.0, 1.0], (len(x), 1))


def pendulum_plant(gain):
    return lambda x: (np.stack([x[:, 1], np.sin(x[:, 0])], 1), np.tile([0.0, gain], (len(x), 1)))


def test_closed_loop_explicit_bit_exact_plant(ab, oracle):
    n, steps, dt = 500, 40, 1e-3
    x0, ud = cf.c1_inputs(n, seed=301)
    x0 *= 0.7
    eng = ab.Engine(ab.FILTER_EXPLICIT, ab.MODEL_DOUBLE_INTEGRATOR, relaxLb=cf.C1_OPTS[0], relaxCost=cf.C1_OPTS[1])
    out = eng.closed_loop(x0, ud, steps, dt, log_stride=1, log_agents=n, log_after_step=False)

    def filt(x, u):
        ua, rl, rc, _ = oracle.filter_batch(1, x, u, cf.C1_OPTS, True)
        return ua, rl, rc
    x, ua, rl, rc, hist, recs = example_loop(filt, di_plant, x0, ud, steps, dt, hold=True)
    assert np.array_equal(out["rc"], rc) and np.array_equal(out["rc_hist"], hist)
    assert np.abs(out["x"] - x).max() < 1e-9
    F = out["fields"]
    log = out["log"]
    assert log.shape == (n, steps, len(F))
    for s in (0, 7, steps - 1):
        assert np.abs(log[:, s, F["t"]] - recs[s]["t"]).max() == 0.0  # tNow accumulates += dt, bit for bit
        assert np.abs(log[:, s, F["x0"]:F["x0"] + 2] - recs[s]["x"]).max() < 1e-9
        assert np.array_equal(log[:, s, F["rc"]].astype(np.int32), recs[s]["rc"])
        assert np.abs(log[:, s, F["uAct0"]] - recs[s]["u_a"][:, 0]).max() < 1e-6


def test_closed_loop_tb_equals_fused_rollout(ab):
    """steps_per_sample = 1, no limiter: the generic loop and the fused TB rollout kernel are the same computation."""
    n, steps, dt = 2000, 25, 1e-3
    x0, ud = cf.c2_inputs(n, seed=302)
    x0 *= 0.5
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS))
    out = eng.closed_loop(x0, ud, steps, dt, log_stride=5, log_agents=100)
    x, u, rc, hist = eng.rollout(x0, ud, steps, dt)
    assert np.array_equal(out["x"], x) and np.array_equal(out["u_act"], u) and np.array_equal(out["rc"], rc)
    assert np.array_equal(out["rc_hist"], hist)
    F = out["fields"]
    log = out["log"]
    assert log.shape[:2] == (100, 5)
    # the TB diagnostics of the logged agents are those of filter_batch at the sampled state
    xe = log[:, 3, F["xEstim0"]:F["xEstim0"] + 2]
    _, _, rc3, diag = eng.filter_batch(np.ascontiguousarray(xe), ud[:100], want_diag=True)
    assert np.array_equal(log[:, 3, F["TTS"]], diag[:, 0]) and np.array_equal(log[:, 3, F["BTorthoBS"]], diag[:, 1])
    assert np.array_equal(log[:, 3, F["rc"]].astype(np.int32), rc3)


def test_closed_loop_sampled_realizable_with_rate_limiter(ab, oracle):
    """dtPerSample = 10 and smoothBounds, on the polytope-kernel pendulum with a plant gain inside [pMin, pMax]."""
    n, steps, dt, k = 300, 60, 1e-3, 10
    x0, ud = cf.c4_inputs(n, seed=303)
    ud = ud * 2.0
    smooth = (-1.5, 1.5, 1.5 * k * 0.001)
    eng = ab.Engine(ab.FILTER_REALIZABLE, ab.MODEL_INVERTED_PENDULUM_KERNEL, **cf.realizable_engine_kwargs(cf.C4_OPTS))
    out = eng.closed_loop(x0, ud, steps, dt, steps_per_sample=k, smooth=smooth, plant_gain=1.05, log_stride=1, log_agents=n,
                          log_after_step=False)

    def filt(x, u):
        ua, rl, rc, _ = oracle.filter_batch(5, x, u, cf.C4_OPTS, True)
        return ua, rl, rc
    x, ua, rl, rc, hist, recs = example_loop(filt, pendulum_plant(1.05), x0, ud, steps, dt, k=k, smooth=smooth, hold=True)
    assert hist.sum() == n * (steps // k) and out["rc_hist"].sum() == hist.sum()
    # The limiter compares uAct with its own previous output (`uAct < smoothBounds[1]`), an exact tie whenever the filter
    # output repeats; a last-bit difference in u between the two QP solvers then takes the other branch.  Those agents
    # (a few per cent) are excluded; all others must agree along the whole run.
    ok = np.abs(out["x"] - x).max(axis=1) < 1e-7
    assert ok.mean() > 0.9
    assert np.array_equal(out["rc"][ok], rc[ok])
    assert np.all(np.abs(out["u_act"][ok] - ua[ok]) <= 1e-6 + 1e-5 * np.abs(ua[ok]))
    F = out["fields"]
    log = out["log"]
    for s in (0, 9, 10, 35, steps - 1):
        assert np.abs(log[ok, s, F["xEstim0"]:F["xEstim0"] + 2] - recs[s]["x_est"][ok]).max() < 1e-7  # held between samples
        assert np.abs(log[ok, s, F["uFilter0"]] - recs[s]["u_f"][ok, 0]).max() < 1e-5
        assert np.abs(log[ok, s, F["uAct0"]] - recs[s]["u_a"][ok, 0]).max() < 1e-5
        assert np.all(log[:, s, F["smoothLo"]] <= log[:, s, F["uAct0"]] + 1e-12)
        assert np.all(log[:, s, F["uAct0"]] <= log[:, s, F["smoothHi"]] + 1e-12)
    assert (np.abs(log[:, :, F["uAct0"]] - log[:, :, F["uFilter0"]]) > 1e-9).any()  # the limiter did act


def test_closed_loop_implicit_pendulum(ab, oracle):
    n, steps, dt = 64, 20, 1e-3
    x0, ud = cf.c3a_inputs(n, seed=304)
    x0 *= 0.3
    eng = ab.Engine(ab.FILTER_IMPLICIT, ab.MODEL_INVERTED_PENDULUM, **cf.implicit_engine_kwargs(cf.C3A_SHORT_OPTS))
    out = eng.closed_loop(x0, ud, steps, dt, log_stride=1, log_agents=8)

    def filt(x, u):
        ua, rl, rc, _ = oracle.filter_batch(3, x, u, cf.C3A_SHORT_OPTS, True)
        return ua, rl, rc
    x, ua, rl, rc, hist, recs = example_loop(filt, pendulum_plant(1.0), x0, ud, steps, dt)
    assert np.array_equal(out["rc_hist"], hist)
    assert np.abs(out["x"] - x).max() < 1e-7
    F = out["fields"]
    # log_after_step (the default): the record holds (t + dt, x after the step) as InvertedPendulum_Implicit.cpp:113-147 writes it
    assert np.abs(out["log"][:, 4, F["t"]] - recs[4]["t_after"]).max() == 0.0
    assert np.abs(out["log"][:, 4, F["x0"]:F["x0"] + 2] - recs[4]["x_after"][:8]).max() < 1e-7
    assert out["log"].shape[2] == len(F) and F["relax1"] == F["relax0"] + 1
