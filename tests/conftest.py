import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def philox(seed):
    """Counter-based generator (Philox4x64) so that inputs are reproducible anywhere."""
    return np.random.Generator(np.random.Philox(key=seed))


# ---- the named workloads ----------------------------------------------------
SEED = 0xA51F

# option vectors in the oracle / reference-harness order
C2_TB_OPTS = [50.0, 10.0, 5.0, 5.0, 10.0, 0.0, 0.1, 0.01, 0.1]          # DoubleIntegrator_implicit_tb, npBT = 101
SEGWAY_TB_OPTS = [10.0, 2.0, 30.0, 60.0, 3.0, 0.05, 0.01, 0.001, 0.1, 1.0]  # example options, centred backup set
C1_OPTS = [5.0, 50.0]


def c1_inputs(n, seed=SEED + 1):
    g = philox(seed)
    return g.uniform(-1, 1, (n, 2)), g.uniform(-1, 1, (n, 1))


def c2_inputs(n, seed=SEED + 2):
    g = philox(seed)
    return g.uniform(-1, 1, (n, 2)), g.uniform(-1, 1, (n, 1))


SEGWAY_THETA0 = 0.1383244254


def c5_inputs(n, seed=SEED + 5):
    g = philox(seed)
    u = g.uniform(-1, 1, (n, 4))
    x = np.stack([0.5 * u[:, 0], 0.5 * u[:, 1], SEGWAY_THETA0 + 0.2 * u[:, 2], 0.5 * u[:, 3]], axis=1)
    return np.ascontiguousarray(x), g.uniform(-5, 5, (n, 1))


def tb_engine_kwargs(opts):
    return dict(relaxCost=opts[0], relaxLb=opts[1], relaxTTS=opts[2], relaxMinOrtho=opts[3], backTrajHorizon=opts[4],
                backTrajExtend=opts[5], backTrajDt=opts[6], backTrajMinOrtho=opts[7], satSharpness=opts[8])


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyref
    if not os.path.exists(pyref.ORACLE_SO):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return pyref.OracleLib()


@pytest.fixture(scope="session")
def reflib():
    from oracle import pyref
    if not os.path.exists(pyref.REF_SO):
        if os.path.isdir("/root/reference"):
            import subprocess
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "ref_build"), "-j8"])
        else:
            pytest.skip("oracle/_ref/libasif_ref.so not built and /root/reference absent")
    L = pyref.RefLib()
    L.set_qp_mode()
    return L


def assert_parity(name, got, want, rc_mask=None):
    """North-star tolerances: |du| <= 1e-6 + 1e-5|u|, identical rc, relax to the same tolerance."""
    u, relax, rc = got
    u0, relax0, rc0 = want
    assert rc.shape == rc0.shape
    mism = np.nonzero(rc != rc0)[0]
    assert mism.size == 0, "%s: %d rc mismatches, first %s: got %s want %s" % (
        name, mism.size, mism[:5], rc[mism[:5]], rc0[mism[:5]])
    tol_u = 1e-6 + 1e-5 * np.abs(u0)
    bad = np.nonzero(np.abs(u - u0) > tol_u)[0]
    assert bad.size == 0, "%s: %d states exceed the u tolerance, worst %g" % (name, bad.size, np.abs(u - u0).max())
    ok = rc0 > 0
    tol_r = 1e-6 + 1e-5 * np.abs(relax0)
    badr = np.nonzero((np.abs(relax - relax0) > tol_r)[ok])[0]
    assert badr.size == 0, "%s: %d states exceed the relax tolerance" % (name, badr.size)


# ---- configs 3a / 3b (InvertedPendulum) -----------------------------------------------------------
# implicit: [relaxCost, relaxReachLb, relaxSafeLb, backTrajHorizon, backTrajDt, satSharpness]
C3A_OPTS = [50.0, 5.0, 10.0, 5.0, 0.001, 0.1]        # example options: npBT = 5001
C3A_SHORT_OPTS = [50.0, 5.0, 10.0, 5.0, 0.05, 0.1]   # same horizon, dt 0.05: npBT = 101 (cheap enough for big CPU samples)
C3B_OPTS = [5.0, 50.0, 0.8, 1.2]                     # robust: [relaxLb, relaxCost, pMin, pMax]


def c3a_inputs(n, seed=SEED + 3):
    g = philox(seed)
    return g.uniform(-np.pi, np.pi, (n, 2)), g.uniform(-1.5, 1.5, (n, 1))


def c3b_inputs(n, seed=SEED + 4):
    g = philox(seed)
    return g.uniform(-2.5, 2.5, (n, 2)), g.uniform(-1.5, 1.5, (n, 1))


def implicit_engine_kwargs(opts):
    return dict(relaxCost=opts[0], relaxReachLb=opts[1], relaxLb=opts[2], backTrajHorizon=opts[3], backTrajDt=opts[4],
                satSharpness=opts[5])


def halfplane_table():
    return np.load(os.path.join(ROOT, "tests", "golden", "halfplanes_70_135.npy"))


def unpinned_mask(rc_ref, rc_exact, relax_exact, qp_status=None):
    """States whose reference return code comes from the OSQP algorithm's inexactness and cannot be matched by an
    exact solver (SURVEY 8c "known gap"): the ADMM loop hit its iteration limit (-2) or stopped at the 10x looser
    "inaccurate" thresholds (2, 3, 4) - known exactly when the raw OSQP status was recorded - and primal-infeasible
    verdicts on problems whose exact optimum exists with a very large relaxation variable."""
    big = np.abs(relax_exact).max(axis=1) > 50.0
    m = np.isin(rc_ref, (-2, 3, 4)) | (np.isin(rc_ref, (-3, -1)) & (rc_exact == 1) & big)
    if qp_status is not None:
        solved_qp = rc_ref != 2  # TB states inside the backup set still solve a (trivial) QP; status is theirs
        m |= np.isin(qp_status, (-2, 2, 3, 4)) & (rc_ref != rc_exact) & solved_qp
    return m


# Disagreements of the committed reference vectors with an EXACT solver, counted with the oracle in this container (round 2):
# every one of them is a state where the reference build's ADMM stand-in left its loop inexactly (unpinned_mask).  The
# golden tests compare EVERY state, require that each disagreement is such a state, and bound their number by these
# measured counts (no percentage masks).
GOLDEN_UNPINNED = {"c1_di_explicit": 0, "c2_di_implicit_tb": 2, "c3a_ip_implicit_short": 15, "c3a_ip_implicit": 0, "c3b_ip_robust": 3,
                   "c4_ip_realizable": 7, "c5_segway_tb": 0, "rb_ip_implicit": 6, "rb_di_implicit": 5}


def disagree(got, want):
    """states on which two result sets differ by the north-star bar (rc; u or relax outside 1e-6 + 1e-5 |.|)"""
    u, relax, rc = got
    u0, relax0, rc0 = want
    du = np.abs(u - u0).max(axis=1) > 1e-6 + 1e-5 * np.abs(u0).max(axis=1)
    dr = np.abs(relax - relax0).max(axis=1) > 1e-6 + 1e-5 * np.abs(relax0).max(axis=1)
    return (rc != rc0) | du | (dr & (rc0 == 1))


def assert_golden_parity(name, got, g, relax_cols=None, knife_edge=0):
    """got = (u, relax, rc) against the golden record g (u_act, relax, rc, qp_status).  All states are compared; a
    disagreement must lie in the unpinned class and there may be at most GOLDEN_UNPINNED[name] (+ knife_edge flips for
    the models whose device libm differs from glibc in the last bit) of them."""
    u, relax, rc = got
    rl, rl0 = relax, g["relax"]
    if relax_cols is not None:
        rl, rl0 = relax[:, relax_cols], g["relax"][:, relax_cols]
    D = disagree((u, rl, rc), (g["u_act"], rl0, g["rc"]))
    unp = unpinned_mask(g["rc"], rc, relax, g["qp_status"] if "qp_status" in g.files else None)
    print("golden %s: %d states, %d disagree with the reference build, %d of them unpinned (budget %d)" % (
        name, len(rc), D.sum(), (D & unp).sum(), GOLDEN_UNPINNED[name]))
    assert (D & ~unp).sum() <= knife_edge, "%s: unexplained disagreements at %s" % (name, np.nonzero(D & ~unp)[0][:10])
    assert (D & unp).sum() <= GOLDEN_UNPINNED[name] + knife_edge
    return ~D


EXTRA_GOLDEN_JOBS = [
    lambda: ("c3a_ip_implicit_short", 3, C3A_SHORT_OPTS, c3a_inputs(1200)),
    lambda: ("c3a_ip_implicit", 3, C3A_OPTS, c3a_inputs(60, seed=SEED + 33)),
    lambda: ("c3b_ip_robust", 4, C3B_OPTS, c3b_inputs(160)),
]

# ---- config 4 (ASIFrealizable, IP dynamics + RealizableKernelData_100Hz_50pt) -----------------------
# [relaxDes, relaxOffset, relaxCost, unc0, unc1, npSSmax]
C4_OPTS = [1.0, 5.0, 50.0, 0.032, 0.027, 2.0]


def realizable_kernel():
    return np.load(os.path.join(ROOT, "tests", "golden", "realizable_kernel_100hz_50pt.npz"))


def c4_inputs(n, seed=SEED + 6):
    """Half of the states uniform on [-3,3]^2, half scattered around the kernel boundary so that
    the critical-facet machinery is exercised (uniform sampling alone makes a facet critical for < 2 % of states)."""
    g = philox(seed)
    k = realizable_kernel()
    V = k["vertices"]
    na = n // 2
    xa = g.uniform(-3, 3, (na, 2))
    idx = g.integers(0, len(V), n - na)
    t = g.uniform(0, 1, (n - na, 1))
    xb = V[idx] * t + V[(idx + 1) % len(V)] * (1 - t) + g.normal(0, 0.03, (n - na, 2))
    return np.ascontiguousarray(np.vstack([xa, xb])), g.uniform(-1.5, 1.5, (n, 1))


def realizable_engine_kwargs(opts, kernel=None):
    k = realizable_kernel() if kernel is None else kernel
    return dict(relaxDes=opts[0], relaxOffset=opts[1], relaxCost=opts[2], uncertaintyBounds=[opts[3], opts[4]],
                npSSmax=int(opts[5]), dynParam=[0.9, 1.1], kernel=k)


EXTRA_GOLDEN_JOBS.append(lambda: ("c4_ip_realizable", 5, C4_OPTS, c4_inputs(1200)))


# ---- ASIFimplicitRB (no example ships): pendulum callbacks (split gradients, cfg 7) and
# DoubleIntegrator_implicit_tb callbacks (fused gradient, cfg 8) ---------------------------------------
# [relaxCost, relaxReachLb, relaxSafeLb, backTrajHorizon, backTrajDt, satSharpness, npBTSS, backContDt, x_unc0, x_unc1]
RB_IP_OPTS = [50.0, 5.0, 10.0, 5.0, 0.05, 0.1, 10.0, 0.2, 0.05, 0.08]   # npBT 101, hold refreshed every 4th step
RB_DI_OPTS = [20.0, 1.0, 2.0, 3.0, 0.02, 0.5, 6.0, 0.1, 0.05, 0.01]     # npBT 151, 6 critical points, hold refreshed every 5th step
RB_DI_NP10_OPTS = [50.0, 5.0, 10.0, 2.0, 0.01, 0.1, 10.0, 0.05, 0.02, 0.03]  # npBT 201, 10 critical points (unrolled kernel)
# (a hold of 0.3 s at dt 0.1 drives this loop into an exactly periodic orbit: tied min-h keys, whose order std::sort leaves open)
RB_IP_NP4_OPTS = [50.0, 5.0, 10.0, 2.0, 0.01, 0.5, 4.0, 0.033, 0.1, 0.0]  # 4 critical points, odd hold period, wide bevel


def rb_engine_kwargs(opts):
    return dict(relaxCost=opts[0], relaxReachLb=opts[1], relaxLb=opts[2], backTrajHorizon=opts[3], backTrajDt=opts[4],
                satSharpness=opts[5], npBTSS=int(opts[6]), backContDt=opts[7], x_unc=[opts[8], opts[9]])


EXTRA_GOLDEN_JOBS.append(lambda: ("rb_ip_implicit", 7, RB_IP_OPTS, c3a_inputs(800, seed=SEED + 71)))
EXTRA_GOLDEN_JOBS.append(lambda: ("rb_di_implicit", 8, RB_DI_OPTS, c2_inputs(800, seed=SEED + 72)))


# ---- learned residual of the implicit classes (Options.use_learning, include/asif_learning_utils.h) -------------
# (d_drift_in, d_act_in, d_drift_hidden, d_act_hidden, d_drift_hidden_2, d_act_hidden_2, d_drift_out, d_act_out)
LEARN_DIMS = [6, 5, 16, 12, 8, 10, 2, 1]


def learning_blob(dims=LEARN_DIMS, seed=7, scale=0.3):
    from oracle import pyref
    return pyref.learning_blob(dims, seed=seed, scale=scale)


# ---- filter(x, H, c, uAct, relax) overloads: a caller-supplied input Hessian block and full linear cost ------------
def custom_cost(ud, nv, seed, relax_c=(-2 * 50 * 7.0, -2 * 50 * 4.0), h00=2.5):
    """H = [h00]; c = [-2 h00 uDes * U(0.5,1.5), relax entries * U(0.8,1.2)] per state."""
    g = philox(seed)
    n = ud.shape[0]
    c = np.concatenate([-2 * h00 * ud * g.uniform(0.5, 1.5, (n, 1)),
                        np.tile(np.asarray(relax_c[:nv - 1]), (n, 1)) * g.uniform(0.8, 1.2, (n, nv - 1))], axis=1)
    return np.array([[h00]]), np.ascontiguousarray(c)
