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
