"""CUDA path vs oracle / golden for config 4 (ASIFrealizable, IP dynamics + 100Hz_50pt polytope kernel); B200 box."""
import os

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


def test_c4_realizable_vs_oracle(ab, oracle):
    n = 50_000
    x, ud = cf.c4_inputs(n)
    eng = ab.Engine(ab.FILTER_REALIZABLE, ab.MODEL_INVERTED_PENDULUM_KERNEL, **cf.realizable_engine_kwargs(cf.C4_OPTS))
    u, relax, rc, diag = eng.filter_batch(x, ud, want_diag=True)
    u0, relax0, rc0, diag0 = oracle.filter_batch(5, x, ud, cf.C4_OPTS, want_diag=True)
    print("realizable rc", dict(zip(*np.unique(rc0, return_counts=True))), "nCrit",
          dict(zip(*np.unique(diag0[:, 0], return_counts=True))), "QP rows/state", eng.last_qp_iterations() / n)
    # everything except sin(x0) in the barrier rows is the same arithmetic: sets must be identical
    assert np.array_equal(diag[:, :6], diag0[:, :6])
    cf.assert_parity("realizable", (u, relax, rc), (u0, relax0, rc0))
    assert np.abs(diag - diag0).max() <= 1e-12
    u2, relax2, rc2 = eng.filter_batch(x, ud)
    assert np.array_equal(rc, rc2) and np.array_equal(u, u2) and np.array_equal(relax, relax2)
    assert (rc0 == -2).sum() > 100 and (rc0 == -1).sum() > 100 and (diag0[:, 0] >= 2).sum() > 100


def test_c4_realizable_golden(ab):
    g = np.load(os.path.join(cf.ROOT, "tests", "golden", "c4_ip_realizable.npz"))
    eng = ab.Engine(ab.FILTER_REALIZABLE, ab.MODEL_INVERTED_PENDULUM_KERNEL, **cf.realizable_engine_kwargs(list(g["opts"])))
    u, relax, rc, diag = eng.filter_batch(g["x"], g["u_des"], want_diag=True)
    tainted = np.isin(g["qp_status"], (-2, 2, 3, 4))  # a facet-feasibility or main QP of the OSQP stand-in did not converge
    assert tainted.mean() < 0.03
    k = ~tainted
    assert np.array_equal(diag[k][:, :6], g["diag"][k][:, :6])  # critical facets and barrier facets
    assert np.abs(diag[k] - g["diag"][k]).max() <= 1e-12
    # relax[0] of the reference is a non-unique LP-dual multiplier: relax[1] is compared, on EVERY state
    cf.assert_golden_parity("c4_ip_realizable", (u, relax, rc), g, relax_cols=[1])
