"""Oracle restatement of ASIFimplicit / ASIFrobust (InvertedPendulum configs 3a, 3b) against golden
vectors produced by the reference's own sources.  CPU only."""
import os

import numpy as np
import pytest

import conftest as cf

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run(oracle, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg, opts = int(g["cfg"]), list(g["opts"])
    u, relax, rc, diag = oracle.filter_batch(cfg, g["x"], g["u_des"], opts, want_diag=True)
    k = cf.assert_golden_parity(name, (u, relax, rc), g)  # every state compared; disagreements counted against the measured budget
    return g, diag, k, rc


def test_c3a_implicit_short_golden(oracle):
    g, diag, k, rc = run(oracle, "c3a_ip_implicit_short")
    m = k & (g["rc"] == 1)
    assert np.array_equal(diag[m], g["diag"][m])  # hSafetyNow, hBackupEnd, critical indices, A_, b_: same bits
    assert (g["rc"] == -1).sum() > 50 and (g["rc"] == 1).sum() > 50


def test_c3a_implicit_example_options_golden(oracle):
    """npBT = 5001 (the example's own horizon and step)."""
    g, diag, k, rc = run(oracle, "c3a_ip_implicit")
    m = k & (g["rc"] == 1)
    assert np.array_equal(diag[m], g["diag"][m])


def test_c3b_robust_golden(oracle):
    """The reduced 2-variable problem against the reference's 402-variable LP-dual formulation (libaffa intervals)."""
    g, diag, k, rc = run(oracle, "c3b_ip_robust")
    # h_k, [Lg-, Lg+], [Lf-, Lf+] read back from the reference's A_ (all states: assembly does not depend on the QP)
    assert np.abs(diag - g["diag"]).max() <= 1e-12
    assert (g["rc"] == -1).sum() > 5


def test_c4_realizable_golden(oracle):
    """Reduced 2-variable problem + exact segment/box test against the reference's 38-variable LP-dual QP with OSQP
    facet-feasibility solves.  States where one of the reference's QPs did not converge are excluded (counted)."""
    g = np.load(os.path.join(GOLD, "c4_ip_realizable.npz"))
    u, relax, rc, diag = oracle.filter_batch(5, g["x"], g["u_des"], list(g["opts"]), want_diag=True)
    tainted = np.isin(g["qp_status"], (-2, 2, 3, 4))
    assert tainted.mean() < 0.03
    k = ~tainted
    assert np.array_equal(diag[k][:, :6], g["diag"][k][:, :6])
    assert np.array_equal(diag[k], g["diag"][k])  # table gathers and barrier rows: same bits
    # relax[0] of the reference is a non-unique LP-dual multiplier: relax[1] (eps) is compared, on EVERY state
    cf.assert_golden_parity("c4_ip_realizable", (u, relax, rc), g, relax_cols=[1])
    assert (g["rc"] == -2).sum() > 50 and (g["diag"][:, 0] >= 1).sum() > 100


@pytest.mark.parametrize("name", ["rb_ip_implicit", "rb_di_implicit"])
def test_implicit_rb_golden(oracle, name):
    """ASIFimplicitRB fixtures from the reference build (libaffa interval safety set, zero-order-hold backup input)."""
    g, diag, k, rc = run(oracle, name)
    assert np.array_equal(diag, g["diag"])  # every state: rows do not depend on the QP
    assert (g["rc"] == -1).sum() > 50 and (g["rc"] == 1).sum() > 50
