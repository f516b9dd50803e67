"""The oracle restatement against the reference compiled from its own sources (oracle/_ref), on
fresh random states.  Runs where libasif_ref.so exists (dev container: built on demand from
/root/reference; GPU box: shipped prebuilt); skipped otherwise.  CPU only."""
import numpy as np
import pytest

import conftest as cf
from oracle import pyref


@pytest.mark.parametrize("cfg,opts,gen,n", [
    (pyref.CFG_DI_EXPLICIT, cf.C1_OPTS, cf.c1_inputs, 20000),
    (pyref.CFG_DI_IMPLICIT_TB, cf.C2_TB_OPTS, cf.c2_inputs, 6000),
    (pyref.CFG_SEGWAY_TB, cf.SEGWAY_TB_OPTS, cf.c5_inputs, 600),
    (pyref.CFG_IP_IMPLICIT, cf.C3A_SHORT_OPTS, cf.c3a_inputs, 1500),
])
def test_oracle_matches_reference(oracle, reflib, cfg, opts, gen, n):
    x, ud = gen(n, seed=cf.SEED + 100 + cfg)
    f = reflib.create(cfg, opts)
    u0, relax0, rc0, diag0, st0, it0 = f.filter_batch_ex(x, ud)
    u, relax, rc, diag = oracle.filter_batch(cfg, x, ud, opts, want_diag=True)
    unpinned = cf.unpinned_mask(rc0, rc, relax, st0)
    print("cfg", cfg, "rc", dict(zip(*np.unique(rc0, return_counts=True))), "unpinned", int(unpinned.sum()))
    assert unpinned.mean() < 0.03
    k = ~unpinned
    cf.assert_parity("cfg%d" % cfg, (u[k], relax[k], rc[k]), (u0[k], relax0[k], rc0[k]))
    m = k & (rc0 == 1)
    assert np.array_equal(diag[m], diag0[m])


def test_rollout_matches_reference(oracle, reflib):
    f = reflib.create(pyref.CFG_DI_IMPLICIT_TB, cf.C2_TB_OPTS)
    x0, ud = cf.c2_inputs(50, seed=99)
    x0 *= 0.3
    xr, ur, rcr, hr = f.rollout(x0, ud, 40, 1e-3)
    xo, uo, rco, ho = oracle.rollout(pyref.CFG_DI_IMPLICIT_TB, x0, ud, 40, 1e-3, cf.C2_TB_OPTS)
    assert np.array_equal(rcr, rco)
    assert np.abs(xr - xo).max() < 1e-9


@pytest.mark.parametrize("lb,ub,sat", [(-2.0, 2.0, 0.1), (-0.7, 1.1, 1.0), (-1.0, 1.0, 1.7)])
def test_tb_other_bounds_and_sharpness(oracle, reflib, lb, ub, sat):
    """initialize(lb, ub) with bounds other than the example's and wide saturation bevels."""
    opts = list(cf.C2_TB_OPTS) + [0.0, lb, ub]
    opts[8] = sat
    x, ud = cf.c2_inputs(1200, seed=int(abs(lb) * 10))
    ud = ud * 2
    f = reflib.create(pyref.CFG_DI_IMPLICIT_TB, opts)
    u0, relax0, rc0, diag0, st0, it0 = f.filter_batch_ex(x, ud)
    u, relax, rc, diag = oracle.filter_batch(pyref.CFG_DI_IMPLICIT_TB, x, ud, opts, want_diag=True)
    k = ~cf.unpinned_mask(rc0, rc, relax, st0)
    assert k.mean() > 0.98
    cf.assert_parity("bounds", (u[k], relax[k], rc[k]), (u0[k], relax0[k], rc0[k]))
    m = k & (rc0 == 1)
    assert np.array_equal(diag[m], diag0[m])


@pytest.mark.parametrize("npSSmax", [1, 2, 3, 9])
def test_explicit_row_selection(oracle, reflib, npSSmax):
    """ASIF::ASIF with npSSmax < npSS keeps the npSSmax smallest-h rows in ascending order (src/asif.cpp:21-22,250-268)."""
    opts = list(cf.C1_OPTS[:2]) + [float(npSSmax)]
    x, ud = cf.c1_inputs(2500, seed=70 + npSSmax)
    f = reflib.create(pyref.CFG_DI_EXPLICIT, opts)
    u0, relax0, rc0, diag0, st0, it0 = f.filter_batch_ex(x, ud)
    u, relax, rc, diag = oracle.filter_batch(pyref.CFG_DI_EXPLICIT, x, ud, opts, want_diag=True)
    nc = min(npSSmax, 4)
    assert diag.shape[1] == nc * 2 + nc and diag0.shape == diag.shape
    assert np.array_equal(diag, diag0)
    cf.assert_parity("npSSmax", (u, relax, rc), (u0, relax0, rc0))
    if nc < 4:  # selected h values (column 1 of A_) ascend
        hsel = diag[:, nc:2 * nc]
        assert np.all(np.diff(hsel, axis=1) >= 0)


@pytest.mark.parametrize("cfg,npBTSS", [(pyref.CFG_DI_IMPLICIT_TB, 1), (pyref.CFG_DI_IMPLICIT_TB, 3), (pyref.CFG_DI_IMPLICIT_TB, 8),
                                        (pyref.CFG_IP_IMPLICIT, 4), (pyref.CFG_IP_IMPLICIT, 16), (pyref.CFG_SEGWAY_TB, 2)])
def test_other_critical_point_counts(oracle, reflib, cfg, npBTSS):
    """npBTSS is a constructor argument of the implicit classes (include/asif_implicit_tb.h:36-48): other values than the
    examples' 4 / 10 change the number of rows, the padding and the diag layout."""
    if cfg == pyref.CFG_DI_IMPLICIT_TB:
        opts, (x, ud) = list(cf.C2_TB_OPTS) + [0.0, -1.0, 1.0, float(npBTSS)], cf.c2_inputs(600, seed=40 + npBTSS)
    elif cfg == pyref.CFG_SEGWAY_TB:
        opts, (x, ud) = list(cf.SEGWAY_TB_OPTS) + [-20.0, 20.0, float(npBTSS)], cf.c5_inputs(150, seed=40 + npBTSS)
    else:
        opts, (x, ud) = list(cf.C3A_SHORT_OPTS) + [float(npBTSS)], cf.c3a_inputs(600, seed=40 + npBTSS)
    f = reflib.create(cfg, opts)
    u0, relax0, rc0, diag0, st0, it0 = f.filter_batch_ex(x, ud)
    u, relax, rc, diag = oracle.filter_batch(cfg, x, ud, opts, want_diag=True)
    npss = 4
    nc = npBTSS * npss + (1 if cfg == pyref.CFG_IP_IMPLICIT else 2)
    assert f.nc == nc and diag.shape == diag0.shape
    k = ~cf.unpinned_mask(rc0, rc, relax, st0)
    assert k.mean() > 0.9
    cf.assert_parity("npBTSS", (u[k], relax[k], rc[k]), (u0[k], relax0[k], rc0[k]))
    m = k & (rc0 == 1)
    assert m.sum() > 20 and np.array_equal(diag[m], diag0[m])


@pytest.mark.parametrize("cfg,opts,gen,n", [
    (pyref.CFG_IP_IMPLICIT_RB, cf.RB_IP_OPTS, cf.c3a_inputs, 1500),
    (pyref.CFG_IP_IMPLICIT_RB, cf.RB_IP_NP4_OPTS, cf.c3a_inputs, 800),
    (pyref.CFG_DI_IMPLICIT_RB, cf.RB_DI_OPTS, cf.c2_inputs, 1500),
    (pyref.CFG_DI_IMPLICIT_RB, cf.RB_DI_NP10_OPTS, cf.c2_inputs, 800),
    (pyref.CFG_IP_IMPLICIT_RB, None, cf.c3a_inputs, 500),  # Options() defaults of the class
])
def test_implicit_rb_matches_reference(oracle, reflib, cfg, opts, gen, n):
    """ASIFimplicitRB (src/asif_implicit_robust.cpp): zero-order-hold backup input, interval lower bound of h through
    libaffa.  Every row of A_, b_ and the critical indices must carry the reference's bits for every state (the
    assembly does not depend on the QP); u / relax / rc within tolerance wherever the OSQP stand-in converged."""
    x, ud = gen(n, seed=cf.SEED + 200 + cfg)
    f = reflib.create(cfg, opts)
    u0, relax0, rc0, diag0, st0, it0 = f.filter_batch_ex(x, ud)
    u, relax, rc, diag = oracle.filter_batch(cfg, x, ud, opts, want_diag=True)
    assert np.array_equal(diag, diag0)
    unpinned = cf.unpinned_mask(rc0, rc, relax, st0)
    print("cfg", cfg, "rc", dict(zip(*np.unique(rc0, return_counts=True))), "unpinned", int(unpinned.sum()))
    assert unpinned.mean() < 0.03
    k = ~unpinned
    cf.assert_parity("rb cfg%d" % cfg, (u[k], relax[k], rc[k]), (u0[k], relax0[k], rc0[k]))
    assert (rc0 == 1).sum() > 50 and (rc0 == -1).sum() > 50


def test_implicit_rb_reduces_to_implicit(oracle, reflib):
    """x_unc = 0 and a hold period equal to the Euler step: ASIFimplicitRB is ASIFimplicit (reference vs reference, and
    oracle vs oracle), bit for bit."""
    x, ud = cf.c3a_inputs(300, seed=5)
    o = list(cf.C3A_SHORT_OPTS) + [10.0]
    f3, f7 = reflib.create(pyref.CFG_IP_IMPLICIT, o), reflib.create(pyref.CFG_IP_IMPLICIT_RB, o + [o[4], 0.0, 0.0])
    d3, d7 = f3.filter_batch(x, ud, True)[3], f7.filter_batch(x, ud, True)[3]
    assert np.array_equal(d3, d7)
    e3 = oracle.filter_batch(pyref.CFG_IP_IMPLICIT, x, ud, o, want_diag=True)[3]
    e7 = oracle.filter_batch(pyref.CFG_IP_IMPLICIT_RB, x, ud, o + [o[4], 0.0, 0.0], want_diag=True)[3]
    assert np.array_equal(e3, e7) and np.array_equal(e3, d3)


@pytest.mark.parametrize("cfg,opts,gen", [
    (pyref.CFG_IP_IMPLICIT, cf.C3A_SHORT_OPTS, cf.c3a_inputs),
    (pyref.CFG_IP_IMPLICIT_RB, cf.RB_IP_OPTS, cf.c3a_inputs),
    (pyref.CFG_DI_IMPLICIT_RB, cf.RB_DI_OPTS, cf.c2_inputs),
])
def test_learned_residual_matches_reference(oracle, reflib, cfg, opts, gen):
    """Options.use_learning with random networks in learning_data_ (include/asif_learning_utils.h:121-155): Lfh[0] and
    Lgh[0] move by the two MLP outputs; rows bit-identical to the reference build for every state."""
    n = 1000
    x, ud = gen(n, seed=cf.SEED + 400 + cfg)
    blob = cf.learning_blob()
    f = reflib.create(cfg, opts)
    d_off = f.filter_batch(x, ud, want_diag=True)[3]
    f.set_learning(cf.LEARN_DIMS, blob)
    u0, relax0, rc0, diag0, st0, it0 = f.filter_batch_ex(x, ud)
    oracle.set_learning(cf.LEARN_DIMS, blob)
    try:
        u, relax, rc, diag = oracle.filter_batch(cfg, x, ud, opts, want_diag=True)
    finally:
        oracle.set_learning()
    assert np.array_equal(diag, diag0)
    changed = diag0 != d_off
    nb, nc = int(round(f.n_diag)) and (f.nc - 1) // 4, f.nc
    # exactly two entries per state differ from the run without the residual: A_[0,0] and b_[0]
    A0, b0 = 2 + nb, 2 + nb + nc * f.nv
    assert changed[:, A0].all() and changed[:, b0].all() and changed.sum() == 2 * n
    k = ~cf.unpinned_mask(rc0, rc, relax, st0)
    assert k.mean() > 0.95
    cf.assert_parity("learning cfg%d" % cfg, (u[k], relax[k], rc[k]), (u0[k], relax0[k], rc0[k]))


@pytest.mark.parametrize("cfg,opts,gen,n", [
    (pyref.CFG_DI_EXPLICIT, cf.C1_OPTS, cf.c1_inputs, 3000),
    (pyref.CFG_DI_IMPLICIT_TB, cf.C2_TB_OPTS, cf.c2_inputs, 2000),
    (pyref.CFG_IP_IMPLICIT, cf.C3A_SHORT_OPTS, cf.c3a_inputs, 800),
    (pyref.CFG_IP_IMPLICIT_RB, cf.RB_IP_OPTS, cf.c3a_inputs, 600),
    (pyref.CFG_SEGWAY_TB, cf.SEGWAY_TB_OPTS, cf.c5_inputs, 200),
])
def test_custom_cost_matches_reference(oracle, reflib, cfg, opts, gen, n):
    """The filter(x, H, c, uAct, relax) overloads of the reference classes (e.g. src/asif_implicit_tb.cpp:252-363)."""
    x, ud = gen(n, seed=cf.SEED + 600 + cfg)
    f = reflib.create(cfg, opts)
    H, c = cf.custom_cost(ud, f.nv, seed=cfg)
    u0, relax0, rc0, diag0, st0 = f.filter_batch_cost(x, c, H)
    u, relax, rc, diag = oracle.filter_batch_cost(cfg, x, c, H, opts, want_diag=True)
    k = ~cf.unpinned_mask(rc0, rc, relax, st0)
    assert k.mean() > 0.97
    # the ADMM stand-in (eps 1e-8 + polish) is itself only accurate to a few 1e-6 on the worst-conditioned of these
    # problems (relaxations of several hundred against an input weight of 2.5): such states are counted and bounded,
    # every other one must meet the north-star tolerance
    loose = k & (rc == rc0) & (np.abs(u - u0).max(axis=1) > 1e-6 + 1e-5 * np.abs(u0).max(axis=1))
    assert loose.mean() < 0.005 and (not loose.any() or np.abs(u - u0)[loose].max() < 5e-5)
    k &= ~loose
    cf.assert_parity("cost cfg%d" % cfg, (u[k], relax[k], rc[k]), (u0[k], relax0[k], rc0[k]))
    m = k & (rc0 == 1)
    assert np.array_equal(diag[m], diag0[m])
    # the cost matters: the same states through filter(x, uDes) give other inputs
    u1 = oracle.filter_batch(cfg, x, ud, opts)[0]
    assert np.abs(u1 - u)[rc == 1].max() > 1e-3


@pytest.mark.parametrize("npSSmax", [0, 2])
def test_explicit_custom_lie_derivatives(oracle, reflib, npSSmax):
    """ASIF::filter(x, uDes, uAct, Lfh, Lgh, relax) (src/asif.cpp:125-141,287-292): caller-supplied Lfh / Lgh rows."""
    opts = list(cf.C1_OPTS) + ([float(npSSmax)] if npSSmax else [])
    x, ud = cf.c1_inputs(3000, seed=cf.SEED + 800 + npSSmax)
    f = reflib.create(pyref.CFG_DI_EXPLICIT, opts)
    g = cf.philox(5)
    Lfh, Lgh = g.normal(0, 1, (3000, f.nc)), g.normal(0, 1, (3000, f.nc))
    u0, relax0, rc0, diag0 = f.filter_batch_lie(x, ud, Lfh, Lgh)
    u, relax, rc, diag = oracle.filter_batch_lie(x, ud, Lfh, Lgh, opts)
    assert np.array_equal(diag, diag0)
    cf.assert_parity("lie", (u, relax, rc), (u0, relax0, rc0))
    assert (rc0 == -1).sum() > 50
