"""The nv > 4 QPWrapper backend (csrc/qp_admm.cuh: cluster-cooperative OSQP-type solver with polish) on the B200 box.

1. The REFERENCE'S OWN ASIFrobust (402-variable LP-dual QP, src/asif_robust.cpp:21-22) and ASIFrealizable (38 variables +
   the 2-variable facet feasibility QP, src/asif_realizable.cpp:50-51) classes, unmodified, constructed on
   ASIF::QPWrapperB200 (oracle/_ref/libasif_ref_b200.so) against the same classes on the OSQP stand-in and against the
   exact oracle: same return codes, u / relax within the north-star tolerance.
2. asif_qp_solve_batch on random problems (semi-definite Hessians, equalities, infeasible ones), checked by a KKT
   certificate computed here (non-negative least squares for the multipliers, an LP for infeasibility) - no solver shared.
3. One problem per cluster of 8 CTAs against one problem per CTA on the same batch."""
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


@pytest.fixture(scope="module")
def refb():
    from oracle import pyref
    if not os.path.exists(pyref.REF_B200_SO):
        pytest.skip("oracle/_ref/libasif_ref_b200.so not built (needs /root/reference at build time)")
    L = pyref.RefLib(pyref.REF_B200_SO)
    L.set_qp_mode()
    return L


@pytest.mark.parametrize("cfg,opts,gen,n", [(5, cf.C4_OPTS, cf.c4_inputs, 400), (4, cf.C3B_OPTS, cf.c3b_inputs, 24)])
def test_reference_lp_dual_classes_run_on_qpwrapper_b200(ab, refb, oracle, cfg, opts, gen, n):
    import time
    from asif_b200 import capi
    x, ud = gen(n, seed=cf.SEED + 1900 + cfg)
    refb.select_backend(0)
    f_osqp = refb.create(cfg, opts)
    refb.select_backend(1)
    f_b200 = refb.create(cfg, opts)
    refb.select_backend(0)
    u0, relax0, rc0, _, st0, it0 = f_osqp.filter_batch_ex(x, ud)
    t0 = time.perf_counter()
    u, relax, rc = f_b200.filter_batch(x, ud)
    per_call = (time.perf_counter() - t0) / n
    assert np.all(rc > -100), "engine error %s passed through a reference class" % rc[rc <= -100][:3]
    uo, relaxo, rco = oracle.filter_batch(cfg, x, ud, opts)
    # realizable has two relax slots; the QP's is the second (the golden tests compare that one too)
    relax, relax0, relaxo = (np.asarray(r).reshape(n, -1)[:, -1:] for r in (relax, relax0, relaxo))
    D = cf.disagree((u, relax, rc), (u0, relax0, rc0))
    Do = cf.disagree((u, relax, rc), (uo, relaxo, rco))
    standin_off = cf.disagree((u0, relax0, rc0), (uo, relaxo, rco))
    print("cfg %d: %d filter() calls of the reference class on QPWrapperB200 (nv > 4 solver), %.2f ms per call; rc %s; last "
          "solve %s; OSQP stand-in iterations median %d; %d disagree with the stand-in build (%d of them: the stand-in is off the "
          "exact optimum), %d disagree with the exact oracle; max |du| vs stand-in %.2e"
          % (cfg, n, per_call * 1e3, dict(zip(*np.unique(rc, return_counts=True))), capi.qp_last_info(), int(np.median(it0)),
             D.sum(), (D & standin_off).sum(), Do.sum(), np.abs(u - u0).max()))
    assert (D & ~standin_off).sum() == 0
    assert Do.sum() == 0


def _rows(A, b, lb, ub, be):
    """all constraints as G v >= h"""
    nv = A.shape[1]
    G = [A, np.eye(nv), -np.eye(nv)]
    h = [b, lb, -ub]
    if be is not None and be.any():
        G.append(-A[be.astype(bool)])
        h.append(-b[be.astype(bool)])
    return np.vstack(G), np.concatenate(h)


def _kkt_ok(H, c, A, b, lb, ub, be, v, tol=2e-6):
    from scipy.optimize import nnls
    G, h = _rows(A, b, lb, ub, be)
    res = G @ v - h
    scale = 1.0 + np.abs(v).max()
    if res.min() < -tol * scale:
        return False, "infeasible by %.2e" % res.min()
    act = res < 1e-5 * scale
    g = 2.0 * H @ v + c
    if not act.any():
        return np.abs(g).max() < tol * (1 + np.abs(c).max()), "gradient %.2e" % np.abs(g).max()
    mu, rn = nnls(G[act].T, g, maxiter=20000)
    return rn < 10 * tol * (1.0 + np.abs(g).max()), "stationarity residual %.2e" % rn


def _random_problems(g, n, nv, nc, n_eq, psd_rank):
    """H diagonal with only psd_rank positive entries (the rest zero: LP directions, bounded by the box)"""
    Hd = np.zeros(nv)
    Hd[:psd_rank] = g.uniform(0.5, 20.0, psd_rank)
    H = np.diag(Hd)
    c = g.normal(0, 3, (n, nv))
    A = g.normal(0, 1, (n, nc, nv))
    A[g.random((n, nc, nv)) < 0.5] = 0.0
    lb = -g.uniform(0.5, 3, nv)
    ub = g.uniform(0.5, 3, nv)
    vstar = g.uniform(0.7 * lb, 0.7 * ub, (n, nv))
    b = np.einsum("kij,kj->ki", A, vstar) - g.exponential(0.5, (n, nc)) * (g.random((n, nc)) < 0.7)
    be = np.zeros(nc, dtype=np.uint8)
    be[:n_eq] = 1
    b[:, :n_eq] = np.einsum("kij,kj->ki", A[:, :n_eq], vstar)
    bad = g.random(n) < 0.15  # contradictory pair of rows: infeasible
    for k in np.nonzero(bad)[0]:
        A[k, nc - 1] = -A[k, nc - 2]
        b[k, nc - 1] = -b[k, nc - 2] + 1.0 + g.random()
    return H, c, A, b, lb, ub, be, bad


@pytest.mark.parametrize("nv,nc,n_eq,rank,n", [(6, 8, 0, 6, 60), (12, 20, 2, 5, 60), (40, 60, 6, 10, 40)])
def test_qp_batch_large_nv_kkt_certificates(ab, nv, nc, n_eq, rank, n):
    from scipy.optimize import linprog
    g = cf.philox(4200 + nv)
    H, c, A, b, lb, ub, be, bad = _random_problems(g, n, nv, nc, n_eq, rank)
    sol, st = ab.qp_solve_batch(H, c, A, b, lb, ub, be=be if n_eq else None)
    n_ok = 0
    for k in range(n):
        G, h = _rows(A[k], b[k], lb, ub, be if n_eq else None)
        lp = linprog(np.zeros(nv), A_ub=-G, b_ub=-h, bounds=[(None, None)] * nv, method="highs")
        if lp.status == 2:
            assert st[k] in (-3, 3), "problem %d is infeasible, status %d" % (k, st[k])
            assert np.isnan(sol[k]).all()
            continue
        assert st[k] == 1, "problem %d is feasible, status %d" % (k, st[k])
        ok, why = _kkt_ok(H, c[k], A[k], b[k], lb, ub, be if n_eq else None, sol[k])
        assert ok, "problem %d: %s" % (k, why)
        n_ok += 1
    print("nv %d nc %d: %d of %d problems feasible and KKT-certified, %d infeasible (status -3)" % (nv, nc, n_ok, n, n - n_ok))
    assert 0 < n_ok < n


def test_cluster_and_single_cta_teams_agree(ab, monkeypatch):
    """the same batch with one problem per 8-CTA cluster (more problems than resident clusters: the persistent stride) and
    with one problem per CTA"""
    g = cf.philox(4300)
    n, nv, nc = 45, 40, 60
    H, c, A, b, lb, ub, be, bad = _random_problems(g, n, nv, nc, 4, 12)
    monkeypatch.setenv("ASIF_B200_QP_CLUSTER", "8")
    s8, t8 = ab.qp_solve_batch(H, c, A, b, lb, ub, be=be)
    monkeypatch.setenv("ASIF_B200_QP_CLUSTER", "1")
    s1, t1 = ab.qp_solve_batch(H, c, A, b, lb, ub, be=be)
    assert np.array_equal(t8, t1)
    f = t1 == 1
    assert f.any() and (~f).any()
    assert np.abs(s8[f] - s1[f]).max() < 1e-7
