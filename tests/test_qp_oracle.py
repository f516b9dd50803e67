"""The exact-QP oracle (oracle/qp_enum.c) against an independent numpy active-set enumeration and
against analytic cases.  CPU only."""
import itertools

import numpy as np

import conftest as cf


def brute(H, c, A, b, lb, ub):
    nv = len(c)
    G = np.vstack([A, np.eye(nv), -np.eye(nv)])
    r = np.concatenate([b, lb, -np.asarray(ub)])
    H2 = 2 * np.diag(np.diag(H))
    best = None
    for k in range(nv + 1):
        for S in itertools.combinations(range(len(r)), k):
            N = G[list(S)]
            if k and np.linalg.matrix_rank(N) < k:
                continue
            K = np.block([[H2, -N.T], [N, np.zeros((k, k))]])
            try:
                sol = np.linalg.solve(K, np.concatenate([-c, r[list(S)]]))
            except np.linalg.LinAlgError:
                continue
            v, mu = sol[:nv], sol[nv:]
            if (mu >= -1e-10).all() and (G @ v - r >= -1e-9 * np.maximum(1, np.abs(G).max(axis=1))).all():
                f = v @ np.diag(np.diag(H)) @ v + c @ v
                if best is None or f < best[0]:
                    best = (f, v)
    return best


def test_random_small_qps(oracle):
    g = cf.philox(31)
    n_feas = n_inf = 0
    for trial in range(300):
        nv = int(g.integers(1, 4))
        nc = int(g.integers(1, 9))
        H = np.diag(g.uniform(0.5, 60.0, nv))
        c = g.normal(0, 5, nv)
        A = g.normal(0, 1, (nc, nv))
        A[g.random((nc, nv)) < 0.2] = 0.0
        b = g.normal(0, 1.5, nc)
        lb = -g.uniform(0.5, 3, nv)
        ub = g.uniform(0.5, 3, nv)
        st, sol = oracle.qp_solve(H, c, A, b, lb, ub)
        ref = brute(H, c, A, b, lb, ub)
        if ref is None:
            assert st == -3
            n_inf += 1
        else:
            assert st == 1
            assert np.abs(sol - ref[1]).max() < 1e-8
            n_feas += 1
    assert n_feas > 50 and n_inf > 10


def test_equalities_and_degenerate(oracle):
    # equality row: v0 + v1 = 1, min v0^2 + v1^2 -> (0.5, 0.5)
    st, sol = oracle.qp_solve(np.eye(2), [0, 0], [[1.0, 1.0]], [1.0], [-5, -5], [5, 5], be=[1])
    assert st == 1 and np.allclose(sol, [0.5, 0.5])
    # duplicated rows and a zero row
    st, sol = oracle.qp_solve(np.eye(2), [-4, 0], [[1, 0], [1, 0], [0, 0]], [1, 1, -1e20], [-5, -5], [5, 5])
    assert st == 1 and np.allclose(sol, [2, 0])
    # zero row with positive rhs: 0 >= 1 is infeasible
    st, _ = oracle.qp_solve(np.eye(2), [0, 0], [[0, 0]], [1.0], [-5, -5], [5, 5])
    assert st == -3
    # pinned variable (lb == ub), as the explicit filter's relax
    st, sol = oracle.qp_solve(np.diag([1.0, 50.0]), [-2, -500], [[1, 1]], [5.5], [-1, 5], [1, 5])
    assert st == 1 and np.allclose(sol, [1.0, 5.0]) or st == -3
    # 1e20 is a number, never active
    st, sol = oracle.qp_solve(np.diag([1.0, 50.0]), [-2, -1000], [[0, 0]], [-1e20], [-1, 10], [1, 1e20])
    assert st == 1 and np.allclose(sol, [1.0, 10.0])
