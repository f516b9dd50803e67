"""The QPWrapper backend (asif_qp_solve_batch) and the C++ host layer on the B200 box."""
import os
import subprocess

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


def random_qps(g, n, nv, nc, eq_frac=0.0):
    H = np.diag(g.uniform(0.5, 60.0, nv))
    c = g.normal(0, 5, (n, nv))
    A = g.normal(0, 1, (n, nc, nv))
    A[g.random((n, nc, nv)) < 0.2] = 0.0
    lb = -g.uniform(0.5, 3, nv)
    ub = g.uniform(0.5, 3, nv)
    # ~70 % feasible by construction (rows satisfied with slack at a random point inside the bounds), rest random
    vstar = g.uniform(0.8 * lb, 0.8 * ub, (n, nv))
    b = np.einsum("kij,kj->ki", A, vstar) - g.exponential(0.5, (n, nc)) * (g.random((n, nc)) < 0.8)
    rnd = g.random(n) < 0.3
    b[rnd] = g.normal(0, 1.5, (int(rnd.sum()), nc))
    return H, c, A, b, lb, ub


@pytest.mark.parametrize("nv,nc", [(1, 3), (2, 6), (2, 18), (3, 9), (4, 12), (2, 200)])
def test_qp_batch_matches_oracle(ab, oracle, nv, nc):
    g = cf.philox(1000 + 10 * nv + nc)
    n = 300 if nc < 100 else 60
    H, c, A, b, lb, ub = random_qps(g, n, nv, nc)
    sol, st = ab.qp_solve_batch(H, c, A, b, lb, ub)
    n_feas = 0
    for k in range(n):
        st0, sol0 = oracle.qp_solve(H, c[k], A[k], b[k], lb, ub)
        assert st[k] == st0, "problem %d: status %d vs oracle %d" % (k, st[k], st0)
        if st0 == 1:
            n_feas += 1
            assert np.abs(sol[k] - sol0).max() < 1e-8
    assert 0 < n_feas < n or nc <= 3


def test_qp_batch_equalities_dense_hessian_and_per_problem_data(ab, oracle):
    g = cf.philox(77)
    n, nv, nc = 200, 3, 5
    H, c, A, b, lb, ub = random_qps(g, n, nv, nc)
    be = np.zeros(nc, dtype=np.uint8)
    be[0] = 1
    sol, st = ab.qp_solve_batch(H, c, A, b, lb, ub, be=be)
    for k in range(n):
        st0, sol0 = oracle.qp_solve(H, c[k], A[k], b[k], lb, ub, be=be)
        assert st[k] == st0
        if st0 == 1:
            assert np.abs(sol[k] - sol0).max() < 1e-8
            assert abs(A[k, 0] @ sol[k] - b[k, 0]) < 1e-8
    # dense SPD Hessian (diagonalCost = false) and per-problem bounds
    M = g.normal(0, 1, (nv, nv))
    Hd = M @ M.T + nv * np.eye(nv)
    lbs = np.tile(lb, (n, 1)) - g.uniform(0, 0.2, (n, nv))
    ubs = np.tile(ub, (n, 1)) + g.uniform(0, 0.2, (n, nv))
    sol, st = ab.qp_solve_batch(Hd, c, A, b, lbs, ubs, diagonal_cost=False)
    for k in range(0, n, 4):
        st0, sol0 = oracle.qp_solve(Hd, c[k], A[k], b[k], lbs[k], ubs[k], diagonal_cost=False)
        assert st[k] == st0
        if st0 == 1:
            assert np.abs(sol[k] - sol0).max() < 1e-8


def test_qp_batch_on_recorded_filter_problems(ab):
    """The (H, c, A, b, lb, ub) of the golden C2 states go through the generic backend and must reproduce the
    filter's own (u, relax): this is what QPWrapperB200 does behind ASIFimplicitTB when QPSOLVER::B200 is selected."""
    g = np.load(os.path.join(cf.ROOT, "tests", "golden", "c2_di_implicit_tb.npz"))
    m = g["rc"] == 1
    d = g["diag"][m]
    n = d.shape[0]
    A = d[:, 8:44].reshape(n, 2, 18).transpose(0, 2, 1)
    b = d[:, 44:62]
    H = np.diag([1.0, 50.0])
    c = np.stack([-2.0 * g["u_des"][m][:, 0], np.full(n, -2.0 * 50.0 * 10.0)], axis=1)
    sol, st = ab.qp_solve_batch(H, c, A, b, np.array([-1.0, 10.0]), np.array([1.0, 1e20]))
    unp = np.isin(g["qp_status"][m], (-2, 2, 3, 4))
    assert np.all(st[~unp] == 1)
    u0, r0 = g["u_act"][m][:, 0], g["relax"][m][:, 0]
    k = ~unp
    assert np.all(np.abs(np.clip(sol[k, 0], -1, 1) - u0[k]) <= 1e-6 + 1e-5 * np.abs(u0[k]))
    assert np.all(np.abs(sol[k, 1] - r0[k]) <= 1e-6 + 1e-5 * np.abs(r0[k]))


def test_cpp_host_layer_runs(ab):
    """asif_b200/host/host_check: QPWrapperB200 behind the abstract interface, the example main loop on the batched
    TB filter with updateOptions, and a filterBatch call - the reference's own usage pattern."""
    host = os.path.join(cf.ROOT, "asif_b200", "host")
    subprocess.check_call(["make", "-C", host, "-s"])
    r = subprocess.run([os.path.join(host, "host_check")], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "host_check ok" in r.stdout
