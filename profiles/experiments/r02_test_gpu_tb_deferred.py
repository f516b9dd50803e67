"""The deferred-QP kernel of the nx = 2 time-to-backup filter (tb_filter_deferred_kernel: persistent warps, QPs that need
the active-set solver are solved 32 at a time from a per-warp stack) against the in-place kernel: the same bits for every
state, for ragged sizes, for a custom cost and for the run-time critical-point count.  ASIF_B200_TB_DEFER is read at
every launch (0: never, n: from n states on)."""
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


def _both(eng, x, ud, call="filter_batch"):
    import torch
    out = []
    xd, udd = torch.from_numpy(x).cuda(), torch.from_numpy(ud).cuda()
    n = x.shape[0]
    for mode in ("0", "2"):  # "2": every batch of two states or more takes the deferred kernel
        os.environ["ASIF_B200_TB_DEFER"] = mode
        ua = torch.full((n, eng.nu), np.nan, dtype=torch.float64, device="cuda")
        rl = torch.full((n, eng.n_relax), np.nan, dtype=torch.float64, device="cuda")
        rc = torch.full((n,), -99, dtype=torch.int32, device="cuda")
        eng.filter_batch_into(n, xd, udd, ua, rl, rc)
        torch.cuda.synchronize()
        out.append((ua.cpu().numpy(), rl.cpu().numpy(), rc.cpu().numpy(), eng.last_qp_iterations()))
    os.environ.pop("ASIF_B200_TB_DEFER", None)
    return out


@pytest.mark.parametrize("n", [2, 31, 32, 33, 1000, 4097, 300_007, 1_000_003])
def test_deferred_equals_in_place(ab, n):
    x, ud = cf.c2_inputs(n, seed=cf.SEED + 700 + n % 97)
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS))
    (u0, r0, c0, it0), (u1, r1, c1, it1) = _both(eng, x, ud)
    assert np.array_equal(c0, c1)
    assert np.array_equal(u0, u1) and np.array_equal(r0, r1)
    assert it0 == it1
    assert not np.any(c1 == -99) and not np.any(np.isnan(u1))


def test_deferred_runtime_point_count_and_general_saturation(ab):
    n = 200_003
    x, ud = cf.c2_inputs(n, seed=cf.SEED + 811)
    for npbtss, lb, ub in ((3, -1.0, 1.0), (8, -1.0, 1.0), (4, -0.7, 1.3), (4, -2.0, 2.0)):
        kw = cf.tb_engine_kwargs(cf.C2_TB_OPTS)
        kw.update(npBTSS=npbtss, lb=[lb], ub=[ub])
        eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **kw)
        (u0, r0, c0, it0), (u1, r1, c1, it1) = _both(eng, x, ud * 1.5)  # uDes outside the bounds as well: bound rows violated at the start
        assert np.array_equal(c0, c1) and np.array_equal(u0, u1) and np.array_equal(r0, r1) and it0 == it1
