"""Engine groups (asif_engine_group_*): one host batch cut into contiguous slices over every visible GPU, results landing
in the caller's arrays.  Runs with any device count (one GPU = a group of one); the N > 1 path is the same code per member.
SURVEY 8b "device list", 8e."""
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


def test_group_matches_single_engine(ab):
    n = 300_001  # ragged: the last slice is shorter
    x, ud = cf.c2_inputs(n, seed=41)
    kw = cf.tb_engine_kwargs(cf.C2_TB_OPTS)
    one = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **kw)
    grp = ab.EngineGroup(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **kw)
    assert grp.size == ab.device_count() and (grp.nx, grp.nu, grp.nc) == (one.nx, one.nu, one.nc)
    lo = [grp.slice(n, i) for i in range(grp.size)]
    assert lo[0][0] == 0 and lo[-1][1] == n and all(a[1] == b[0] for a, b in zip(lo, lo[1:]))
    want = one.filter_batch(x, ud, want_diag=True)
    got = grp.filter_batch(x, ud, want_diag=True)          # pageable arrays
    for p, q in zip(want, got):
        assert np.array_equal(p, q)
    # pinned arrays: same bits, and every member reports a device-addressable path
    bufs = [ab.PinnedArray(s, d) for s, d in (((n, 2), np.float64), ((n, 1), np.float64), ((n, 1), np.float64), ((n, 1), np.float64), (n, np.int32))]
    bufs[0].array[:] = x
    bufs[1].array[:] = ud
    grp.filter_batch_into(n, *[b.array for b in bufs])
    assert np.array_equal(bufs[2].array, want[0]) and np.array_equal(bufs[3].array, want[1]) and np.array_equal(bufs[4].array, want[2])
    assert all(grp.member_host_io(i) in (ab.HOST_IO_INOUT, ab.HOST_IO_STAGED, ab.HOST_IO_OUT) for i in range(grp.size))
    # the (H, c) overload through the group
    H, c = cf.custom_cost(ud, one.nv, seed=5)
    a = one.filter_batch_cost(x, c, H)
    b = grp.filter_batch_cost(x, c, H)
    for p, q in zip(a, b):
        assert np.array_equal(p, q)
    # tiny batches: fewer states than devices
    for m in (1, 2, 3):
        a = one.filter_batch(x[:m], ud[:m])
        b = grp.filter_batch(x[:m], ud[:m])
        for p, q in zip(a, b):
            assert np.array_equal(p, q)
    for b_ in bufs:
        b_.free()
    grp.close()


def test_group_rollout_and_explicit_device_list(ab):
    n, steps = 4097, 40
    x0, ud = cf.c5_inputs(n, seed=43)
    kw = cf.tb_engine_kwargs(cf.SEGWAY_TB_OPTS)
    one = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_SEGWAY, **kw)
    grp = ab.EngineGroup(ab.FILTER_IMPLICIT_TB, ab.MODEL_SEGWAY, devices=list(range(ab.device_count()))[::-1], **kw)
    a = one.rollout(x0, ud, steps, 1e-3)
    b = grp.rollout(x0, ud, steps, 1e-3)
    for p, q in zip(a, b):
        assert np.array_equal(p, q)
    assert b[3].sum() == n * steps
    with pytest.raises(ab.AsifError):
        ab.EngineGroup(ab.FILTER_IMPLICIT_TB, ab.MODEL_SEGWAY, devices=[0, 0], **kw)
    with pytest.raises(ab.AsifError):
        ab.EngineGroup(ab.FILTER_IMPLICIT_TB, ab.MODEL_SEGWAY, devices=[ab.device_count()], **kw)
