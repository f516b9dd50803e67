"""BASELINE-size runs (1e7 states of config 2, 1e6 of the others) checked through size-independent properties:
output invariants of the QP, determinism, slicing and permutation invariance, agreement of the host-memory and
device-memory entry points.  B200 box."""
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


def test_c2_full_size_properties(ab, oracle):
    n = 10_000_000
    x, ud = cf.c2_inputs(n, seed=cf.SEED + 200)
    o = cf.C2_TB_OPTS
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(o))
    u, relax, rc = eng.filter_batch(x, ud)
    # return codes are the reference's, nothing else
    assert set(np.unique(rc)) <= {1, 2, -1, -3}
    hist = {int(k): int(v) for k, v in zip(*np.unique(rc, return_counts=True))}
    print("C2 1e7 rc histogram", hist)
    assert 0.80 < hist[1] / n < 0.86 and 0.14 < hist[-3] / n < 0.20
    # invariants of the solution: input bounds, relaxation lower bound, pass-through inside the backup set
    assert u.min() >= -1.0 and u.max() <= 1.0
    ok = rc > 0
    assert relax[ok].min() >= o[1] - 1e-9
    inside = rc == 2
    assert np.array_equal(u[inside, 0], np.clip(ud[inside, 0], -1, 1)) and np.all(relax[inside] == o[1])
    # failure path: saturated backup controller K x
    fail = rc < 0
    assert np.array_equal(u[fail, 0], np.clip(-10.0 * x[fail, 0] + -20.0 * x[fail, 1], -1, 1))
    # determinism and slicing invariance (each state is independent; no cross-state effects of chunking / tail CTAs)
    u2, relax2, rc2 = eng.filter_batch(x, ud)
    assert np.array_equal(u, u2) and np.array_equal(relax, relax2) and np.array_equal(rc, rc2)
    for lo, hi in ((0, 1), (12345, 12345 + 1048577), (n - 777, n)):
        us, rs, cs = eng.filter_batch(x[lo:hi], ud[lo:hi])
        assert np.array_equal(us, u[lo:hi]) and np.array_equal(rs, relax[lo:hi]) and np.array_equal(cs, rc[lo:hi])
    # permutation invariance on a sample
    g = cf.philox(5)
    idx = g.permutation(n)[:500_000]
    up, rp, cp = eng.filter_batch(x[idx], ud[idx])
    assert np.array_equal(up, u[idx]) and np.array_equal(cp, rc[idx])
    # a seeded sub-sample against the oracle (the oracle finishes 2e4 states in seconds)
    sub = g.permutation(n)[:20_000]
    u0, relax0, rc0 = oracle.filter_batch(2, x[sub], ud[sub], o)
    cf.assert_parity("C2 sub-sample of 1e7", (u[sub], relax[sub], rc[sub]), (u0, relax0, rc0))


def test_device_pointer_entry_matches_host_entry(ab):
    import torch
    n = 300_001
    x, ud = cf.c2_inputs(n, seed=3)
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS))
    u, relax, rc = eng.filter_batch(x, ud)
    dev = torch.device("cuda", 0)
    xd, udd = torch.from_numpy(x).to(dev), torch.from_numpy(ud).to(dev)
    ua = torch.empty((n, 1), dtype=torch.float64, device=dev)
    rl = torch.empty((n, 1), dtype=torch.float64, device=dev)
    rcd = torch.empty((n,), dtype=torch.int32, device=dev)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        eng.filter_batch_into(n, xd, udd, ua, rl, rcd, stream=st.cuda_stream)
    st.synchronize()
    assert np.array_equal(ua.cpu().numpy(), u) and np.array_equal(rl.cpu().numpy(), relax) and np.array_equal(rcd.cpu().numpy(), rc)


def test_other_configs_full_size_invariants(ab):
    n = 1_000_000
    # C1
    x, ud = cf.c1_inputs(n)
    u, relax, rc = ab.Engine(ab.FILTER_EXPLICIT, ab.MODEL_DOUBLE_INTEGRATOR).filter_batch(x, ud)
    assert set(np.unique(rc)) <= {1, -1} and np.abs(u).max() <= 1.0 and np.all(relax[rc == 1] == 5.0)
    # C3b
    o = cf.C3B_OPTS
    x, ud = cf.c3b_inputs(n)
    u, relax, rc = ab.Engine(ab.FILTER_ROBUST, ab.MODEL_INVERTED_PENDULUM_TABLE, relaxLb=o[0], relaxCost=o[1],
                             dynParam=[o[2], o[3]], halfplanes=cf.halfplane_table()).filter_batch(x, ud)
    assert set(np.unique(rc)) <= {1, -1} and np.abs(u).max() <= 1.5 and relax[rc == 1].min() >= o[0] - 1e-9
    # C4
    x, ud = cf.c4_inputs(n)
    u, relax, rc = ab.Engine(ab.FILTER_REALIZABLE, ab.MODEL_INVERTED_PENDULUM_KERNEL,
                             **cf.realizable_engine_kwargs(cf.C4_OPTS)).filter_batch(x, ud)
    assert set(np.unique(rc)) <= {1, -1, -2} and np.abs(u).max() <= 1.5 and relax[:, 1].min() >= -1e-12
    # C5 filter call
    x, ud = cf.c5_inputs(n)
    u, relax, rc = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_SEGWAY, **cf.tb_engine_kwargs(cf.SEGWAY_TB_OPTS)).filter_batch(x, ud)
    assert set(np.unique(rc)) <= {1, 2, -1, -2, -3} and np.abs(u).max() <= 20.0
    assert (rc == -2).sum() <= 5  # iteration cap of the active-set solver: essentially never


@pytest.mark.parametrize("mode", ["staged", "out", "inout", "auto"])
def test_pinned_host_io_modes_match_pageable(ab, mode, monkeypatch):
    """ASIF_B200_HOST_IO: with pinned (device-addressable) arrays the kernels may store results straight into the
    caller's memory ("out", the default) or also read the inputs from it ("inout"); every mode returns the bits of
    the staged pageable path, over a size that spans the chunk ramp and ends in a ragged chunk."""
    import torch
    n = 4 * (1 << 19) + 12_345
    x, ud = cf.c2_inputs(n, seed=11)
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS))
    monkeypatch.setenv("ASIF_B200_HOST_IO", "staged")
    u, relax, rc, diag = eng.filter_batch(x, ud, want_diag=True)
    monkeypatch.setenv("ASIF_B200_HOST_IO", mode)
    xp, up = torch.from_numpy(x).pin_memory(), torch.from_numpy(ud).pin_memory()
    ua = torch.full((n, eng.nu), np.nan, dtype=torch.float64).pin_memory()
    rl = torch.full((n, eng.n_relax), np.nan, dtype=torch.float64).pin_memory()
    rcp = torch.full((n,), -99, dtype=torch.int32).pin_memory()
    dg = torch.full((n, eng.n_diag), np.nan, dtype=torch.float64).pin_memory()
    for with_diag in (False, True):
        ua.fill_(np.nan), rl.fill_(np.nan), rcp.fill_(-99)
        eng.filter_batch_into(n, xp, up, ua, rl, rcp, dg if with_diag else None)
        assert np.array_equal(ua.numpy(), u) and np.array_equal(rl.numpy(), relax) and np.array_equal(rcp.numpy(), rc)
        if with_diag:
            assert np.array_equal(dg.numpy(), diag, equal_nan=True)
    # a sub-range of the pinned arrays (alias arithmetic with an offset) and a mix of pinned inputs / pageable outputs
    lo, hi = 1001, 1001 + 700_001
    ua.fill_(np.nan)
    eng.filter_batch_into(hi - lo, xp[lo:hi], up[lo:hi], ua[lo:hi], rl[lo:hi], rcp[lo:hi])
    assert np.array_equal(ua.numpy()[lo:hi], u[lo:hi]) and np.isnan(ua.numpy()[:lo]).all() and np.isnan(ua.numpy()[hi:]).all()
    u2, r2, c2 = np.empty_like(u), np.empty_like(relax), np.empty_like(rc)
    eng.filter_batch_into(n, xp, up, u2, r2, c2)
    assert np.array_equal(u2, u) and np.array_equal(r2, relax) and np.array_equal(c2, rc)
    # another filter class through the same entry point
    x4, ud4 = cf.c4_inputs(300_001)
    e4 = ab.Engine(ab.FILTER_REALIZABLE, ab.MODEL_INVERTED_PENDULUM_KERNEL, **cf.realizable_engine_kwargs(cf.C4_OPTS))
    monkeypatch.setenv("ASIF_B200_HOST_IO", "staged")
    u4, r4, c4 = e4.filter_batch(x4, ud4)
    monkeypatch.setenv("ASIF_B200_HOST_IO", mode)
    a4 = torch.empty(u4.shape, dtype=torch.float64).pin_memory()
    b4 = torch.empty(r4.shape, dtype=torch.float64).pin_memory()
    d4 = torch.empty(c4.shape, dtype=torch.int32).pin_memory()
    e4.filter_batch_into(x4.shape[0], torch.from_numpy(x4).pin_memory(), torch.from_numpy(ud4).pin_memory(), a4, b4, d4)
    assert np.array_equal(a4.numpy(), u4) and np.array_equal(b4.numpy(), r4) and np.array_equal(d4.numpy(), c4)


def test_pinned_arrays_of_the_c_abi_take_the_direct_path(ab, monkeypatch):
    """asif_host_alloc memory is device-addressable: the TB filter reads and writes it in place (auto -> inout), the
    explicit filter stages it (auto -> staged), pageable numpy arrays are always staged; a range pinned with
    asif_host_register behaves like allocated memory.  Same bits every way."""
    monkeypatch.delenv("ASIF_B200_HOST_IO", raising=False)
    import ctypes as C
    L = ab.load_library()
    n = 200_003
    x, ud = cf.c2_inputs(n, seed=21)
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS))
    assert eng.last_host_io() == -1
    u, relax, rc = eng.filter_batch(x, ud)
    assert eng.last_host_io() == ab.HOST_IO_STAGED
    bufs = [ab.PinnedArray((n, 2)), ab.PinnedArray((n, 1)), ab.PinnedArray((n, 1)), ab.PinnedArray((n, 1)), ab.PinnedArray(n, np.int32)]
    xp, up, ua, rl, rcp = (b.array for b in bufs)
    xp[:], up[:] = x, ud
    eng.filter_batch_into(n, xp, up, ua, rl, rcp)
    assert eng.last_host_io() == ab.HOST_IO_INOUT
    assert np.array_equal(ua, u) and np.array_equal(rl, relax) and np.array_equal(rcp, rc)
    # pinned outputs, pageable inputs: the kernels still store in place, the inputs are copied
    ua[:], rcp[:] = np.nan, -99
    eng.filter_batch_into(n, x, ud, ua, rl, rcp)
    assert eng.last_host_io() == ab.HOST_IO_OUT
    assert np.array_equal(ua, u) and np.array_equal(rcp, rc)
    # explicit filter: staged by default even with pinned arrays
    e1 = ab.Engine(ab.FILTER_EXPLICIT, ab.MODEL_DOUBLE_INTEGRATOR)
    u1, r1, c1 = e1.filter_batch(x, ud)
    e1.filter_batch_into(n, xp, up, ua, rl, rcp)
    assert e1.last_host_io() == ab.HOST_IO_STAGED
    assert np.array_equal(ua, u1) and np.array_equal(rl, r1) and np.array_equal(rcp, c1)
    # registering a caller-owned range
    own = [np.zeros((n, 2)), np.zeros((n, 1)), np.zeros((n, 1)), np.zeros((n, 1)), np.zeros(n, dtype=np.int32)]
    own[0][:], own[1][:] = x, ud
    for a in own:
        ab.capi._check(L.asif_host_register(a.ctypes.data, a.nbytes))
    try:
        eng.filter_batch_into(n, *own)
        assert eng.last_host_io() == ab.HOST_IO_INOUT
        assert np.array_equal(own[2], u) and np.array_equal(own[3], relax) and np.array_equal(own[4], rc)
    finally:
        for a in own:
            assert L.asif_host_unregister(a.ctypes.data) == 0
    for b in bufs:
        b.free()


def test_pageable_bounce_path_matches_driver_staging(ab, monkeypatch):
    """Large pageable batches go through the slots' pinned staging buffers, filled and drained by host threads
    (csrc/host_copier.hpp); ASIF_B200_BOUNCE=0 leaves the staging to the driver.  Same bits, with and without the diag
    record, with pinned inputs and pageable outputs, and for the (x, c) overload whose second array is nv wide."""
    monkeypatch.delenv("ASIF_B200_HOST_IO", raising=False)
    n = 4 * (1 << 19) + 54_321
    x, ud = cf.c2_inputs(n, seed=31)
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS))
    monkeypatch.setenv("ASIF_B200_BOUNCE", "0")
    ref = eng.filter_batch(x, ud, want_diag=True)
    c = np.concatenate([-2.0 * ud, np.full((n, 1), -2.0 * cf.C2_TB_OPTS[0] * cf.C2_TB_OPTS[1])], axis=1)
    ref_c = eng.filter_batch_cost(x, c)
    monkeypatch.delenv("ASIF_B200_BOUNCE")
    for want_diag in (True, False):
        out = eng.filter_batch(x, ud, want_diag=want_diag)
        assert eng.last_host_io() == ab.HOST_IO_STAGED
        for a, b in zip(out, ref):
            assert np.array_equal(a, b, equal_nan=True)
    out_c = eng.filter_batch_cost(x, c)
    for a, b in zip(out_c, ref_c):
        assert np.array_equal(a, b)
    for a, b in zip(ref_c, ref[:3]):  # the (x, c) overload with updateCost's own c is filter(x, uDes)
        assert np.array_equal(a, b)
    # pinned inputs, pageable outputs
    xp, up = ab.PinnedArray((n, 2)), ab.PinnedArray((n, 1))
    xp.array[:], up.array[:] = x, ud
    u, relax, rc = np.empty((n, 1)), np.empty((n, 1)), np.empty(n, dtype=np.int32)
    eng.filter_batch_into(n, xp.array, up.array, u, relax, rc)
    assert np.array_equal(u, ref[0]) and np.array_equal(relax, ref[1]) and np.array_equal(rc, ref[2])
    xp.free(), up.free()


def test_small_batches_in_place_match_the_copy_path(ab, monkeypatch):
    """Host batches up to 64 KB (the single-state filter() calls above all) run in place on the engine's mapped scratch:
    same bits as the copy path, with and without the diag record, at the size limit and just past it."""
    monkeypatch.delenv("ASIF_B200_HOST_IO", raising=False)
    x, ud = cf.c2_inputs(3000, seed=41)
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS))
    row = 44
    for n in (1, 2, 31, 33, 1000, (64 << 10) // row, (64 << 10) // row + 1, 3000):
        for want_diag in (False, True):
            monkeypatch.setenv("ASIF_B200_SMALL_INPLACE", "0")
            ref = eng.filter_batch(x[:n], ud[:n], want_diag=want_diag)
            assert eng.last_host_io() == ab.HOST_IO_STAGED
            monkeypatch.setenv("ASIF_B200_SMALL_INPLACE", "1")
            out = eng.filter_batch(x[:n], ud[:n], want_diag=want_diag)
            small = n * (row + (8 * eng.n_diag if want_diag else 0)) <= (64 << 10)
            assert eng.last_host_io() == (ab.HOST_IO_INOUT if small else ab.HOST_IO_STAGED)
            for a, b in zip(out, ref):
                assert np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("which", ["explicit", "tb"])
def test_latency_server_matches_launch_path(which):
    """asif_engine_latency_server: host batches of <= 32 states through the resident warp return the bits of the launch
    path; larger batches, diag requests and the (H, c) overload keep the launch path while it runs; stop / start / H change."""
    import asif_b200 as ab
    if which == "explicit":
        eng = ab.Engine(ab.FILTER_EXPLICIT, ab.MODEL_DOUBLE_INTEGRATOR, relaxLb=cf.C1_OPTS[0], relaxCost=cf.C1_OPTS[1])
        x, ud = cf.c1_inputs(4000, seed=321)
    else:
        eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS))
        x, ud = cf.c2_inputs(4000, seed=322)
    want = eng.filter_batch(x, ud)
    eng.latency_server(True)
    for lo, hi in ((0, 1), (1, 2), (2, 34), (34, 66), (66, 97), (100, 131)):  # 1, 1, 32, 32, 31, 31 states
        for a in range(lo, hi, 32):
            b = min(hi, a + 32)
            got = eng.filter_batch(x[a:b], ud[a:b])
            for p, q in zip(got, want):
                assert np.array_equal(p, q[a:b]), (which, a, b)
    for k in range(200, 700):  # the example main loop's pattern: one state per call
        got = eng.filter_batch(x[k:k + 1], ud[k:k + 1])
        assert got[2][0] == want[2][k] and got[0][0, 0] == want[0][k, 0] and got[1][0, 0] == want[1][k, 0]
    big = eng.filter_batch(x, ud)  # a large batch while the server is resident: launch path, same bits
    for p, q in zip(big, want):
        assert np.array_equal(p, q)
    d1 = eng.filter_batch(x[:8], ud[:8], want_diag=True)  # diag: launch path
    assert np.array_equal(d1[0], want[0][:8])
    if which == "tb":
        H, c = cf.custom_cost(ud[:16], eng.nv, seed=9)
        a = eng.filter_batch_cost(x[:16], c, H)             # (H, c): launch path, and the server restarts with the new H
        eng.latency_server(False)
        b = eng.filter_batch_cost(x[:16], c, None)
        for p, q in zip(a, b):
            assert np.array_equal(p, q)
        eng.latency_server(True)
        s1 = eng.filter_batch(x[:16], ud[:16])              # server with H = 2.5 in force
        eng.latency_server(False)
        s2 = eng.filter_batch(x[:16], ud[:16])
        for p, q in zip(s1, s2):
            assert np.array_equal(p, q)
    eng.latency_server(False)
    eng.latency_server(False)
    eng.close()


def test_latency_server_refused_for_other_classes():
    import asif_b200 as ab
    eng = ab.Engine(ab.FILTER_IMPLICIT, ab.MODEL_INVERTED_PENDULUM, **cf.implicit_engine_kwargs(cf.C3A_SHORT_OPTS))
    with pytest.raises(ab.AsifError):
        eng.latency_server(True)
