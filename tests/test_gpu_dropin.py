"""The drop-in claim of SURVEY 8f rank 1, shown rather than asserted: the REFERENCE'S OWN filter classes (unmodified
/root/reference/src compiled where they lie) constructed on ASIF::QPWrapperB200 through the abstract QP interface
(oracle/_ref/libasif_ref_b200.so, oracle/ref_build/qp_select_shim.h) and driven state by state as the example mains do,
against the same classes on the OSQP stand-in (eps 1e-8, polish).  Plus initialize -> updateOptions -> filter of the host
mirror class against the reference class (src/asif_implicit_tb.cpp:365-405: npBT recomputed WITHOUT backTrajExtend, :377;
satSharpness clamped with codes 2 / 3).  B200 box."""
import os
import struct
import subprocess
import time

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


CASES = [
    (1, cf.C1_OPTS, cf.c1_inputs, 4000),
    (2, cf.C2_TB_OPTS, cf.c2_inputs, 4000),
    (3, cf.C3A_SHORT_OPTS, cf.c3a_inputs, 2000),
    (6, cf.SEGWAY_TB_OPTS, cf.c5_inputs, 1500),
]


@pytest.mark.parametrize("cfg,opts,gen,n", CASES)
def test_reference_classes_run_on_qpwrapper_b200(ab, refb, oracle, cfg, opts, gen, n):
    x, ud = gen(n, seed=cf.SEED + 900 + cfg)
    refb.select_backend(0)
    f_osqp = refb.create(cfg, opts)
    refb.select_backend(1)
    f_b200 = refb.create(cfg, opts)
    refb.select_backend(0)
    u0, relax0, rc0, _, st0, _ = f_osqp.filter_batch_ex(x, ud)
    t0 = time.perf_counter()
    u, relax, rc = f_b200.filter_batch(x, ud)  # n single-state filter() calls of the reference class, QP on the GPU
    per_call = (time.perf_counter() - t0) / n
    assert np.all(rc > -100), "engine error %s passed through a reference class" % rc[rc <= -100][:3]
    uo, relaxo, rco = oracle.filter_batch(cfg, x, ud, opts)
    D = cf.disagree((u, relax, rc), (u0, relax0, rc0))
    standin_off = cf.disagree((u0, relax0, rc0), (uo, relaxo, rco))
    b200_ok = ~cf.disagree((u, relax, rc), (uo, relaxo, rco))
    print("cfg %d: %d single-state filter() calls of the reference class on QPWrapperB200, %.1f us per call; rc %s; "
          "%d disagree with the OSQP stand-in build, %d of them are the stand-in deviating from the exact optimum, %d unexplained"
          % (cfg, n, per_call * 1e6, dict(zip(*np.unique(rc, return_counts=True))), D.sum(), (D & standin_off & b200_ok).sum(),
             (D & ~(standin_off & b200_ok)).sum()))
    assert (D & ~(standin_off & b200_ok)).sum() == 0
    assert D.mean() < 0.02
    assert (~b200_ok).sum() == 0  # the reference class on the B200 backend agrees with the exact oracle on every state


def test_update_options_matches_reference_class(ab, refb, tmp_path):
    """initialize(opts1) -> filter -> updateOptions(opts2) -> filter: FilterBatchImplicitTB (host mirror, C++) against the
    reference's ASIFimplicitTB.  opts1 has backTrajExtend = 0.3 (npBT 131); updateOptions drops the extension (npBT 81 for
    the new horizon 8) and clamps satSharpness 3.0 to 2 with return code 2."""
    n = 3000
    x, ud = cf.c2_inputs(n, seed=cf.SEED + 950)
    o1 = [50.0, 10.0, 5.0, 5.0, 10.0, 0.3, 0.1, 0.01, 0.1]
    o2 = [40.0, 8.0, 6.0, 4.0, 8.0, 0.3, 0.1, 0.02, 3.0]
    inp, outp = str(tmp_path / "uo_in.bin"), str(tmp_path / "uo_out.bin")
    with open(inp, "wb") as f:
        f.write(struct.pack("<q", n))
        f.write(np.asarray(o1, dtype=np.float64).tobytes())
        f.write(np.asarray(o2, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(x).tobytes())
        f.write(np.ascontiguousarray(ud).tobytes())
    host = os.path.join(cf.ROOT, "asif_b200", "host")
    subprocess.check_call(["make", "-C", host, "-s"])
    r = subprocess.run([os.path.join(host, "host_check"), "--update-options", inp, outp], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(outp, "rb").read()
    code = struct.unpack_from("<i", raw, 0)[0]
    off = 4
    got = []
    for _ in range(2):
        u = np.frombuffer(raw, dtype=np.float64, count=n, offset=off); off += 8 * n
        rl = np.frombuffer(raw, dtype=np.float64, count=n, offset=off); off += 8 * n
        rc = np.frombuffer(raw, dtype=np.int32, count=n, offset=off); off += 4 * n
        got.append((u.reshape(-1, 1), rl.reshape(-1, 1), rc))
    refb.select_backend(0)
    f = refb.create(2, o1)
    want1 = f.filter_batch_ex(x, ud)
    code_ref = f.update_options(o2)
    want2 = f.filter_batch_ex(x, ud)
    assert code == code_ref == 2
    for k, (g, w) in enumerate(zip(got, (want1, want2))):
        D = cf.disagree(g, (w[0], w[1], w[2]))
        unp = cf.unpinned_mask(w[2], g[2], g[1], w[4])
        print("updateOptions sequence, call %d: rc %s, %d disagree, %d of them unpinned" % (
            k + 1, dict(zip(*np.unique(w[2], return_counts=True))), D.sum(), (D & unp).sum()))
        assert (D & ~unp).sum() == 0 and D.mean() < 0.02
    # the two calls really ran different problems (horizon, weights, saturation changed)
    assert (np.abs(got[0][0] - got[1][0]) > 1e-3).mean() > 0.05
