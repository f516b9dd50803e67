"""C3a / RB / C5 device-resident timings for kernel iteration (documentation numbers, not the judged bench)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import bench_configs as bc
import conftest as cf
import asif_b200 as ab
which = sys.argv[1:] or ["c3a"]
if "c3a" in which:
    x, ud = cf.c3a_inputs(1_000_000)
    bc.run("C3a ASIFimplicit / InvertedPendulum npBT=5001, 1e6 states",
           ab.Engine(ab.FILTER_IMPLICIT, ab.MODEL_INVERTED_PENDULUM, **cf.implicit_engine_kwargs(cf.C3A_OPTS)), x, ud, reps=3)
if "rb" in which:
    x, ud = cf.c3a_inputs(1_000_000)
    rb_opts = [50.0, 5.0, 10.0, 5.0, 0.001, 0.1, 10.0, 0.01, 0.05, 0.08]
    bc.run("RB ASIFimplicitRB / InvertedPendulum npBT=5001, hold 10 ms, 1e6 states",
           ab.Engine(ab.FILTER_IMPLICIT_RB, ab.MODEL_INVERTED_PENDULUM, **cf.rb_engine_kwargs(rb_opts)), x, ud, reps=2)
if "c5" in which:
    x, ud = cf.c5_inputs(1_000_000)
    bc.run("C5-filter ASIFimplicitTB / segway npBT=316, 1e6 states (one control step)",
           ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_SEGWAY, **cf.tb_engine_kwargs(cf.SEGWAY_TB_OPTS)), x, ud, reps=3)
if "c2" in which:
    x, ud = cf.c2_inputs(10_000_000)
    bc.run("C2 ASIFimplicitTB / DoubleIntegrator npBT=101, 1e7 states",
           ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS)), x, ud)
if "c1big" in which:
    x, ud = cf.c1_inputs(100_000_000)
    bc.run("C1 ASIF explicit / DoubleIntegrator, 1e8 states",
           ab.Engine(ab.FILTER_EXPLICIT, ab.MODEL_DOUBLE_INTEGRATOR, relaxLb=cf.C1_OPTS[0], relaxCost=cf.C1_OPTS[1]), x, ud, reps=5)
if "c5roll" in which:
    import time, torch
    n, steps = 100_000, 1000
    x, ud = cf.c5_inputs(n)
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_SEGWAY, **cf.tb_engine_kwargs(cf.SEGWAY_TB_OPTS))
    xd, udd = torch.from_numpy(x).cuda(), torch.from_numpy(ud).cuda()
    ua = torch.empty((n, 1), dtype=torch.float64, device="cuda")
    rc = torch.empty((n,), dtype=torch.int32, device="cuda")
    xw = xd.clone()
    eng.rollout_into(n, 10, 1e-3, xw, udd, ua, rc)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hist = eng.rollout_into(n, steps, 1e-3, xd, udd, ua, rc, want_hist=True)
    torch.cuda.synchronize()
    print(json.dumps({"config": "C5 rollout 1e5 x 1000", "seconds": time.perf_counter() - t0, "hist": hist.tolist()}), flush=True)
