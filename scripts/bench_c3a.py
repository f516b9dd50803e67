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
