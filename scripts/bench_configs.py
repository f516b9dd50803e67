#!/usr/bin/env python
"""Device-resident throughput of every built config (not the judged bench; documentation numbers).
Writes one JSON object per config to stdout."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import asif_b200 as ab  # noqa: E402
import conftest as cf  # noqa: E402


def run(name, eng, x, ud, reps=5):
    n = x.shape[0]
    dev = torch.device("cuda", 0)
    xd, udd = torch.from_numpy(x).to(dev), torch.from_numpy(ud).to(dev)
    ua = torch.empty((n, eng.nu), dtype=torch.float64, device=dev)
    rl = torch.empty((n, eng.n_relax), dtype=torch.float64, device=dev)
    rc = torch.empty((n,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    for _ in range(2):
        eng.filter_batch_into(n, xd, udd, ua, rl, rc, stream=st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        eng.filter_batch_into(n, xd, udd, ua, rl, rc, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    h = {int(k): int(v) for k, v in zip(*np.unique(rc.cpu().numpy(), return_counts=True))}
    out = {"config": name, "states": n, "ms": ms, "states_per_s": n / (ms * 1e-3), "rc_histogram": h,
           "qp_rows_per_state": eng.last_qp_iterations() / n}
    print(json.dumps(out), flush=True)


def run_e2e(name, eng, x, ud, reps=5):
    """Host leg: pinned arrays through the C ABI, once per ASIF_B200_HOST_IO mode (wall clock around the calls)."""
    n = x.shape[0]
    xp, up = torch.from_numpy(x).pin_memory(), torch.from_numpy(ud).pin_memory()
    ua = torch.empty((n, eng.nu), dtype=torch.float64).pin_memory()
    rl = torch.empty((n, eng.n_relax), dtype=torch.float64).pin_memory()
    rc = torch.empty((n,), dtype=torch.int32).pin_memory()
    out = {"config": name, "states": n, "leg": "e2e (pinned host arrays, wall clock)",
           "bytes_per_state": 8 * (eng.nx + 2 * eng.nu + eng.n_relax) + 4}
    for mode in ("staged", "out", "inout"):
        os.environ["ASIF_B200_HOST_IO"] = mode
        for _ in range(2):
            eng.filter_batch_into(n, xp, up, ua, rl, rc)
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.filter_batch_into(n, xp, up, ua, rl, rc)
        out[mode + "_ms"] = 1e3 * (time.perf_counter() - t0) / reps
    os.environ.pop("ASIF_B200_HOST_IO", None)
    print(json.dumps(out), flush=True)


def main():
    global run
    if "--e2e" in sys.argv:
        run = run_e2e
    x, ud = cf.c1_inputs(1_000_000)
    run("C1 ASIF explicit / DoubleIntegrator, 1e6 states",
        ab.Engine(ab.FILTER_EXPLICIT, ab.MODEL_DOUBLE_INTEGRATOR, relaxLb=cf.C1_OPTS[0], relaxCost=cf.C1_OPTS[1]), x, ud, reps=20)
    x, ud = cf.c2_inputs(10_000_000)
    run("C2 ASIFimplicitTB / DoubleIntegrator npBT=101, 1e7 states",
        ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS)), x, ud)
    x, ud = cf.c3a_inputs(1_000_000)
    run("C3a ASIFimplicit / InvertedPendulum npBT=5001, 1e6 states",
        ab.Engine(ab.FILTER_IMPLICIT, ab.MODEL_INVERTED_PENDULUM, **cf.implicit_engine_kwargs(cf.C3A_OPTS)), x, ud, reps=2)
    # ASIFimplicitRB (no config of BASELINE.json names it): pendulum callbacks at the example's
    # horizon (5000 Euler steps), hold period 10 ms, x_unc = (0.05, 0.08); and with the learned residual switched on
    rb_opts = [50.0, 5.0, 10.0, 5.0, 0.001, 0.1, 10.0, 0.01, 0.05, 0.08]
    eng = ab.Engine(ab.FILTER_IMPLICIT_RB, ab.MODEL_INVERTED_PENDULUM, **cf.rb_engine_kwargs(rb_opts))
    run("RB ASIFimplicitRB / InvertedPendulum npBT=5001, hold 10 ms, 1e6 states", eng, x, ud, reps=2)
    eng.set_learning(cf.LEARN_DIMS, cf.learning_blob())
    run("RB + learned residual (16/8 and 12/10 hidden units), 1e6 states", eng, x, ud, reps=2)
    x, ud = cf.c3b_inputs(1_000_000)
    o = cf.C3B_OPTS
    run("C3b ASIFrobust / InvertedPendulum + 100 half-planes, 1e6 states",
        ab.Engine(ab.FILTER_ROBUST, ab.MODEL_INVERTED_PENDULUM_TABLE, relaxLb=o[0], relaxCost=o[1], dynParam=[o[2], o[3]],
                  halfplanes=cf.halfplane_table()), x, ud, reps=10)
    x, ud = cf.c4_inputs(1_000_000)
    run("C4 ASIFrealizable / InvertedPendulum + 100Hz_50pt polytope kernel, 1e6 states",
        ab.Engine(ab.FILTER_REALIZABLE, ab.MODEL_INVERTED_PENDULUM_KERNEL, **cf.realizable_engine_kwargs(cf.C4_OPTS)), x, ud, reps=10)
    x, ud = cf.c5_inputs(1_000_000)
    run("C5-filter ASIFimplicitTB / segway npBT=316, 1e6 states (one control step)",
        ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_SEGWAY, **cf.tb_engine_kwargs(cf.SEGWAY_TB_OPTS)), x, ud, reps=3)
    if "--e2e" in sys.argv:
        return
    # C5 proper: fleet rollout, 1e5 agents x 1000 control steps, state resident on the device
    n, steps = 100_000, 1000
    x, ud = cf.c5_inputs(n)
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_SEGWAY, **cf.tb_engine_kwargs(cf.SEGWAY_TB_OPTS))
    dev = torch.device("cuda", 0)
    xd, udd = torch.from_numpy(x).to(dev), torch.from_numpy(ud).to(dev)
    ua = torch.empty((n, 1), dtype=torch.float64, device=dev)
    rc = torch.empty((n,), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hist = eng.rollout_into(n, steps, 1e-3, xd, udd, ua, rc, want_hist=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"config": "C5 segway fleet rollout: 1e5 agents x 1000 control steps", "agents": n, "steps": steps,
                      "seconds": dt, "state_steps_per_s": n * steps / dt, "rc_histogram_index_rc_plus_3": hist.tolist(),
                      "x_finite": bool(torch.isfinite(xd).all().item())}), flush=True)


if __name__ == "__main__":
    main()
