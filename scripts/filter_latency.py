#!/usr/bin/env python
"""Latency of one host-memory filter call at small n (the single-state filter() of the host classes), per config,
with the small-batch in-place path on and off (ASIF_B200_SMALL_INPLACE)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import asif_b200 as ab  # noqa: E402
import conftest as cf  # noqa: E402

engines = (
    ("C1 explicit", ab.Engine(ab.FILTER_EXPLICIT, ab.MODEL_DOUBLE_INTEGRATOR), cf.c1_inputs(1024)),
    ("C2 TB npBT=101", ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS)), cf.c2_inputs(1024)),
    ("C5 segway npBT=316", ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_SEGWAY, **cf.tb_engine_kwargs(cf.SEGWAY_TB_OPTS)), cf.c5_inputs(1024)),
)
for name, eng, (x, ud) in engines:
    for n in (1, 32, 1024):
        xs, us = np.ascontiguousarray(x[:n]), np.ascontiguousarray(ud[:n])
        out = {}
        for mode in ("0", "1"):
            os.environ["ASIF_B200_SMALL_INPLACE"] = mode
            u, rl, rc = np.empty((n, eng.nu)), np.empty((n, eng.n_relax)), np.empty(n, dtype=np.int32)
            for _ in range(50):
                eng.filter_batch_into(n, xs, us, u, rl, rc)
            reps = 500
            t0 = time.perf_counter()
            for _ in range(reps):
                eng.filter_batch_into(n, xs, us, u, rl, rc)
            out[mode] = (1e6 * (time.perf_counter() - t0) / reps, u.copy(), rl.copy(), rc.copy())
        same = all(np.array_equal(a, b) for a, b in zip(out["0"][1:], out["1"][1:]))
        print("%-20s n=%-5d copies %.1f us   in place %.1f us   same bits %s" % (name, n, out["0"][0], out["1"][0], same), flush=True)
