#!/usr/bin/env python
"""End-to-end time of a host batch from pageable (plain numpy) arrays against pinned ones, C1 and C2."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import asif_b200 as ab  # noqa: E402
import conftest as cf  # noqa: E402


def timed(eng, n, arrs, reps=5):
    for _ in range(2):
        eng.filter_batch_into(n, *arrs)
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.filter_batch_into(n, *arrs)
    return 1e3 * (time.perf_counter() - t0) / reps


for name, eng, (x, ud) in (
        ("C1 1e6", ab.Engine(ab.FILTER_EXPLICIT, ab.MODEL_DOUBLE_INTEGRATOR), cf.c1_inputs(1_000_000)),
        ("C2 1e7", ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS)), cf.c2_inputs(10_000_000))):
    n = x.shape[0]
    page = [x, ud, np.empty((n, 1)), np.empty((n, 1)), np.empty(n, dtype=np.int32)]
    pins = [ab.PinnedArray((n, 2)), ab.PinnedArray((n, 1)), ab.PinnedArray((n, 1)), ab.PinnedArray((n, 1)), ab.PinnedArray(n, np.int32)]
    pin = [p.array for p in pins]
    pin[0][:], pin[1][:] = x, ud
    t_page = timed(eng, n, page)
    io_page = eng.last_host_io()
    t_pin = timed(eng, n, pin)
    io_pin = eng.last_host_io()
    same = all(np.array_equal(a, b) for a, b in zip(page[2:], pin[2:]))
    print("%s: pageable %.3f ms (path %d), pinned %.3f ms (path %d), same bits %s" % (name, t_page, io_page, t_pin, io_pin, same), flush=True)
