#!/usr/bin/env python
"""Generate the golden fixtures from the reference itself (oracle/_ref/libasif_ref.so =
unmodified reference sources + example callbacks + OSQP stand-in at the oracle setting:
polish on, eps_abs = eps_rel = 1e-8, cold start).  Run in the dev container, where
/root/reference exists:   make -C oracle/ref_build && python tests/golden/make_golden.py
The .npz files are committed; nothing at test time needs /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import conftest as cf  # noqa: E402
from oracle import pyref  # noqa: E402


def main():
    L = pyref.RefLib()
    L.set_qp_mode()  # eps 1e-8, polish, cold start, max_iter 20000, polish_refine_iter 10
    jobs = [
        ("c1_di_explicit", pyref.CFG_DI_EXPLICIT, cf.C1_OPTS, cf.c1_inputs(1500)),
        ("c2_di_implicit_tb", pyref.CFG_DI_IMPLICIT_TB, cf.C2_TB_OPTS, cf.c2_inputs(1500)),
        ("c5_segway_tb", pyref.CFG_SEGWAY_TB, cf.SEGWAY_TB_OPTS, cf.c5_inputs(400)),
    ]
    for extra in getattr(cf, "EXTRA_GOLDEN_JOBS", []):
        jobs.append(extra())
    for name, cfg, opts, (x, ud) in jobs:
        f = L.create(cfg, opts)
        u, relax, rc, diag, qp_status, qp_iters = f.filter_batch_ex(x, ud)
        out = os.path.join(HERE, name + ".npz")
        np.savez_compressed(out, cfg=cfg, opts=np.asarray(opts, dtype=np.float64), x=x, u_des=ud, u_act=u, relax=relax,
                            rc=rc, diag=diag, qp_status=qp_status, qp_iters=qp_iters)
        print(name, "n =", len(rc), "rc histogram", dict(zip(*np.unique(rc, return_counts=True))),
              "%.0f kB" % (os.path.getsize(out) / 1e3))
    # a closed-loop rollout (segway, 20 agents x 50 control steps) through the reference main-loop arithmetic
    f = L.create(pyref.CFG_SEGWAY_TB, cf.SEGWAY_TB_OPTS)
    x0, ud = cf.c5_inputs(20, seed=cf.SEED + 55)
    xe, ue, rce, hist = f.rollout(x0, ud, 50, 1e-3)
    np.savez_compressed(os.path.join(HERE, "c5_segway_rollout.npz"), opts=np.asarray(cf.SEGWAY_TB_OPTS), x0=x0, u_des=ud,
                        steps=50, dt=1e-3, x_end=xe, u_act_last=ue, rc_last=rce, rc_hist=hist)
    print("c5_segway_rollout rc histogram (index rc+3):", hist)


if __name__ == "__main__":
    main()
