"""Phase timing of the nv > 4 QP solver through the reference's own ASIFrobust / ASIFrealizable classes on
QPWrapperB200 (oracle/_ref/libasif_ref_b200.so): per solve, the microseconds the kernel spent in equilibration,
factorisations, ADMM iterations and polish (asif_qp_last_info), and the wall time per filter() call.  GPU box only."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest as cf  # noqa: E402
from asif_b200 import capi  # noqa: E402
from oracle import pyref  # noqa: E402

L = pyref.RefLib(pyref.REF_B200_SO)
L.set_qp_mode()
for csize in sys.argv[1:] or ["default"]:
    if csize != "default":
        os.environ["ASIF_B200_QP_CLUSTER"] = csize
    for cfg, opts, gen, n in ((4, cf.C3B_OPTS, cf.c3b_inputs, 6), (5, cf.C4_OPTS, cf.c4_inputs, 6)):
        if os.environ.get("PROBE_CFG") and int(os.environ["PROBE_CFG"]) != cfg:
            continue
        x, ud = gen(n, seed=cf.SEED + 77)
        L.select_backend(1)
        f = L.create(cfg, opts)
        f.filter_batch(x[:1], ud[:1])  # warm-up: allocations
        for k in range(n):
            t0 = time.perf_counter()
            u, relax, rc = f.filter_batch(x[k:k + 1], ud[k:k + 1])
            dt = time.perf_counter() - t0
            it, nrho, pol, nact, t_sc, t_fa, t_it, t_po = capi.qp_last_info()
            print("cluster %s cfg %d: rc %d, %.2f ms per filter() | iterations %d, rho updates %d, polish %d (%d active rows) | "
                  "us: scale %d, factor %d (%d factorisations), iterate %d (%.1f us each), polish %d"
                  % (csize, cfg, rc[0], dt * 1e3, it, nrho, pol, nact, t_sc, t_fa, nrho + 1, t_it, t_it / max(it, 1), t_po))
