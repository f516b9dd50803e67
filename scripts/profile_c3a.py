"""One C3a (InvertedPendulum, ASIFimplicit, npBT 5001) filter launch over device-resident states, for ncu."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest as cf  # noqa: E402
import asif_b200 as ab  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
x, ud = cf.c3a_inputs(n)
eng = ab.Engine(ab.FILTER_IMPLICIT, ab.MODEL_INVERTED_PENDULUM, **cf.implicit_engine_kwargs(cf.C3A_OPTS))
xd = torch.from_numpy(x).cuda()
udd = torch.from_numpy(ud).cuda()
ua = torch.empty((n, eng.nu), dtype=torch.float64, device="cuda")
rl = torch.empty((n, eng.n_relax), dtype=torch.float64, device="cuda")
rc = torch.empty((n,), dtype=torch.int32, device="cuda")
for _ in range(2):
    eng.filter_batch_into(n, xd, udd, ua, rl, rc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.filter_batch_into(n, xd, udd, ua, rl, rc)
e1.record()
torch.cuda.synchronize()
print("C3a filter", n, "states", e0.elapsed_time(e1), "ms", np.bincount(rc.cpu().numpy() + 3))
