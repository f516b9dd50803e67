"""One C5 (segway, ASIFimplicitTB, npBT 316) filter launch over 1e6 device-resident states, for ncu."""
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
x, ud = cf.c5_inputs(n)
eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_SEGWAY, **cf.tb_engine_kwargs(cf.SEGWAY_TB_OPTS))
xd = torch.from_numpy(x).cuda()
udd = torch.from_numpy(ud).cuda()
ua = torch.empty((n, 1), dtype=torch.float64, device="cuda")
rl = torch.empty((n, 1), dtype=torch.float64, device="cuda")
rc = torch.empty((n,), dtype=torch.int32, device="cuda")
for _ in range(3):
    eng.filter_batch_into(n, xd, udd, ua, rl, rc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.filter_batch_into(n, xd, udd, ua, rl, rc)
e1.record()
torch.cuda.synchronize()
print("C5 filter", n, "states", e0.elapsed_time(e1), "ms", np.bincount(rc.cpu().numpy() + 3))
