"""A few launches of one config's filter kernel over device-resident states, for ncu (scripts/ncu_configs.sh).
usage: python scripts/profile_cfg.py c1|c2|c3a|c3b|c4|c5 [n_states]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest as cf  # noqa: E402
import asif_b200 as ab  # noqa: E402

cfg = sys.argv[1]
DEF = {"c1": 10_000_000, "c2": 2_000_000, "c3a": 200_000, "c3b": 1_000_000, "c4": 1_000_000, "c5": 200_000}
n = int(sys.argv[2]) if len(sys.argv) > 2 else DEF[cfg]
if cfg == "c1":
    x, ud = cf.c1_inputs(n)
    eng = ab.Engine(ab.FILTER_EXPLICIT, ab.MODEL_DOUBLE_INTEGRATOR, relaxLb=cf.C1_OPTS[0], relaxCost=cf.C1_OPTS[1])
elif cfg == "c2":
    x, ud = cf.c2_inputs(n)
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS))
elif cfg == "c3a":
    x, ud = cf.c3a_inputs(n)
    eng = ab.Engine(ab.FILTER_IMPLICIT, ab.MODEL_INVERTED_PENDULUM, **cf.implicit_engine_kwargs(cf.C3A_OPTS))
elif cfg == "c3b":
    x, ud = cf.c3b_inputs(n)
    o = cf.C3B_OPTS
    eng = ab.Engine(ab.FILTER_ROBUST, ab.MODEL_INVERTED_PENDULUM_TABLE, relaxLb=o[0], relaxCost=o[1], dynParam=[o[2], o[3]],
                    halfplanes=cf.halfplane_table())
elif cfg == "c4":
    x, ud = cf.c4_inputs(n)
    eng = ab.Engine(ab.FILTER_REALIZABLE, ab.MODEL_INVERTED_PENDULUM_KERNEL, **cf.realizable_engine_kwargs(cf.C4_OPTS))
elif cfg == "c5":
    x, ud = cf.c5_inputs(n)
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_SEGWAY, **cf.tb_engine_kwargs(cf.SEGWAY_TB_OPTS))
else:
    raise SystemExit("unknown config " + cfg)
xd, udd = torch.from_numpy(x).cuda(), torch.from_numpy(ud).cuda()
ua = torch.empty((n, eng.nu), dtype=torch.float64, device="cuda")
rl = torch.empty((n, eng.n_relax), dtype=torch.float64, device="cuda")
rc = torch.empty((n,), dtype=torch.int32, device="cuda")
for _ in range(3):
    eng.filter_batch_into(n, xd, udd, ua, rl, rc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.filter_batch_into(n, xd, udd, ua, rl, rc)
e1.record()
torch.cuda.synchronize()
print(cfg, n, "states", e0.elapsed_time(e1), "ms", np.bincount(rc.cpu().numpy() + 3))
