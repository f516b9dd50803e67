"""C1 (explicit filter) device-resident timing at several batch sizes: mean of `reps` launches between two CUDA events.
usage: python scripts/c1_sweep.py [n ...]"""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest as cf
import asif_b200 as ab
sizes = [int(float(a)) for a in sys.argv[1:]] or [1_000_000, 10_000_000, 100_000_000]
eng = ab.Engine(ab.FILTER_EXPLICIT, ab.MODEL_DOUBLE_INTEGRATOR, relaxLb=cf.C1_OPTS[0], relaxCost=cf.C1_OPTS[1])
for n in sizes:
    x, ud = cf.c1_inputs(n)
    xd, udd = torch.from_numpy(x).cuda(), torch.from_numpy(ud).cuda()
    ua = torch.empty((n, 1), dtype=torch.float64, device="cuda"); rl = torch.empty((n, 1), dtype=torch.float64, device="cuda")
    rc = torch.empty((n,), dtype=torch.int32, device="cuda")
    reps = max(5, min(200, int(2e9 / n / 10)))
    for _ in range(3): eng.filter_batch_into(n, xd, udd, ua, rl, rc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream()
    e0.record()
    for _ in range(reps): eng.filter_batch_into(n, xd, udd, ua, rl, rc, stream=st.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"n": n, "reps": reps, "ms": ms, "GB_per_s": 44.0 * n / ms / 1e6, "rc_m1": int((rc == -1).sum())}), flush=True)
