#!/usr/bin/env python
"""Single-state filter() latency, GPU path beside the reference's own per-call time on the same box.

  * GPU, C++ host classes, no Python in the timed path: asif_b200/host/host_check --latency (launch path and latency server)
  * reference build (oracle/_ref/libasif_ref.so, one thread): us per filter() call at the reference's own OSQP settings
    (eps 1e-3, warm start, no polish - what an unmodified user runs) and at the oracle settings (eps 1e-8, polish, cold)
Writes one JSON object; run on the GPU box:  python scripts/filter_latency.py > profiles/rNN_filter_latency.json"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest as cf  # noqa: E402
from oracle import pyref  # noqa: E402  (reference timing only)

out = {"gpu_cpp_host_classes": [], "reference_build_one_thread_us_per_call": {}}
host = os.path.join(ROOT, "asif_b200", "host", "host_check")
r = subprocess.run([host, "--latency"], capture_output=True, text=True, timeout=300)
for line in r.stdout.splitlines():
    if line.startswith("latency "):
        name, rest = line[8:].split("median")
        f = rest.replace("us", "").split()
        out["gpu_cpp_host_classes"].append({"call": name.strip(), "median_us": float(f[0]), "p90_us": float(f[2]), "min_us": float(f[4])})
if r.returncode != 0:
    out["gpu_error"] = (r.stdout + r.stderr)[-500:]
if os.path.exists(pyref.REF_SO):
    L = pyref.RefLib()
    for cfg, opts, gen, name in ((1, cf.C1_OPTS, cf.c1_inputs, "C1 ASIF::filter"), (2, cf.C2_TB_OPTS, cf.c2_inputs, "C2 ASIFimplicitTB::filter (npBT 101)")):
        x, ud = gen(20000)
        for label, mode in (("reference OSQP settings (eps 1e-3, warm start)", (-1.0, -1, -1, -1, -1)), ("oracle settings (eps 1e-8, polish, cold)", ())):
            L.set_qp_mode(*mode)
            f = L.create(cfg, opts)
            f.filter_batch(x[:2000], ud[:2000])
            t0 = time.perf_counter()
            f.filter_batch(x, ud)  # one C call looping over the states: per-call time of the reference class itself
            out["reference_build_one_thread_us_per_call"]["%s, %s" % (name, label)] = 1e6 * (time.perf_counter() - t0) / len(x)
out["note"] = ("GPU numbers: wall clock around FilterBatch*::filter(x, uDes, uAct, relax) with one state, 2000 calls after 200 warm-up calls. "
               "Reference numbers: the unmodified classes with the OSQP-algorithm stand-in (real OSQP is not in this image).")
print(json.dumps(out, indent=1))
