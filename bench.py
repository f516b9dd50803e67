#!/usr/bin/env python
"""bench.py -- filtered states/s of the batched safety-filter path on B200.

One "step" = one pass of the hot path (ASIFimplicitTB::filter on every state of the batch) over
one batch of synthetic input.  Workload = BASELINE.json configs[1]: DoubleIntegrator_implicit_tb,
backup horizon 100 Euler steps (npBT = 101), 10^7 states per GPU (weak scaling: every rank owns
its own 10^7-state slice, no collective on the data path).

  value      device-resident throughput: inputs already in HBM, CUDA-event time of K launches
  e2e        the same metric through the C-ABI call with HOST (pinned) buffers, H2D/D2H inside
  roofline   FP64-pipe roofline of tb_filter_kernel (algorithmic flops / event time vs the DFMA
             peak measured in the same run); roofline_hbm gives the HBM view of the same launch
  cpu_baseline  the reference's own filter() loop (oracle/_ref = reference sources + OSQP stand-in)
             on the host cores, bounded sample

  extra_keys  per-config device-resident numbers with a roofline each (C1 at 1e6 and 1e8 states, C3a, C3b, C4, C5 filter,
             C5 fleet rollout), a >= 2 s sustained leg of the headline kernel, the strong-scaling split of 1e7 states
             over the N ranks (BASELINE configs[1] as worded), the copy-only H2D+D2H ceiling of the box at this N,
             the pageable-caller leg and - with several visible GPUs in one process - the engine-group leg

`--impl reference` times that CPU loop alone (all host threads) and prints the same JSON shape.
Only the cpu_baseline leg and --impl reference touch oracle/; the GPU path is the C ABI of
asif_b200/libasif_b200.so and fails loudly if the library or the device is missing.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STATES_PER_GPU = 10_000_000
# ASIFimplicitTB options of the C2 workload: [relaxCost, relaxSafeLb, relaxTTS, relaxMinOrtho,
# backTrajHorizon, backTrajExtend, backTrajDt, backTrajMinOrtho, satSharpness]
C2_OPTS = [50.0, 10.0, 5.0, 5.0, 10.0, 0.0, 0.1, 0.01, 0.1]
NPBT = 101
SEED = 0xA51F + 2

# ---- algorithmic work per state (restated in DESIGN.md) --------------------------------
BYTES_PER_STATE = 8 * (2 + 1) + 8 * (1 + 1) + 4  # x, uDes in; uAct, relax, rc out = 44 B
F_STEP = 88.0   # flops per Euler step of the DoubleIntegrator closed-loop + sensitivity rhs
F_ASM = 400.0   # row assembly incl. TTS / orthogonality rows
M_ROWS, NV = 18 + 4, 2


def flops_per_state(frac_integrated, qp_rows_per_state):
    # dual active-set QP: one violation scan per processed row + 1, m*(2nv+1) flops each, ~60 flops per step
    f_qp = (qp_rows_per_state + 1.0) * M_ROWS * (2 * NV + 1) + qp_rows_per_state * 60.0
    return frac_integrated * ((NPBT - 1) * F_STEP + F_ASM) + f_qp


def make_inputs(n, seed):
    g = np.random.Generator(np.random.Philox(key=seed))
    return g.uniform(-1, 1, (n, 2)), g.uniform(-1, 1, (n, 1))


def bind_to_gpu_numa_node(dev_index):
    """Run this rank (and place its pinned staging buffers: first touch) on the CPU socket the GPU hangs off.
    With several ranks per box the host<->device legs of `e2e` otherwise cross the socket interconnect.
    Returns a short description for the JSON line; never fails the run."""
    try:
        import torch
        p = torch.cuda.get_device_properties(dev_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return "numa node unknown for %s" % bdf
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "numa node %d has no allowed cpus" % node
        os.sched_setaffinity(0, cpus)
        return "gpu %s -> numa node %d (%d cpus)" % (bdf, node, len(cpus))
    except Exception as e:  # sysfs layout differs, containers without the node files, ...
        return "not bound (%s)" % type(e).__name__


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.1:
                continue
            f = [s.strip() for s in line.split(",")]
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# =================================================================================================
def cpu_reference_run(n_states, threads, qp_oracle_mode):
    """Time the reference's filter() loop on n_states C2 states with `threads` host threads."""
    from oracle import pyref
    L = pyref.RefLib()
    if qp_oracle_mode:
        L.set_qp_mode()
    else:
        L.set_qp_mode(-1.0, -1, -1, -1, -1)  # reference defaults: eps 1e-3, no polish, warm start (qpwrapper_osqp.cpp:67-69)
    x, ud = make_inputs(n_states, SEED + 1000)
    per = (n_states + threads - 1) // threads
    filt = [L.create(pyref.CFG_DI_IMPLICIT_TB, C2_OPTS) for _ in range(threads)]
    times = [0.0] * threads

    def work(i):
        a, b = i * per, min(n_states, (i + 1) * per)
        if a >= b:
            return
        t = time.perf_counter()
        filt[i].filter_batch(x[a:b], ud[a:b])  # ctypes drops the GIL for the duration of the C call
        times[i] = time.perf_counter() - t

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    wall = time.perf_counter() - t0
    return wall, max(times)


def run_reference_arm(args):
    """--impl reference: the reference CPU implementation, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import pyref
    cores = os.cpu_count() or 1
    # size a step for ~2 s of wall clock
    w, _ = cpu_reference_run(200 * cores, cores, False)
    per_state = w / (200 * cores)
    n_step = int(max(cores * 200, min(2_000_000, 2.0 / per_state)))
    for _ in range(args.warmup):
        cpu_reference_run(n_step, cores, False)
    t_tot = 0.0
    for _ in range(args.steps):
        w, _ = cpu_reference_run(n_step, cores, False)
        t_tot += w
    value = n_step * args.steps / t_tot
    line = {
        "impl": "reference", "metric": "filtered_states_per_sec", "value": value, "unit": "states/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "states/s", "cores": cores, "kind": "reference",
                         "sample": "%d states per step (x ~ U[-1,1]^2, uDes ~ U[-1,1]); reference sources + OSQP "
                                   "stand-in at the reference's OSQP settings (eps 1e-3, warm start, no polish); "
                                   "threads=%d" % (n_step, cores)},
        "e2e": {"value": value, "unit": "states/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(n_gpus):
    return {"workload": "DoubleIntegrator_implicit_tb: ASIFimplicitTB, backup horizon 100 Euler steps (H 10, dt 0.1, "
                        "extend 0 -> npBT 101), x ~ U[-1,1]^2, uDes ~ U[-1,1]",
            "states_per_gpu": STATES_PER_GPU, "global_batch": STATES_PER_GPU * n_gpus, "parallelism": "dp%d (independent slices, "
            "no collective)" % n_gpus, "l2": "inputs larger than L2 (240 MB in, 200 MB out per step)"}


# ---- per-config legs (extra_keys): workloads of SURVEY 8d, algorithmic work per state as DESIGN.md section 4 states it ----
def _wl():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import conftest as cf  # the named workloads (inputs, option vectors, tables) live with the parity tests
    return cf


def qp_flops(rows_per_state, m, nv):
    """direct dual active-set method ("use that method's operation count"): one violation scan of the m rows per
    processed row + 1, (2 nv + 1) flops per row, and ~30 nv flops per active-set step"""
    return (rows_per_state + 1.0) * m * (2 * nv + 1) + rows_per_state * 30.0 * nv


def config_leg(torch, ab, dev, name, eng, x, ud, reps, flops_fn, bytes_per_state, p64, hbm_peak, note=""):
    n = x.shape[0]
    xd, udd = torch.from_numpy(x).to(dev), torch.from_numpy(ud).to(dev)
    ua = torch.empty((n, eng.nu), dtype=torch.float64, device=dev)
    rl = torch.empty((n, eng.n_relax), dtype=torch.float64, device=dev)
    rc = torch.empty((n,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    for _ in range(3):
        eng.filter_batch_into(n, xd, udd, ua, rl, rc, stream=st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        eng.filter_batch_into(n, xd, udd, ua, rl, rc, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rows = eng.last_qp_iterations() / n
    fl = flops_fn(rows)
    tf, gbs = fl * n / (ms * 1e-3) / 1e12, bytes_per_state * n / (ms * 1e-3) / 1e9
    f64, fhbm = tf / p64, gbs / hbm_peak
    hist = {int(k): int(v) for k, v in zip(*np.unique(rc.cpu().numpy(), return_counts=True))}
    out = {"config": name, "states": n, "ms": ms, "states_per_s": n / (ms * 1e-3), "launches_timed": reps,
           "rc_histogram": hist, "qp_rows_per_state": rows,
           "roofline": {"bound": "fp64" if f64 >= fhbm else "hbm", "flops_per_state": fl, "bytes_per_state": bytes_per_state,
                        "achieved_tflops": tf, "frac_fp64": f64, "achieved_gbs": gbs, "frac_hbm": fhbm,
                        "frac": max(f64, fhbm)}}
    if note:
        out["note"] = note
    del xd, udd, ua, rl, rc
    return out


def all_config_legs(torch, ab, dev, device_index, p64, hbm_peak):
    cf = _wl()
    legs = []
    mk = lambda *a, **k: ab.Engine(*a, device=device_index, **k)  # noqa: E731
    # C1: the one HBM-bound config; 1e6 states is launch/tail bound (40 us), 1e8 shows the asymptote
    e = mk(ab.FILTER_EXPLICIT, ab.MODEL_DOUBLE_INTEGRATOR, relaxLb=cf.C1_OPTS[0], relaxCost=cf.C1_OPTS[1])
    for n, reps in ((1_000_000, 50), (100_000_000, 5)):
        x, ud = cf.c1_inputs(n)
        legs.append(config_leg(torch, ab, dev, "C1 ASIF explicit / DoubleIntegrator, %.0e states" % n, e, x, ud, reps,
                               lambda r: 2 * 4 * 2 * 2 + 4 * 6 + 20.0, 44, p64, hbm_peak,
                               "closed-form 1-D QP (relax pinned): 2 npSS nx (1+nu) assembly flops + interval intersection"))
        del x, ud
    e.close()
    # C3a: (N-1) (70 + sincos) + assembly + QP; transcendental calls are not counted as flops
    x, ud = cf.c3a_inputs(1_000_000)
    e = mk(ab.FILTER_IMPLICIT, ab.MODEL_INVERTED_PENDULUM, **cf.implicit_engine_kwargs(cf.C3A_OPTS))
    legs.append(config_leg(torch, ab, dev, "C3a ASIFimplicit / InvertedPendulum npBT 5001, 1e6 states", e, x, ud, 3,
                           lambda r: 5000 * 70.0 + (2 * 10 * 4 * 4 + 2 * 41 * 2 * 2) + qp_flops(r, 41 + 6, 3), 52, p64, hbm_peak,
                           "5000 sincos calls per state are NOT credited as flops (they execute ~33 FP64 instructions each)"))
    e.close()
    x, ud = cf.c3b_inputs(1_000_000)
    o = cf.C3B_OPTS
    e = mk(ab.FILTER_ROBUST, ab.MODEL_INVERTED_PENDULUM_TABLE, relaxLb=o[0], relaxCost=o[1], dynParam=[o[2], o[3]],
           halfplanes=cf.halfplane_table())
    legs.append(config_leg(torch, ab, dev, "C3b ASIFrobust / InvertedPendulum + 100 half-planes, 1e6 states", e, x, ud, 20,
                           lambda r: 100 * 10.0 + qp_flops(r, 200 + 4, 2), 44, p64, hbm_peak,
                           "rows are recomputed from the table on every solver scan; credited once"))
    e.close()
    x, ud = cf.c4_inputs(1_000_000)
    e = mk(ab.FILTER_REALIZABLE, ab.MODEL_INVERTED_PENDULUM_KERNEL, **cf.realizable_engine_kwargs(cf.C4_OPTS))
    legs.append(config_leg(torch, ab, dev, "C4 ASIFrealizable / InvertedPendulum + 100Hz_50pt kernel, 1e6 states", e, x, ud, 20,
                           lambda r: 50 * (4 + 8 + 12.0) + 2 * 20.0 + qp_flops(r, 18 + 2 + 4, 2), 52, p64, hbm_peak,
                           "per facet: h (4), bounding-box test (8), exact segment/box test (12, credited for every facet)"))
    e.close()
    x, ud = cf.c5_inputs(1_000_000)
    e = mk(ab.FILTER_IMPLICIT_TB, ab.MODEL_SEGWAY, **cf.tb_engine_kwargs(cf.SEGWAY_TB_OPTS))
    leg = config_leg(torch, ab, dev, "C5 filter: ASIFimplicitTB / segway npBT 316, 1e6 states (one control step)", e, x, ud, 3,
                     lambda r: 315 * 580.0 + 1500.0 + qp_flops(r, 18 + 4, 2), 60, p64, hbm_peak,
                     "4 trig + tanh per Euler step not credited; lanes stop at their first hit, credited for the full horizon")
    legs.append(leg)
    # C5 proper: fleet rollout, state resident on the device
    n, steps = 100_000, 1000
    xd, udd = torch.from_numpy(x[:n].copy()).to(dev), torch.from_numpy(ud[:n].copy()).to(dev)
    ua = torch.empty((n, 1), dtype=torch.float64, device=dev)
    rc = torch.empty((n,), dtype=torch.int32, device=dev)
    xw = xd.clone()
    e.rollout_into(n, 10, 1e-3, xw, udd, ua, rc)  # warm-up on a copy
    torch.cuda.synchronize()
    st = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    hist = e.rollout_into(n, steps, 1e-3, xd, udd, ua, rc, want_hist=True, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    fl = leg["roofline"]["flops_per_state"]
    legs.append({"config": "C5 segway fleet rollout: 1e5 agents x 1000 control steps, state resident on the device", "agents": n,
                 "steps": steps, "seconds": sec, "state_steps_per_s": n * steps / sec, "rc_histogram_index_rc_plus_3": hist.tolist(),
                 "x_finite": bool(torch.isfinite(xd).all().item()),
                 "roofline": {"bound": "fp64", "flops_per_state": fl, "achieved_tflops": fl * n * steps / sec / 1e12,
                              "frac": fl * n * steps / sec / 1e12 / p64}})
    e.close()
    legs[-2]["note"] += ("; credited = the operation count of the reference's generated model expressions (SURVEY 8d); the kernel "
                         "evaluates them in collected form (literals folded per monomial, two reciprocals instead of seven)")
    legs.extend(qp_backend_legs(ab, device_index))
    return legs


def qp_backend_legs(ab, device_index):
    """The QPWrapper backend for nv > 4 (csrc/qp_admm.cuh): batches of synthetic QPs with a semi-definite Hessian (only a
    quarter of the variables carry cost, the rest are LP directions inside a box), through asif_qp_solve_batch with host arrays."""
    from asif_b200 import capi
    out = []
    for nv, nc, n, label in ((40, 60, 296, "one CTA per problem, workspace in shared memory"),
                             (400, 300, 18, "one thread-block cluster per problem (the size of ASIFrobust's LP-dual QP)")):
        g = np.random.Generator(np.random.Philox(key=1234 + nv))
        Hd = np.zeros(nv)
        Hd[:nv // 4] = g.uniform(0.5, 20.0, nv // 4)
        c = g.normal(0, 3, (n, nv))
        A = g.normal(0, 1, (n, nc, nv))
        A[g.random((n, nc, nv)) < 0.8] = 0.0
        lb, ub = -g.uniform(0.5, 3, nv), g.uniform(0.5, 3, nv)
        vstar = g.uniform(0.7 * lb, 0.7 * ub, (n, nv))
        b = np.einsum("kij,kj->ki", A, vstar) - g.exponential(0.5, (n, nc)) * (g.random((n, nc)) < 0.7)
        ab.qp_solve_batch(np.diag(Hd), c[:2], A[:2], b[:2], lb, ub, device=device_index)  # warm-up: allocations
        t0 = time.perf_counter()
        sol, st = ab.qp_solve_batch(np.diag(Hd), c, A, b, lb, ub, device=device_index)
        sec = time.perf_counter() - t0
        info = capi.qp_last_info()
        out.append({"config": "QPWrapper backend, nv = %d, nc = %d, %d problems (%s)" % (nv, nc, n, label), "problems": n,
                    "seconds": sec, "problems_per_s": n / sec, "ms_per_problem_if_serial": 1e3 * sec / n,
                    "status_histogram": {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))},
                    "first_problem": {"admm_iterations": info[0], "rho_updates": info[1], "polish": info[2], "active_rows": info[3],
                                      "us_equilibrate_factor_iterate_polish": list(info[4:8])},
                    "note": "wall clock through the C ABI with host arrays (copies included); eps 1e-8 + polish"})
    return out


def copy_ceiling(torch, dev, n):
    """H2D of one step's inputs and D2H of one step's outputs at the same time, nothing else: what the link allows"""
    hx = torch.empty(n * 3, dtype=torch.float64).pin_memory()
    ho = torch.empty(n * 5 // 2, dtype=torch.float64).pin_memory()
    dx, do = torch.empty_like(hx, device=dev), torch.empty_like(ho, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def both():
        with torch.cuda.stream(s1):
            dx.copy_(hx, non_blocking=True)
        with torch.cuda.stream(s2):
            ho.copy_(do, non_blocking=True)
    return both, (hx, ho, dx, do)


# =================================================================================================
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--states", type=int, default=STATES_PER_GPU, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_keys legs (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import asif_b200 as ab

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    n = args.states
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "single rank: not bound"

    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, device=local_rank,
                    relaxCost=C2_OPTS[0], relaxLb=C2_OPTS[1], relaxTTS=C2_OPTS[2], relaxMinOrtho=C2_OPTS[3],
                    backTrajHorizon=C2_OPTS[4], backTrajExtend=C2_OPTS[5], backTrajDt=C2_OPTS[6],
                    backTrajMinOrtho=C2_OPTS[7], satSharpness=C2_OPTS[8])

    # synthetic inputs: every rank owns a different slice (seed offset by rank)
    xh_np, uh_np = make_inputs(n, SEED + rank)
    xh = torch.from_numpy(xh_np).pin_memory()
    uh = torch.from_numpy(uh_np).pin_memory()
    ua_h = torch.empty((n, 1), dtype=torch.float64).pin_memory()
    rl_h = torch.empty((n, 1), dtype=torch.float64).pin_memory()
    rc_h = torch.empty((n,), dtype=torch.int32).pin_memory()
    xd, ud = xh.to(dev), uh.to(dev)
    ua_d = torch.empty((n, 1), dtype=torch.float64, device=dev)
    rl_d = torch.empty((n, 1), dtype=torch.float64, device=dev)
    rc_d = torch.empty((n,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- FP64 peak of this device, measured in the same run
    p64_tflops, p64_clock = ab.measure_fp64_peak(local_rank)

    # ---- device-resident: K launches, CUDA events on the launching stream
    for _ in range(args.warmup):
        eng.filter_batch_into(n, xd, ud, ua_d, rl_d, rc_d, stream=sp)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record(stream)
    for a, b in evs:
        a.record(stream)
        eng.filter_batch_into(n, xd, ud, ua_d, rl_d, rc_d, stream=sp)
        b.record(stream)
    e1.record(stream)
    barrier()
    t_wall1 = time.time()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    kern_ms = statistics.mean(a.elapsed_time(b) for a, b in evs)
    qp_rows = eng.last_qp_iterations() / n
    rc_np = rc_d.cpu().numpy()
    hist = {int(k): int(v) for k, v in zip(*np.unique(rc_np, return_counts=True))}
    frac_integrated = 1.0 - hist.get(2, 0) / n

    # ---- end to end through the C ABI with host (pinned) buffers: H2D + kernel + D2H per step
    for _ in range(max(args.warmup, 5)):  # >= 4: the "auto" host-IO policy tries staged and in-place twice each before it settles
        eng.filter_batch_into(n, xh, uh, ua_h, rl_h, rc_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.filter_batch_into(n, xh, uh, ua_h, rl_h, rc_h)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    t_wall2 = time.time()
    rc_h_np_first = rc_h.numpy().copy()
    ua_h_np_first = ua_h.numpy().copy()
    clocks = sampler.stop(t_wall0, t_wall2)
    same = bool(np.array_equal(rc_h_np_first, rc_np) and np.array_equal(ua_h_np_first, ua_d.cpu().numpy()))

    host_io_used = eng.last_host_io()
    extra = {}
    if not args.no_extra:
        # ---- (1) sustained: the headline kernel back to back for >= 2 s of device time, clocks sampled throughout
        k_sus = max(args.steps, int(2200.0 / kern_ms) + 1)
        s2 = ClockSampler(local_rank)
        s2.start()
        time.sleep(0.25)
        barrier()
        tw0 = time.time()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(k_sus):
            eng.filter_batch_into(n, xd, ud, ua_d, rl_d, rc_d, stream=sp)
        b.record(stream)
        barrier()
        tw1 = time.time()
        sus_ms = max_over_ranks(a.elapsed_time(b))
        extra["sustained"] = {"launches": k_sus, "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / k_sus,
                              "states_per_s": n * world * k_sus / (sus_ms * 1e-3), "clocks": s2.stop(tw0, tw1),
                              "vs_short_run_ms_per_step": (sus_ms / k_sus) / (ms_total / args.steps)}
        # ---- (2) strong scaling: BASELINE configs[1] as worded - 1e7 states TOTAL, sharded over the N ranks
        n_s = STATES_PER_GPU // world
        for _ in range(2):
            eng.filter_batch_into(n_s, xd[:n_s], ud[:n_s], ua_d[:n_s], rl_d[:n_s], rc_d[:n_s], stream=sp)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.steps):
            eng.filter_batch_into(n_s, xd[:n_s], ud[:n_s], ua_d[:n_s], rl_d[:n_s], rc_d[:n_s], stream=sp)
        b.record(stream)
        barrier()
        st_ms = max_over_ranks(a.elapsed_time(b))
        for _ in range(2):
            eng.filter_batch_into(n_s, xh[:n_s], uh[:n_s], ua_h[:n_s], rl_h[:n_s], rc_h[:n_s])
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.filter_batch_into(n_s, xh[:n_s], uh[:n_s], ua_h[:n_s], rl_h[:n_s], rc_h[:n_s])
        torch.cuda.synchronize()
        st_e2e = max_over_ranks(time.perf_counter() - t0)
        extra["strong_scaling"] = {"total_states": n_s * world, "states_per_rank": n_s, "scaling": "strong",
                                   "value": n_s * world * args.steps / (st_ms * 1e-3), "ms_per_step": st_ms / args.steps,
                                   "e2e_value": n_s * world * args.steps / st_e2e, "e2e_ms_per_step": 1e3 * st_e2e / args.steps,
                                   "unit": "states/s"}
        # ---- (3) copy-only ceiling of the box at this N: every rank moves one step's inputs H2D and outputs D2H at once
        both, keep = copy_ceiling(torch, dev, n)
        for _ in range(2):
            both()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            both()
        torch.cuda.synchronize()
        cp_s = max_over_ranks(time.perf_counter() - t0)
        del keep
        extra["copy_ceiling"] = {"ms_per_step": 1e3 * cp_s / args.steps, "aggregate_GB_per_s": BYTES_PER_STATE * n * world * args.steps / cp_s / 1e9,
                                 "states_per_s_if_copies_were_everything": n * world * args.steps / cp_s,
                                 "e2e_fraction_of_ceiling": (n * world * args.steps / e2e_s) / (n * world * args.steps / cp_s),
                                 "what": "H2D 24 B/state + D2H 20 B/state concurrently on two streams per rank, all ranks at once, pinned buffers"}
        # ---- (4) a caller with plain (pageable) arrays: std::vector / numpy memory, bounced through pinned staging
        ua_p, rl_p, rc_p = np.empty((n, 1)), np.empty((n, 1)), np.empty(n, dtype=np.int32)
        eng.filter_batch_into(n, xh_np, uh_np, ua_p, rl_p, rc_p)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            eng.filter_batch_into(n, xh_np, uh_np, ua_p, rl_p, rc_p)
        pg_s = max_over_ranks(time.perf_counter() - t0)
        extra["e2e_pageable_caller"] = {"value": n * world * 3 / pg_s, "unit": "states/s", "ms_per_step": 1e3 * pg_s / 3,
                                        "matches": bool(np.array_equal(rc_p, rc_np))}
        del ua_p, rl_p, rc_p
        extra["host_io_auto_policy_ms_per_1e6_states"] = eng.host_io_stats()
        # ---- (5) one process, one host batch, every visible GPU: the engine-group entry (asif_engine_group_*)
        if world == 1 and ab.device_count() > 1:
            grp = ab.EngineGroup(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, relaxCost=C2_OPTS[0], relaxLb=C2_OPTS[1],
                                 relaxTTS=C2_OPTS[2], relaxMinOrtho=C2_OPTS[3], backTrajHorizon=C2_OPTS[4], backTrajExtend=C2_OPTS[5],
                                 backTrajDt=C2_OPTS[6], backTrajMinOrtho=C2_OPTS[7], satSharpness=C2_OPTS[8])
            for _ in range(3):
                grp.filter_batch_into(n, xh, uh, ua_h, rl_h, rc_h)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                grp.filter_batch_into(n, xh, uh, ua_h, rl_h, rc_h)
            g_s = time.perf_counter() - t0
            extra["engine_group_e2e"] = {"devices": grp.size, "states": n, "value": n * args.steps / g_s, "unit": "states/s",
                                         "ms_per_step": 1e3 * g_s / args.steps, "matches_device_resident": bool(np.array_equal(rc_h.numpy(), rc_np))}
            grp.close()

    value = n * world * args.steps / (ms_total * 1e-3)
    e2e_value = n * world * args.steps / e2e_s
    fl = flops_per_state(frac_integrated, qp_rows)
    ach_tf = fl * n / (kern_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    ach_gbs = BYTES_PER_STATE * n / (kern_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("tb_filter_kernel_bytes_per_launch")
    except Exception:
        pass

    line = {
        "metric": "filtered_states_per_sec", "value": value, "unit": "states/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world),
        "e2e": {"value": e2e_value, "unit": "states/s", "h2d_bytes_per_step": 24 * n, "d2h_bytes_per_step": 20 * n,
                "ms_per_step": 1e3 * e2e_s / args.steps, "matches_device_resident": same, "host_binding": numa,
                "host_io": {0: "staged", 1: "out", 2: "inout"}.get(host_io_used, "?") + " (ASIF_B200_HOST_IO=%s)" % os.environ.get("ASIF_B200_HOST_IO", "auto")},
        "gpu_launches": args.steps,
        "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": p64_tflops, "unit": "TFLOP/s", "frac": ach_tf / p64_tflops,
                     "traffic": traffic, "kernel": "tb_filter_kernel<DoubleIntegratorTB,4,false>",
                     "kernel_ms": kern_ms, "flops_per_state": fl,
                     "peak_source": "asif_measure_fp64_peak (DFMA microbenchmark, same run; implied SM clock %.0f MHz)" % p64_clock},
        "roofline_hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                         "bytes_per_state": BYTES_PER_STATE, "peak_source": hbm_src},
        "rc_histogram": hist, "qp_rows_per_state": qp_rows, "clocks": clocks,
    }
    line["roofline"]["credited_vs_executed"] = (
        "credited %.0f flop/state (SURVEY 8d: 88 per Euler step + assembly + QP); the kernel executes ~33 FP64 instructions per "
        "Euler step with FMA contraction off and structural zeros skipped, i.e. credited/executed ~ 2.1: the hardware-side view "
        "is ncu's FP64-pipe utilisation in profiles/r02_c2_1e7_summary.md (68.4 %% busy, 70.7 %% issue slots, 23.9 of 32 lanes)" % fl)
    if traffic is not None:
        line["roofline"]["traffic_source"] = "profiles/traffic.json (one ncu --set full capture; see its 'source' field for the build)"
    if extra:
        line["extra_keys"] = extra
    if rank == 0 and world == 1 and not args.no_extra:
        try:
            line.setdefault("extra_keys", {})["configs"] = all_config_legs(torch, ab, dev, local_rank, p64_tflops, hbm_peak)
        except Exception as ex:  # a missing fixture must not lose the headline line
            line.setdefault("extra_keys", {})["configs"] = {"error": repr(ex)}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)
        try:
            cores = os.cpu_count() or 1
            w, _ = cpu_reference_run(100 * cores, cores, False)
            n_s = int(max(100 * cores, min(4_000_000, 12.0 / (w / (100 * cores)))))
            w, _ = cpu_reference_run(n_s, cores, False)
            line["cpu_baseline"] = {"value": n_s / w, "unit": "states/s", "cores": cores, "kind": "reference",
                                    "sample": "%d states of the same workload, reference sources + OSQP stand-in at the "
                                              "reference's OSQP settings (eps 1e-3, warm start, no polish), %d threads"
                                              % (n_s, cores)}
        except Exception as ex:  # the checker library is absent: say so, never substitute
            line["cpu_baseline"] = {"value": None, "unit": "states/s", "cores": 0, "kind": "reference",
                                    "sample": "unavailable: %s" % ex}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
