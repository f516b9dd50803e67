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


# =================================================================================================
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--states", type=int, default=STATES_PER_GPU, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    for _ in range(2):
        eng.filter_batch_into(n, xh, uh, ua_h, rl_h, rc_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.filter_batch_into(n, xh, uh, ua_h, rl_h, rc_h)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    t_wall2 = time.time()
    clocks = sampler.stop(t_wall0, t_wall2)
    same = bool(np.array_equal(rc_h.numpy(), rc_np) and np.array_equal(ua_h.numpy(), ua_d.cpu().numpy()))

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
                "host_io": {0: "staged", 1: "out", 2: "inout"}.get(eng.last_host_io(), "?") + " (ASIF_B200_HOST_IO=%s)" % os.environ.get("ASIF_B200_HOST_IO", "auto")},
        "gpu_launches": args.steps,
        "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": p64_tflops, "unit": "TFLOP/s", "frac": ach_tf / p64_tflops,
                     "traffic": traffic, "kernel": "tb_filter_kernel<DoubleIntegratorTB,4,false>",
                     "kernel_ms": kern_ms, "flops_per_state": fl,
                     "peak_source": "asif_measure_fp64_peak (DFMA microbenchmark, same run; implied SM clock %.0f MHz)" % p64_clock},
        "roofline_hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                         "bytes_per_state": BYTES_PER_STATE, "peak_source": hbm_src},
        "rc_histogram": hist, "qp_rows_per_state": qp_rows, "clocks": clocks,
    }

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
