#!/usr/bin/env python
"""Parity run of SURVEY 8d: for every named config a fixed sub-sample goes through the CUDA path (C ABI) and through the
CPU oracle restatement; writes one JSON record per config with the counts the parity bar is stated in
(|du| <= 1e-6 + 1e-5|u|, identical rc, rows / TTS / BTorthoBS / hSafetyNow within 1e-9) and reports flips separately.
Runs on the GPU box:  python scripts/parity_report.py [states_per_config] > profiles/rNN_parity_report.jsonl"""
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest as cf  # noqa: E402
import asif_b200 as ab  # noqa: E402
from oracle import pyref  # noqa: E402  (checker only)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
O = pyref.OracleLib()
CORES = os.cpu_count() or 1


def oracle_batch(cfg, x, ud, opts):
    """the single-threaded C oracle over chunks on all host cores (ctypes releases the GIL)"""
    chunks = np.array_split(np.arange(len(x)), max(1, min(CORES, len(x) // 64)))
    with ThreadPoolExecutor(CORES) as ex:
        parts = list(ex.map(lambda idx: O.filter_batch(cfg, x[idx], ud[idx], opts, True), chunks))
    return [np.concatenate([p[i] for p in parts]) for i in range(4)]


def report(name, n, got, want, diag_cols=None, note="", crit_cols=None):
    u, relax, rc, diag = got
    u0, relax0, rc0, diag0 = want
    same_rc = rc == rc0
    ok = same_rc & (rc0 == 1)
    du = np.abs(u - u0).max(axis=1)
    tol = 1e-6 + 1e-5 * np.abs(u0).max(axis=1)
    rec = {
        "config": name, "states": int(n), "rc_histogram_oracle": {str(k): int(v) for k, v in zip(*np.unique(rc0, return_counts=True))},
        "rc_flips": int((~same_rc).sum()),
        "u_outside_tolerance": int((du[ok] > tol[ok]).sum()), "max_abs_du": float(du[ok].max()) if ok.any() else 0.0,
        "p999_abs_du": float(np.quantile(du[ok], 0.999)) if ok.any() else 0.0,
        "max_abs_drelax": float(np.abs(relax - relax0)[ok].max()) if ok.any() else 0.0,
    }
    if diag_cols is not None and ok.any():
        m = ok
        if crit_cols is not None:  # a different critical trajectory point is a discrete flip: counted, not compared row by row
            same_idx = np.all(diag[:, crit_cols] == diag0[:, crit_cols], axis=1)
            rec["critical_index_flips"] = int((ok & ~same_idx).sum())
            m = ok & same_idx
        d = np.abs(diag[m][:, diag_cols] - diag0[m][:, diag_cols]) / (1.0 + np.abs(diag0[m][:, diag_cols]))
        rec["rows_and_barriers_max_diff_rel_to_1_plus_abs"] = float(d.max())
        rec["rows_and_barriers_bit_identical"] = bool(np.array_equal(diag[m][:, diag_cols], diag0[m][:, diag_cols]))
        rec["rows_outside_1e-9"] = int((d.max(axis=1) > 1e-9).sum())
    if note:
        rec["note"] = note
    rec["git_head"] = os.environ.get("ASIF_GIT_HEAD", "unknown")  # the box has no .git: the caller passes the commit it snapshots
    print(json.dumps(rec), flush=True)


def main():
    # C1
    x, ud = cf.c1_inputs(N)
    eng = ab.Engine(ab.FILTER_EXPLICIT, ab.MODEL_DOUBLE_INTEGRATOR, relaxLb=cf.C1_OPTS[0], relaxCost=cf.C1_OPTS[1])
    report("C1 ASIF / DoubleIntegrator", N, eng.filter_batch(x, ud, want_diag=True), oracle_batch(1, x, ud, cf.C1_OPTS), slice(None))
    # C2
    x, ud = cf.c2_inputs(N)
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS))
    cols = [0, 1, 2] + list(range(4, eng.n_diag))  # hBackupEnd_ (col 3) is a previous-call diagnostic in the reference
    report("C2 ASIFimplicitTB / DoubleIntegrator npBT 101", N, eng.filter_batch(x, ud, want_diag=True),
           oracle_batch(2, x, ud, cf.C2_TB_OPTS), cols, crit_cols=slice(4, 8))
    # C3a at the example options is 5000 Euler steps per state: a smaller sample
    n3 = max(2000, N // 20)
    x, ud = cf.c3a_inputs(n3)
    eng = ab.Engine(ab.FILTER_IMPLICIT, ab.MODEL_INVERTED_PENDULUM, **cf.implicit_engine_kwargs(cf.C3A_OPTS))
    report("C3a ASIFimplicit / InvertedPendulum npBT 5001", n3, eng.filter_batch(x, ud, want_diag=True),
           oracle_batch(3, x, ud, cf.C3A_OPTS), slice(None), note="CUDA sincos vs glibc sin/cos: rows agree to tolerance, not bit for bit")
    # C3b
    x, ud = cf.c3b_inputs(N)
    o = cf.C3B_OPTS
    eng = ab.Engine(ab.FILTER_ROBUST, ab.MODEL_INVERTED_PENDULUM_TABLE, relaxLb=o[0], relaxCost=o[1], dynParam=[o[2], o[3]],
                    halfplanes=cf.halfplane_table())
    report("C3b ASIFrobust / InvertedPendulum + 100 half-planes", N, eng.filter_batch(x, ud, want_diag=True),
           oracle_batch(4, x, ud, o), slice(None))
    # C4
    x, ud = cf.c4_inputs(N)
    eng = ab.Engine(ab.FILTER_REALIZABLE, ab.MODEL_INVERTED_PENDULUM_KERNEL, **cf.realizable_engine_kwargs(cf.C4_OPTS))
    got = list(eng.filter_batch(x, ud, want_diag=True))
    want = oracle_batch(5, x, ud, cf.C4_OPTS)
    got[1] = got[1].copy()
    want[1] = want[1].copy()
    got[1][:, 0] = want[1][:, 0] = 0.0  # relax[0] of the reference is a non-unique LP-dual multiplier
    report("C4 ASIFrealizable / InvertedPendulum + 100Hz_50pt kernel", N, got, want, slice(None))
    # C5 (one control step)
    n5 = max(5000, N // 5)
    x, ud = cf.c5_inputs(n5)
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_SEGWAY, **cf.tb_engine_kwargs(cf.SEGWAY_TB_OPTS))
    cols = [0, 1, 2] + list(range(4, eng.n_diag))
    report("C5 ASIFimplicitTB / segway npBT 316", n5, eng.filter_batch(x, ud, want_diag=True), oracle_batch(6, x, ud, cf.SEGWAY_TB_OPTS),
           cols, note="CUDA sincos/tanh vs glibc, FMA contraction on: tolerance parity", crit_cols=slice(4, 8))
    # ASIFimplicitRB: pendulum callbacks (split gradients, tolerance parity) and double-integrator callbacks
    # (fused gradient, bit parity)
    nr = max(5000, N // 5)
    x, ud = cf.c3a_inputs(nr, seed=cf.SEED + 91)
    eng = ab.Engine(ab.FILTER_IMPLICIT_RB, ab.MODEL_INVERTED_PENDULUM, **cf.rb_engine_kwargs(cf.RB_IP_OPTS))
    report("RB ASIFimplicitRB / InvertedPendulum npBT 101, hold 0.2 s, x_unc (0.05, 0.08)", nr, eng.filter_batch(x, ud, want_diag=True),
           oracle_batch(7, x, ud, cf.RB_IP_OPTS), slice(None), crit_cols=slice(2, 12))
    x, ud = cf.c2_inputs(nr, seed=cf.SEED + 92)
    eng = ab.Engine(ab.FILTER_IMPLICIT_RB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.rb_engine_kwargs(cf.RB_DI_OPTS))
    report("RB ASIFimplicitRB / DoubleIntegrator npBT 151, hold 0.1 s, x_unc (0.05, 0.01)", nr, eng.filter_batch(x, ud, want_diag=True),
           oracle_batch(8, x, ud, cf.RB_DI_OPTS), slice(None), crit_cols=slice(2, 8))
    # the filter(x, H, c, ...) overloads on the headline config
    x, ud = cf.c2_inputs(N, seed=cf.SEED + 93)
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS))
    H, c = cf.custom_cost(ud, eng.nv, seed=2)
    want = O.filter_batch_cost(2, x, c, H, cf.C2_TB_OPTS, True)  # single thread: the oracle's cost override is a process-wide switch
    cols = [0, 1, 2] + list(range(4, eng.n_diag))
    report("C2 through filter(x, H, c): H = 2.5, c random", N, eng.filter_batch_cost(x, c, H, want_diag=True), want, cols,
           crit_cols=slice(4, 8))


if __name__ == "__main__":
    main()
