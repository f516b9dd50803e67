#!/usr/bin/env python
"""Parity run of SURVEY 8d at its own sample sizes.  Runs on the GPU box:

    ASIF_GIT_HEAD=$(git rev-parse --short HEAD) python scripts/parity_report.py > profiles/rNN_parity_report.jsonl

Two legs per config, one JSON record each, every record stamped with the commit it ran on:

  "vs_oracle"    the CUDA path (C ABI) against the CPU restatement (oracle/liboracle.so, exact QP) - >= 1e5 states per
                 config, and the whole 1e6 draw for C1-C4 so that every state with a return code other than 1 is compared.
  "vs_reference" the CUDA path against the REFERENCE BUILD (oracle/_ref/libasif_ref.so: unmodified src/*.cpp + libaffa +
                 example callbacks, QP = the OSQP-algorithm stand-in at eps 1e-8 with polish).  States whose reference
                 return code comes from the ADMM loop's inexactness (iteration limit, "inaccurate" exits, an infeasibility
                 verdict on a problem whose exact optimum exists with a huge relaxation) cannot be matched by an exact
                 solver (SURVEY 8c known gap).  Nothing is masked: EVERY state is compared, and every state on which the CUDA
                 path and the reference build disagree (return code, or u / relax outside the tolerance) is then put to the
                 exact oracle as the arbiter - "reference_build_deviates_from_exact_optimum" counts the disagreements on which
                 the oracle sides with the CUDA path and against the reference build (with the stand-in's raw QP status
                 histogram of those states), "unexplained" counts the rest and must be 0.

The bar (BASELINE.json): |du| <= 1e-6 + 1e-5|u|, identical rc, relax to the same tolerance, rows / TTS / BTorthoBS /
hSafetyNow within 1e-9.  Flips are listed individually (index, both codes, the relaxation the exact optimum needs).
"""
import json
import multiprocessing as mp
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest as cf  # noqa: E402
from oracle import pyref  # noqa: E402  (checker only)

SCALE = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0  # 1.0 = the SURVEY 8d sizes; smaller for a quick look
CORES = os.cpu_count() or 1
HEAD = os.environ.get("ASIF_GIT_HEAD", "unknown")  # the box has no .git: the caller passes the commit it snapshots


def sz(n):
    return max(64, int(n * SCALE))


def emit(rec):
    rec["git_head"] = HEAD
    print(json.dumps(rec), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# reference-build leg: forked workers (libaffa's symbol counter and the stand-in's statistics are process-wide), started
# before anything touches CUDA in this process
def _ref_worker(job):
    kind, cfg, opts, a, b, extra = job
    R = pyref.RefLib()
    R.set_qp_mode()  # oracle setting: eps 1e-8, polish, cold start
    f = R.create(cfg, opts)
    if kind == "filter":
        u, relax, rc, _, st, _ = f.filter_batch_ex(a, b)
        return u, relax, rc, st
    if kind == "cost":
        u, relax, rc, _, st = f.filter_batch_cost(a, b, extra)
        return u, relax, rc, st
    if kind == "rollout":
        steps, dt = extra
        return f.rollout_log(a, b, steps, dt)
    raise ValueError(kind)


def ref_parallel(pool, kind, cfg, opts, a, b, extra=None):
    n = len(a)
    parts = np.array_split(np.arange(n), max(1, min(4 * CORES, n // 8)))
    res = pool.map(_ref_worker, [(kind, cfg, opts, a[i], b[i], extra) for i in parts])
    k = len(res[0])
    if kind == "rollout":  # (x, u, rc, hist, inexact steps per agent, per-call logs: x, u, rc, inexact)
        return [np.concatenate([r[0] for r in res]), np.concatenate([r[1] for r in res]), np.concatenate([r[2] for r in res]),
                np.sum([r[3] for r in res], axis=0)] + [np.concatenate([r[j] for r in res]) for j in range(4, 9)]
    return [np.concatenate([r[j] for r in res]) for j in range(k)]


# ---------------------------------------------------------------------------------------------------------------------
def oracle_batch(O, cfg, x, ud, opts, want_diag):
    """the single-threaded C oracle over chunks on all host cores (ctypes releases the GIL)"""
    chunks = np.array_split(np.arange(len(x)), max(1, min(4 * CORES, len(x) // 64)))
    with ThreadPoolExecutor(CORES) as ex:
        parts = list(ex.map(lambda idx: O.filter_batch(cfg, x[idx], ud[idx], opts, want_diag), chunks))
    return [np.concatenate([p[i] for p in parts]) if parts[0][i] is not None else None for i in range(len(parts[0]))]


def disagree(a, b):
    """states on which two result sets differ by the north-star bar: rc, or (rc == 1 on both) u / relax outside tolerance,
    or (same failure code) a different fallback input"""
    u, relax, rc = a
    u0, relax0, rc0 = b
    du = np.abs(u - u0).max(axis=1) > 1e-6 + 1e-5 * np.abs(u0).max(axis=1)
    dr = np.abs(relax - relax0).max(axis=1) > 1e-6 + 1e-5 * np.abs(relax0).max(axis=1)
    return (rc != rc0) | du | (dr & (rc0 == 1))


def exact_qp(A, b, c, hdiag, lb, ub):
    """The QP  min sum_i hdiag_i v_i^2 + c.v,  A v >= b,  lb <= v <= ub  solved in exact rational arithmetic by KKT
    enumeration (the inputs are doubles, i.e. rationals): the arbiter for states where the CUDA solver and the oracle's
    floating-point enumeration differ by more than the tolerance.  Returns the optimum as floats, or None (infeasible)."""
    from fractions import Fraction as F
    from itertools import combinations
    nv = len(c)
    arows = [([F(float(v)) for v in A[i]], F(float(b[i]))) for i in range(len(b))]
    brows = []
    for j in range(nv):
        e = [F(0)] * nv
        e[j] = F(1)
        brows.append((e, F(float(lb[j]))))
        if abs(ub[j]) < 1e19:  # an open bound (options_.inf = 1e20) is never active
            brows.append(([-t for t in e], F(-float(ub[j]))))
    rows = [r for r in arows if any(t != 0 for t in r[0])] + brows  # candidates for the active set
    allrows = arows + brows
    H2 = [F(2) * F(float(h)) for h in hdiag]
    cF = [F(float(v)) for v in c]

    def solve(M, r):
        n = len(r)
        M = [row[:] + [r[i]] for i, row in enumerate(M)]
        for col in range(n):
            piv = next((i for i in range(col, n) if M[i][col] != 0), None)
            if piv is None:
                return None
            M[col], M[piv] = M[piv], M[col]
            for i in range(n):
                if i != col and M[i][col] != 0:
                    f = M[i][col] / M[col][col]
                    M[i] = [a - f * bb for a, bb in zip(M[i], M[col])]
        return [M[i][n] / M[i][i] for i in range(n)]

    for k in range(nv + 1):
        for S in combinations(range(len(rows)), k):
            n = nv + k
            K = [[F(0)] * n for _ in range(n)]
            r = [F(0)] * n
            for i in range(nv):
                K[i][i] = H2[i]
                r[i] = -cF[i]
            for a, si in enumerate(S):
                for i in range(nv):
                    K[i][nv + a] = -rows[si][0][i]
                    K[nv + a][i] = rows[si][0][i]
                r[nv + a] = rows[si][1]
            sol = solve(K, r)
            if sol is None or any(m < 0 for m in sol[nv:]):
                continue
            v = sol[:nv]
            if all(sum(g * t for g, t in zip(row, v)) >= rhs for row, rhs in allrows):
                return [float(t) for t in v]
    return None


def compare(name, leg, got, want, diag=None, diag0=None, diag_cols=None, crit_cols=None, arbiter=None, qp_status=None, note="",
            relax_cols=None, exact=None, knife=None):
    u, relax, rc = got
    u0, relax0, rc0 = want
    if relax_cols is not None:
        relax, relax0 = relax[:, relax_cols], relax0[:, relax_cols]
    n = len(rc0)
    unp = np.zeros(n, bool)
    same_rc = rc == rc0
    flips = np.nonzero(~same_rc)[0]
    ok = same_rc & (rc0 == 1)
    du = np.abs(u - u0).max(axis=1)
    tol = 1e-6 + 1e-5 * np.abs(u0).max(axis=1)
    dr = np.abs(relax - relax0).max(axis=1)
    tolr = 1e-6 + 1e-5 * np.abs(relax0).max(axis=1)
    fb = same_rc & (rc0 != 1) & ~unp  # failure codes: uAct is the saturated backup controller (or untouched): compare it too
    rec = {
        "config": name, "leg": leg, "states": int(n),
        "rc_histogram_" + ("oracle" if leg == "vs_oracle" else "reference"): {str(k): int(v) for k, v in zip(*np.unique(rc0, return_counts=True))},
        "non_1_rc_states_compared": int((rc0 != 1).sum()),
        "rc_flips": int(flips.size),
        "u_outside_tolerance": int((du[ok] > tol[ok]).sum()), "max_abs_du": float(du[ok].max()) if ok.any() else 0.0,
        "p999_abs_du": float(np.quantile(du[ok], 0.999)) if ok.any() else 0.0,
        "relax_outside_tolerance": int((dr[ok] > tolr[ok]).sum()), "max_abs_drelax": float(dr[ok].max()) if ok.any() else 0.0,
        "fallback_u_outside_tolerance": int((du[fb] > tol[fb]).sum()) if fb.any() else 0,
    }
    bad_u = np.nonzero(ok & ((du > tol) | (dr > tolr)))[0]
    other = "oracle" if leg == "vs_oracle" else "reference"
    if bad_u.size:
        rec["outside_tolerance_states"] = [{"index": int(i), "u_cuda": [float(v) for v in u[i]], "u_" + other: [float(v) for v in u0[i]],
                                            "relax_cuda": [float(v) for v in np.atleast_1d(relax[i])],
                                            "relax_" + other: [float(v) for v in np.atleast_1d(relax0[i])]} for i in bad_u[:10]]
        if exact is not None:  # exact rational optimum of the QP these states pose (rows bit-identical on both sides)
            for e in rec["outside_tolerance_states"][:4]:
                v = exact(e["index"])
                if v is not None:
                    e["u_exact_rational"] = v[:u.shape[1]]
                    e["abs_err_cuda_vs_exact"] = float(np.abs(np.asarray(v[:u.shape[1]]) - u[e["index"]]).max())
                    e["abs_err_" + other + "_vs_exact"] = float(np.abs(np.asarray(v[:u.shape[1]]) - u0[e["index"]]).max())
    if arbiter is not None:  # three-way: who does the exact oracle side with on the states where CUDA and the reference build differ?
        ua, ra, rca = arbiter
        if relax_cols is not None:
            ra = ra[:, relax_cols]
        D = disagree((u, relax, rc), (u0, relax0, rc0))
        cuda_ok = ~disagree((u, relax, rc), (ua, ra, rca))
        ref_off = disagree((u0, relax0, rc0), (ua, ra, rca))
        rec["states_where_cuda_and_reference_build_disagree"] = int(D.sum())
        rec["of_which_reference_build_deviates_from_exact_optimum"] = int((D & cuda_ok & ref_off).sum())
        rec["unexplained"] = int((D & ~(cuda_ok & ref_off)).sum())
        if qp_status is not None and D.any():
            rec["reference_qp_status_of_disagreeing_states"] = {str(k): int(v) for k, v in zip(*np.unique(qp_status[D], return_counts=True))}
        if (D & ~(cuda_ok & ref_off)).any():
            rec["unexplained_indices"] = [int(i) for i in np.nonzero(D & ~(cuda_ok & ref_off))[0][:20]]
    if flips.size:
        rec["flips"] = [{"index": int(i), "rc_cuda": int(rc[i]), "rc_" + other: int(rc0[i]),
                         "relax_cuda": [float(v) for v in np.atleast_1d(relax[i])]} for i in flips[:20]]
        if knife is not None:  # is the state on a knife edge?  the ORACLE's own return code over 512 copies of it moved by 1e-15 (relative)
            for e in rec["flips"][:8]:
                e["oracle_rc_histogram_over_512_copies_moved_by_1e-15_relative"] = knife(e["index"])
    if diag is not None and ok.any():
        m = ok[:len(diag)]
        if crit_cols is not None:  # a different critical trajectory point is a discrete flip: counted, not compared row by row
            same_idx = np.all(diag[:, crit_cols] == diag0[:, crit_cols], axis=1)
            rec["critical_index_flips"] = int((m & ~same_idx).sum())
            m = m & same_idx
        rec["states_with_rows_compared"] = int(m.sum())
        d = np.abs(diag[m][:, diag_cols] - diag0[m][:, diag_cols]) / (1.0 + np.abs(diag0[m][:, diag_cols]))
        rec["rows_and_barriers_max_diff_rel_to_1_plus_abs"] = float(d.max())
        rec["rows_and_barriers_bit_identical"] = bool(np.array_equal(diag[m][:, diag_cols], diag0[m][:, diag_cols]))
        rec["rows_outside_1e-9"] = int((d.max(axis=1) > 1e-9).sum())
    if note:
        rec["note"] = note
    emit(rec)
    return rec


def main():
    t_start = time.time()
    NDIAG = sz(100_000)  # states whose full diag record (A_, b_, TTS, ...) is compared; u / relax / rc are compared on all
    # ---- inputs -------------------------------------------------------------------------------------------------
    inp = {}
    inp["C1"] = (1, cf.C1_OPTS, *cf.c1_inputs(sz(1_000_000)))
    inp["C2"] = (2, cf.C2_TB_OPTS, *cf.c2_inputs(sz(1_000_000)))
    inp["C3a"] = (3, cf.C3A_OPTS, *cf.c3a_inputs(sz(100_000)))
    inp["C3b"] = (4, cf.C3B_OPTS, *cf.c3b_inputs(sz(1_000_000)))
    inp["C4"] = (5, cf.C4_OPTS, *cf.c4_inputs(sz(1_000_000)))
    inp["C5"] = (6, cf.SEGWAY_TB_OPTS, *cf.c5_inputs(sz(100_000)))
    inp["RB-IP"] = (7, cf.RB_IP_OPTS, *cf.c3a_inputs(sz(100_000), seed=cf.SEED + 91))
    inp["RB-DI"] = (8, cf.RB_DI_OPTS, *cf.c2_inputs(sz(100_000), seed=cf.SEED + 92))
    # ASIF_PARITY_ONLY=C3b,C1 restricts the run to those configs (a kernel that changed after the full report); "cost" and
    # "rollout" name the two extra legs
    only = {t for t in os.environ.get("ASIF_PARITY_ONLY", "").split(",") if t}
    if only:
        inp = {k: v for k, v in inp.items() if k in only}
    do_cost = do_roll = (not only) or bool(only & {"cost", "rollout"})  # the two extra legs run together
    xc, udc = cf.c2_inputs(sz(100_000), seed=cf.SEED + 93)
    Hc, cc = cf.custom_cost(udc, 2, seed=2)
    n_roll, steps_roll = sz(1000), 1000 if SCALE >= 1 else 100
    xr0, udr = cf.c5_inputs(n_roll, seed=cf.SEED + 55)

    # ---- reference build first (forked workers; this process has not touched CUDA yet) ----------------------------
    # sample sizes: what the reference build finishes in seconds on the box's cores (C3b: 0.9 s per state - the 402-variable
    # LP-dual QP through a dense ADMM - hence a few hundred states)
    ref_n = {"C1": sz(1_000_000), "C2": sz(1_000_000), "C3a": sz(100_000), "C3b": sz(320), "C4": sz(200_000), "C5": sz(100_000),
             "RB-IP": sz(100_000), "RB-DI": sz(100_000)}
    ref = {}
    have_ref = os.path.exists(pyref.REF_SO)
    if have_ref:
        with mp.get_context("fork").Pool(CORES) as pool:
            for k, (cfg, opts, x, ud) in inp.items():
                m = min(ref_n[k], len(x))
                t0 = time.time()
                ref[k] = ref_parallel(pool, "filter", cfg, opts, x[:m], ud[:m])
                sys.stderr.write("reference build %s: %d states in %.1f s\n" % (k, m, time.time() - t0))
            if do_cost:
                ref["cost"] = ref_parallel(pool, "cost", 2, cf.C2_TB_OPTS, xc, cc, Hc)
            t0 = time.time()
            if do_roll:
                ref["rollout"] = ref_parallel(pool, "rollout", 6, cf.SEGWAY_TB_OPTS, xr0, udr, (steps_roll, 1e-3))
                # yardstick: the reference build against ITSELF from initial states moved by one part in 1e13
                ref["rollout_perturbed"] = ref_parallel(pool, "rollout", 6, cf.SEGWAY_TB_OPTS, xr0 * (1.0 + 1e-13), udr, (steps_roll, 1e-3))
                sys.stderr.write("reference build rollout: %d agents x %d steps in %.1f s\n" % (n_roll, steps_roll, time.time() - t0))
    else:
        sys.stderr.write("oracle/_ref/libasif_ref.so missing: vs_reference legs skipped\n")

    import asif_b200 as ab  # noqa: E402
    O = pyref.OracleLib()

    def engine(k):
        cfg, opts = inp[k][0], inp[k][1]
        if k == "C1":
            return ab.Engine(ab.FILTER_EXPLICIT, ab.MODEL_DOUBLE_INTEGRATOR, relaxLb=opts[0], relaxCost=opts[1])
        if k == "C2":
            return ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(opts))
        if k == "C3a":
            return ab.Engine(ab.FILTER_IMPLICIT, ab.MODEL_INVERTED_PENDULUM, **cf.implicit_engine_kwargs(opts))
        if k == "C3b":
            return ab.Engine(ab.FILTER_ROBUST, ab.MODEL_INVERTED_PENDULUM_TABLE, relaxLb=opts[0], relaxCost=opts[1],
                             dynParam=[opts[2], opts[3]], halfplanes=cf.halfplane_table())
        if k == "C4":
            return ab.Engine(ab.FILTER_REALIZABLE, ab.MODEL_INVERTED_PENDULUM_KERNEL, **cf.realizable_engine_kwargs(opts))
        if k == "C5":
            return ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_SEGWAY, **cf.tb_engine_kwargs(opts))
        if k == "RB-IP":
            return ab.Engine(ab.FILTER_IMPLICIT_RB, ab.MODEL_INVERTED_PENDULUM, **cf.rb_engine_kwargs(opts))
        if k == "RB-DI":
            return ab.Engine(ab.FILTER_IMPLICIT_RB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.rb_engine_kwargs(opts))
        raise KeyError(k)

    names = {"C1": "C1 ASIF / DoubleIntegrator", "C2": "C2 ASIFimplicitTB / DoubleIntegrator npBT 101",
             "C3a": "C3a ASIFimplicit / InvertedPendulum npBT 5001 (example options)",
             "C3b": "C3b ASIFrobust / InvertedPendulum + 100 half-planes",
             "C4": "C4 ASIFrealizable / InvertedPendulum + 100Hz_50pt kernel", "C5": "C5 ASIFimplicitTB / segway npBT 316",
             "RB-IP": "RB ASIFimplicitRB / InvertedPendulum npBT 101, hold 0.2 s, x_unc (0.05, 0.08)",
             "RB-DI": "RB ASIFimplicitRB / DoubleIntegrator npBT 151, hold 0.1 s, x_unc (0.05, 0.01)"}
    notes = {"C3a": "CUDA sincos vs glibc sin/cos: rows agree to tolerance, not bit for bit",
             "C5": "CUDA sincos/tanh vs glibc, FMA contraction on: tolerance parity",
             "C4": "relax[0] of the reference is a non-unique LP-dual multiplier: relax[1] (eps) is compared"}
    for k, (cfg, opts, x, ud) in inp.items():
        eng = engine(k)
        n = len(x)
        nd = min(NDIAG, n)
        t0 = time.time()
        u, relax, rc = eng.filter_batch(x, ud)
        _, _, _, diag = eng.filter_batch(x[:nd], ud[:nd], want_diag=True)
        t1 = time.time()
        u0, relax0, rc0 = oracle_batch(O, cfg, x, ud, opts, False)[:3]
        diag0 = oracle_batch(O, cfg, x[:nd], ud[:nd], opts, True)[3]
        sys.stderr.write("%s: CUDA %.1f s, oracle %.1f s\n" % (k, t1 - t0, time.time() - t1))
        cols, crit = slice(None), None
        if k in ("C2", "C5"):  # hBackupEnd_ (col 3) is a previous-call diagnostic in the reference
            cols, crit = [0, 1, 2] + list(range(4, eng.n_diag)), slice(4, 8)
        elif k == "RB-IP":
            crit = slice(2, 12)
        elif k == "RB-DI":
            crit = slice(2, 8)
        elif k == "C3a":
            crit = slice(2, 12)
        rcols = [1] if k == "C4" else None
        exact = None
        if k == "C2":
            nc, nv = eng.nc, eng.nv

            def exact(i, diag0=diag0, ud=ud, opts=opts, nc=nc, nv=nv):
                if i >= len(diag0):
                    return None
                A = diag0[i, 8:8 + nc * nv].reshape(nv, nc).T
                return exact_qp(A, diag0[i, 8 + nc * nv:8 + nc * nv + nc], [-2.0 * ud[i, 0], -2.0 * opts[0] * opts[1]], [1.0, opts[0]],
                                [-1.0, opts[1]], [1.0, 1e20])
        def knife(i, cfg=cfg, x=x, ud=ud, opts=opts):
            g = np.random.Generator(np.random.Philox(key=i))
            xs = x[i] * (1.0 + 1e-15 * g.standard_normal((512, x.shape[1])))
            rck = O.filter_batch(cfg, xs, np.repeat(ud[i:i + 1], 512, axis=0), opts)[2]
            return {str(a): int(b) for a, b in zip(*np.unique(rck, return_counts=True))}
        compare(names[k], "vs_oracle", (u, relax, rc), (u0, relax0, rc0), diag, diag0, cols, crit, note=notes.get(k, ""), relax_cols=rcols,
                exact=exact, knife=knife)
        if k in ref:
            ur, rr, rcr, st = ref[k]
            m = len(rcr)
            compare(names[k], "vs_reference", (u[:m], relax[:m], rc[:m]), (ur, rr, rcr), arbiter=(u0[:m], relax0[:m], rc0[:m]), qp_status=st,
                    relax_cols=rcols, note="reference build = unmodified sources + OSQP-algorithm stand-in (eps 1e-8, polish, cold start)")
        eng.close()
        # ---- the reference's OWN class on ASIF::QPWrapperB200 (nv > 4: csrc/qp_admm.cuh), state by state -----------------
        if k in ("C3b", "C4") and k in ref and os.path.exists(pyref.REF_B200_SO):
            ur, rr, rcr, st = ref[k]
            mw = min(len(rcr), sz(320) if k == "C3b" else sz(5000))
            RB = pyref.RefLib(pyref.REF_B200_SO)
            RB.set_qp_mode()
            RB.select_backend(1)
            fw = RB.create(cfg, opts)
            RB.select_backend(0)
            t0 = time.time()
            uw, rw, rcw = fw.filter_batch(x[:mw], ud[:mw])
            per_call = (time.time() - t0) / mw
            rw = np.asarray(rw).reshape(mw, -1)
            nm = names[k] + ": the reference's own class on QPWrapperB200 (%d-variable QP, cluster solver), %.2f ms per filter()" % (
                402 if k == "C3b" else 38, per_call * 1e3)
            compare(nm, "vs_oracle", (uw, rw, rcw), (u0[:mw], relax0[:mw], rc0[:mw]), relax_cols=rcols,
                    note="this backend IS the OSQP algorithm (ADMM to eps 1e-8 + polish), so it shares the OSQP stand-in's deviations from the "
                         "exact optimum on degenerate states; the statement for it is the vs_reference leg")
            compare(nm, "vs_reference", (uw, rw, rcw), (ur[:mw], rr[:mw], rcr[:mw]), arbiter=(u0[:mw], relax0[:mw], rc0[:mw]),
                    qp_status=st[:mw], relax_cols=rcols,
                    note="same unmodified class, QP backend switched: OSQP-algorithm stand-in (CPU) vs QPWrapperB200 (GPU)")

    if not do_cost and not do_roll:
        sys.stderr.write("parity report: %.0f s\n" % (time.time() - t_start))
        return
    # ---- the filter(x, H, c, ...) overloads on the headline config: fresh engine, first call -------------------------
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS))
    got = eng.filter_batch_cost(xc, cc, Hc, want_diag=True)
    want = O.filter_batch_cost(2, xc, cc, Hc, cf.C2_TB_OPTS, True)  # single thread: the oracle's cost override is a process-wide switch
    cols = [0, 1, 2] + list(range(4, eng.n_diag))
    nc, nv = eng.nc, eng.nv

    def exact_c(i):
        d0 = want[3]
        A = d0[i, 8:8 + nc * nv].reshape(nv, nc).T
        return exact_qp(A, d0[i, 8 + nc * nv:8 + nc * nv + nc], cc[i], [float(Hc[0, 0]), cf.C2_TB_OPTS[0]], [-1.0, cf.C2_TB_OPTS[1]], [1.0, 1e20])
    compare("C2 through filter(x, H, c): H = 2.5, c random; fresh engine, first call", "vs_oracle", got[:3], want[:3], got[3], want[3],
            cols, slice(4, 8), exact=exact_c)
    if "cost" in ref:
        ur, rr, rcr, st = ref["cost"]
        compare("C2 through filter(x, H, c): H = 2.5, c random; fresh engine, first call", "vs_reference", got[:3], (ur, rr, rcr),
                arbiter=want[:3], qp_status=st)
    eng.close()

    # ---- C5 closed-loop rollout against the reference build's example main loop (examples/segway_implicit_tb.cpp:251-283)
    eng = engine("C5")
    x, u, rc, hist = eng.rollout(xr0, udr, steps_roll, 1e-3)
    eng.close()
    t0 = time.time()
    chunks = np.array_split(np.arange(n_roll), max(1, min(4 * CORES, n_roll // 4)))
    with ThreadPoolExecutor(CORES) as ex:
        parts = list(ex.map(lambda idx: O.rollout(6, xr0[idx], udr[idx], steps_roll, 1e-3, cf.SEGWAY_TB_OPTS), chunks))
    xo, uo, rco = (np.concatenate([p[i] for p in parts]) for i in range(3))
    histo = np.sum([p[3] for p in parts], axis=0)
    sys.stderr.write("oracle rollout: %.1f s\n" % (time.time() - t0))

    def roll_rec(leg, xe, ue, rce, hist_e, extra):
        dx = np.abs(x - xe).max(axis=1)
        du = np.abs(u - ue).max(axis=1)
        r = {"config": "C5 rollout: %d agents x %d control steps, plant Euler dt 1e-3" % (n_roll, steps_roll), "leg": leg,
             "filter_calls": int(n_roll) * int(steps_roll),
             "rc_hist_cuda[-3..2,other]": [int(v) for v in hist], "rc_hist_other": [int(v) for v in hist_e],
             "rc_hist_total_abs_diff": int(np.abs(hist - hist_e).sum()), "rc_last_flips": int((rc != rce).sum()),
             "x_end_max_abs_diff": float(dx.max()), "x_end_p99_abs_diff": float(np.quantile(dx, 0.99)),
             "x_end_median_abs_diff": float(np.median(dx)), "agents_with_x_end_within_1e-6": int((dx <= 1e-6).sum()),
             "u_last_outside_tolerance": int((du > 1e-6 + 1e-5 * np.abs(ue).max(axis=1)).sum()), "u_last_max_abs_diff": float(du.max())}
        r.update(extra)
        emit(r)
        return dx

    dxo = roll_rec("vs_oracle", xo, uo, rco, histo, {"note": "same arithmetic except the QP step (dual active set vs KKT enumeration); "
                   "closed loop feeds per-call differences of 1e-8..1e-12 back through the plant for 1000 steps"})
    if "rollout" in ref:
        xr, ur, rcr, histr, bad, xlog, ulog, rclog, stlog = ref["rollout"]
        # (1) call by call on the states the reference's closed loop visited ("teacher forcing"): per-call parity is well posed
        xs = np.ascontiguousarray(xlog.reshape(-1, xlog.shape[2]))
        uds = np.ascontiguousarray(np.repeat(udr, steps_roll, axis=0))
        eng = engine("C5")
        uc, rlc, rcc = eng.filter_batch(xs, uds)
        eng.close()
        uo2, rlo2, rco2 = oracle_batch(O, 6, xs, uds, cf.SEGWAY_TB_OPTS, False)[:3]
        zr = np.zeros_like(rlc)  # the reference loop does not log relax: compare u and rc
        compare("C5 closed loop, call by call on the %d x %d states the reference build's loop visited" % (n_roll, steps_roll), "vs_reference",
                (uc, zr, rcc), (ulog.reshape(-1, ulog.shape[2]), zr, rclog.reshape(-1)), arbiter=(uo2, zr, rco2), qp_status=stlog.reshape(-1),
                note="reference_qp_status here is 0 / 1 = exact / inexact ADMM exit of that call")
        def knife_cl(i):
            g = np.random.Generator(np.random.Philox(key=i))
            xk = xs[i] * (1.0 + 1e-15 * g.standard_normal((512, xs.shape[1])))
            rck = O.filter_batch(6, xk, np.repeat(uds[i:i + 1], 512, axis=0), cf.SEGWAY_TB_OPTS)[2]
            return {str(a): int(b) for a, b in zip(*np.unique(rck, return_counts=True))}
        compare("C5 closed loop, call by call on the %d x %d states the reference build's loop visited" % (n_roll, steps_roll), "vs_oracle",
                (uc, rlc, rcc), (uo2, rlo2, rco2), knife=knife_cl,
                note="a flip here is a state where CUDA's sin/cos/tanh and glibc's differ in the last bit and a discrete decision sits on it")
        # (2) whole trajectories
        xp = ref["rollout_perturbed"][0]
        dself = np.abs(xr - xp).max(axis=1)
        clean = bad == 0
        dxr = np.abs(x - xr).max(axis=1)
        off = dxr > 1e-6
        roll_rec("vs_reference", xr, ur, rcr, histr, {
            "agents_whose_reference_trajectory_had_an_inexact_qp_exit": int((~clean).sum()),
            "agents_off_by_more_than_1e-6": int(off.sum()),
            "of_which_reference_had_an_inexact_qp_exit": int((off & ~clean).sum()),
            "of_which_cuda_agrees_with_the_exact_oracle_rollout_to_1e-6": int((off & (dxo <= 1e-6)).sum()),
            "yardstick_reference_vs_itself_from_x0_times_1_plus_1e-13": {
                "agents_off_by_more_than_1e-6": int((dself > 1e-6).sum()), "x_end_median_abs_diff": float(np.median(dself)),
                "x_end_p99_abs_diff": float(np.quantile(dself, 0.99)), "x_end_max_abs_diff": float(dself.max())},
            "note": "Trajectory-level agreement over 1000 control steps is not a well-posed test of this closed loop: the filter's discrete "
                    "decisions (first hit index, critical points, active set) make it discontinuous in x, and the reference build run "
                    "against ITSELF from initial states moved by 1e-13 (yardstick above) separates by more than 1e-6 for as many agents as "
                    "the CUDA path does.  The well-posed statement is the call-by-call record above (every state the reference's loop "
                    "visited, compared per call)."})
    sys.stderr.write("parity report: %.0f s\n" % (time.time() - t_start))


if __name__ == "__main__":
    main()
