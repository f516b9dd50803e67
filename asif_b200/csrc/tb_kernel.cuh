// tb_kernel.cuh -- batched ASIFimplicitTB::filter, one state per thread, everything fused:
// backup-trajectory + sensitivity integration -> streaming critical-point selection ->
// time-to-safety -> constraint rows -> exact QP -> saturation / fallback.
//
// Reference path replaced: src/asif_implicit_tb.cpp:261-363 (filter), :407-714
// (updateConstraints), :833-909 (closed-loop rhs), plus the OSQP solve behind
// src/qpwrapper_osqp.cpp:217-239.
//
// Design (B200): the per-state work is a strictly sequential FP64 recurrence of a few thousand
// operations on 6..20 doubles against 44 B of HBM traffic per state, so the kernel is bound by
// the FP64 pipe (and by the latency of its dependent chains), never by memory.  Consequences:
//   * one state per *thread*: all 32 lanes of a warp stay on the FP64 pipe (one state per warp
//     would idle >= 26 of 32 lanes for nx = 2);
//   * the trajectory is never stored: each thread keeps a running list of the NPBTSS smallest
//     min-h points seen so far (keys in registers, (x_i, Q_i) snapshots in shared memory laid
//     out [slot][element][thread], so every access is bank-conflict free whatever slot a lane
//     writes), frozen at the first point inside the backup set;
//   * constraint rows are NOT materialised: the QP's row functor recomputes the four safety rows
//     of a critical point from its snapshot each time the active-set solver scans them (about 1.6
//     scans per state).  That costs ~4 % extra flops and removes 54 doubles of shared memory per
//     thread, which is what limits occupancy - and occupancy (warps to cover the FP64 dependent-
//     issue latency) is what limits this kernel (ncu: stall "wait" dominant at 2 warps/SMSP).
#pragma once
#include "filter_common.cuh"
#include "qp_gi.cuh"

namespace asifb {

constexpr int TB_THREADS = 128;
#ifndef ASIF_TB_STEP_UNROLL
#define ASIF_TB_STEP_UNROLL 1
#endif
constexpr int TB_STEP_UNROLL = ASIF_TB_STEP_UNROLL;
#ifndef ASIF_TB_MINBLOCKS_NX4
#define ASIF_TB_MINBLOCKS_NX4 4 // 128 registers, 16 warps/SM (snapshots in global scratch); 13.2 ms vs 16.5 ms at 2 for 1e6 C5 states
#endif
#ifndef ASIF_TB_MINBLOCKS_NX2
#define ASIF_TB_MINBLOCKS_NX2 6 // 80 registers: 24 warps/SM; measured 5.41 ms vs 6.23 ms at 4 (16 warps/SM) for 1e7 C2 states
#endif

constexpr int TB_NPBTSS_RUNTIME = -8; // generic TB instantiation: any npBTSS in 1..8

// diag record: [TTS, BTorthoBS, hSafetyNow, hBackupEnd, critIdx[npBTSS], A (nc*nv col-major), b (nc)]
__host__ __device__ constexpr int tb_diag_head(const int np) { return 4 + np; }

// shared memory doubles per thread: NPBTSS critical-point snapshots + the hit point
template <class M, int NPBTSS>
__host__ __device__ constexpr int tb_smem_doubles_per_thread()
{
	return (np_capacity(NPBTSS) + 1) * (M::NX + M::NX * M::NX);
}

// Rows of the TB QP (src/asif_implicit_tb.cpp:554-674), computed on demand.
template <class M, int NPBTSS>
struct TbRows {
#ifndef ASIF_TB_SCAN_INDEX_ONLY
#define ASIF_TB_SCAN_INDEX_ONLY 1 // qp_gi.cuh index-only scan: C2 3.55 -> 3.43 ms per 1e7 states, with the point pre-filter off (below) 3.36 ms; C5 neutral
#endif
	static constexpr bool SCAN_INDEX_ONLY = ASIF_TB_SCAN_INDEX_ONLY != 0;
	static constexpr int NX = M::NX, NU = M::NU, NPSS = M::NPSS, NS = NX + NX * NX;
	static constexpr int CAP = np_capacity(NPBTSS), NV = NU + 1;
	__device__ __forceinline__ int count_np() const { return np_runtime(NPBTSS) ? np : NPBTSS; }
	__device__ __forceinline__ int nc() const { return count_np() * NPSS + 2; }
	const double *snap; // this thread's view, element stride T
	static constexpr int T = TB_THREADS;
	double f[NX], g[NX * NU]; // open-loop dynamics at the current state (:416-418)
	int kslot[CAP];
	int nkept;
	bool trivial;         // inside the backup set: A = 0, b = -inf (:716-733)
	double lgT[NU], rhsT; // time-to-safety row
	double lgO[NU], rhsO; // orthogonality row
	double lb[NV], ub[NV];
	double neg_inf;
	int np; // critical points in use, set for the run-time-count instantiations only

	// the NPSS safety rows of the critical point stored in `slot` (:567-585, :643-674)
	__device__ __forceinline__ void point_rows(const int slot, double (&n)[NPSS][NV], double (&rhs)[NPSS]) const
	{
		double xs[NS], hs[NPSS], lf[NPSS], lg[NPSS * NU];
#pragma unroll
		for (int e = 0; e < NS; e++) xs[e] = snap[(slot * NS + e) * T];
		safety_point_rows<M>(xs, f, g, hs, lf, lg);
#pragma unroll
		for (int j = 0; j < NPSS; j++) {
#pragma unroll
			for (int i = 0; i < NU; i++) n[j][i] = lg[j * NU + i];
			n[j][NU] = hs[j];
			rhs[j] = -lf[j];
		}
	}
	__device__ __forceinline__ void bound_row(const int k, double (&n)[NV], double &rhs) const
	{
		const int var = k >> 1;
		const bool upper = k & 1;
		double bnd = 0.0;
#pragma unroll
		for (int i = 0; i < NV; i++) {
			n[i] = (i == var) ? (upper ? -1.0 : 1.0) : 0.0;
			if (i == var) bnd = upper ? -ub[i] : lb[i];
		}
		rhs = bnd;
	}
	// the slots of the kept points, 4 bits each (CAP <= 8): looked up by shift and mask in the row loops, where the
	// loop index is not a compile-time constant (an array would live in local memory)
	unsigned kpack;
	__device__ __forceinline__ void pack_slots()
	{
		kpack = 0u;
#pragma unroll
		for (int t = 0; t < CAP; t++) kpack |= (unsigned)kslot[t] << (4 * t);
	}
	__device__ __forceinline__ int slot_of(const int s) const { return (int)((kpack >> (4 * s)) & 15u); }
	// every row once, in the reference's row order, then the 2*NV variable bounds
	template <class F, class FB>
	__device__ __forceinline__ void scan(F &&fn, FB &&fb) const
	{
		double v0[NV];
#pragma unroll
		for (int i = 0; i < NV; i++) v0[i] = 0.0;
		scan_at<false>(v0, fn, fb);
	}
	// PREFILTER: rows are presented at the iterate v; the four rows of a critical point are handed to fn only when
	// the smallest of their residuals is below -QP_FEAS_TOL, which is the first thing fn tests for each of them
	// (qp_gi.cuh), so the solver's decisions are unchanged while the common case costs one branch per point
	template <bool PREFILTER, class F, class FB>
	__device__ __forceinline__ void scan_at(const double (&v)[NV], F &&fn, FB &&fb) const
	{
		if (!trivial) {
			const int NC = nc();
			const int nk = nkept < count_np() ? nkept : count_np();
#pragma unroll 1
			for (int s = 0; s < nk; s++) { // not unrolled: one copy of point_rows keeps registers down
				double n[NPSS][NV], rhs[NPSS];
				point_rows(slot_of(s), n, rhs);
				bool any = true;
				if (PREFILTER) {
					double smin = 0.0;
#pragma unroll
					for (int j = 0; j < NPSS; j++) {
						double r = -rhs[j]; // same expression as the solver's residual
#pragma unroll
						for (int i = 0; i < NV; i++) r += n[j][i] * v[i];
						smin = (j == 0 || r < smin) ? r : smin;
					}
					any = smin < -QP_FEAS_TOL;
				}
				if (any) {
#pragma unroll
					for (int j = 0; j < NPSS; j++) fn(s * NPSS + j, n[j], rhs[j]);
				}
			}
#pragma unroll 1
			for (int s = nk; s < count_np(); s++) { // missing points: h = 1, Dh = 0 (:556-566)
				double n[NV];
#pragma unroll
				for (int i = 0; i < NU; i++) n[i] = 0.0;
				n[NU] = 1.0;
#pragma unroll
				for (int j = 0; j < NPSS; j++) fn(s * NPSS + j, n, -0.0);
			}
			double n[NV];
#pragma unroll
			for (int i = 0; i < NU; i++) n[i] = lgT[i];
			n[NU] = 0.0;
			fn(NC - 2, n, rhsT);
#pragma unroll
			for (int i = 0; i < NU; i++) n[i] = lgO[i];
			fn(NC - 1, n, rhsO);
		}
#pragma unroll
		for (int k = 0; k < 2 * NV; k++) fb(nc() + k, k >> 1, (k & 1) != 0, (k & 1) ? -ub[k >> 1] : lb[k >> 1]);
	}
	// The solver's index-only scan needs the smallest residual and its row number, not the rows.  The residual of safety row j
	// of a critical point (x_i, Q_i) at the iterate (u, d) is  Lgh u + h d + Lfh  with  Lfh = Dh Q f,  Lgh = Dh Q g,  i.e.
	//     Dh_j . (Q_i (f + g u)) + h_j d:
	// one nx x nx product per POINT on a vector formed once per scan, then one patterned inner product per row - instead of Dh Q
	// (per row), two inner products for Lfh and Lgh, and the residual's own.  Same number up to its last bits; the winner is
	// fetched again with get() and the solver works on the exact row (robust_kernel.cuh: scan_min).
#ifndef ASIF_TB_SCAN_MIN
#define ASIF_TB_SCAN_MIN 0 // measured neutral: C2 3.242 vs 3.238 ms per 1e7 states, C5 9.29 vs 9.31 ms per 1e6 (the scans of these kernels are bound by the snapshot loads and by lane divergence, not by the row arithmetic)
#endif
	static constexpr bool HAS_SCAN_MIN = ASIF_TB_SCAN_MIN != 0;
	__device__ __forceinline__ void scan_min(const double (&v)[NV], double &sr, int &pr) const
	{
		auto upd = [&](const double s, const int j) {
			const bool better = s < sr;
			sr = better ? s : sr;
			pr = better ? j : pr;
		};
		if (!trivial) {
			const int NC = nc();
			const int nk = nkept < count_np() ? nkept : count_np();
			const double dl = v[NU];
			double wv[NX];
#pragma unroll
			for (int m = 0; m < NX; m++) {
				double a = (M::f_pat(m) == PZ) ? 0.0 : f[m];
#pragma unroll
				for (int i = 0; i < NU; i++) {
					const int pg = M::g_pat(m + i * NX);
					if (pg == PZ) continue;
					a = (pg == P1) ? a + v[i] : fma(g[m + i * NX], v[i], a);
				}
				wv[m] = a;
			}
#pragma unroll 1
			for (int s = 0; s < nk; s++) {
				const int slot = slot_of(s);
				double xs[NS], hs[NPSS], Dhs[NPSS * NX], y[NX];
#pragma unroll
				for (int e = 0; e < NS; e++) xs[e] = snap[(slot * NS + e) * T];
				M::safety_set(xs, hs, Dhs);
#pragma unroll
				for (int m = 0; m < NX; m++) {
					double a = 0.0;
#pragma unroll
					for (int cc = 0; cc < NX; cc++) a = fma(xs[NX + m + cc * NX], wv[cc], a);
					y[m] = a;
				}
#pragma unroll
				for (int j = 0; j < NPSS; j++) {
					double r = hs[j] * dl;
#pragma unroll
					for (int m = 0; m < NX; m++) {
						const int pt = M::dhs_pat(j + m * NPSS);
						if (pt == PZ) continue;
						r = (pt == P1) ? r + y[m] : ((pt == PM1) ? r - y[m] : fma(Dhs[j + m * NPSS], y[m], r));
					}
					upd(r, s * NPSS + j);
				}
			}
			if (nk < count_np()) upd(dl, nk * NPSS); // missing points: rows (0, 1), rhs 0, all with the same residual - the first one counts
			double rT = -rhsT, rO = -rhsO;
#pragma unroll
			for (int i = 0; i < NU; i++) {
				rT = fma(lgT[i], v[i], rT);
				rO = fma(lgO[i], v[i], rO);
			}
			upd(rT, NC - 2);
			upd(rO, NC - 1);
		}
		const int NCb = nc();
#pragma unroll
		for (int k = 0; k < 2 * NV; k++) {
			const bool upper = (k & 1) != 0;
			const double vv = v[k >> 1];
			upd((upper ? -vv : vv) - (upper ? -ub[k >> 1] : lb[k >> 1]), NCb + k);
		}
	}
	__device__ __forceinline__ void get(const int j, double (&n)[NV], double &rhs) const
	{
		const int NC = nc();
		if (j >= NC) {
			bound_row(j - NC, n, rhs);
		} else if (trivial) {
#pragma unroll
			for (int i = 0; i < NV; i++) n[i] = 0.0;
			rhs = neg_inf;
		} else if (j >= NC - 2) {
			const bool o = (j == NC - 1);
#pragma unroll
			for (int i = 0; i < NU; i++) n[i] = o ? lgO[i] : lgT[i];
			n[NU] = 0.0;
			rhs = o ? rhsO : rhsT;
		} else {
			const int s = j / NPSS, jj = j - s * NPSS;
			if (s < nkept) {
				double nn[NPSS][NV], rr[NPSS];
				point_rows(slot_of(s), nn, rr);
#pragma unroll
				for (int t = 0; t < NPSS; t++) {
					if (t == jj) {
#pragma unroll
						for (int i = 0; i < NV; i++) n[i] = nn[t][i];
						rhs = rr[t];
					}
				}
			} else {
#pragma unroll
				for (int i = 0; i < NU; i++) n[i] = 0.0;
				n[NU] = 1.0;
				rhs = -0.0;
			}
		}
	}
};

// Linear cost of one state: updateCost(uDes) plus the relax entry of initialize() (src/asif_implicit_tb.cpp:206-212,
// 735-746), or the caller's c of the filter(x, H, c, ...) overloads (:252-259) when p.custom_cost is set.
template <int NU>
__device__ __forceinline__ void tb_cost_vector(const TbParams &p, const double *__restrict__ u_des, const int64_t kk,
                                               double (&cin)[NU + 1])
{
	if (p.custom_cost) {
#pragma unroll
		for (int i = 0; i < NU + 1; i++) cin[i] = u_des[kk * (NU + 1) + i];
	} else {
#pragma unroll
		for (int i = 0; i < NU; i++) cin[i] = -2.0 * u_des[kk * NU + i];
		cin[NU] = -2.0 * p.relaxCost * p.relaxSafeLb;
	}
}

// the solver's row scan for the TB rows (found by ADL, see qp_scan_rows in qp_gi.cuh): scan at the iterate with the
// per-point pre-filter
template <class M, int NPBTSS, class F, class FB>
__device__ __forceinline__ void qp_scan_rows(const TbRows<M, NPBTSS> &rows, const double (&v)[M::NU + 1], F &&fn, FB &&fb)
{
#ifndef ASIF_TB_SCAN_PREFILTER
#define ASIF_TB_SCAN_PREFILTER 0 // with the index-only scan the solver's per-row work is a compare and two selects: the per-point pre-test and its branch cost more than they save
#endif
	rows.template scan_at<(ASIF_TB_SCAN_PREFILTER != 0)>(v, fn, fb);
}

template <int NPBTSS>
struct TbDiagRec {
	double TTS, ortho, hSafetyNow, hBackupEnd;
	int critIdx[np_capacity(NPBTSS)];
	bool have_rows;
};

// One reference filter() call for one state (src/asif_implicit_tb.cpp:261-363).  snap is this
// thread's shared-memory view (already offset by threadIdx.x, element stride T).  Must be called
// by all 32 lanes of a warp together (warp-uniform early exit from the trajectory loop).
// R is left describing the state's QP rows (the diag writer re-scans them).
template <class M, int NPBTSS, bool WITH_DIAG, int SATMODE>
__device__ __forceinline__ int32_t tb_filter_one(const TbParams &p, const double (&x0)[M::NX], const double (&cin)[M::NU + 1],
                                                 double *snap, const int T, double (&uo)[M::NU], double &relax,
                                                 int &qp_iters, TbDiagRec<NPBTSS> &dg, TbRows<M, NPBTSS> &R)
{
	constexpr int NX = M::NX, NU = M::NU, NPSS = M::NPSS;
	constexpr int NS = NX + NX * NX;
	constexpr int NV = NU + 1;
	constexpr int CAP = np_capacity(NPBTSS);
	const int np = np_runtime(NPBTSS) ? p.npBTSS : NPBTSS;
	if (np_runtime(NPBTSS)) R.np = np;

	// ---- filter(): h_BS(x), hSafetyNow (:278-283)
	double hs[NPSS], Dhs[NPSS * NX];
	M::safety_set(x0, hs, Dhs);
	double hSafetyNow = hs[0];
#pragma unroll
	for (int j = 1; j < NPSS; j++) hSafetyNow = (hs[j] < hSafetyNow) ? hs[j] : hSafetyNow;
	const bool inside = M::backup_set_reached(x0);

	// ---- backup trajectory (Euler, :464-487) with streaming selection and hit detection (:505-528)
	double X[NS];
#pragma unroll
	for (int i = 0; i < NS; i++) X[i] = 0.0;
#pragma unroll
	for (int i = 0; i < NX; i++) X[i] = x0[i];
#pragma unroll
	for (int i = 0; i < NX; i++) X[NX + i * (NX + 1)] = 1.0;

	double key[CAP]; // ascending; +inf = empty
	int kidx[CAP];
#pragma unroll
	for (int s = 0; s < CAP; s++) {
		key[s] = INFINITY;
		kidx[s] = -1;
		R.kslot[s] = s;
	}
	// point 0
	key[0] = hSafetyNow;
	kidx[0] = 0;
	int nkept = 1;
#pragma unroll
	for (int e = 0; e < NS; e++) snap[(R.kslot[0] * NS + e) * T] = X[e];

	bool hit = false;
	int idxHit = 0;
	double hBackupEnd = 0.0;
	const int N = p.npBT;
	bool active = !inside; // still selecting critical points and looking for the first hit
	if (inside) key[CAP - 1] = -INFINITY; // trivial rows: the list is never read
	double bs_level = active ? M::backup_set_level() : -INFINITY;
	// Two nested loops: the inner one is the per-step instruction stream and leaves as soon as ANY lane of the
	// warp hits the backup set; the hit bookkeeping (once per state) lives between the loops, so it costs
	// nothing per step.  (A plain `if (hit_now) {...}` in the step body is if-converted by ptxas into ~11
	// predicated instructions that issue on every step.)
	// One loop; the hit bookkeeping (once per state) sits behind a warp-uniform branch, so it costs one vote and one
	// branch per step and nothing else.  (A plain `if (hit_now) {...}` is if-converted by ptxas into ~11 predicated
	// instructions that issue on every step; leaving the loop at every hit and re-entering it costs ~45 instructions
	// per event, 25-45 events per warp.)
	bool all_done = !WITH_DIAG && __all_sync(0xffffffffu, !active);
#pragma unroll TB_STEP_UNROLL
	for (int i = 1; i < N && !all_done; i++) {
		// rhs (:899-909) and Euler step: (rhs*dt) + prev, two roundings (:477-480)
		double Xd[NS], DfCL[NX * NX];
		backup_cl_dynamics<M, SATMODE>(p.sat, p.lb, p.ub, X, Xd, DfCL);
		sensitivity_rhs<M>(DfCL, X + NX, Xd + NX);
#pragma unroll
		for (int e = 0; e < NS; e++) X[e] = Xd[e] * p.backTrajDt + X[e];
		// selection and hit scan are frozen after the first hit (:507,539)
		double hmin;
		if (M::HAS_SAFETY_MIN) {
			hmin = M::safety_min(X);
		} else {
			M::safety_set(X, hs, Dhs);
			hmin = hs[0];
#pragma unroll
			for (int j = 1; j < NPSS; j++) hmin = (hs[j] < hmin) ? hs[j] : hmin;
		}
		if (hmin < key[CAP - 1]) { // lanes that are no longer selecting hold key[CAP-1] = -inf
			// evict the largest key, insert (hmin, i) keeping ascending order; ties keep the earlier index first
			const int slot = R.kslot[CAP - 1];
#pragma unroll
			for (int e = 0; e < NS; e++) snap[(slot * NS + e) * T] = X[e];
			double ck = hmin;
			int ci = i, cs = slot;
			bool ins = false; // once placed, everything behind shifts by one
#pragma unroll
			for (int s = 0; s < CAP; s++) {
				const bool sw = ins || (ck < key[s]);
				ins = sw;
				const double tk = key[s];
				const int ti = kidx[s], ts = R.kslot[s];
				key[s] = sw ? ck : tk;
				kidx[s] = sw ? ci : ti;
				R.kslot[s] = sw ? cs : ts;
				ck = sw ? tk : ck;
				ci = sw ? ti : ci;
				cs = sw ? ts : cs;
			}
			nkept = nkept < CAP ? nkept + 1 : CAP;
		}
		// models with a level form fold `active` into the level (-inf once the lane stopped looking: never reached)
		const bool hit_now = M::HAS_BACKUP_SET_REACHED ? M::backup_set_reached(X, bs_level) : (active && M::backup_set_reached(X));
		if (__any_sync(0xffffffffu, hit_now)) {
			if (hit_now) {
				hit = true;
				active = false;
				key[CAP - 1] = -INFINITY; // freezes the selection (:507,539): no later point can enter the list
				bs_level = -INFINITY;
				idxHit = i;
#pragma unroll
				for (int e = 0; e < NS; e++) snap[(CAP * NS + e) * T] = X[e];
			}
			all_done = !WITH_DIAG && __all_sync(0xffffffffu, !active);
		}
	}
	if (WITH_DIAG) hBackupEnd = M::backup_set_value(X);

	// ---- cost (:198-212, 735-746), bounds
	double c[NV];
	DiagMetric<NV> mt;
	R.snap = snap;
	R.nkept = (np_runtime(NPBTSS) && nkept > np) ? np : nkept;
	R.pack_slots();
	R.trivial = inside;
	R.neg_inf = -p.inf;
#pragma unroll
	for (int i = 0; i < NU; i++) {
		R.lb[i] = p.lb[i];
		R.ub[i] = p.ub[i];
		R.lgT[i] = 0.0;
		R.lgO[i] = 0.0;
	}
#pragma unroll
	for (int i = 0; i < NV; i++) c[i] = cin[i]; // linear cost: from uDes (tb_cost_vector) or the caller's c
#pragma unroll
	for (int i = 0; i < NX; i++) R.f[i] = 0.0;
#pragma unroll
	for (int i = 0; i < NX * NU; i++) R.g[i] = 0.0;
	R.rhsT = R.rhsO = -p.inf;
	R.lb[NU] = p.relaxSafeLb;
	R.ub[NU] = p.inf;
#pragma unroll
	for (int i = 0; i < NV; i++) {
		mt.gi[i] = p.gi[i];
		mt.gih[i] = p.gih[i];
	}

	double TTS = 0.0, ortho = 1.0;
	int32_t rc;
	bool solve = true;
	if (inside) {
		rc = 2;
	} else if (!hit) {
		ortho = 0.0;
		rc = -3; // backup set not reached (:354-361)
		solve = false;
	} else {
		rc = 1;
		if (M::FUSED_GRADIENT) {
			double u0[NU], dtmp[NX * NX];
#pragma unroll
			for (int i = 0; i < NU; i++) u0[i] = 0.0;
			M::dynamics_with_gradient(x0, u0, R.f, R.g, dtmp); // the dynamics_ lambda of :79-86
		} else {
			M::dynamics(x0, R.f, R.g);
		}
		// time-to-safety and orthogonality rows (:588-641)
		double xh[NS];
#pragma unroll
		for (int e = 0; e < NS; e++) xh[e] = snap[(CAP * NS + e) * T];
		double hBS, DhBS[NX], DDhBS[NX * NX], fCl[NX], DfCl[NX * NX];
		M::backup_set(xh, hBS, DhBS, DDhBS);
		backup_cl_dynamics<M, SATMODE>(p.sat, p.lb, p.ub, xh, fCl, DfCl);
#pragma unroll
		for (int r = 0; r < NX; r++)
#pragma unroll
			for (int cc = 0; cc < NX; cc++) { // structural entries are not written by backup_cl_dynamics
				const int pt = dfcl_pattern<M>(r, cc);
				if (pt == PZ) DfCl[r + cc * NX] = 0.0;
				if (pt == P1) DfCl[r + cc * NX] = 1.0;
			}
		double cosT = DhBS[0] * fCl[0];
#pragma unroll
		for (int m = 1; m < NX; m++) cosT = cosT + DhBS[m] * fCl[m];
		double n1 = 0.0, n2 = 0.0;
#pragma unroll
		for (int m = 0; m < NX; m++) {
			n1 += DhBS[m] * DhBS[m];
			n2 += fCl[m] * fCl[m];
		}
		const double den1 = sqrt(n1), den2 = sqrt(n2);
		const double den = den1 * den2;
		ortho = cosT / den;
		// t_i accumulates += dt in the reference (:475); it depends on i only, so the host tabulates it once
		const double tHit = p.t_of_index[idxHit];
		TTS = tHit;
		const double hReach = p.backTrajHorizon - tHit;
		double DhBSDx[NX];
#pragma unroll
		for (int cc = 0; cc < NX; cc++) {
			double acc = DhBS[0] * xh[NX + cc * NX];
#pragma unroll
			for (int m = 1; m < NX; m++) acc = acc + DhBS[m] * xh[NX + m + cc * NX];
			DhBSDx[cc] = acc;
		}
		double dhT[NX], dhO[NX];
#pragma unroll
		for (int i = 0; i < NX; i++) dhT[i] = DhBSDx[i] / cosT;
		const double hOrtho = ortho - p.backTrajMinOrtho;
		double DxHit[NX * NX];
#pragma unroll
		for (int r = 0; r < NX; r++)
#pragma unroll
			for (int cc = 0; cc < NX; cc++) DxHit[r + cc * NX] = xh[NX + r + cc * NX] - fCl[r] * DhBSDx[cc];
		const double denSq = den * den;
#pragma unroll
		for (int i = 0; i < NX; i++) {
			double Dnum = 0.0, Dden1 = 0.0, Dden2 = 0.0;
#pragma unroll
			for (int kq = 0; kq < NX; kq++) {
				double t1 = 0.0, t2 = 0.0;
#pragma unroll
				for (int l = 0; l < NX; l++) {
					t1 += DDhBS[kq + l * NX] * DxHit[l + i * NX];
					t2 += DfCl[kq + l * NX] * DxHit[l + i * NX];
				}
				const double t3 = DhBS[kq] * t2;
				const double t4 = t1 * fCl[kq];
				Dden1 += t3;
				Dden2 += t4;
				Dnum += t3 + t4;
			}
			const double Dden = den2 * Dden1 / den1 + den1 * Dden2 / den2;
			dhO[i] = (Dnum * den - cosT * Dden) / denSq;
		}
		// TTS row and ortho row: Lgh u >= -Lfh - relax*h, no relax-variable column (:655-674)
		double lfT = dhT[0] * R.f[0], lfO = dhO[0] * R.f[0];
#pragma unroll
		for (int m = 1; m < NX; m++) {
			lfT = lfT + dhT[m] * R.f[m];
			lfO = lfO + dhO[m] * R.f[m];
		}
#pragma unroll
		for (int i = 0; i < NU; i++) {
			double a = dhT[0] * R.g[i * NX], b = dhO[0] * R.g[i * NX];
#pragma unroll
			for (int m = 1; m < NX; m++) {
				a = a + dhT[m] * R.g[m + i * NX];
				b = b + dhO[m] * R.g[m + i * NX];
			}
			R.lgT[i] = a;
			R.lgO[i] = b;
		}
		R.rhsT = -lfT - p.relaxTTS * hReach;
		R.rhsO = -lfO - p.relaxMinOrtho * hOrtho;
	}

	// ---- QP + post-solve (:324-352)
	double v[NV];
	relax = 0.0;
	qp_iters = 0;
	int st = QP_PRIMAL_INFEASIBLE;
	if (solve) st = qp_gi_solve<NV>(mt, c, R, v, &qp_iters);
	if (solve && st == QP_OK) {
#pragma unroll
		for (int i = 0; i < NU; i++) uo[i] = input_saturate(v[i], p.lb[i], p.ub[i]);
		relax = v[NU];
	} else {
		double Du[NU * NX];
		M::backup_controller(x0, uo, Du);
#pragma unroll
		for (int i = 0; i < NU; i++) uo[i] = input_saturate(uo[i], p.lb[i], p.ub[i]);
		if (solve) rc = (rc == 2) ? -1 : st;
	}
	if (WITH_DIAG) {
		dg.TTS = TTS;
		dg.ortho = ortho;
		dg.hSafetyNow = hSafetyNow;
		dg.hBackupEnd = hBackupEnd;
#pragma unroll
		for (int s = 0; s < CAP; s++) dg.critIdx[s] = (!inside && hit && s < nkept) ? kidx[s] : -1;
		dg.have_rows = inside || hit;
	}
	return rc;
}

// resident CTAs per SM the register allocation is tuned for (nx = 2 models: 16 warps/SM)
template <class M>
__host__ __device__ constexpr int tb_min_blocks()
{
	return M::NX <= 2 ? ASIF_TB_MINBLOCKS_NX2 : ASIF_TB_MINBLOCKS_NX4;
}
// the fused rollout kernel carries the plant state and the loop bookkeeping on top of the filter: at 168 registers (3 CTAs per
// SM) the 1e5-agent fleet of BASELINE config 5 runs in 1.11 s against 1.20 s at 128 (its 782 tiles fill 2.6 CTAs per SM either
// way); a fleet large enough to fill the SM would lose ~1.5 % (measured on the filter kernel: 9.47 vs 9.33 ms per 1e6 states)
#ifndef ASIF_TB_ROLLOUT_MINBLOCKS_NX4
#define ASIF_TB_ROLLOUT_MINBLOCKS_NX4 3
#endif
template <class M>
__host__ __device__ constexpr int tb_rollout_min_blocks()
{
	return M::NX <= 2 ? ASIF_TB_MINBLOCKS_NX2 : ASIF_TB_ROLLOUT_MINBLOCKS_NX4;
}

// Where the critical-point snapshots live.  nx = 2: shared memory (30 doubles per thread).  nx = 4: 100 doubles per
// thread would cap the SM at 8 warps, and the FP64 pipe then idles on dependent-issue latency (ncu: issue slots 54 %
// busy); the snapshots are written only when a point enters the running k-smallest list and read only by the QP, so
// they go to a global scratch sized for the resident grid (L2 resident) and the kernel becomes persistent:
// 16 warps/SM, measured 16.5 -> 13.2 ms for 1e6 segway states.
constexpr int TB_SCRATCH_HEADER = 16; // doubles in front of the snapshot scratch: [0] = tile counter (zeroed per launch)

template <class M>
__host__ __device__ constexpr bool tb_global_snapshots()
{
	return M::NX > 2;
}

template <class M, int NPBTSS, bool WITH_DIAG, int SATMODE>
__device__ __forceinline__ void tb_filter_tile(const TbParams &p, const int64_t n, const int64_t k, double *snap, const int T,
                                               const double *__restrict__ x_in, const double *__restrict__ u_des,
                                               double *__restrict__ u_act, double *__restrict__ relax_out,
                                               int32_t *__restrict__ rc_out, double *__restrict__ diag,
                                               unsigned long long *__restrict__ qp_iter_sum)
{
	constexpr int NX = M::NX, NU = M::NU, NPSS = M::NPSS;
	constexpr int NV = NU + 1;
	const int np = np_runtime(NPBTSS) ? p.npBTSS : NPBTSS;
	const int NC = np * NPSS + 2;
	const int NDIAG = tb_diag_head(np) + NC * NV + NC;

	const bool live = k < n;
	const int64_t kk = live ? k : (n - 1); // tail lanes redo the last state and do not store

	double x0[NX], cin[NU + 1];
#pragma unroll
	for (int i = 0; i < NX; i++) x0[i] = x_in[kk * NX + i];
	tb_cost_vector<NU>(p, u_des, kk, cin);

	double uo[NU], relax;
	int qp_iters;
	TbDiagRec<NPBTSS> dg;
	TbRows<M, NPBTSS> R;
	const int32_t rc = tb_filter_one<M, NPBTSS, WITH_DIAG, SATMODE>(p, x0, cin, snap, T, uo, relax, qp_iters, dg, R);

	if (live) {
#pragma unroll
		for (int i = 0; i < NU; i++) u_act[k * NU + i] = uo[i];
		relax_out[k] = relax;
		rc_out[k] = rc;
		if (WITH_DIAG) {
			double *d = diag + k * NDIAG;
			d[0] = dg.TTS;
			d[1] = dg.ortho;
			d[2] = dg.hSafetyNow;
			d[3] = dg.hBackupEnd;
#pragma unroll
			for (int s = 0; s < np_capacity(NPBTSS); s++)
				if (s < np) d[4 + s] = (double)dg.critIdx[s];
			double *A = d + 4 + np, *b = A + NC * NV;
			for (int j = 0; j < NC; j++) {
				for (int i = 0; i < NV; i++) A[j + i * NC] = 0.0;
				b[j] = (dg.have_rows && R.trivial) ? -p.inf : 0.0;
			}
			if (dg.have_rows && !R.trivial) {
				R.scan(
				    [&](const int j, const double(&nn)[NV], const double rhs) {
#pragma unroll
					    for (int i = 0; i < NV; i++) A[j + i * NC] = nn[i];
					    b[j] = rhs;
				    },
				    [](const int, const int, const bool, const double) {});
			}
		}
	}
	if (qp_iter_sum) {
		unsigned int it = live ? (unsigned int)qp_iters : 0u;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) it += __shfl_xor_sync(0xffffffffu, it, o);
		if ((threadIdx.x & 31) == 0 && it) qp_rows_add(qp_iter_sum, (unsigned long long)it);
	}
}

// gsnap: global snapshot scratch, gridDim.x * tb_smem_doubles_per_thread * blockDim.x doubles (nx = 4 models), else unused
template <class M, int NPBTSS, bool WITH_DIAG, int SATMODE>
__global__ void __launch_bounds__(TB_THREADS, tb_min_blocks<M>())
tb_filter_kernel(const TbParams p, const int64_t n, const double *__restrict__ x_in, const double *__restrict__ u_des,
                 double *__restrict__ u_act, double *__restrict__ relax_out, int32_t *__restrict__ rc_out,
                 double *__restrict__ diag, unsigned long long *__restrict__ qp_iter_sum, double *__restrict__ gsnap)
{
	constexpr int T = TB_THREADS; // every launch uses TB_THREADS threads (engine_internal.cuh): compile-time strides
	if (tb_global_snapshots<M>()) {
		// persistent: every warp pulls tiles of 32 states from a counter that lives in front of the scratch
		unsigned long long *next_tile = reinterpret_cast<unsigned long long *>(gsnap);
		double *snap = gsnap + TB_SCRATCH_HEADER + (int64_t)blockIdx.x * tb_smem_doubles_per_thread<M, NPBTSS>() * T + threadIdx.x;
		for (;;) {
			unsigned long long tile = 0;
			if ((threadIdx.x & 31) == 0) tile = atomicAdd(next_tile, 1ull);
			tile = __shfl_sync(0xffffffffu, tile, 0);
			if ((int64_t)tile * 32 >= n) break;
			tb_filter_tile<M, NPBTSS, WITH_DIAG, SATMODE>(p, n, (int64_t)tile * 32 + (threadIdx.x & 31), snap, T, x_in, u_des, u_act,
			                                              relax_out, rc_out, diag, qp_iter_sum);
		}
	} else {
		extern __shared__ double smem[];
		double *snap = smem + threadIdx.x; // [(NPBTSS+1)*NS][T]
		tb_filter_tile<M, NPBTSS, WITH_DIAG, SATMODE>(p, n, (int64_t)blockIdx.x * T + threadIdx.x, snap, T, x_in, u_des, u_act,
		                                              relax_out, rc_out, diag, qp_iter_sum);
	}
}

// Closed-loop rollout (examples/segway_implicit_tb.cpp:251-283): the state never leaves the
// registers between control steps; one launch covers all steps of every agent.
template <class M, int NPBTSS, int SATMODE>
__global__ void __launch_bounds__(TB_THREADS, tb_rollout_min_blocks<M>())
tb_rollout_kernel(const TbParams p, const int64_t n, const int32_t steps, const double dt_plant, double *__restrict__ x_io,
                  const double *__restrict__ u_des, double *__restrict__ u_act_last, int32_t *__restrict__ rc_last,
                  unsigned long long *__restrict__ rc_hist, unsigned long long *__restrict__ qp_iter_sum, double *__restrict__ gsnap)
{
	constexpr int NX = M::NX, NU = M::NU;
	extern __shared__ double smem[];
	constexpr int T = TB_THREADS;
	constexpr bool GS = tb_global_snapshots<M>();
	double *snap = GS ? gsnap + TB_SCRATCH_HEADER + (int64_t)blockIdx.x * tb_smem_doubles_per_thread<M, NPBTSS>() * T + threadIdx.x
	                  : smem + threadIdx.x;
	unsigned long long *next_tile = reinterpret_cast<unsigned long long *>(gsnap);
	unsigned int hist[8];
#pragma unroll
	for (int i = 0; i < 8; i++) hist[i] = 0;
	unsigned long long iters = 0;
	for (bool first = true;; first = false) { // nx = 4: warps pull tiles of 32 agents; nx = 2: one tile per CTA
	int64_t k;
	if (GS) {
		unsigned long long tile = 0;
		if ((threadIdx.x & 31) == 0) tile = atomicAdd(next_tile, 1ull);
		tile = __shfl_sync(0xffffffffu, tile, 0);
		if ((int64_t)tile * 32 >= n) break;
		k = (int64_t)tile * 32 + (threadIdx.x & 31);
	} else {
		if (!first) break;
		k = (int64_t)blockIdx.x * T + threadIdx.x;
	}
	const bool live = k < n;
	const int64_t kk = live ? k : (n - 1);
	double x[NX], ud[NU + 1], uo[NU];
#pragma unroll
	for (int i = 0; i < NX; i++) x[i] = x_io[kk * NX + i];
	tb_cost_vector<NU>(p, u_des, kk, ud); // the rollout holds uDes per agent; ud is its cost vector
#pragma unroll
	for (int i = 0; i < NU; i++) uo[i] = 0.0;
	int32_t rc = 0;
	for (int32_t s = 0; s < steps; s++) {
		double relax;
		int qp_iters;
		TbDiagRec<NPBTSS> dg;
		TbRows<M, NPBTSS> R;
		rc = tb_filter_one<M, NPBTSS, false, SATMODE>(p, x, ud, snap, T, uo, relax, qp_iters, dg, R);
		iters += live ? (unsigned long long)qp_iters : 0ull;
		const int slot = (rc >= -3 && rc <= 2) ? rc + 3 : 7;
#pragma unroll
		for (int i = 0; i < 8; i++) hist[i] += (live && i == slot) ? 1u : 0u;
		// plant step: fCl = f + g uAct ; x += dt*fCl  (:265-283)
		double f[NX], g[NX * NU];
		M::dynamics(x, f, g);
#pragma unroll
		for (int i = 0; i < NX; i++) {
			double fcl = f[i];
#pragma unroll
			for (int j = 0; j < NU; j++) fcl += g[i + j * NX] * uo[j];
			x[i] += dt_plant * fcl;
		}
	}
	if (live) {
#pragma unroll
		for (int i = 0; i < NX; i++) x_io[k * NX + i] = x[i];
#pragma unroll
		for (int i = 0; i < NU; i++) u_act_last[k * NU + i] = uo[i];
		rc_last[k] = rc;
	}
	} // tiles
	if (rc_hist) {
#pragma unroll
		for (int i = 0; i < 8; i++) {
			unsigned int c = hist[i];
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
			if ((threadIdx.x & 31) == 0 && c) atomicAdd(rc_hist + i, (unsigned long long)c);
		}
	}
	if (qp_iter_sum) {
		unsigned long long it = iters;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) it += __shfl_xor_sync(0xffffffffu, it, o);
		if ((threadIdx.x & 31) == 0 && it) qp_rows_add(qp_iter_sum, it);
	}
}

} // namespace asifb
