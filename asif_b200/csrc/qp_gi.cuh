// qp_gi.cuh -- exact per-thread solver for the tiny strictly convex QPs of the safety filters.
//
// Replaces the reference's QPWrapperOsqp::solve (src/qpwrapper_osqp.cpp:217-239, OSQP ADMM) for
//      min v'Hv + c'v   s.t.  rows(v) >= rhs,   nv <= 4, H diagonal (diagonalCost, the default)
// with the Goldfarb-Idnani dual active-set method: start at the unconstrained minimiser, add the
// most violated row, take primal/dual steps that keep every multiplier non-negative, drop blocking
// rows.  It terminates at the unique optimum (what "OSQP with polish, eps 1e-8" converges to) in
// a handful of iterations and detects infeasibility exactly (OSQP_PRIMAL_INFEASIBLE = -3).
//
// Everything is per thread and in registers: NV is a template constant, all loops over NV are
// unrolled, the active set holds at most NV rows.  Rows are presented by a functor with two
// members -- scan(f, fb): call f(j, normal, rhs) for every general row j in order and
// fb(j, var, upper, bound) for every variable-bound row; get(j, normal, rhs): one row -- so that the caller decides where they live (shared memory for the implicit filters, computed on the fly
// from a table for the robust/realizable ones, global memory for the generic batch entry).
// The solver's own inner products and updates are written with fma(): the bit-exact units are compiled with -fmad=false
// for the sake of the trajectory and the ROWS (which are compared bit for bit with the reference), but the QP step has no
// bit parity to keep (the reference's OSQP result is an iterate at eps 1e-3), and one DFMA per term instead of DMUL + DADD
// is both more accurate and a third fewer FP64 instructions in the solver.
// Work is done in the metric of the Hessian (v-hat = sqrt(2H) v) with every processed row
// normalised to unit length, so the thresholds below are scale free.
#pragma once
#include <cuda_runtime.h>

namespace asifb {

constexpr int QP_OK = 1;
constexpr int QP_MAX_ITER = -2;          // OSQP_MAX_ITER_REACHED
constexpr int QP_PRIMAL_INFEASIBLE = -3; // OSQP_PRIMAL_INFEASIBLE

constexpr double QP_FEAS_TOL = 1e-9;   // a row is violated when row.v - rhs < -tol * max(1, |row|_inf)
constexpr double QP_INDEP_TOL = 1e-18; // sin^2 of the angle to the active span below this: dependent
constexpr double QP_DUAL_TOL = 1e-12;  // dual direction entries below this are treated as <= 0

// Metric of a diagonal Hessian (diagonalCost = true, the reference default):
// gi[i] = 1/(2 H_ii), gih[i] = sqrt(gi[i]).
template <int NV>
struct DiagMetric {
	double gi[NV], gih[NV];
	__device__ __forceinline__ void unconstrained(const double (&c)[NV], double (&v)[NV]) const
	{
#pragma unroll
		for (int i = 0; i < NV; i++) v[i] = -(gi[i] * c[i]);
	}
	__device__ __forceinline__ void to_hat(const double (&n)[NV], double (&nh)[NV]) const
	{
#pragma unroll
		for (int i = 0; i < NV; i++) nh[i] = gih[i] * n[i];
	}
	__device__ __forceinline__ void from_hat(const double (&zh)[NV], double (&z)[NV]) const
	{
#pragma unroll
		for (int i = 0; i < NV; i++) z[i] = gih[i] * zh[i];
	}
};

// Metric of a dense SPD Hessian: 2H = L L' (lower Cholesky factor, L[i][j], j <= i).
template <int NV>
struct CholMetric {
	double L[NV][NV];
	// returns false when 2H is not positive definite
	__device__ __forceinline__ bool factor(const double *H /* column-major NV x NV */)
	{
		bool ok = true;
#pragma unroll
		for (int j = 0; j < NV; j++) {
			double d = 2.0 * H[j + j * NV];
#pragma unroll
			for (int k = 0; k < NV; k++)
				if (k < j) d -= L[j][k] * L[j][k];
			ok = ok && (d > 0.0);
			d = sqrt(d);
			L[j][j] = d;
#pragma unroll
			for (int i = 0; i < NV; i++) {
				if (i > j) {
					double t = 2.0 * H[i + j * NV];
#pragma unroll
					for (int k = 0; k < NV; k++)
						if (k < j) t -= L[i][k] * L[j][k];
					L[i][j] = t / d;
				}
			}
		}
		return ok;
	}
	__device__ __forceinline__ void fwd(double (&y)[NV]) const // y <- L^-1 y
	{
#pragma unroll
		for (int i = 0; i < NV; i++) {
			double t = y[i];
#pragma unroll
			for (int k = 0; k < NV; k++)
				if (k < i) t -= L[i][k] * y[k];
			y[i] = t / L[i][i];
		}
	}
	__device__ __forceinline__ void bwd(double (&y)[NV]) const // y <- L^-T y
	{
#pragma unroll
		for (int i = NV - 1; i >= 0; i--) {
			double t = y[i];
#pragma unroll
			for (int k = 0; k < NV; k++)
				if (k > i) t -= L[k][i] * y[k];
			y[i] = t / L[i][i];
		}
	}
	__device__ __forceinline__ void unconstrained(const double (&c)[NV], double (&v)[NV]) const
	{
#pragma unroll
		for (int i = 0; i < NV; i++) v[i] = -c[i];
		fwd(v);
		bwd(v);
	}
	__device__ __forceinline__ void to_hat(const double (&n)[NV], double (&nh)[NV]) const
	{
#pragma unroll
		for (int i = 0; i < NV; i++) nh[i] = n[i];
		fwd(nh);
	}
	__device__ __forceinline__ void from_hat(const double (&zh)[NV], double (&z)[NV]) const
	{
#pragma unroll
		for (int i = 0; i < NV; i++) z[i] = zh[i];
		bwd(z);
	}
};

template <int NV>
struct QpWork {
	double v[NV];
	double mu[NV];
	double Q[NV][NV]; // Q[a][:] orthonormal basis vector a (hat space)
	double R[NV][NV]; // upper triangular, R[b][a] for b <= a
	int act[NV];
	int q;
};

// fetch row j, scale into hat space and normalise.  returns the hat-space length.
template <int NV, class Rows, class Metric>
__device__ __forceinline__ double qp_fetch_unit_row(const Rows &rows, int j, const Metric &mt, double (&nh)[NV],
                                                    double &rhs)
{
	double n[NV];
	rows.get(j, n, rhs);
	mt.to_hat(n, nh);
	double len2 = 0.0;
#pragma unroll
	for (int i = 0; i < NV; i++) len2 = fma(nh[i], nh[i], len2);
	const double inv = rsqrt(len2); // one rsqrt (18 instructions, no slow path) instead of sqrt + division (45 + two subroutines)
#pragma unroll
	for (int i = 0; i < NV; i++) nh[i] *= inv;
	return len2 * inv;
}

// rebuild Q, R from the current active set (modified Gram-Schmidt in hat space)
template <int NV, class Rows, class Metric>
__device__ __forceinline__ void qp_rebuild(QpWork<NV> &w, const Rows &rows, const Metric &mt)
{
#pragma unroll
	for (int a = 0; a < NV; a++) {
		if (a < w.q) {
			double nh[NV], rhs;
			qp_fetch_unit_row<NV>(rows, w.act[a], mt, nh, rhs);
#pragma unroll
			for (int b = 0; b < NV; b++) {
				if (b < a) {
					double d = 0.0;
#pragma unroll
					for (int i = 0; i < NV; i++) d = fma(w.Q[b][i], nh[i], d);
					w.R[b][a] = d;
#pragma unroll
					for (int i = 0; i < NV; i++) nh[i] = fma(-d, w.Q[b][i], nh[i]);
				}
			}
			double l2 = 0.0;
#pragma unroll
			for (int i = 0; i < NV; i++) l2 = fma(nh[i], nh[i], l2);
			const double inv = rsqrt(l2);
			w.R[a][a] = l2 * inv;
#pragma unroll
			for (int i = 0; i < NV; i++) w.Q[a][i] = nh[i] * inv;
		}
	}
}

// Row scan at the current iterate.  The default ignores the iterate; a row functor whose rows come in families of which
// only one member can be the most violated at a given point overloads this (found by ADL) to present that member only.
template <int NV, class Rows, class F, class FB>
__device__ __forceinline__ void qp_scan_rows(const Rows &rows, const double (&)[NV], F &&fn, FB &&fb)
{
	rows.scan(fn, fb);
}

// Row functors whose scans are long and whose single rows are cheap to fetch again (the 100 half-planes of the robust
// filter) opt into the index-only scan: declare  static constexpr bool SCAN_INDEX_ONLY = true;
template <class Rows, class = void>
struct qp_scan_index_only {
	static constexpr bool value = false;
};
template <class Rows>
struct qp_scan_index_only<Rows, decltype((void)Rows::SCAN_INDEX_ONLY)> {
	static constexpr bool value = Rows::SCAN_INDEX_ONLY;
};

// ... and may bring their own search for the smallest residual:  static constexpr bool HAS_SCAN_MIN = true;  scan_min(v, sr, pr)
template <class Rows, class = void>
struct qp_has_scan_min {
	static constexpr bool value = false;
};
template <class Rows>
struct qp_has_scan_min<Rows, decltype((void)Rows::HAS_SCAN_MIN)> {
	static constexpr bool value = Rows::HAS_SCAN_MIN;
};

// Vertex polish.  At a vertex (NV active rows) the minimiser is fixed by the rows alone: N v = rhs.  The dual method reaches
// it by steps along directions orthogonalised in hat space, whose error is eps / sin(angle between active rows) times the
// length of the step - for a safety row with h ~ 1e-6 against the orthogonality row (3.7e-7 rad apart in hat space, the
// relax variable at 1.8e4) that was 6e-5 in u, found against an exact rational solve in the round-2 parity run (3 states
// in 2.7e6).  Such vertices are re-solved from the original rows by Gaussian elimination with partial pivoting, whose error
// is eps times the condition number of N only; the result replaces v when it satisfies the active rows at least as well.
// Rare (a few states in 1e4), so it is kept out of line and works on local arrays.
constexpr double QP_POLISH_SIN = 1e-2; // polish when some R[a][a] (sine of the angle to the earlier active rows) is below this

#ifndef QP_POLISH_CALL
#define QP_POLISH_CALL __noinline__
#endif
template <int NV>
static __device__ QP_POLISH_CALL bool qp_vertex_solve(double (&M)[NV][NV + 1], double (&out)[NV])
{
	for (int c = 0; c < NV; c++) {
		int piv = c;
		double best = fabs(M[c][c]);
		for (int i = c + 1; i < NV; i++)
			if (fabs(M[i][c]) > best) {
				best = fabs(M[i][c]);
				piv = i;
			}
		if (!(best > 0.0)) return false;
		if (piv != c)
			for (int j = 0; j <= NV; j++) {
				const double t = M[c][j];
				M[c][j] = M[piv][j];
				M[piv][j] = t;
			}
		for (int i = c + 1; i < NV; i++) {
			const double f = M[i][c] / M[c][c];
			for (int j = c; j <= NV; j++) M[i][j] -= f * M[c][j];
		}
	}
	for (int i = NV - 1; i >= 0; i--) {
		double t = M[i][NV];
		for (int j = i + 1; j < NV; j++) t -= M[i][j] * out[j];
		out[i] = t / M[i][i];
	}
	bool fin = true;
	for (int i = 0; i < NV; i++) fin = fin && (fabs(out[i]) < INFINITY);
	return fin;
}

// N = the active rows [normal | rhs]; v is replaced by the direct solve when that satisfies the rows at least as well
template <int NV>
static __device__ QP_POLISH_CALL void qp_vertex_polish_rows(const double (&N)[NV][NV + 1], double (&v)[NV])
{
	double M[NV][NV + 1], vp[NV];
	for (int a = 0; a < NV; a++)
		for (int i = 0; i <= NV; i++) M[a][i] = N[a][i];
	if (!qp_vertex_solve<NV>(M, vp)) return;
	double r_old = 0.0, r_new = 0.0; // worst active-row residual, relative to max(1, |row|_inf)
	for (int a = 0; a < NV; a++) {
		double so = -N[a][NV], sn = -N[a][NV], nmax = 1.0;
		for (int i = 0; i < NV; i++) {
			so += N[a][i] * v[i];
			sn += N[a][i] * vp[i];
			nmax = fabs(N[a][i]) > nmax ? fabs(N[a][i]) : nmax;
		}
		r_old = fabs(so) / nmax > r_old ? fabs(so) / nmax : r_old;
		r_new = fabs(sn) / nmax > r_new ? fabs(sn) / nmax : r_new;
	}
	if (r_new <= r_old)
		for (int i = 0; i < NV; i++) v[i] = vp[i];
}

// The same with the rows fetched out of line: the row functor arrives as a pointer to a COPY the caller makes inside its rare branch: the solver's own functor never has
// its address taken and stays in registers (passing it by reference cost the hot kernels 2-7 %, round-2 A/B).
template <int NV, class Rows>
static __device__ __noinline__ void qp_vertex_polish(const Rows *rows, const int *act, double *v)
{
	double N[NV][NV + 1], vv[NV];
	for (int a = 0; a < NV; a++) {
		double n[NV], rhs = 0.0;
		rows->get(act[a], n, rhs);
		for (int i = 0; i < NV; i++) N[a][i] = n[i];
		N[a][NV] = rhs;
	}
	for (int i = 0; i < NV; i++) vv[i] = v[i];
	qp_vertex_polish_rows<NV>(N, vv);
	for (int i = 0; i < NV; i++) v[i] = vv[i];
}

// v returns the minimiser when the result is QP_OK.
// iters (optional) returns the number of rows processed (for the K-bar statistic).
template <int NV, class Rows, class Metric>
__device__ __forceinline__ int qp_gi_solve(const Metric &mt, const double (&c)[NV], const Rows &rows, double (&v)[NV],
                                           int *iters = nullptr)
{
	QpWork<NV> w;
	w.q = 0;
	mt.unconstrained(c, w.v);
#pragma unroll
	for (int i = 0; i < NV; i++) {
		w.mu[i] = 0.0;
		w.act[i] = -1;
	}
	int status = QP_MAX_ITER;
	int it = 0;
	const int max_outer = 4 * NV + 24;
	for (; it < max_outer; it++) {
		// ---- most violated row that is not active (rows.scan visits every row once, in order)
		int p = -1;
		double sp = 0.0;
		double np_[NV];
#pragma unroll
		for (int i = 0; i < NV; i++) np_[i] = 0.0;
		bool scanned = false;
		if (qp_scan_index_only<Rows>::value) {
			// Index-only scan.  In the scan below the "a more violated row" body runs ~ln(rows) times per lane, each time for
			// that lane alone (55-60 times per warp and scan over 100 half-planes: a quarter of the scan's issue slots at 1-2
			// lanes).  Here the loop only tracks the smallest residual and its row number with selects; the tolerance and
			// active-set tests are applied once, to the winner.  If the winner passes them it is exactly the row the scan
			// below would pick (the smallest residual among the rows that pass, first index on ties); if its residual is not
			// below -tol no row's is and the solve is finished; only a winner that is violated AND fails a test (active, or
			// inside its own scaled tolerance - a knife edge) sends the lane through the full scan.
			int pr = -1;
			double sr = 0.0;
			if constexpr (qp_has_scan_min<Rows>::value) {
				rows.scan_min(w.v, sr, pr); // the row functor's own minimum search (robust_kernel.cuh)
			} else
			qp_scan_rows(
			    rows, w.v,
			    [&](const int j, const double(&n)[NV], const double rhs) {
				    double s = -rhs;
#pragma unroll
				    for (int i = 0; i < NV; i++) s = fma(n[i], w.v[i], s);
				    const bool better = s < sr;
				    sr = better ? s : sr;
				    pr = better ? j : pr;
			    },
			    [&](const int j, const int var, const bool upper, const double bnd) {
				    double vv = 0.0;
#pragma unroll
				    for (int i = 0; i < NV; i++) vv = (i == var) ? w.v[i] : vv;
				    const double s = (upper ? -vv : vv) - bnd;
				    const bool better = s < sr;
				    sr = better ? s : sr;
				    pr = better ? j : pr;
			    });
			scanned = true;
			if (pr >= 0 && sr < -QP_FEAS_TOL) {
				double n[NV], rhs;
				rows.get(pr, n, rhs);
				double nmax = 1.0;
#pragma unroll
				for (int i = 0; i < NV; i++) nmax = (fabs(n[i]) > nmax) ? fabs(n[i]) : nmax;
				bool is_act = false;
#pragma unroll
				for (int a = 0; a < NV; a++) is_act |= (a < w.q) && (w.act[a] == pr);
				if (!is_act && sr < -QP_FEAS_TOL * nmax) {
					p = pr;
					sp = sr;
#pragma unroll
					for (int i = 0; i < NV; i++) np_[i] = n[i];
				} else {
					scanned = false; // knife edge: the full scan decides
				}
			}
		}
		if (!scanned)
		qp_scan_rows(
		    rows, w.v,
		    [&](const int j, const double(&n)[NV], const double rhs) {
			    double s = -rhs;
#pragma unroll
			    for (int i = 0; i < NV; i++) s += n[i] * w.v[i];
			    if (s < -QP_FEAS_TOL && s < sp) { // rarely true: the expensive part of the test only runs then
				    double nmax = 1.0; // max(1, |n|_inf) by compare and select (fmax would add NaN handling to every term)
#pragma unroll
				    for (int i = 0; i < NV; i++) nmax = (fabs(n[i]) > nmax) ? fabs(n[i]) : nmax;
				    bool is_act = false;
#pragma unroll
				    for (int a = 0; a < NV; a++) is_act |= (a < w.q) && (w.act[a] == j);
				    if (!is_act && s < -QP_FEAS_TOL * nmax) {
					    sp = s;
					    p = j;
#pragma unroll
					    for (int i = 0; i < NV; i++) np_[i] = n[i];
				    }
			    }
		    },
		    // variable bound: +v[var] >= bnd (lower) or -v[var] >= bnd (upper); same value as the generic row
		    [&](const int j, const int var, const bool upper, const double bnd) {
			    double vv = 0.0;
#pragma unroll
			    for (int i = 0; i < NV; i++) vv = (i == var) ? w.v[i] : vv;
			    const double s = (upper ? -vv : vv) - bnd;
			    if (s < -QP_FEAS_TOL && s < sp) {
				    bool is_act = false;
#pragma unroll
				    for (int a = 0; a < NV; a++) is_act |= (a < w.q) && (w.act[a] == j);
				    if (!is_act) {
					    sp = s;
					    p = j;
#pragma unroll
					    for (int i = 0; i < NV; i++) np_[i] = (i == var) ? (upper ? -1.0 : 1.0) : 0.0;
				    }
			    }
		    });
		if (p < 0) {
			status = QP_OK;
			break;
		}
		double nh[NV];
		mt.to_hat(np_, nh);
		double len2 = 0.0;
#pragma unroll
		for (int i = 0; i < NV; i++) len2 = fma(nh[i], nh[i], len2);
		if (!(len2 > 0.0)) { // 0 >= rhs with rhs > 0
			status = QP_PRIMAL_INFEASIBLE;
			break;
		}
		{
			const double inv = rsqrt(len2);
#pragma unroll
			for (int i = 0; i < NV; i++) nh[i] *= inv;
			sp = sp * inv;
		}
		double mu_p = 0.0;
		bool failed = false;
		// ---- primal/dual steps until row p becomes active (at most q drops)
		for (int inner = 0; inner <= NV + 1; inner++) {
			double d[NV], zh[NV], r[NV];
#pragma unroll
			for (int i = 0; i < NV; i++) zh[i] = nh[i];
#pragma unroll
			for (int a = 0; a < NV; a++) {
				d[a] = 0.0;
				if (a < w.q) {
					double t = 0.0;
#pragma unroll
					for (int i = 0; i < NV; i++) t = fma(w.Q[a][i], zh[i], t);
					d[a] = t;
#pragma unroll
					for (int i = 0; i < NV; i++) zh[i] = fma(-t, w.Q[a][i], zh[i]);
				}
			}
			double zz = 0.0;
#pragma unroll
			for (int i = 0; i < NV; i++) zz = fma(zh[i], zh[i], zz);
			// r = R^-1 d (back substitution over the active part)
#pragma unroll
			for (int a = NV - 1; a >= 0; a--) {
				r[a] = 0.0;
				if (a < w.q) {
					double t = d[a];
#pragma unroll
					for (int b = a + 1; b < NV; b++)
						if (b < w.q) t = fma(-w.R[a][b], r[b], t);
					// a new row orthogonal to an active one (a variable bound against a row without that variable) gives an
					// exact 0 here, and a zero numerator sends CUDA's division through its slow-path subroutine
					r[a] = (t == 0.0) ? t : t / w.R[a][a];
				}
			}
			double t1 = INFINITY;
			int l = -1;
#pragma unroll
			for (int a = 0; a < NV; a++) {
				if (a < w.q && r[a] > QP_DUAL_TOL) {
					const double t = w.mu[a] / r[a];
					if (t < t1) {
						t1 = t;
						l = a;
					}
				}
			}
			const bool indep = (w.q < NV) && (zz > QP_INDEP_TOL);
			const double t2 = indep ? (-sp / zz) : INFINITY;
			if (l < 0 && !indep) {
				failed = true;
				break;
			}
			if (t2 <= t1) { // full step: row p becomes active
				double z[NV];
				mt.from_hat(zh, z);
#pragma unroll
				for (int i = 0; i < NV; i++) w.v[i] = fma(t2, z[i], w.v[i]);
#pragma unroll
				for (int a = 0; a < NV; a++)
					if (a < w.q) w.mu[a] = fma(-t2, r[a], w.mu[a]);
				mu_p += t2;
				const double inv = rsqrt(zz);
				const double zl = zz * inv;
#pragma unroll
				for (int a = 0; a < NV; a++) {
					if (a == w.q) {
						w.act[a] = p;
						w.mu[a] = mu_p;
						w.R[a][a] = zl;
#pragma unroll
						for (int i = 0; i < NV; i++) w.Q[a][i] = zh[i] * inv;
#pragma unroll
						for (int b = 0; b < NV; b++)
							if (b < a) w.R[b][a] = d[b];
					}
				}
				w.q++;
				break;
			}
			// partial step: blocking row l leaves the active set
			if (indep) {
				double z[NV];
				mt.from_hat(zh, z);
#pragma unroll
				for (int i = 0; i < NV; i++) w.v[i] = fma(t1, z[i], w.v[i]);
				sp = fma(t1, zz, sp);
			}
#pragma unroll
			for (int a = 0; a < NV; a++)
				if (a < w.q) w.mu[a] = fma(-t1, r[a], w.mu[a]);
			mu_p += t1;
#pragma unroll
			for (int a = 0; a < NV - 1; a++) {
				if (a >= l && a + 1 < w.q) {
					w.act[a] = w.act[a + 1];
					w.mu[a] = w.mu[a + 1];
				}
			}
			w.q--;
			qp_rebuild<NV>(w, rows, mt);
		}
		if (failed) {
			status = QP_PRIMAL_INFEASIBLE;
			break;
		}
	}
	if (iters) *iters = it;
#ifndef ASIF_QP_NO_POLISH // (A/B switch of scripts/build_variant.sh; the product is always built with the polish)
	if (status == QP_OK && w.q == NV) {
		double rmin = w.R[0][0];
#pragma unroll
		for (int a = 1; a < NV; a++) rmin = (w.R[a][a] < rmin) ? w.R[a][a] : rmin;
		if (rmin < QP_POLISH_SIN) { // rare
			double vv[NV];
#pragma unroll
			for (int i = 0; i < NV; i++) vv[i] = w.v[i];
#ifdef ASIF_QP_POLISH_INLINE_ROWS
			// rows fetched here, only plain arrays go out of line.  Which of the two forms costs the hot kernels less was
			// measured per translation unit (round-2 A/B, 1e7 C2 states / 1e6 C5 states; no polish 3.45 / 12.10 ms):
			// this form 3.54 / 12.92 ms, the copy form below 3.59 / 12.73 ms - engine.cu defines the macro, kernels_contract.cu not.
			double N[NV][NV + 1];
#pragma unroll 1
			for (int a = 0; a < NV; a++) {
				double n[NV], rhs = 0.0;
				int ja = 0;
#pragma unroll
				for (int t = 0; t < NV; t++) ja = (t == a) ? w.act[t] : ja;
				rows.get(ja, n, rhs);
#pragma unroll
				for (int i = 0; i < NV; i++) N[a][i] = n[i];
				N[a][NV] = rhs;
			}
			qp_vertex_polish_rows<NV>(N, vv);
#else
			const Rows rows_copy = rows; // copies go to local memory here and only here
			int act[NV];
#pragma unroll
			for (int i = 0; i < NV; i++) act[i] = w.act[i];
			qp_vertex_polish<NV, Rows>(&rows_copy, act, vv);
#endif
#pragma unroll
			for (int i = 0; i < NV; i++) w.v[i] = vv[i];
		}
	}
#endif
#pragma unroll
	for (int i = 0; i < NV; i++) v[i] = w.v[i];
	return status;
}

} // namespace asifb
