// qp_admm.cuh -- the QPWrapper backend for nv > 4: a team-cooperative dense operator-splitting solver with an
// active-set polish (the north star's item 3; BASELINE.json).  It replaces QPWrapperOsqp + OSQP
// (src/qpwrapper_osqp.cpp:55-261) for the QPs the per-thread dual active-set solver (qp_gi.cuh) cannot take: the
// LP-dual formulations of ASIFrobust (nv = 402, nc = 300; src/asif_robust.cpp:21-22) and ASIFrealizable (nv = 38),
// whose Hessian is only positive SEMI-definite in the multipliers.
//
//      min v'Hv + c'v   s.t.  A v >= b (row i an equality where be[i]),  lb <= v <= ub
//
// is posed exactly as the wrapper poses it to OSQP (P = 2H, rows [A ; I] with l = [b ; lb], u = [b|inf ; ub],
// src/qpwrapper_osqp.cpp:262-375) and solved by the published OSQP algorithm (Stellato et al. 2020: Ruiz
// equilibration, ADMM with relaxation and per-row step sizes, adaptive rho, infeasibility certificates, polish),
// restated for ONE TEAM of threads per problem:
//   * the team is a thread-block cluster (8 CTAs x 512 threads on 8 SMs of one die; cluster.sync() is the phase
//     barrier, ~0.2 us) for large problems and one CTA for small ones; everything lives in an L2-resident global
//     workspace, so eight SMs' worth of L2 bandwidth serves one problem;
//   * nothing is sparse here: A is kept dense in both orientations (columns for A'w, rows for Av), every product is
//     "one warp per output element, lanes along the contiguous direction, shuffle reduction";
//   * the linear system of the x-update is NOT solved by substitution (2n dependent steps = 2n barriers per
//     iteration): K = P + sigma I + A'RA is factored once per rho (left-looking LDL', two barriers per column),
//     inverted explicitly (K^-1 = L^-T D^-1 L^-1, thread-per-column substitution + one product) and the iteration's
//     x-update is a dense symmetric matrix-vector product - three barriers per ADMM iteration in all;
//   * the polish solves OSQP's regularised KKT system of the guessed active set (dense LDL', substitution by one warp
//     over row- and column-major copies of L) with iterative refinement against the unregularised system.
// The code is written against a Team concept (thread rank, warp rank, warp reductions, barrier) so that the same text
// runs (a) on the device with ClusterTeam and (b) with a one-thread team on the host inside tests/hostemu/ -- a
// TEST-ONLY build that lets the CPU suite check the algorithm against the reference build; the product
// (libasif_b200.so) contains the device instantiation only and has no CPU path.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define QA_FN __host__ __device__ __forceinline__
#define QA_NOINLINE __host__ __device__ __noinline__
#else
#define QA_FN inline
#define QA_NOINLINE
#endif

namespace qpadmm {

constexpr double QA_INFTY = 1e30; // OSQP_INFTY
constexpr double RHO_MIN = 1e-6, RHO_MAX = 1e6, RHO_EQ_OVER_RHO_INEQ = 1e3, RHO_TOL = 1e-4;
constexpr double MIN_SCALING = 1e-4, MAX_SCALING = 1e4;
constexpr int NRED = 8;          // values per team reduction
constexpr int MAX_TEAM_WARPS = 256;

enum : int32_t {
	ST_SOLVED = 1,
	ST_SOLVED_INACCURATE = 2,
	ST_PRIMAL_INFEASIBLE_INACCURATE = 3,
	ST_DUAL_INFEASIBLE_INACCURATE = 4,
	ST_MAX_ITER = -2,
	ST_PRIMAL_INFEASIBLE = -3,
	ST_DUAL_INFEASIBLE = -4,
	ST_NON_CVX = -7,
	ST_UNSOLVED = -10
};

struct Settings {
	double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, delta, adaptive_rho_tolerance;
	int32_t scaling, max_iter, check_termination, adaptive_rho_interval, polish, polish_refine_iter;
};

// OSQP's defaults with the accuracy BASELINE.json defines parity at (eps 1e-8, polish on)
inline Settings default_settings()
{
	Settings s;
	s.rho = 0.1;
	s.sigma = 1e-6;
	s.alpha = 1.6;
	s.eps_abs = 1e-8;
	s.eps_rel = 1e-8;
	s.eps_prim_inf = 1e-4;
	s.eps_dual_inf = 1e-4;
	s.delta = 1e-6;
	s.adaptive_rho_tolerance = 5.0;
	s.scaling = 10;
	s.max_iter = 20000;
	s.check_termination = 25;
	s.adaptive_rho_interval = 50;
	s.polish = 1;
	s.polish_refine_iter = 10;
	return s;
}

// one problem in the reference wrapper's layout (include/asif_b200.h, asif_qp_solve_batch)
struct Problem {
	int32_t nv, nc, diag_cost;
	const double *H, *c, *A, *b, *lb, *ub;
	const uint8_t *be;
	double *sol;
	int32_t *status;
	int32_t *info; // optional [4]: ADMM iterations, rho updates, polish (1 accepted / -1 rejected / 0 not run), active rows
};

// ---- workspace ------------------------------------------------------------------------------------------------------
struct Work {
	// matrices (doubles)
	double *P, *A, *At, *K, *Xinv, *Kinv, *LR, *LC;
	// length-n vectors
	double *q, *D, *Dinv, *Ib, *x0, *x1, *xt, *dx, *rhs, *Dt, *pcol, *Px, *Aty, *xp, *kd;
	// length-m vectors
	double *l, *u, *E, *Einv, *rho_vec, *rho_inv, *z0, *z1, *y, *zt, *dy, *Et, *Ax, *tmpm, *yp, *zp;
	// length-N (= n + m) vectors of the polish
	double *prhs, *psol, *pres, *tcol, *pd;
	double *red; // 2 x NRED x MAX_TEAM_WARPS
	int32_t *ctype, *rows, *ints; // ints[0] = n_low, ints[1] = na, ints[2] = failure flag of the factorisation
};

QA_FN size_t work_doubles(int n, int mA)
{
	const size_t m = (size_t)mA + n, N = (size_t)n + m;
	size_t d = 0;
	d += (size_t)n * n * 4;        // P, K, Xinv, Kinv
	d += (size_t)mA * n * 2;       // A, At
	d += N * N * 2;                // LR, LC
	d += (size_t)n * 15 + m * 16 + N * 5;
	d += 2 * NRED * MAX_TEAM_WARPS;
	d += (2 * m + 8 + 1) / 2 + 2;  // ints
	return d + 256;                // alignment slack (every vector is rounded up to 16 bytes)
}

QA_FN Work carve(double *base, int n, int mA)
{
	const size_t m = (size_t)mA + n, N = (size_t)n + m;
	Work w;
	double *p = base;
	auto take = [&](size_t k) {
		double *r = p;
		p += (k + 1) & ~(size_t)1; // keep 16-byte alignment
		return r;
	};
	w.P = take((size_t)n * n);
	w.K = take((size_t)n * n);
	w.Xinv = take((size_t)n * n);
	w.Kinv = take((size_t)n * n);
	w.A = take((size_t)mA * n);
	w.At = take((size_t)mA * n);
	w.LR = take(N * N);
	w.LC = take(N * N);
	w.q = take(n); w.D = take(n); w.Dinv = take(n); w.Ib = take(n); w.x0 = take(n); w.x1 = take(n); w.xt = take(n);
	w.dx = take(n); w.rhs = take(n); w.Dt = take(n); w.pcol = take(n); w.Px = take(n); w.Aty = take(n); w.xp = take(n);
	w.kd = take(n);
	w.l = take(m); w.u = take(m); w.E = take(m); w.Einv = take(m); w.rho_vec = take(m); w.rho_inv = take(m);
	w.z0 = take(m); w.z1 = take(m); w.y = take(m); w.zt = take(m); w.dy = take(m); w.Et = take(m); w.Ax = take(m);
	w.tmpm = take(m); w.yp = take(m); w.zp = take(m);
	w.prhs = take(N); w.psol = take(N); w.pres = take(N); w.tcol = take(N); w.pd = take(N);
	w.red = take(2 * NRED * MAX_TEAM_WARPS);
	w.ctype = (int32_t *)p;
	w.rows = w.ctype + m;
	w.ints = w.rows + m;
	return w;
}

// ---- the solver -----------------------------------------------------------------------------------------------------
template <class T>
struct Solver {
	T &tm;
	const Settings st;
	const int n, mA, m;
	Work w;
	double cscale, cinv, rho;
	int red_slot;

	QA_FN Solver(T &team, const Settings &s, int nv, int nc, const Work &wk)
	    : tm(team), st(s), n(nv), mA(nc), m(nc + nv), w(wk), cscale(1.0), cinv(1.0), rho(s.rho), red_slot(0)
	{
	}

	// -- team reduction of NRED values; bit k of summask: sum, else max.  Every thread returns the same bits.
	QA_FN void reduce(double (&v)[NRED], unsigned summask)
	{
		double *slot = w.red + (size_t)red_slot * NRED * MAX_TEAM_WARPS;
		red_slot ^= 1;
#pragma unroll
		for (int k = 0; k < NRED; k++) {
			const double r = ((summask >> k) & 1u) ? tm.warp_sum(v[k]) : tm.warp_max(v[k]);
			if (tm.lane == 0) slot[k * MAX_TEAM_WARPS + tm.warp] = r;
		}
		tm.sync();
#pragma unroll
		for (int k = 0; k < NRED; k++) {
			const bool sum = (summask >> k) & 1u;
			double a = sum ? 0.0 : -INFINITY;
			for (int i = tm.lane; i < tm.nwarps; i += T::LANES) {
				const double e = slot[k * MAX_TEAM_WARPS + i];
				a = sum ? a + e : fmax(a, e);
			}
			v[k] = sum ? tm.warp_sum(a) : tm.warp_max(a);
		}
	}

	// -- products with the constraint matrix [A ; diag(Ib)] and with P: one warp per output element, no barrier inside
	QA_FN void mv_A(const double *v, double *out) const
	{
		for (int i = tm.warp; i < mA; i += tm.nwarps) {
			const double *row = w.At + (size_t)i * n;
			double s = 0.0;
			for (int j = tm.lane; j < n; j += T::LANES) s = fma(row[j], v[j], s);
			s = tm.warp_sum(s);
			if (tm.lane == 0) out[i] = s;
		}
		for (int j = tm.tid; j < n; j += tm.nthreads) out[mA + j] = w.Ib[j] * v[j];
	}
	QA_FN void mv_At(const double *y, double *out) const
	{
		for (int j = tm.warp; j < n; j += tm.nwarps) {
			const double *col = w.A + (size_t)j * mA;
			double s = 0.0;
			for (int i = tm.lane; i < mA; i += T::LANES) s = fma(col[i], y[i], s);
			s = tm.warp_sum(s);
			if (tm.lane == 0) out[j] = s + w.Ib[j] * y[mA + j];
		}
	}
	QA_FN void mv_P(const double *v, double *out) const
	{
		for (int j = tm.warp; j < n; j += tm.nwarps) {
			const double *col = w.P + (size_t)j * n;
			double s = 0.0;
			for (int i = tm.lane; i < n; i += T::LANES) s = fma(col[i], v[i], s);
			s = tm.warp_sum(s);
			if (tm.lane == 0) out[j] = s;
		}
	}

	static QA_FN double limit_scaling(double v)
	{
		v = v < MIN_SCALING ? 1.0 : v;
		return v > MAX_SCALING ? MAX_SCALING : v;
	}

	// -- load the problem in OSQP form and equilibrate it (OSQP scaling.c::scale_data)
	QA_FN void load_and_scale(const Problem &pb)
	{
		for (size_t e = tm.tid; e < (size_t)n * n; e += tm.nthreads) {
			const int i = (int)(e % n), j = (int)(e / n);
			double v = 0.0;
			if (pb.diag_cost) v = (i == j) ? 2.0 * pb.H[e] : 0.0;
			else v = 2.0 * pb.H[e];
			w.P[e] = v;
		}
		for (size_t e = tm.tid; e < (size_t)mA * n; e += tm.nthreads) {
			const int i = (int)(e % mA), j = (int)(e / mA);
			const double v = pb.A[e];
			w.A[e] = v;
			w.At[(size_t)i * n + j] = v;
		}
		for (int j = tm.tid; j < n; j += tm.nthreads) {
			w.q[j] = pb.c[j];
			w.D[j] = 1.0;
			w.Ib[j] = 1.0;
			w.l[mA + j] = pb.lb[j];
			w.u[mA + j] = pb.ub[j];
			w.E[mA + j] = 1.0;
		}
		for (int i = tm.tid; i < mA; i += tm.nthreads) {
			w.l[i] = pb.b[i];
			w.u[i] = (pb.be != nullptr && pb.be[i]) ? pb.b[i] : QA_INFTY;
			w.E[i] = 1.0;
		}
		cscale = 1.0;
		tm.sync();
		for (int it = 0; it < st.scaling; it++) {
			// column norms of [P ; A ; I] and row norms of [A ; I], both from the matrices as they stand
			for (int j = tm.warp; j < n; j += tm.nwarps) {
				const double *pc = w.P + (size_t)j * n, *ac = w.A + (size_t)j * mA;
				double v = 0.0;
				for (int i = tm.lane; i < n; i += T::LANES) v = fmax(v, fabs(pc[i]));
				for (int i = tm.lane; i < mA; i += T::LANES) v = fmax(v, fabs(ac[i]));
				v = fmax(tm.warp_max(v), fabs(w.Ib[j]));
				if (tm.lane == 0) {
					w.Dt[j] = 1.0 / sqrt(limit_scaling(v));
					w.Et[mA + j] = 1.0 / sqrt(limit_scaling(fabs(w.Ib[j])));
				}
			}
			for (int i = tm.warp; i < mA; i += tm.nwarps) {
				const double *ar = w.At + (size_t)i * n;
				double v = 0.0;
				for (int j = tm.lane; j < n; j += T::LANES) v = fmax(v, fabs(ar[j]));
				v = tm.warp_max(v);
				if (tm.lane == 0) w.Et[i] = 1.0 / sqrt(limit_scaling(v));
			}
			tm.sync();
			for (size_t e = tm.tid; e < (size_t)n * n; e += tm.nthreads) w.P[e] *= w.Dt[e % n] * w.Dt[e / n];
			for (size_t e = tm.tid; e < (size_t)mA * n; e += tm.nthreads) {
				w.A[e] *= w.Et[e % mA] * w.Dt[e / mA];
				w.At[e] *= w.Et[e / n] * w.Dt[e % n];
			}
			for (int j = tm.tid; j < n; j += tm.nthreads) {
				w.Ib[j] *= w.Et[mA + j] * w.Dt[j];
				w.q[j] *= w.Dt[j];
				w.D[j] *= w.Dt[j];
			}
			for (int i = tm.tid; i < m; i += tm.nthreads) w.E[i] *= w.Et[i];
			tm.sync();
			// cost normalisation: mean column norm of P against |q|_inf
			double red[NRED] = {0, 0, 0, 0, 0, 0, 0, 0};
			for (int j = tm.warp; j < n; j += tm.nwarps) {
				const double *pc = w.P + (size_t)j * n;
				double v = 0.0;
				for (int i = tm.lane; i < n; i += T::LANES) v = fmax(v, fabs(pc[i]));
				v = tm.warp_max(v);
				if (tm.lane == 0) red[0] += v;
			}
			for (int j = tm.tid; j < n; j += tm.nthreads) red[1] = fmax(red[1], fabs(w.q[j]));
			reduce(red, 1u);
			double ct = fmax(red[0] / n, limit_scaling(red[1]));
			ct = 1.0 / limit_scaling(ct);
			for (size_t e = tm.tid; e < (size_t)n * n; e += tm.nthreads) w.P[e] *= ct;
			for (int j = tm.tid; j < n; j += tm.nthreads) w.q[j] *= ct;
			cscale *= ct;
			tm.sync();
		}
		cinv = 1.0 / cscale;
		for (int j = tm.tid; j < n; j += tm.nthreads) w.Dinv[j] = 1.0 / w.D[j];
		for (int i = tm.tid; i < m; i += tm.nthreads) {
			w.Einv[i] = 1.0 / w.E[i];
			w.l[i] *= w.E[i];
			w.u[i] *= w.E[i];
		}
		tm.sync();
	}

	// -- per-row step sizes (OSQP auxil.c::set_rho_vec)
	QA_FN void set_rho_vec()
	{
		rho = fmin(fmax(rho, RHO_MIN), RHO_MAX);
		for (int i = tm.tid; i < m; i += tm.nthreads) {
			double r;
			int t;
			if (w.l[i] < -QA_INFTY * MIN_SCALING && w.u[i] > QA_INFTY * MIN_SCALING) {
				t = -1;
				r = RHO_MIN;
			} else if (w.u[i] - w.l[i] < RHO_TOL) {
				t = 1;
				r = RHO_EQ_OVER_RHO_INEQ * rho;
			} else {
				t = 0;
				r = rho;
			}
			w.ctype[i] = t;
			w.rho_vec[i] = r;
			w.rho_inv[i] = 1.0 / r;
		}
		tm.sync();
	}

	// -- in-place LDL' of the symmetric N x N matrix held (fully) in LRm (row-major == column-major); on return the
	//    strict lower triangle of LRm holds L by rows, LCm (if given) holds L by columns, d the pivots.  Left-looking:
	//    column j needs one sweep of dot products over the finished columns - two barriers per column.
	//    Returns false (on every thread) when a pivot is zero or not finite.
	QA_FN bool ldl_factor(double *LRm, double *LCm, double *d, double *tcol, const int N, const bool want_positive)
	{
		bool ok = true;
		for (int j = 0; j < N; j++) {
			const double *rj = LRm + (size_t)j * N;
			for (int i = j + tm.warp; i < N; i += tm.nwarps) {
				const double *ri = LRm + (size_t)i * N;
				double s = 0.0;
				for (int k = tm.lane; k < j; k += T::LANES) s = fma(ri[k] * d[k], rj[k], s);
				s = tm.warp_sum(s);
				if (tm.lane == 0) tcol[i] = ri[j] - s;
			}
			tm.sync();
			const double dj = tcol[j];
			if (!(fabs(dj) > 0.0) || !(fabs(dj) < INFINITY) || (want_positive && dj < 0.0)) ok = false;
			const double dinv = 1.0 / dj;
			for (int i = j + 1 + tm.tid; i < N; i += tm.nthreads) {
				const double lij = tcol[i] * dinv;
				LRm[(size_t)i * N + j] = lij;
				if (LCm) LCm[(size_t)j * N + i] = lij;
			}
			if (tm.tid == 0) d[j] = dj;
			tm.sync();
			if (!ok) break; // dj is the same value on every thread
		}
		return ok;
	}

	// -- solve L D L' s = b in place by ONE warp (rows of L for the forward sweep, columns for the backward one)
	QA_FN void ldl_solve_warp0(const double *LRm, const double *LCm, const double *d, double *b, const int N)
	{
		if (tm.warp == 0) {
			for (int i = 0; i < N; i++) {
				const double *ri = LRm + (size_t)i * N;
				double s = 0.0;
				for (int k = tm.lane; k < i; k += T::LANES) s = fma(ri[k], b[k], s);
				s = tm.warp_sum(s);
				if (tm.lane == 0) b[i] -= s;
				tm.warp_sync();
			}
			for (int i = tm.lane; i < N; i += T::LANES) b[i] /= d[i];
			tm.warp_sync();
			for (int i = N - 1; i >= 0; i--) {
				const double *ci = LCm + (size_t)i * N;
				double s = 0.0;
				for (int k = i + 1 + tm.lane; k < N; k += T::LANES) s = fma(ci[k], b[k], s);
				s = tm.warp_sum(s);
				if (tm.lane == 0) b[i] -= s;
				tm.warp_sync();
			}
		}
		tm.sync();
	}

	// -- K = P + sigma I + [A;I]' R [A;I], factored and inverted explicitly; false when K is not positive definite
	QA_FN bool factor()
	{
		for (size_t p = tm.warp; p < (size_t)n * n; p += tm.nwarps) {
			const int a = (int)(p / n), b = (int)(p % n);
			if (b > a) continue;
			const double *ca = w.A + (size_t)a * mA, *cb = w.A + (size_t)b * mA;
			double s = 0.0;
			for (int i = tm.lane; i < mA; i += T::LANES) s = fma(ca[i] * w.rho_vec[i], cb[i], s);
			s = tm.warp_sum(s);
			if (tm.lane == 0) {
				double v = w.P[(size_t)a * n + b] + s;
				if (a == b) v += st.sigma + w.Ib[a] * w.Ib[a] * w.rho_vec[mA + a];
				w.K[(size_t)a * n + b] = v;
				w.K[(size_t)b * n + a] = v;
			}
		}
		for (size_t e = tm.tid; e < (size_t)n * n; e += tm.nthreads) w.Xinv[e] = 0.0;
		tm.sync();
		if (!ldl_factor(w.K, nullptr, w.kd, w.tcol, n, true)) return false;
		for (int i = tm.tid; i < n; i += tm.nthreads) w.pcol[i] = 1.0 / w.kd[i];
		// X = L^-1 by rows (X[i][j] at Xinv[i*n + j]): thread j owns column j; the lanes of a warp walk the same
		// (i, k) so that L[i][k] is one broadcast load and X[k][j] a coalesced one
		for (int j = tm.tid; j < n; j += tm.nthreads) {
			const int k0 = j - (j % T::LANES);
			for (int i = k0; i < n; i++) {
				const double *li = w.K + (size_t)i * n;
				double s0 = 0.0, s1 = 0.0;
				int k = k0;
				for (; k + 1 < i; k += 2) {
					s0 = fma(li[k], w.Xinv[(size_t)k * n + j], s0);
					s1 = fma(li[k + 1], w.Xinv[(size_t)(k + 1) * n + j], s1);
				}
				if (k < i) s0 = fma(li[k], w.Xinv[(size_t)k * n + j], s0);
				w.Xinv[(size_t)i * n + j] = (i == j ? 1.0 : 0.0) - (s0 + s1);
			}
		}
		tm.sync();
		// K^-1[a][b] = sum_{i >= max(a,b)} X[i][a] X[i][b] / d_i
		for (size_t e = tm.tid; e < (size_t)n * n; e += tm.nthreads) {
			const int a = (int)(e / n), b = (int)(e % n);
			double s = 0.0;
			for (int i = (a > b ? a : b); i < n; i++) s = fma(w.Xinv[(size_t)i * n + a] * w.pcol[i], w.Xinv[(size_t)i * n + b], s);
			w.Kinv[e] = s;
		}
		tm.sync();
		return true;
	}

	struct Res {
		double pri, dua, eps_pri_norm, eps_dua_norm, obj;
	};

	// -- unscaled residuals and the norms of the tolerances (OSQP auxil.c::compute_*_res / compute_*_tol)
	QA_FN Res residuals(const double *xx, const double *zz, const double *yy)
	{
		mv_A(xx, w.Ax);
		mv_P(xx, w.Px);
		mv_At(yy, w.Aty);
		tm.sync();
		double r[NRED] = {0, 0, 0, 0, 0, 0, 0, 0};
		for (int i = tm.tid; i < m; i += tm.nthreads) {
			r[0] = fmax(r[0], fabs(w.Einv[i] * (w.Ax[i] - zz[i])));
			r[1] = fmax(r[1], fabs(w.Einv[i] * w.Ax[i]));
			r[2] = fmax(r[2], fabs(w.Einv[i] * zz[i]));
		}
		for (int j = tm.tid; j < n; j += tm.nthreads) {
			r[3] = fmax(r[3], fabs(w.Dinv[j] * (w.Px[j] + w.q[j] + w.Aty[j])));
			r[4] = fmax(r[4], fabs(w.Dinv[j] * w.Px[j]));
			r[5] = fmax(r[5], fabs(w.Dinv[j] * w.Aty[j]));
			r[6] = fmax(r[6], fabs(w.Dinv[j] * w.q[j]));
			r[7] += xx[j] * (0.5 * w.Px[j] + w.q[j]);
		}
		reduce(r, 1u << 7);
		Res o;
		o.pri = r[0];
		o.dua = cinv * r[3];
		o.eps_pri_norm = fmax(r[1], r[2]);
		o.eps_dua_norm = cinv * fmax(fmax(r[4], r[5]), r[6]);
		o.obj = cinv * r[7];
		return o;
	}

	QA_FN bool primal_infeasible(const double eps)
	{
		const double infval = QA_INFTY * MIN_SCALING;
		double r[NRED] = {0, 0, 0, 0, 0, 0, 0, 0};
		for (int i = tm.tid; i < m; i += tm.nthreads) {
			double d = w.dy[i];
			if (w.u[i] > infval) {
				if (w.l[i] < -infval) d = 0.0;
				else d = fmin(d, 0.0);
			} else if (w.l[i] < -infval) {
				d = fmax(d, 0.0);
			}
			w.tmpm[i] = d;
			r[0] = fmax(r[0], fabs(w.E[i] * d));
			r[1] += w.u[i] * fmax(d, 0.0) + w.l[i] * fmin(d, 0.0);
		}
		reduce(r, 2u);
		const double nd = r[0], lhs = r[1];
		if (!(nd > eps) || !(lhs < -eps * nd)) return false;
		mv_At(w.tmpm, w.Aty);
		tm.sync();
		double r2[NRED] = {0, 0, 0, 0, 0, 0, 0, 0};
		for (int j = tm.tid; j < n; j += tm.nthreads) r2[0] = fmax(r2[0], fabs(w.Dinv[j] * w.Aty[j]));
		reduce(r2, 0u);
		return r2[0] < eps * nd;
	}

	QA_FN bool dual_infeasible(const double eps)
	{
		double r[NRED] = {0, 0, 0, 0, 0, 0, 0, 0};
		for (int j = tm.tid; j < n; j += tm.nthreads) {
			r[0] = fmax(r[0], fabs(w.D[j] * w.dx[j]));
			r[1] += w.q[j] * w.dx[j];
		}
		reduce(r, 2u);
		const double ndx = r[0], qdx = r[1];
		if (!(ndx > eps) || !(qdx < -cscale * eps * ndx)) return false;
		mv_P(w.dx, w.Px);
		mv_A(w.dx, w.Ax);
		tm.sync();
		const double infval = QA_INFTY * MIN_SCALING;
		double r2[NRED] = {0, 0, 0, 0, 0, 0, 0, 0};
		for (int j = tm.tid; j < n; j += tm.nthreads) r2[0] = fmax(r2[0], fabs(w.Dinv[j] * w.Px[j]));
		for (int i = tm.tid; i < m; i += tm.nthreads) {
			const double a = w.Einv[i] * w.Ax[i];
			if ((w.u[i] < infval && a > eps * ndx) || (w.l[i] > -infval && a < -eps * ndx)) r2[1] = 1.0;
		}
		reduce(r2, 0u);
		return r2[0] < cscale * eps * ndx && r2[1] == 0.0;
	}

	// -- OSQP auxil.c::check_termination
	QA_FN int32_t check(const Res &r, const bool approximate)
	{
		double ea = st.eps_abs, er = st.eps_rel, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
		if (approximate) {
			ea *= 10;
			er *= 10;
			epi *= 10;
			edi *= 10;
		}
		bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
		if (r.pri < ea + er * r.eps_pri_norm) prim_ok = true;
		else prim_inf = primal_infeasible(epi);
		if (r.dua < ea + er * r.eps_dua_norm) dual_ok = true;
		else dual_inf = dual_infeasible(edi);
		if (prim_ok && dual_ok) return approximate ? ST_SOLVED_INACCURATE : ST_SOLVED;
		if (prim_inf) return approximate ? ST_PRIMAL_INFEASIBLE_INACCURATE : ST_PRIMAL_INFEASIBLE;
		if (dual_inf) return approximate ? ST_DUAL_INFEASIBLE_INACCURATE : ST_DUAL_INFEASIBLE;
		return ST_UNSOLVED;
	}

	// -- OSQP auxil.c::compute_rho_estimate (scaled quantities)
	QA_FN double rho_estimate(const double *x, const double *z, const double *y)
	{
		mv_A(x, w.Ax);
		mv_P(x, w.Px);
		mv_At(y, w.Aty);
		tm.sync();
		double r[NRED] = {0, 0, 0, 0, 0, 0, 0, 0};
		for (int i = tm.tid; i < m; i += tm.nthreads) {
			r[0] = fmax(r[0], fabs(w.Ax[i] - z[i]));
			r[1] = fmax(r[1], fabs(z[i]));
			r[2] = fmax(r[2], fabs(w.Ax[i]));
		}
		for (int j = tm.tid; j < n; j += tm.nthreads) {
			r[3] = fmax(r[3], fabs(w.Px[j] + w.q[j] + w.Aty[j]));
			r[4] = fmax(r[4], fabs(w.q[j]));
			r[5] = fmax(r[5], fabs(w.Aty[j]));
			r[6] = fmax(r[6], fabs(w.Px[j]));
		}
		reduce(r, 0u);
		const double pri = r[0] / (fmax(r[1], r[2]) + 1e-10);
		const double dua = r[3] / (fmax(fmax(r[4], r[5]), r[6]) + 1e-10);
		const double e = rho * sqrt(pri / (dua + 1e-10));
		return fmin(fmax(e, RHO_MIN), RHO_MAX);
	}

	// -- OSQP polish.c::polish on the iterate (x, z, y); true when the polished point replaced it (then in xp, zp, yp)
	QA_FN bool polish(const double *x, const double *z, const double *y, Res &info, int32_t &n_active)
	{
		if (tm.tid == 0) { // the two ordered row lists: a few hundred tests, not worth a scan
			int na = 0;
			for (int i = 0; i < m; i++)
				if (z[i] - w.l[i] < -y[i]) w.rows[na++] = i;
			w.ints[0] = na;
			for (int i = 0; i < m; i++)
				if (w.u[i] - z[i] < y[i]) w.rows[na++] = i;
			w.ints[1] = na;
		}
		tm.sync();
		const int n_low = w.ints[0], na = w.ints[1], N = n + na;
		n_active = na;
		// regularised KKT matrix [P + delta I, A_act' ; A_act, -delta I], full, in LR
		for (size_t e = tm.tid; e < (size_t)N * N; e += tm.nthreads) {
			int i = (int)(e / N), j = (int)(e % N);
			if (i < j) {
				const int t = i;
				i = j;
				j = t;
			}
			double v;
			if (i < n) v = w.P[(size_t)i * n + j] + (i == j ? st.delta : 0.0);
			else if (j < n) {
				const int row = w.rows[i - n];
				v = row < mA ? w.At[(size_t)row * n + j] : (row - mA == j ? w.Ib[j] : 0.0);
			} else v = (i == j) ? -st.delta : 0.0;
			w.LR[e] = v;
		}
		for (int j = tm.tid; j < n; j += tm.nthreads) w.prhs[j] = -w.q[j];
		for (int a = tm.tid; a < na; a += tm.nthreads) w.prhs[n + a] = a < n_low ? w.l[w.rows[a]] : w.u[w.rows[a]];
		tm.sync();
		if (!ldl_factor(w.LR, w.LC, w.pd, w.tcol, N, false)) return false;
		for (int i = tm.tid; i < N; i += tm.nthreads) w.psol[i] = w.prhs[i];
		tm.sync();
		ldl_solve_warp0(w.LR, w.LC, w.pd, w.psol, N);
		for (int it = 0; it < st.polish_refine_iter; it++) {
			// residual of the UNregularised system: [P x + A_act' y ; A_act x]
			for (int i = tm.tid; i < m; i += tm.nthreads) w.tmpm[i] = 0.0;
			tm.sync();
			if (tm.tid == 0)
				for (int a = 0; a < na; a++) w.tmpm[w.rows[a]] += w.psol[n + a];
			tm.sync();
			mv_P(w.psol, w.Px);
			mv_At(w.tmpm, w.Aty);
			mv_A(w.psol, w.Ax);
			tm.sync();
			for (int j = tm.tid; j < n; j += tm.nthreads) w.pres[j] = w.prhs[j] - w.Px[j] - w.Aty[j];
			for (int a = tm.tid; a < na; a += tm.nthreads) w.pres[n + a] = w.prhs[n + a] - w.Ax[w.rows[a]];
			tm.sync();
			ldl_solve_warp0(w.LR, w.LC, w.pd, w.pres, N);
			double r[NRED] = {0, 0, 0, 0, 0, 0, 0, 0};
			for (int i = tm.tid; i < N; i += tm.nthreads) {
				const double s = w.psol[i] + w.pres[i];
				r[0] = fmax(r[0], fabs(w.pres[i]));
				r[1] = fmax(r[1], fabs(s));
				w.psol[i] = s;
			}
			reduce(r, 0u);
			if (!(r[0] > 1e-15 * r[1])) break; // the correction no longer changes the solution
		}
		double bad[NRED] = {0, 0, 0, 0, 0, 0, 0, 0};
		for (int j = tm.tid; j < n; j += tm.nthreads) {
			w.xp[j] = w.psol[j];
			if (w.psol[j] != w.psol[j]) bad[0] = 1.0;
		}
		for (int i = tm.tid; i < m; i += tm.nthreads) w.yp[i] = 0.0;
		reduce(bad, 0u);
		if (bad[0] != 0.0) return false;
		if (tm.tid == 0)
			for (int a = 0; a < na; a++) w.yp[w.rows[a]] = w.psol[n + a]; // a later (upper) entry wins, as OSQP's get_ypol_from_yred
		mv_A(w.xp, w.zp);
		tm.sync();
		for (int i = tm.tid; i < m; i += tm.nthreads) w.zp[i] = fmin(fmax(w.zp[i], w.l[i]), w.u[i]);
		tm.sync();
		const Res pr = residuals(w.xp, w.zp, w.yp);
		const bool ok = (pr.pri < info.pri && pr.dua < info.dua) || (pr.pri < info.pri && info.dua < 1e-10) ||
		                (pr.dua < info.dua && info.pri < 1e-10);
		if (ok) info = pr;
		return ok;
	}

	// -- the whole solve: status code, solution written to pb.sol (NaN when there is none, as OSQP)
	QA_FN int32_t solve(const Problem &pb)
	{
		load_and_scale(pb);
		rho = st.rho;
		set_rho_vec();
		int32_t status = ST_UNSOLVED, iter = 0, rho_updates = 0, polish_state = 0, n_active = 0;
		const double *xf = nullptr;
		if (!factor()) {
			status = ST_NON_CVX;
		} else {
			double *x = w.x0, *xprev = w.x1, *z = w.z0, *zprev = w.z1;
			for (int j = tm.tid; j < n; j += tm.nthreads) x[j] = xprev[j] = w.dx[j] = 0.0;
			for (int i = tm.tid; i < m; i += tm.nthreads) z[i] = zprev[i] = w.y[i] = w.dy[i] = 0.0;
			tm.sync();
			const double alpha = st.alpha, sigma = st.sigma;
			Res r = {0, 0, 0, 0, 0};
			bool checked = false;
			for (iter = 1; iter <= st.max_iter; iter++) {
				{ // the iterate of the last pass becomes "prev"
					double *t = x;
					x = xprev;
					xprev = t;
					t = z;
					z = zprev;
					zprev = t;
				}
				// rhs = sigma x_prev - q + [A;I]'(R z_prev - y)
				for (int j = tm.warp; j < n; j += tm.nwarps) {
					const double *col = w.A + (size_t)j * mA;
					double s = 0.0;
					for (int i = tm.lane; i < mA; i += T::LANES) s = fma(col[i], w.rho_vec[i] * zprev[i] - w.y[i], s);
					s = tm.warp_sum(s);
					if (tm.lane == 0)
						w.rhs[j] = s + w.Ib[j] * (w.rho_vec[mA + j] * zprev[mA + j] - w.y[mA + j]) + sigma * xprev[j] - w.q[j];
				}
				tm.sync();
				// x-tilde = K^-1 rhs
				for (int j = tm.warp; j < n; j += tm.nwarps) {
					const double *row = w.Kinv + (size_t)j * n;
					double s = 0.0;
					for (int k = tm.lane; k < n; k += T::LANES) s = fma(row[k], w.rhs[k], s);
					s = tm.warp_sum(s);
					if (tm.lane == 0) w.xt[j] = s;
				}
				tm.sync();
				// z-tilde = [A;I] x-tilde, then the relaxed x, the projected z and the dual step, row by row
				for (int i = tm.warp; i < mA; i += tm.nwarps) {
					const double *row = w.At + (size_t)i * n;
					double s = 0.0;
					for (int j = tm.lane; j < n; j += T::LANES) s = fma(row[j], w.xt[j], s);
					s = tm.warp_sum(s);
					if (tm.lane == 0) {
						const double zr = alpha * s + (1.0 - alpha) * zprev[i];
						double zi = zr + w.y[i] * w.rho_inv[i];
						zi = fmin(fmax(zi, w.l[i]), w.u[i]);
						z[i] = zi;
						const double d = w.rho_vec[i] * (zr - zi);
						w.dy[i] = d;
						w.y[i] += d;
					}
				}
				for (int j = tm.tid; j < n; j += tm.nthreads) {
					const double xtj = w.xt[j];
					const double xn = alpha * xtj + (1.0 - alpha) * xprev[j];
					x[j] = xn;
					w.dx[j] = xn - xprev[j];
					const int i = mA + j;
					const double zr = alpha * (w.Ib[j] * xtj) + (1.0 - alpha) * zprev[i];
					double zi = zr + w.y[i] * w.rho_inv[i];
					zi = fmin(fmax(zi, w.l[i]), w.u[i]);
					z[i] = zi;
					const double d = w.rho_vec[i] * (zr - zi);
					w.dy[i] = d;
					w.y[i] += d;
				}
				tm.sync();
				checked = st.check_termination > 0 && (iter % st.check_termination == 0);
				if (checked) {
					r = residuals(x, z, w.y);
					status = check(r, false);
					if (status != ST_UNSOLVED) break;
				}
				if (st.adaptive_rho_interval > 0 && (iter % st.adaptive_rho_interval == 0)) {
					const double rn = rho_estimate(x, z, w.y);
					if (rn > rho * st.adaptive_rho_tolerance || rn < rho / st.adaptive_rho_tolerance) {
						rho = rn;
						set_rho_vec();
						rho_updates++;
						if (!factor()) {
							status = ST_NON_CVX;
							break;
						}
					}
				}
			}
			if (iter > st.max_iter) iter = st.max_iter;
			if (status != ST_NON_CVX) {
				if (!checked) {
					r = residuals(x, z, w.y);
					status = check(r, false);
				}
				if (status == ST_UNSOLVED) {
					status = check(r, true);
					if (status == ST_UNSOLVED) status = ST_MAX_ITER;
				}
				xf = x;
				if (st.polish && status == ST_SOLVED) {
					const bool ok = polish(x, z, w.y, r, n_active);
					polish_state = ok ? 1 : -1;
					if (ok) xf = w.xp;
				}
			}
		}
		const bool has_sol = (status == ST_SOLVED || status == ST_SOLVED_INACCURATE || status == ST_MAX_ITER);
		for (int j = tm.tid; j < n; j += tm.nthreads) pb.sol[j] = has_sol ? w.D[j] * xf[j] : NAN;
		if (tm.tid == 0) {
			*pb.status = status;
			if (pb.info) {
				pb.info[0] = iter;
				pb.info[1] = rho_updates;
				pb.info[2] = polish_state;
				pb.info[3] = n_active;
			}
		}
		tm.sync();
		return status;
	}
};

} // namespace qpadmm
