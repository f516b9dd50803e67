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

// wall clock for the phase statistics (ns); 0 on the host
QA_FN unsigned long long now_ns()
{
#if defined(__CUDA_ARCH__)
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
#else
	return 0ull;
#endif
}
constexpr int NINFO = 8; // per problem: iterations, rho updates, polish state, active rows, us in: scaling, factorisations, iterations, polish

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
	int32_t *info; // optional [NINFO]: ADMM iterations, rho updates, polish (1 accepted / -1 rejected / 0 not run), active rows, then us per phase
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
	double *red; // 2 x NRED x red_stride (>= the team's warps)
	int32_t red_stride;
	int32_t *ctype, *rows, *ints; // ints[0] = n_low, ints[1] = na, ints[2] = failure flag of the factorisation
};

// The workspace has two parts: the polish's KKT factor (LR, LC: 2 (2n + nc)^2 doubles), which always lives in global
// memory, and everything else ("near" part: P, A in both orientations, K, its inverse, every vector), which a
// one-CTA team keeps in SHARED memory when it fits (the 38-variable problem: 117 KB) - the phases of a small problem are
// chains of dependent loads, and the vectors one phase writes are exactly what the next one reads.
QA_FN size_t work_far_doubles(int n, int mA)
{
	const size_t N = (size_t)2 * n + mA;
	return 2 * N * N + 8;
}
QA_FN size_t work_near_doubles(int n, int mA, int team_warps)
{
	const size_t m = (size_t)mA + n, N = (size_t)n + m;
	size_t d = 0;
	d += (size_t)n * n * 4;                // P, K, Xinv, Kinv
	d += (size_t)mA * n * 2;               // A, At
	d += (size_t)n * 15 + m * 16 + N * 8;  // tcol holds N x PB (= 4) doubles
	d += 2 * NRED * (size_t)team_warps;
	d += (2 * m + 8 + 1) / 2 + 2;          // ints
	return d + 128;                        // every array is rounded up to 16 bytes
}
QA_FN size_t work_doubles(int n, int mA) { return work_near_doubles(n, mA, MAX_TEAM_WARPS) + work_far_doubles(n, mA); }

QA_FN Work carve(double *near_base, double *far_base, int n, int mA, int team_warps)
{
	const size_t m = (size_t)mA + n, N = (size_t)n + m;
	Work w;
	double *p = near_base;
	auto take = [&](size_t k) {
		double *r = p;
		p += (k + 1) & ~(size_t)1; // keep 16-byte alignment
		return r;
	};
	w.LR = far_base;
	w.LC = far_base + ((N * N + 1) & ~(size_t)1);
	w.P = take((size_t)n * n);
	w.K = take((size_t)n * n);
	w.Xinv = take((size_t)n * n);
	w.Kinv = take((size_t)n * n);
	w.A = take((size_t)mA * n);
	w.At = take((size_t)mA * n);
	w.q = take(n); w.D = take(n); w.Dinv = take(n); w.Ib = take(n); w.x0 = take(n); w.x1 = take(n); w.xt = take(n);
	w.dx = take(n); w.rhs = take(n); w.Dt = take(n); w.pcol = take(n); w.Px = take(n); w.Aty = take(n); w.xp = take(n);
	w.kd = take(n);
	w.l = take(m); w.u = take(m); w.E = take(m); w.Einv = take(m); w.rho_vec = take(m); w.rho_inv = take(m);
	w.z0 = take(m); w.z1 = take(m); w.y = take(m); w.zt = take(m); w.dy = take(m); w.Et = take(m); w.Ax = take(m);
	w.tmpm = take(m); w.yp = take(m); w.zp = take(m);
	w.prhs = take(N); w.psol = take(N); w.pres = take(N); w.tcol = take(N * 4); w.pd = take(N);
	w.red = take(2 * NRED * (size_t)team_warps);
	w.red_stride = team_warps;
	w.ctype = (int32_t *)p;
	w.rows = w.ctype + m;
	w.ints = w.rows + m;
	return w;
}
// one block of global memory for both parts
QA_FN Work carve(double *base, int n, int mA)
{
	return carve(base + ((work_far_doubles(n, mA) + 1) & ~(size_t)1), base, n, mA, MAX_TEAM_WARPS);
}

// ---- the solver -----------------------------------------------------------------------------------------------------
template <class T>
struct Solver {
	T &tm;
	const Settings &st; // device: the kernel's __grid_constant__ parameter (constant bank)
	const int n, mA, m;
	const Work &w;      // device: one copy per CTA in shared memory - 60 pointers do not fit in registers, and left to
	                    // itself the compiler re-derives them from the base address at every use (a third of all
	                    // executed instructions in the first ncu capture)
	double cscale, cinv, rho;
	int red_slot;

	QA_FN Solver(T &team, const Settings &s, int nv, int nc, const Work &wk)
	    : tm(team), st(s), n(nv), mA(nc), m(nc + nv), w(wk), cscale(1.0), cinv(1.0), rho(s.rho), red_slot(0)
	{
	}

	// -- team reduction of NRED values; bit k of summask: sum, else max.  Every thread returns the same bits.
	QA_FN void reduce(double (&v)[NRED], unsigned summask)
	{
		const int rs = w.red_stride;
		double *slot = w.red + (size_t)red_slot * NRED * rs;
		red_slot ^= 1;
#pragma unroll
		for (int k = 0; k < NRED; k++) {
			const double r = ((summask >> k) & 1u) ? tm.warp_sum(v[k]) : tm.warp_max(v[k]);
			if (tm.lane == 0) slot[k * rs + tm.warp] = r;
		}
		tm.sync();
#pragma unroll
		for (int k = 0; k < NRED; k++) {
			const bool sum = (summask >> k) & 1u;
			double a = sum ? 0.0 : -INFINITY;
			for (int i = tm.lane; i < tm.nwarps; i += T::LANES) {
				const double e = slot[k * rs + i];
				a = sum ? a + e : fmax(a, e);
			}
			v[k] = sum ? tm.warp_sum(a) : tm.warp_max(a);
		}
	}

	// -- `count` dot products of length `len` against one vector: a warp takes R rows at a time and four steps of the
	//    lane loop at once, so that 4R + 4 independent loads are in flight per lane (the phases are bound by the latency
	//    of L2 - every cluster barrier invalidates L1 - not by bandwidth); rowptr(i) -> first element of row i,
	//    store(i, sum) runs on lane 0
	template <int R, class RowPtr, class Store>
	QA_FN void warp_dots(const int count, const int len, RowPtr rowptr, const double *vec, Store store) const
	{
		constexpr int L = T::LANES;
		for (int base = tm.warp * R; base < count; base += tm.nwarps * R) {
			const double *rp[R];
			double s[R];
#pragma unroll
			for (int r = 0; r < R; r++) {
				rp[r] = rowptr(base + r < count ? base + r : count - 1);
				s[r] = 0.0;
			}
			int k = tm.lane;
			for (; k + 3 * L < len; k += 4 * L) {
				const double v0 = vec[k], v1 = vec[k + L], v2 = vec[k + 2 * L], v3 = vec[k + 3 * L];
				double a[R][4];
#pragma unroll
				for (int r = 0; r < R; r++) {
					a[r][0] = rp[r][k];
					a[r][1] = rp[r][k + L];
					a[r][2] = rp[r][k + 2 * L];
					a[r][3] = rp[r][k + 3 * L];
				}
#pragma unroll
				for (int r = 0; r < R; r++) s[r] = fma(a[r][3], v3, fma(a[r][2], v2, fma(a[r][1], v1, fma(a[r][0], v0, s[r]))));
			}
			for (; k < len; k += L) {
				const double v0 = vec[k];
#pragma unroll
				for (int r = 0; r < R; r++) s[r] = fma(rp[r][k], v0, s[r]);
			}
#pragma unroll
			for (int r = 0; r < R; r++) s[r] = tm.warp_sum(s[r]);
			if (tm.lane == 0) {
#pragma unroll
				for (int r = 0; r < R; r++)
					if (base + r < count) store(base + r, s[r]);
			}
		}
	}

	// -- products with the constraint matrix [A ; diag(Ib)] and with P, no barrier inside
	QA_FN void mv_A(const double *v, double *out) const
	{
		warp_dots<4>(mA, n, [&](int i) { return w.At + (size_t)i * n; }, v, [&](int i, double s) { out[i] = s; });
		for (int j = tm.tid; j < n; j += tm.nthreads) out[mA + j] = w.Ib[j] * v[j];
	}
	QA_FN void mv_At(const double *y, double *out) const
	{
		warp_dots<4>(n, mA, [&](int j) { return w.A + (size_t)j * mA; }, y, [&](int j, double s) { out[j] = s + w.Ib[j] * y[mA + j]; });
	}
	QA_FN void mv_P(const double *v, double *out) const
	{
		warp_dots<4>(n, n, [&](int j) { return w.P + (size_t)j * n; }, v, [&](int j, double s) { out[j] = s; });
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
	//    strict lower triangle of LRm holds L by rows, LCm (if given) holds L by columns, d the pivots.  Left-looking by
	//    PANELS of PB columns: one sweep of dot products over the finished columns gives the panel's PB columns at once
	//    (a row's loads are shared by the PB sums); the coupling inside the panel is a PB x PB factorisation that every
	//    thread repeats in registers for the rows it owns - two barriers per panel instead of two per column.
	//    tbuf: N x PB doubles.  Returns false (on every thread) when a pivot is zero, not finite or (want_positive) negative.
	static constexpr int PB = 4;
	QA_FN bool ldl_factor(double *LRm, double *LCm, double *d, double *tbuf, const int N, const bool want_positive)
	{
		constexpr int L = T::LANES;
		bool ok = true;
		for (int j0 = 0; j0 < N && ok; j0 += PB) {
			const int nb = (N - j0 < PB) ? N - j0 : PB;
			const double *rj[PB];
#pragma unroll
			for (int c = 0; c < PB; c++) rj[c] = LRm + (size_t)(j0 + (c < nb ? c : nb - 1)) * N;
			for (int i = j0 + tm.warp; i < N; i += tm.nwarps) {
				const double *ri = LRm + (size_t)i * N;
				double s[PB];
#pragma unroll
				for (int c = 0; c < PB; c++) s[c] = 0.0;
				int k = tm.lane;
				for (; k + L < j0; k += 2 * L) {
					const double a0 = ri[k] * d[k], a1 = ri[k + L] * d[k + L];
					double b0[PB], b1[PB];
#pragma unroll
					for (int c = 0; c < PB; c++) {
						b0[c] = rj[c][k];
						b1[c] = rj[c][k + L];
					}
#pragma unroll
					for (int c = 0; c < PB; c++) s[c] = fma(a1, b1[c], fma(a0, b0[c], s[c]));
				}
				if (k < j0) {
					const double a0 = ri[k] * d[k];
#pragma unroll
					for (int c = 0; c < PB; c++) s[c] = fma(a0, rj[c][k], s[c]);
				}
#pragma unroll
				for (int c = 0; c < PB; c++) s[c] = tm.warp_sum(s[c]);
				if (tm.lane == 0) {
#pragma unroll
					for (int c = 0; c < PB; c++)
						if (c < nb) tbuf[(size_t)i * PB + c] = ri[j0 + c] - s[c];
				}
			}
			tm.sync();
			// the panel's own PB x PB block: Lp (strict lower), dd - the same numbers on every thread
			double Lp[PB][PB], dd[PB], di[PB];
#pragma unroll
			for (int c = 0; c < PB; c++) {
				dd[c] = 1.0;
				di[c] = 1.0;
#pragma unroll
				for (int r = 0; r < PB; r++) Lp[r][c] = 0.0;
			}
#pragma unroll
			for (int c = 0; c < PB; c++) {
				if (c < nb) {
					double v = tbuf[(size_t)(j0 + c) * PB + c];
#pragma unroll
					for (int e = 0; e < c; e++) v = fma(-Lp[c][e] * dd[e], Lp[c][e], v);
					dd[c] = v;
					if (!(fabs(v) > 0.0) || !(fabs(v) < INFINITY) || (want_positive && v < 0.0)) ok = false;
					di[c] = 1.0 / v;
#pragma unroll
					for (int r = c + 1; r < PB; r++) {
						if (r < nb) {
							double t = tbuf[(size_t)(j0 + r) * PB + c];
#pragma unroll
							for (int e = 0; e < c; e++) t = fma(-Lp[r][e] * dd[e], Lp[c][e], t);
							Lp[r][c] = t * di[c];
						}
					}
				}
			}
			if (ok) {
				for (int i = j0 + 1 + tm.tid; i < N; i += tm.nthreads) {
					const int cmax = (i - j0 < nb) ? i - j0 : nb; // columns j0 .. j0 + cmax - 1 lie left of the diagonal
					double li[PB];
#pragma unroll
					for (int c = 0; c < PB; c++) {
						li[c] = 0.0;
						if (c < cmax) {
							double t = tbuf[(size_t)i * PB + c];
#pragma unroll
							for (int e = 0; e < c; e++) t = fma(-li[e] * dd[e], Lp[c][e], t);
							li[c] = t * di[c];
							LRm[(size_t)i * N + j0 + c] = li[c];
							if (LCm) LCm[(size_t)(j0 + c) * N + i] = li[c];
						}
					}
				}
				if (tm.tid == 0)
					for (int c = 0; c < nb; c++) d[j0 + c] = dd[c];
			}
			tm.sync();
		}
		return ok;
	}

	// -- solve L D L' s = b in place, blocked by LANES: the diagonal block is solved by warp 0 with one row per lane
	//    (its elements loaded ahead of the shuffle chain), the rest of the right-hand side is updated by the whole team,
	//    one row per thread reading ITS contiguous LANES elements - rows of L (LRm) going forward, columns (LCm) going back
	QA_FN void ldl_solve(const double *LRm, const double *LCm, const double *d, double *b, const int N)
	{
		constexpr int BS = T::LANES;
		const int nblk = (N + BS - 1) / BS;
		for (int sweep = 0; sweep < 2; sweep++) {
			const double *M = sweep == 0 ? LRm : LCm;
			for (int q = 0; q < nblk; q++) {
				const int blk = sweep == 0 ? q : nblk - 1 - q;
				const int k0 = blk * BS, nb = (N - k0 < BS) ? N - k0 : BS;
				if (tm.warp == 0) {
					const bool live = tm.lane < nb;
					const double *mine = M + (size_t)(k0 + (live ? tm.lane : 0)) * N + k0;
					double br = live ? b[k0 + tm.lane] : 0.0;
#pragma unroll 8
					for (int q2 = 0; q2 < BS; q2++) {
						const int c = sweep == 0 ? q2 : BS - 1 - q2;
						const bool use = live && c < nb && (sweep == 0 ? c < tm.lane : c > tm.lane);
						const double l = use ? mine[c] : 0.0;
						const double xc = tm.warp_bcast(br, c);
						br = fma(-l, xc, br);
					}
					if (live) b[k0 + tm.lane] = br;
				}
				tm.sync();
				const int lo = sweep == 0 ? k0 + nb : 0, hi = sweep == 0 ? N : k0;
				for (int i = lo + tm.tid; i < hi; i += tm.nthreads) {
					const double *mi = M + (size_t)i * N + k0;
					double s0 = 0.0, s1 = 0.0;
					int c = 0;
					for (; c + 1 < nb; c += 2) {
						s0 = fma(mi[c], b[k0 + c], s0);
						s1 = fma(mi[c + 1], b[k0 + c + 1], s1);
					}
					if (c < nb) s0 = fma(mi[c], b[k0 + c], s0);
					b[i] -= s0 + s1;
				}
				tm.sync();
			}
			if (sweep == 0) {
				for (int i = tm.tid; i < N; i += tm.nthreads) b[i] /= d[i];
				tm.sync();
			}
		}
	}

	// -- K = P + sigma I + [A;I]' R [A;I], factored and inverted explicitly; false when K is not positive definite
	QA_FN bool factor()
	{
		// lower triangle of A'RA, a warp per column a, four columns b at a time sharing the loads of R a
		for (int a = tm.warp; a < n; a += tm.nwarps) {
			const double *ca = w.A + (size_t)a * mA;
			for (int b0 = 0; b0 <= a; b0 += 4) {
				const double *cb[4];
				double s[4];
#pragma unroll
				for (int r = 0; r < 4; r++) {
					cb[r] = w.A + (size_t)(b0 + r <= a ? b0 + r : a) * mA;
					s[r] = 0.0;
				}
				for (int i = tm.lane; i < mA; i += T::LANES) {
					const double t = ca[i] * w.rho_vec[i];
#pragma unroll
					for (int r = 0; r < 4; r++) s[r] = fma(t, cb[r][i], s[r]);
				}
#pragma unroll
				for (int r = 0; r < 4; r++) s[r] = tm.warp_sum(s[r]);
				if (tm.lane == 0) {
#pragma unroll
					for (int r = 0; r < 4; r++) {
						const int b = b0 + r;
						if (b <= a) {
							double v = w.P[(size_t)a * n + b] + s[r];
							if (a == b) v += st.sigma + w.Ib[a] * w.Ib[a] * w.rho_vec[mA + a];
							w.K[(size_t)a * n + b] = v;
							w.K[(size_t)b * n + a] = v;
						}
					}
				}
			}
		}
		for (size_t e = tm.tid; e < (size_t)n * n; e += tm.nthreads) w.Xinv[e] = 0.0;
		tm.sync();
		if (!ldl_factor(w.K, nullptr, w.kd, w.tcol, n, true)) return false;
		for (int i = tm.tid; i < n; i += tm.nthreads) w.pcol[i] = 1.0 / w.kd[i];
		// X = L^-1 (X[i][j] at Xinv[i*n + j]): a warp takes up to four adjacent columns at a time (as many as its scratch in
		// shared memory holds), the columns under construction live in the scratch, every row is one coalesced sweep over
		// L's row i feeding the four dot products and their interleaved shuffle reductions
		{
			const int ld = (n + 2) & ~1;
			int G = tm.scratch_len() / ld;
			G = G < 1 ? 1 : (G > 4 ? 4 : G);
			double *xs = tm.warp_scratch();
			for (int j0 = tm.warp * G; j0 < n; j0 += tm.nwarps * G) {
				const int ng = (n - j0 < G) ? n - j0 : G;
				for (int i = j0; i < n; i++) {
					const double *li = w.K + (size_t)i * n;
					double s[4] = {0.0, 0.0, 0.0, 0.0};
					for (int k = j0 + tm.lane; k < i; k += T::LANES) {
						const double l = li[k];
#pragma unroll
						for (int c = 0; c < 4; c++)
							if (c < ng) s[c] = fma(l, xs[c * ld + k], s[c]);
					}
#pragma unroll
					for (int c = 0; c < 4; c++) s[c] = tm.warp_sum(s[c]);
					if (tm.lane == 0) {
#pragma unroll
						for (int c = 0; c < 4; c++)
							if (c < ng) xs[c * ld + i] = (i == j0 + c ? 1.0 : 0.0) - s[c];
					}
					tm.warp_sync();
				}
				for (int c = 0; c < ng; c++)
					for (int i = j0 + c + tm.lane; i < n; i += T::LANES) w.Xinv[(size_t)i * n + j0 + c] = xs[c * ld + i];
				tm.warp_sync();
			}
		}
		tm.sync();
		// K^-1[a][b] = sum_{i >= max(a,b)} X[i][a] X[i][b] / d_i
		for (size_t e = tm.tid; e < (size_t)n * n; e += tm.nthreads) {
			const int a = (int)(e / n), b = (int)(e % n);
			if (b > a) continue; // the lower triangle, mirrored
			double s0 = 0.0, s1 = 0.0;
			int i = a;
			for (; i + 1 < n; i += 2) {
				s0 = fma(w.Xinv[(size_t)i * n + a] * w.pcol[i], w.Xinv[(size_t)i * n + b], s0);
				s1 = fma(w.Xinv[(size_t)(i + 1) * n + a] * w.pcol[i + 1], w.Xinv[(size_t)(i + 1) * n + b], s1);
			}
			if (i < n) s0 = fma(w.Xinv[(size_t)i * n + a] * w.pcol[i], w.Xinv[(size_t)i * n + b], s0);
			w.Kinv[e] = s0 + s1;
			w.Kinv[(size_t)b * n + a] = s0 + s1;
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
		ldl_solve(w.LR, w.LC, w.pd, w.psol, N);
		for (int it = 0; it < st.polish_refine_iter; it++) {
			// residual of the UNregularised system: [P x + A_act' y ; A_act x]
			// (a row is in at most one of the two lists: "lower" needs y < -(z - l) <= 0, "upper" y > u - z >= 0)
			for (int i = tm.tid; i < m; i += tm.nthreads) w.tmpm[i] = 0.0;
			tm.sync();
			for (int a = tm.tid; a < na; a += tm.nthreads) w.tmpm[w.rows[a]] = w.psol[n + a];
			tm.sync();
			mv_P(w.psol, w.Px);
			mv_At(w.tmpm, w.Aty);
			mv_A(w.psol, w.Ax);
			tm.sync();
			for (int j = tm.tid; j < n; j += tm.nthreads) w.pres[j] = w.prhs[j] - w.Px[j] - w.Aty[j];
			for (int a = tm.tid; a < na; a += tm.nthreads) w.pres[n + a] = w.prhs[n + a] - w.Ax[w.rows[a]];
			tm.sync();
			ldl_solve(w.LR, w.LC, w.pd, w.pres, N);
			double r[NRED] = {0, 0, 0, 0, 0, 0, 0, 0};
			for (int i = tm.tid; i < N; i += tm.nthreads) {
				const double s = w.psol[i] + w.pres[i];
				r[0] = fmax(r[0], fabs(w.pres[i]));
				r[1] = fmax(r[1], fabs(s));
				w.psol[i] = s;
			}
			reduce(r, 0u);
			if (!(r[0] > 1e-13 * r[1])) break; // the correction is five orders below the solver's own tolerance: done
		}
		double bad[NRED] = {0, 0, 0, 0, 0, 0, 0, 0};
		for (int j = tm.tid; j < n; j += tm.nthreads) {
			w.xp[j] = w.psol[j];
			if (w.psol[j] != w.psol[j]) bad[0] = 1.0;
		}
		for (int i = tm.tid; i < m; i += tm.nthreads) w.yp[i] = 0.0;
		reduce(bad, 0u);
		if (bad[0] != 0.0) return false;
		for (int a = tm.tid; a < na; a += tm.nthreads) w.yp[w.rows[a]] = w.psol[n + a]; // OSQP's get_ypol_from_yred (no row is listed twice)
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
		unsigned long long t0 = now_ns(), t_factor = 0, t_polish = 0;
		load_and_scale(pb);
		rho = st.rho;
		set_rho_vec();
		const unsigned long long t_scale = now_ns() - t0;
		int32_t status = ST_UNSOLVED, iter = 0, rho_updates = 0, polish_state = 0, n_active = 0;
		const double *xf = nullptr;
		const unsigned long long t_loop0 = now_ns();
		t0 = t_loop0;
		const bool f0 = factor();
		t_factor += now_ns() - t0;
		if (!f0) {
			status = ST_NON_CVX;
		} else {
			double *x = w.x0, *xprev = w.x1, *z = w.z0, *zprev = w.z1;
			for (int j = tm.tid; j < n; j += tm.nthreads) x[j] = xprev[j] = w.dx[j] = 0.0;
			for (int i = tm.tid; i < m; i += tm.nthreads) z[i] = zprev[i] = w.y[i] = w.dy[i] = w.zt[i] = 0.0;
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
				// rhs = sigma x_prev - q + [A;I]'(R z_prev - y); the vector R z - y is carried in w.zt from the last pass
				warp_dots<4>(n, mA, [&](int j) { return w.A + (size_t)j * mA; }, w.zt, [&](int j, double s) {
					w.rhs[j] = s + w.Ib[j] * w.zt[mA + j] + sigma * xprev[j] - w.q[j];
				});
				tm.sync();
				// x-tilde = K^-1 rhs
				warp_dots<4>(n, n, [&](int j) { return w.Kinv + (size_t)j * n; }, w.rhs, [&](int j, double s) { w.xt[j] = s; });
				tm.sync();
				// z-tilde = [A;I] x-tilde, then the relaxed x, the projected z and the dual step, row by row
				warp_dots<4>(mA, n, [&](int i) { return w.At + (size_t)i * n; }, w.xt, [&](int i, double s) {
					const double zr = alpha * s + (1.0 - alpha) * zprev[i];
					double zi = zr + w.y[i] * w.rho_inv[i];
					zi = fmin(fmax(zi, w.l[i]), w.u[i]);
					z[i] = zi;
					const double d = w.rho_vec[i] * (zr - zi);
					w.dy[i] = d;
					const double yn = w.y[i] + d;
					w.y[i] = yn;
					w.zt[i] = w.rho_vec[i] * zi - yn;
				});
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
					const double yn = w.y[i] + d;
					w.y[i] = yn;
					w.zt[i] = w.rho_vec[i] * zi - yn;
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
						for (int i = tm.tid; i < m; i += tm.nthreads) w.zt[i] = w.rho_vec[i] * z[i] - w.y[i];
						rho_updates++;
						t0 = now_ns();
						const bool fk = factor();
						t_factor += now_ns() - t0;
						if (!fk) {
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
					t0 = now_ns();
					const bool ok = polish(x, z, w.y, r, n_active);
					t_polish = now_ns() - t0;
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
				const unsigned long long t_all = now_ns() - t_loop0;
				pb.info[4] = (int32_t)(t_scale / 1000);
				pb.info[5] = (int32_t)(t_factor / 1000);
				pb.info[6] = (int32_t)((t_all - t_factor - t_polish) / 1000);
				pb.info[7] = (int32_t)(t_polish / 1000);
			}
		}
		tm.sync();
		return status;
	}
};

} // namespace qpadmm
