/*
 * osqp_shim.cpp -- dense restatement of the OSQP 0.6.x algorithm (TEST INFRASTRUCTURE ONLY).
 *
 * Why this exists: the reference solves its QPs with OSQP (src/qpwrapper_osqp.cpp:114-121,
 * 164,177,188,223), a third-party dependency that is absent from /root/reference and from
 * this image (no network).  Parity for the safety-filter path is defined against "OSQP with
 * polishing, eps_abs = eps_rel = 1e-8", so the published algorithm is restated here and
 * linked under the unmodified reference sources (oracle/ref_build/Makefile).
 *
 * What is restated (Stellato, Banjac, Goulart, Bemporad, Boyd: "OSQP: an operator splitting
 * solver for quadratic programs", 2020, Alg. 1 + sec. 3.4 infeasibility, 4.1 polishing,
 * 5.1 Ruiz equilibration, 5.2 adaptive rho; constants follow OSQP 0.6.x constants.h):
 *   - modified Ruiz equilibration with cost scaling (scaling = 10 passes)
 *   - ADMM iteration with relaxation alpha, per-constraint rho (1e3*rho on equalities,
 *     rho_min on free rows), sigma regularisation
 *   - unscaled termination test every check_termination iterations, primal / dual
 *     infeasibility certificates from delta_y / delta_x
 *   - adaptive rho with the 5x hysteresis
 *   - solution polishing (active-set guess from (z,y), regularised KKT, 3 refinement steps)
 * What differs from the real library (none of it changes the optimum that is converged to):
 *   - dense Cholesky of P + sigma I + A' diag(rho) A replaces QDLDL on the sparse KKT matrix
 *     (same x-tilde; z-tilde = A x-tilde is algebraically identical to the nu-update)
 *   - adaptive_rho_interval is fixed (OSQP picks it from measured setup time, which makes the
 *     real library's iteration count timing-dependent)
 *   - with warm_start = 0 the step size rho is also reset per solve, so results are a pure
 *     function of the problem data (the real library lets rho carry over between solves)
 */
#include "osqp.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

namespace {

const double RHO_MIN = 1e-6, RHO_MAX = 1e6, RHO_EQ_OVER_RHO_INEQ = 1e3, RHO_TOL = 1e-4;
const double MIN_SCALING = 1e-4, MAX_SCALING = 1e4;

double g_cfg_eps = -1.0;
int g_cfg_polish = -1, g_cfg_warm = -1, g_cfg_max_iter = -1, g_cfg_refine = -1;
long long g_n_solves = 0, g_n_iters = 0, g_n_polish_ok = 0;
int g_last_status = 0, g_last_iter = 0;
long long g_n_inexact = 0; /* solves that ended in -2 / 2 / 3 / 4 (iteration limit or 'inaccurate' verdicts) */

typedef std::vector<double> vec;

inline double norm_inf(const vec &v)
{
	double r = 0;
	for (double e : v) r = std::max(r, std::fabs(e));
	return r;
}

inline double limit_scaling(double v)
{
	v = v < MIN_SCALING ? 1.0 : v;
	return v > MAX_SCALING ? MAX_SCALING : v;
}

/* in-place dense Cholesky (lower), column-major n x n; returns false if not PD */
bool chol_factor(vec &M, int n)
{
	for (int j = 0; j < n; j++) {
		double d = M[j + j * n];
		for (int k = 0; k < j; k++) d -= M[j + k * n] * M[j + k * n];
		if (!(d > 0)) return false;
		d = std::sqrt(d);
		M[j + j * n] = d;
		for (int i = j + 1; i < n; i++) {
			double s = M[i + j * n];
			for (int k = 0; k < j; k++) s -= M[i + k * n] * M[j + k * n];
			M[i + j * n] = s / d;
		}
	}
	return true;
}

void chol_solve(const vec &L, int n, vec &b)
{
	for (int i = 0; i < n; i++) {
		double s = b[i];
		for (int k = 0; k < i; k++) s -= L[i + k * n] * b[k];
		b[i] = s / L[i + i * n];
	}
	for (int i = n - 1; i >= 0; i--) {
		double s = b[i];
		for (int k = i + 1; k < n; k++) s -= L[k + i * n] * b[k];
		b[i] = s / L[i + i * n];
	}
}

/* LDL' (no pivoting) of a symmetric quasi-definite matrix, column-major; D on the diagonal */
bool ldl_factor(vec &M, int n)
{
	for (int j = 0; j < n; j++) {
		double d = M[j + j * n];
		for (int k = 0; k < j; k++) d -= M[j + k * n] * M[j + k * n] * M[k + k * n];
		if (d == 0.0 || d != d) return false;
		M[j + j * n] = d;
		for (int i = j + 1; i < n; i++) {
			double s = M[i + j * n];
			for (int k = 0; k < j; k++) s -= M[i + k * n] * M[j + k * n] * M[k + k * n];
			M[i + j * n] = s / d;
		}
	}
	return true;
}

void ldl_solve(const vec &M, int n, vec &b)
{
	for (int i = 0; i < n; i++) {
		double s = b[i];
		for (int k = 0; k < i; k++) s -= M[i + k * n] * b[k];
		b[i] = s;
	}
	for (int i = 0; i < n; i++) b[i] /= M[i + i * n];
	for (int i = n - 1; i >= 0; i--) {
		double s = b[i];
		for (int k = i + 1; k < n; k++) s -= M[k + i * n] * b[k];
		b[i] = s;
	}
}

struct Solver {
	int n, m;
	OSQPSettings st;
	/* sparsity patterns as handed to osqp_setup (needed by the index-based updates) */
	std::vector<c_int> Pp, Pi, Ap, Ai;
	vec Px, Ax;          /* unscaled CSC values */
	vec q0, l0, u0;      /* unscaled vectors */
	/* dense scaled data */
	vec P, A, q, l, u;   /* P n x n symmetric full, A m x n, column-major */
	vec D, E, Dinv, Einv;
	double c, cinv;
	vec rho_vec;
	std::vector<int> ctype;
	double rho;
	vec L;               /* Cholesky factor of P + sigma I + A' R A */
	bool factor_ok;
	/* iterates (scaled) */
	vec x, z, y, x_prev, z_prev, xt, zt, dx, dy;
	bool have_prev_solution;

	void densify()
	{
		P.assign((size_t)n * n, 0.0);
		for (int j = 0; j < n; j++)
			for (c_int k = Pp[j]; k < Pp[j + 1]; k++) {
				int i = (int)Pi[k];
				P[i + (size_t)j * n] = Px[k];
				P[j + (size_t)i * n] = Px[k];
			}
		A.assign((size_t)m * n, 0.0);
		for (int j = 0; j < n; j++)
			for (c_int k = Ap[j]; k < Ap[j + 1]; k++) A[Ai[k] + (size_t)j * m] = Ax[k];
		q = q0;
		l = l0;
		u = u0;
	}

	/* OSQP scaling.c::scale_data */
	void scale()
	{
		D.assign(n, 1.0);
		E.assign(m, 1.0);
		c = 1.0;
		vec Dt(n), Et(m);
		for (int it = 0; it < (int)st.scaling; it++) {
			for (int j = 0; j < n; j++) {
				double v = 0;
				for (int i = 0; i < n; i++) v = std::max(v, std::fabs(P[i + (size_t)j * n]));
				for (int i = 0; i < m; i++) v = std::max(v, std::fabs(A[i + (size_t)j * m]));
				Dt[j] = v;
			}
			for (int i = 0; i < m; i++) {
				double v = 0;
				for (int j = 0; j < n; j++) v = std::max(v, std::fabs(A[i + (size_t)j * m]));
				Et[i] = v;
			}
			for (int j = 0; j < n; j++) Dt[j] = 1.0 / std::sqrt(limit_scaling(Dt[j]));
			for (int i = 0; i < m; i++) Et[i] = 1.0 / std::sqrt(limit_scaling(Et[i]));
			for (int j = 0; j < n; j++)
				for (int i = 0; i < n; i++) P[i + (size_t)j * n] *= Dt[i] * Dt[j];
			for (int j = 0; j < n; j++)
				for (int i = 0; i < m; i++) A[i + (size_t)j * m] *= Et[i] * Dt[j];
			for (int j = 0; j < n; j++) {
				q[j] *= Dt[j];
				D[j] *= Dt[j];
			}
			for (int i = 0; i < m; i++) E[i] *= Et[i];
			/* cost normalisation */
			double mean = 0;
			for (int j = 0; j < n; j++) {
				double v = 0;
				for (int i = 0; i < n; i++) v = std::max(v, std::fabs(P[i + (size_t)j * n]));
				mean += v;
			}
			mean /= n;
			double ct = std::max(mean, limit_scaling(norm_inf(q)));
			ct = 1.0 / limit_scaling(ct);
			for (double &e : P) e *= ct;
			for (double &e : q) e *= ct;
			c *= ct;
		}
		cinv = 1.0 / c;
		Dinv.resize(n);
		Einv.resize(m);
		for (int j = 0; j < n; j++) Dinv[j] = 1.0 / D[j];
		for (int i = 0; i < m; i++) {
			Einv[i] = 1.0 / E[i];
			l[i] *= E[i];
			u[i] *= E[i];
		}
	}

	/* returns true if any constraint type changed (OSQP auxil.c::set_rho_vec / update_rho_vec) */
	bool set_rho_vec()
	{
		bool changed = false;
		rho = std::min(std::max(rho, RHO_MIN), RHO_MAX);
		rho_vec.resize(m);
		ctype.resize(m, -2);
		for (int i = 0; i < m; i++) {
			int t;
			if (l[i] < -OSQP_INFTY * MIN_SCALING && u[i] > OSQP_INFTY * MIN_SCALING) {
				t = -1;
				rho_vec[i] = RHO_MIN;
			} else if (u[i] - l[i] < RHO_TOL) {
				t = 1;
				rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * rho;
			} else {
				t = 0;
				rho_vec[i] = rho;
			}
			if (t != ctype[i]) changed = true;
			ctype[i] = t;
		}
		return changed;
	}

	void factor()
	{
		L.assign((size_t)n * n, 0.0);
		for (int j = 0; j < n; j++)
			for (int i = j; i < n; i++) {
				double s = P[i + (size_t)j * n];
				for (int k = 0; k < m; k++) s += A[k + (size_t)i * m] * rho_vec[k] * A[k + (size_t)j * m];
				L[i + (size_t)j * n] = s;
			}
		for (int j = 0; j < n; j++) L[j + (size_t)j * n] += st.sigma;
		factor_ok = chol_factor(L, n);
	}

	void setup_numeric(bool reset_rho)
	{
		densify();
		if (st.scaling) scale();
		else {
			D.assign(n, 1.0);
			E.assign(m, 1.0);
			Dinv = D;
			Einv = E;
			c = cinv = 1.0;
		}
		if (reset_rho) rho = st.rho;
		set_rho_vec();
		factor();
	}

	void matvecA(const vec &v, vec &out) const
	{
		out.assign(m, 0.0);
		for (int j = 0; j < n; j++) {
			double vj = v[j];
			if (vj == 0.0) continue;
			for (int i = 0; i < m; i++) out[i] += A[i + (size_t)j * m] * vj;
		}
	}
	void matvecAt(const vec &v, vec &out) const
	{
		out.assign(n, 0.0);
		for (int j = 0; j < n; j++) {
			double s = 0;
			for (int i = 0; i < m; i++) s += A[i + (size_t)j * m] * v[i];
			out[j] = s;
		}
	}
	void matvecP(const vec &v, vec &out) const
	{
		out.assign(n, 0.0);
		for (int j = 0; j < n; j++) {
			double vj = v[j];
			if (vj == 0.0) continue;
			for (int i = 0; i < n; i++) out[i] += P[i + (size_t)j * n] * vj;
		}
	}

	struct Res {
		double pri, dua, eps_pri_norm, eps_dua_norm, obj;
	};

	/* unscaled residuals and the norms entering the tolerances (auxil.c::compute_*_res/tol) */
	Res residuals(const vec &xx, const vec &zz, const vec &yy) const
	{
		Res r;
		vec Axv, Pxv, Aty;
		matvecA(xx, Axv);
		matvecP(xx, Pxv);
		matvecAt(yy, Aty);
		double pri = 0, nAx = 0, nz = 0;
		for (int i = 0; i < m; i++) {
			pri = std::max(pri, std::fabs(Einv[i] * (Axv[i] - zz[i])));
			nAx = std::max(nAx, std::fabs(Einv[i] * Axv[i]));
			nz = std::max(nz, std::fabs(Einv[i] * zz[i]));
		}
		double dua = 0, nPx = 0, nAty = 0, nq = 0, obj = 0;
		for (int j = 0; j < n; j++) {
			dua = std::max(dua, std::fabs(Dinv[j] * (Pxv[j] + q[j] + Aty[j])));
			nPx = std::max(nPx, std::fabs(Dinv[j] * Pxv[j]));
			nAty = std::max(nAty, std::fabs(Dinv[j] * Aty[j]));
			nq = std::max(nq, std::fabs(Dinv[j] * q[j]));
			obj += xx[j] * (0.5 * Pxv[j] + q[j]);
		}
		r.pri = pri;
		r.dua = cinv * dua;
		r.eps_pri_norm = std::max(nAx, nz);
		r.eps_dua_norm = cinv * std::max(std::max(nPx, nAty), nq);
		r.obj = cinv * obj;
		return r;
	}

	bool primal_infeasible(double eps)
	{
		vec d = dy;
		const double infval = OSQP_INFTY * MIN_SCALING;
		for (int i = 0; i < m; i++) {
			if (u[i] > infval) {
				if (l[i] < -infval) d[i] = 0.0;
				else d[i] = std::min(d[i], 0.0);
			} else if (l[i] < -infval) {
				d[i] = std::max(d[i], 0.0);
			}
		}
		double nd = 0;
		for (int i = 0; i < m; i++) nd = std::max(nd, std::fabs(E[i] * d[i]));
		if (nd > eps) {
			double lhs = 0;
			for (int i = 0; i < m; i++) lhs += u[i] * std::max(d[i], 0.0) + l[i] * std::min(d[i], 0.0);
			if (lhs < -eps * nd) {
				vec Atd;
				matvecAt(d, Atd);
				double na = 0;
				for (int j = 0; j < n; j++) na = std::max(na, std::fabs(Dinv[j] * Atd[j]));
				return na < eps * nd;
			}
		}
		return false;
	}

	bool dual_infeasible(double eps)
	{
		double ndx = 0;
		for (int j = 0; j < n; j++) ndx = std::max(ndx, std::fabs(D[j] * dx[j]));
		if (ndx > eps) {
			double qdx = 0;
			for (int j = 0; j < n; j++) qdx += q[j] * dx[j];
			if (qdx < -c * eps * ndx) {
				vec Pdx;
				matvecP(dx, Pdx);
				double np = 0;
				for (int j = 0; j < n; j++) np = std::max(np, std::fabs(Dinv[j] * Pdx[j]));
				if (np < c * eps * ndx) {
					vec Adx;
					matvecA(dx, Adx);
					const double infval = OSQP_INFTY * MIN_SCALING;
					for (int i = 0; i < m; i++) {
						double a = Einv[i] * Adx[i];
						if ((u[i] < infval && a > eps * ndx) || (l[i] > -infval && a < -eps * ndx)) return false;
					}
					return true;
				}
			}
		}
		return false;
	}

	/* auxil.c::check_termination; returns status or OSQP_UNSOLVED */
	int check(const Res &r, bool approximate)
	{
		double ea = st.eps_abs, er = st.eps_rel, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
		if (approximate) {
			ea *= 10;
			er *= 10;
			epi *= 10;
			edi *= 10;
		}
		bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
		if (m == 0) prim_ok = true;
		else {
			if (r.pri < ea + er * r.eps_pri_norm) prim_ok = true;
			else prim_inf = primal_infeasible(epi);
		}
		if (r.dua < ea + er * r.eps_dua_norm) dual_ok = true;
		else dual_inf = dual_infeasible(edi);
		if (prim_ok && dual_ok) return approximate ? OSQP_SOLVED_INACCURATE : OSQP_SOLVED;
		if (prim_inf) return approximate ? OSQP_PRIMAL_INFEASIBLE_INACCURATE : OSQP_PRIMAL_INFEASIBLE;
		if (dual_inf) return approximate ? OSQP_DUAL_INFEASIBLE_INACCURATE : OSQP_DUAL_INFEASIBLE;
		return OSQP_UNSOLVED;
	}

	double rho_estimate()
	{
		vec Axv, Pxv, Aty;
		matvecA(x, Axv);
		matvecP(x, Pxv);
		matvecAt(y, Aty);
		double pri = 0, nz = norm_inf(z), nAx = norm_inf(Axv);
		for (int i = 0; i < m; i++) pri = std::max(pri, std::fabs(Axv[i] - z[i]));
		double dua = 0;
		for (int j = 0; j < n; j++) dua = std::max(dua, std::fabs(Pxv[j] + q[j] + Aty[j]));
		pri /= (std::max(nz, nAx) + 1e-10);
		dua /= (std::max(std::max(norm_inf(q), norm_inf(Aty)), norm_inf(Pxv)) + 1e-10);
		double r = rho * std::sqrt(pri / (dua + 1e-10));
		return std::min(std::max(r, RHO_MIN), RHO_MAX);
	}

	/* polish.c::polish; returns true when the polished point was accepted */
	bool polish(Res &info_res)
	{
		std::vector<int> low, upp;
		for (int i = 0; i < m; i++)
			if (z[i] - l[i] < -y[i]) low.push_back(i);
		for (int i = 0; i < m; i++)
			if (u[i] - z[i] < y[i]) upp.push_back(i);
		int na = (int)(low.size() + upp.size());
		int N = n + na;
		std::vector<int> rows(low);
		rows.insert(rows.end(), upp.begin(), upp.end());
		vec K((size_t)N * N, 0.0); /* unregularised KKT (full symmetric) */
		for (int j = 0; j < n; j++)
			for (int i = 0; i < n; i++) K[i + (size_t)j * N] = P[i + (size_t)j * n];
		for (int a = 0; a < na; a++)
			for (int j = 0; j < n; j++) {
				double v = A[rows[a] + (size_t)j * m];
				K[(n + a) + (size_t)j * N] = v;
				K[j + (size_t)(n + a) * N] = v;
			}
		vec Kreg = K;
		for (int j = 0; j < n; j++) Kreg[j + (size_t)j * N] += st.delta;
		for (int a = 0; a < na; a++) Kreg[(n + a) + (size_t)(n + a) * N] -= st.delta;
		if (!ldl_factor(Kreg, N)) return false;
		vec rhs(N);
		for (int j = 0; j < n; j++) rhs[j] = -q[j];
		for (size_t a = 0; a < low.size(); a++) rhs[n + a] = l[low[a]];
		for (size_t a = 0; a < upp.size(); a++) rhs[n + low.size() + a] = u[upp[a]];
		vec sol = rhs;
		ldl_solve(Kreg, N, sol);
		for (int it = 0; it < (int)st.polish_refine_iter; it++) {
			vec d = rhs;
			for (int j = 0; j < N; j++) {
				double sj = sol[j];
				if (sj == 0.0) continue;
				for (int i = 0; i < N; i++) d[i] -= K[i + (size_t)j * N] * sj;
			}
			ldl_solve(Kreg, N, d);
			for (int i = 0; i < N; i++) sol[i] += d[i];
		}
		vec xp(sol.begin(), sol.begin() + n), zp, yp(m, 0.0);
		matvecA(xp, zp);
		for (int a = 0; a < na; a++) yp[rows[a]] = sol[n + a]; /* later (upper) entry wins on duplicates, as in get_ypol_from_yred */
		for (int i = 0; i < m; i++) zp[i] = std::min(std::max(zp[i], l[i]), u[i]);
		for (double e : xp)
			if (e != e) return false;
		Res pr = residuals(xp, zp, yp);
		if (getenv("OSQP_SHIM_DEBUG")) {
			fprintf(stderr, "polish: na=%d admm pri %g dua %g | pol pri %g dua %g | xp", na, info_res.pri, info_res.dua, pr.pri, pr.dua);
			for (int j = 0; j < n; j++) fprintf(stderr, " %.15g", D[j] * xp[j]);
			fprintf(stderr, " | x");
			for (int j = 0; j < n; j++) fprintf(stderr, " %.15g", D[j] * x[j]);
			fprintf(stderr, " | D %g %g c %g P00 %g P11 %g\n", D[0], D[n-1], c, P[0], P[(size_t)n*n-1]);
		}
		bool ok = (pr.pri < info_res.pri && pr.dua < info_res.dua) || (pr.pri < info_res.pri && info_res.dua < 1e-10) ||
		          (pr.dua < info_res.dua && info_res.pri < 1e-10);
		if (ok) {
			x = xp;
			z = zp;
			y = yp;
			info_res = pr;
		}
		return ok;
	}

	int solve(OSQPWorkspace *w)
	{
		g_n_solves++;
		if (getenv("OSQP_SHIM_TRACE") && n == 2 && atoi(getenv("OSQP_SHIM_TRACE")) > 1) {
			fprintf(stderr, "  data: q0 %g %g rho %g\n", q0[0], q0[1], rho);
			for (int i = 0; i < m; i++) {
				double a0 = 0, a1 = 0;
				for (c_int k = Ap[0]; k < Ap[1]; k++) if (Ai[k] == i) a0 = Ax[k];
				for (c_int k = Ap[1]; k < Ap[2]; k++) if (Ai[k] == i) a1 = Ax[k];
				fprintf(stderr, "  row %d: %.17g <= %.17g l0 + %.17g l1 <= %.17g\n", i, l0[i], a0, a1, u0[i]);
			}
		}
		if (!factor_ok) {
			w->info->status_val = OSQP_NON_CVX;
			return 1;
		}
		if (!st.warm_start || !have_prev_solution) {
			x.assign(n, 0.0);
			z.assign(m, 0.0);
			y.assign(m, 0.0);
			if (!st.warm_start && rho != st.rho) {
				rho = st.rho;
				set_rho_vec();
				factor();
			}
		}
		x_prev.assign(n, 0.0);
		z_prev.assign(m, 0.0);
		xt.resize(n);
		zt.resize(m);
		dx.assign(n, 0.0);
		dy.assign(m, 0.0);
		const double alpha = st.alpha, sigma = st.sigma;
		const int interval = st.adaptive_rho_interval > 0 ? (int)st.adaptive_rho_interval : 50;
		int status = OSQP_UNSOLVED, iter = 0, rho_updates = 0;
		Res r = {0, 0, 0, 0, 0};
		bool checked = false;
		vec tmp(m);
		for (iter = 1; iter <= (int)st.max_iter; iter++) {
			x_prev.swap(x);
			z_prev.swap(z);
			/* x-tilde: (P + sigma I + A'RA) xt = sigma x_prev - q + A'(R z_prev - y) */
			for (int i = 0; i < m; i++) tmp[i] = rho_vec[i] * z_prev[i] - y[i];
			matvecAt(tmp, xt);
			for (int j = 0; j < n; j++) xt[j] += sigma * x_prev[j] - q[j];
			chol_solve(L, n, xt);
			matvecA(xt, zt);
			for (int j = 0; j < n; j++) {
				x[j] = alpha * xt[j] + (1.0 - alpha) * x_prev[j];
				dx[j] = x[j] - x_prev[j];
			}
			for (int i = 0; i < m; i++) {
				double zr = alpha * zt[i] + (1.0 - alpha) * z_prev[i];
				double zi = zr + y[i] / rho_vec[i];
				zi = std::min(std::max(zi, l[i]), u[i]);
				z[i] = zi;
				dy[i] = rho_vec[i] * (zr - zi);
				y[i] += dy[i];
			}
			checked = st.check_termination && (iter % st.check_termination == 0);
			if (checked) {
				r = residuals(x, z, y);
				status = check(r, false);
				if (status != OSQP_UNSOLVED) break;
			}
			if (st.adaptive_rho && (iter % interval == 0)) {
				double rn = rho_estimate();
				if (rn > rho * st.adaptive_rho_tolerance || rn < rho / st.adaptive_rho_tolerance) {
					rho = rn;
					set_rho_vec();
					factor();
					rho_updates++;
				}
			}
		}
		if (iter > (int)st.max_iter) iter = (int)st.max_iter;
		if (!checked) {
			r = residuals(x, z, y);
			status = check(r, false);
		}
		if (status == OSQP_UNSOLVED) {
			status = check(r, true);
			if (status == OSQP_UNSOLVED) status = OSQP_MAX_ITER_REACHED;
		}
		g_n_iters += iter;
		w->info->iter = iter;
		w->info->rho_updates = rho_updates;
		w->info->status_polish = 0;
		if (st.polish && status == OSQP_SOLVED) {
			bool ok = polish(r);
			w->info->status_polish = ok ? 1 : -1;
			if (ok) g_n_polish_ok++;
		}
		w->info->status_val = status;
		g_last_status = status;
		if (status == OSQP_MAX_ITER_REACHED || status == OSQP_SOLVED_INACCURATE || status == OSQP_PRIMAL_INFEASIBLE_INACCURATE ||
		    status == OSQP_DUAL_INFEASIBLE_INACCURATE)
			g_n_inexact++;
		g_last_iter = iter;
		if (getenv("OSQP_SHIM_TRACE")) fprintf(stderr, "osqp_solve n=%d m=%d status=%d iter=%d pri=%g dua=%g\n", n, m, status, iter, r.pri, r.dua);
		w->info->pri_res = r.pri;
		w->info->dua_res = r.dua;
		w->info->obj_val = r.obj;
		w->info->rho_estimate = rho;
		const bool has_sol = (status == OSQP_SOLVED || status == OSQP_SOLVED_INACCURATE || status == OSQP_MAX_ITER_REACHED);
		for (int j = 0; j < n; j++) w->solution->x[j] = has_sol ? D[j] * x[j] : std::numeric_limits<double>::quiet_NaN();
		for (int i = 0; i < m; i++) w->solution->y[i] = has_sol ? cinv * E[i] * y[i] : std::numeric_limits<double>::quiet_NaN();
		have_prev_solution = has_sol;
		if (!has_sol) { /* OSQP cold-starts after an infeasible solve (osqp.c: "cold start if infeasible") */
			x.assign(n, 0.0);
			z.assign(m, 0.0);
			y.assign(m, 0.0);
		}
		return 0;
	}
};

} // namespace

extern "C" {

void osqp_shim_configure_refine(int polish_refine_iter) { g_cfg_refine = polish_refine_iter; }

void osqp_shim_configure(double eps_abs_rel, int polish, int warm_start, int max_iter)
{
	g_cfg_eps = eps_abs_rel;
	g_cfg_polish = polish;
	g_cfg_warm = warm_start;
	g_cfg_max_iter = max_iter;
}

long long osqp_shim_inexact_count(void) { return g_n_inexact; }

int osqp_shim_last_status(int *iters)
{
	if (iters) *iters = g_last_iter;
	return g_last_status;
}

void osqp_shim_stats(long long *n_solves, long long *n_iters, long long *n_polish_ok)
{
	if (n_solves) *n_solves = g_n_solves;
	if (n_iters) *n_iters = g_n_iters;
	if (n_polish_ok) *n_polish_ok = g_n_polish_ok;
}

void osqp_set_default_settings(OSQPSettings *s)
{
	s->rho = 0.1;
	s->sigma = 1e-6;
	s->scaling = 10;
	s->adaptive_rho = 1;
	s->adaptive_rho_interval = 0;
	s->adaptive_rho_tolerance = 5;
	s->adaptive_rho_fraction = 0.4;
	s->max_iter = 4000;
	s->eps_abs = 1e-3;
	s->eps_rel = 1e-3;
	s->eps_prim_inf = 1e-4;
	s->eps_dual_inf = 1e-4;
	s->alpha = 1.6;
	s->linsys_solver = 0;
	s->delta = 1e-6;
	s->polish = 0;
	s->polish_refine_iter = 3;
	s->verbose = 1;
	s->scaled_termination = 0;
	s->check_termination = 25;
	s->warm_start = 1;
	s->time_limit = 0;
}

csc *csc_matrix(c_int m, c_int n, c_int nzmax, c_float *x, c_int *i, c_int *p)
{
	csc *M = (csc *)c_malloc(sizeof(csc));
	if (!M) return OSQP_NULL;
	M->m = m;
	M->n = n;
	M->nz = -1;
	M->nzmax = nzmax;
	M->x = x;
	M->i = i;
	M->p = p;
	return M;
}

c_int osqp_setup(OSQPWorkspace **workp, const OSQPData *data, const OSQPSettings *settings)
{
	/* validate_data: P must be upper triangular (osqp auxil.c) */
	for (c_int j = 0; j < data->n; j++)
		for (c_int k = data->P->p[j]; k < data->P->p[j + 1]; k++)
			if (data->P->i[k] > j) {
				fprintf(stderr, "osqp_shim: P is not upper triangular\n");
				*workp = OSQP_NULL;
				return 3; /* OSQP_DATA_VALIDATION_ERROR */
			}
	OSQPWorkspace *w = (OSQPWorkspace *)c_calloc(1, sizeof(OSQPWorkspace));
	Solver *s = new Solver();
	s->n = (int)data->n;
	s->m = (int)data->m;
	s->st = *settings;
	if (g_cfg_eps > 0) s->st.eps_abs = s->st.eps_rel = g_cfg_eps;
	if (g_cfg_polish >= 0) s->st.polish = g_cfg_polish;
	if (g_cfg_warm >= 0) s->st.warm_start = g_cfg_warm;
	if (g_cfg_max_iter > 0) s->st.max_iter = g_cfg_max_iter;
	if (g_cfg_refine >= 0) s->st.polish_refine_iter = g_cfg_refine;
	s->Pp.assign(data->P->p, data->P->p + data->n + 1);
	s->Pi.assign(data->P->i, data->P->i + data->P->p[data->n]);
	s->Px.assign(data->P->x, data->P->x + data->P->p[data->n]);
	s->Ap.assign(data->A->p, data->A->p + data->n + 1);
	s->Ai.assign(data->A->i, data->A->i + data->A->p[data->n]);
	s->Ax.assign(data->A->x, data->A->x + data->A->p[data->n]);
	s->q0.assign(data->q, data->q + data->n);
	s->l0.assign(data->l, data->l + data->m);
	s->u0.assign(data->u, data->u + data->m);
	s->have_prev_solution = false;
	s->setup_numeric(true);
	w->impl = s;
	w->settings = (OSQPSettings *)c_malloc(sizeof(OSQPSettings));
	*w->settings = s->st;
	w->solution = (OSQPSolution *)c_malloc(sizeof(OSQPSolution));
	w->solution->x = (c_float *)c_calloc(data->n, sizeof(c_float));
	w->solution->y = (c_float *)c_calloc(data->m, sizeof(c_float));
	w->info = (OSQPInfo *)c_calloc(1, sizeof(OSQPInfo));
	w->info->status_val = OSQP_UNSOLVED;
	*workp = w;
	return 0;
}

c_int osqp_solve(OSQPWorkspace *w)
{
	return ((Solver *)w->impl)->solve(w);
}

c_int osqp_cleanup(OSQPWorkspace *w)
{
	if (!w) return 0;
	delete (Solver *)w->impl;
	c_free(w->settings);
	c_free(w->solution->x);
	c_free(w->solution->y);
	c_free(w->solution);
	c_free(w->info);
	c_free(w);
	return 0;
}

c_int osqp_update_lin_cost(OSQPWorkspace *w, const c_float *q_new)
{
	Solver *s = (Solver *)w->impl;
	s->q0.assign(q_new, q_new + s->n);
	for (int j = 0; j < s->n; j++) s->q[j] = s->c * s->D[j] * q_new[j];
	return 0;
}

static c_int update_bounds(Solver *s)
{
	for (int i = 0; i < s->m; i++) {
		if (s->l0[i] > s->u0[i]) return 1;
	}
	for (int i = 0; i < s->m; i++) {
		s->l[i] = s->E[i] * s->l0[i];
		s->u[i] = s->E[i] * s->u0[i];
	}
	if (s->set_rho_vec()) s->factor();
	return 0;
}

c_int osqp_update_lower_bound(OSQPWorkspace *w, const c_float *l_new)
{
	Solver *s = (Solver *)w->impl;
	s->l0.assign(l_new, l_new + s->m);
	return update_bounds(s);
}

c_int osqp_update_upper_bound(OSQPWorkspace *w, const c_float *u_new)
{
	Solver *s = (Solver *)w->impl;
	s->u0.assign(u_new, u_new + s->m);
	return update_bounds(s);
}

c_int osqp_update_P(OSQPWorkspace *w, const c_float *Px_new, const c_int *Px_new_idx, c_int P_new_n)
{
	Solver *s = (Solver *)w->impl;
	if (Px_new_idx) {
		for (c_int k = 0; k < P_new_n; k++) s->Px[Px_new_idx[k]] = Px_new[k];
	} else {
		if (P_new_n != (c_int)s->Px.size()) return 1;
		for (c_int k = 0; k < P_new_n; k++) s->Px[k] = Px_new[k];
	}
	s->setup_numeric(false);
	return 0;
}

c_int osqp_update_A(OSQPWorkspace *w, const c_float *Ax_new, const c_int *Ax_new_idx, c_int A_new_n)
{
	Solver *s = (Solver *)w->impl;
	if (Ax_new_idx) {
		for (c_int k = 0; k < A_new_n; k++) s->Ax[Ax_new_idx[k]] = Ax_new[k];
	} else {
		if (A_new_n != (c_int)s->Ax.size()) return 1;
		for (c_int k = 0; k < A_new_n; k++) s->Ax[k] = Ax_new[k];
	}
	s->setup_numeric(false);
	return 0;
}

} /* extern "C" */
