/*
 * qp_enum.c -- exact small-QP oracle by KKT active-set enumeration (TEST INFRASTRUCTURE ONLY).
 *
 * Contract restated: include/qpwrapper_abstract.h:11-15 of the reference
 *      min  v'Hv + c'v   s.t.  A v >= b (row i is an equality when be[i]),  lb <= v <= ub
 * with the conventions of src/qpwrapper_osqp.cpp:263-376 (dense column-major A[i + j*nc];
 * only diag(H) is used when diagonal_cost; variable bounds are just more rows; a bound of
 * magnitude >= 1e20 is a real number, never +-infinity).
 *
 * Method: for a strictly convex objective the optimum is the unique KKT point.  Every subset
 * S of at most nv linearly independent rows is tried as the active set: solve the equality-
 * constrained problem on S, accept it when all multipliers of inequality rows are >= 0 and
 * every other row is satisfied.  Deliberately brute force - it shares no logic with the
 * dual active-set method the CUDA kernels use, so agreement between the two is evidence.
 * The OSQP behaviour it stands for ("polish on, eps 1e-8") converges to the same point;
 * oracle/_ref links an ADMM restatement against which this file is pinned (tests/test_oracle_vs_ref.py).
 *
 * Returns 1 (optimal found) or -3 (no KKT point => primal infeasible; OSQP_PRIMAL_INFEASIBLE).
 */
#include "asif_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NVMAX ORACLE_QP_NVMAX

/* solve the k x k system M y = r in place by Gaussian elimination with partial pivoting.
 * returns 0 when M is (numerically) singular */
static int solve_small(int k, double M[NVMAX][NVMAX], double r[NVMAX])
{
	for (int c = 0; c < k; c++) {
		int p = c;
		double best = fabs(M[c][c]);
		for (int i = c + 1; i < k; i++)
			if (fabs(M[i][c]) > best) {
				best = fabs(M[i][c]);
				p = i;
			}
		if (best < 1e-13) return 0;
		if (p != c) {
			for (int j = 0; j < k; j++) {
				double t = M[c][j];
				M[c][j] = M[p][j];
				M[p][j] = t;
			}
			double t = r[c];
			r[c] = r[p];
			r[p] = t;
		}
		for (int i = c + 1; i < k; i++) {
			double f = M[i][c] / M[c][c];
			for (int j = c; j < k; j++) M[i][j] -= f * M[c][j];
			r[i] -= f * r[c];
		}
	}
	for (int i = k - 1; i >= 0; i--) {
		double s = r[i];
		for (int j = i + 1; j < k; j++) s -= M[i][j] * r[j];
		r[i] = s / M[i][i];
	}
	return 1;
}

typedef struct {
	int nv, m;       /* m = total rows incl. bounds */
	const double *G; /* m x nv row-major normals (row >= rhs form) */
	const double *r; /* rhs */
	const char *eq;  /* equality flag */
	double G2[NVMAX][NVMAX];   /* 2H */
	const double *c;
	double v0[NVMAX];          /* unconstrained minimiser */
} qp_t;

/* dense LU with partial pivoting for the (nv+k) x (nv+k) KKT system; returns 0 when singular */
#define KMAX (2 * NVMAX)
static int solve_kkt(int n, double M[KMAX][KMAX], double r[KMAX])
{
	for (int c = 0; c < n; c++) {
		int p = c;
		double best = fabs(M[c][c]);
		for (int i = c + 1; i < n; i++)
			if (fabs(M[i][c]) > best) {
				best = fabs(M[i][c]);
				p = i;
			}
		if (best < 1e-14) return 0;
		if (p != c) {
			for (int j = 0; j < n; j++) {
				double t = M[c][j];
				M[c][j] = M[p][j];
				M[p][j] = t;
			}
			double t = r[c];
			r[c] = r[p];
			r[p] = t;
		}
		for (int i = c + 1; i < n; i++) {
			double f = M[i][c] / M[c][c];
			for (int j = c; j < n; j++) M[i][j] -= f * M[c][j];
			r[i] -= f * r[c];
		}
	}
	for (int i = n - 1; i >= 0; i--) {
		double s = r[i];
		for (int j = i + 1; j < n; j++) s -= M[i][j] * r[j];
		r[i] = s / M[i][i];
	}
	return 1;
}

static int try_set(const qp_t *q, const int *S, int k, double *v_out, const double *rown, double ftol)
{
	const int nv = q->nv;
	double v[NVMAX];
	if (k == 0) {
		memcpy(v, q->v0, sizeof(double) * nv);
	} else if (k == nv) {
		/* a vertex: v is fixed by the rows alone (N' v = r_S), the multipliers follow from 2Hv + c = N mu.  Two
		 * k x k solves instead of the (nv+k) KKT system, whose Schur complement N'(2H)^-1 N squares the sine of
		 * the angle between near-parallel rows: a row pair at 1.7e-6 rad (a safety row with h = 1.7e-6 against the
		 * orthogonality row; relax = 1.8e4 at the optimum) put its pivot at 2.8e-14, under the round-off of the
		 * elimination, and the optimum was reported as "no KKT point" (found by the round-2 parity run). */
		double M[NVMAX][NVMAX], rr[NVMAX];
		for (int a = 0; a < k; a++) {
			const double *na = q->G + (size_t)S[a] * nv;
			const double sc = 1.0 / rown[S[a]];
			for (int i = 0; i < nv; i++) M[a][i] = na[i] * sc;
			rr[a] = q->r[S[a]] * sc;
		}
		if (!solve_small(k, M, rr)) return 0;
		for (int i = 0; i < nv; i++) v[i] = rr[i];
		double mu[NVMAX];
		for (int i = 0; i < nv; i++) {
			double t = q->c[i];
			for (int j = 0; j < nv; j++) t += q->G2[i][j] * v[j];
			mu[i] = t;
			for (int a = 0; a < k; a++) M[i][a] = q->G[(size_t)S[a] * nv + i] / rown[S[a]];
		}
		if (!solve_small(k, M, mu)) return 0;
		for (int a = 0; a < k; a++)
			if (!q->eq[S[a]] && mu[a] < 0.0) return 0;
	} else {
		/* KKT system of the equality-constrained problem on S, rows scaled to unit length:
		 *   [ 2H  -N ] [v ]   [ -c  ]
		 *   [ N'   0 ] [mu] = [ r_S ]      (N = scaled normals as columns)
		 * solved by LU with partial pivoting; one step of iterative refinement. */
		double K[KMAX][KMAX], K0[KMAX][KMAX], rhs[KMAX], rhs0[KMAX], sol[KMAX];
		const int n = nv + k;
		for (int i = 0; i < n; i++)
			for (int j = 0; j < n; j++) K[i][j] = 0.0;
		for (int i = 0; i < nv; i++) {
			for (int j = 0; j < nv; j++) K[i][j] = q->G2[i][j];
			rhs[i] = -q->c[i];
		}
		for (int a = 0; a < k; a++) {
			const double *na = q->G + (size_t)S[a] * nv;
			const double sc = 1.0 / rown[S[a]];
			for (int i = 0; i < nv; i++) {
				K[i][nv + a] = -na[i] * sc;
				K[nv + a][i] = na[i] * sc;
			}
			rhs[nv + a] = q->r[S[a]] * sc;
		}
		memcpy(K0, K, sizeof(K));
		memcpy(rhs0, rhs, sizeof(rhs));
		if (!solve_kkt(n, K, rhs)) return 0;
		memcpy(sol, rhs, sizeof(sol));
		/* refinement: residual with the unfactored matrix */
		double res[KMAX];
		for (int i = 0; i < n; i++) {
			double t = rhs0[i];
			for (int j = 0; j < n; j++) t -= K0[i][j] * sol[j];
			res[i] = t;
		}
		memcpy(K, K0, sizeof(K));
		if (solve_kkt(n, K, res))
			for (int i = 0; i < n; i++) sol[i] += res[i];
		for (int a = 0; a < k; a++)
			if (!q->eq[S[a]] && sol[nv + a] < 0.0) return 0;
		for (int i = 0; i < nv; i++) v[i] = sol[i];
	}
	/* primal feasibility of every row */
	for (int i = 0; i < q->m; i++) {
		const double *gi = q->G + (size_t)i * nv;
		double s = 0;
		for (int j = 0; j < nv; j++) s += gi[j] * v[j];
		double viol = s - q->r[i];
		double tol = ftol * (rown[i] > 1.0 ? rown[i] : 1.0);
		if (q->eq[i]) {
			if (fabs(viol) > tol) return 0;
		} else if (viol < -tol)
			return 0;
	}
	memcpy(v_out, v, sizeof(double) * nv);
	return 1;
}

int oracle_qp_solve(int nv, int nc, int diagonal_cost, const double *H, const double *c, const double *A,
                    const double *b, const double *lb, const double *ub, const unsigned char *be, double *sol)
{
	if (nv < 1 || nv > NVMAX) return -100;
	const int m = nc + 2 * nv;
	double *G = (double *)calloc((size_t)m * nv, sizeof(double));
	double *r = (double *)calloc(m, sizeof(double));
	double *rown = (double *)calloc(m, sizeof(double));
	char *eq = (char *)calloc(m, 1);
	int *cand = (int *)calloc(m, sizeof(int));
	qp_t q;
	q.nv = nv;
	q.m = m;
	for (int i = 0; i < nc; i++) {
		for (int j = 0; j < nv; j++) G[(size_t)i * nv + j] = A[i + (size_t)j * nc];
		r[i] = b[i];
		eq[i] = be ? (be[i] != 0) : 0;
	}
	for (int j = 0; j < nv; j++) {
		G[(size_t)(nc + j) * nv + j] = 1.0;
		r[nc + j] = lb[j];
		G[(size_t)(nc + nv + j) * nv + j] = -1.0;
		r[nc + nv + j] = -ub[j];
	}
	/* a pinned variable (lb == ub, e.g. the explicit filter's relax, src/asif.cpp:88-91) is an equality */
	for (int j = 0; j < nv; j++)
		if (lb[j] == ub[j]) eq[nc + j] = 1;
	q.G = G;
	q.r = r;
	q.eq = eq;
	/* 2H and v0 = -(2H)^-1 c */
	{
		double M[NVMAX][NVMAX], e[NVMAX];
		for (int i = 0; i < nv; i++) {
			for (int j = 0; j < nv; j++) {
				q.G2[i][j] = (diagonal_cost && i != j) ? 0.0 : 2.0 * H[i + j * nv];
				M[i][j] = q.G2[i][j];
			}
			e[i] = -c[i];
		}
		q.c = c;
		if (!solve_small(nv, M, e)) {
			free(G); free(r); free(rown); free(eq); free(cand);
			return -7; /* OSQP_NON_CVX */
		}
		for (int i = 0; i < nv; i++) q.v0[i] = e[i];
	}
	/* candidate rows: non-zero normal.  Zero rows (pads / trivial rows) only matter for feasibility. */
	int ncand = 0, neq = 0, infeasible = 0;
	for (int i = 0; i < m; i++) {
		double s = 0;
		for (int j = 0; j < nv; j++) s += G[(size_t)i * nv + j] * G[(size_t)i * nv + j];
		rown[i] = sqrt(s);
		if (rown[i] > 0.0) {
			/* a right-hand side of magnitude >= 1e19 (options_.inf = 1e20: trivial rows, the open upper bound
			 * of a relax variable) is a real number for feasibility but is never an active constraint: at
			 * that scale the multiplier test is pure round-off */
			if (fabs(r[i]) < 1e19 * rown[i]) {
				cand[ncand++] = i;
				if (eq[i]) neq++;
			}
		} else {
			rown[i] = 1.0;
			if (eq[i] ? (r[i] != 0.0) : (r[i] > 0.0)) infeasible = 1;
		}
	}
	const double ftol = 1e-9;
	int found = 0;
	int S[NVMAX];
	if (!infeasible) {
		/* all equalities must be in the active set; enumerate the rest by size */
		for (int k = neq > 0 ? neq : 0; k <= nv && !found; k++) {
			/* combinations of k rows out of ncand */
			int idx[NVMAX];
			for (int a = 0; a < k; a++) idx[a] = a;
			if (k > ncand) break;
			for (;;) {
				int eqcount = 0;
				for (int a = 0; a < k; a++) {
					S[a] = cand[idx[a]];
					if (eq[S[a]]) eqcount++;
				}
				if (eqcount == neq && try_set(&q, S, k, sol, rown, ftol)) {
					found = 1;
					break;
				}
				int a = k - 1;
				while (a >= 0 && idx[a] == ncand - k + a) a--;
				if (a < 0) break;
				idx[a]++;
				for (int t = a + 1; t < k; t++) idx[t] = idx[t - 1] + 1;
			}
		}
	}
	free(G); free(r); free(rown); free(eq); free(cand);
	return found ? 1 : -3;
}
