/*
 * osqp.h -- link-time stand-in for the OSQP 0.6.x C API (TEST INFRASTRUCTURE, not product code).
 *
 * The reference (DrewSingletary/asif) calls OSQP through src/qpwrapper_osqp.cpp
 * (find_package(osqp), CMakeLists.txt:38-50; version unpinned, API = 0.5/0.6 series).
 * OSQP is a third-party dependency that is neither vendored under /root/reference nor
 * installed in this image, so the ten entry points the wrapper binds
 * (src/qpwrapper_osqp.cpp:10-11,68,114-121,138,152,164,177,188,202,212,223,249) are
 * provided here and implemented in osqp_shim.cpp as a dense-linear-algebra restatement
 * of the published OSQP algorithm (Stellato et al., "OSQP: an operator splitting solver
 * for quadratic programs", Math. Prog. Comp. 2020; constants as in OSQP 0.6.x
 * include/constants.h).  Only the declarations the wrapper touches are present.
 */
#ifndef OSQP_SHIM_H
#define OSQP_SHIM_H

#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef long long c_int;
typedef double c_float;

#define c_malloc malloc
#define c_calloc calloc
#define c_free free
#define OSQP_NULL 0
#define OSQP_INFTY ((c_float)1e30)

/* status values, OSQP 0.6 constants.h */
#define OSQP_DUAL_INFEASIBLE_INACCURATE (4)
#define OSQP_PRIMAL_INFEASIBLE_INACCURATE (3)
#define OSQP_SOLVED_INACCURATE (2)
#define OSQP_SOLVED (1)
#define OSQP_MAX_ITER_REACHED (-2)
#define OSQP_PRIMAL_INFEASIBLE (-3)
#define OSQP_DUAL_INFEASIBLE (-4)
#define OSQP_SIGINT (-5)
#define OSQP_NON_CVX (-7)
#define OSQP_UNSOLVED (-10)

typedef struct {
	c_int nzmax, m, n;
	c_int *p, *i;
	c_float *x;
	c_int nz;
} csc;

typedef struct {
	c_float rho, sigma;
	c_int scaling;
	c_int adaptive_rho, adaptive_rho_interval;
	c_float adaptive_rho_tolerance, adaptive_rho_fraction;
	c_int max_iter;
	c_float eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, alpha;
	c_int linsys_solver;
	c_float delta;
	c_int polish, polish_refine_iter, verbose, scaled_termination, check_termination, warm_start;
	c_float time_limit;
} OSQPSettings;

typedef struct {
	c_int n, m;
	csc *P, *A;
	c_float *q, *l, *u;
} OSQPData;

typedef struct {
	c_float *x, *y;
} OSQPSolution;

typedef struct {
	c_int iter;
	char status[32];
	c_int status_val, status_polish;
	c_float obj_val, pri_res, dua_res;
	c_int rho_updates;
	c_float rho_estimate;
} OSQPInfo;

typedef struct {
	void *impl; /* dense solver state (osqp_shim.cpp) */
	OSQPSettings *settings;
	OSQPSolution *solution;
	OSQPInfo *info;
} OSQPWorkspace;

void osqp_set_default_settings(OSQPSettings *settings);
csc *csc_matrix(c_int m, c_int n, c_int nzmax, c_float *x, c_int *i, c_int *p);
c_int osqp_setup(OSQPWorkspace **workp, const OSQPData *data, const OSQPSettings *settings);
c_int osqp_solve(OSQPWorkspace *work);
c_int osqp_cleanup(OSQPWorkspace *work);
c_int osqp_update_lin_cost(OSQPWorkspace *work, const c_float *q_new);
c_int osqp_update_lower_bound(OSQPWorkspace *work, const c_float *l_new);
c_int osqp_update_upper_bound(OSQPWorkspace *work, const c_float *u_new);
c_int osqp_update_P(OSQPWorkspace *work, const c_float *Px_new, const c_int *Px_new_idx, c_int P_new_n);
c_int osqp_update_A(OSQPWorkspace *work, const c_float *Ax_new, const c_int *Ax_new_idx, c_int A_new_n);

/*
 * Shim-only control (deviation D4, oracle/README.md): the reference wrapper keeps its
 * OSQPSettings protected and only overrides max_iter (src/qpwrapper_osqp.cpp:67-69), so
 * the north-star oracle setting (polish on, eps_abs = eps_rel = 1e-8, cold start) is
 * injected here.  Values < 0 leave the corresponding OSQP default untouched.
 * Applied to every workspace created by osqp_setup after the call.
 */
void osqp_shim_configure(double eps_abs_rel, int polish, int warm_start, int max_iter);
/* polish_refine_iter (OSQP default 3).  With the default, a polished point can be WORSE than the
 * ADMM iterate it replaces when the Ruiz-scaled P is small against delta = 1e-6 (the refinement
 * contracts by delta/(P_ii+delta) per step) and OSQP still accepts it (polish.c acceptance rule,
 * second clause); observed |du| up to 2e-5 on the DoubleIntegrator TB config.  The oracle setting
 * therefore raises it (10) so that "OSQP with polish" means the converged KKT point. */
void osqp_shim_configure_refine(int polish_refine_iter);
/* raw status_val (and iteration count) of the most recent osqp_solve in this process */
int osqp_shim_last_status(int *iters);
/* number of solves so far that ended at the iteration limit or with an 'inaccurate' verdict (-2, 2, 3, 4) */
long long osqp_shim_inexact_count(void);
/* cumulative ADMM iterations / solves since process start (for reporting K-bar) */
void osqp_shim_stats(long long *n_solves, long long *n_iters, long long *n_polish_ok);

#ifdef __cplusplus
}
#endif
#endif
