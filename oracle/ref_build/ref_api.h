/*
 * ref_api.h -- C entry points of oracle/_ref/libasif_ref.so (TEST INFRASTRUCTURE ONLY).
 *
 * libasif_ref.so = the UNMODIFIED reference sources (/root/reference/src/*.cpp,
 * lib/libaffa/src/*.cpp) + the reference's own example callbacks (included with `main`
 * fenced into a namespace) + the OSQP stand-in (osqp_shim/).  Nothing here is product code;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
 */
#ifndef ASIF_REF_API_H
#define ASIF_REF_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* model/filter configurations */
enum {
	REF_CFG_DI_EXPLICIT = 1,    /* ASIF           + examples/DoubleIntegrator.cpp             */
	REF_CFG_DI_IMPLICIT_TB = 2, /* ASIFimplicitTB + examples/DoubleIntegrator_implicit_tb.cpp */
	REF_CFG_IP_IMPLICIT = 3,    /* ASIFimplicit   + examples/InvertedPendulum_Implicit.cpp    */
	REF_CFG_IP_ROBUST = 4,      /* ASIFrobust     + examples/InvertedPendulum_Robust.cpp + KernelData_70-135kg.h */
	REF_CFG_IP_REALIZABLE = 5,  /* ASIFrealizable + IP dynamics + RealizableKernelData_100Hz_50pt.h */
	REF_CFG_SEGWAY_TB = 6,      /* ASIFimplicitTB + examples/segway_implicit_tb.cpp           */
	REF_CFG_IP_IMPLICIT_RB = 7, /* ASIFimplicitRB + examples/InvertedPendulum_Implicit.cpp callbacks (split gradients) */
	REF_CFG_DI_IMPLICIT_RB = 8  /* ASIFimplicitRB + examples/DoubleIntegrator_implicit_tb.cpp callbacks (fused gradient) */
};

/* QP mode for every filter created afterwards:
 *   eps<0   -> reference defaults (eps 1e-3, no polish, warm start; src/qpwrapper_osqp.cpp:67-69)
 *   eps>0   -> oracle setting: eps_abs=eps_rel=eps, polish, cold start, max_iter */
void ref_set_qp_mode(double eps, int polish, int warm_start, int max_iter);
/* OSQP polish_refine_iter override (default 3; see osqp_shim/osqp.h) */
void ref_set_polish_refine(int iters);
void ref_qp_stats(long long *n_solves, long long *n_iters, long long *n_polish_ok);

/* opts: config-specific option vector (see ref_harness.cpp), NULL = the options of the example's main() */
void *ref_create(int cfg, const double *opts, int n_opts);
void ref_destroy(void *h);
/* dims[0..5] = nx, nu, n_relax, nc (rows of A_), nv, n_diag */
int ref_dims(void *h, int32_t *dims);
/* implicit classes (cfg 3, 7, 8): switch on Options.use_learning with these networks (include/asif_learning_utils.h:8-32).
 * dims[8] = d_drift_in, d_act_in, d_drift_hidden, d_act_hidden, d_drift_hidden_2, d_act_hidden_2, d_drift_out, d_act_out;
 * blob = drift net (w1 [hidden x in] column-major, b1, w2, b2, w3, b3) followed by the actuation net, same order */
int ref_set_learning(void *h, const uint32_t *dims, const double *blob);
/* x[n*nx], u_des[n*nu] row-major per state; outputs u_act[n*nu], relax[n*n_relax], rc[n];
 * diag (optional, n*n_diag doubles): config-specific diagnostics followed by A_ (nc*nv, col-major) and b_ (nc). */
int ref_filter_batch(void *h, int64_t n, const double *x, const double *u_des,
                     double *u_act, double *relax, int32_t *rc, double *diag);
/* same, also returning the raw OSQP status_val / iteration count of each state's QP (single-threaded use only;
 * for calls that solve no QP the values are those of the previous solve) */
int ref_filter_batch_ex(void *h, int64_t n, const double *x, const double *u_des, double *u_act, double *relax,
                        int32_t *rc, double *diag, int32_t *qp_status, int32_t *qp_iters);
/* the classes' filter(x, H, c, uAct, relax) overloads per state (cfg 1, 2, 3, 6, 7, 8): H nu x nu column-major or NULL
 * (the class then keeps its current H_), c[n*nv]; qp_status as ref_filter_batch_ex */
int ref_filter_batch_cost(void *h, int64_t n, const double *x, const double *H, const double *c, double *u_act, double *relax,
                          int32_t *rc, double *diag, int32_t *qp_status);
/* cfg 1 only: ASIF::filter(x, uDes, uAct, Lfh, Lgh, relax) per state with Lfh[n*nc], Lgh[n*nc*nu].  The reference keeps
 * the caller's pointers and the custom mode for good (src/asif.cpp:137-139): use a filter object created for this purpose */
int ref_filter_batch_lie(void *h, int64_t n, const double *x, const double *u_des, const double *Lfh, const double *Lgh,
                         double *u_act, double *relax, int32_t *rc, double *diag);
/* closed-loop rollout exactly as the example main loops do (x += dt*(f+g*uAct)); x is updated in place */
int ref_rollout(void *h, int64_t n, int32_t steps, double dt, double *x, const double *u_des,
                double *u_act_last, int32_t *rc_last, int64_t *rc_hist /* [8] or NULL */);
#ifdef __cplusplus
}
#endif
#endif
