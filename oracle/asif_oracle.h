/*
 * asif_oracle.h -- CPU restatement of the reference's safety-filter path (TEST INFRASTRUCTURE ONLY).
 *
 * Plain C99, FP64, compiled with -ffp-contract=off so that every product/sum rounds exactly
 * as the reference's GCC x86-64 build does.  Each function cites the reference file:line it
 * follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it;
 * nothing under asif_b200/ includes, links or loads anything from oracle/.
 *
 * Pinning: the reference ships no tests or golden vectors, so this
 * restatement is pinned against the reference's own sources compiled here
 * (oracle/_ref/libasif_ref.so, recipe oracle/ref_build/Makefile): tests/test_oracle_vs_ref.py
 * when /root/reference is present, and the committed fixtures tests/golden/ (.npz) (generated
 * by tests/golden/make_golden.py from libasif_ref.so) everywhere else.
 */
#ifndef ASIF_ORACLE_H
#define ASIF_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_QP_NVMAX 4
#define ORACLE_NXMAX 4
#define ORACLE_NUMAX 2

/* same numbering as oracle/ref_build/ref_api.h */
enum {
	ORACLE_CFG_DI_EXPLICIT = 1,
	ORACLE_CFG_DI_IMPLICIT_TB = 2,
	ORACLE_CFG_IP_IMPLICIT = 3,
	ORACLE_CFG_IP_ROBUST = 4,
	ORACLE_CFG_IP_REALIZABLE = 5,
	ORACLE_CFG_SEGWAY_TB = 6,
	ORACLE_CFG_IP_IMPLICIT_RB = 7, /* ASIFimplicitRB, InvertedPendulum callbacks (split gradients) */
	ORACLE_CFG_DI_IMPLICIT_RB = 8  /* ASIFimplicitRB, DoubleIntegrator_implicit_tb callbacks (fused gradient) */
};

/* exact QP by active-set enumeration (qp_enum.c).  Returns 1 or -3. */
int oracle_qp_solve(int nv, int nc, int diagonal_cost, const double *H, const double *c, const double *A,
                    const double *b, const double *lb, const double *ub, const unsigned char *be, double *sol);

/* model callbacks, same signatures and layouts as the reference's std::function members
 * (include/asif_implicit_tb.h:44-86): column-major Dh[npSS x nx], g[nx x nu], Du[nu x nx],
 * Dg[i + k*nx + j*nx*nu] */
typedef struct oracle_model {
	int nx, nu, npSS, npBS;
	double lb[ORACLE_NUMAX], ub[ORACLE_NUMAX];
	void (*safety_set)(const double *x, double *h, double *Dh);
	void (*backup_set)(const double *x, double *h, double *Dh, double *DDh); /* DDh may be NULL */
	void (*dynamics)(const double *x, double *f, double *g);
	void (*dynamics_gradients)(const double *x, double *Df, double *Dg);     /* NULL when fused */
	void (*dynamics_with_gradient)(const double *x, const double *u, double *f, double *g, double *d_fcl_dx); /* or NULL */
	void (*backup_controller)(const double *x, double *u, double *Du);
	/* ASIFimplicitRB only: h_int[j].convert().left() of safetySet_int over the box [x - x_unc, x + x_unc]
	 * (src/asif_implicit_robust.cpp:640-647), restating the libaffa operations the interval callback performs */
	void (*safety_set_lower)(const double *x, const double *x_unc, double *h_lo);
} oracle_model;

const oracle_model *oracle_get_model(int cfg, int variant);

/* option vectors are the same plain-double arrays oracle/ref_build/ref_api.h takes:
 * explicit: [relaxLb, relaxCost]
 * TB      : [relaxCost, relaxSafeLb, relaxTTS, relaxMinOrtho, backTrajHorizon, backTrajExtend,
 *            backTrajDt, backTrajMinOrtho, satSharpness, (segway) centred, lb, ub, npBTSS]
 * implicit: [relaxCost, relaxReachLb, relaxSafeLb, backTrajHorizon, backTrajDt, satSharpness, npBTSS]
 * implicitRB: the same seven, then [backContDt, x_unc[0..nx-1]]
 */
/* dims[0..5] = nx, nu, n_relax, nc, nv, n_diag ; returns 0 or -1 */
int oracle_dims(int cfg, const double *opts, int n_opts, int32_t *dims);
/* one call = the reference's filter(x, uDes, uAct, relax) on each of the n states, stateless */
int oracle_filter_batch(int cfg, const double *opts, int n_opts, int64_t n, const double *x, const double *u_des,
                        double *u_act, double *relax, int32_t *rc, double *diag);
/* the filter(x, H, c, uAct, relax) overloads on each of the n states: H nu x nu column-major (or NULL = identity
 * block), cvec[n*nv].  Classes ASIF, ASIFimplicitTB, ASIFimplicit, ASIFimplicitRB. */
int oracle_filter_batch_cost(int cfg, const double *opts, int n_opts, int64_t n, const double *x, const double *H,
                             const double *cvec, double *u_act, double *relax, int32_t *rc, double *diag);
/* ASIF::filter(x, uDes, uAct, Lfh, Lgh, relax) per state (cfg 1): Lfh[n*nc], Lgh[n*nc*nu] (per state column-major nc x nu) */
int oracle_filter_batch_lie(const double *opts, int n_opts, int64_t n, const double *x, const double *u_des, const double *Lfh,
                            const double *Lgh, double *u_act, double *relax, int32_t *rc, double *diag);
/* learned residual of the implicit classes (cfg 3, 7, 8), include/asif_learning_utils.h: process-wide switch, NULL = off.
 * dims[8] / blob as ref_set_learning (oracle/ref_build/ref_api.h) */
int oracle_set_learning(const uint32_t *dims, const double *blob);
/* closed-loop rollout (example main loops): x += dt*(f + g*uAct), in place */
int oracle_rollout(int cfg, const double *opts, int n_opts, int64_t n, int32_t steps, double dt, double *x,
                   const double *u_des, double *u_act_last, int32_t *rc_last, int64_t *rc_hist);

#ifdef __cplusplus
}
#endif
#endif
