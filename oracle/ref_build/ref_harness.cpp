/* ref_harness.cpp -- C entry points of libasif_ref.so (TEST INFRASTRUCTURE ONLY; see ref_api.h). */
#include "ref_api.h"
#include "ref_common.h"
#include "osqp.h"
#include <vector>
#ifdef ASIF_REF_REAL_OSQP /* a real OSQP's osqp.h does not declare the stand-in's hooks: osqp_real_glue.cpp defines them as no-ops */
extern "C" {
void osqp_shim_configure(double eps_abs_rel, int polish, int warm_start, int max_iter);
void osqp_shim_configure_refine(int polish_refine_iter);
int osqp_shim_last_status(int *iters);
long long osqp_shim_inexact_count(void);
void osqp_shim_stats(long long *n_solves, long long *n_iters, long long *n_polish_ok);
}
#endif

extern "C" {

void ref_set_qp_mode(double eps, int polish, int warm_start, int max_iter)
{
	osqp_shim_configure(eps, polish, warm_start, max_iter);
}

void ref_set_polish_refine(int iters) { osqp_shim_configure_refine(iters); }

void ref_qp_stats(long long *n_solves, long long *n_iters, long long *n_polish_ok)
{
	osqp_shim_stats(n_solves, n_iters, n_polish_ok);
}

void *ref_create(int cfg, const double *opts, int n_opts)
{
	switch (cfg) {
	case REF_CFG_DI_EXPLICIT: return make_di_explicit(opts, n_opts);
	case REF_CFG_DI_IMPLICIT_TB: return make_di_implicit_tb(opts, n_opts);
	case REF_CFG_IP_IMPLICIT: return make_ip_implicit(opts, n_opts);
	case REF_CFG_IP_ROBUST: return make_ip_robust(opts, n_opts);
	case REF_CFG_IP_REALIZABLE: return make_ip_realizable(opts, n_opts);
	case REF_CFG_SEGWAY_TB: return make_segway_tb(opts, n_opts);
	case REF_CFG_IP_IMPLICIT_RB: return make_ip_implicit_rb(opts, n_opts);
	case REF_CFG_DI_IMPLICIT_RB: return make_di_implicit_rb(opts, n_opts);
	default: return 0;
	}
}

void ref_destroy(void *h) { delete (RefFilter *)h; }

int ref_set_learning(void *h, const uint32_t *dims, const double *blob)
{
	RefFilter *f = (RefFilter *)h;
	return f ? f->set_learning(dims, blob) : -1;
}

int32_t ref_update_options(void *h, const double *opts, int n_opts)
{
	RefFilter *f = (RefFilter *)h;
	return f ? f->update_options(opts, n_opts) : -101;
}

int ref_dims(void *h, int32_t *dims)
{
	RefFilter *f = (RefFilter *)h;
	if (!f) return -1;
	dims[0] = f->nx; dims[1] = f->nu; dims[2] = f->n_relax; dims[3] = f->nc; dims[4] = f->nv; dims[5] = f->n_diag;
	return 0;
}

int ref_filter_batch(void *h, int64_t n, const double *x, const double *u_des,
                     double *u_act, double *relax, int32_t *rc, double *diag)
{
	RefFilter *f = (RefFilter *)h;
	if (!f) return -1;
	for (int64_t k = 0; k < n; k++) {
		double r[2] = {0.0, 0.0};
		rc[k] = f->filter(x + k * f->nx, u_des + k * f->nu, u_act + k * f->nu, r, diag ? diag + k * f->n_diag : 0);
		for (int j = 0; j < f->n_relax; j++) relax[k * f->n_relax + j] = r[j];
	}
	return 0;
}

int ref_filter_batch_ex(void *h, int64_t n, const double *x, const double *u_des, double *u_act, double *relax,
                        int32_t *rc, double *diag, int32_t *qp_status, int32_t *qp_iters)
{
	RefFilter *f = (RefFilter *)h;
	if (!f) return -1;
	for (int64_t k = 0; k < n; k++) {
		double r[2] = {0.0, 0.0};
		const long long inexact0 = osqp_shim_inexact_count();
		rc[k] = f->filter(x + k * f->nx, u_des + k * f->nu, u_act + k * f->nu, r, diag ? diag + k * f->n_diag : 0);
		for (int j = 0; j < f->n_relax; j++) relax[k * f->n_relax + j] = r[j];
		int it = 0;
		int st = osqp_shim_last_status(&it); /* of the (last) QP this filter() call solved */
		/* any QP of this call (e.g. a facet-feasibility solve of ASIFrealizable) that ended inexactly taints the state */
		if (osqp_shim_inexact_count() != inexact0 && st != -2 && st != 2 && st != 3 && st != 4) st = -2;
		if (qp_status) qp_status[k] = st;
		if (qp_iters) qp_iters[k] = it;
	}
	return 0;
}

int ref_filter_batch_cost(void *h, int64_t n, const double *x, const double *H, const double *c, double *u_act, double *relax,
                          int32_t *rc, double *diag, int32_t *qp_status)
{
	RefFilter *f = (RefFilter *)h;
	if (!f) return -1;
	std::vector<double> ud0(f->nu, 0.0);
	for (int64_t k = 0; k < n; k++) {
		double r[2] = {0.0, 0.0};
		const long long inexact0 = osqp_shim_inexact_count();
		f->costH = H;
		f->costC = c + k * f->nv;
		rc[k] = f->filter(x + k * f->nx, ud0.data(), u_act + k * f->nu, r, diag ? diag + k * f->n_diag : 0);
		f->costH = 0;
		f->costC = 0;
		for (int j = 0; j < f->n_relax; j++) relax[k * f->n_relax + j] = r[j];
		int it = 0;
		int st = osqp_shim_last_status(&it);
		if (osqp_shim_inexact_count() != inexact0 && st != -2 && st != 2 && st != 3 && st != 4) st = -2;
		if (qp_status) qp_status[k] = st;
	}
	return 0;
}

int ref_filter_batch_lie(void *h, int64_t n, const double *x, const double *u_des, const double *Lfh, const double *Lgh,
                         double *u_act, double *relax, int32_t *rc, double *diag)
{
	RefFilter *f = (RefFilter *)h;
	if (!f) return -1;
	std::vector<double> lf(f->nc), lg(f->nc * f->nu);
	for (int64_t k = 0; k < n; k++) {
		double r[2] = {0.0, 0.0};
		for (int i = 0; i < f->nc; i++) lf[i] = Lfh[k * f->nc + i];
		for (int i = 0; i < f->nc * f->nu; i++) lg[i] = Lgh[k * f->nc * f->nu + i];
		f->lieLfh = lf.data();
		f->lieLgh = lg.data();
		rc[k] = f->filter(x + k * f->nx, u_des + k * f->nu, u_act + k * f->nu, r, diag ? diag + k * f->n_diag : 0);
		relax[k] = r[0];
	}
	return 0;
}

/* inexact_steps (optional, per agent): control steps of that agent whose QP ended in one of the ADMM stand-in's inexact
 * exits (iteration limit, "inaccurate" thresholds) - after such a step the agent's trajectory is no longer pinned */
int ref_rollout_ex(void *h, int64_t n, int32_t steps, double dt, double *x, const double *u_des,
                   double *u_act_last, int32_t *rc_last, int64_t *rc_hist, int32_t *inexact_steps);

int ref_rollout(void *h, int64_t n, int32_t steps, double dt, double *x, const double *u_des,
                double *u_act_last, int32_t *rc_last, int64_t *rc_hist)
{
	return ref_rollout_ex(h, n, steps, dt, x, u_des, u_act_last, rc_last, rc_hist, 0);
}

/* x_log[n][steps][nx], u_log[n][steps][nu], rc_log[n][steps], st_log[n][steps] (optional): the state every filter() call of the
 * loop saw, what it returned, and whether that call's QP took an inexact ADMM exit (0/1) - "teacher forcing" data: the CUDA
 * path is then evaluated call by call on exactly the states the reference's closed loop visited */
int ref_rollout_log(void *h, int64_t n, int32_t steps, double dt, double *x, const double *u_des, double *u_act_last,
                    int32_t *rc_last, int64_t *rc_hist, int32_t *inexact_steps, double *x_log, double *u_log, int32_t *rc_log,
                    int32_t *st_log);

int ref_rollout_ex(void *h, int64_t n, int32_t steps, double dt, double *x, const double *u_des,
                   double *u_act_last, int32_t *rc_last, int64_t *rc_hist, int32_t *inexact_steps)
{
	return ref_rollout_log(h, n, steps, dt, x, u_des, u_act_last, rc_last, rc_hist, inexact_steps, 0, 0, 0, 0);
}

int ref_rollout_log(void *h, int64_t n, int32_t steps, double dt, double *x, const double *u_des, double *u_act_last,
                    int32_t *rc_last, int64_t *rc_hist, int32_t *inexact_steps, double *x_log, double *u_log, int32_t *rc_log,
                    int32_t *st_log)
{
	RefFilter *f = (RefFilter *)h;
	if (!f) return -1;
	const int nx = f->nx, nu = f->nu;
	std::vector<double> fo(nx), go(nx * nu), fcl(nx), ua(nu);
	if (rc_hist) for (int i = 0; i < 8; i++) rc_hist[i] = 0;
	for (int64_t k = 0; k < n; k++) {
		double *xk = x + k * nx;
		int32_t rc = 0;
		for (int j = 0; j < nu; j++) ua[j] = 0.0;
		int32_t inexact = 0;
		for (int32_t s = 0; s < steps; s++) {
			double r[2];
			const long long inexact0 = osqp_shim_inexact_count();
			if (x_log) for (int i = 0; i < nx; i++) x_log[(k * steps + s) * nx + i] = xk[i];
			rc = f->filter(xk, u_des + k * nu, ua.data(), r, 0);
			const int bad = osqp_shim_inexact_count() != inexact0;
			if (bad) inexact++;
			if (u_log) for (int j = 0; j < nu; j++) u_log[(k * steps + s) * nu + j] = ua[j];
			if (rc_log) rc_log[k * steps + s] = rc;
			if (st_log) st_log[k * steps + s] = bad;
			if (rc_hist) rc_hist[(rc >= -3 && rc <= 2) ? rc + 3 : 7]++;
			/* plant step exactly as the example mains: fCl = 0; fCl += f; fCl += g*uAct; x += dt*fCl
			   (examples/segway_implicit_tb.cpp:265-283) */
			f->plant(xk, fo.data(), go.data());
			for (int i = 0; i < nx; i++) {
				fcl[i] = 0.0;
				fcl[i] += fo[i];
				for (int j = 0; j < nu; j++) fcl[i] += go[i + j * nx] * ua[j];
			}
			for (int i = 0; i < nx; i++) xk[i] += dt * fcl[i];
		}
		for (int j = 0; j < nu; j++) u_act_last[k * nu + j] = ua[j];
		rc_last[k] = rc;
		if (inexact_steps) inexact_steps[k] = inexact;
	}
	return 0;
}

} /* extern "C" */
