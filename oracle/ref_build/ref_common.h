/* ref_common.h -- shared plumbing of the reference harness (TEST INFRASTRUCTURE ONLY). */
#ifndef ASIF_REF_COMMON_H
#define ASIF_REF_COMMON_H
#include <stdint.h>
#include <string.h>

struct RefFilter {
	int nx, nu, n_relax, nc, nv, n_diag;
	virtual ~RefFilter() {}
	/* one reference filter() call; diag may be NULL */
	virtual int32_t filter(const double *x, const double *u_des, double *u_act, double *relax, double *diag) = 0;
	/* open-loop plant f(x), g(x) used by the example main loops */
	virtual void plant(const double *x, double *f, double *g) = 0;
};

RefFilter *make_di_explicit(const double *opts, int n_opts);
RefFilter *make_di_implicit_tb(const double *opts, int n_opts);
RefFilter *make_ip_implicit(const double *opts, int n_opts);
RefFilter *make_ip_robust(const double *opts, int n_opts);
RefFilter *make_ip_realizable(const double *opts, int n_opts);
RefFilter *make_segway_tb(const double *opts, int n_opts);
RefFilter *make_ip_implicit_rb(const double *opts, int n_opts);
RefFilter *make_di_implicit_rb(const double *opts, int n_opts);
#endif
