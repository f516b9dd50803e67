/* ref_common.h -- shared plumbing of the reference harness (TEST INFRASTRUCTURE ONLY). */
#ifndef ASIF_REF_COMMON_H
#define ASIF_REF_COMMON_H
#include <stdint.h>
#include <string.h>

struct RefFilter {
	int nx, nu, n_relax, nc, nv, n_diag;
	/* when costC is set, filter() calls the class's filter(x, H, c, uAct, relax) overload with these instead of uDes */
	const double *costH = 0, *costC = 0;
	/* ASIF only: when set, filter() calls filter(x, uDes, uAct, Lfh, Lgh, relax) (sticky in the reference: use a fresh object) */
	double *lieLfh = 0, *lieLgh = 0;
	virtual ~RefFilter() {}
	/* one reference filter() call; diag may be NULL */
	virtual int32_t filter(const double *x, const double *u_des, double *u_act, double *relax, double *diag) = 0;
	/* open-loop plant f(x), g(x) used by the example main loops */
	virtual void plant(const double *x, double *f, double *g) = 0;
	/* learned residual of the implicit classes (include/asif_learning_utils.h): dims[8] in the order of the
	 * LearningData fields, blob = w1,b1,w2,b2,w3,b3 of the drift net then of the actuation net (column-major weights) */
	virtual int set_learning(const uint32_t *dims, const double *blob) { (void)dims; (void)blob; return -1; }
	/* the class's updateOptions(options) with a new option vector (same layout as at creation); returns its code, -100 = not wired */
	virtual int32_t update_options(const double *opts, int n_opts) { (void)opts; (void)n_opts; return -100; }
};

#ifdef ASIF_REF_WITH_LEARNING_HELPER
#include <vector>
/* fills an ASIF::LearningData from (dims, blob); keeps the copy alive */
struct LearnStore {
	std::vector<double> w;
	template <class LD> void fill(LD &d, const uint32_t *dims, const double *blob)
	{
		d.d_drift_in = dims[0]; d.d_act_in = dims[1]; d.d_drift_hidden = dims[2]; d.d_act_hidden = dims[3];
		d.d_drift_hidden_2 = dims[4]; d.d_act_hidden_2 = dims[5]; d.d_drift_out = dims[6]; d.d_act_out = dims[7];
		const size_t nd = (size_t)dims[2] * dims[0] + dims[2] + (size_t)dims[4] * dims[2] + dims[4] + (size_t)dims[6] * dims[4] + dims[6];
		const size_t na = (size_t)dims[3] * dims[1] + dims[3] + (size_t)dims[5] * dims[3] + dims[5] + (size_t)dims[7] * dims[5] + dims[7];
		w.assign(blob, blob + nd + na);
		const double *p = w.data();
		d.w_1_drift = p; p += (size_t)dims[2] * dims[0];
		d.b_1_drift = p; p += dims[2];
		d.w_2_drift = p; p += (size_t)dims[4] * dims[2];
		d.b_2_drift = p; p += dims[4];
		d.w_3_drift = p; p += (size_t)dims[6] * dims[4];
		d.b_3_drift = p; p += dims[6];
		d.w_1_act = p; p += (size_t)dims[3] * dims[1];
		d.b_1_act = p; p += dims[3];
		d.w_2_act = p; p += (size_t)dims[5] * dims[3];
		d.b_2_act = p; p += dims[5];
		d.w_3_act = p; p += (size_t)dims[7] * dims[5];
		d.b_3_act = p; p += dims[7];
	}
};
#endif

RefFilter *make_di_explicit(const double *opts, int n_opts);
RefFilter *make_di_implicit_tb(const double *opts, int n_opts);
RefFilter *make_ip_implicit(const double *opts, int n_opts);
RefFilter *make_ip_robust(const double *opts, int n_opts);
RefFilter *make_ip_realizable(const double *opts, int n_opts);
RefFilter *make_segway_tb(const double *opts, int n_opts);
RefFilter *make_ip_implicit_rb(const double *opts, int n_opts);
RefFilter *make_di_implicit_rb(const double *opts, int n_opts);
#endif
