/* ref_model_di.cpp -- reference filters on the DoubleIntegrator example callbacks
 * (TEST INFRASTRUCTURE ONLY).  The example files are included verbatim; their main() lands
 * in a namespace and is never called. */
#include "ref_std_includes.h"

namespace ex_di {
#include "examples/DoubleIntegrator.cpp"
}
namespace ex_di_tb {
#include "examples/DoubleIntegrator_implicit_tb.cpp"
}

namespace {

struct AsifAccess : ASIF::ASIF {
	using ASIF::ASIF::ASIF;
	const double *A() const { return A_; }
	const double *b() const { return b_; }
};

struct DiExplicit : RefFilter {
	AsifAccess f;
	static uint32_t sel(const double *opts, int n_opts) { return (opts && n_opts >= 3 && opts[2] > 0) ? (uint32_t)opts[2] : (uint32_t)-1; }
	DiExplicit(const double *opts, int n_opts)
	    : f(ex_di::nx, ex_di::nu, ex_di::npSS, ex_di::safetySet, ex_di::dynamics, sel(opts, n_opts)) /* opts[2] = npSSmax */
	{
		ASIF::ASIF::Options o; /* defaults: relaxLb 5, relaxCost 50 (include/asif.h:11-17) */
		if (opts && n_opts >= 2) {
			o.relaxLb = opts[0];
			o.relaxCost = opts[1];
		}
		f.initialize(ex_di::lb, ex_di::ub, o);
		nx = 2; nu = 1; n_relax = 1; nv = 2;
		nc = (sel(opts, n_opts) < 4) ? (int)sel(opts, n_opts) : 4;
		n_diag = nc * nv + nc;
	}
	int32_t filter(const double *x, const double *u_des, double *u_act, double *relax, double *diag) override
	{
		int32_t rc = costC ? f.filter(x, costH, costC, u_act, relax[0])
		                   : (lieLfh ? f.filter(x, u_des, u_act, lieLfh, lieLgh, relax[0]) : f.filter(x, u_des, u_act, relax[0]));
		if (diag) {
			memcpy(diag, f.A(), sizeof(double) * nc * nv);
			memcpy(diag + nc * nv, f.b(), sizeof(double) * nc);
		}
		return rc;
	}
	void plant(const double *x, double *fo, double *go) override { ex_di::dynamics(x, fo, go); }
	int32_t update_options(const double *opts, int n_opts) override
	{
		ASIF::ASIF::Options o;
		if (opts && n_opts >= 2) {
			o.relaxLb = opts[0];
			o.relaxCost = opts[1];
		}
		return f.updateOptions(o);
	}
};

struct TbAccess : ASIF::ASIFimplicitTB {
	using ASIF::ASIFimplicitTB::ASIFimplicitTB;
	const double *A() const { return A_; }
	const double *b() const { return b_; }
};

/* Deviation D3: examples/DoubleIntegrator_implicit_tb.cpp:49 writes only DDh[0..1]
 * of the 2x2 Hessian (DDh[i] with i<nx) and the library reads all four entries
 * (src/asif_implicit_tb.cpp:623).  The callback below calls the shipped one and then
 * completes DDh = mPpPt. */
void di_tb_backup_set_fixed(const double *x, double *h, double *Dh, double *DDh)
{
	ex_di_tb::backupSet(x, h, Dh, DDh);
	for (uint32_t k = 0; k < ex_di_tb::nx * ex_di_tb::nx; k++) DDh[k] = ex_di_tb::mPpPt[k];
}

} // namespace

/* diag layout shared by the TB configs:
 * [TTS, BTorthoBS, hSafetyNow, hBackupEnd, critIdx[npBTSS] (-1 = absent), A_[nc*nv], b_[nc]] */
void ref_tb_fill_diag(const ASIF::ASIFimplicitTB &f, const double *A, const double *b, int npBTSS, int nc, int nv, double *diag)
{
	diag[0] = f.TTS_;
	diag[1] = f.BTorthoBS_;
	diag[2] = f.hSafetyNow_;
	diag[3] = f.hBackupEnd_;
	for (int i = 0; i < npBTSS; i++) diag[4 + i] = i < (int)f.backTrajCritIdx_.size() ? (double)f.backTrajCritIdx_[i] : -1.0;
	memcpy(diag + 4 + npBTSS, A, sizeof(double) * nc * nv);
	memcpy(diag + 4 + npBTSS + nc * nv, b, sizeof(double) * nc);
}

void ref_tb_options(const double *opts, int n_opts, ASIF::ASIFimplicitTB::Options &o)
{
	/* [relaxCost, relaxSafeLb, relaxTTS, relaxMinOrtho, backTrajHorizon, backTrajExtend, backTrajDt, backTrajMinOrtho, satSharpness] */
	if (!opts || n_opts < 9) return;
	o.relaxCost = opts[0];
	o.relaxSafeLb = opts[1];
	o.relaxTTS = opts[2];
	o.relaxMinOrtho = opts[3];
	o.backTrajHorizon = opts[4];
	o.backTrajExtend = opts[5];
	o.backTrajDt = opts[6];
	o.backTrajMinOrtho = opts[7];
	o.satSharpness = opts[8];
}

/* opts[12]: the constructor argument npBTSS (default: the example's) */
uint32_t ref_tb_npbtss(const double *opts, int n_opts, uint32_t dflt)
{
	return (opts && n_opts >= 13 && opts[12] >= 1.0 && opts[12] <= 16.0) ? (uint32_t)opts[12] : dflt;
}

namespace {
struct DiTb : RefFilter {
	TbAccess f;
	int npBTSS;
	DiTb(const double *opts, int n_opts)
	    : f(ex_di_tb::nx, ex_di_tb::nu, ex_di_tb::npSS, ref_tb_npbtss(opts, n_opts, ex_di_tb::npBTSS), ex_di_tb::safetySet, di_tb_backup_set_fixed,
	        ex_di_tb::dynamicsWithGradient, ex_di_tb::backupController)
	{
		ASIF::ASIFimplicitTB::Options o;
		/* options of the example's main() (examples/DoubleIntegrator_implicit_tb.cpp:90-95) */
		o.backTrajHorizon = 2.0;
		o.backTrajDt = 0.001;
		o.relaxSafeLb = 10.0;
		o.relaxTTS = 5.0;
		o.relaxMinOrtho = 5.0;
		ref_tb_options(opts, n_opts, o);
		double lb[1] = {ex_di_tb::lb[0]}, ub[1] = {ex_di_tb::ub[0]};
		if (opts && n_opts >= 12) { /* initialize(lb, ub) with other input bounds: opts[10], opts[11] */
			lb[0] = opts[10];
			ub[0] = opts[11];
		}
		f.initialize(lb, ub, o);
		npBTSS = (int)ref_tb_npbtss(opts, n_opts, ex_di_tb::npBTSS);
		nx = 2; nu = 1; n_relax = 1; nc = npBTSS * 4 + 2; nv = 2; n_diag = 4 + npBTSS + nc * nv + nc;
	}
	int32_t filter(const double *x, const double *u_des, double *u_act, double *relax, double *diag) override
	{
		int32_t rc = costC ? f.filter(x, costH, costC, u_act, relax[0]) : f.filter(x, u_des, u_act, relax[0]);
		if (diag) ref_tb_fill_diag(f, f.A(), f.b(), npBTSS, nc, nv, diag);
		return rc;
	}
	void plant(const double *x, double *fo, double *go) override { ex_di_tb::dynamics(x, fo, go); }
	int32_t update_options(const double *opts, int n_opts) override /* src/asif_implicit_tb.cpp:365-405 */
	{
		ASIF::ASIFimplicitTB::Options o;
		ref_tb_options(opts, n_opts, o);
		return f.updateOptions(o);
	}
};
} // namespace


namespace {
/* ------------------------------------------------------------------------------------------------
 * ASIFimplicitRB (src/asif_implicit_robust.cpp), fused-gradient constructor, on the
 * DoubleIntegrator_implicit_tb callbacks (no example of this class is shipped).  Interval callbacks: the
 * example's expressions (examples/DoubleIntegrator_implicit_tb.cpp:30-36,55-85) re-typed on interval_t.
 * ---------------------------------------------------------------------------------------------- */
void di_safety_set_int(const interval_t *x, interval_t *h, interval_t *Dh)
{
	using namespace ex_di_tb;
	h[0] = -x[0] + xBound[1]; Dh[0] = -1.0; Dh[4] =  0.0;
	h[1] =  x[0] - xBound[0]; Dh[1] =  1.0; Dh[5] =  0.0;
	h[2] =  x[1] - vBound[0]; Dh[2] =  0.0; Dh[6] =  1.0;
	h[3] = -x[1] + vBound[1]; Dh[3] =  0.0; Dh[7] = -1.0;
}
void di_backup_set3(const double *x, double *h, double *Dh)
{
	double DDh[4];
	ex_di_tb::backupSet(x, h, Dh, DDh);
}
void di_backup_set_int(const interval_t *x, interval_t *h, interval_t *Dh)
{
	using namespace ex_di_tb;
	h[0] = Pv * Pv;
	for (uint32_t i = 0; i < nx; i++)
		for (uint32_t j = 0; j < nx; j++) h[0] = h[0] - P[i + j * nx] * x[i] * x[j];
	for (uint32_t i = 0; i < nx; i++) {
		Dh[i] = 0.0;
		for (uint32_t k = 0; k < nx; k++) Dh[i] = Dh[i] + mPpPt[i + k * nx] * x[k];
	}
}
void di_dynamics_with_gradient_int(const interval_t *x, const double *u, interval_t *f, interval_t *g, interval_t *d)
{
	using namespace ex_di_tb;
	for (uint32_t i = 0; i < nx; i++) { /* matrixVectorMultiply(A, x) on interval_t: accumulate from 0, k ascending */
		f[i] = 0.0;
		for (uint32_t k = 0; k < nx; k++) f[i] = f[i] + A[i + k * nx] * x[k];
	}
	for (uint32_t i = 0; i < nx * nu; i++) g[i] = B[i];
	for (uint32_t i = 0; i < nx * nx; i++) d[i] = A[i];
}

struct RbAccessDi : ASIF::ASIFimplicitRB {
	using ASIF::ASIFimplicitRB::ASIFimplicitRB;
	const double *A() const { return A_; }
	const double *b() const { return b_; }
};

/* opts: [relaxCost, relaxReachLb, relaxSafeLb, backTrajHorizon, backTrajDt, satSharpness, npBTSS, backContDt, x_unc0, x_unc1]
 * diag: [hSafetyNow, hBackupEnd (this call's trajectory), critIdx[npBTSS], A_[nc*nv], b_[nc]] */
struct DiImplicitRB : RefFilter {
	RbAccessDi f;
	int npBTSS;
	double xunc[2];
	ASIF::ASIFimplicitRB::Options o_;
	LearnStore learn_;
	int set_learning(const uint32_t *dims, const double *blob) override
	{
		learn_.fill(f.learning_data_, dims, blob);
		o_.use_learning = true;
		f.updateOptions(o_);
		return 0;
	}
	static uint32_t np(const double *opts, int n_opts)
	{
		return (opts && n_opts >= 7 && opts[6] >= 1.0 && opts[6] <= 16.0) ? (uint32_t)opts[6] : 10u;
	}
	DiImplicitRB(const double *opts, int n_opts)
	    : f(ex_di_tb::nx, ex_di_tb::nu, ex_di_tb::npSS, 1, np(opts, n_opts), ex_di_tb::safetySet, di_safety_set_int,
	        di_backup_set3, di_backup_set_int, ex_di_tb::dynamicsWithGradient, di_dynamics_with_gradient_int,
	        ex_di_tb::backupController)
	{
		ASIF::ASIFimplicitRB::Options o; /* defaults: include/asif_implicit_robust.h:22-38 */
		xunc[0] = xunc[1] = 0.0;
		if (opts && n_opts >= 10) {
			o.relaxCost = opts[0];
			o.relaxReachLb = opts[1];
			o.relaxSafeLb = opts[2];
			o.backTrajHorizon = opts[3];
			o.backTrajDt = opts[4];
			o.satSharpness = opts[5];
			o.backContDt = opts[7];
			xunc[0] = opts[8];
			xunc[1] = opts[9];
		}
		o.x_unc = xunc;
		f.initialize(ex_di_tb::lb, ex_di_tb::ub, o);
		o_ = o;
		npBTSS = (int)np(opts, n_opts);
		nx = 2; nu = 1; n_relax = 2; nc = npBTSS * 4 + 1; nv = 3; n_diag = 2 + npBTSS + nc * nv + nc;
	}
	int32_t filter(const double *x, const double *u_des, double *u_act, double *relax, double *diag) override
	{
		AAF::set_default(0);
		int32_t rc = costC ? f.filter(x, costH, costC, u_act, relax) : f.filter(x, u_des, u_act, relax);
		if (diag) {
			double h[1], Dh[2];
			diag[0] = f.hSafetyNow_;
			di_backup_set3(f.backTraj_.back().second.data(), h, Dh);
			diag[1] = h[0];
			for (int i = 0; i < npBTSS; i++) diag[2 + i] = (double)f.backTrajCritIdx_[i];
			memcpy(diag + 2 + npBTSS, f.A(), sizeof(double) * nc * nv);
			memcpy(diag + 2 + npBTSS + nc * nv, f.b(), sizeof(double) * nc);
		}
		return rc;
	}
	void plant(const double *x, double *fo, double *go) override { ex_di_tb::dynamics(x, fo, go); }
};
} // namespace

RefFilter *make_di_implicit_rb(const double *opts, int n_opts) { return new DiImplicitRB(opts, n_opts); }
RefFilter *make_di_explicit(const double *opts, int n_opts) { return new DiExplicit(opts, n_opts); }
RefFilter *make_di_implicit_tb(const double *opts, int n_opts) { return new DiTb(opts, n_opts); }
