/* ref_model_ip.cpp -- reference filters on the InvertedPendulum example callbacks
 * (TEST INFRASTRUCTURE ONLY): ASIFimplicit (config 3a) and ASIFrobust (config 3b). */
#include "ref_std_includes.h"

namespace ex_ip_implicit {
#include "examples/InvertedPendulum_Implicit.cpp"
}
namespace ex_ip_robust {
#include "examples/InvertedPendulum_Robust.cpp"
}
namespace kd_70_135 {
using namespace std;
#include "KernelData_70-135kg.h"
}

namespace {

struct ImplicitAccess : ASIF::ASIFimplicit {
	using ASIF::ASIFimplicit::ASIFimplicit;
	const double *A() const { return A_; }
	const double *b() const { return b_; }
};

/* diag: [hSafetyNow, hBackupEnd (h_BS at the end of THIS call's trajectory), critIdx[npBTSS], A_[nc*nv], b_[nc]] */
struct IpImplicit : RefFilter {
	ImplicitAccess f;
	int npBTSS;
	ASIF::ASIFimplicit::Options o_;
	LearnStore learn_;
	int set_learning(const uint32_t *dims, const double *blob) override
	{
		learn_.fill(f.learning_data_, dims, blob);
		o_.use_learning = true;
		f.updateOptions(o_);
		return 0;
	}
	static uint32_t np(const double *opts, int n_opts) /* opts[6]: the constructor argument npBTSS */
	{
		return (opts && n_opts >= 7 && opts[6] >= 1.0 && opts[6] <= 16.0) ? (uint32_t)opts[6] : ex_ip_implicit::npBTSS;
	}
	int32_t update_options(const double *opts, int n_opts) override /* src/asif_implicit.cpp:369-401 */
	{
		ASIF::ASIFimplicit::Options o = o_;
		if (opts && n_opts >= 6) {
			o.relaxCost = opts[0];
			o.relaxReachLb = opts[1];
			o.relaxSafeLb = opts[2];
			o.backTrajHorizon = opts[3];
			o.backTrajDt = opts[4];
			o.satSharpness = opts[5];
		}
		o_ = o;
		return f.updateOptions(o);
	}
	IpImplicit(const double *opts, int n_opts)
	    : f(ex_ip_implicit::nx, ex_ip_implicit::nu, ex_ip_implicit::npSS, ex_ip_implicit::npBS, np(opts, n_opts),
	        ex_ip_implicit::safetySet, ex_ip_implicit::backupSet, ex_ip_implicit::dynamics,
	        ex_ip_implicit::dynamicsGradients, ex_ip_implicit::backupController)
	{
		ASIF::ASIFimplicit::Options o;
		/* options of the example's main() (examples/InvertedPendulum_Implicit.cpp:93-97) */
		o.backTrajHorizon = 5.0;
		o.backTrajDt = 0.001;
		o.relaxReachLb = 5.0;
		o.relaxSafeLb = 10.0;
		if (opts && n_opts >= 6) { /* [relaxCost, relaxReachLb, relaxSafeLb, backTrajHorizon, backTrajDt, satSharpness] */
			o.relaxCost = opts[0];
			o.relaxReachLb = opts[1];
			o.relaxSafeLb = opts[2];
			o.backTrajHorizon = opts[3];
			o.backTrajDt = opts[4];
			o.satSharpness = opts[5];
		}
		f.initialize(ex_ip_implicit::lb, ex_ip_implicit::ub, o);
		o_ = o;
		npBTSS = (int)np(opts, n_opts);
		nx = 2; nu = 1; n_relax = 2; nc = npBTSS * 4 + 1; nv = 3; n_diag = 2 + npBTSS + nc * nv + nc;
	}
	int32_t filter(const double *x, const double *u_des, double *u_act, double *relax, double *diag) override
	{
		int32_t rc = costC ? f.filter(x, costH, costC, u_act, relax) : f.filter(x, u_des, u_act, relax);
		if (diag) {
			double h[1], Dh[2];
			diag[0] = f.hSafetyNow_;
			ex_ip_implicit::backupSet(f.backTraj_.back().second.data(), h, Dh);
			diag[1] = h[0];
			for (int i = 0; i < npBTSS; i++) diag[2 + i] = (double)f.backTrajCritIdx_[i];
			memcpy(diag + 2 + npBTSS, f.A(), sizeof(double) * nc * nv);
			memcpy(diag + 2 + npBTSS + nc * nv, f.b(), sizeof(double) * nc);
		}
		return rc;
	}
	void plant(const double *x, double *fo, double *go) override { ex_ip_implicit::dynamics(x, fo, go); }
};


/* ------------------------------------------------------------------------------------------------
 * ASIFimplicitRB (src/asif_implicit_robust.cpp) on the InvertedPendulum_Implicit callbacks.
 * The reference ships no example of this class; the interval callbacks below are the example's own
 * expressions (examples/InvertedPendulum_Implicit.cpp:27-59) re-typed on interval_t, operator for operator.
 * Only safetySet_int and dynamics_int are ever called by the class (src/asif_implicit_robust.cpp:500,643).
 * ---------------------------------------------------------------------------------------------- */
void ip_safety_set_int(const interval_t *x, interval_t *h, interval_t *Dh)
{
	using namespace ex_ip_implicit;
	h[0] = -x[0] + xBound[1]; Dh[0] = -1.0; Dh[4] =  0.0;
	h[1] =  x[0] - xBound[0]; Dh[1] =  1.0; Dh[5] =  0.0;
	h[2] =  x[1] - vBound[0]; Dh[2] =  0.0; Dh[6] =  1.0;
	h[3] = -x[1] + vBound[1]; Dh[3] =  0.0; Dh[7] = -1.0;
}
void ip_backup_set_int(const interval_t *x, interval_t *h, interval_t *Dh)
{
	using namespace ex_ip_implicit;
	h[0] = Pv;
	for (uint32_t i = 0; i < nx; i++)
		for (uint32_t j = 0; j < nx; j++) h[0] = h[0] - P[i + j * nx] * x[i] * x[j];
	for (uint32_t i = 0; i < nx; i++) {
		Dh[i] = 0.0;
		for (uint32_t k = 0; k < nx; k++) Dh[i] = Dh[i] + mPpPt[i + k * nx] * x[k];
	}
}
void ip_dynamics_int(const interval_t *x, interval_t *f, interval_t *g)
{
	f[0] = x[1];
	f[1] = sin(x[0]);
	g[0] = 0.;
	g[1] = 1.;
}
void ip_dynamics_gradients_int(const interval_t *x, interval_t *Df, interval_t *Dg)
{
	Df[0] = 0.;        Df[2] = 1.;
	Df[1] = cos(x[0]); Df[3] = 0.;
	for (uint32_t i = 0; i < 4; i++) Dg[i] = 0.0;
}

struct RbAccess : ASIF::ASIFimplicitRB {
	using ASIF::ASIFimplicitRB::ASIFimplicitRB;
	const double *A() const { return A_; }
	const double *b() const { return b_; }
};

/* opts: [relaxCost, relaxReachLb, relaxSafeLb, backTrajHorizon, backTrajDt, satSharpness, npBTSS, backContDt, x_unc0, x_unc1]
 * diag: as IpImplicit */
struct IpImplicitRB : RefFilter {
	RbAccess f;
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
		return (opts && n_opts >= 7 && opts[6] >= 1.0 && opts[6] <= 16.0) ? (uint32_t)opts[6] : ex_ip_implicit::npBTSS;
	}
	IpImplicitRB(const double *opts, int n_opts)
	    : f(ex_ip_implicit::nx, ex_ip_implicit::nu, ex_ip_implicit::npSS, ex_ip_implicit::npBS, np(opts, n_opts),
	        ex_ip_implicit::safetySet, ip_safety_set_int, ex_ip_implicit::backupSet, ip_backup_set_int,
	        ex_ip_implicit::dynamics, ip_dynamics_int, ex_ip_implicit::dynamicsGradients, ip_dynamics_gradients_int,
	        ex_ip_implicit::backupController)
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
		o.x_unc = xunc; /* borrowed pointer (src/asif_implicit_robust.cpp:276-279) */
		f.initialize(ex_ip_implicit::lb, ex_ip_implicit::ub, o);
		o_ = o;
		npBTSS = (int)np(opts, n_opts);
		nx = 2; nu = 1; n_relax = 2; nc = npBTSS * 4 + 1; nv = 3; n_diag = 2 + npBTSS + nc * nv + nc;
	}
	int32_t filter(const double *x, const double *u_des, double *u_act, double *relax, double *diag) override
	{
		AAF::set_default(0); /* libaffa's global noise-symbol counter (F12) */
		int32_t rc = costC ? f.filter(x, costH, costC, u_act, relax) : f.filter(x, u_des, u_act, relax);
		if (diag) {
			double h[1], Dh[2];
			diag[0] = f.hSafetyNow_;
			ex_ip_implicit::backupSet(f.backTraj_.back().second.data(), h, Dh);
			diag[1] = h[0];
			for (int i = 0; i < npBTSS; i++) diag[2 + i] = (double)f.backTrajCritIdx_[i];
			memcpy(diag + 2 + npBTSS, f.A(), sizeof(double) * nc * nv);
			memcpy(diag + 2 + npBTSS + nc * nv, f.b(), sizeof(double) * nc);
		}
		return rc;
	}
	void plant(const double *x, double *fo, double *go) override { ex_ip_implicit::dynamics(x, fo, go); }
};

/* Config 3b.  The shipped example compiles its STANDARD block (pMin = pMax = 1, :12,30-33) and never fills
 * SafetySetData (:51).  Deviation D5: the table comes from include/KernelData_70-135kg.h and the interval
 * dynamics callback below is the example's (:62-69) with [pMin, pMax] as a parameter (the ROBUST block's
 * values 0.8 / 1.2 by default, :35-38). */
double g_pMin = 0.8, g_pMax = 1.2;
void ip_dynamics_interval(const interval_t *x, interval_t *f, interval_t *g)
{
	f[0] = x[1];
	f[1] = sin(x[0]);
	g[0] = 0.;
	g[1] = interval(g_pMin, g_pMax);
}

struct RobustAccess : ASIF::ASIFrobust {
	using ASIF::ASIFrobust::ASIFrobust;
	const double *A() const { return A_; }
};

/* diag: per half-plane k: [h_k, LgLo_k, LgHi_k, LfLo_k, LfHi_k]  (read back from the reference's A_, src/asif_robust.cpp:339-358) */
struct IpRobust : RefFilter {
	RobustAccess *f;
	int npSS, ncFull, nvFull;
	double pmid;
	IpRobust(const double *opts, int n_opts)
	{
		if (ex_ip_robust::SafetySetData.empty()) ex_ip_robust::SafetySetData = kd_70_135::SafetySetData;
		npSS = (int)ex_ip_robust::SafetySetData.size();
		ASIF::ASIFrobust::Options o; /* relaxLb 5, relaxCost 50 = the example's non-STANDARD options (:119-120) */
		g_pMin = 0.8;
		g_pMax = 1.2;
		if (opts && n_opts >= 4) { /* [relaxLb, relaxCost, pMin, pMax] */
			o.relaxLb = opts[0];
			o.relaxCost = opts[1];
			g_pMin = opts[2];
			g_pMax = opts[3];
		}
		pmid = 0.5 * (g_pMin + g_pMax);
		f = new RobustAccess(2, 1, npSS, ex_ip_robust::safetySet, ip_dynamics_interval);
		f->initialize(ex_ip_robust::lb, ex_ip_robust::ub, o);
		nx = 2; nu = 1; n_relax = 1;
		ncFull = 3 * npSS;          /* (nu+2) rows per safety function, src/asif_robust.cpp:21-22 */
		nvFull = 2 + 4 * npSS;      /* nu + 1 + 2 (nu+1) npSS */
		nc = 2 * npSS; nv = 2;      /* dimensions of the REDUCED problem the engine solves */
		n_diag = 5 * npSS;
	}
	~IpRobust() { delete f; }
	int32_t filter(const double *x, const double *u_des, double *u_act, double *relax, double *diag) override
	{
		AAF::set_default(0); /* keep libaffa's global noise-symbol counter from growing without bound (F12) */
		int32_t rc = f->filter(x, u_des, u_act, relax[0]);
		if (diag) {
			const double *A = f->A();
			int col = 2;
			for (int k = 0; k < npSS; k++, col += 4) {
				const int row = 3 * k;
				diag[5 * k + 0] = A[row + 1 * ncFull];
				diag[5 * k + 1] = A[row + (col + 0) * ncFull];
				diag[5 * k + 2] = -A[row + (col + 2) * ncFull];
				diag[5 * k + 3] = A[row + (col + 1) * ncFull];
				diag[5 * k + 4] = -A[row + (col + 3) * ncFull];
			}
		}
		return rc;
	}
	void plant(const double *x, double *fo, double *go) override
	{
		fo[0] = x[1];
		fo[1] = sin(x[0]);
		go[0] = 0.;
		go[1] = pmid;
	}
};
} // namespace

RefFilter *make_ip_implicit(const double *opts, int n_opts) { return new IpImplicit(opts, n_opts); }
RefFilter *make_ip_robust(const double *opts, int n_opts) { return new IpRobust(opts, n_opts); }
RefFilter *make_ip_implicit_rb(const double *opts, int n_opts) { return new IpImplicitRB(opts, n_opts); }
