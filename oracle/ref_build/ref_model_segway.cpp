/* ref_model_segway.cpp -- ASIFimplicitTB on the segway example callbacks (TEST INFRASTRUCTURE ONLY). */
#include "ref_std_includes.h"

namespace ex_segway {
#include "examples/segway_implicit_tb.cpp"
}

void ref_tb_fill_diag(const ASIF::ASIFimplicitTB &f, const double *A, const double *b, int npBTSS, int nc, int nv, double *diag);
uint32_t ref_tb_npbtss(const double *opts, int n_opts, uint32_t dflt);
void ref_tb_options(const double *opts, int n_opts, ASIF::ASIFimplicitTB::Options &o);

namespace {

struct TbAccess : ASIF::ASIFimplicitTB {
	using ASIF::ASIFimplicitTB::ASIFimplicitTB;
	const double *A() const { return A_; }
	const double *b() const { return b_; }
};

/* Deviation D6: backup set centred on the backup controller's equilibrium
 * (examples/segway_implicit_tb.cpp:58 shifts theta by -0.1383244254 inside the controller, but
 * the shipped backupSet :41-55 is centred on the origin, so almost no sampled state ever reaches it).
 * Same arithmetic as the shipped callback with x replaced by x - xe. */
const double kSegwayXe[4] = {0.0, 0.0, 0.1383244254, 0.0};
void segway_backup_set_centred(const double *x, double *h, double *Dh, double *DDh)
{
	double xs[4];
	for (uint32_t i = 0; i < 4; i++) xs[i] = x[i] - kSegwayXe[i];
	ex_segway::backupSet(xs, h, Dh, DDh);
}

struct SegwayTb : RefFilter {
	TbAccess *f;
	int npBTSS;
	SegwayTb(const double *opts, int n_opts)
	{
		bool centred = true;
		if (opts && n_opts >= 10) centred = opts[9] != 0.0;
		npBTSS = (int)ref_tb_npbtss(opts, n_opts, ex_segway::npBTSS);
		f = new TbAccess(ex_segway::nx, ex_segway::nu, ex_segway::npSS, (uint32_t)npBTSS, ex_segway::safetySet,
		                 centred ? segway_backup_set_centred : ex_segway::backupSet, ex_segway::dynamics,
		                 ex_segway::dynamicsGradients, ex_segway::backupController);
		ASIF::ASIFimplicitTB::Options o;
		/* options of the example's main() (examples/segway_implicit_tb.cpp:223-230) */
		o.backTrajHorizon = 3.0;
		o.backTrajDt = 0.01;
		o.relaxCost = 10;
		o.relaxSafeLb = 2.0;
		o.relaxTTS = 30.0;
		o.relaxMinOrtho = 60.0;
		o.backTrajMinOrtho = 0.001;
		ref_tb_options(opts, n_opts, o);
		f->initialize(ex_segway::lb, ex_segway::ub, o);
		nx = 4; nu = 1; n_relax = 1; nc = npBTSS * 4 + 2; nv = 2; n_diag = 4 + npBTSS + nc * nv + nc;
	}
	~SegwayTb() { delete f; }
	int32_t filter(const double *x, const double *u_des, double *u_act, double *relax, double *diag) override
	{
		int32_t rc = costC ? f->filter(x, costH, costC, u_act, relax[0]) : f->filter(x, u_des, u_act, relax[0]);
		if (diag) ref_tb_fill_diag(*f, f->A(), f->b(), npBTSS, nc, nv, diag);
		return rc;
	}
	void plant(const double *x, double *fo, double *go) override { ex_segway::dynamics(x, fo, go); }
	int32_t update_options(const double *opts, int n_opts) override
	{
		ASIF::ASIFimplicitTB::Options o;
		ref_tb_options(opts, n_opts, o);
		return f->updateOptions(o);
	}
};
} // namespace

RefFilter *make_segway_tb(const double *opts, int n_opts) { return new SegwayTb(opts, n_opts); }
