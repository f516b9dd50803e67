/*
 * oracle_models.c -- the model callbacks of the named configs, restated in C
 * (TEST INFRASTRUCTURE ONLY).  The reference ships its models as example programs; each
 * function below follows the cited example line for line in the order of its floating-point
 * operations (terms that multiply an exact 0.0 and are then added to a +0.0-initialised
 * accumulator are dropped - they cannot change the IEEE result for finite inputs).
 */
#include "asif_oracle.h"

#include <math.h>
#include <string.h>

/* ---------------------------------------------------------------- DoubleIntegrator (explicit)
 * examples/DoubleIntegrator.cpp:12-61 */
static void di_safety(const double *x, double *h, double *Dh)
{
	/* :24-39, xBound = vBound = {-1, 1} */
	if (x[1] > 0) {
		h[0] = 1.0 - x[0] - (x[1] * x[1]) / 2.0; Dh[0] = -1.0; Dh[4] = -x[1];
		h[1] = x[0] - (-1.0);                    Dh[1] = 1.0;  Dh[5] = 0.0;
	} else {
		h[0] = -x[0] + 1.0;                         Dh[0] = -1.0; Dh[4] = 0.0;
		h[1] = x[0] - (-1.0) - (x[1] * x[1]) / 2.0; Dh[1] = 1.0;  Dh[5] = -x[1];
	}
	h[2] = x[1] - (-1.0); Dh[2] = 0.0; Dh[6] = 1.0;
	h[3] = -x[1] + 1.0;   Dh[3] = 0.0; Dh[7] = -1.0;
}

static void di_dynamics(const double *x, double *f, double *g)
{
	/* :41-61  f = A x with A = [0 1; 0 0] (column-major {0,0,1,0}), g = {0,1} */
	f[0] = 0.0 + 0.0 * x[0];
	f[0] += 1.0 * x[1];
	f[1] = 0.0 + 0.0 * x[0];
	f[1] += 0.0 * x[1];
	g[0] = 0.0;
	g[1] = 1.0;
}

static const oracle_model k_di_explicit = {
	2, 1, 4, 0, {-1.0, 0.0}, {1.0, 0.0}, di_safety, 0, di_dynamics, 0, 0, 0};

/* ---------------------------------------------------------------- DoubleIntegrator (implicit TB)
 * examples/DoubleIntegrator_implicit_tb.cpp:13-85 */
static void ditb_safety(const double *x, double *h, double *Dh)
{
	/* :32-38 */
	h[0] = -x[0] + 1.0;    Dh[0] = -1.0; Dh[4] = 0.0;
	h[1] = x[0] - (-1.0);  Dh[1] = 1.0;  Dh[5] = 0.0;
	h[2] = x[1] - (-1.0);  Dh[2] = 0.0;  Dh[6] = 1.0;
	h[3] = -x[1] + 1.0;    Dh[3] = 0.0;  Dh[7] = -1.0;
}

static void ditb_backup_set(const double *x, double *h, double *Dh, double *DDh)
{
	/* :40-55, P = I, mPpPt = -2I, Pv = 0.01; DDh completed to the full Hessian (deviation D3) */
	static const double P[4] = {1.0, 0.0, 0.0, 1.0};
	static const double mPpPt[4] = {-2.0, 0.0, 0.0, -2.0};
	h[0] = 0.01 * 0.01;
	for (int i = 0; i < 2; i++)
		for (int j = 0; j < 2; j++) h[0] -= P[i + j * 2] * x[i] * x[j];
	if (DDh)
		for (int k = 0; k < 4; k++) DDh[k] = mPpPt[k];
	for (int i = 0; i < 2; i++) {
		Dh[i] = 0.0;
		for (int k = 0; k < 2; k++) Dh[i] = Dh[i] + mPpPt[i + k * 2] * x[k];
	}
}

static void ditb_dynamics(const double *x, double *f, double *g)
{
	/* :57-63 matrixVectorMultiply(A, x) with A = {0,0,1,0}; g = B = {0,1} */
	static const double A[4] = {0.0, 0.0, 1.0, 0.0};
	for (int i = 0; i < 2; i++) {
		f[i] = 0.0;
		for (int k = 0; k < 2; k++) f[i] = f[i] + A[i + k * 2] * x[k];
	}
	g[0] = 0.0;
	g[1] = 1.0;
}

static void ditb_backup_controller(const double *x, double *u, double *Du)
{
	/* :65-72 K = {-10,-20} */
	u[0] = 0.0;
	u[0] = u[0] + (-10.0) * x[0];
	u[0] = u[0] + (-20.0) * x[1];
	Du[0] = -10.0;
	Du[1] = -20.0;
}

static void ditb_dynamics_with_gradient(const double *x, const double *u, double *f, double *g, double *d)
{
	/* :81-85 */
	(void)u;
	ditb_dynamics(x, f, g);
	d[0] = 0.0; d[1] = 0.0; d[2] = 1.0; d[3] = 0.0;
}

static void ditb_dynamics_gradients(const double *x, double *Df, double *Dg)
{
	/* :74-79 */
	(void)x;
	Df[0] = 0.0; Df[1] = 0.0; Df[2] = 1.0; Df[3] = 0.0;
	for (int i = 0; i < 4; i++) Dg[i] = 0.0;
}

/* variant 0: fused-gradient constructor (DYNAMICS_WITH_GRADIENT, the shipped default :9,97-100);
 * variant 1: split dynamics / dynamicsGradients constructor (:101-104) */
static const oracle_model k_di_tb = {
	2, 1, 4, 1, {-1.0, 0.0}, {1.0, 0.0}, ditb_safety, ditb_backup_set, ditb_dynamics, 0,
	ditb_dynamics_with_gradient, ditb_backup_controller};
static const oracle_model k_di_tb_split = {
	2, 1, 4, 1, {-1.0, 0.0}, {1.0, 0.0}, ditb_safety, ditb_backup_set, ditb_dynamics, ditb_dynamics_gradients,
	0, ditb_backup_controller};

/* ---------------------------------------------------------------- Segway (implicit TB)
 * examples/segway_implicit_tb.cpp:13-212 */
static const double k_seg_bound[4] = {3.0, 3.0, M_PI / 6, M_PI};

static void seg_safety(const double *x, double *h, double *Dh)
{
	/* :27-39 */
	for (int i = 0; i < 16; i++) Dh[i] = 0.0;
	for (int i = 0; i < 4; i++) {
		h[i] = (k_seg_bound[i] * k_seg_bound[i]) - (x[i] * x[i]);
		Dh[i * 5] = -2.0 * x[i];
	}
}

static void seg_backup_set_at(const double *x, double *h, double *Dh, double *DDh)
{
	/* :41-55, Pv = 0.05 */
	if (DDh)
		for (int i = 0; i < 16; i++) DDh[i] = 0.0;
	h[0] = 0.05 * 0.05;
	for (int i = 0; i < 4; i++) {
		h[0] -= (x[i] / k_seg_bound[i]) * (x[i] / k_seg_bound[i]);
		Dh[i] = -2.0 * x[i] / (k_seg_bound[i] * k_seg_bound[i]);
		if (DDh) DDh[i * 5] = -2.0 / (k_seg_bound[i] * k_seg_bound[i]);
	}
}

static void seg_backup_set_shipped(const double *x, double *h, double *Dh, double *DDh)
{
	seg_backup_set_at(x, h, Dh, DDh);
}

static void seg_backup_set_centred(const double *x, double *h, double *Dh, double *DDh)
{
	/* deviation D6: the shipped set evaluated at x - xe, xe = (0,0,0.1383244254,0) */
	double xs[4];
	xs[0] = x[0] - 0.0;
	xs[1] = x[1] - 0.0;
	xs[2] = x[2] - 0.1383244254;
	xs[3] = x[3] - 0.0;
	seg_backup_set_at(xs, h, Dh, DDh);
}

static void seg_backup_controller(const double *x, double *u, double *Du)
{
	/* :57-67 */
	static const double K[4] = {44.7214, 44.6528, 150.1612, 37.6492};
	double xt[4] = {0., 0., -0.1383244254, 0.};
	for (int i = 0; i < 4; i++) xt[i] += x[i];
	u[0] = 0.0;
	for (int k = 0; k < 4; k++) u[0] = u[0] + K[k] * xt[k];
	memcpy(Du, K, sizeof(K));
}

static void seg_dynamics(const double *X, double *f, double *g)
{
	/* :69-116 (MATLAB Coder output; the friction term carries a literal factor 0.0) */
	double Fric = 0.0 * 2.595498 * tanh(X[1] / 0.001);
	f[0] = X[1];
	double w2 = X[3] * X[3];
	double s1 = sin(X[2]);
	double s2 = sin(2.0 * X[2]);
	double c2 = cos(2.0 * X[2]);
	double c1 = cos(X[2]);
	double den = 1.0 / ((14.553176960783997 + -2.0831375273848773 * c2) + -0.59146430898882 * s2);
	f[1] = 0.0975 *
	       ((((((((((-23.195670626755415 * Fric + -0.0043160179477503974 * Fric * 44.798) +
	                -0.22270033964034344 * Fric * 44.798) +
	               44.798 *
	                   (((-1.3347669149041519 * Fric + -0.2693850964936445 * w2) + -0.0022454764220255392 * w2) +
	                    -0.11586336477125109 * w2) *
	                   0.195 * c1) +
	              59.510408935182809 * c2) +
	             -0.185817500742 * Fric * 44.798 * 0.195 * s1) +
	            86.686408318784913 * w2 * 0.195 * s1) +
	           0.72258001100852454 * w2 * 0.195 * s1) +
	          37.284092841364554 * w2 * 0.195 * s1) +
	         4.1423245261005457 * s2) +
	        -213.73800805067131 * s2) *
	       den;
	f[2] = X[3];
	f[3] = den * ((((((((((8.0 * Fric * 0.055936595310797 + 4.0 * Fric * 44.798 * 0.038025) +
	                      8.0 * Fric * 2.485 * 0.038025) +
	                     89.596 * (0.333691728726038 * Fric * 0.195 + -0.45669752988922296) * c1) +
	                    15.554616935932147 * w2 * 0.038025 * c2) +
	                   16.405863695295427 * s1) +
	                  0.092908750371 * Fric * 44.798 * 0.195 * s1) +
	                 249.80488266222164 * s1) +
	                27.713966400983114 * s1) +
	               1.0827059060875992 * w2 * 0.038025 * s2) +
	              -55.866072832711595 * w2 * 0.038025 * s2);
	g[0] = 0.0;
	double a = 1.4575004011882324 * c1;
	double b = 0.20290365220710288 * s1;
	g[1] = 0.551244194154502 * ((4.1706936767483551 + a) + b) *
	       (1.0 / (((8.3593271361634187 + -2.1243074194638587 * (c1 * c1)) + -0.04116989207898096 * (s1 * s1)) +
	               -0.29573215449441 * s2));
	g[2] = 0.0;
	g[3] = -5.65378660671284 * ((2.0043013906215941 + a) + b) * den;
}

static void seg_dynamics_gradients(const double *x, double *Df, double *Dg)
{
	/* :118-212 */
	double c1 = cos(x[2]);
	double s1 = sin(x[2]);
	double a2 = x[2] * 2.0;
	double w2 = x[3] * x[3];
	double c2 = cos(a2);
	double s2 = sin(a2);
	double th = tanh(x[1] * 1000.0);
	double th2 = th * th;
	double t25 = th * 15.13175750513302 - 40.918271887954823;
	double t26 = w2 * 3.3849959169972448 + th * 30.26351501026604;
	double t23 = 1.0 / ((c2 * 2.0831375273848769 + s2 * 0.59146430898882) - 14.553176960784);
	Df[0] = 0.0;
	Df[1] = 0.0;
	Df[2] = 0.0;
	Df[3] = 0.0;
	Df[4] = 1.0;
	double p = s1 * (th2 * 1000.0 - 1000.0);
	Df[5] = -t23 * (((th2 * 8443.5211353581435 + p * 0.41077609832706019) +
	                 c1 * (th2 * 30263.515010266041 - 30263.515010266041) * 0.0975) -
	                8443.5211353581435);
	Df[6] = 0.0;
	Df[7] = t23 * (((th2 * 20808.641003022261 + p * 2.1065440939849238) +
	                c1 * (th2 * 15131.75750513302 - 15131.75750513302)) -
	               20808.641003022261);
	Df[8] = 0.0;
	double cth = c1 * th;
	double sth = s1 * th;
	double cc = (c2 * 1.18292861797764 + -(s2 * 4.1662750547697547)) * (t23 * t23);
	Df[9] = t23 * ((((c2 * 40.8711582872913 + s2 * 11.604529742360651) - c1 * w2 * 2.3707272057666411) +
	                cth * 0.41077609832706019) -
	               s1 * t26 * 0.0975) -
	        cc * (((((c2 * -5.8022648711803244 + s2 * 20.435579143645651) + th * 8.443521135358143) -
	                s1 * w2 * 2.3707272057666411) +
	               sth * 0.41077609832706019) +
	              c1 * t26 * 0.0975);
	Df[10] = 0.0;
	double wc = w2 * c2;
	double ws = w2 * s2;
	Df[11] = t23 * ((((c1 * -293.92471275850022 - cth * 2.1065440939849238) + wc * 4.1662750547697547) +
	                 ws * 1.18292861797764) +
	                s1 * t25) +
	         cc * (((((s1 * 293.92471275850022 + th * 20.808641003022259) + wc * 0.59146430898881985) +
	                 sth * 2.1065440939849238) -
	                ws * 2.0831375273848769) +
	               c1 * t25);
	Df[12] = 0.0;
	Df[13] = t23 * (c1 * x[3] * 0.6600742038144628 - s1 * x[3] * 4.7414544115332831);
	Df[14] = 1.0;
	Df[15] = -t23 * (c2 * x[3] * 1.18292861797764 - s2 * x[3] * 4.1662750547697547);

	double a2b = x[2] * 2.0;
	double c2b = cos(a2b);
	double s2b = sin(a2b);
	double d4 = (c2b * 2.0831375273848769 + s2b * 0.59146430898882) - 14.553176960784;
	double d26 = ((c1 * c1 * 2.1243074194638591 + s2b * 0.29573215449441) + s1 * s1 * 0.04116989207898096) -
	             8.3593271361634187;
	for (int i = 0; i < 16; i++) Dg[i] = 0.0;
	Dg[9] = -(c1 * 0.1118494602519098 - s1 * 0.80343863413287053) / d26 +
	        1.0 / (d26 * d26) * (c2b * 0.59146430898882 - c1 * s1 * 4.1662750547697547) *
	            ((c1 * 0.80343863413287053 + s1 * 0.1118494602519098) + 2.2990706749044238);
	Dg[11] = (c1 * 1.1471739513016379 - s1 * 8.24039624751662) / d4 -
	         1.0 / (d4 * d4) * (c2b * 1.18292861797764 - s2b * 4.1662750547697547) *
	             ((c1 * 8.24039624751662 + s1 * 1.1471739513016379) + 11.33189235811229);
}

static const oracle_model k_segway_centred = {
	4, 1, 4, 1, {-20.0, 0.0}, {20.0, 0.0}, seg_safety, seg_backup_set_centred, seg_dynamics,
	seg_dynamics_gradients, 0, seg_backup_controller};
static const oracle_model k_segway_shipped = {
	4, 1, 4, 1, {-20.0, 0.0}, {20.0, 0.0}, seg_safety, seg_backup_set_shipped, seg_dynamics,
	seg_dynamics_gradients, 0, seg_backup_controller};


/* ---------------------------------------------------------------- InvertedPendulum (implicit)
 * examples/InvertedPendulum_Implicit.cpp:13-80 */
static void ip_safety(const double *x, double *h, double *Dh)
{
	/* :31-37, xBound = vBound = {-pi, pi} */
	h[0] = -x[0] + M_PI;    Dh[0] = -1.0; Dh[4] = 0.0;
	h[1] = x[0] - (-M_PI);  Dh[1] = 1.0;  Dh[5] = 0.0;
	h[2] = x[1] - (-M_PI);  Dh[2] = 0.0;  Dh[6] = 1.0;
	h[3] = -x[1] + M_PI;    Dh[3] = 0.0;  Dh[7] = -1.0;
}

static void ip_backup_set(const double *x, double *h, double *Dh, double *DDh)
{
	/* :39-52, P = {1.25,.25,.25,.25}, mPpPt = {-2.5,-.5,-.5,-.5}, Pv = 0.05 (not squared) */
	static const double P[4] = {1.25, 0.25, 0.25, 0.25};
	static const double mPpPt[4] = {-2.5, -0.5, -0.5, -0.5};
	(void)DDh;
	h[0] = 0.05;
	for (int i = 0; i < 2; i++)
		for (int j = 0; j < 2; j++) h[0] -= P[i + j * 2] * x[i] * x[j];
	for (int i = 0; i < 2; i++) {
		Dh[i] = 0.0;
		for (int k = 0; k < 2; k++) Dh[i] = Dh[i] + mPpPt[i + k * 2] * x[k];
	}
}

static void ip_dynamics(const double *x, double *f, double *g)
{
	/* :55-62 */
	f[0] = x[1];
	f[1] = sin(x[0]);
	g[0] = 0.;
	g[1] = 1.;
}

static void ip_backup_controller(const double *x, double *u, double *Du)
{
	/* :64-71, K = {-3,-3} */
	u[0] = 0.0;
	u[0] = u[0] + (-3.0) * x[0];
	u[0] = u[0] + (-3.0) * x[1];
	Du[0] = -3.0;
	Du[1] = -3.0;
}

static void ip_dynamics_gradients(const double *x, double *Df, double *Dg)
{
	/* :73-80 */
	Df[0] = 0.;        Df[2] = 1.;
	Df[1] = cos(x[0]); Df[3] = 0.;
	for (int i = 0; i < 4; i++) Dg[i] = 0.0;
}

static const oracle_model k_ip_implicit = {
	2, 1, 4, 1, {-1.5, 0.0}, {1.5, 0.0}, ip_safety, ip_backup_set, ip_dynamics, ip_dynamics_gradients, 0,
	ip_backup_controller};

/* ---------------------------------------------------------------- interval lower bounds for ASIFimplicitRB
 * safetySet_int on a box safety set {-x_i + hi_i, x_i - lo_i} evaluated in libaffa, then .convert().left():
 *   x_i   = AAF(interval(x_i - unc_i, x_i + unc_i)):  centre (right + left)/2, one noise term (right - left)/2
 *           (lib/libaffa/src/aa_aafcommon.cpp:80-101)
 *   -x_i  : centre and coefficient negated (aa_aafarithm.cpp:173-184);  (+-x_i) +- AAF(constant): centres added /
 *           subtracted, the coefficient carried over (aa_aafarithm.cpp:35-98, :103-163)
 *   left  = centre - rad, rad = 0 + |coefficient| (aa_aafcommon.cpp:217-245) */
static double aa_centre(double lo, double hi) { return (hi + lo) / 2; }
static double aa_radius(double lo, double hi)
{
	double coef = (hi - lo) / 2, sum = 0;
	if (coef >= 0.0) sum += coef;
	else sum += -coef;
	return sum;
}
static void box_safety_lower(const double *x, const double *unc, double lo0, double hi0, double lo1, double hi1, double *hl)
{
	const double c0 = aa_centre(x[0] - unc[0], x[0] + unc[0]), r0 = aa_radius(x[0] - unc[0], x[0] + unc[0]);
	const double c1 = aa_centre(x[1] - unc[1], x[1] + unc[1]), r1 = aa_radius(x[1] - unc[1], x[1] + unc[1]);
	hl[0] = (-c0 + hi0) - r0;
	hl[1] = (c0 - lo0) - r0;
	hl[2] = (c1 - lo1) - r1;
	hl[3] = (-c1 + hi1) - r1;
}
static void ip_safety_lower(const double *x, const double *unc, double *hl) { box_safety_lower(x, unc, -M_PI, M_PI, -M_PI, M_PI, hl); }
static void ditb_safety_lower(const double *x, const double *unc, double *hl) { box_safety_lower(x, unc, -1.0, 1.0, -1.0, 1.0, hl); }

static const oracle_model k_ip_implicit_rb = {
	2, 1, 4, 1, {-1.5, 0.0}, {1.5, 0.0}, ip_safety, ip_backup_set, ip_dynamics, ip_dynamics_gradients, 0,
	ip_backup_controller, ip_safety_lower};
static const oracle_model k_di_implicit_rb = {
	2, 1, 4, 1, {-1.0, 0.0}, {1.0, 0.0}, ditb_safety, ditb_backup_set, ditb_dynamics, 0,
	ditb_dynamics_with_gradient, ditb_backup_controller, ditb_safety_lower};

const oracle_model *oracle_get_model(int cfg, int variant)
{
	switch (cfg) {
	case ORACLE_CFG_DI_EXPLICIT: return &k_di_explicit;
	case ORACLE_CFG_DI_IMPLICIT_TB: return variant ? &k_di_tb_split : &k_di_tb;
	case ORACLE_CFG_SEGWAY_TB: return variant ? &k_segway_shipped : &k_segway_centred;
	case ORACLE_CFG_IP_IMPLICIT: return &k_ip_implicit;
	case ORACLE_CFG_IP_IMPLICIT_RB: return &k_ip_implicit_rb;
	case ORACLE_CFG_DI_IMPLICIT_RB: return &k_di_implicit_rb;
	default: return 0;
	}
}
