// models.cuh -- device functors for the models of the named configs.
//
// A model keeps the reference's five callback signatures and layouts:
// column-major Dh[NPSS x NX], g[NX x NU], Du[NU x NX], Df[NX x NX], Dg[i + k*NX + j*NX*NU],
// so that a user model ports by copy-paste from its std::function form.  Constants are those
// of the reference's example programs (cited per functor); the floating-point operation order
// of each example is preserved (terms that are an exact 0.0 product added to a +0.0
// accumulator are dropped, which cannot change an IEEE result for finite inputs).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace asifb {

// Lower bound of an affine box safety set {-x_i + hi, x_i - lo} over the box [x - unc, x + unc] exactly as libaffa
// produces it for ASIFimplicitRB (src/asif_implicit_robust.cpp:636-647: safetySet_int, then h_int.convert().left()):
//   AAF(interval(l, r)): centre (r + l)/2, one noise coefficient (r - l)/2   (lib/libaffa/src/aa_aafcommon.cpp:80-101)
//   negation / adding a constant: centre negated / shifted, coefficient negated / kept (aa_aafarithm.cpp:35-98,103-184)
//   convert().left() = centre - rad(), rad() = 0 + |coefficient|                (aa_aafcommon.cpp:217-245)
__device__ __forceinline__ void box_safety_lower(const double *x, const double *unc, const double lo, const double hi, double *hl)
{
	const double l0 = x[0] - unc[0], r0 = x[0] + unc[0], l1 = x[1] - unc[1], r1 = x[1] + unc[1];
	const double c0 = (r0 + l0) / 2, c1 = (r1 + l1) / 2;
	const double d0 = fabs((r0 - l0) / 2), d1 = fabs((r1 - l1) / 2);
	hl[0] = (-c0 + hi) - d0;
	hl[1] = (c0 - lo) - d0;
	hl[2] = (c1 - lo) - d1;
	hl[3] = (-c1 + hi) - d1;
}

// sin and cos of one argument for the per-step dynamics of the pendulum and segway models.
// CUDA's sincos() is accurate but, inlined, materialises each of its ~17 polynomial / reduction constants with two
// UMOV instructions on every call (~34 issue slots per Euler step in ncu's source view); here the constants sit in
// constant memory and are read as FMA operands.  Argument reduction: k = rint(x * 2/pi), r = x - k*pi/2 with pi/2
// split into three doubles (Cody-Waite, FMA); polynomials: the fdlibm kernels (public domain) on [-pi/4, pi/4],
// evaluated with FMA.  Error <= ~1 ulp for |x| <= 1e5 (k is exact and the three-term reduction holds); beyond
// that the library routine is used.  These models have no bit parity with the reference anyway (CUDA's sin/cos
// differ from glibc's in the last bit): their parity is the 1e-9 / 1e-6 tolerance, which this keeps (tests).
static __constant__ double kSinCosC[18] = {
    6.36619772367581382433e-01,                                                                   // 2/pi
    1.57079632679489655800e+00, 6.12323399573676603587e-17, -1.49738490485916983294e-33,          // pi/2 = hi + mid + lo
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,         // S1..S3
    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,          // S4..S6
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,          // C1..C3
    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11,         // C4..C6
    1.0e5, 0.0};

__device__ __forceinline__ void sincos_model(const double x, double *sp, double *cp)
{
	// The fast path is computed unconditionally and the huge / NaN case (never on a sane trajectory) is a fix-up branch
	// BEHIND it: the polynomial chains then sit in the caller's basic block, where ptxas can interleave them with whatever
	// else the step computes that does not depend on them (a guard branch in front would fence them off).
	// k = rint(x 2/pi) by the add-and-subtract of 1.5 * 2^52: one FMA rounds the exact product to an integer in the last
	// place, the low word of the sum IS k (two's complement), and kd comes back by one subtraction - two fixed-latency FP64
	// instructions instead of DMUL + F2I + I2F, whose conversions went through the short-scoreboard path at the head of
	// every step's dependency chain.
	const double t = fma(x, kSinCosC[0], 6755399441055744.0);
	const int k = __double2loint(t);
	const double kd = t - 6755399441055744.0;
	double r = fma(-kd, kSinCosC[1], x);
	r = fma(-kd, kSinCosC[2], r);
	r = fma(-kd, kSinCosC[3], r);
	const double z = r * r;
	double ps = fma(z, kSinCosC[9], kSinCosC[8]);
	double pc = fma(z, kSinCosC[15], kSinCosC[14]);
	ps = fma(z, ps, kSinCosC[7]);
	pc = fma(z, pc, kSinCosC[13]);
	ps = fma(z, ps, kSinCosC[6]);
	pc = fma(z, pc, kSinCosC[12]);
	ps = fma(z, ps, kSinCosC[5]);
	pc = fma(z, pc, kSinCosC[11]);
	ps = fma(z, ps, kSinCosC[4]);
	pc = fma(z, pc, kSinCosC[10]);
	const double s = fma(r * z, ps, r);          // r + r z (S1 + z (...))
	const double c = fma(z, fma(z, pc, -0.5), 1.0); // 1 - z/2 + z^2 (C1 + z (...))
	const double a = (k & 1) ? c : s, b = (k & 1) ? s : c; // odd quadrant: sin <- cos, cos <- sin
	// quadrants 2,3 negate the sine, quadrants 1,2 the cosine: flip the sign bit with integer logic (one LOP3 each)
	*sp = __hiloint2double(__double2hiint(a) ^ ((k & 2) << 30), __double2loint(a));
	*cp = __hiloint2double(__double2hiint(b) ^ (((k + 1) & 2) << 30), __double2loint(b));
	// |x| >= 1e5, infinite or NaN: the library path.  1e5 = 0x40F86A00'00000000, so the test is one integer compare of
	// the high word (ALU) instead of a DSETP on the FP64 pipe.
	if ((__double2hiint(x) & 0x7fffffff) >= 0x40F86A00) sincos(x, sp, cp);
}

// ------------------------------------------------------------------ DoubleIntegrator (explicit)
// examples/DoubleIntegrator.cpp:12-61
struct DoubleIntegratorExplicit {
	static constexpr int NX = 2, NU = 1, NPSS = 4;
	// structural patterns (filter_common.cuh) read by explicit_kernel.cuh: DhSS column-major {-1, 1, 0, 0 | *, *, 1, -1}
	// (entries 4 and 5 are -x1 or 0 depending on the sign of x1: general), f = {x1, 0}, g = {0, 1}
	static constexpr bool HAS_PATTERNS = true;
	__host__ __device__ static constexpr int dhs_pat(int i)
	{
		constexpr int t[8] = {3, 1, 0, 0, 2, 2, 1, 3};
		return t[i];
	}
	__host__ __device__ static constexpr int f_pat(int i)
	{
		constexpr int t[2] = {2, 0};
		return t[i];
	}
	__host__ __device__ static constexpr int g_pat(int i)
	{
		constexpr int t[2] = {0, 1};
		return t[i];
	}
	__device__ static void safety_set(const double *x, double *h, double *Dh)
	{
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
	__device__ static void dynamics(const double *x, double *f, double *g)
	{
		f[0] = x[1]; // A = [0 1; 0 0]
		f[1] = 0.0;
		g[0] = 0.0;
		g[1] = 1.0;
	}
};

// ------------------------------------------------------------------ DoubleIntegrator (implicit TB)
// examples/DoubleIntegrator_implicit_tb.cpp:13-85 ; backup-set Hessian completed
struct DoubleIntegratorTB {
	static constexpr int NX = 2, NU = 1, NPSS = 4, NPBS = 1;
	static constexpr bool FUSED_GRADIENT = true; // DYNAMICS_WITH_GRADIENT, :9
	// structural patterns (filter_common.cuh): B = {0,1}; d_fcl_dx = A = {0,0,1,0} column-major
	__host__ __device__ static constexpr int g_pat(int i)
	{
		constexpr int t[2] = {0, 1};
		return t[i];
	}
	__host__ __device__ static constexpr int df_pat(int i)
	{
		constexpr int t[4] = {0, 0, 1, 0};
		return t[i];
	}
	__host__ __device__ static constexpr int dg_pat(int i)
	{
		constexpr int t[4] = {0, 0, 0, 0};
		return t[i];
	}
	__host__ __device__ static constexpr int f_pat(int i)
	{
		constexpr int t[2] = {2, 0};
		return t[i];
	}
	// DhSS column-major {-1,1,0,0 | 0,0,1,-1}
	__host__ __device__ static constexpr int dhs_pat(int i)
	{
		constexpr int t[8] = {3, 1, 0, 0, 0, 0, 1, 3};
		return t[i];
	}
	// min_j h_j(x) without forming the four values: min(1-x, x+1) = 1-|x| and rounding is monotone, so
	// min_j h_j = 1 - max(|x0|, |x1|) with the same bits as the reference's min_element over h
	static constexpr bool HAS_SAFETY_MIN = true;
	// (the larger magnitude by one compare and a select: fmax() adds NaN handling - five more instructions per Euler step)
	__device__ static double safety_min(const double *x)
	{
		const double m = (fabs(x[0]) > fabs(x[1])) ? x[0] : x[1]; // |.| rides on the compare and on the subtraction as operand modifiers
		return 1.0 - fabs(m);
	}
	__device__ static void safety_set(const double *x, double *h, double *Dh)
	{
		h[0] = -x[0] + 1.0;   Dh[0] = -1.0; Dh[4] = 0.0;
		h[1] = x[0] - (-1.0); Dh[1] = 1.0;  Dh[5] = 0.0;
		h[2] = x[1] - (-1.0); Dh[2] = 0.0;  Dh[6] = 1.0;
		h[3] = -x[1] + 1.0;   Dh[3] = 0.0;  Dh[7] = -1.0;
	}
	__device__ static double backup_set_value(const double *x)
	{
		// h = Pv^2 - sum P_ij x_i x_j with P = I: ((h - x0*x0) - 0) - 0) - x1*x1
		double h = 0.01 * 0.01;
		h -= x[0] * x[0];
		h -= x[1] * x[1];
		return h;
	}
	// backup_set_value(x) >= 0 without the last subtraction: an IEEE difference p - q has exactly the sign of the real
	// p - q and is zero only for p == q (gradual underflow), so ((c - x0^2) - x1^2) >= 0  <=>  (c - x0^2) >= x1^2
	static constexpr bool HAS_BACKUP_SET_REACHED = true;
	__device__ static double backup_set_level() { return 0.01 * 0.01; }
	// level = backup_set_level(), or -inf for a lane that must never report a hit (then the test is false by itself)
	__device__ static bool backup_set_reached(const double *x, const double level)
	{
		double h = level;
		h -= x[0] * x[0];
		return h >= x[1] * x[1];
	}
	__device__ static bool backup_set_reached(const double *x) { return backup_set_reached(x, backup_set_level()); }
	__device__ static void backup_set(const double *x, double &h, double *Dh, double *DDh)
	{
		h = backup_set_value(x);
		Dh[0] = -2.0 * x[0];
		Dh[1] = -2.0 * x[1];
		DDh[0] = -2.0; DDh[1] = 0.0; DDh[2] = 0.0; DDh[3] = -2.0;
	}
	// the 3-argument backup set of the non-TB implicit classes (h, Dh only)
	__device__ static void backup_set_rows(const double *x, double *h, double *Dh)
	{
		h[0] = backup_set_value(x);
		Dh[0] = -2.0 * x[0];
		Dh[1] = -2.0 * x[1];
	}
	__device__ static void safety_set_lower(const double *x, const double *unc, double *hl) { box_safety_lower(x, unc, -1.0, 1.0, hl); }
	__device__ static void backup_controller(const double *x, double *u, double *Du)
	{
		u[0] = (-10.0) * x[0] + (-20.0) * x[1];
		Du[0] = -10.0;
		Du[1] = -20.0;
	}
	__device__ static void dynamics(const double *x, double *f, double *g)
	{
		f[0] = x[1];
		f[1] = 0.0;
		g[0] = 0.0;
		g[1] = 1.0;
	}
	__device__ static void dynamics_with_gradient(const double *x, const double *u, double *f, double *g, double *d)
	{
		dynamics(x, f, g);
		d[0] = 0.0; d[1] = 0.0; d[2] = 1.0; d[3] = 0.0;
	}
	__device__ static void dynamics_gradients(const double *x, double *Df, double *Dg)
	{
		Df[0] = 0.0; Df[1] = 0.0; Df[2] = 1.0; Df[3] = 0.0;
		Dg[0] = Dg[1] = Dg[2] = Dg[3] = 0.0;
	}
	static constexpr bool HAS_DYNAMICS_ALL = false;
	__device__ static void dynamics_all(const double *, double *, double *, double *, double *) {}
};

// ------------------------------------------------------------------ Segway (implicit TB)
// examples/segway_implicit_tb.cpp:13-212.  CENTRED selects the backup set centred on the backup
// controller's equilibrium (SURVEY 8c deviation D6); false = the set as shipped.
template <bool CENTRED>
struct SegwayTB {
	static constexpr int NX = 4, NU = 1, NPSS = 4;
	static constexpr bool FUSED_GRADIENT = false;
	// structural patterns: g = {0,*,0,*}; Df rows 0 and 2 are unit rows, column 0 is zero (:145-201);
	// Dg is zero except entries 9 and 11 (:203-211)
	__host__ __device__ static constexpr int g_pat(int i)
	{
		constexpr int t[4] = {0, 2, 0, 2};
		return t[i];
	}
	__host__ __device__ static constexpr int df_pat(int i)
	{
		constexpr int t[16] = {0, 0, 0, 0, 1, 2, 0, 2, 0, 2, 0, 2, 0, 2, 1, 2};
		return t[i];
	}
	__host__ __device__ static constexpr int dg_pat(int i)
	{
		constexpr int t[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 0, 2, 0, 0, 0, 0};
		return t[i];
	}
	__host__ __device__ static constexpr int f_pat(int i)
	{
		constexpr int t[4] = {2, 2, 2, 2};
		return t[i];
	}
	// DhSS = diag(-2 x_i)
	__host__ __device__ static constexpr int dhs_pat(int i)
	{
		constexpr int t[16] = {2, 0, 0, 0, 0, 2, 0, 0, 0, 0, 2, 0, 0, 0, 0, 2};
		return t[i];
	}
	static constexpr bool HAS_SAFETY_MIN = false;
	__device__ static double safety_min(const double *) { return 0.0; }
	__device__ static double bound(int i)
	{
		return i == 0 ? 3.0 : (i == 1 ? 3.0 : (i == 2 ? (M_PI / 6) : M_PI));
	}
	__device__ static void safety_set(const double *x, double *h, double *Dh)
	{
#pragma unroll
		for (int i = 0; i < 16; i++) Dh[i] = 0.0;
#pragma unroll
		for (int i = 0; i < 4; i++) {
			h[i] = (bound(i) * bound(i)) - (x[i] * x[i]);
			Dh[i * 5] = -2.0 * x[i];
		}
	}
	__device__ static double backup_set_value(const double *xin)
	{
		double h = 0.05 * 0.05;
#pragma unroll
		for (int i = 0; i < 4; i++) {
			const double xi = (CENTRED && i == 2) ? (xin[i] - 0.1383244254) : (CENTRED ? (xin[i] - 0.0) : xin[i]);
			h -= (xi / bound(i)) * (xi / bound(i));
		}
		return h;
	}
	static constexpr bool HAS_BACKUP_SET_REACHED = false;
	__device__ static bool backup_set_reached(const double *x) { return backup_set_value(x) >= 0.0; }
	__device__ static double backup_set_level() { return 0.0; }
	__device__ static bool backup_set_reached(const double *x, const double) { return backup_set_reached(x); }
	__device__ static void backup_set(const double *xin, double &h, double *Dh, double *DDh)
	{
		h = backup_set_value(xin);
#pragma unroll
		for (int i = 0; i < 16; i++) DDh[i] = 0.0;
#pragma unroll
		for (int i = 0; i < 4; i++) {
			const double xi = (CENTRED && i == 2) ? (xin[i] - 0.1383244254) : (CENTRED ? (xin[i] - 0.0) : xin[i]);
			Dh[i] = -2.0 * xi / (bound(i) * bound(i));
			DDh[i * 5] = -2.0 / (bound(i) * bound(i));
		}
	}
	__device__ static void backup_controller(const double *x, double *u, double *Du)
	{
		const double K[4] = {44.7214, 44.6528, 150.1612, 37.6492};
		double xt[4] = {0., 0., -0.1383244254, 0.};
#pragma unroll
		for (int i = 0; i < 4; i++) xt[i] += x[i];
		double s = K[0] * xt[0];
#pragma unroll
		for (int k = 1; k < 4; k++) s = s + K[k] * xt[k];
		u[0] = s;
#pragma unroll
		for (int k = 0; k < 4; k++) Du[k] = K[k];
	}
#ifndef ASIF_SEGWAY_COMPACT
#define ASIF_SEGWAY_COMPACT 1
#endif
#ifndef ASIF_SEGWAY_DOUBLE_ANGLE
#define ASIF_SEGWAY_DOUBLE_ANGLE 1 // sin 2t, cos 2t from sin t, cos t: C5 9.86 -> 9.38 ms, rows 1.25e-10 -> 1.30e-10 off the oracle
#endif
#ifndef ASIF_SEGWAY_TANH_FAST
#define ASIF_SEGWAY_TANH_FAST 0 // measured neutral (9.353 vs 9.327 ms per 1e6 C5 states): CUDA's tanh has its own early exit
#endif
	// sin/cos of theta and 2 theta are shared by f, g and their gradients: one sincos pair per evaluation point
	static constexpr bool HAS_DYNAMICS_ALL = true;
	__device__ static void dynamics_all(const double *x, double *f, double *g, double *Df, double *Dg)
	{
		double s1, c1, s2, c2;
		sincos_model(x[2], &s1, &c1);
#if ASIF_SEGWAY_DOUBLE_ANGLE
		s2 = 2.0 * s1 * c1;
		c2 = fma(-2.0 * s1, s1, 1.0);
#else
		sincos_model(2.0 * x[2], &s2, &c2);
#endif
#if ASIF_SEGWAY_COMPACT
		all_compact(x, s1, c1, s2, c2, f, g, Df, Dg);
#else
		dynamics_core(x, s1, c1, s2, c2, f, g);
		gradients_core(x, s1, c1, s2, c2, Df, Dg);
#endif
	}
	// The same model with the generated expressions of dynamics_core / gradients_core collected: literals that multiply the
	// same monomial are folded at compile time (f[1] alone carried 14 of them for four monomials), the seven reciprocals
	// are two (1/D shared by f, g[3], every Df entry and Dg[11] - the generated gradient's own denominator is -D with a
	// literal that differs in the 16th digit - and 1/G shared by g[1] and Dg[9]), and (tanh^2 - 1) is factored out of the
	// friction terms.  Mathematically identical, rounded differently (a few ulp per entry): this translation unit is
	// compiled with FMA contraction and uses its own sincos, so it is in the tolerance class anyway (rows within 1e-9 of the
	// reference build; tests/test_gpu_parity.py, the golden rollout, scripts/parity_report.py C5).  The plant step of the
	// rollout (dynamics(), once per control step) keeps the generated form.
	__device__ static void all_compact(const double *x, const double s1, const double c1, const double s2, const double c2,
	                                   double *f, double *g, double *Df, double *Dg)
	{
		constexpr double B1 = 0.0975 * (44.798 * ((-0.2693850964936445 + -0.0022454764220255392) + -0.11586336477125109) * 0.195);
		constexpr double B2 = 0.0975 * 59.510408935182809;
		constexpr double B3 = 0.0975 * (((86.686408318784913 + 0.72258001100852454) + 37.284092841364554) * 0.195);
		constexpr double B4 = 0.0975 * (4.1423245261005457 + -213.73800805067131);
		constexpr double C1 = 89.596 * -0.45669752988922296;
		constexpr double C2 = 15.554616935932147 * 0.038025;
		constexpr double C3 = (16.405863695295427 + 249.80488266222164) + 27.713966400983114;
		constexpr double C4 = (1.0827059060875992 + -55.866072832711595) * 0.038025;
		constexpr double E0 = 0.551244194154502 * 4.1706936767483551, E1 = 0.551244194154502 * 1.4575004011882324,
		                 E2 = 0.551244194154502 * 0.20290365220710288;
		constexpr double F0 = -5.65378660671284 * 2.0043013906215941, F1 = -5.65378660671284 * 1.4575004011882324,
		                 F2 = -5.65378660671284 * 0.20290365220710288;
		const double w = x[3], w2 = w * w;
		const double D = (14.553176960783997 + -2.0831375273848773 * c2) + -0.59146430898882 * s2;
		const double den = 1.0 / D;
		const double cc1 = c1 * c1, ss1 = s1 * s1;
		const double G = ((8.3593271361634187 + -2.1243074194638587 * cc1) + -0.04116989207898096 * ss1) + -0.29573215449441 * s2;
		const double ginv = 1.0 / G;
		f[0] = x[1];
		f[1] = den * (w2 * (B1 * c1 + B3 * s1) + (B2 * c2 + B4 * s2));
		f[2] = w;
		f[3] = den * ((C1 * c1 + C3 * s1) + w2 * (C2 * c2 + C4 * s2));
		g[0] = 0.0;
		g[1] = ((E0 + E1 * c1) + E2 * s1) * ginv;
		g[2] = 0.0;
		g[3] = ((F0 + F1 * c1) + F2 * s1) * den;
		// gradients; t23 of the generated code is -den.  tanh(1000 v): |argument| >= 20 rounds to +-1 exactly in double
		// (1 - tanh 20 = 8.5e-18 < half an ulp of 1, in glibc and in CUDA alike), which is every state whose wheel speed is
		// not within 0.02 of zero: those lanes skip the library call
		const double tz = x[1] * 1000.0;
#if ASIF_SEGWAY_TANH_FAST
		double th = copysign(1.0, tz);
		if (fabs(tz) < 20.0) th = tanh(tz);
#else
		const double th = tanh(tz);
#endif
		const double q = th * th - 1.0;
		const double t25 = th * 15.13175750513302 - 40.918271887954823;
		const double t26s = w2 * (3.3849959169972448 * 0.0975) + th * (30.26351501026604 * 0.0975);
		const double den2 = den * den;
		Df[0] = 0.0; Df[1] = 0.0; Df[2] = 0.0; Df[3] = 0.0;
		Df[4] = 1.0;
		Df[5] = den * q * ((8443.5211353581435 + s1 * (1000.0 * 0.41077609832706019)) + c1 * (30263.515010266041 * 0.0975));
		Df[6] = 0.0;
		Df[7] = -den * q * ((20808.641003022261 + s1 * (1000.0 * 2.1065440939849238)) + c1 * 15131.75750513302);
		Df[8] = 0.0;
		const double cth = c1 * th, sth = s1 * th;
		const double cc = (c2 * 1.18292861797764 - s2 * 4.1662750547697547) * den2;
		Df[9] = -den * ((((c2 * 40.8711582872913 + s2 * 11.604529742360651) - c1 * w2 * 2.3707272057666411) + cth * 0.41077609832706019) -
		                s1 * t26s) -
		        cc * (((((c2 * -5.8022648711803244 + s2 * 20.435579143645651) + th * 8.443521135358143) - s1 * w2 * 2.3707272057666411) +
		               sth * 0.41077609832706019) +
		              c1 * t26s);
		Df[10] = 0.0;
		const double wc = w2 * c2, ws = w2 * s2;
		Df[11] = -den * ((((c1 * -293.92471275850022 - cth * 2.1065440939849238) + wc * 4.1662750547697547) + ws * 1.18292861797764) +
		                 s1 * t25) +
		         cc * (((((s1 * 293.92471275850022 + th * 20.808641003022259) + wc * 0.59146430898881985) + sth * 2.1065440939849238) -
		                ws * 2.0831375273848769) +
		               c1 * t25);
		Df[12] = 0.0;
		const double dw = den * w;
		Df[13] = -dw * (c1 * 0.6600742038144628 - s1 * 4.7414544115332831);
		Df[14] = 1.0;
		Df[15] = dw * (c2 * 1.18292861797764 - s2 * 4.1662750547697547);
#pragma unroll
		for (int i = 0; i < 16; i++) Dg[i] = 0.0;
		// d26 of the generated code is -G, d4 is -D
		Dg[9] = (c1 * 0.1118494602519098 - s1 * 0.80343863413287053) * ginv +
		        (ginv * ginv) * (c2 * 0.59146430898882 - c1 * s1 * 4.1662750547697547) *
		            ((c1 * 0.80343863413287053 + s1 * 0.1118494602519098) + 2.2990706749044238);
		Dg[11] = -(c1 * 1.1471739513016379 - s1 * 8.24039624751662) * den -
		         den2 * (c2 * 1.18292861797764 - s2 * 4.1662750547697547) *
		             ((c1 * 8.24039624751662 + s1 * 1.1471739513016379) + 11.33189235811229);
	}
	__device__ static void dynamics(const double *X, double *f, double *g)
	{
		double s1, c1, s2, c2;
		sincos(X[2], &s1, &c1);
		sincos(2.0 * X[2], &s2, &c2);
		dynamics_core(X, s1, c1, s2, c2, f, g);
	}
	__device__ static void dynamics_core(const double *X, const double s1, const double c1, const double s2, const double c2,
	                                     double *f, double *g)
	{
		// The friction term of the shipped model carries a literal factor 0.0 (:80), so every
		// product with it is an exact zero; those terms are dropped (adding +-0.0 to a non-zero
		// partial sum is the identity in IEEE arithmetic).
		f[0] = X[1];
		const double w2 = X[3] * X[3];
		const double den = 1.0 / ((14.553176960783997 + -2.0831375273848773 * c2) + -0.59146430898882 * s2);
		f[1] = 0.0975 *
		       (((((((44.798 *
		                  ((-0.2693850964936445 * w2 + -0.0022454764220255392 * w2) + -0.11586336477125109 * w2) *
		                  0.195 * c1 +
		              59.510408935182809 * c2) +
		             86.686408318784913 * w2 * 0.195 * s1) +
		            0.72258001100852454 * w2 * 0.195 * s1) +
		           37.284092841364554 * w2 * 0.195 * s1) +
		          4.1423245261005457 * s2) +
		         -213.73800805067131 * s2)) *
		       den;
		f[2] = X[3];
		f[3] = den * (((((((89.596 * -0.45669752988922296 * c1 + 15.554616935932147 * w2 * 0.038025 * c2) +
		                   16.405863695295427 * s1) +
		                  249.80488266222164 * s1) +
		                 27.713966400983114 * s1) +
		                1.0827059060875992 * w2 * 0.038025 * s2) +
		               -55.866072832711595 * w2 * 0.038025 * s2));
		g[0] = 0.0;
		const double a = 1.4575004011882324 * c1;
		const double b = 0.20290365220710288 * s1;
		g[1] = 0.551244194154502 * ((4.1706936767483551 + a) + b) *
		       (1.0 / (((8.3593271361634187 + -2.1243074194638587 * (c1 * c1)) + -0.04116989207898096 * (s1 * s1)) +
		               -0.29573215449441 * s2));
		g[2] = 0.0;
		g[3] = -5.65378660671284 * ((2.0043013906215941 + a) + b) * den;
	}
	__device__ static void dynamics_gradients(const double *x, double *Df, double *Dg)
	{
		double s1, c1, s2, c2;
		sincos(x[2], &s1, &c1);
		sincos(x[2] * 2.0, &s2, &c2);
		gradients_core(x, s1, c1, s2, c2, Df, Dg);
	}
	__device__ static void gradients_core(const double *x, const double s1, const double c1, const double s2, const double c2,
	                                      double *Df, double *Dg)
	{
		const double w2 = x[3] * x[3];
		const double th = tanh(x[1] * 1000.0);
		const double th2 = th * th;
		const double t25 = th * 15.13175750513302 - 40.918271887954823;
		const double t26 = w2 * 3.3849959169972448 + th * 30.26351501026604;
		const double t23 = 1.0 / ((c2 * 2.0831375273848769 + s2 * 0.59146430898882) - 14.553176960784);
		Df[0] = 0.0; Df[1] = 0.0; Df[2] = 0.0; Df[3] = 0.0;
		Df[4] = 1.0;
		const double p = s1 * (th2 * 1000.0 - 1000.0);
		Df[5] = -t23 * (((th2 * 8443.5211353581435 + p * 0.41077609832706019) +
		                 c1 * (th2 * 30263.515010266041 - 30263.515010266041) * 0.0975) -
		                8443.5211353581435);
		Df[6] = 0.0;
		Df[7] = t23 * (((th2 * 20808.641003022261 + p * 2.1065440939849238) +
		                c1 * (th2 * 15131.75750513302 - 15131.75750513302)) -
		               20808.641003022261);
		Df[8] = 0.0;
		const double cth = c1 * th;
		const double sth = s1 * th;
		const double cc = (c2 * 1.18292861797764 + -(s2 * 4.1662750547697547)) * (t23 * t23);
		Df[9] = t23 * ((((c2 * 40.8711582872913 + s2 * 11.604529742360651) - c1 * w2 * 2.3707272057666411) +
		                cth * 0.41077609832706019) -
		               s1 * t26 * 0.0975) -
		        cc * (((((c2 * -5.8022648711803244 + s2 * 20.435579143645651) + th * 8.443521135358143) -
		                s1 * w2 * 2.3707272057666411) +
		               sth * 0.41077609832706019) +
		              c1 * t26 * 0.0975);
		Df[10] = 0.0;
		const double wc = w2 * c2;
		const double ws = w2 * s2;
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
		const double d4 = (c2 * 2.0831375273848769 + s2 * 0.59146430898882) - 14.553176960784;
		const double d26 =
		    ((c1 * c1 * 2.1243074194638591 + s2 * 0.29573215449441) + s1 * s1 * 0.04116989207898096) - 8.3593271361634187;
#pragma unroll
		for (int i = 0; i < 16; i++) Dg[i] = 0.0;
		Dg[9] = -(c1 * 0.1118494602519098 - s1 * 0.80343863413287053) / d26 +
		        1.0 / (d26 * d26) * (c2 * 0.59146430898882 - c1 * s1 * 4.1662750547697547) *
		            ((c1 * 0.80343863413287053 + s1 * 0.1118494602519098) + 2.2990706749044238);
		Dg[11] = (c1 * 1.1471739513016379 - s1 * 8.24039624751662) / d4 -
		         1.0 / (d4 * d4) * (c2 * 1.18292861797764 - s2 * 4.1662750547697547) *
		             ((c1 * 8.24039624751662 + s1 * 1.1471739513016379) + 11.33189235811229);
	}
	__device__ static void dynamics_with_gradient(const double *, const double *, double *, double *, double *) {}
};

// ------------------------------------------------------------------ InvertedPendulum (implicit)
// examples/InvertedPendulum_Implicit.cpp:13-80
struct InvertedPendulumImplicit {
	static constexpr int NX = 2, NU = 1, NPSS = 4, NPBS = 1;
	static constexpr bool FUSED_GRADIENT = false;
	// g = {0,1}; Df = {0, cos x0, 1, 0} column-major; Dg = 0
	__host__ __device__ static constexpr int g_pat(int i)
	{
		constexpr int t[2] = {0, 1};
		return t[i];
	}
	__host__ __device__ static constexpr int df_pat(int i)
	{
		constexpr int t[4] = {0, 2, 1, 0};
		return t[i];
	}
	__host__ __device__ static constexpr int dg_pat(int i)
	{
		constexpr int t[4] = {0, 0, 0, 0};
		return t[i];
	}
	__host__ __device__ static constexpr int f_pat(int i)
	{
		constexpr int t[2] = {2, 2};
		return t[i];
	}
	// DhSS column-major {-1,1,0,0 | 0,0,1,-1}
	__host__ __device__ static constexpr int dhs_pat(int i)
	{
		constexpr int t[8] = {3, 1, 0, 0, 0, 0, 1, 3};
		return t[i];
	}
	// box +-pi on both components: min_j h_j = pi - max(|x0|, |x1|), same bits as min_element over h
	static constexpr bool HAS_SAFETY_MIN = true;
	__device__ static double safety_min(const double *x)
	{
		const double m = (fabs(x[0]) > fabs(x[1])) ? x[0] : x[1];
		return M_PI - fabs(m);
	}
	__device__ static void safety_set(const double *x, double *h, double *Dh)
	{
		h[0] = -x[0] + M_PI;   Dh[0] = -1.0; Dh[4] = 0.0;
		h[1] = x[0] - (-M_PI); Dh[1] = 1.0;  Dh[5] = 0.0;
		h[2] = x[1] - (-M_PI); Dh[2] = 0.0;  Dh[6] = 1.0;
		h[3] = -x[1] + M_PI;   Dh[3] = 0.0;  Dh[7] = -1.0;
	}
	__device__ static void safety_set_lower(const double *x, const double *unc, double *hl) { box_safety_lower(x, unc, -M_PI, M_PI, hl); }
	// h = Pv - x'Px, P = {1.25,.25,.25,.25}, Pv = 0.05 (:39-52): the four products in loop order
	__device__ static void backup_set_rows(const double *x, double *h, double *Dh)
	{
		double v = 0.05;
		v -= 1.25 * x[0] * x[0];
		v -= 0.25 * x[0] * x[1];
		v -= 0.25 * x[1] * x[0];
		v -= 0.25 * x[1] * x[1];
		h[0] = v;
		Dh[0] = (-2.5) * x[0] + (-0.5) * x[1];
		Dh[1] = (-0.5) * x[0] + (-0.5) * x[1];
	}
	__device__ static void backup_controller(const double *x, double *u, double *Du)
	{
		u[0] = (-3.0) * x[0] + (-3.0) * x[1];
		Du[0] = -3.0;
		Du[1] = -3.0;
	}
	__device__ static void dynamics(const double *x, double *f, double *g)
	{
		f[0] = x[1];
		f[1] = sin(x[0]);
		g[0] = 0.;
		g[1] = 1.;
	}
	__device__ static void dynamics_gradients(const double *x, double *Df, double *Dg)
	{
		Df[0] = 0.;        Df[2] = 1.;
		Df[1] = cos(x[0]); Df[3] = 0.;
		Dg[0] = Dg[1] = Dg[2] = Dg[3] = 0.0;
	}
	// sin x0 (drift) and cos x0 (its gradient) share one argument reduction
	static constexpr bool HAS_DYNAMICS_ALL = true;
	__device__ static void dynamics_all(const double *x, double *f, double *g, double *Df, double *Dg)
	{
		double s, c;
		sincos_model(x[0], &s, &c);
		f[0] = x[1];
		f[1] = s;
		g[0] = 0.;
		g[1] = 1.;
		Df[0] = 0.; Df[2] = 1.;
		Df[1] = c;  Df[3] = 0.;
		Dg[0] = Dg[1] = Dg[2] = Dg[3] = 0.0;
	}
	// the same with sin x0, cos x0 supplied by the caller (filter_common.cuh: TrigSC)
	static constexpr bool HAS_TRIG_STATE = true;
	static constexpr int TRIG_ANGLE = 0;
	__device__ static void dynamics_all_sc(const double *x, const double s, const double c, double *f, double *g, double *Df, double *Dg)
	{
		f[0] = x[1];
		f[1] = s;
		g[0] = 0.;
		g[1] = 1.;
		Df[0] = 0.; Df[2] = 1.;
		Df[1] = c;  Df[3] = 0.;
		Dg[0] = Dg[1] = Dg[2] = Dg[3] = 0.0;
	}
	__device__ static void dynamics_with_gradient(const double *, const double *, double *, double *, double *) {}
};

} // namespace asifb
