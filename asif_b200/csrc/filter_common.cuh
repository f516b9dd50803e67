// filter_common.cuh -- parameter blocks and device helpers shared by the filter kernels.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace asifb {

constexpr int MAX_NU = 2;
constexpr int MAX_NV = 4;

// Soft input saturation constants (src/asif_implicit_tb.cpp:764-785).  Everything that does not
// depend on the state is evaluated once on the host with the same double operations (and the
// host libm's tan / cos, as the reference does), so the device never calls tan() or cos() here.
struct SoftSat {
	double r;          // satSharpness
	double r2;         // r*r
	double bevelStart; // 1 - cos(pi/4) * (r*tan(pi/8))
	double bevelStop;  // 1 + r*tan(pi/8)      (= bevelXc)
	double bevelYc;    // 1 - r
	double range[MAX_NU];  // ub - lb
	double middle[MAX_NU]; // (ub + lb)/2
	// 2/range when range is a power of two (then 2*(u-mid)/range == (u-mid)*(2/range) bit for bit and the
	// FP64 division, ~10 dependent DFMAs per Euler step, is not needed); 0 otherwise
	double uc_scale_exact[MAX_NU];
	double uc_scale[MAX_NU]; // 2/range rounded (SAT_RECIP)
};

// How the normalised input uc = 2 (u - middle) / range is evaluated; chosen on the host from (lb, ub).
// The first three modes produce the same bits as the reference expression.
enum : int {
	SAT_GENERAL = 0,  // 2*(u-middle)/range
	SAT_POW2 = 1,     // range is a power of two: (u-middle)*(2/range), all operations exact
	SAT_IDENTITY = 2, // middle == 0 and range == 2 (e.g. u in [-1,1]): uc == u
	SAT_RECIP = 3,    // (u-middle)*fl(2/range): last-bit differences; only for models without bit parity (kernels_contract.cu)
	SAT_RECIP_SYM = 4 // SAT_RECIP with lb == -ub (middle == 0): uc = u*fl(2/range), bound = copysign(ub, uc) - the same bits as
	                  // SAT_RECIP (u - 0 == u, +-ub are the bounds themselves) without the subtraction and the select of two loads
};

// Structural patterns a model may declare for its callback outputs: an entry that is the literal
// 0.0 or 1.0 for every state.  The kernels skip the multiplications by such entries; for finite
// operands this cannot change an IEEE result (x*0 = +-0, acc + +-0 = acc, 1*x = x) except for the
// sign of an exact zero.  PG = general (no assumption).
enum : int { PZ = 0, P1 = 1, PG = 2, PM1 = 3 }; // PM1: the literal -1.0 (only used for the safety-set Jacobian)

// npBTSS as a template value: N > 0 is the count itself (everything unrolled: the tuned instantiations, 4 for the TB
// examples, 10 for the implicit one); N < 0 is a list CAPACITY of -N with the count taken from the options at run time (1 <= npBTSS <= -N).  The
// first npBTSS entries of an ascending list of the -N smallest points are the npBTSS smallest points in the same
// order and with the same tie handling, so the run-time variant only has to look at fewer entries when it builds rows.
__host__ __device__ constexpr int np_capacity(const int n) { return n > 0 ? n : -n; }
__host__ __device__ constexpr bool np_runtime(const int n) { return n < 0; }

// Options of ASIFimplicitTB (include/asif_implicit_tb.h:19-33) + what initialize() derives from them
// (src/asif_implicit_tb.cpp:169-223).
struct TbParams {
	double lb[MAX_NU], ub[MAX_NU];
	double relaxCost, relaxSafeLb, relaxTTS, relaxMinOrtho;
	double backTrajHorizon, backTrajDt, backTrajMinOrtho, inf;
	int32_t npBT; // trajectory points (npBT-1 Euler steps)
	int32_t sat_mode;
	int32_t npBTSS; // critical trajectory points (read by the run-time-count instantiations only)
	// filter(x, H, c, ...) overloads (src/asif_implicit_tb.cpp:252-259): != 0 means the u_des argument of the kernel
	// holds the caller's linear cost c[n][nv] instead of uDes[n][nu] (H enters through gi / gih)
	int32_t custom_cost;
	const double *t_of_index; // device: t_0 = 0, t_i = t_{i-1} + dt (src/asif_implicit_tb.cpp:465,475)
	SoftSat sat;
	// QP metric: gi = 1/(2 H_ii), gih = sqrt(gi) for v = (u, relax)
	double gi[MAX_NV], gih[MAX_NV];
};

// Options of ASIF (include/asif.h:11-17) as used by initialize() (src/asif.cpp:64-110)
struct ExplicitParams {
	double lb[MAX_NU], ub[MAX_NU];
	double relaxLb, relaxCost;
	double gi[MAX_NV], gih[MAX_NV];
	int32_t npSSmax; // rows kept (min(npSSmax, npSS), src/asif.cpp:21-22)
	int32_t custom_cost; // filter(x, H, c, ...) (src/asif.cpp:153-174): u_des holds c[n][nv]
	// filter(x, uDes, uAct, Lfh, Lgh[, relax]) (src/asif.cpp:125-141, 287-292): caller-supplied Lie derivatives per state,
	// Lfh[n][nc], Lgh[n][nc x nu column-major], replacing the computed ones row for row; nullptr = computed
	const double *lfh, *lgh;
};

// Engine counters (device, unsigned long long): [0] unused, [1..8] return-code histogram of the rollouts, and from
// QP_CTR_BASE on QP_CTR_SPREAD partial sums of "rows processed by the active-set solver" (the K-bar statistic).  One
// warp-aggregated atomicAdd per warp to ONE address was the bottleneck of the explicit filter at 1e8 states (3.1e6
// same-address atomics serialise in L2 at about 1.2 ns each = the whole 3.7 ms of the launch); spread over 256 addresses
// by block index they no longer queue.  The host sums the partials (asif_engine_last_qp_iterations).
constexpr int QP_CTR_BASE = 16, QP_CTR_SPREAD = 256, N_COUNTERS = QP_CTR_BASE + QP_CTR_SPREAD;
__device__ __forceinline__ void qp_rows_add(unsigned long long *counters, const unsigned long long it)
{
	atomicAdd(counters + QP_CTR_BASE + (blockIdx.x & (QP_CTR_SPREAD - 1)), it);
}

// src/asif_implicit_tb.cpp:821-830
__device__ __forceinline__ double input_saturate(double u, double lb, double ub)
{
	if (u > ub) return ub;
	if (u < lb) return lb;
	return u;
}

// src/asif_implicit_tb.cpp:764-819, one input.  Same decisions and operation order; the five-way branch is
// folded onto |uc| (two compares instead of four; a NaN input takes the pass-through leg as in the reference).
template <int SATMODE>
__device__ __forceinline__ void input_saturate_soft(const SoftSat &s, int i, double lb, double ub, double u, double &uSat,
                                                    double &DuSat)
{
	const double range = s.range[i], middle = s.middle[i];
	double uc;
	if (SATMODE == SAT_IDENTITY) uc = u;
	else if (SATMODE == SAT_POW2) uc = (u - middle) * s.uc_scale_exact[i];
	else if (SATMODE == SAT_RECIP) uc = (u - middle) * s.uc_scale[i];
	else if (SATMODE == SAT_RECIP_SYM) uc = u * s.uc_scale[i];
	else uc = 2 * (u - middle) / range;
	const double a = fabs(uc);
	// clamp and pass-through legs as selects (no branch, no register shuffling between the legs) ...
	const bool sat = a >= s.bevelStop;
	const bool bevel = !sat && (a > s.bevelStart); // false for NaN: pass through, as the reference's last else
	// SAT_IDENTITY means lb = -1, ub = 1: the bound is the sign of uc on 1.0 (uc = 0 never saturates)
	const double bound = (SATMODE == SAT_IDENTITY) ? copysign(1.0, uc) : ((SATMODE == SAT_RECIP_SYM) ? copysign(ub, uc) : ((uc > 0) ? ub : lb));
	uSat = sat ? bound : u;
	DuSat = sat ? 0.0 : 1.0;
	// ... the circular bevel (sqrt + division) is rare: a real branch, entered only by the lanes that need it
	if (bevel) {
		// uc > bevelStart: d = uc - Xc ; uc < -bevelStart: d = uc + Xc.  (uc -+ Xc)^2 is even in the sign flip,
		// so both legs of the reference share sq; the signs below reproduce each leg's expressions exactly.
		const double d = (uc > 0) ? (uc - s.bevelStop) : (uc + s.bevelStop);
		const double sq = sqrt(s.r2 - d * d);
		const double y = (uc > 0) ? (sq + s.bevelYc) : (-sq - s.bevelYc);
		DuSat = ((uc > 0) ? (s.bevelStop - uc) : (s.bevelStop + uc)) / sq;
		uSat = 0.5 * y * range + middle;
	}
}

// structural pattern of DfCL = d_fcl_dx (or Df + Dg uSat) + g diag(DuSat) Du, entry (i,j)
template <class M>
__host__ __device__ constexpr int dfcl_pattern(int i, int j)
{
	bool g_zero = true;
	for (int k = 0; k < M::NU; k++) g_zero = g_zero && (M::g_pat(i + k * M::NX) == PZ);
	bool dg_zero = true;
	if (!M::FUSED_GRADIENT)
		for (int k = 0; k < M::NU; k++) dg_zero = dg_zero && (M::dg_pat(i + k * M::NX + j * M::NX * M::NU) == PZ);
	if (g_zero && dg_zero) return M::df_pat(i + j * M::NX);
	return PG;
}

// Closed-loop backup dynamics and their Jacobian, src/asif_implicit_tb.cpp:833-897.
// Entries of DfCL whose dfcl_pattern is PZ / P1 are not written (callers must not read them).
// f_pat(i) == PZ marks a drift component that is the literal 0.0 (then g uSat + 0.0 == g uSat).
// The part after the backup controller: u is the input handed to the saturation, Du the controller Jacobian that
// enters DfCL (for ASIFimplicitRB these can be the zero-order-held values, see backup_cl_dynamics_zoh).
// Optional carrier of sin / cos of the model's angle for the Euler loops that advance them by the angle-addition recurrence
// (implicit_kernel.cuh, ASIF_IMP_TRIG_RECURRENCE); NoTrig = the model evaluates its own trigonometry.
struct NoTrig {};
struct TrigSC {
	double s, c, x0;
};
template <class M>
__device__ __forceinline__ void model_dynamics_all(const double *x, const NoTrig &, double *f, double *g, double *d, double *Dg)
{
	M::dynamics_all(x, f, g, d, Dg);
}
template <class M>
__device__ __forceinline__ void model_dynamics_all(const double *x, const TrigSC &t, double *f, double *g, double *d, double *Dg)
{
	M::dynamics_all_sc(x, t.s, t.c, f, g, d, Dg);
}

template <class M, int SATMODE, class TR = NoTrig>
__device__ __forceinline__ void backup_cl_from_input(const SoftSat &sat, const double *lb, const double *ub, const double *x,
                                                     const double *u, const double *Du, double *fCL, double *DfCL,
                                                     const TR &tr = TR())
{
	constexpr int NX = M::NX, NU = M::NU;
	double f[NX], g[NX * NU], uSat[NU], DuSat[NU];
	double d[NX * NX], Dg[M::FUSED_GRADIENT ? 1 : NX * NU * NX];
	// control-affine models: f, g and their gradients do not depend on the input, so they are evaluated BEFORE the soft
	// saturation - its rare bevel branch then comes after the model's long chains instead of fencing them off
	if (!M::FUSED_GRADIENT) {
		if (M::HAS_DYNAMICS_ALL) {
			model_dynamics_all<M>(x, tr, f, g, d, Dg); // shares sub-expressions (trig of the same argument) between f, g and Df, Dg
		} else {
			M::dynamics(x, f, g);
			M::dynamics_gradients(x, d, Dg);
		}
	}
#pragma unroll
	for (int k = 0; k < NU; k++) input_saturate_soft<SATMODE>(sat, k, lb[k], ub[k], u[k], uSat[k], DuSat[k]);
	if (M::FUSED_GRADIENT) M::dynamics_with_gradient(x, uSat, f, g, d);
#pragma unroll
	for (int i = 0; i < NX; i++)
#pragma unroll
		for (int j = 0; j < NX; j++) {
			if (dfcl_pattern<M>(i, j) != PG) continue;
			// DfCL = d; DfCL += Dg*uSat + g*DuSat*Du   (:857-888), structural zeros skipped
			const int pd = M::df_pat(i + j * NX);
			bool have = (pd != PZ);
			double acc = (pd == P1) ? 1.0 : d[i + j * NX];
#pragma unroll
			for (int k = 0; k < NU; k++) {
				const int pg = M::g_pat(i + k * NX);
				bool tz = true;
				double term = 0.0;
				if (!M::FUSED_GRADIENT) {
					if (M::dg_pat(i + k * NX + j * NX * NU) != PZ) {
						term = Dg[i + k * NX + j * NX * NU] * uSat[k];
						tz = false;
					}
				}
				if (pg != PZ) {
					const double gd = (pg == P1) ? DuSat[k] : g[i + k * NX] * DuSat[k];
					const double t2 = gd * Du[k + j * NU];
					term = tz ? t2 : term + t2;
					tz = false;
				}
				if (!tz) {
					acc = have ? acc + term : term;
					have = true;
				}
			}
			DfCL[i + j * NX] = acc;
		}
	// fCL = g uSat + f   (matrixVectorMultiply then += f, :892-896)
#pragma unroll
	for (int i = 0; i < NX; i++) {
		bool have = false;
		double acc = 0.0;
#pragma unroll
		for (int k = 0; k < NU; k++) {
			const int pg = M::g_pat(i + k * NX);
			if (pg == PZ) continue;
			const double t = (pg == P1) ? uSat[k] : g[i + k * NX] * uSat[k];
			acc = have ? acc + t : t;
			have = true;
		}
		fCL[i] = have ? ((M::f_pat(i) == PZ) ? acc : acc + f[i]) : f[i];
	}
}

template <class M, int SATMODE, class TR = NoTrig>
__device__ __forceinline__ void backup_cl_dynamics(const SoftSat &sat, const double *lb, const double *ub, const double *x,
                                                   double *fCL, double *DfCL, const TR &tr = TR())
{
	double u[M::NU], Du[M::NU * M::NX];
	M::backup_controller(x, u, Du);
	backup_cl_from_input<M, SATMODE, TR>(sat, lb, ub, x, u, Du, fCL, DfCL, tr);
}

// Zero-order-hold backup controller of ASIFimplicitRB (src/asif_implicit_robust.cpp:878-903): the controller is
// evaluated at every rhs call, but the input that is saturated and applied is refreshed only when
// t >= t_last + backContDt - 0.0001; t_last is reset to -1 while t <= backTrajDt, i.e. on the first Euler step of
// every trajectory, so the hold carries no state from one filter() call to the next as long as backContDt < 1
// (enforced at engine creation).  Quirk kept: the fused-gradient branch builds DfCL from the HELD Du (:921), the
// split branch from the CURRENT one (:939).
template <class M>
struct ZohState {
	double u[M::NU], Du[M::NU * M::NX];
	double t_last;
};

template <class M, int SATMODE, class TR = NoTrig>
__device__ __forceinline__ void backup_cl_dynamics_zoh(const SoftSat &sat, const double *lb, const double *ub, const double *x,
                                                       const double t, const double backTrajDt, const double backContDt,
                                                       ZohState<M> &z, double *fCL, double *DfCL, const TR &tr = TR())
{
	constexpr int NX = M::NX, NU = M::NU;
	double u[NU], Du[NU * NX];
	M::backup_controller(x, u, Du);
	if (t <= backTrajDt) z.t_last = -1.;
	if (t >= (z.t_last + backContDt - 0.0001)) {
#pragma unroll
		for (int i = 0; i < NU; i++) z.u[i] = u[i];
#pragma unroll
		for (int i = 0; i < NU * NX; i++) z.Du[i] = Du[i];
		z.t_last = t;
	}
	backup_cl_from_input<M, SATMODE, TR>(sat, lb, ub, x, z.u, M::FUSED_GRADIENT ? z.Du : Du, fCL, DfCL, tr);
}

// Q-dot = DfCL Q (src/asif_implicit_tb.cpp:906-908) with the structural pattern of DfCL
template <class M>
__device__ __forceinline__ void sensitivity_rhs(const double *DfCL, const double *Q, double *Qd)
{
	constexpr int NX = M::NX;
#pragma unroll
	for (int r = 0; r < NX; r++)
#pragma unroll
		for (int c = 0; c < NX; c++) {
			bool have = false;
			double acc = 0.0;
#pragma unroll
			for (int m = 0; m < NX; m++) {
				const int pt = dfcl_pattern<M>(r, m);
				if (pt == PZ) continue;
				const double t = (pt == P1) ? Q[m + c * NX] : DfCL[r + m * NX] * Q[m + c * NX];
				acc = have ? acc + t : t;
				have = true;
			}
			Qd[r + c * NX] = acc;
		}
}

// Safety rows of one critical point (src/asif_implicit_tb.cpp:567-585, :643-653): h_j, Dh_row = DhSS(x_i) Q_i,
// Lfh_j = Dh_row f, Lgh_j = Dh_row g.  xs = [x_i; Q_i] (column-major Q), f, g = open-loop dynamics at the CURRENT state.
// Structural patterns of DhSS (dhs_pat), f (f_pat) and g (g_pat) remove the products with literal 0 / +-1 entries;
// what remains is evaluated in the reference's order.
template <class M>
__device__ __forceinline__ void safety_point_rows(const double *xs, const double *f, const double *g, double *hs, double *lf,
                                                  double *lg /* [NPSS][NU] */)
{
	constexpr int NX = M::NX, NU = M::NU, NPSS = M::NPSS;
	double Dhs[NPSS * NX];
	M::safety_set(xs, hs, Dhs);
#pragma unroll
	for (int j = 0; j < NPSS; j++) {
		double dh[NX];
		bool dhz[NX]; // structurally zero
#pragma unroll
		for (int cc = 0; cc < NX; cc++) {
			bool have = false;
			double acc = 0.0;
#pragma unroll
			for (int m = 0; m < NX; m++) {
				const int pt = M::dhs_pat(j + m * NPSS);
				if (pt == PZ) continue;
				const double q = xs[NX + m + cc * NX];
				const double t = (pt == P1) ? q : ((pt == PM1) ? -q : Dhs[j + m * NPSS] * q);
				acc = have ? acc + t : t;
				have = true;
			}
			dh[cc] = acc;
			dhz[cc] = !have;
		}
		{
			bool have = false;
			double acc = 0.0;
#pragma unroll
			for (int m = 0; m < NX; m++) {
				if (dhz[m] || M::f_pat(m) == PZ) continue;
				const double t = dh[m] * f[m];
				acc = have ? acc + t : t;
				have = true;
			}
			lf[j] = acc;
		}
#pragma unroll
		for (int i = 0; i < NU; i++) {
			bool have = false;
			double acc = 0.0;
#pragma unroll
			for (int m = 0; m < NX; m++) {
				const int pg = M::g_pat(m + i * NX);
				if (dhz[m] || pg == PZ) continue;
				const double t = (pg == P1) ? dh[m] : dh[m] * g[m + i * NX];
				acc = have ? acc + t : t;
				have = true;
			}
			lg[j * NU + i] = acc;
		}
	}
}

} // namespace asifb
