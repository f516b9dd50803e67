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
	// 1/range when range is a power of two (then x/range == x*(1/range) bit for bit and the FP64
	// division, ~10 dependent DFMAs per Euler step, is not needed); 0 otherwise
	double inv_range_exact[MAX_NU];
};

// Structural patterns a model may declare for its callback outputs: an entry that is the literal
// 0.0 or 1.0 for every state.  The kernels skip the multiplications by such entries; for finite
// operands this cannot change an IEEE result (x*0 = +-0, acc + +-0 = acc, 1*x = x) except for the
// sign of an exact zero.  PG = general (no assumption).
enum : int { PZ = 0, P1 = 1, PG = 2 };

// Options of ASIFimplicitTB (include/asif_implicit_tb.h:19-33) + what initialize() derives from them
// (src/asif_implicit_tb.cpp:169-223).
struct TbParams {
	double lb[MAX_NU], ub[MAX_NU];
	double relaxCost, relaxSafeLb, relaxTTS, relaxMinOrtho;
	double backTrajHorizon, backTrajDt, backTrajMinOrtho, inf;
	int32_t npBT; // trajectory points (npBT-1 Euler steps)
	int32_t pad_;
	SoftSat sat;
	// QP metric: gi = 1/(2 H_ii), gih = sqrt(gi) for v = (u, relax)
	double gi[MAX_NV], gih[MAX_NV];
};

// Options of ASIF (include/asif.h:11-17) as used by initialize() (src/asif.cpp:64-110)
struct ExplicitParams {
	double lb[MAX_NU], ub[MAX_NU];
	double relaxLb, relaxCost;
	double gi[MAX_NV], gih[MAX_NV];
};

// src/asif_implicit_tb.cpp:821-830
__device__ __forceinline__ double input_saturate(double u, double lb, double ub)
{
	if (u > ub) return ub;
	if (u < lb) return lb;
	return u;
}

// src/asif_implicit_tb.cpp:764-819, one input.  Same branch order and operation order.
__device__ __forceinline__ void input_saturate_soft(const SoftSat &s, int i, double lb, double ub, double u, double &uSat,
                                                    double &DuSat)
{
	const double range = s.range[i], middle = s.middle[i];
	const double inv = s.inv_range_exact[i];
	const double uc = (inv != 0.0) ? (2 * (u - middle)) * inv : 2 * (u - middle) / range;
	if (uc >= s.bevelStop) {
		uSat = ub;
		DuSat = 0;
	} else if (uc <= -s.bevelStop) {
		uSat = lb;
		DuSat = 0;
	} else if (uc <= s.bevelStart && uc >= -s.bevelStart) {
		uSat = u;
		DuSat = 1;
	} else if (uc > s.bevelStart) {
		const double d = uc - s.bevelStop;
		const double sq = sqrt(s.r2 - d * d);
		uSat = sq + s.bevelYc;
		DuSat = (s.bevelStop - uc) / sq;
		uSat = 0.5 * uSat * range + middle;
	} else if (uc < -s.bevelStart) {
		const double d = uc + s.bevelStop;
		const double sq = sqrt(s.r2 - d * d);
		uSat = -sq - s.bevelYc;
		DuSat = (s.bevelStop + uc) / sq;
		uSat = 0.5 * uSat * range + middle;
	} else { // NaN input
		DuSat = 1;
		uSat = u;
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
template <class M>
__device__ __forceinline__ void backup_cl_dynamics(const SoftSat &sat, const double *lb, const double *ub, const double *x,
                                                   double *fCL, double *DfCL)
{
	constexpr int NX = M::NX, NU = M::NU;
	double f[NX], g[NX * NU], u[NU], Du[NU * NX], uSat[NU], DuSat[NU];
	double d[NX * NX], Dg[M::FUSED_GRADIENT ? 1 : NX * NU * NX];
	M::backup_controller(x, u, Du);
#pragma unroll
	for (int k = 0; k < NU; k++) input_saturate_soft(sat, k, lb[k], ub[k], u[k], uSat[k], DuSat[k]);
	if (M::FUSED_GRADIENT) {
		M::dynamics_with_gradient(x, uSat, f, g, d);
	} else {
		M::dynamics(x, f, g);
		M::dynamics_gradients(x, d, Dg);
	}
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
		fCL[i] = have ? acc + f[i] : f[i];
	}
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

} // namespace asifb
