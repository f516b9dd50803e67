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
};

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
	const double uc = 2 * (u - middle) / range;
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

// Closed-loop backup dynamics and their Jacobian, src/asif_implicit_tb.cpp:833-897.
template <class M>
__device__ __forceinline__ void backup_cl_dynamics(const SoftSat &sat, const double *lb, const double *ub, const double *x,
                                                   double *fCL, double *DfCL)
{
	constexpr int NX = M::NX, NU = M::NU;
	double f[NX], g[NX * NU], u[NU], Du[NU * NX], uSat[NU], DuSat[NU];
	M::backup_controller(x, u, Du);
#pragma unroll
	for (int k = 0; k < NU; k++) input_saturate_soft(sat, k, lb[k], ub[k], u[k], uSat[k], DuSat[k]);
	if (M::FUSED_GRADIENT) {
		double d[NX * NX];
		M::dynamics_with_gradient(x, uSat, f, g, d);
#pragma unroll
		for (int i = 0; i < NX; i++)
#pragma unroll
			for (int j = 0; j < NX; j++) {
				double acc = d[i + j * NX];
#pragma unroll
				for (int k = 0; k < NU; k++) acc += g[i + k * NX] * DuSat[k] * Du[k + j * NU];
				DfCL[i + j * NX] = acc;
			}
	} else {
		double Df[NX * NX], Dg[NX * NU * NX];
		M::dynamics(x, f, g);
		M::dynamics_gradients(x, Df, Dg);
#pragma unroll
		for (int i = 0; i < NX; i++)
#pragma unroll
			for (int j = 0; j < NX; j++) {
				double acc = Df[i + j * NX];
#pragma unroll
				for (int k = 0; k < NU; k++)
					acc += Dg[i + k * NX + j * NX * NU] * uSat[k] + g[i + k * NX] * DuSat[k] * Du[k + j * NU];
				DfCL[i + j * NX] = acc;
			}
	}
	// fCL = g uSat + f   (matrixVectorMultiply then += f, :892-896)
#pragma unroll
	for (int i = 0; i < NX; i++) {
		double acc = g[i] * uSat[0];
#pragma unroll
		for (int k = 1; k < NU; k++) acc = acc + g[i + k * NX] * uSat[k];
		fCL[i] = acc + f[i];
	}
}

} // namespace asifb
