// closed_loop_kernel.cuh -- the part of the example main loops that sits around filter(): sample-and-hold of the
// filtered input, the smoothBounds rate limiter, the Euler plant step and the per-step log record.
// Reference: examples/segway_implicit_tb.cpp:251-331, examples/InvertedPendulum_Implicit.cpp:113-147,
// examples/InvertedPendulum_RealizableSampled.cpp:258-304, examples/DoubleIntegrator_RealizableSampled.cpp:110-205.
// One thread per agent; the state stays on the device between control steps.
#pragma once
#include "filter_common.cuh"
#include "models.cuh"

namespace asifb {

constexpr int LOOP_THREADS = 128;

// plant of the table / kernel pendulum examples: dynamicsExact with the true input gain p
// (examples/InvertedPendulum_RealizableSampled.cpp:56-62)
struct PendulumExactPlant {
	static constexpr int NX = 2, NU = 1;
	__device__ static void dynamics(const double *x, const double gain, double *f, double *g)
	{
		f[0] = x[1];
		f[1] = sin(x[0]);
		g[0] = 0.;
		g[1] = gain;
	}
};

template <class M>
struct ModelPlant {
	static constexpr int NX = M::NX, NU = M::NU;
	__device__ static void dynamics(const double *x, const double, double *f, double *g) { M::dynamics(x, f, g); }
};

struct LoopParams {
	double dt;          // plant step
	double t0;          // time at the first plant step of this launch (accumulated on the host as the examples do)
	int32_t k;          // plant steps in this launch (<= steps_per_sample)
	int32_t step0;      // global index of the first plant step
	int32_t smooth;     // smoothBounds on/off
	int32_t hold_on_failure; // ASIF / ASIFrobust / ASIFrealizable leave uAct untouched when the QP fails
	                         // (src/asif.cpp:199-209, src/asif_robust.cpp:249-251, src/asif_realizable.cpp:324-351):
	                         // the example loops then keep applying the previous filter output
	int32_t log_stride; // 0 = no log
	int32_t log_after;  // record after the plant step (segway / implicit examples) or before it (sampled examples)
	int32_t n_relax, n_diag, tb_diag; // tb_diag: diag holds the TB record (TTS, ortho, ..., critIdx at 4)
	int32_t log_width, log_records;
	int64_t log_agents;
	double smooth_lb, smooth_ub, smooth_rate, plant_gain;
};

__global__ void fill_pairs_kernel(double *p, const int64_t n, const double a, const double b)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		p[2 * i] = a;
		p[2 * i + 1] = b;
	}
}

// log record: t, x[nx], xEstim[nx], uDes[nu], uFilter[nu], uAct[nu], relax[n_relax], rc, smoothLo, smoothHi, TTS, ortho, crit0
template <class P>
__global__ void __launch_bounds__(LOOP_THREADS)
closed_loop_step_kernel(const LoopParams p, const int64_t n, double *__restrict__ x_io, const double *__restrict__ x_estim,
                        const double *__restrict__ u_des, const double *__restrict__ u_filter, double *__restrict__ u_hold,
                        double *__restrict__ u_act,
                        const double *__restrict__ relax, const int32_t *__restrict__ rc, double *__restrict__ smooth,
                        const double *__restrict__ diag, double *__restrict__ log, unsigned long long *__restrict__ rc_hist)
{
	constexpr int NX = P::NX, NU = P::NU;
	const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const bool live = a < n;
	int slot = 7;
	bool failed = false;
	if (live) {
		const int32_t r = rc[a];
		slot = (r >= -3 && r <= 2) ? r + 3 : 7;
		failed = r < 1;
	}
#pragma unroll
	for (int i = 0; i < 8; i++) { // one count per filter call
		unsigned int c = (live && slot == i) ? 1u : 0u;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
		if ((threadIdx.x & 31) == 0 && c) atomicAdd(rc_hist + i, (unsigned long long)c);
	}
	if (!live) return;
	double x[NX], ud[NU], uf[NU], ua[NU];
#pragma unroll
	for (int i = 0; i < NX; i++) x[i] = x_io[a * NX + i];
#pragma unroll
	for (int i = 0; i < NU; i++) {
		ud[i] = u_des[a * NU + i];
		uf[i] = (p.hold_on_failure && failed) ? u_hold[a * NU + i] : u_filter[a * NU + i];
		u_hold[a * NU + i] = uf[i];
		ua[i] = uf[i];
	}
	double sLo = 0.0, sHi = 0.0;
	if (p.smooth) { // examples/DoubleIntegrator_RealizableSampled.cpp:120-150, input 0
		sLo = smooth[a * 2];
		sHi = smooth[a * 2 + 1];
		const double sc = p.smooth_rate;
		if (ua[0] > ud[0] && ua[0] > sLo) {
			sLo = ua[0];
			if (sLo > sHi) sHi = sLo;
			else sHi += sc;
		} else if (ua[0] < ud[0] && ua[0] < sHi) {
			sHi = ua[0];
			if (sLo > sHi) sLo = sHi;
			else sLo -= sc;
		} else {
			sLo -= sc;
			sHi += sc;
		}
		if (sLo < p.smooth_lb) sLo = p.smooth_lb;
		if (sHi > p.smooth_ub) sHi = p.smooth_ub;
		if (ua[0] > sHi) ua[0] = sHi;
		if (ua[0] < sLo) ua[0] = sLo;
		smooth[a * 2] = sLo;
		smooth[a * 2 + 1] = sHi;
	}
#pragma unroll
	for (int i = 0; i < NU; i++) u_act[a * NU + i] = ua[i];

	const bool logged = p.log_stride > 0 && a < p.log_agents;
	double t = p.t0;
	for (int j = 0; j < p.k; j++) {
		const int g_idx = p.step0 + j;
		const bool rec = logged && (g_idx % p.log_stride == 0) && (g_idx / p.log_stride < p.log_records);
		double *L = rec ? log + ((int64_t)a * p.log_records + g_idx / p.log_stride) * p.log_width : nullptr;
		if (rec && !p.log_after) {
			L[0] = t;
#pragma unroll
			for (int i = 0; i < NX; i++) L[1 + i] = x[i];
		}
		double f[NX], g[NX * NU];
		P::dynamics(x, p.plant_gain, f, g);
#pragma unroll
		for (int i = 0; i < NX; i++) { // fCl = f + g uAct ; x += dt fCl
			double fcl = 0.0;
			fcl += f[i];
#pragma unroll
			for (int q = 0; q < NU; q++) fcl += g[i + q * NX] * ua[q];
			x[i] += p.dt * fcl;
		}
		t += p.dt;
		if (rec) {
			if (p.log_after) {
				L[0] = t;
#pragma unroll
				for (int i = 0; i < NX; i++) L[1 + i] = x[i];
			}
			int o = 1 + NX;
#pragma unroll
			for (int i = 0; i < NX; i++) L[o + i] = x_estim[a * NX + i];
			o += NX;
#pragma unroll
			for (int i = 0; i < NU; i++) {
				L[o + i] = ud[i];
				L[o + NU + i] = uf[i];
				L[o + 2 * NU + i] = ua[i];
			}
			o += 3 * NU;
			for (int i = 0; i < p.n_relax; i++) L[o + i] = relax[a * p.n_relax + i];
			o += p.n_relax;
			L[o] = (double)rc[a];
			L[o + 1] = sLo;
			L[o + 2] = sHi;
			const bool td = p.tb_diag && diag;
			L[o + 3] = td ? diag[a * p.n_diag] : 0.0;
			L[o + 4] = td ? diag[a * p.n_diag + 1] : 0.0;
			L[o + 5] = td ? diag[a * p.n_diag + 4] : 0.0;
		}
	}
#pragma unroll
	for (int i = 0; i < NX; i++) x_io[a * NX + i] = x[i];
}

} // namespace asifb
