// explicit_kernel.cuh -- batched ASIF::filter (explicit CBF filter), one state per thread:
// safety set + dynamics -> rows  Lgh u + h delta >= -Lfh  -> exact QP -> saturation.
// Reference path replaced: src/asif.cpp:176-210 (filter), :233-312 (updateConstraints) and the
// OSQP solve behind it.  44 B of HBM traffic and ~150 FP64 operations per state: this is the one
// config where HBM can bind, so loads/stores are the coalesced state-major streams and nothing
// else touches global memory.
#pragma once
#include "filter_common.cuh"
#include "qp_gi.cuh"

namespace asifb {

constexpr int EXPL_THREADS = 256;

template <int NV, int NC>
struct RegRows {
	// rows in shared memory [(NC)*(NV+1)][T]; bounds appended
	const double *rows;
	int stride;
	int nc; // rows in use (npSSmax <= NC)
	double lb[NV], ub[NV];
	__device__ __forceinline__ int count() const { return nc + 2 * NV; }
	template <class F, class FB>
	__device__ __forceinline__ void scan(F &&fn, FB &&) const
	{
		const int m = count();
		for (int j = 0; j < m; j++) {
			double n[NV], rhs;
			get(j, n, rhs);
			fn(j, n, rhs);
		}
	}
	__device__ __forceinline__ void get(int j, double (&n)[NV], double &rhs) const
	{
		if (j < nc) {
#pragma unroll
			for (int i = 0; i < NV; i++) n[i] = rows[(j * (NV + 1) + i) * stride];
			rhs = rows[(j * (NV + 1) + NV) * stride];
		} else {
			const int k = j - nc;
			const int var = k >> 1;
			const bool upper = k & 1;
			double bnd = 0.0;
#pragma unroll
			for (int i = 0; i < NV; i++) {
				n[i] = (i == var) ? (upper ? -1.0 : 1.0) : 0.0;
				if (i == var) bnd = upper ? -ub[i] : lb[i];
			}
			rhs = bnd;
		}
	}
};

// one state (index k) of the explicit filter; shared by the batch kernel and the latency server (latency_server.cuh)
// SELECT = false is the common case npSSmax >= npSS (every safety function keeps its own row): the rank computation and the
// per-row "is it among the selected" tests are compiled out; the launcher picks the instantiation from p.npSSmax.
template <class M, bool WITH_DIAG, bool SELECT = true>
__device__ __forceinline__ void explicit_filter_tile(const ExplicitParams &p, const int64_t n, const int64_t k, const double *__restrict__ x_in,
                                                     const double *__restrict__ u_des, double *__restrict__ u_act,
                                                     double *__restrict__ relax_out, int32_t *__restrict__ rc_out,
                                                     double *__restrict__ diag, unsigned long long *__restrict__ qp_iter_sum)
{
	constexpr int NX = M::NX, NU = M::NU, NPSS = M::NPSS, NV = NU + 1, NC = NPSS;
	__shared__ double smem[NC * (NV + 1) * EXPL_THREADS];
	const int T = EXPL_THREADS;
	double *rows = smem + threadIdx.x;
	const bool live = k < n;
	const int64_t kk = live ? k : (n - 1);
	double x[NX], c[NU + 1];
#pragma unroll
	for (int i = 0; i < NX; i++) x[i] = x_in[kk * NX + i];
	if (p.custom_cost) { // filter(x, H, c, ...) (src/asif.cpp:153-174): the caller's c, all nv entries
#pragma unroll
		for (int i = 0; i < NU + 1; i++) c[i] = u_des[kk * (NU + 1) + i];
	} else { // updateCost(uDes) (src/asif.cpp:314-323) + the relax entry of initialize() (:93-98)
#pragma unroll
		for (int i = 0; i < NU; i++) c[i] = -2.0 * u_des[kk * NU + i];
		c[NU] = -2.0 * p.relaxCost * p.relaxLb;
	}

	double h[NPSS], Dh[NPSS * NX], f[NX], g[NX * NU];
	M::safety_set(x, h, Dh);
	M::dynamics(x, f, g);
	// optional row selection: the npSSmax smallest h, ascending, ties keep the lower index (src/asif.cpp:250-268;
	// std::sort leaves tie order unspecified).  pos[j] = rank of safety function j.
	const int ncUse = SELECT ? p.npSSmax : NPSS;
	int pos[NPSS];
#pragma unroll
	for (int j = 0; j < NPSS; j++) {
		int r = 0;
		if (SELECT) {
#pragma unroll
			for (int q = 0; q < NPSS; q++) r += (h[q] < h[j] || (h[q] == h[j] && q < j)) ? 1 : 0;
		}
		pos[j] = (SELECT && ncUse < NPSS) ? r : j;
	}
	// Lfh = Dh f, Lgh = Dh g (src/asif.cpp:276-285), A = [Lgh | h], b = -Lfh (:295-303); safety function j lands in row
	// pos[j] when it is among the selected ones
	double lgv[NPSS][NU], rhsv[NPSS];
	if constexpr (M::HAS_PATTERNS) {
		// products with the literal 0 / +-1 entries of DhSS, f and g are not formed (x * 0 = +-0 and acc + +-0 = acc, 1 * x = x:
		// the same values as the reference's full sums, up to the sign of an exact zero); what remains is in the reference's order
#pragma unroll
		for (int j = 0; j < NPSS; j++) {
			bool have = false;
			double lf = 0.0;
#pragma unroll
			for (int m = 0; m < NX; m++) {
				const int pt = M::dhs_pat(j + m * NPSS);
				if (pt == PZ || M::f_pat(m) == PZ) continue;
				const double t = (pt == P1) ? f[m] : ((pt == PM1) ? -f[m] : Dh[j + m * NPSS] * f[m]);
				lf = have ? lf + t : t;
				have = true;
			}
			rhsv[j] = -lf;
#pragma unroll
			for (int i = 0; i < NU; i++) {
				bool hv = false;
				double lg = 0.0;
#pragma unroll
				for (int m = 0; m < NX; m++) {
					const int pt = M::dhs_pat(j + m * NPSS), pg = M::g_pat(m + i * NX);
					if (pt == PZ || pg == PZ) continue;
					const double d = (pt == P1) ? 1.0 : ((pt == PM1) ? -1.0 : Dh[j + m * NPSS]);
					const double t = (pg == P1) ? d : ((pt == P1) ? g[m + i * NX] : ((pt == PM1) ? -g[m + i * NX] : d * g[m + i * NX]));
					lg = hv ? lg + t : t;
					hv = true;
				}
				lgv[j][i] = lg;
			}
		}
	} else {
#pragma unroll
	for (int j = 0; j < NPSS; j++) {
		double lf = Dh[j] * f[0];
#pragma unroll
		for (int m = 1; m < NX; m++) lf = lf + Dh[j + m * NPSS] * f[m];
		rhsv[j] = -lf;
#pragma unroll
		for (int i = 0; i < NU; i++) {
			double lg = Dh[j] * g[i * NX];
#pragma unroll
			for (int m = 1; m < NX; m++) lg = lg + Dh[j + m * NPSS] * g[m + i * NX];
			lgv[j][i] = lg;
		}
	}
	}
	if (p.lfh != nullptr) { // use_custom_ineq_: row pos[j] takes the caller's values (src/asif.cpp:287-292)
#pragma unroll
		for (int j = 0; j < NPSS; j++) {
			if (pos[j] >= ncUse) continue;
			rhsv[j] = -p.lfh[kk * ncUse + pos[j]];
#pragma unroll
			for (int i = 0; i < NU; i++) lgv[j][i] = p.lgh[kk * ncUse * NU + pos[j] + i * ncUse];
		}
	}
	constexpr bool CLOSED_FORM = (NU == 1); // one input, relax variable fixed: the QP is an interval intersection
	if (!CLOSED_FORM || WITH_DIAG) {
#pragma unroll
		for (int j = 0; j < NPSS; j++) {
			if (pos[j] >= ncUse) continue;
#pragma unroll
			for (int i = 0; i < NU; i++) rows[(pos[j] * (NV + 1) + i) * T] = lgv[j][i];
			rows[(pos[j] * (NV + 1) + NU) * T] = h[j];
			rows[(pos[j] * (NV + 1) + NV) * T] = rhsv[j];
		}
	}
	double v[NV];
	int iters = 0, st;
	if (CLOSED_FORM) {
		// delta is pinned to relaxLb by both bounds (src/asif.cpp:88-91), so row j reads  Lgh_j u >= -Lfh_j - h_j relaxLb:
		// a lower bound on u when Lgh_j > 0, an upper bound when < 0, a feasibility condition when == 0, and
		// u* = clamp(uDes) onto the intersection with [lb, ub] -- the exact optimum of the QP the reference hands to OSQP
		// (SURVEY 8c known-answer anchor ii).  Bounds are kept as fractions r/a and compared by cross-multiplication:
		// one division per state.
		double alo = 1.0, rlo = p.lb[0], ahi = -1.0, rhi = -p.ub[0];
		bool feasible = true;
		bool row_lo = false, row_hi = false;
#pragma unroll
		for (int j = 0; j < NPSS; j++) {
			if (pos[j] >= ncUse) continue;
			const double a = lgv[j][0];
			const double r = rhsv[j] - h[j] * p.relaxLb;
			if (a > 0.0) {
				if (r * alo > rlo * a) {
					alo = a;
					rlo = r;
					row_lo = true;
				}
			} else if (a < 0.0) {
				if (r * ahi < rhi * a) {
					ahi = a;
					rhi = r;
					row_hi = true;
				}
			} else {
				feasible = feasible && !(r > 0.0);
			}
		}
		feasible = feasible && (rlo * ahi >= rhi * alo);
		// unconstrained minimiser of H00 u^2 + c0 u; with the default cost (H00 = 1, c0 = -2 uDes) this is uDes bit for bit
		double u = -(p.gi[0] * c[0]);
		// (one division site for both bounds: the two legs are the same instruction sequence on different operands, and a warp
		// usually holds lanes of both kinds)
		const bool below = u * alo < rlo, above = !below && (u * ahi < rhi);
		if (below || above) {
			u = (below ? rlo : rhi) / (below ? alo : ahi);
			iters = (below ? row_lo : row_hi) ? 1 : 0;
		}
		v[0] = u;
		v[NU] = p.relaxLb;
		st = feasible ? QP_OK : QP_PRIMAL_INFEASIBLE;
	} else {
		RegRows<NV, NC> R;
		R.rows = rows;
		R.stride = T;
		R.nc = ncUse;
		DiagMetric<NV> mt;
#pragma unroll
		for (int i = 0; i < NU; i++) {
			R.lb[i] = p.lb[i];
			R.ub[i] = p.ub[i];
		}
		R.lb[NU] = p.relaxLb; // both bounds pinned (src/asif.cpp:88-91)
		R.ub[NU] = p.relaxLb;
#pragma unroll
		for (int i = 0; i < NV; i++) {
			mt.gi[i] = p.gi[i];
			mt.gih[i] = p.gih[i];
		}
		st = qp_gi_solve<NV>(mt, c, R, v, &iters);
	}
	if (live) {
		if (st == QP_OK) {
#pragma unroll
			for (int i = 0; i < NU; i++) u_act[k * NU + i] = input_saturate(v[i], p.lb[i], p.ub[i]);
			relax_out[k] = v[NU];
			rc_out[k] = 1;
		} else {
			// the reference leaves uAct and relax untouched on failure (src/asif.cpp:207-209);
			// a batch has no "previous value", so the outputs are defined as 0 here
#pragma unroll
			for (int i = 0; i < NU; i++) u_act[k * NU + i] = 0.0;
			relax_out[k] = 0.0;
			rc_out[k] = -1;
		}
		if (WITH_DIAG) {
			const int NDIAG = ncUse * NV + ncUse; // A_[nc*nv] column-major, b_[nc] with nc = npSSmax
			double *d = diag + k * NDIAG;
			for (int j = 0; j < ncUse; j++) {
				for (int i = 0; i < NV; i++) d[j + i * ncUse] = rows[(j * (NV + 1) + i) * T];
				d[ncUse * NV + j] = rows[(j * (NV + 1) + NV) * T];
			}
		}
	}
	if (qp_iter_sum) {
		const unsigned int it = __reduce_add_sync(0xffffffffu, live ? (unsigned int)iters : 0u); // one REDUX instead of a shuffle tree
		if ((threadIdx.x & 31) == 0 && it) qp_rows_add(qp_iter_sum, (unsigned long long)it);
	}
}

template <class M, bool WITH_DIAG, bool SELECT = true>
__global__ void __launch_bounds__(EXPL_THREADS)
explicit_filter_kernel(const ExplicitParams p, const int64_t n, const double *__restrict__ x_in,
                       const double *__restrict__ u_des, double *__restrict__ u_act, double *__restrict__ relax_out,
                       int32_t *__restrict__ rc_out, double *__restrict__ diag,
                       unsigned long long *__restrict__ qp_iter_sum)
{
	// (An L2 prefetch of the input lines of the CTA 888 / 1776 / 3552 blocks ahead was measured and is not in: 1e7 states
	// 113 -> 109 us, 1e8 states 0.956 -> 0.91-1.03 ms, inside the run-to-run spread; gpurun_out/r02_c1_prefetch_ab.log.)
	explicit_filter_tile<M, WITH_DIAG, SELECT>(p, n, (int64_t)blockIdx.x * EXPL_THREADS + threadIdx.x, x_in, u_des, u_act, relax_out, rc_out,
	                                           diag, qp_iter_sum);
}

} // namespace asifb
