// robust_kernel.cuh -- batched ASIFrobust::filter on a half-plane safety table, one state per thread.
// Reference path replaced: src/asif_robust.cpp:218-252 (filter), :275-367 (updateConstraints) with
// the interval Lie derivatives libaffa produces, and the 402-variable LP-dual QP behind it.
//
// The reference carries 2(nu+1) multipliers per safety function (nv = nu+1+2 npSS (nu+1), dense).
// They have zero cost and can be eliminated exactly: row k holds for some lambda >= 0 iff
//     h_k d + min(Lg-_k u, Lg+_k u) + Lf-_k >= 0     (nu = 1)
// i.e. two ordinary rows in (u, d).  The optimum (u*, d*) is identical; the multipliers
// are not outputs.  So per state: K half-planes -> 2K rows, nothing stored: the active-set solver's
// scan recomputes each row from the table entry (shared memory, same address for all lanes ->
// broadcast) and four per-state scalars.  HBM traffic stays 44 B per state; the table is read once
// per CTA.
//
// Interval values for the InvertedPendulum dynamics on a point state (examples/InvertedPendulum_Robust.cpp:62-69):
// f = [x1, sin x0] exact, g = [0, [pMin, pMax]] = centre (pMin+pMax)/2 +- (pMax-pMin)/2, hence
// Lf_k = Dh_k0 f0 + Dh_k1 f1 (a point) and Lg_k = Dh_k1 gc +- |Dh_k1 gr|.
#pragma once
#include "filter_common.cuh"
#include "qp_gi.cuh"

namespace asifb {

constexpr int ROB_THREADS = 128;

struct RobustParams {
	double lb[MAX_NU], ub[MAX_NU];
	double relaxLb, relaxCost, inf;
	double gc, gr; // centre and radius of the input-gain interval
	double gi[MAX_NV], gih[MAX_NV];
	int32_t n_halfplanes;
	int32_t pad_;
	const double *table; // device pointer, [n_halfplanes][2]
};

struct RobustRows {
	static constexpr int NV = 2;
#ifndef ROB_SCAN_INDEX_ONLY
#define ROB_SCAN_INDEX_ONLY 1
#endif
	static constexpr bool SCAN_INDEX_ONLY = ROB_SCAN_INDEX_ONLY != 0; // qp_gi.cuh: 100 half-planes per scan, rows cheap to fetch again
	const double *tab; // shared memory, [K][4] = a0, a1, Lg- / Dh-independent part, Lg+ (the last two staged per CTA)
	int K;
	double x0, x1, f0, f1;
	double lb[NV], ub[NV];
	__device__ __forceinline__ void plane(const int k, double &h, double &lo, double &hi, double &lf) const
	{
		const double a0 = tab[4 * k], a1 = tab[4 * k + 1];
		h = 1. - a0 * x0 - a1 * x1;
		const double Dh0 = -a0, Dh1 = -a1;
		lf = Dh0 * f0 + Dh1 * f1;
		lo = tab[4 * k + 2]; // Lg = Dh1 * [gc - gr, gc + gr] does not depend on the state
		hi = tab[4 * k + 3];
	}
	__device__ __forceinline__ void bound_row(const int k, double (&n)[NV], double &rhs) const
	{
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
	template <class F, class FB>
	__device__ __forceinline__ void scan(F &&fn, FB &&fb) const
	{
		for (int k = 0; k < K; k++) {
			double h, lo, hi, lf;
			plane(k, h, lo, hi, lf);
			double n[NV] = {lo, h};
			fn(2 * k, n, -lf);
			n[0] = hi;
			fn(2 * k + 1, n, -lf);
		}
#pragma unroll
		for (int k = 0; k < 2 * NV; k++) fb(2 * K + k, k >> 1, (k & 1) != 0, (k & 1) ? -ub[k >> 1] : lb[k >> 1]);
	}
	// The two rows of a half-plane differ only in the u coefficient (Lg- <= Lg+): at an iterate with u >= 0 the Lg- row
	// has the smaller residual, with u < 0 the Lg+ row.  The other one can be neither the most violated row nor
	// violated at all when its sibling is active, so the solver's scan evaluates one row per half-plane.
	template <class F, class FB>
	__device__ __forceinline__ void scan_at(const double (&v)[NV], F &&fn, FB &&fb) const
	{
		const bool lower = v[0] >= 0.0;
		for (int k = 0; k < K; k++) {
			double h, lo, hi, lf;
			plane(k, h, lo, hi, lf);
			const double n[NV] = {lower ? lo : hi, h};
			fn(2 * k + (lower ? 0 : 1), n, -lf);
		}
#pragma unroll
		for (int k = 0; k < 2 * NV; k++) fb(2 * K + k, k >> 1, (k & 1) != 0, (k & 1) ? -ub[k >> 1] : lb[k >> 1]);
	}
	// The index-only scan of the solver (qp_gi.cuh) needs only the smallest residual and its row number.  The residual of a
	// half-plane row at the iterate (u, d) is  Lg u + h d + Lfh  with  h = 1 - a.x  and  Lfh = -a.f,  i.e.
	//     (Lg u + d) - a0 (x0 d + f0) - a1 (x1 d + f1):
	// three fused multiply-adds per half-plane on two numbers formed once per scan, instead of forming h and Lfh (seven
	// operations) and then the inner product.  Same number up to its last bits; the solver fetches the winner again with get()
	// and works on the exact row, so only a tie within rounding could pick another row, and then both are equally violated.
#ifndef ROB_SCAN_MIN
#define ROB_SCAN_MIN 1
#endif
	static constexpr bool HAS_SCAN_MIN = ROB_SCAN_MIN != 0;
	__device__ __forceinline__ void scan_min(const double (&v)[NV], double &sr, int &pr) const
	{
		const bool lower = v[0] >= 0.0;
		const double P = fma(x0, v[1], f0), Q = fma(x1, v[1], f1);
		const double *t = tab + (lower ? 2 : 3);
		const int off = lower ? 0 : 1;
		for (int k = 0; k < K; k++) {
			const double s = fma(-tab[4 * k + 1], Q, fma(-tab[4 * k], P, fma(t[4 * k], v[0], v[1])));
			const bool better = s < sr;
			sr = better ? s : sr;
			pr = better ? 2 * k + off : pr;
		}
#pragma unroll
		for (int k = 0; k < 2 * NV; k++) {
			const bool upper = (k & 1) != 0;
			const double vv = v[k >> 1];
			const double s = (upper ? -vv : vv) - (upper ? -ub[k >> 1] : lb[k >> 1]);
			const bool better = s < sr;
			sr = better ? s : sr;
			pr = better ? 2 * K + k : pr;
		}
	}
	__device__ __forceinline__ void get(const int j, double (&n)[NV], double &rhs) const
	{
		if (j >= 2 * K) {
			bound_row(j - 2 * K, n, rhs);
		} else {
			double h, lo, hi, lf;
			plane(j >> 1, h, lo, hi, lf);
			n[0] = (j & 1) ? hi : lo;
			n[1] = h;
			rhs = -lf;
		}
	}
};

template <class F, class FB>
__device__ __forceinline__ void qp_scan_rows(const RobustRows &rows, const double (&v)[2], F &&fn, FB &&fb)
{
	rows.scan_at(v, fn, fb);
}

template <bool WITH_DIAG>
#ifndef ROB_MIN_BLOCKS
#define ROB_MIN_BLOCKS 6 // 80 registers, 24 warps/SM: 0.653 -> 0.615 ms per 1e6 C3b states (1 / 6 / 8 CTAs: 0.653 / 0.615 / 0.616)
#endif
__global__ void __launch_bounds__(ROB_THREADS, ROB_MIN_BLOCKS)
robust_ip_filter_kernel(const RobustParams p, const int64_t n, const double *__restrict__ x_in,
                        const double *__restrict__ u_des, double *__restrict__ u_act, double *__restrict__ relax_out,
                        int32_t *__restrict__ rc_out, double *__restrict__ diag, unsigned long long *__restrict__ qp_iter_sum)
{
	extern __shared__ double tab[];
	const int K = p.n_halfplanes;
	for (int i = threadIdx.x; i < K; i += blockDim.x) {
		const double a0 = p.table[2 * i], a1 = p.table[2 * i + 1];
		const double Dh1 = -a1;
		const double lgc = Dh1 * p.gc, lgr = fabs(Dh1 * p.gr);
		tab[4 * i] = a0;
		tab[4 * i + 1] = a1;
		tab[4 * i + 2] = lgc - lgr;
		tab[4 * i + 3] = lgc + lgr;
	}
	__syncthreads();
	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const bool live = k < n;
	const int64_t kk = live ? k : (n - 1);
	RobustRows R;
	R.tab = tab;
	R.K = K;
	R.x0 = x_in[kk * 2];
	R.x1 = x_in[kk * 2 + 1];
	R.f0 = R.x1;
	R.f1 = sin(R.x0);
	R.lb[0] = p.lb[0];
	R.ub[0] = p.ub[0];
	R.lb[1] = p.relaxLb;
	R.ub[1] = p.inf;
	double c[2] = {-2.0 * u_des[kk], -2.0 * p.relaxCost * p.relaxLb}, v[2];
	DiagMetric<2> mt;
#pragma unroll
	for (int i = 0; i < 2; i++) {
		mt.gi[i] = p.gi[i];
		mt.gih[i] = p.gih[i];
	}
	int iters = 0;
	const int st = qp_gi_solve<2>(mt, c, R, v, &iters);
	if (live) {
		if (st == QP_OK) {
			u_act[k] = input_saturate(v[0], p.lb[0], p.ub[0]);
			relax_out[k] = v[1];
			rc_out[k] = 1;
		} else { // the reference leaves uAct untouched (src/asif_robust.cpp:249-251); a batch defines it as 0
			u_act[k] = 0.0;
			relax_out[k] = 0.0;
			rc_out[k] = -1;
		}
		if (WITH_DIAG) {
			double *d = diag + k * (5 * (int64_t)K);
			for (int j = 0; j < K; j++) {
				double h, lo, hi, lf;
				R.plane(j, h, lo, hi, lf);
				d[5 * j] = h;
				d[5 * j + 1] = lo;
				d[5 * j + 2] = hi;
				d[5 * j + 3] = lf;
				d[5 * j + 4] = lf;
			}
		}
	}
	if (qp_iter_sum) {
		unsigned int it = live ? (unsigned int)iters : 0u;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) it += __shfl_xor_sync(0xffffffffu, it, o);
		if ((threadIdx.x & 31) == 0 && it) qp_rows_add(qp_iter_sum, (unsigned long long)it);
	}
}

} // namespace asifb
