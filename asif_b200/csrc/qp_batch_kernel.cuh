// qp_batch_kernel.cuh -- the QPWrapper backend: n dense QPs of identical shape, one per thread.
// Replaces QPWrapperOsqp::{initialize,updateCost,updateA,updateb,solve,getSolution}
// (src/qpwrapper_osqp.cpp:55-261).  Problem data arrive in the reference's layout (per problem:
// A[nc*nv] column-major, b[nc], c[nv], ...; problems consecutive), which is problem-major and
// therefore not coalesced for a thread-per-problem mapping.  Each CTA first stages its slab of
// A and b through shared memory with fully coalesced loads ([problem][element] in HBM ->
// [element][problem] in smem, conflict free), then every thread runs the register-resident dual
// active-set solver on its own column of the slab.
#pragma once
#include "qp_gi.cuh"
#include <stdint.h>

namespace asifb {

template <int NV>
struct BatchRows {
	const double *A; // smem view [(nc*NV)][T] offset by thread: element (i + j*nc)
	const double *b; // smem view [nc][T]
	const uint8_t *be;
	int stride, nc;
	double lb[NV], ub[NV];
	bool has_eq;
	// rows: [0,nc) A rows; [nc, nc+2NV) bounds; [nc+2NV, 2nc+2NV) negated equality rows
	__device__ __forceinline__ int count() const { return has_eq ? 2 * nc + 2 * NV : nc + 2 * NV; }
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
			for (int i = 0; i < NV; i++) n[i] = A[(j + i * nc) * stride];
			rhs = b[j * stride];
		} else if (j < nc + 2 * NV) {
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
		} else {
			const int r = j - nc - 2 * NV;
			if (be[r]) { // A v = b  <=>  A v >= b  and  -A v >= -b
#pragma unroll
				for (int i = 0; i < NV; i++) n[i] = -A[(r + i * nc) * stride];
				rhs = -b[r * stride];
			} else {
#pragma unroll
				for (int i = 0; i < NV; i++) n[i] = 0.0;
				rhs = -1.0;
			}
		}
	}
};

constexpr int QPB_THREADS = 64;

template <int NV>
__global__ void __launch_bounds__(QPB_THREADS)
qp_batch_kernel(const int64_t n, const int nc, const int diagonal_cost, const double *__restrict__ H,
                const double *__restrict__ c_in, const double *__restrict__ A_in, const double *__restrict__ b_in,
                const double *__restrict__ lb, const double *__restrict__ ub, const uint8_t *__restrict__ be,
                double *__restrict__ sol, int32_t *__restrict__ status, const int shared_H, const int shared_bounds)
{
	extern __shared__ double smem[];
	const int T = blockDim.x;  // 64, or 32 when the slab of a 64-problem CTA would not fit in shared memory
	const int TS = T + 1;      // padded element stride: conflict-free transposing stores
	const int64_t base = (int64_t)blockIdx.x * T;
	const int nprob = (int)((n - base) < T ? (n - base) : T);
	const int ea = nc * NV;
	// stage A: slab is contiguous in HBM: [nprob][ea]
	{
		const double *src = A_in + base * ea;
		const int total = nprob * ea;
		for (int idx = threadIdx.x; idx < total; idx += T) {
			const int pr = idx / ea, e = idx - pr * ea;
			smem[e * TS + pr] = src[idx];
		}
		const double *srcb = b_in + base * nc;
		double *sb = smem + (size_t)ea * TS;
		const int totalb = nprob * nc;
		for (int idx = threadIdx.x; idx < totalb; idx += T) {
			const int pr = idx / nc, e = idx - pr * nc;
			sb[e * TS + pr] = srcb[idx];
		}
	}
	__syncthreads();
	const int64_t k = base + threadIdx.x;
	if (k >= n) return;
	BatchRows<NV> R;
	R.A = smem + threadIdx.x;
	R.b = smem + (size_t)ea * TS + threadIdx.x;
	R.be = be;
	R.stride = TS;
	R.nc = nc;
	R.has_eq = (be != nullptr);
	const double *lbk = shared_bounds ? lb : lb + k * NV;
	const double *ubk = shared_bounds ? ub : ub + k * NV;
	const double *Hk = shared_H ? H : H + k * NV * NV;
	double c[NV], v[NV];
#pragma unroll
	for (int i = 0; i < NV; i++) {
		v[i] = 0.0; // what getSolution() returns when H is rejected (-7) and the solver never runs
		R.lb[i] = lbk[i];
		R.ub[i] = ubk[i];
		c[i] = c_in[k * NV + i];
	}
	int st;
	if (diagonal_cost) {
		DiagMetric<NV> mt;
		bool ok = true;
#pragma unroll
		for (int i = 0; i < NV; i++) {
			const double h2 = 2.0 * Hk[i + i * NV];
			ok = ok && (h2 > 0.0);
			mt.gi[i] = (h2 > 0.0) ? 1.0 / h2 : 1.0;
			mt.gih[i] = sqrt(mt.gi[i]);
		}
		st = ok ? qp_gi_solve<NV>(mt, c, R, v) : -7; // OSQP_NON_CVX
	} else {
		CholMetric<NV> mt;
		const bool ok = mt.factor(Hk);
		st = ok ? qp_gi_solve<NV>(mt, c, R, v) : -7;
	}
#pragma unroll
	for (int i = 0; i < NV; i++) sol[k * NV + i] = v[i];
	status[k] = st;
}

} // namespace asifb
