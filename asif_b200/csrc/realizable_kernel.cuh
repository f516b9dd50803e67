// realizable_kernel.cuh -- batched ASIFrealizable::filter (nx = 2, nu = 1), one state per thread.
// Reference path replaced: src/asif_realizable.cpp:314-352 (filter), :375-610 (updateConstraints):
// up to nFacets OSQP feasibility solves + 9 libaffa dynamics evaluations + a 38-variable QP per call.
//
// What happens per state here:
//   (i)   hFull_i = 1 - n_i.x over all facets, with a running list of the npSSmax smallest (barrier rows)
//   (ii)  critical facets in index order (at most maxCrit): bounding box inflated by the uncertainty
//         bounds, then the EXACT segment/box intersection that the reference poses as a feasibility QP
//         ({lambda in [0,1]^2, sum = 1, |V lambda - x| <= unc}  <=>  the interval of t in [0,1] for which
//         t v0 + (1-t) v1 lies in the box is non-empty)
//   (iii) rows from the x-independent facet table [LfLo, LfHi, LgLo, LgHi] (gathered by facet id and active
//         constraint; computed once at initialize by the caller's interval dynamics), LP-dual multipliers
//         eliminated exactly as in robust_kernel.cuh:   LgLo u >= -LfLo,  LgHi u >= -LfLo
//   (iv)  barrier rows with the mid-point dynamics:  Lgh u + eps >= -Lfh - relaxDes (h - relaxOffset)
//   (v)   QP in v = (u, eps), eps in [0, inf], cost (u-uDes)^2 + relaxCost eps^2; rc -2 / 1 / -1.
// All tables (<= 5 KB) are staged in shared memory once per CTA; every lane reads the same facet in
// (i)-(ii) (broadcast), the gathers of (iii) are per lane.
#pragma once
#include "filter_common.cuh"
#include "qp_gi.cuh"

namespace asifb {

constexpr int RZ_THREADS = 128;
constexpr int RZ_MAX_CRIT = 8, RZ_MAX_ACT = 4, RZ_MAX_BAR = 4;

struct RealizableParams {
	double lb, ub;
	double relaxDes, relaxOffset, relaxCost, inf;
	double unc[2];
	double gmid; // mid of the input-gain interval as libaffa reports it
	double gi[2], gih[2];
	int32_t n_vertices, n_facets, max_crit, max_act, npSSmax, pad_;
	// device pointers
	const double *vertices;      // [nV][2]
	const double *normals;       // [nF][2]
	const int32_t *facet_vertices; // [nF][2]
	const int32_t *facet_active;   // [nF][max_act], -1 = absent
	const double *facet_lie;     // [nF][max_act][4] = LfLo, LfHi, LgLo, LgHi
};

struct RealizableRows {
	static constexpr int NV = 2;
#ifndef ASIF_RZ_SCAN_INDEX_ONLY
#define ASIF_RZ_SCAN_INDEX_ONLY 1 // C4 0.853 -> 0.828 ms per 1e6 states
#endif
	static constexpr bool SCAN_INDEX_ONLY = ASIF_RZ_SCAN_INDEX_ONLY != 0; // qp_gi.cuh
	// The facet rows are NOT stored per state: a slot is the index of its entry [LfLo, LfHi, LgLo, LgHi] in the shared-memory
	// facet table (16 bits, list [slot][thread] in shared memory), and the solver's scans read the three numbers from the
	// table.  (Round 1 kept 3 x 32 doubles per thread, dynamically indexed: a 1 KB stack frame, 242 registers, 8 warps/SM.)
	const double *lie;    // shared-memory facet table, 4 doubles per (facet, active constraint)
	const uint16_t *slot; // this thread's list, stride RZ_THREADS
	int nslots;
	__device__ __forceinline__ const double *entry(const int s) const { return lie + 4 * (int)slot[s * RZ_THREADS]; }
	double barL[RZ_MAX_BAR], barB[RZ_MAX_BAR];
	int nbar;
	double lb[NV], ub[NV];
	// row numbering: 2*s, 2*s+1 for facet slot s; then barrier rows; then the 4 bounds
	template <class F, class FB>
	__device__ __forceinline__ void scan(F &&fn, FB &&fb) const
	{
		for (int s = 0; s < nslots; s++) {
			const double *t = entry(s);
			double n[NV] = {t[2], 0.0};
			const double r = -t[0];
			fn(2 * s, n, r);
			n[0] = t[3];
			fn(2 * s + 1, n, r);
		}
		for (int i = 0; i < nbar; i++) {
			const double n[NV] = {barL[i], 1.0};
			fn(2 * nslots + i, n, barB[i]);
		}
		const int base = 2 * nslots + nbar;
#pragma unroll
		for (int k = 0; k < 2 * NV; k++) fb(base + k, k >> 1, (k & 1) != 0, (k & 1) ? -ub[k >> 1] : lb[k >> 1]);
	}
	// the two rows of a facet slot differ only in the u coefficient (Lg- <= Lg+): one of them per scan, as in robust_kernel.cuh
	template <class F, class FB>
	__device__ __forceinline__ void scan_at(const double (&v)[NV], F &&fn, FB &&fb) const
	{
		const bool lower = v[0] >= 0.0;
		for (int s = 0; s < nslots; s++) {
			const double *t = entry(s);
			const double n[NV] = {lower ? t[2] : t[3], 0.0};
			fn(2 * s + (lower ? 0 : 1), n, -t[0]);
		}
		for (int i = 0; i < nbar; i++) {
			const double n[NV] = {barL[i], 1.0};
			fn(2 * nslots + i, n, barB[i]);
		}
		const int base = 2 * nslots + nbar;
#pragma unroll
		for (int k = 0; k < 2 * NV; k++) fb(base + k, k >> 1, (k & 1) != 0, (k & 1) ? -ub[k >> 1] : lb[k >> 1]);
	}
	__device__ __forceinline__ void get(const int j, double (&n)[NV], double &r) const
	{
		const int base = 2 * nslots + nbar;
		if (j >= base) {
			const int k = j - base, var = k >> 1;
			const bool upper = k & 1;
			n[0] = (var == 0) ? (upper ? -1.0 : 1.0) : 0.0;
			n[1] = (var == 1) ? (upper ? -1.0 : 1.0) : 0.0;
			r = upper ? -ub[var] : lb[var];
		} else if (j >= 2 * nslots) {
			n[0] = barL[j - 2 * nslots];
			n[1] = 1.0;
			r = barB[j - 2 * nslots];
		} else {
			const double *t = entry(j >> 1);
			n[0] = (j & 1) ? t[3] : t[2];
			n[1] = 0.0;
			r = -t[0];
		}
	}
};

template <class F, class FB>
__device__ __forceinline__ void qp_scan_rows(const RealizableRows &rows, const double (&v)[2], F &&fn, FB &&fb)
{
	rows.scan_at(v, fn, fb);
}

// diag: [nCrit, critFacet[max_crit] (-1 absent), barrierFacet[npSSmax], per slot LgLo, LgHi, LfLo, LfHi, per barrier row Lgh, b]
template <bool WITH_DIAG>
#ifndef RZ_MIN_BLOCKS
// Round 1 / early round 2: rows stored per thread, 242 registers, 8 warps/SM, 0.81 ms per 1e6 C4 states (forcing occupancy
// spilled the facet pass: 0.884 / 1.09 / 1.65 ms at 3 / 4 / 5 CTAs).  With the rows read from the shared-memory table through a
// slot list: 134 registers; 1 / 4 / 5 / 6 CTAs per SM = 0.547 / 0.454 / 0.407 / 0.399 ms (profiles/experiments/r02_ab_c4_slots.txt)
#define RZ_MIN_BLOCKS 6
#endif
__global__ void __launch_bounds__(RZ_THREADS, RZ_MIN_BLOCKS)
realizable_ip_filter_kernel(const RealizableParams p, const int64_t n, const double *__restrict__ x_in,
                            const double *__restrict__ u_des, double *__restrict__ u_act, double *__restrict__ relax_out,
                            int32_t *__restrict__ rc_out, double *__restrict__ diag, unsigned long long *__restrict__ qp_iter_sum)
{
	extern __shared__ double sm[];
	const int nV = p.n_vertices, nF = p.n_facets, mA = p.max_act;
	double *sV = sm;                 // [nV][2]
	double *sN = sV + 2 * nV;        // [nF][2]
	double *sLie = sN + 2 * nF;      // [nF][mA][4]
	int32_t *sFV = (int32_t *)(sLie + 4 * nF * mA); // [nF][2]
	int32_t *sFA = sFV + 2 * nF;                     // [nF][mA]
	uint16_t *sSlot = (uint16_t *)(sFA + nF * mA) + threadIdx.x; // [RZ_MAX_CRIT * RZ_MAX_ACT][RZ_THREADS]: slot lists
	// bounding box of every facet inflated by the uncertainty bounds (:407-415): state independent, so formed once per CTA -
	// the same four operations per facet the per-state test did, hence the same bits
	double *sBox = (double *)(((uintptr_t)((uint16_t *)(sFA + nF * mA) + RZ_MAX_CRIT * RZ_MAX_ACT * RZ_THREADS) + 7) & ~(uintptr_t)7); // [nF][4]
	for (int i = threadIdx.x; i < 2 * nV; i += blockDim.x) sV[i] = p.vertices[i];
	for (int i = threadIdx.x; i < 2 * nF; i += blockDim.x) {
		sN[i] = p.normals[i];
		sFV[i] = p.facet_vertices[i];
	}
	for (int i = threadIdx.x; i < 4 * nF * mA; i += blockDim.x) sLie[i] = p.facet_lie[i];
	for (int i = threadIdx.x; i < nF * mA; i += blockDim.x) sFA[i] = p.facet_active[i];
	for (int i = threadIdx.x; i < nF; i += blockDim.x) {
		const double *v0 = p.vertices + 2 * p.facet_vertices[2 * i], *v1 = p.vertices + 2 * p.facet_vertices[2 * i + 1];
		const double b0lo = fmin(v0[0], v1[0]), b0hi = fmax(v0[0], v1[0]);
		const double b1lo = fmin(v0[1], v1[1]), b1hi = fmax(v0[1], v1[1]);
		sBox[4 * i] = b0lo - p.unc[0];
		sBox[4 * i + 1] = b0hi + p.unc[0];
		sBox[4 * i + 2] = b1lo - p.unc[1];
		sBox[4 * i + 3] = b1hi + p.unc[1];
	}
	__syncthreads();

	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const bool live = k < n;
	const int64_t kk = live ? k : (n - 1);
	const double x0 = x_in[kk * 2], x1 = x_in[kk * 2 + 1];

	// (i) + (ii): one pass over the facets
	int crit[RZ_MAX_CRIT], nCrit = 0;
	double bkey[RZ_MAX_BAR];
	int bidx[RZ_MAX_BAR];
#pragma unroll
	for (int i = 0; i < RZ_MAX_BAR; i++) {
		bkey[i] = INFINITY;
		bidx[i] = -1;
	}
#pragma unroll
	for (int i = 0; i < RZ_MAX_CRIT; i++) crit[i] = -1;
	bool anyNeg = false;
	const int nb = p.npSSmax;
	for (int i = 0; i < nF; i++) {
		// hFull = 1; hFull -= n0 x0; hFull -= n1 x1   (:391-398)
		double h = 1.;
		h -= sN[2 * i] * x0;
		h -= sN[2 * i + 1] * x1;
		anyNeg |= (h < 0.);
		// npSSmax smallest, ascending, ties keep the lower index
		if (nb > 0 && h < bkey[RZ_MAX_BAR - 1]) {
			double ck = h;
			int ci = i;
			bool ins = false;
#pragma unroll
			for (int s = 0; s < RZ_MAX_BAR; s++) {
				const bool sw = (s < nb) && (ins || ck < bkey[s]);
				ins = sw;
				const double tk = bkey[s];
				const int ti = bidx[s];
				bkey[s] = sw ? ck : tk;
				bidx[s] = sw ? ci : ti;
				ck = sw ? tk : ck;
				ci = sw ? ti : ci;
			}
		}
		if (nCrit < p.max_crit) {
			// bounding box inflated by the uncertainty (:407-415), staged per CTA
			const bool potential = !(x0 < sBox[4 * i] || x0 > sBox[4 * i + 1] || x1 < sBox[4 * i + 2] || x1 > sBox[4 * i + 3]);
			if (potential) {
				const double *v0 = sV + 2 * sFV[2 * i], *v1 = sV + 2 * sFV[2 * i + 1];
				// p(t) = t v0 + (1-t) v1 inside [x - unc, x + unc] for some t in [0,1]?   (:419-440, exact)
				double tlo = 0.0, thi = 1.0;
				bool ok = true;
#pragma unroll
				for (int c = 0; c < 2; c++) {
					const double xc = c ? x1 : x0;
					const double d = v0[c] - v1[c], lo = xc - p.unc[c] - v1[c], hi = xc + p.unc[c] - v1[c];
					if (d == 0.0) {
						ok = ok && !(lo > 0.0 || hi < 0.0);
					} else {
						const double a = lo / d, b = hi / d;
						tlo = fmax(tlo, fmin(a, b));
						thi = fmin(thi, fmax(a, b));
					}
				}
				if (ok && tlo <= thi) {
#pragma unroll
					for (int s = 0; s < RZ_MAX_CRIT; s++)
						if (s == nCrit) crit[s] = i;
					nCrit++;
				}
			}
		}
	}
	// barrier keys only hold the first nb entries
	// (iii) rows from the facet table
	RealizableRows R;
	R.lie = sLie;
	R.slot = sSlot;
	R.nslots = 0;
	R.nbar = nb;
#pragma unroll
	for (int c = 0; c < RZ_MAX_CRIT; c++) {
		if (c < nCrit) {
			const int fc = crit[c];
			for (int j = 0; j < mA; j++) {
				if (sFA[mA * fc + j] < 0) continue;
				sSlot[R.nslots * RZ_THREADS] = (uint16_t)(mA * fc + j);
				R.nslots++;
			}
		}
	}
	// (iv) barrier rows, mid-point dynamics f = [x1, sin x0], g = [0, gmid]   (:545-604)
	const double f0 = x1, f1 = sin(x0);
#pragma unroll
	for (int i = 0; i < RZ_MAX_BAR; i++) {
		if (i < nb) {
			const int bi = bidx[i];
			const double Dh0 = -sN[2 * bi], Dh1 = -sN[2 * bi + 1];
			const double Lfh = Dh0 * f0 + Dh1 * f1;
			const double Lgh = Dh0 * 0. + Dh1 * p.gmid;
			R.barL[i] = Lgh;
			R.barB[i] = -Lfh - p.relaxDes * (bkey[i] - p.relaxOffset);
		}
	}
	R.lb[0] = p.lb;
	R.ub[0] = p.ub;
	R.lb[1] = 0.0;
	R.ub[1] = p.inf;
	int32_t rc;
	double uo = 0.0, eps = 0.0;
	int iters = 0;
	if (nCrit == 0 && anyNeg) {
		rc = -2; // outside the kernel with no critical facet (:606-607, :324-326)
	} else {
		double c[2] = {-2.0 * u_des[kk], 0.0}, v[2];
		DiagMetric<2> mt;
		mt.gi[0] = p.gi[0];
		mt.gih[0] = p.gih[0];
		mt.gi[1] = p.gi[1];
		mt.gih[1] = p.gih[1];
		const int st = qp_gi_solve<2>(mt, c, R, v, &iters);
		if (st == QP_OK) {
			uo = input_saturate(v[0], p.lb, p.ub);
			eps = v[1];
			rc = 1;
		} else
			rc = -1;
	}
	if (live) {
		u_act[k] = uo; // the reference leaves uAct untouched on failure; a batch defines it as 0
		relax_out[2 * k] = 0.0; // relax[0] of the reference is a multiplier of the LP-dual form: not defined here
		relax_out[2 * k + 1] = eps;
		rc_out[k] = rc;
		if (WITH_DIAG) {
			const int npSS = p.max_crit * mA;
			double *d = diag + k * (int64_t)(1 + p.max_crit + nb + 4 * npSS + 2 * nb);
			int o = 0;
			d[o++] = (double)nCrit;
			for (int i = 0; i < p.max_crit; i++) {
				int cv = -1;
#pragma unroll
				for (int s = 0; s < RZ_MAX_CRIT; s++) cv = (s == i) ? crit[s] : cv;
				d[o++] = (double)cv;
			}
			for (int i = 0; i < nb; i++) {
				int bv = -1;
#pragma unroll
				for (int s = 0; s < RZ_MAX_BAR; s++) bv = (s == i) ? bidx[s] : bv;
				d[o++] = (double)bv;
			}
			// table entries of the filled slots, in slot order
			int slot = 0;
			for (int c = 0; c < nCrit; c++) {
				int fc = -1;
#pragma unroll
				for (int s = 0; s < RZ_MAX_CRIT; s++) fc = (s == c) ? crit[s] : fc;
				for (int j = 0; j < mA; j++) {
					if (sFA[mA * fc + j] < 0) continue;
					const double *t = sLie + 4 * (mA * fc + j);
					d[o + 4 * slot + 0] = t[2];
					d[o + 4 * slot + 1] = t[3];
					d[o + 4 * slot + 2] = t[0];
					d[o + 4 * slot + 3] = t[1];
					slot++;
				}
			}
			for (; slot < npSS; slot++) d[o + 4 * slot] = d[o + 4 * slot + 1] = d[o + 4 * slot + 2] = d[o + 4 * slot + 3] = 0.0;
			o += 4 * npSS;
			for (int i = 0; i < nb; i++) {
				double bl = 0.0, bb = 0.0;
#pragma unroll
				for (int s = 0; s < RZ_MAX_BAR; s++) {
					bl = (s == i) ? R.barL[s] : bl;
					bb = (s == i) ? R.barB[s] : bb;
				}
				d[o++] = bl;
				d[o++] = bb;
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
