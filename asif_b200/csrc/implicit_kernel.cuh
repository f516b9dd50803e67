// implicit_kernel.cuh -- batched ASIFimplicit::filter, one state per thread.
// Reference path replaced: src/asif_implicit.cpp:305-356 (filter), :403-611 (updateConstraints),
// closed-loop rhs :751-827 (identical to the TB class), and the OSQP solve behind it.
// Same design as tb_kernel.cuh: streaming selection of the NPBTSS smallest min-h points over the
// WHOLE horizon (no hit test, :487), snapshots in shared memory, rows recomputed on demand, plus
// NPBS backup-set rows taken at the end of the trajectory (:542-554) with their own relax variable.
// QP in v = (u, delta_safe, delta_reach).
//
// RB = true selects ASIFimplicitRB (src/asif_implicit_robust.cpp:363-441 filter, :481-778 updateConstraints,
// :878-965 closed-loop rhs): the same filter with (a) the backup input under a zero-order hold of backContDt
// (backup_cl_dynamics_zoh) and (b) the h entries of the safety rows replaced by the lower bound of the interval
// safety set over x_i +- x_unc (:636-647; the Dh rows stay nominal, and the interval Lfh / Lgh the reference also
// computes are never used, :689-738).  Selection of the critical points still uses the nominal min h (:572).
#pragma once
#include "filter_common.cuh"
#include "models.cuh" // sincos_model (imp_advance)
#include "qp_gi.cuh"

namespace asifb {

constexpr int IMP_THREADS = 128;

// sin / cos of the model's angle are advanced by the angle-addition recurrence between full evaluations every 16 steps
// (models that declare HAS_TRIG_STATE: the pendulum).  The increment d = x_i - x_{i-1} of the ROUNDED state is exact (Sterbenz),
// sin d and cos d are three- and four-term Taylor sums (|d| < 0.0125: truncation below 1e-17; larger steps take the full
// evaluation), so the only error is the rounding of the rotation itself, ~3 ulp per step, reset every 16 steps.  Both passes of
// the checkpoint kernel re-synchronise at the same indices ((i - 1) % 16 == 0), hence produce the same bits.  The first half of
// the round had rejected this on paper ("multiplies the rate of critical-index flips"); measured: C3a 21.85 -> 19.11 ms per 1e6
// states, and over 1e5 states against the oracle 0 return-code flips, 0 critical-index flips, rows within 1.8e-14 as before,
// max |du| 2.4e-9 (6.8e-10 before) of a 1e-6 bar.  -DASIF_IMP_TRIG_RECURRENCE=0 restores one sincos per step.
#ifndef ASIF_IMP_TRIG_RECURRENCE
#define ASIF_IMP_TRIG_RECURRENCE 1
#endif
#ifndef ASIF_IMP_PEEL_UNROLL
#define ASIF_IMP_PEEL_UNROLL 3
#endif
constexpr int IMP_PEEL_UNROLL = ASIF_IMP_PEEL_UNROLL; // (#pragma unroll does not expand macros)
#ifndef ASIF_IMP_PEEL
#define ASIF_IMP_PEEL 1 // pass A of the checkpoint kernel: re-synchronisation step peeled out of the step loop (C3a 19.04 -> 18.44 ms, same bits)
#endif
template <class M, class = void>
struct model_has_trig {
	static constexpr bool value = false;
};
template <class M>
struct model_has_trig<M, decltype((void)M::HAS_TRIG_STATE)> {
	static constexpr bool value = M::HAS_TRIG_STATE;
};
template <class M>
__host__ __device__ constexpr bool imp_use_trig()
{
	return ASIF_IMP_TRIG_RECURRENCE != 0 && model_has_trig<M>::value;
}

// one step of that recurrence: the angle now is a, tr holds sin / cos of the angle of the step before
// KIND: 0 = decided from i (every caller but the peeled pass-A loop), 1 = i is known to be a re-synchronisation step,
// 2 = i is known not to be one (the |d| guard stays); the same arithmetic for a given i in all three
template <class M, int KIND = 0>
__device__ __forceinline__ void imp_trig_step(const int i, const double a, TrigSC &tr)
{
	const double dl = a - tr.x0;
	if (KIND == 1 || (KIND == 0 && (((i - 1) & 15) == 0)) || !(fabs(dl) < 0.0125)) {
		sincos_model(a, &tr.s, &tr.c);
	} else {
		const double d2 = dl * dl;
		const double sd = dl * fma(d2, fma(d2, 1.0 / 120.0, -1.0 / 6.0), 1.0);
		const double cd = fma(d2, fma(d2, fma(d2, -1.0 / 720.0, 1.0 / 24.0), -0.5), 1.0);
		const double ns = fma(tr.c, sd, tr.s * cd), nc = fma(-tr.s, sd, tr.c * cd);
		tr.s = ns;
		tr.c = nc;
	}
	tr.x0 = a;
}


// The two residual networks of include/asif_learning_utils.h:8-32 on the device: weights column-major [rows x cols]
// as matrixVectorMultiply reads them, in one blob: drift net (w1, b1, w2, b2, w3, b3) then actuation net.
constexpr int LEARN_MAX_WIDTH = 64; // widest input / hidden layer accepted
struct LearnNets {
	const double *blob; // nullptr: use_learning == false
	int32_t d_in[2], d_h1[2], d_h2[2], d_out[2]; // [0] drift net, [1] actuation net
	int32_t off[2];                              // offset of each net inside the blob
};

// driftNN / actNN (include/asif_learning_utils.h:34-119): out = W3 relu(W2 relu(W1 in + b1) + b2) + b3, every product
// accumulated from 0.0 with k ascending (include/asif_utils.h:46-62).  Only the first n_out outputs are formed (the
// reference forms all and uses the first / the first nu).  Kept out of line: it runs once per state.
static __device__ __noinline__ void learn_mlp(const double *__restrict__ w, const int din, const int dh1, const int dh2, const int dout,
                                       const double *in, const int n_out, double *out)
{
	double o1[LEARN_MAX_WIDTH], o2[LEARN_MAX_WIDTH];
	const double *w1 = w, *b1 = w1 + dh1 * din, *w2 = b1 + dh1, *b2 = w2 + dh2 * dh1, *w3 = b2 + dh2, *b3 = w3 + dout * dh2;
	for (int i = 0; i < dh1; i++) {
		double acc = 0.0;
		for (int k = 0; k < din; k++) acc = acc + __ldg(w1 + i + k * dh1) * in[k];
		o1[i] = fmax(0., acc + __ldg(b1 + i));
	}
	for (int i = 0; i < dh2; i++) {
		double acc = 0.0;
		for (int k = 0; k < dh1; k++) acc = acc + __ldg(w2 + i + k * dh2) * o1[k];
		o2[i] = fmax(0., acc + __ldg(b2 + i));
	}
	for (int i = 0; i < n_out; i++) {
		double acc = 0.0;
		for (int k = 0; k < dh2; k++) acc = acc + __ldg(w3 + i + k * dout) * o2[k];
		out[i] = acc + __ldg(b3 + i);
	}
}

// update_weights (include/asif_learning_utils.h:121-155): input [x ; Dh_index_[0..nx-1] ; 0 ...] for both nets
template <int NX, int NU>
__device__ __forceinline__ void learned_residual(const LearnNets &L, const double *x, const double *dh_index, double &dLf,
                                                 double (&dLg)[NU])
{
	double in[LEARN_MAX_WIDTH];
#pragma unroll 1
	for (int i = 0; i < LEARN_MAX_WIDTH; i++) in[i] = 0.0;
#pragma unroll
	for (int i = 0; i < NX; i++) {
		in[i] = x[i];
		in[i + NX] = dh_index[i];
	}
	double od[1];
	learn_mlp(L.blob + L.off[0], L.d_in[0], L.d_h1[0], L.d_h2[0], L.d_out[0], in, 1, od);
	dLf = od[0];
	learn_mlp(L.blob + L.off[1], L.d_in[1], L.d_h1[1], L.d_h2[1], L.d_out[1], in, NU, dLg);
}

// Options of ASIFimplicit (include/asif_implicit.h:20-34) + what initialize() derives (src/asif_implicit.cpp:211-254)
struct ImplicitParams {
	double lb[MAX_NU], ub[MAX_NU];
	double relaxCost, relaxSafeLb, relaxReachLb;
	double backTrajDt, inf;
	int32_t npBT;
	int32_t sat_mode;
	int32_t npBTSS; // critical trajectory points (read by the run-time-count instantiation only)
	int32_t custom_cost; // filter(x, H, c, ...) (src/asif_implicit.cpp:296-303): u_des holds c[n][nv]
	int32_t ck_log;      // checkpoint spacing of implicit_ckpt_kernel (imp_ck_log(npBT)), set by the launcher
	SoftSat sat;
	double gi[MAX_NV], gih[MAX_NV];
	// ASIFimplicitRB only (include/asif_implicit_robust.h:22-38)
	double backContDt;
	double x_unc[4];
	// learned residual (Options.use_learning, include/asif_learning_utils.h): device blob or nullptr
	LearnNets learn;
};

constexpr int IMP_NPBTSS_RUNTIME = -16; // generic instantiation: any npBTSS in 1..16 (see np_capacity in tb_kernel.cuh)

// Shared memory per thread: CAP snapshot slots of (x_i, Q_i), the CAP min-h keys of those slots and (diagnostic kernel
// only) their trajectory indices.  The ORDER of the slots lives in one 64-bit register (ImpOrder below).
template <class M, int NPBTSS, bool WITH_IDX = false>
__host__ __device__ constexpr int imp_smem_doubles_per_thread()
{
	return np_capacity(NPBTSS) * (M::NX + M::NX * M::NX + 1) + (WITH_IDX ? (np_capacity(NPBTSS) + 1) / 2 : 0);
}

// The running list of the CAP smallest min-h points, ascending.  Keys and snapshots sit in shared memory at fixed
// physical slots; the sorted order is a permutation packed four bits per entry: nibble s = slot of the s-th smallest.
// A new point always takes the slot of the entry it evicts (the largest), so an insert is one snapshot store, one key
// store and a few 64-bit shifts - no data moves.  The common cases are O(1): a new overall minimum (every step of a
// trajectory whose min h is still falling, e.g. a pendulum that the saturated backup input cannot catch) shifts the
// whole word by one nibble; anything else searches its place among the stored keys (rare: at most CAP - 1 appends
// while the list fills, then only when the trajectory comes back below an earlier excursion).
// Ties keep the earlier trajectory index first (a new key goes behind equal ones), as before.
template <int CAP>
struct ImpOrder {
	static_assert(CAP >= 1 && CAP <= 16, "slot numbers are packed in four bits, sixteen to a word");
	static constexpr unsigned long long MASK = (CAP == 16) ? ~0ull : ((1ull << (4 * (CAP % 16))) - 1ull);
	unsigned long long perm;
	__device__ __forceinline__ void init()
	{
		unsigned long long v = 0;
#pragma unroll
		for (int s = 0; s < CAP; s++) v |= (unsigned long long)s << (4 * s);
		perm = v;
	}
	__device__ __forceinline__ int slot(const int s) const { return (int)(perm >> (4 * s)) & 15; }
	__device__ __forceinline__ int last() const { return (int)(perm >> (4 * (CAP - 1))) & 15; }
	__device__ __forceinline__ void push_front(const int sl) { perm = ((perm << 4) | (unsigned long long)sl) & MASK; }
	__device__ __forceinline__ void insert_at(const int p, const int sl)
	{
		const unsigned long long low = (1ull << (4 * p)) - 1ull;
		perm = ((perm & low) | ((unsigned long long)sl << (4 * p)) | ((perm & ~low) << 4)) & MASK;
	}
};

template <class M, int NPBTSS, bool RB = false>
struct ImpRows {
#ifndef ASIF_IMP_SCAN_INDEX_ONLY
#define ASIF_IMP_SCAN_INDEX_ONLY 1
#endif
	static constexpr bool SCAN_INDEX_ONLY = ASIF_IMP_SCAN_INDEX_ONLY != 0; // qp_gi.cuh
	static constexpr int NX = M::NX, NU = M::NU, NPSS = M::NPSS, NPBS = M::NPBS, NS = NX + NX * NX;
	static constexpr int CAP = np_capacity(NPBTSS), NV = NU + 2;
	int np; // critical points in use (== NPBTSS when that is a compile-time count)
	__device__ __forceinline__ int count_np() const { return np_runtime(NPBTSS) ? np : NPBTSS; }
	__device__ __forceinline__ int nsafe() const { return count_np() * NPSS; }
	__device__ __forceinline__ int nc() const { return count_np() * NPSS + NPBS; }
	const double *snap;
	int T;
	double f[NX], g[NX * NU];
	ImpOrder<CAP> order; // slot of the s-th smallest point
	double lgB[NPBS][NU], hB[NPBS], rhsB[NPBS]; // backup rows
	double lb[NV], ub[NV];
	double x_unc[RB ? NX : 1];
	bool learn;           // Options.use_learning
	double dLf, dLg[NU];  // residual on Lfh[0] / Lgh[0..nu-1] (flat, column-major npTC x nu: rows 0..nu-1 of column 0)

	__device__ __forceinline__ void point_rows(const int slot, double (&n)[NPSS][NV], double (&rhs)[NPSS], const bool first = false) const
	{
		double xs[NS], hs[NPSS], lf[NPSS], lg[NPSS * NU];
#pragma unroll
		for (int e = 0; e < NS; e++) xs[e] = snap[(slot * NS + e) * T];
		safety_point_rows<M>(xs, f, g, hs, lf, lg);
		if (RB) M::safety_set_lower(xs, x_unc, hs);
		if (learn && first) { // Lfh[0] += ..., Lgh[i] += ... before A_ and b_ are filled (src/asif_implicit.cpp:585-611)
			lf[0] += dLf;
#pragma unroll
			for (int i = 0; i < NU; i++) lg[i * NU] += dLg[i];
		}
#pragma unroll
		for (int j = 0; j < NPSS; j++) {
#pragma unroll
			for (int i = 0; i < NU; i++) n[j][i] = lg[j * NU + i];
			n[j][NU] = hs[j];
			n[j][NU + 1] = 0.0;
			rhs[j] = -lf[j];
		}
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
	__device__ __forceinline__ void backup_row(const int r, double (&n)[NV], double &rhs) const
	{
#pragma unroll
		for (int t = 0; t < NPBS; t++) {
			if (t == r) {
#pragma unroll
				for (int i = 0; i < NU; i++) n[i] = lgB[t][i];
				n[NU] = 0.0;
				n[NU + 1] = hB[t];
				rhs = rhsB[t];
			}
		}
	}
	template <class F, class FB>
	__device__ __forceinline__ void scan(F &&fn, FB &&fb) const
	{
		const int NSAFE = nsafe(), NC = nc();
#pragma unroll 1
		for (int s = 0; s < count_np(); s++) {
			const int slot = order.slot(s);
			double n[NPSS][NV], rhs[NPSS];
			point_rows(slot, n, rhs, s == 0);
#pragma unroll
			for (int j = 0; j < NPSS; j++) fn(s * NPSS + j, n[j], rhs[j]);
		}
#pragma unroll
		for (int r = 0; r < NPBS; r++) {
			double n[NV], rhs = 0.0;
			backup_row(r, n, rhs);
			fn(NSAFE + r, n, rhs);
		}
#pragma unroll
		for (int k = 0; k < 2 * NV; k++) fb(NC + k, k >> 1, (k & 1) != 0, (k & 1) ? -ub[k >> 1] : lb[k >> 1]);
	}
	__device__ __forceinline__ void get(const int j, double (&n)[NV], double &rhs) const
	{
		const int NSAFE = nsafe(), NC = nc();
		if (j >= NC) {
			bound_row(j - NC, n, rhs);
		} else if (j >= NSAFE) {
			backup_row(j - NSAFE, n, rhs);
		} else {
			const int s = j / NPSS, jj = j - s * NPSS;
			const int slot = order.slot(s);
			double nn[NPSS][NV], rr[NPSS];
			point_rows(slot, nn, rr, s == 0);
#pragma unroll
			for (int t = 0; t < NPSS; t++) {
				if (t == jj) {
#pragma unroll
					for (int i = 0; i < NV; i++) n[i] = nn[t][i];
					rhs = rr[t];
				}
			}
		}
	}
};

template <class M, int NPBTSS, bool WITH_DIAG, int SATMODE, bool RB = false>
__global__ void __launch_bounds__(IMP_THREADS, 3)
implicit_filter_kernel(const ImplicitParams p, const int64_t n, const double *__restrict__ x_in,
                       const double *__restrict__ u_des, double *__restrict__ u_act, double *__restrict__ relax_out,
                       int32_t *__restrict__ rc_out, double *__restrict__ diag, unsigned long long *__restrict__ qp_iter_sum)
{
	constexpr int NX = M::NX, NU = M::NU, NPSS = M::NPSS, NPBS = M::NPBS;
	constexpr int NS = NX + NX * NX;
	constexpr int NV = NU + 2, CAP = np_capacity(NPBTSS);
	const int np = np_runtime(NPBTSS) ? p.npBTSS : NPBTSS;
	const int NC = np * NPSS + NPBS;
	const int NDIAG = 2 + np + NC * NV + NC;

	extern __shared__ double smem[];
	const int T = blockDim.x;
	double *snap = smem + threadIdx.x;
	const int64_t k = (int64_t)blockIdx.x * T + threadIdx.x;
	const bool live = k < n;
	const int64_t kk = live ? k : (n - 1);
	double x0[NX], c[NU + 2];
#pragma unroll
	for (int i = 0; i < NX; i++) x0[i] = x_in[kk * NX + i];
	if (p.custom_cost) { // the caller's c, all nv entries
#pragma unroll
		for (int i = 0; i < NU + 2; i++) c[i] = u_des[kk * (NU + 2) + i];
	} else { // updateCost(uDes) + the relax entries of initialize() (:238-254, :653-664)
#pragma unroll
		for (int i = 0; i < NU; i++) c[i] = -2.0 * u_des[kk * NU + i];
		c[NU] = -2.0 * p.relaxCost * p.relaxSafeLb;
		c[NU + 1] = -2.0 * p.relaxCost * p.relaxReachLb;
	}

	double hs[NPSS], Dhs[NPSS * NX];
	M::safety_set(x0, hs, Dhs);
	double hSafetyNow = hs[0];
#pragma unroll
	for (int j = 1; j < NPSS; j++) hSafetyNow = (hs[j] < hSafetyNow) ? hs[j] : hSafetyNow;

	ImpRows<M, NPBTSS, RB> R;
	if (RB) {
#pragma unroll
		for (int i = 0; i < NX; i++) R.x_unc[i] = p.x_unc[i];
	}
	R.np = np;
	R.snap = snap;
	R.T = T;
	double X[NS];
#pragma unroll
	for (int i = 0; i < NS; i++) X[i] = 0.0;
#pragma unroll
	for (int i = 0; i < NX; i++) X[i] = x0[i];
#pragma unroll
	for (int i = 0; i < NX; i++) X[NX + i * (NX + 1)] = 1.0;
	// keys by physical slot (shared memory), trajectory indices likewise (diagnostic kernel only)
	double *keys = snap + CAP * NS * T;
	int *kidx_s = reinterpret_cast<int *>(smem + CAP * (NS + 1) * T) + threadIdx.x;
#pragma unroll
	for (int s = 1; s < CAP; s++) keys[s * T] = INFINITY;
	keys[0] = hSafetyNow;
	if (WITH_DIAG) {
#pragma unroll
		for (int s = 1; s < CAP; s++) kidx_s[s * T] = -1;
		kidx_s[0] = 0;
	}
	R.order.init();
	double kmin = hSafetyNow, kmax = (CAP > 1) ? INFINITY : hSafetyNow; // smallest / largest key in the list
#pragma unroll
	for (int e = 0; e < NS; e++) snap[e * T] = X[e]; // slot 0

	const int N = p.npBT;
	ZohState<M> zoh;
	if (RB) {
#pragma unroll
		for (int i = 0; i < NU; i++) zoh.u[i] = 0.0;
#pragma unroll
		for (int i = 0; i < NU * NX; i++) zoh.Du[i] = 0.0;
		zoh.t_last = -1.;
	}
	TrigSC trig = {0.0, 1.0, 0.0};
	for (int i = 1; i < N; i++) {
		double Xd[NS], DfCL[NX * NX];
		if constexpr (imp_use_trig<M>()) {
			imp_trig_step<M>(i, X[M::TRIG_ANGLE], trig);
			if (RB)
				backup_cl_dynamics_zoh<M, SATMODE, TrigSC>(p.sat, p.lb, p.ub, X, (double)(unsigned)i * p.backTrajDt, p.backTrajDt,
				                                           p.backContDt, zoh, Xd, DfCL, trig);
			else
				backup_cl_dynamics<M, SATMODE, TrigSC>(p.sat, p.lb, p.ub, X, Xd, DfCL, trig);
		} else
		if (RB) // the reference hands t = i*backTrajDt to the rhs (src/asif_implicit_robust.cpp:550)
			backup_cl_dynamics_zoh<M, SATMODE>(p.sat, p.lb, p.ub, X, (double)(unsigned)i * p.backTrajDt, p.backTrajDt,
			                                   p.backContDt, zoh, Xd, DfCL);
		else
			backup_cl_dynamics<M, SATMODE>(p.sat, p.lb, p.ub, X, Xd, DfCL);
		sensitivity_rhs<M>(DfCL, X + NX, Xd + NX);
#pragma unroll
		for (int e = 0; e < NS; e++) X[e] = Xd[e] * p.backTrajDt + X[e];
		double hmin;
		if (M::HAS_SAFETY_MIN) {
			hmin = M::safety_min(X);
		} else {
			M::safety_set(X, hs, Dhs);
			hmin = hs[0];
#pragma unroll
			for (int j = 1; j < NPSS; j++) hmin = (hs[j] < hmin) ? hs[j] : hmin;
		}
		if (hmin < kmax) {
			const int slot = R.order.last(); // the evicted (largest) entry's slot takes the new point
#pragma unroll
			for (int e = 0; e < NS; e++) snap[(slot * NS + e) * T] = X[e];
			if (hmin < kmin) {
				R.order.push_front(slot);
				kmin = hmin;
			} else {
				// first place whose key is larger (hmin >= key of place 0 and hmin < key of place CAP-1 are known)
				int pl = 1;
				while (pl < CAP - 1 && !(hmin < keys[R.order.slot(pl) * T])) pl++;
				R.order.insert_at(pl, slot);
			}
			keys[slot * T] = hmin;
			if (WITH_DIAG) kidx_s[slot * T] = i;
			kmax = keys[R.order.last() * T];
		}
	}
	// open-loop dynamics at the current state (:414-416) and the backup rows at the trajectory end (:542-554)
	M::dynamics(x0, R.f, R.g);
	R.learn = p.learn.blob != nullptr;
	R.dLf = 0.0;
#pragma unroll
	for (int i = 0; i < NU; i++) R.dLg[i] = 0.0;
	if (R.learn) {
		// Dh_index_ = DhSS(x_c) Q_c of the first critical point, npSS x nx column-major; its first nx entries
		// (src/asif_implicit.cpp:533-537, include/asif_learning_utils.h:127-129)
		double xs[NS], hq[NPSS], Dq[NPSS * NX], dhi[NX];
#pragma unroll
		for (int e = 0; e < NS; e++) xs[e] = snap[(R.order.slot(0) * NS + e) * T];
		M::safety_set(xs, hq, Dq);
#pragma unroll
		for (int i = 0; i < NX; i++) {
			const int row = i % NPSS, col = i / NPSS;
			double acc = 0.0;
#pragma unroll
			for (int m = 0; m < NX; m++) acc = acc + Dq[row + m * NPSS] * xs[NX + m + col * NX];
			dhi[i] = acc;
		}
		learned_residual<NX, NU>(p.learn, x0, dhi, R.dLf, R.dLg);
	}
	double hBackupEnd;
	{
		double hB[NPBS], DhB[NPBS * NX];
		M::backup_set_rows(X, hB, DhB);
		hBackupEnd = hB[0];
#pragma unroll
		for (int r = 1; r < NPBS; r++) hBackupEnd = (hB[r] < hBackupEnd) ? hB[r] : hBackupEnd;
#pragma unroll
		for (int r = 0; r < NPBS; r++) {
			double dh[NX];
#pragma unroll
			for (int cc = 0; cc < NX; cc++) {
				double acc = DhB[r] * X[NX + cc * NX];
#pragma unroll
				for (int m = 1; m < NX; m++) acc = acc + DhB[r + m * NPBS] * X[NX + m + cc * NX];
				dh[cc] = acc;
			}
			double lf = dh[0] * R.f[0];
#pragma unroll
			for (int m = 1; m < NX; m++) lf = lf + dh[m] * R.f[m];
#pragma unroll
			for (int i = 0; i < NU; i++) {
				double lg = dh[0] * R.g[i * NX];
#pragma unroll
				for (int m = 1; m < NX; m++) lg = lg + dh[m] * R.g[m + i * NX];
				R.lgB[r][i] = lg;
			}
			R.hB[r] = hB[r];
			R.rhsB[r] = -lf;
		}
	}
	// cost, bounds (:238-254), QP, post-solve (:334-353)
	double v[NV];
	DiagMetric<NV> mt;
#pragma unroll
	for (int i = 0; i < NU; i++) {
		R.lb[i] = p.lb[i];
		R.ub[i] = p.ub[i];
	}
	R.lb[NU] = p.relaxSafeLb;
	R.lb[NU + 1] = p.relaxReachLb;
	R.ub[NU] = p.inf;
	R.ub[NU + 1] = p.inf;
#pragma unroll
	for (int i = 0; i < NV; i++) {
		mt.gi[i] = p.gi[i];
		mt.gih[i] = p.gih[i];
	}
	int iters = 0;
	const int st = qp_gi_solve<NV>(mt, c, R, v, &iters);
	double uo[NU], r0 = 0.0, r1 = 0.0;
	int32_t rc;
	if (st == QP_OK) {
#pragma unroll
		for (int i = 0; i < NU; i++) uo[i] = input_saturate(v[i], p.lb[i], p.ub[i]);
		r0 = v[NU];
		r1 = v[NU + 1];
		rc = 1;
	} else {
		double Du[NU * NX];
		M::backup_controller(x0, uo, Du);
#pragma unroll
		for (int i = 0; i < NU; i++) uo[i] = input_saturate(uo[i], p.lb[i], p.ub[i]);
		rc = -1;
	}
	if (live) {
#pragma unroll
		for (int i = 0; i < NU; i++) u_act[k * NU + i] = uo[i];
		relax_out[k * 2] = r0;
		relax_out[k * 2 + 1] = r1;
		rc_out[k] = rc;
		if (WITH_DIAG) {
			double *d = diag + k * NDIAG;
			d[0] = hSafetyNow;
			d[1] = hBackupEnd;
#pragma unroll
			for (int s = 0; s < CAP; s++)
				if (s < np) d[2 + s] = (double)kidx_s[R.order.slot(s) * T];
			double *A = d + 2 + np, *b = A + NC * NV;
			R.scan(
			    [&](const int j, const double(&nn)[NV], const double rhs) {
#pragma unroll
				    for (int i = 0; i < NV; i++) A[j + i * NC] = nn[i];
				    b[j] = rhs;
			    },
			    [](const int, const int, const bool, const double) {});
		}
	}
	if (qp_iter_sum) {
		unsigned int it = live ? (unsigned int)iters : 0u;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) it += __shfl_xor_sync(0xffffffffu, it, o);
		if ((threadIdx.x & 31) == 0 && it) qp_rows_add(qp_iter_sum, (unsigned long long)it);
	}
}

// =====================================================================================================================
// Checkpointed variant (round 2): the same filter with NO snapshots in the hot loop.
//
// The kernel above keeps the (x_i, Q_i) snapshot of every point that enters the running list of the CAP smallest min-h
// points: 480 B of shared memory per state for CAP = 10, which caps the SM at 12 warps, and a trajectory whose min h keeps
// falling (a pendulum the saturated backup input cannot catch: a third of the C3a draw) stores a snapshot on every one
// of its 5000 steps.  ncu (profiles/r01_c3a_*): FP64 pipe 43 % busy, 0.86 eligible warps per cycle - latency bound.
//
// Here pass A integrates (x, Q) and keeps only the list itself - keys and trajectory indices, 120 B of shared memory per
// state - plus one CHECKPOINT of (x, Q) [and the zero-order-hold state of ASIFimplicitRB] every 16 / 32 / 64 steps in a
// global scratch (coalesced 256 B stores per warp, 15 KB per resident thread for C3a, written once per tile).  Pass B then re-integrates, from the
// checkpoint before it, up to each selected point and stores its snapshot (global scratch as well: written once, read by
// the one or two row scans of the QP).  The Euler step is the same inlined function in both passes, so pass B retraces
// pass A's trajectory bit for bit.  In pass B every lane works on its own segments (the lanes of a warp gather different
// checkpoints and run the same instructions on them).  Cost: at most np segments of <= 64 steps per state against the
// npBT - 1 steps of pass A (typically 1-2 segments: the selected points cluster at the start, at the end or around one excursion).
// The scratch belongs to the resident grid, not to the batch: the kernel is persistent (grid = SMs x CTAs per SM, tiles
// strided), every thread re-uses its own area for all its tiles.  Used up to npBT = 20481; longer horizons run the kernel above.
// Checkpoint spacing 2^ck_log steps, chosen per engine (imp_ck_log): 16 steps up to npBT 5121 (C3a: 313 checkpoints, 15 KB of
// scratch per resident thread), 32 up to 10241, 64 up to 20481 - pass B costs a warp at most (2^ck_log - 1) steps per round.
#ifndef IMP_CK_LOG_MIN
#define IMP_CK_LOG_MIN 4
#endif
constexpr int IMP_MAX_CKPT = 320;
__host__ __device__ inline int imp_ck_log(const int npBT)
{
	for (int l = IMP_CK_LOG_MIN; l <= 6; l++)
		if ((((npBT - 1) >> l) + 1) <= IMP_MAX_CKPT) return l;
	return -1; // too long for the checkpoint table: the shared-memory kernel runs
}
constexpr int IMP2_THREADS = 128;

template <class M, bool RB>
__host__ __device__ constexpr int imp_ckpt_doubles()
{
	return M::NX + M::NX * M::NX + (RB ? M::NU + M::NU * M::NX + 1 : 0);
}
// doubles of global scratch per thread for a trajectory of npBT points
template <class M, int NPBTSS, bool RB>
__host__ __device__ constexpr size_t imp_scratch_doubles_per_thread(const int npBT, const int ck_log)
{
	return (size_t)(((npBT - 1) >> ck_log) + 1) * imp_ckpt_doubles<M, RB>() + (size_t)np_capacity(NPBTSS) * (M::NX + M::NX * M::NX);
}
template <int NPBTSS>
__host__ __device__ constexpr int imp2_smem_doubles_per_thread()
{
	return 2 * (np_capacity(NPBTSS) + (np_capacity(NPBTSS) + 1) / 2); // block list + point list: keys and indices
}

// one Euler step of the augmented backup flow, X_i = X_{i-1} + dt rhs(X_{i-1}) (src/asif_implicit.cpp:461-484), and min_j h_j(x_i)
template <class M, int SATMODE, bool RB, int KIND = 0>
__device__ __forceinline__ double imp_advance(const ImplicitParams &p, const int i, double (&X)[M::NX + M::NX * M::NX], ZohState<M> &zoh,
                                              TrigSC &tr)
{
	constexpr int NX = M::NX, NPSS = M::NPSS, NS = NX + NX * NX;
	double Xd[NS], DfCL[NX * NX];
	if constexpr (imp_use_trig<M>()) {
		imp_trig_step<M, KIND>(i, X[M::TRIG_ANGLE], tr);
		if (RB)
			backup_cl_dynamics_zoh<M, SATMODE, TrigSC>(p.sat, p.lb, p.ub, X, (double)(unsigned)i * p.backTrajDt, p.backTrajDt, p.backContDt, zoh, Xd, DfCL, tr);
		else
			backup_cl_dynamics<M, SATMODE, TrigSC>(p.sat, p.lb, p.ub, X, Xd, DfCL, tr);
	} else
	if (RB) // the reference hands t = i*backTrajDt to the rhs (src/asif_implicit_robust.cpp:550)
		backup_cl_dynamics_zoh<M, SATMODE>(p.sat, p.lb, p.ub, X, (double)(unsigned)i * p.backTrajDt, p.backTrajDt, p.backContDt, zoh, Xd, DfCL);
	else
		backup_cl_dynamics<M, SATMODE>(p.sat, p.lb, p.ub, X, Xd, DfCL);
	sensitivity_rhs<M>(DfCL, X + NX, Xd + NX);
#pragma unroll
	for (int e = 0; e < NS; e++) X[e] = Xd[e] * p.backTrajDt + X[e];
	if (M::HAS_SAFETY_MIN) return M::safety_min(X);
	double hs[NPSS], Dhs[NPSS * NX];
	M::safety_set(X, hs, Dhs);
	double hmin = hs[0];
#pragma unroll
	for (int j = 1; j < NPSS; j++) hmin = (hs[j] < hmin) ? hs[j] : hmin;
	return hmin;
}

#ifndef IMP2_MIN_BLOCKS
#define IMP2_MIN_BLOCKS 5 // 5 CTAs = 20 warps per SM at <= 102 registers: measured 25.5 ms against 31.1 (4) and 28.6 (3) on C3a
#endif
#ifndef IMP2_MIN_BLOCKS_RB
#define IMP2_MIN_BLOCKS_RB 6 // with the zero-order hold: 80 registers, 24 warps per SM: 21.78 -> 21.33 ms at npBT 5001 (C3a: 18.00 / 18.06, kept at 5)
#endif
template <class M, int NPBTSS, bool WITH_DIAG, int SATMODE, bool RB = false>
__global__ void __launch_bounds__(IMP2_THREADS, RB ? IMP2_MIN_BLOCKS_RB : IMP2_MIN_BLOCKS)
implicit_ckpt_kernel(const ImplicitParams p, const int64_t n, const double *__restrict__ x_in, const double *__restrict__ u_des,
                     double *__restrict__ u_act, double *__restrict__ relax_out, int32_t *__restrict__ rc_out,
                     double *__restrict__ diag, unsigned long long *__restrict__ qp_iter_sum, double *gscratch)
{
	constexpr int NX = M::NX, NU = M::NU, NPSS = M::NPSS, NPBS = M::NPBS;
	constexpr int NS = NX + NX * NX, NSC = imp_ckpt_doubles<M, RB>();
	constexpr int NV = NU + 2, CAP = np_capacity(NPBTSS), T = IMP2_THREADS;
	const int np = np_runtime(NPBTSS) ? p.npBTSS : NPBTSS;
	const int NC = np * NPSS + NPBS;
	const int NDIAG = 2 + np + NC * NV + NC;
	const int N = p.npBT;
	const int IMP_CK_LOG = p.ck_log, IMP_CK = 1 << IMP_CK_LOG; // run-time spacing (see imp_ck_log)
	const int NCK = ((N - 1) >> IMP_CK_LOG) + 1;

	extern __shared__ double smem[];
	constexpr int LST = CAP + (CAP + 1) / 2; // doubles of one list per thread
	double *keys = smem + threadIdx.x;                                              // point list: min h by slot
	int *kidx = reinterpret_cast<int *>(smem + CAP * T) + threadIdx.x;              //             trajectory index by slot
	double *bkeys = smem + LST * T + threadIdx.x;                                   // block list: min over the block by slot
	int *bidx = reinterpret_cast<int *>(smem + (LST + CAP) * T) + threadIdx.x;      //             block number by slot
	// this thread's scratch: checkpoints [c][e], then snapshots [slot][e]; element stride T (coalesced across the warp)
	double *ck = gscratch + (size_t)blockIdx.x * ((size_t)NCK * NSC + (size_t)CAP * NS) * T + threadIdx.x;
	double *snap = ck + (size_t)NCK * NSC * T;

	const int64_t tiles = (n + T - 1) / T;
	for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
		const int64_t k = tile * T + threadIdx.x;
		const bool live = k < n;
		const int64_t kk = live ? k : (n - 1);
		double x0[NX], c[NU + 2];
#pragma unroll
		for (int i = 0; i < NX; i++) x0[i] = x_in[kk * NX + i];
		if (p.custom_cost) { // the caller's c, all nv entries
#pragma unroll
			for (int i = 0; i < NU + 2; i++) c[i] = u_des[kk * (NU + 2) + i];
		} else { // updateCost(uDes) + the relax entries of initialize() (:238-254, :653-664)
#pragma unroll
			for (int i = 0; i < NU; i++) c[i] = -2.0 * u_des[kk * NU + i];
			c[NU] = -2.0 * p.relaxCost * p.relaxSafeLb;
			c[NU + 1] = -2.0 * p.relaxCost * p.relaxReachLb;
		}
		double hSafetyNow;
		{
			double hs[NPSS], Dhs[NPSS * NX];
			M::safety_set(x0, hs, Dhs);
			hSafetyNow = hs[0];
#pragma unroll
			for (int j = 1; j < NPSS; j++) hSafetyNow = (hs[j] < hSafetyNow) ? hs[j] : hSafetyNow;
		}
		ImpRows<M, NPBTSS, RB> R;
		if (RB) {
#pragma unroll
			for (int i = 0; i < NX; i++) R.x_unc[i] = p.x_unc[i];
		}
		R.np = np;
		R.snap = snap;
		R.T = T;
		double X[NS];
#pragma unroll
		for (int i = 0; i < NS; i++) X[i] = 0.0;
#pragma unroll
		for (int i = 0; i < NX; i++) X[i] = x0[i];
#pragma unroll
		for (int i = 0; i < NX; i++) X[NX + i * (NX + 1)] = 1.0;
		ZohState<M> zoh;
		TrigSC trig = {0.0, 1.0, 0.0}; // (imp_advance: re-synchronised at the first step after every checkpoint)
#pragma unroll
		for (int i = 0; i < NU; i++) zoh.u[i] = 0.0;
#pragma unroll
		for (int i = 0; i < NU * NX; i++) zoh.Du[i] = 0.0;
		zoh.t_last = -1.;
#pragma unroll
		for (int s = 0; s < CAP; s++) {
			bkeys[s * T] = INFINITY;
			bidx[s * T] = -1;
		}
		ImpOrder<CAP> ob; // order of the block list
		ob.init();
		double bkmin = INFINITY, bkmax = INFINITY;
		// checkpoint 0 = the initial point
#pragma unroll
		for (int e = 0; e < NS; e++) ck[e * T] = X[e];
		if (RB) {
#pragma unroll
			for (int i = 0; i < NU; i++) ck[(NS + i) * T] = zoh.u[i];
#pragma unroll
			for (int i = 0; i < NU * NX; i++) ck[(NS + NU + i) * T] = zoh.Du[i];
			ck[(NS + NU + NU * NX) * T] = zoh.t_last;
		}

		// ---- pass A: the whole horizon.  Per step only the running minimum of min h over the current BLOCK (the IMP_CK points
		// that follow a checkpoint); per block one insert into the list of the CAP blocks with the smallest minimum, ties
		// keeping the earlier block first, and the checkpoint store.  The CAP smallest POINTS (by min h, ties by index) all
		// lie in those blocks: a point p of block B among them has (min of B, B) <= p, and CAP blocks ahead of B in the list
		// would each hold a point ahead of p.  Point 0 is not in any block; pass B starts the point list with it.
		for (int c0 = 0; c0 < N - 1; c0 += IMP_CK) {
			const int iend = (c0 + IMP_CK < N - 1) ? c0 + IMP_CK : N - 1;
			double bmin = INFINITY;
#if ASIF_IMP_PEEL
			if constexpr (imp_use_trig<M>()) {
				// groups of 16 steps (c0 is a multiple of 16): the re-synchronisation step peeled off, then a loop without the test on i
				// unrolled by three without the zero-order hold (C3a 18.53 -> 18.01 ms; by five 18.76); with it the plain loop is
				// the fastest (ASIFimplicitRB 21.84 / 22.55 / 23.23 ms at 1 / 3 / 5)
				constexpr int PEEL_UNROLL = RB ? 1 : IMP_PEEL_UNROLL;
				for (int g0 = c0; g0 < iend; g0 += 16) {
					const int gend = (g0 + 16 < iend) ? g0 + 16 : iend;
					const double h1 = imp_advance<M, SATMODE, RB, 1>(p, g0 + 1, X, zoh, trig);
					bmin = (h1 < bmin) ? h1 : bmin;
#pragma unroll PEEL_UNROLL
					for (int i = g0 + 2; i <= gend; i++) {
						const double hmin = imp_advance<M, SATMODE, RB, 2>(p, i, X, zoh, trig);
						bmin = (hmin < bmin) ? hmin : bmin;
					}
				}
			} else
#endif
			for (int i = c0 + 1; i <= iend; i++) {
				const double hmin = imp_advance<M, SATMODE, RB>(p, i, X, zoh, trig);
				bmin = (hmin < bmin) ? hmin : bmin;
			}
			if (bmin < bkmax) {
				const int slot = ob.last(); // the evicted (largest) entry's slot takes the new block
				if (bmin < bkmin) {
					ob.push_front(slot);
					bkmin = bmin;
				} else {
					int pl = 1; // first place whose key is larger (behind equal keys: ties keep the earlier block first)
					while (pl < CAP - 1 && !(bmin < bkeys[ob.slot(pl) * T])) pl++;
					ob.insert_at(pl, slot);
				}
				bkeys[slot * T] = bmin;
				bidx[slot * T] = c0 >> IMP_CK_LOG;
				bkmax = bkeys[ob.last() * T];
			}
			if (iend == c0 + IMP_CK) { // the state after step iend = a multiple of the spacing: checkpoint iend >> ck_log
				double *q = ck + (size_t)(iend >> IMP_CK_LOG) * NSC * T;
#pragma unroll
				for (int e = 0; e < NS; e++) q[e * T] = X[e];
				if (RB) {
#pragma unroll
					for (int t = 0; t < NU; t++) q[(NS + t) * T] = zoh.u[t];
#pragma unroll
					for (int t = 0; t < NU * NX; t++) q[(NS + NU + t) * T] = zoh.Du[t];
					q[(NS + NU + NU * NX) * T] = zoh.t_last;
				}
			}
		}
		// open-loop dynamics at the current state (:414-416) and the backup rows at the trajectory end (:542-554)
		M::dynamics(x0, R.f, R.g);
		double hBackupEnd;
		{
			double hB[NPBS], DhB[NPBS * NX];
			M::backup_set_rows(X, hB, DhB);
			hBackupEnd = hB[0];
#pragma unroll
			for (int r = 1; r < NPBS; r++) hBackupEnd = (hB[r] < hBackupEnd) ? hB[r] : hBackupEnd;
#pragma unroll
			for (int r = 0; r < NPBS; r++) {
				double dh[NX];
#pragma unroll
				for (int cc = 0; cc < NX; cc++) {
					double acc = DhB[r] * X[NX + cc * NX];
#pragma unroll
					for (int m = 1; m < NX; m++) acc = acc + DhB[r + m * NPBS] * X[NX + m + cc * NX];
					dh[cc] = acc;
				}
				double lf = dh[0] * R.f[0];
#pragma unroll
				for (int m = 1; m < NX; m++) lf = lf + dh[m] * R.f[m];
#pragma unroll
				for (int i = 0; i < NU; i++) {
					double lg = dh[0] * R.g[i * NX];
#pragma unroll
					for (int m = 1; m < NX; m++) lg = lg + dh[m] * R.g[m + i * NX];
					R.lgB[r][i] = lg;
				}
				R.hB[r] = hB[r];
				R.rhsB[r] = -lf;
			}
		}

		// ---- pass B: the point list (the CAP smallest points by (min h, index)) and its snapshots, from the listed blocks
		// re-integrated out of their checkpoints.  Blocks are taken in ascending order of their minimum; once the point list is
		// full, a block whose minimum cannot get in ends the lane's work (every later block has a larger minimum or the same
		// minimum and a later index).  A visited block always leaves its minimum in the final list (later blocks only hold
		// points behind it), so the number of rounds is the number of blocks that hold final points: typically 1-2, the
		// selected points cluster at the start, at the end or around one excursion.  Every lane works on ITS OWN block (the
		// lanes gather their checkpoints from different places and run the same instruction stream on them); a round costs
		// the warp one block's worth of steps.  The Euler step is the same inlined function as in pass A: bit for bit the
		// same trajectory.
#ifndef IMP2_SKIP_PASS_B // (timing experiments only: results are garbage without it)
		{
#pragma unroll
			for (int s = 1; s < CAP; s++) {
				keys[s * T] = INFINITY;
				kidx[s * T] = -1;
			}
			keys[0] = hSafetyNow; // point 0 = the current state, Q = I
			kidx[0] = 0;
			R.order.init();
#pragma unroll
			for (int e = 0; e < NS; e++) snap[(size_t)e * T] = 0.0;
#pragma unroll
			for (int e = 0; e < NX; e++) snap[(size_t)e * T] = x0[e];
#pragma unroll
			for (int e = 0; e < NX; e++) snap[(size_t)(NX + e * (NX + 1)) * T] = 1.0;
			double kmax = (CAP > 1) ? INFINITY : hSafetyNow;
			int nb = 0; // blocks of the list looked at so far
			for (;;) {
				int mine = -1;
				if (nb < CAP) {
					const int sl = ob.slot(nb);
					const int b = bidx[sl * T];
					const double bk = bkeys[sl * T];
					nb++;
					const bool wanted = b >= 0 && ((bk < kmax) || (bk == kmax && (b << IMP_CK_LOG) + 1 < kidx[R.order.last() * T]));
					if (wanted)
						mine = b;
					else
						nb = CAP;
				}
				const bool have = mine >= 0;
				if (!__any_sync(0xffffffffu, have)) break;
				int nsteps = 0;
				const int base_i = have ? (mine << IMP_CK_LOG) : 0;
				if (have) {
					const double *q = ck + (size_t)mine * NSC * T;
#pragma unroll
					for (int e = 0; e < NS; e++) X[e] = q[e * T];
					if (RB) {
#pragma unroll
						for (int t = 0; t < NU; t++) zoh.u[t] = q[(NS + t) * T];
#pragma unroll
						for (int t = 0; t < NU * NX; t++) zoh.Du[t] = q[(NS + NU + t) * T];
						zoh.t_last = q[(NS + NU + NU * NX) * T];
					}
					nsteps = (N - 1 - base_i < IMP_CK) ? (N - 1 - base_i) : IMP_CK;
				}
				const int wsteps = __reduce_max_sync(0xffffffffu, nsteps);
				for (int j = 1; j <= wsteps; j++) {
					if (j <= nsteps) {
						const int i = base_i + j;
						const double hmin = imp_advance<M, SATMODE, RB>(p, i, X, zoh, trig);
						const int slot = R.order.last(); // the evicted (largest) entry's slot takes the new point
						if (hmin < kmax || (hmin == kmax && i < kidx[slot * T])) {
							int pl = 0; // first place that (hmin, i) comes before: ties keep the earlier index first
							while (pl < CAP - 1) {
								const int sp = R.order.slot(pl);
								const double kp = keys[sp * T];
								if (hmin < kp || (hmin == kp && i < kidx[sp * T])) break;
								pl++;
							}
							R.order.insert_at(pl, slot);
							keys[slot * T] = hmin;
							kidx[slot * T] = i;
#pragma unroll
							for (int e = 0; e < NS; e++) snap[(size_t)(slot * NS + e) * T] = X[e];
							kmax = keys[R.order.last() * T];
						}
					}
				}
			}
		}
#endif
		R.learn = p.learn.blob != nullptr;
		R.dLf = 0.0;
#pragma unroll
		for (int i = 0; i < NU; i++) R.dLg[i] = 0.0;
		if (R.learn) {
			// Dh_index_ = DhSS(x_c) Q_c of the first critical point, npSS x nx column-major; its first nx entries
			// (src/asif_implicit.cpp:533-537, include/asif_learning_utils.h:127-129)
			double xs[NS], hq[NPSS], Dq[NPSS * NX], dhi[NX];
#pragma unroll
			for (int e = 0; e < NS; e++) xs[e] = snap[(size_t)(R.order.slot(0) * NS + e) * T];
			M::safety_set(xs, hq, Dq);
#pragma unroll
			for (int i = 0; i < NX; i++) {
				const int row = i % NPSS, col = i / NPSS;
				double acc = 0.0;
#pragma unroll
				for (int m = 0; m < NX; m++) acc = acc + Dq[row + m * NPSS] * xs[NX + m + col * NX];
				dhi[i] = acc;
			}
			learned_residual<NX, NU>(p.learn, x0, dhi, R.dLf, R.dLg);
		}
		// cost, bounds (:238-254), QP, post-solve (:334-353)
		double v[NV];
		DiagMetric<NV> mt;
#pragma unroll
		for (int i = 0; i < NU; i++) {
			R.lb[i] = p.lb[i];
			R.ub[i] = p.ub[i];
		}
		R.lb[NU] = p.relaxSafeLb;
		R.lb[NU + 1] = p.relaxReachLb;
		R.ub[NU] = p.inf;
		R.ub[NU + 1] = p.inf;
#pragma unroll
		for (int i = 0; i < NV; i++) {
			mt.gi[i] = p.gi[i];
			mt.gih[i] = p.gih[i];
		}
		int iters = 0;
		const int st = qp_gi_solve<NV>(mt, c, R, v, &iters);
		double uo[NU], r0 = 0.0, r1 = 0.0;
		int32_t rc;
		if (st == QP_OK) {
#pragma unroll
			for (int i = 0; i < NU; i++) uo[i] = input_saturate(v[i], p.lb[i], p.ub[i]);
			r0 = v[NU];
			r1 = v[NU + 1];
			rc = 1;
		} else {
			double Du[NU * NX];
			M::backup_controller(x0, uo, Du);
#pragma unroll
			for (int i = 0; i < NU; i++) uo[i] = input_saturate(uo[i], p.lb[i], p.ub[i]);
			rc = -1;
		}
		if (live) {
#pragma unroll
			for (int i = 0; i < NU; i++) u_act[k * NU + i] = uo[i];
			relax_out[k * 2] = r0;
			relax_out[k * 2 + 1] = r1;
			rc_out[k] = rc;
			if (WITH_DIAG) {
				double *d = diag + k * NDIAG;
				d[0] = hSafetyNow;
				d[1] = hBackupEnd;
#pragma unroll
				for (int s = 0; s < CAP; s++)
					if (s < np) d[2 + s] = (double)kidx[R.order.slot(s) * T];
				double *A = d + 2 + np, *b = A + NC * NV;
				R.scan(
				    [&](const int j, const double(&nn)[NV], const double rhs) {
#pragma unroll
					    for (int i = 0; i < NV; i++) A[j + i * NC] = nn[i];
					    b[j] = rhs;
				    },
				    [](const int, const int, const bool, const double) {});
			}
		}
		if (qp_iter_sum) {
			unsigned int it = live ? (unsigned int)iters : 0u;
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) it += __shfl_xor_sync(0xffffffffu, it, o);
			if ((threadIdx.x & 31) == 0 && it) qp_rows_add(qp_iter_sum, (unsigned long long)it);
		}
	}
}

} // namespace asifb
