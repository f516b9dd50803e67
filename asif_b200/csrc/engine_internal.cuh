// engine_internal.cuh -- what the translation units of libasif_b200.so share: the engine object, error plumbing and the
// launch helpers of the TB kernels.  engine.cu holds the C ABI and every kernel whose arithmetic is bit-identical to
// the reference build (compiled with -fmad=false); kernels_contract.cu holds the kernels of the models that call
// libm-class functions (sin, cos, tanh), compiled with FMA contraction (see the note there).
#pragma once
#include "../../include/asif_b200.h"

#include "explicit_kernel.cuh"
#include "filter_common.cuh"
#include "implicit_kernel.cuh"
#include "robust_kernel.cuh"
#include "realizable_kernel.cuh"
#include "models.cuh"
#include "tb_kernel.cuh"

#include <cstdint>

using namespace asifb;

namespace asifb {
// records the message asif_last_error() returns (thread local) and hands the code back; defined in engine.cu
int fail(int code, const char *fmt, ...);
}
using asifb::fail;

#include "host_copier.hpp"

namespace {

#define CUDA_TRY(expr)                                                                                         \
	do {                                                                                                       \
		cudaError_t e__ = (expr);                                                                              \
		if (e__ != cudaSuccess)                                                                                \
			return fail(ASIF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
	} while (0)

constexpr size_t SMALL_BATCH_BYTES = 64 << 10; // host batches up to this size run in place on the engine's mapped scratch
constexpr int N_SLOTS = 4;               // pipeline depth of the host-memory path
constexpr int64_t CHUNK_STATES = 1 << 19; // states per pipeline chunk (env ASIF_B200_CHUNK_STATES overrides); measured
                                          // e2e for 1e7 C2 states: 2^17 6.7 ms, 2^18 6.1, 2^19 5.7, 2^20 6.1, 2^21 6.3

inline int64_t chunk_states()
{
	static int64_t v = 0;
	if (v == 0) {
		const char *e = getenv("ASIF_B200_CHUNK_STATES");
		const long long x = e ? atoll(e) : 0;
		v = (x >= 1024) ? (int64_t)x : CHUNK_STATES;
	}
	return v;
}

constexpr int N_SNAP_BUFS = 6;

struct Slot {
	cudaStream_t stream = nullptr;
	double *x = nullptr, *ud = nullptr, *ua = nullptr, *relax = nullptr, *diag = nullptr;
	int32_t *rc = nullptr;
	int64_t cap = 0, cap_diag = 0;
	// pinned staging for pageable caller arrays (bounce path): [x | ud | ua | relax | diag | rc], and the chunk whose
	// results sit in it waiting to be copied out
	char *h = nullptr;
	size_t h_cap = 0;
	int64_t pend_off = 0, pend_m = 0;
};

} // namespace

struct asif_engine {
	asif_engine_config cfg;
	int nx, nu, n_relax, nc, nv, n_diag;
	TbParams tb;
	ExplicitParams ex;
	ImplicitParams im;
	RobustParams rb;
	RealizableParams rz;
	double *d_ttable = nullptr; // t_i of the backup trajectory (TB filter)
	void *d_kernel = nullptr; // polytope kernel tables (realizable filter), one allocation
	size_t rz_smem = 0;
	double *d_table = nullptr; // half-plane table (robust filter)
	double *d_learn = nullptr; // learned-residual networks (implicit filters)
	const double *lie_lfh = nullptr, *lie_lgh = nullptr; // device arrays of the current filter_batch_lie call (explicit filter)
	Slot slot[N_SLOTS];
	cudaStream_t stream = nullptr; // device-memory calls without a caller stream
	unsigned long long *d_counters = nullptr; // counters on the device (filter_common.cuh: rc histogram, QP-work partial sums)
	unsigned long long *ctr = nullptr;        // what the filter launches pass: d_counters, or nullptr while a small (latency) batch runs
	uint64_t last_qp_rows = 0;
	HostCopier *copier = nullptr; // created by the first large pageable batch
	int copy_threads = 0;         // its thread count; 0 = HostCopier::default_threads() (a group shares the cores out)
	char *small_h = nullptr;      // pinned, device-addressable scratch of the small-batch path (SMALL_BATCH_BYTES)
	int last_host_io = -1; // ASIF_HOST_IO_* actually used by the last host-memory batch (-1: none yet)
	// "auto" host-IO policy: wall time per state of large pinned batches per mode (exponential average), batches seen
	double io_ms_per_state[3] = {0.0, 0.0, 0.0};
	int io_samples[3] = {0, 0, 0};
	unsigned io_calls = 0;
	int num_sms = 148;
	long long l2_persist_max = -1, l2_window_max = 0; // persisting-L2 limits of the device (-1: not queried yet)
	// latency server (latency_server.cuh): pinned mailbox the device can address, its stream, the last sequence number
	double *srv_mailbox = nullptr, *srv_mailbox_dev = nullptr, *srv_snap = nullptr;
	cudaStream_t srv_stream = nullptr;
	bool srv_on = false;
	unsigned long long srv_seq = 0;
	// snapshot scratch of the persistent nx = 4 kernels: a ring, so that launches on different streams never share
	// a buffer that may still be in use (the next user waits on the previous user's event)
	struct SnapBuf {
		double *p = nullptr;
		size_t cap = 0;
		cudaEvent_t ev = nullptr;
		cudaStream_t last = nullptr;
		bool used = false;
	} snapbuf[N_SNAP_BUFS];
	int snap_next = 0;
};


template <class K>
int set_smem(K kernel, size_t bytes)
{
	if (bytes > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
	return ASIF_OK;
}

// L2 residency of a scratch buffer for the launches that follow on stream st (bytes = 0 clears the window again).
// The snapshot scratch of the persistent nx = 4 kernels is rewritten tile after tile by the same threads and only ever read
// back by them: nothing in it needs to reach HBM, but with the default policy its lines were evicted between two uses and
// 119 of the 142 MB of DRAM traffic of a 2e5-state segway launch were such write-backs (profiles/r01_c5_*; 11.8x the 12 MB
// of algorithmic traffic).  Marking the window persisting (set-aside L2, cudaLimitPersistingL2CacheSize) keeps the lines on
// chip until they are overwritten.  ASIF_B200_L2_PERSIST=0 switches it off.  Failures are ignored: this is a hint.
inline void scratch_l2_window(asif_engine *e, cudaStream_t st, void *p, size_t bytes)
{
	static const bool enabled = []() {
		const char *v = getenv("ASIF_B200_L2_PERSIST");
		return !(v && v[0] == '0');
	}();
	if (!enabled) return;
	if (e->l2_persist_max < 0) { // once per engine: the device's limits, and the set-aside at its maximum
		cudaDeviceProp prop;
		e->l2_persist_max = 0;
		if (cudaGetDeviceProperties(&prop, e->cfg.device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
			if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)prop.persistingL2CacheMaxSize) == cudaSuccess) {
				e->l2_persist_max = (long long)prop.persistingL2CacheMaxSize;
				e->l2_window_max = (long long)prop.accessPolicyMaxWindowSize;
			}
		}
		cudaGetLastError();
	}
	if (e->l2_persist_max <= 0) return;
	cudaStreamAttrValue a;
	memset(&a, 0, sizeof(a));
	if (bytes > 0) {
		const size_t win = bytes < (size_t)e->l2_window_max ? bytes : (size_t)e->l2_window_max;
		a.accessPolicyWindow.base_ptr = p;
		a.accessPolicyWindow.num_bytes = win;
		a.accessPolicyWindow.hitRatio = win <= (size_t)e->l2_persist_max ? 1.0f : (float)((double)e->l2_persist_max / (double)win);
		a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
		a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
	}
	cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &a);
	cudaGetLastError();
}

// A scratch buffer of `need` bytes for a launch on stream st, out of the engine's ring: one last used on this same stream
// (stream order protects it), else an idle one, else an unallocated one, else the next in the ring after waiting for its
// last user.  The caller launches and then calls tb_release(e, buf, st).
inline int acquire_scratch(asif_engine *e, const size_t need, cudaStream_t st, int &buf)
{
	int pick = -1;
	for (int i = 0; i < N_SNAP_BUFS && pick < 0; i++) {
		asif_engine::SnapBuf &c = e->snapbuf[i];
		if (c.p && c.cap >= need && c.used && c.last == st) pick = i;
	}
	for (int i = 0; i < N_SNAP_BUFS && pick < 0; i++) {
		asif_engine::SnapBuf &c = e->snapbuf[i];
		if (c.p && c.cap >= need && (!c.used || cudaEventQuery(c.ev) == cudaSuccess)) {
			c.used = false;
			pick = i;
		}
	}
	for (int i = 0; i < N_SNAP_BUFS && pick < 0; i++)
		if (!e->snapbuf[i].p) pick = i;
	if (pick < 0) {
		pick = e->snap_next;
		e->snap_next = (e->snap_next + 1) % N_SNAP_BUFS;
	}
	buf = pick;
	asif_engine::SnapBuf &b = e->snapbuf[buf];
	if (b.cap < need) {
		if (b.p) CUDA_TRY(cudaFree(b.p)); // synchronises with any kernel still using it
		b.p = nullptr;
		b.cap = 0;
		b.used = false;
		CUDA_TRY(cudaMalloc(&b.p, need));
		b.cap = need;
	}
	if (!b.ev) CUDA_TRY(cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming));
	if (b.used && b.last != st) CUDA_TRY(cudaStreamWaitEvent(st, b.ev, 0));
	return ASIF_OK;
}

// ---- kernel dispatch -------------------------------------------------------------------------
// Launch geometry of the TB kernels.  nx = 2: one CTA per tile of TB_THREADS states, snapshots in shared memory.
// nx = 4: persistent grid over a global snapshot scratch; the grid is the largest one that fits the SMs, shrunk so
// that every CTA runs the same number of tiles (no ragged last wave).
template <class M, int NPBTSS>
int tb_geometry(asif_engine *e, int64_t n, cudaStream_t st, unsigned &blocks, size_t &smem, double *&gsnap, int &buf,
                const int per_sm = tb_min_blocks<M>())
{
	const int64_t tiles = (n + TB_THREADS - 1) / TB_THREADS;
	const size_t per_cta = sizeof(double) * tb_smem_doubles_per_thread<M, NPBTSS>() * TB_THREADS;
	gsnap = nullptr;
	buf = -1;
	if (!tb_global_snapshots<M>()) {
		blocks = (unsigned)tiles;
		smem = per_cta;
		return ASIF_OK;
	}
	const int64_t resident = (int64_t)e->num_sms * per_sm;
	const int64_t rounds = (tiles + resident - 1) / resident;
	blocks = (unsigned)((tiles + rounds - 1) / rounds);
	if (blocks < 1) blocks = 1;
	smem = 0;
	const size_t need = sizeof(double) * TB_SCRATCH_HEADER + per_cta * (size_t)resident;
	int r_;
	r_ = acquire_scratch(e, need, st, buf);
	if (r_) return r_;
	asif_engine::SnapBuf &b = e->snapbuf[buf];
	CUDA_TRY(cudaMemsetAsync(b.p, 0, sizeof(double) * TB_SCRATCH_HEADER, st)); // tile counter
	gsnap = b.p;
	scratch_l2_window(e, st, b.p, need);
	return ASIF_OK;
}

inline int tb_release(asif_engine *e, int buf, cudaStream_t st)
{
	if (buf >= 0) {
		scratch_l2_window(e, st, nullptr, 0); // the window applies to the launch just made, not to what the stream runs next
		CUDA_TRY(cudaEventRecord(e->snapbuf[buf].ev, st));
		e->snapbuf[buf].used = true;
		e->snapbuf[buf].last = st;
	}
	return ASIF_OK;
}

template <class M, int NPBTSS, int SATMODE>
int launch_tb_mode(asif_engine *e, int64_t n, const double *x, const double *ud, double *ua, double *relax, int32_t *rc,
                   double *diag, cudaStream_t st)
{
	unsigned blocks;
	size_t smem;
	double *gsnap;
	int buf;
	int r = tb_geometry<M, NPBTSS>(e, n, st, blocks, smem, gsnap, buf);
	if (r) return r;
	if (diag) {
		auto k = tb_filter_kernel<M, NPBTSS, true, SATMODE>;
		r = set_smem(k, smem);
		if (r) return r;
		k<<<blocks, TB_THREADS, smem, st>>>(e->tb, n, x, ud, ua, relax, rc, diag, e->ctr, gsnap);
	} else {
		auto k = tb_filter_kernel<M, NPBTSS, false, SATMODE>;
		r = set_smem(k, smem);
		if (r) return r;
		k<<<blocks, TB_THREADS, smem, st>>>(e->tb, n, x, ud, ua, relax, rc, nullptr, e->ctr, gsnap);
	}
	CUDA_TRY(cudaGetLastError());
	return tb_release(e, buf, st);
}

// SAT_POW2 and SAT_IDENTITY are exact shortcuts of SAT_GENERAL; IDENTITY is only instantiated for the nx = 2 models
template <class M, int NPBTSS>
int launch_tb(asif_engine *e, int64_t n, const double *x, const double *ud, double *ua, double *relax, int32_t *rc,
              double *diag, cudaStream_t st)
{
	if (e->tb.sat_mode == SAT_IDENTITY && M::NX <= 2) return launch_tb_mode<M, NPBTSS, SAT_IDENTITY>(e, n, x, ud, ua, relax, rc, diag, st);
	if (e->tb.sat_mode >= SAT_POW2) return launch_tb_mode<M, NPBTSS, SAT_POW2>(e, n, x, ud, ua, relax, rc, diag, st);
	return launch_tb_mode<M, NPBTSS, SAT_GENERAL>(e, n, x, ud, ua, relax, rc, diag, st);
}


template <class M, int NPBTSS>
int launch_tb_rollout(asif_engine *e, int64_t n, int32_t steps, double dt, double *x, const double *ud, double *ua,
                      int32_t *rc, cudaStream_t st)
{
	unsigned blocks;
	size_t smem;
	double *gsnap;
	int buf;
	int r = tb_geometry<M, NPBTSS>(e, n, st, blocks, smem, gsnap, buf, tb_rollout_min_blocks<M>());
	if (r) return r;
	auto k = (e->tb.sat_mode >= SAT_POW2) ? tb_rollout_kernel<M, NPBTSS, SAT_POW2> : tb_rollout_kernel<M, NPBTSS, SAT_GENERAL>;
	r = set_smem(k, smem);
	if (r) return r;
	k<<<blocks, TB_THREADS, smem, st>>>(e->tb, n, steps, dt, x, ud, ua, rc, e->d_counters + 1, e->d_counters, gsnap);
	CUDA_TRY(cudaGetLastError());
	return tb_release(e, buf, st);
}

// ASIFimplicit / ASIFimplicitRB launch for one model.  SAT_LO is the saturation mode used when the input range is not a
// power of two: SAT_GENERAL in the bit-exact unit (engine.cu), SAT_RECIP in the contracted one (kernels_contract.cu).
template <class M, int NPBTSS, bool RB, int SAT_LO>
int launch_implicit_t(asif_engine *e, int64_t n, const double *x, const double *ud, double *ua, double *relax, int32_t *rc,
                      double *diag, cudaStream_t st)
{
	const bool pow2 = e->im.sat_mode >= SAT_POW2;
	// checkpointed kernel (no snapshots in the hot loop, persistent grid over a global scratch) whenever the horizon fits
	// its checkpoint table and the scratch stays below 2 GB; ASIF_B200_IMPLICIT_SMEM=1 forces the shared-memory kernel
	static const bool force_smem = []() {
		const char *v = getenv("ASIF_B200_IMPLICIT_SMEM");
		return v && v[0] == '1';
	}();
	const int ck_log = imp_ck_log(e->im.npBT);
	if (!force_smem && ck_log > 0) {
		const size_t smem2 = sizeof(double) * imp2_smem_doubles_per_thread<NPBTSS>() * IMP2_THREADS;
		const int64_t tiles = (n + IMP2_THREADS - 1) / IMP2_THREADS;
		auto launch = [&](auto kern) -> int {
			int per_sm = 0;
			CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, IMP2_THREADS, smem2));
			if (per_sm < 1) per_sm = 1;
			const int64_t resident = (int64_t)e->num_sms * per_sm;
			// every CTA runs the same number of tiles (no ragged last wave), as for the persistent TB kernels
			const int64_t rounds = (tiles + resident - 1) / resident;
			const unsigned blocks2 = (unsigned)((tiles + rounds - 1) / rounds);
			// a scratch beyond 2 GB (ASIFimplicitRB at npBT 5001: 10 doubles per checkpoint, 2.4 GB at 16 steps) takes a coarser
			// spacing first: any multiple of 16 keeps the checkpoints on the re-synchronisation steps of the trig recurrence
			int l = ck_log;
			size_t need;
			for (;; l++) {
				need = sizeof(double) * imp_scratch_doubles_per_thread<M, NPBTSS, RB>(e->im.npBT, l) * IMP2_THREADS * (size_t)blocks2;
				if (need <= ((size_t)2 << 30) || l == 6) break;
			}
			if (need > ((size_t)2 << 30)) return 1; // fall back
			e->im.ck_log = l;
			int buf;
			int r = acquire_scratch(e, need, st, buf);
			if (r) return r;
			kern<<<blocks2, IMP2_THREADS, smem2, st>>>(e->im, n, x, ud, ua, relax, rc, diag, e->ctr, e->snapbuf[buf].p);
			CUDA_TRY(cudaGetLastError());
			return tb_release(e, buf, st);
		};
		// symmetric input bounds (lb == -ub for every input) in the contracted unit: the cheaper form of SAT_RECIP, same bits
		bool sym = (SAT_LO == SAT_RECIP);
		for (int i = 0; i < e->nu; i++) sym = sym && (e->im.lb[i] == -e->im.ub[i]);
		constexpr int SAT_SYM = (SAT_LO == SAT_RECIP) ? SAT_RECIP_SYM : SAT_LO;
		int r;
		if (diag)
			r = pow2 ? launch(implicit_ckpt_kernel<M, NPBTSS, true, SAT_POW2, RB>) : launch(implicit_ckpt_kernel<M, NPBTSS, true, SAT_LO, RB>);
		else if (sym && !pow2)
			r = launch(implicit_ckpt_kernel<M, NPBTSS, false, SAT_SYM, RB>);
		else
			r = pow2 ? launch(implicit_ckpt_kernel<M, NPBTSS, false, SAT_POW2, RB>) : launch(implicit_ckpt_kernel<M, NPBTSS, false, SAT_LO, RB>);
		if (r != 1) return r;
	}
	const size_t smem = sizeof(double) * (diag ? imp_smem_doubles_per_thread<M, NPBTSS, true>() : imp_smem_doubles_per_thread<M, NPBTSS, false>()) * IMP_THREADS;
	const unsigned blocks = (unsigned)((n + IMP_THREADS - 1) / IMP_THREADS);
	if (diag) {
		auto k = pow2 ? implicit_filter_kernel<M, NPBTSS, true, SAT_POW2, RB> : implicit_filter_kernel<M, NPBTSS, true, SAT_LO, RB>;
		int r = set_smem(k, smem);
		if (r) return r;
		k<<<blocks, IMP_THREADS, smem, st>>>(e->im, n, x, ud, ua, relax, rc, diag, e->ctr);
	} else {
		auto k = pow2 ? implicit_filter_kernel<M, NPBTSS, false, SAT_POW2, RB> : implicit_filter_kernel<M, NPBTSS, false, SAT_LO, RB>;
		int r = set_smem(k, smem);
		if (r) return r;
		k<<<blocks, IMP_THREADS, smem, st>>>(e->im, n, x, ud, ua, relax, rc, nullptr, e->ctr);
	}
	CUDA_TRY(cudaGetLastError());
	return ASIF_OK;
}


// defined in engine.cu: the latency server's stop / start (latency_server.cuh)
namespace asifb {
int server_suspend(asif_engine *e);
int server_resume(asif_engine *e);
}
// RAII: stops the latency server for the duration of a call that may allocate, restarts it on the way out (see engine.cu)
struct ServerPause {
	asif_engine *e;
	bool was_on;
	explicit ServerPause(asif_engine *e_) : e(e_), was_on(e_->srv_on)
	{
		if (was_on) asifb::server_suspend(e);
	}
	~ServerPause()
	{
		if (was_on) asifb::server_resume(e);
	}
	ServerPause(const ServerPause &) = delete;
	ServerPause &operator=(const ServerPause &) = delete;
};

// defined in engine.cu: dispatch of one filter launch on device pointers
namespace asifb {
int launch_filter_any(asif_engine *e, int64_t n, const double *x, const double *ud, double *ua, double *relax, int32_t *rc,
                      double *diag, cudaStream_t st);
}

// defined in kernels_contract.cu
namespace asifb {
int launch_tb_segway(asif_engine *e, bool shipped, int64_t n, const double *x, const double *ud, double *ua, double *relax,
                     int32_t *rc, double *diag, cudaStream_t st);
int launch_tb_rollout_segway(asif_engine *e, bool shipped, int64_t n, int32_t steps, double dt, double *x, const double *ud,
                             double *ua, int32_t *rc, cudaStream_t st);
int launch_implicit_ip(asif_engine *e, int64_t n, const double *x, const double *ud, double *ua, double *relax, int32_t *rc,
                       double *diag, cudaStream_t st);
int launch_implicit_rb_ip(asif_engine *e, int64_t n, const double *x, const double *ud, double *ua, double *relax, int32_t *rc,
                          double *diag, cudaStream_t st);
} // namespace asifb
