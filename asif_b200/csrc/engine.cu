// engine.cu -- C ABI of libasif_b200.so (declared in include/asif_b200.h).
// Host side: option handling exactly as the reference's initialize() methods, device buffers,
// chunked H2D / kernel / D2H pipelining for host-memory batches, kernel dispatch.  No CPU
// fallback anywhere: every compute entry point needs a CUDA device and says so if there is none.
#define ASIF_QP_POLISH_INLINE_ROWS 1 // qp_gi.cuh: the form of the vertex polish that costs this unit's kernels least
#include "../../include/asif_b200.h"

#include "explicit_kernel.cuh"
#include "filter_common.cuh"
#include "implicit_kernel.cuh"
#include "robust_kernel.cuh"
#include "realizable_kernel.cuh"
#include "models.cuh"
#include "qp_batch_kernel.cuh"
#include "tb_kernel.cuh"
#include "latency_server.cuh"

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <stdexcept>

using namespace asifb;

#include "engine_internal.cuh"

namespace {
thread_local char g_err[512] = "";
int server_stop(asif_engine *e);  // latency server, defined with the host-batch code below
int server_start(asif_engine *e);
}

namespace asifb {
int fail(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}
} // namespace asifb

namespace asifb {
int launch_qp_admm(int device, int nv, int nc, int64_t n, int diag_cost, const double *H, const double *c, const double *A,
                   const double *b, const double *lb, const double *ub, const uint8_t *be, double *sol, int32_t *status,
                   int share_flags, cudaStream_t st, bool fetch_info); // qp_admm.cu
}

namespace {

int ensure_slot(asif_engine *e, Slot &s, int64_t n, bool want_diag)
{
	if (!s.stream) CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
	if (n > s.cap) {
		cudaFree(s.x); cudaFree(s.ud); cudaFree(s.ua); cudaFree(s.relax); cudaFree(s.rc);
		s.x = s.ud = s.ua = s.relax = nullptr;
		s.rc = nullptr;
		s.cap = 0;
		CUDA_TRY(cudaMalloc(&s.x, sizeof(double) * n * e->nx));
		CUDA_TRY(cudaMalloc(&s.ud, sizeof(double) * n * e->nv)); // uDes[n][nu] or, for the (H, c) overloads, c[n][nv]
		CUDA_TRY(cudaMalloc(&s.ua, sizeof(double) * n * e->nu));
		CUDA_TRY(cudaMalloc(&s.relax, sizeof(double) * n * e->n_relax));
		CUDA_TRY(cudaMalloc(&s.rc, sizeof(int32_t) * n));
		s.cap = n;
	}
	if (want_diag && n > s.cap_diag) {
		cudaFree(s.diag);
		s.diag = nullptr;
		s.cap_diag = 0;
		CUDA_TRY(cudaMalloc(&s.diag, sizeof(double) * n * e->n_diag));
		s.cap_diag = n;
	}
	return ASIF_OK;
}

// ---- soft saturation constants, src/asif_implicit_tb.cpp:768-784 (host libm, same operations)
int make_softsat(double r, const double *lb, const double *ub, int nu, SoftSat &s)
{
	const double alpha = M_PI / 8;
	const double beta = M_PI / 4;
	const double bevelL = r * tan(alpha);
	s.r = r;
	s.r2 = r * r;
	s.bevelStart = 1 - cos(beta) * bevelL;
	s.bevelStop = 1 + bevelL;
	s.bevelYc = 1 - r;
	for (int i = 0; i < MAX_NU; i++) {
		s.range[i] = 1.0;
		s.middle[i] = 0.0;
		s.uc_scale_exact[i] = 0.0;
		s.uc_scale[i] = 2.0;
	}
	int mode = SAT_IDENTITY;
	for (int i = 0; i < nu; i++) {
		s.range[i] = ub[i] - lb[i];
		s.middle[i] = (ub[i] + lb[i]) / 2;
		s.uc_scale[i] = 2.0 / s.range[i];
		int ex;
		const bool pow2 = frexp(s.range[i], &ex) == 0.5;
		if (pow2) s.uc_scale_exact[i] = 2.0 / s.range[i]; // exact: every operation of 2*(u-mid)/range is then exact
		const int m = (pow2 && s.middle[i] == 0.0 && s.range[i] == 2.0) ? SAT_IDENTITY : (pow2 ? SAT_POW2 : SAT_GENERAL);
		if (m < mode) mode = m;
	}
	return mode;
}

template <class M>
int launch_explicit(asif_engine *e, int64_t n, const double *x, const double *ud, double *ua, double *relax,
                    int32_t *rc, double *diag, cudaStream_t st)
{
	const unsigned blocks = (unsigned)((n + EXPL_THREADS - 1) / EXPL_THREADS);
	if (diag)
		explicit_filter_kernel<M, true><<<blocks, EXPL_THREADS, 0, st>>>(e->ex, n, x, ud, ua, relax, rc, diag, e->ctr);
	else if (e->ex.npSSmax >= M::NPSS) // every safety function keeps its row: the instantiation without the rank computation
		explicit_filter_kernel<M, false, false><<<blocks, EXPL_THREADS, 0, st>>>(e->ex, n, x, ud, ua, relax, rc, nullptr, e->ctr);
	else
		explicit_filter_kernel<M, false><<<blocks, EXPL_THREADS, 0, st>>>(e->ex, n, x, ud, ua, relax, rc, nullptr, e->ctr);
	CUDA_TRY(cudaGetLastError());
	return ASIF_OK;
}

int launch_robust(asif_engine *e, int64_t n, const double *x, const double *ud, double *ua, double *relax, int32_t *rc,
                  double *diag, cudaStream_t st)
{
	const size_t smem = sizeof(double) * 4 * (size_t)e->rb.n_halfplanes;
	const unsigned blocks = (unsigned)((n + ROB_THREADS - 1) / ROB_THREADS);
	int r = set_smem(robust_ip_filter_kernel<true>, smem); // tables beyond 1536 half-planes need the opt-in shared-memory size
	if (!r) r = set_smem(robust_ip_filter_kernel<false>, smem);
	if (r) return r;
	if (diag)
		robust_ip_filter_kernel<true><<<blocks, ROB_THREADS, smem, st>>>(e->rb, n, x, ud, ua, relax, rc, diag, e->ctr);
	else
		robust_ip_filter_kernel<false><<<blocks, ROB_THREADS, smem, st>>>(e->rb, n, x, ud, ua, relax, rc, nullptr, e->ctr);
	CUDA_TRY(cudaGetLastError());
	return ASIF_OK;
}

int launch_realizable(asif_engine *e, int64_t n, const double *x, const double *ud, double *ua, double *relax, int32_t *rc,
                      double *diag, cudaStream_t st)
{
	const unsigned blocks = (unsigned)((n + RZ_THREADS - 1) / RZ_THREADS);
	if (diag)
		realizable_ip_filter_kernel<true><<<blocks, RZ_THREADS, e->rz_smem, st>>>(e->rz, n, x, ud, ua, relax, rc, diag, e->ctr);
	else
		realizable_ip_filter_kernel<false><<<blocks, RZ_THREADS, e->rz_smem, st>>>(e->rz, n, x, ud, ua, relax, rc, nullptr, e->ctr);
	CUDA_TRY(cudaGetLastError());
	return ASIF_OK;
}

int launch_filter(asif_engine *e, int64_t n, const double *x, const double *ud, double *ua, double *relax, int32_t *rc,
                  double *diag, cudaStream_t st)
{
	if (n <= 0) return ASIF_OK;
	switch (e->cfg.filter) {
	case ASIF_FILTER_EXPLICIT:
		return launch_explicit<DoubleIntegratorExplicit>(e, n, x, ud, ua, relax, rc, diag, st);
	case ASIF_FILTER_IMPLICIT_TB:
		switch (e->cfg.model) {
		case ASIF_MODEL_DOUBLE_INTEGRATOR_TB:
			if (e->cfg.npBTSS != 4) return launch_tb<DoubleIntegratorTB, TB_NPBTSS_RUNTIME>(e, n, x, ud, ua, relax, rc, diag, st);
			return launch_tb<DoubleIntegratorTB, 4>(e, n, x, ud, ua, relax, rc, diag, st);
		case ASIF_MODEL_SEGWAY: return launch_tb_segway(e, false, n, x, ud, ua, relax, rc, diag, st);
		case ASIF_MODEL_SEGWAY_SHIPPED: return launch_tb_segway(e, true, n, x, ud, ua, relax, rc, diag, st);
		}
		break;
	case ASIF_FILTER_IMPLICIT:
		if (e->cfg.model == ASIF_MODEL_INVERTED_PENDULUM)
			return launch_implicit_ip(e, n, x, ud, ua, relax, rc, diag, st);
		break;
	case ASIF_FILTER_IMPLICIT_RB:
		if (e->cfg.model == ASIF_MODEL_INVERTED_PENDULUM) return launch_implicit_rb_ip(e, n, x, ud, ua, relax, rc, diag, st);
		if (e->cfg.model == ASIF_MODEL_DOUBLE_INTEGRATOR_TB) { // fused-gradient constructor; bit-exact unit
			using M = DoubleIntegratorTB;
			if (e->cfg.npBTSS != 10) return launch_implicit_t<M, IMP_NPBTSS_RUNTIME, true, SAT_GENERAL>(e, n, x, ud, ua, relax, rc, diag, st);
			return launch_implicit_t<M, 10, true, SAT_GENERAL>(e, n, x, ud, ua, relax, rc, diag, st);
		}
		break;
	case ASIF_FILTER_ROBUST:
		if (e->cfg.model == ASIF_MODEL_INVERTED_PENDULUM_TABLE) return launch_robust(e, n, x, ud, ua, relax, rc, diag, st);
		break;
	case ASIF_FILTER_REALIZABLE:
		if (e->cfg.model == ASIF_MODEL_INVERTED_PENDULUM_KERNEL) return launch_realizable(e, n, x, ud, ua, relax, rc, diag, st);
		break;
	}
	return fail(ASIF_ERR_UNSUPPORTED, "no kernel for filter %d / model %d", e->cfg.filter, e->cfg.model);
}

} // namespace

namespace asifb {
int server_suspend(asif_engine *e) { return server_stop(e); }
int server_resume(asif_engine *e) { return server_start(e); }
int launch_filter_any(asif_engine *e, int64_t n, const double *x, const double *ud, double *ua, double *relax, int32_t *rc,
                      double *diag, cudaStream_t st)
{
	return launch_filter(e, n, x, ud, ua, relax, rc, diag, st);
}
} // namespace asifb

namespace {

int launch_rollout(asif_engine *e, int64_t n, int32_t steps, double dt, double *x, const double *ud, double *ua,
                   int32_t *rc, cudaStream_t st)
{
	if (n <= 0) return ASIF_OK;
	if (e->cfg.filter == ASIF_FILTER_IMPLICIT_TB) {
		switch (e->cfg.model) {
		case ASIF_MODEL_DOUBLE_INTEGRATOR_TB:
			if (e->cfg.npBTSS != 4) return launch_tb_rollout<DoubleIntegratorTB, TB_NPBTSS_RUNTIME>(e, n, steps, dt, x, ud, ua, rc, st);
			return launch_tb_rollout<DoubleIntegratorTB, 4>(e, n, steps, dt, x, ud, ua, rc, st);
		case ASIF_MODEL_SEGWAY: return launch_tb_rollout_segway(e, false, n, steps, dt, x, ud, ua, rc, st);
		case ASIF_MODEL_SEGWAY_SHIPPED: return launch_tb_rollout_segway(e, true, n, steps, dt, x, ud, ua, rc, st);
		}
	}
	return fail(ASIF_ERR_UNSUPPORTED, "rollout is implemented for the implicit-TB filter (filter %d / model %d given)",
	            e->cfg.filter, e->cfg.model);
}

void model_dims(int model, int &nx, int &nu, int &npSS)
{
	switch (model) {
	case ASIF_MODEL_SEGWAY:
	case ASIF_MODEL_SEGWAY_SHIPPED: nx = 4; nu = 1; npSS = 4; break;
	default: nx = 2; nu = 1; npSS = 4; break;
	}
}

} // namespace

// =================================================================================================
extern "C" {

int32_t asif_b200_abi_version(void) { return ASIF_B200_ABI_VERSION; }

const char *asif_last_error(void) { return g_err; }

int32_t asif_device_count(void)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) return fail(ASIF_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
	return n;
}

int32_t asif_engine_config_init(asif_engine_config *cfg, int32_t filter, int32_t model)
{
	if (!cfg) return fail(ASIF_ERR_INVALID_ARGUMENT, "cfg is NULL");
	memset(cfg, 0, sizeof(*cfg));
	cfg->struct_size = sizeof(*cfg);
	cfg->filter = filter;
	cfg->model = model;
	cfg->device = 0;
	cfg->npSSmax = -1;
	cfg->inf = 1e20;
	cfg->relaxCost = 50.0;
	switch (filter) {
	case ASIF_FILTER_EXPLICIT: /* include/asif.h:11-17 */
		cfg->relaxLb = 5.0;
		cfg->satSharpness = 5.0;
		break;
	case ASIF_FILTER_IMPLICIT_TB: /* include/asif_implicit_tb.h:19-33 */
		cfg->relaxLb = 5.0;
		cfg->relaxTTS = 5.0;
		cfg->relaxMinOrtho = 5.0;
		cfg->backTrajHorizon = 1.0;
		cfg->backTrajExtend = 0.05;
		cfg->backTrajDt = 0.01;
		cfg->backTrajMinOrtho = 0.01;
		cfg->satSharpness = 0.1;
		cfg->npBTSS = 4;
		break;
	case ASIF_FILTER_IMPLICIT: /* include/asif_implicit.h:20-34 */
		cfg->relaxLb = 5.0;
		cfg->relaxReachLb = 5.0;
		cfg->backTrajHorizon = 1.0;
		cfg->backTrajDt = 0.01;
		cfg->satSharpness = 0.1;
		cfg->npBTSS = 10;
		break;
	case ASIF_FILTER_IMPLICIT_RB: /* include/asif_implicit_robust.h:22-38 */
		cfg->relaxLb = 5.0;
		cfg->relaxReachLb = 5.0;
		cfg->backTrajHorizon = 1.0;
		cfg->backTrajDt = 0.01;
		cfg->backContDt = 0.01;
		cfg->satSharpness = 0.1;
		cfg->npBTSS = 10;
		break;
	case ASIF_FILTER_ROBUST: /* include/asif_robust.h:14-19 */
		cfg->relaxLb = 5.0;
		break;
	case ASIF_FILTER_REALIZABLE: /* include/asif_realizable.h:14-20 */
		cfg->relaxDes = 5.0;
		cfg->relaxOffset = 5.0;
		break;
	default:
		return fail(ASIF_ERR_INVALID_ARGUMENT, "unknown filter %d", filter);
	}
	switch (model) {
	case ASIF_MODEL_DOUBLE_INTEGRATOR:
	case ASIF_MODEL_DOUBLE_INTEGRATOR_TB:
		cfg->lb[0] = -1.0; cfg->ub[0] = 1.0; /* examples/DoubleIntegrator*.cpp: lb/ub */
		break;
	case ASIF_MODEL_INVERTED_PENDULUM_KERNEL:
		cfg->lb[0] = -1.5; cfg->ub[0] = 1.5; /* examples/InvertedPendulum_RealizableSampled.cpp:19-20 */
		cfg->dynParam[0] = 0.9; cfg->dynParam[1] = 1.1; /* :27-28 */
		cfg->uncertaintyBounds[0] = 0.032; cfg->uncertaintyBounds[1] = 0.027; /* :23 */
		cfg->npSSmax = 0; /* the default of the example's constructor call (:239) */
		break;
	case ASIF_MODEL_INVERTED_PENDULUM:
	case ASIF_MODEL_INVERTED_PENDULUM_TABLE:
		cfg->lb[0] = -1.5; cfg->ub[0] = 1.5; /* examples/InvertedPendulum_Implicit.cpp:19-20 */
		cfg->dynParam[0] = 0.8; cfg->dynParam[1] = 1.2; /* input-gain interval, examples/InvertedPendulum_Robust.cpp:35-38 */
		break;
	case ASIF_MODEL_SEGWAY:
	case ASIF_MODEL_SEGWAY_SHIPPED:
		cfg->lb[0] = -20.0; cfg->ub[0] = 20.0; /* examples/segway_implicit_tb.cpp:18-19 */
		break;
	default:
		return fail(ASIF_ERR_INVALID_ARGUMENT, "unknown model %d", model);
	}
	return ASIF_OK;
}

int32_t asif_engine_create(const asif_engine_config *cfg, asif_engine **out)
{
	if (!cfg || !out) return fail(ASIF_ERR_INVALID_ARGUMENT, "cfg/out is NULL");
	if (cfg->struct_size != sizeof(asif_engine_config))
		return fail(ASIF_ERR_INVALID_ARGUMENT, "asif_engine_config size mismatch (%u vs %zu): use asif_engine_config_init",
		            cfg->struct_size, sizeof(asif_engine_config));
	*out = nullptr;
	int ndev = 0;
	cudaError_t ce = cudaGetDeviceCount(&ndev);
	if (ce != cudaSuccess || ndev <= 0)
		return fail(ASIF_ERR_NO_DEVICE, "no CUDA device (%s); this engine has no CPU fallback",
		            ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
	if (cfg->device < 0 || cfg->device >= ndev) return fail(ASIF_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)", cfg->device, ndev);
	CUDA_TRY(cudaSetDevice(cfg->device));

	asif_engine *e = new (std::nothrow) asif_engine();
	if (!e) return fail(ASIF_ERR_INVALID_ARGUMENT, "out of host memory");
	e->cfg = *cfg;
	if (cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, cfg->device) != cudaSuccess || e->num_sms <= 0) e->num_sms = 148;
	int npSS;
	model_dims(cfg->model, e->nx, e->nu, npSS);
	e->nv = e->nu + 1;
	e->n_relax = 1;
	const int nu = e->nu;
	for (int i = 0; i < nu; i++)
		if (!(cfg->lb[i] < cfg->ub[i])) {
			delete e;
			return fail(ASIF_ERR_INVALID_ARGUMENT, "lb[%d] must be < ub[%d]", i, i);
		}
	if (!(cfg->relaxCost > 0)) {
		delete e;
		return fail(ASIF_ERR_INVALID_ARGUMENT, "relaxCost must be > 0");
	}
	switch (cfg->filter) {
	case ASIF_FILTER_EXPLICIT: {
		if (cfg->model != ASIF_MODEL_DOUBLE_INTEGRATOR) {
			delete e;
			return fail(ASIF_ERR_UNSUPPORTED, "explicit filter: model %d not compiled in", cfg->model);
		}
		e->nc = (cfg->npSSmax > 0 && cfg->npSSmax < npSS) ? cfg->npSSmax : npSS; // npSSmax_ = min(npSSmax, npSS), nc_ = npSSmax_
		e->n_diag = e->nc * e->nv + e->nc;
		ExplicitParams &p = e->ex;
		memset(&p, 0, sizeof(p));
		p.npSSmax = e->nc;
		for (int i = 0; i < nu; i++) {
			p.lb[i] = cfg->lb[i];
			p.ub[i] = cfg->ub[i];
			p.gi[i] = 1.0 / 2.0; // H = diag(I_nu, relaxCost), src/asif.cpp:71-83
			p.gih[i] = sqrt(p.gi[i]);
		}
		p.relaxLb = cfg->relaxLb;
		p.relaxCost = cfg->relaxCost;
		p.gi[nu] = 1.0 / (2.0 * cfg->relaxCost);
		p.gih[nu] = sqrt(p.gi[nu]);
		break;
	}
	case ASIF_FILTER_IMPLICIT_TB: {
		if (cfg->model != ASIF_MODEL_DOUBLE_INTEGRATOR_TB && cfg->model != ASIF_MODEL_SEGWAY &&
		    cfg->model != ASIF_MODEL_SEGWAY_SHIPPED) {
			delete e;
			return fail(ASIF_ERR_UNSUPPORTED, "implicit-TB filter: model %d not compiled in", cfg->model);
		}
		if (cfg->npBTSS < 1 || cfg->npBTSS > np_capacity(TB_NPBTSS_RUNTIME)) {
			delete e;
			return fail(ASIF_ERR_UNSUPPORTED, "implicit-TB filter: npBTSS = %d outside 1..%d", cfg->npBTSS, np_capacity(TB_NPBTSS_RUNTIME));
		}
		if (!(cfg->backTrajDt > 0) || !(cfg->backTrajHorizon > 0)) {
			delete e;
			return fail(ASIF_ERR_INVALID_ARGUMENT, "backTrajDt and backTrajHorizon must be > 0");
		}
		e->nc = cfg->npBTSS * npSS + 2;
		e->n_diag = 4 + cfg->npBTSS + e->nc * e->nv + e->nc;
		TbParams &p = e->tb;
		memset(&p, 0, sizeof(p));
		for (int i = 0; i < nu; i++) {
			p.lb[i] = cfg->lb[i];
			p.ub[i] = cfg->ub[i];
			p.gi[i] = 1.0 / 2.0;
			p.gih[i] = sqrt(p.gi[i]);
		}
		p.relaxCost = cfg->relaxCost;
		p.relaxSafeLb = cfg->relaxLb;
		p.relaxTTS = cfg->relaxTTS;
		p.relaxMinOrtho = cfg->relaxMinOrtho;
		p.backTrajHorizon = cfg->backTrajHorizon;
		p.backTrajDt = cfg->backTrajDt;
		p.backTrajMinOrtho = cfg->backTrajMinOrtho;
		p.inf = cfg->inf;
		// src/asif_implicit_tb.cpp:177-182
		int64_t npBT = (int64_t)round(cfg->backTrajHorizon * (1.0 + cfg->backTrajExtend) / cfg->backTrajDt) + 1;
		if (npBT < cfg->npBTSS) {
			npBT = cfg->npBTSS;
			p.backTrajDt = cfg->backTrajHorizon * (1.0 + cfg->backTrajExtend) / (double)(npBT - 1);
		}
		if (npBT > (1 << 24)) {
			delete e;
			return fail(ASIF_ERR_INVALID_ARGUMENT, "backup trajectory of %lld points is not sensible", (long long)npBT);
		}
		p.npBT = (int32_t)npBT;
		p.npBTSS = cfg->npBTSS;
		// updateOptions() clamps satSharpness to [0.01, 2] (src/asif_implicit_tb.cpp:391-400); initialize() does not.
		p.sat_mode = make_softsat(cfg->satSharpness, cfg->lb, cfg->ub, nu, p.sat);
		p.gi[nu] = 1.0 / (2.0 * cfg->relaxCost);
		p.gih[nu] = sqrt(p.gi[nu]);
		{ // t_i exactly as the reference accumulates it
			double *tt = new (std::nothrow) double[npBT];
			cudaError_t te = tt ? cudaMalloc(&e->d_ttable, sizeof(double) * npBT) : cudaErrorMemoryAllocation;
			if (te == cudaSuccess) {
				tt[0] = 0.0;
				for (int64_t i = 1; i < npBT; i++) tt[i] = tt[i - 1] + p.backTrajDt;
				te = cudaMemcpy(e->d_ttable, tt, sizeof(double) * npBT, cudaMemcpyHostToDevice);
			}
			delete[] tt;
			if (te != cudaSuccess) {
				cudaFree(e->d_ttable);
				delete e;
				return fail(ASIF_ERR_CUDA, "time table upload failed: %s", cudaGetErrorString(te));
			}
			p.t_of_index = e->d_ttable;
		}
		break;
	}
	case ASIF_FILTER_IMPLICIT_RB:
		if (cfg->model != ASIF_MODEL_INVERTED_PENDULUM && cfg->model != ASIF_MODEL_DOUBLE_INTEGRATOR_TB) {
			delete e;
			return fail(ASIF_ERR_UNSUPPORTED, "implicit-RB filter: model %d not compiled in", cfg->model);
		}
		// the hold is re-armed on the first Euler step of every trajectory only while backContDt - 0.0001 <= 1 + dt
		// (src/asif_implicit_robust.cpp:895-903); beyond that the reference would carry the held input from one
		// filter() call into the next, which a batch of independent states cannot reproduce
		if (!(cfg->backContDt > 0) || !(cfg->backContDt < 1.0)) {
			delete e;
			return fail(ASIF_ERR_INVALID_ARGUMENT, "implicit-RB filter: backContDt must be in (0, 1)");
		}
		for (int i = 0; i < e->nx; i++)
			if (!(cfg->x_unc[i] >= 0)) {
				delete e;
				return fail(ASIF_ERR_INVALID_ARGUMENT, "implicit-RB filter: x_unc[%d] must be >= 0", i);
			}
		/* fall through: the rest is ASIFimplicit's initialize (src/asif_implicit_robust.cpp:264-344 == src/asif_implicit.cpp:194-266) */
	case ASIF_FILTER_IMPLICIT: {
		if (cfg->filter == ASIF_FILTER_IMPLICIT && cfg->model != ASIF_MODEL_INVERTED_PENDULUM) {
			delete e;
			return fail(ASIF_ERR_UNSUPPORTED, "implicit filter: model %d not compiled in", cfg->model);
		}
		if (cfg->npBTSS < 1 || cfg->npBTSS > np_capacity(IMP_NPBTSS_RUNTIME)) {
			delete e;
			return fail(ASIF_ERR_UNSUPPORTED, "implicit filter: npBTSS = %d outside 1..%d", cfg->npBTSS, np_capacity(IMP_NPBTSS_RUNTIME));
		}
		if (!(cfg->backTrajDt > 0) || !(cfg->backTrajHorizon > 0)) {
			delete e;
			return fail(ASIF_ERR_INVALID_ARGUMENT, "backTrajDt and backTrajHorizon must be > 0");
		}
		const int npBS = 1;
		e->n_relax = 2;
		e->nv = nu + 2;
		e->nc = cfg->npBTSS * npSS + npBS;
		e->n_diag = 2 + cfg->npBTSS + e->nc * e->nv + e->nc;
		ImplicitParams &p = e->im;
		memset(&p, 0, sizeof(p));
		for (int i = 0; i < nu; i++) {
			p.lb[i] = cfg->lb[i];
			p.ub[i] = cfg->ub[i];
			p.gi[i] = 1.0 / 2.0;
			p.gih[i] = sqrt(p.gi[i]);
		}
		p.relaxCost = cfg->relaxCost;
		p.relaxSafeLb = cfg->relaxLb;
		p.relaxReachLb = cfg->relaxReachLb;
		p.backTrajDt = cfg->backTrajDt;
		p.inf = cfg->inf;
		// src/asif_implicit.cpp:211-216 (this class has no backTrajExtend)
		int64_t npBT = (int64_t)round(cfg->backTrajHorizon / cfg->backTrajDt) + 1;
		if (npBT < cfg->npBTSS) {
			npBT = cfg->npBTSS;
			p.backTrajDt = cfg->backTrajHorizon / (double)(npBT - 1);
		}
		if (npBT > (1 << 24)) {
			delete e;
			return fail(ASIF_ERR_INVALID_ARGUMENT, "backup trajectory of %lld points is not sensible", (long long)npBT);
		}
		p.npBT = (int32_t)npBT;
		p.npBTSS = cfg->npBTSS;
		p.backContDt = cfg->backContDt;
		for (int i = 0; i < 4; i++) p.x_unc[i] = (i < e->nx) ? cfg->x_unc[i] : 0.0;
		p.sat_mode = make_softsat(cfg->satSharpness, cfg->lb, cfg->ub, nu, p.sat);
		for (int i = nu; i < nu + 2; i++) {
			p.gi[i] = 1.0 / (2.0 * cfg->relaxCost);
			p.gih[i] = sqrt(p.gi[i]);
		}
		break;
	}
	case ASIF_FILTER_ROBUST: {
		if (cfg->model != ASIF_MODEL_INVERTED_PENDULUM_TABLE) {
			delete e;
			return fail(ASIF_ERR_UNSUPPORTED, "robust filter: model %d not compiled in", cfg->model);
		}
		if (!cfg->halfplanes || cfg->n_halfplanes < 1 || cfg->n_halfplanes > 4096) {
			delete e;
			return fail(ASIF_ERR_INVALID_ARGUMENT, "robust filter needs a half-plane table (1..4096 rows)");
		}
		if (!(cfg->dynParam[0] <= cfg->dynParam[1])) {
			delete e;
			return fail(ASIF_ERR_INVALID_ARGUMENT, "dynParam[0..1] = [pMin, pMax] must be ordered");
		}
		const int K = cfg->n_halfplanes;
		e->nc = 2 * K; // rows of the reduced problem
		e->n_diag = 5 * K;
		RobustParams &p = e->rb;
		memset(&p, 0, sizeof(p));
		p.lb[0] = cfg->lb[0];
		p.ub[0] = cfg->ub[0];
		p.relaxLb = cfg->relaxLb;
		p.relaxCost = cfg->relaxCost;
		p.inf = cfg->inf;
		p.gc = (cfg->dynParam[0] + cfg->dynParam[1]) / 2; // AAF(interval): centre, radius (aa_aafcommon.cpp:81-101)
		p.gr = (cfg->dynParam[1] - cfg->dynParam[0]) / 2;
		p.gi[0] = 0.5;
		p.gih[0] = sqrt(0.5);
		p.gi[1] = 1.0 / (2.0 * cfg->relaxCost);
		p.gih[1] = sqrt(p.gi[1]);
		p.n_halfplanes = K;
		cudaError_t te = cudaMalloc(&e->d_table, sizeof(double) * 2 * K);
		if (te == cudaSuccess) te = cudaMemcpy(e->d_table, cfg->halfplanes, sizeof(double) * 2 * K, cudaMemcpyHostToDevice);
		if (te != cudaSuccess) {
			cudaFree(e->d_table);
			delete e;
			return fail(ASIF_ERR_CUDA, "table upload failed: %s", cudaGetErrorString(te));
		}
		p.table = e->d_table;
		e->cfg.halfplanes = nullptr; // borrowed pointer is not kept
		break;
	}
	case ASIF_FILTER_REALIZABLE: {
		if (cfg->model != ASIF_MODEL_INVERTED_PENDULUM_KERNEL) {
			delete e;
			return fail(ASIF_ERR_UNSUPPORTED, "realizable filter: model %d not compiled in", cfg->model);
		}
		const int nV = cfg->n_vertices, nF = cfg->n_facets, mC = cfg->max_critical_facets, mA = cfg->max_active_constraints;
		const int nb = cfg->npSSmax < 0 ? 0 : (cfg->npSSmax > nF ? nF : cfg->npSSmax); // src/asif_realizable.cpp:20
		if (!cfg->kernel_vertices || !cfg->facet_normals || !cfg->facet_vertices || !cfg->facet_active || !cfg->facet_lie ||
		    nV < 2 || nF < 1 || nV > 4096 || nF > 4096) {
			delete e;
			return fail(ASIF_ERR_INVALID_ARGUMENT, "realizable filter needs the polytope kernel tables (vertices, normals, "
			                                       "facet_vertices, facet_active, facet_lie)");
		}
		if (mC < 1 || mC > RZ_MAX_CRIT || mA < 1 || mA > RZ_MAX_ACT || nb > RZ_MAX_BAR) {
			delete e;
			return fail(ASIF_ERR_UNSUPPORTED, "realizable filter: maxCriticalFacets <= %d, maxActiveConstraints <= %d, npSSmax <= %d",
			            RZ_MAX_CRIT, RZ_MAX_ACT, RZ_MAX_BAR);
		}
		for (int i = 0; i < 2 * nF; i++)
			if (cfg->facet_vertices[i] < 0 || cfg->facet_vertices[i] >= nV) {
				delete e;
				return fail(ASIF_ERR_INVALID_ARGUMENT, "facet_vertices[%d] out of range", i);
			}
		for (int i = 0; i < mA * nF; i++)
			if (cfg->facet_active[i] >= nF) {
				delete e;
				return fail(ASIF_ERR_INVALID_ARGUMENT, "facet_active[%d] out of range", i);
			}
		e->n_relax = 2;
		e->nv = 2;
		e->nc = 2 * mC * mA + nb;
		e->n_diag = 1 + mC + nb + 4 * mC * mA + 2 * nb;
		RealizableParams &p = e->rz;
		memset(&p, 0, sizeof(p));
		p.lb = cfg->lb[0];
		p.ub = cfg->ub[0];
		p.relaxDes = cfg->relaxDes;
		p.relaxOffset = cfg->relaxOffset;
		p.relaxCost = cfg->relaxCost;
		p.inf = cfg->inf;
		p.unc[0] = cfg->uncertaintyBounds[0];
		p.unc[1] = cfg->uncertaintyBounds[1];
		{ // mid of AAF(interval(pMin, pMax)).convert()  (aa_aafcommon.cpp:81-101, 217-227)
			const double gc = (cfg->dynParam[1] + cfg->dynParam[0]) / 2, gr = (cfg->dynParam[1] - cfg->dynParam[0]) / 2;
			p.gmid = ((gc - gr) + (gc + gr)) / 2;
		}
		p.gi[0] = 0.5;
		p.gih[0] = sqrt(0.5);
		p.gi[1] = 1.0 / (2.0 * (nb > 0 ? cfg->relaxCost : 1.0)); // H(eps) = relaxCost only when npSSmax > 0 (:188-189)
		p.gih[1] = sqrt(p.gi[1]);
		p.n_vertices = nV;
		p.n_facets = nF;
		p.max_crit = mC;
		p.max_act = mA;
		p.npSSmax = nb;
		const size_t bV = sizeof(double) * 2 * nV, bN = sizeof(double) * 2 * nF, bL = sizeof(double) * 4 * nF * mA;
		const size_t bFV = sizeof(int32_t) * 2 * nF, bFA = sizeof(int32_t) * nF * mA;
		const size_t bTable = bV + bN + bL + bFV + bFA;
		// + the per-thread slot lists of the critical-facet rows (realizable_kernel.cuh)
		e->rz_smem = bTable + sizeof(uint16_t) * RZ_MAX_CRIT * RZ_MAX_ACT * RZ_THREADS + sizeof(double) * (4 * nF + 1); // + inflated boxes (8-byte aligned)
		if (e->rz_smem > 96 * 1024 || (size_t)nF * mA > 65535) {
			delete e;
			return fail(ASIF_ERR_UNSUPPORTED, "polytope kernel too large for the shared-memory staging (%zu B)", e->rz_smem);
		}
		cudaError_t te = cudaMalloc(&e->d_kernel, bTable);
		char *base = (char *)e->d_kernel;
		if (te == cudaSuccess) te = cudaMemcpy(base, cfg->kernel_vertices, bV, cudaMemcpyHostToDevice);
		if (te == cudaSuccess) te = cudaMemcpy(base + bV, cfg->facet_normals, bN, cudaMemcpyHostToDevice);
		if (te == cudaSuccess) te = cudaMemcpy(base + bV + bN, cfg->facet_lie, bL, cudaMemcpyHostToDevice);
		if (te == cudaSuccess) te = cudaMemcpy(base + bV + bN + bL, cfg->facet_vertices, bFV, cudaMemcpyHostToDevice);
		if (te == cudaSuccess) te = cudaMemcpy(base + bV + bN + bL + bFV, cfg->facet_active, bFA, cudaMemcpyHostToDevice);
		if (te == cudaSuccess && e->rz_smem > 48 * 1024) {
			te = cudaFuncSetAttribute(realizable_ip_filter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->rz_smem);
			if (te == cudaSuccess)
				te = cudaFuncSetAttribute(realizable_ip_filter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->rz_smem);
		}
		if (te != cudaSuccess) {
			cudaFree(e->d_kernel);
			delete e;
			return fail(ASIF_ERR_CUDA, "kernel-table upload failed: %s", cudaGetErrorString(te));
		}
		p.vertices = (const double *)base;
		p.normals = (const double *)(base + bV);
		p.facet_lie = (const double *)(base + bV + bN);
		p.facet_vertices = (const int32_t *)(base + bV + bN + bL);
		p.facet_active = (const int32_t *)(base + bV + bN + bL + bFV);
		e->cfg.kernel_vertices = e->cfg.facet_normals = e->cfg.facet_lie = nullptr; // borrowed pointers are not kept
		e->cfg.facet_vertices = e->cfg.facet_active = nullptr;
		break;
	}
	default:
		delete e;
		return fail(ASIF_ERR_UNSUPPORTED, "filter %d not implemented yet", cfg->filter);
	}
	cudaError_t err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
	if (err == cudaSuccess) err = cudaMalloc(&e->d_counters, N_COUNTERS * sizeof(unsigned long long));
	if (err == cudaSuccess) err = cudaMemset(e->d_counters, 0, N_COUNTERS * sizeof(unsigned long long));
	if (err != cudaSuccess) {
		asif_engine_destroy(e);
		return fail(ASIF_ERR_CUDA, "engine allocation failed: %s", cudaGetErrorString(err));
	}
	e->ctr = e->d_counters;
	*out = e;
	return ASIF_OK;
}

int32_t asif_engine_destroy(asif_engine *e)
{
	if (!e) return ASIF_OK;
	cudaSetDevice(e->cfg.device);
	server_stop(e);
	if (e->srv_stream) cudaStreamDestroy(e->srv_stream);
	if (e->srv_mailbox) cudaFreeHost(e->srv_mailbox);
	for (Slot &s : e->slot) {
		if (s.stream) {
			cudaStreamSynchronize(s.stream);
			cudaStreamDestroy(s.stream);
		}
		cudaFree(s.x); cudaFree(s.ud); cudaFree(s.ua); cudaFree(s.relax); cudaFree(s.rc); cudaFree(s.diag);
		if (s.h) cudaFreeHost(s.h);
	}
	if (e->stream) {
		cudaStreamSynchronize(e->stream);
		cudaStreamDestroy(e->stream);
	}
	cudaFree(e->d_counters);
	for (auto &b : e->snapbuf) {
		if (b.p) cudaFree(b.p);
		if (b.ev) cudaEventDestroy(b.ev);
	}
	cudaFree(e->d_table);
	cudaFree(e->d_learn);
	cudaFree(e->d_kernel);
	cudaFree(e->d_ttable);
	if (e->small_h) cudaFreeHost(e->small_h);
	delete e->copier;
	delete e;
	return ASIF_OK;
}

int32_t asif_engine_dims(const asif_engine *e, int32_t dims[6])
{
	if (!e || !dims) return fail(ASIF_ERR_INVALID_ARGUMENT, "NULL argument");
	dims[0] = e->nx; dims[1] = e->nu; dims[2] = e->n_relax; dims[3] = e->nc; dims[4] = e->nv; dims[5] = e->n_diag;
	return ASIF_OK;
}

} // extern "C"

namespace {
// ---- latency server (latency_server.cuh) ---------------------------------------------------------------------------
constexpr size_t SRV_MAILBOX_BYTES = 16 << 10;

int server_stop(asif_engine *e)
{
	if (!e->srv_on) return ASIF_OK;
	volatile unsigned long long *hdr = reinterpret_cast<volatile unsigned long long *>(e->srv_mailbox);
	__atomic_store_n(const_cast<unsigned long long *>(hdr), SRV_EXIT, __ATOMIC_RELEASE);
	e->srv_on = false;
	// bounded wait: a resident kernel that no longer polls its mailbox cannot be killed, but it must not hang the caller
	const auto t0 = std::chrono::steady_clock::now();
	for (;;) {
		const cudaError_t q = cudaStreamQuery(e->srv_stream);
		if (q == cudaSuccess) return ASIF_OK;
		if (q != cudaErrorNotReady) return fail(ASIF_ERR_CUDA, "latency server ended with: %s", cudaGetErrorString(q));
		if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 5.0)
			return fail(ASIF_ERR_CUDA, "latency server did not stop within 5 s");
	}
}

int server_start(asif_engine *e)
{
	if (e->srv_on) return ASIF_OK;
	const bool tb_di = e->cfg.filter == ASIF_FILTER_IMPLICIT_TB && e->cfg.model == ASIF_MODEL_DOUBLE_INTEGRATOR_TB && e->cfg.npBTSS == 4;
	const bool expl = e->cfg.filter == ASIF_FILTER_EXPLICIT && e->cfg.model == ASIF_MODEL_DOUBLE_INTEGRATOR;
	if (!tb_di && !expl)
		return fail(ASIF_ERR_UNSUPPORTED, "the latency server is built for ASIF / DoubleIntegrator and ASIFimplicitTB / DoubleIntegrator "
		                                  "(npBTSS 4); other classes keep the launch path");
	CUDA_TRY(cudaSetDevice(e->cfg.device));
	if (!e->srv_mailbox) {
		CUDA_TRY(cudaHostAlloc((void **)&e->srv_mailbox, SRV_MAILBOX_BYTES, cudaHostAllocMapped));
		CUDA_TRY(cudaHostGetDevicePointer((void **)&e->srv_mailbox_dev, e->srv_mailbox, 0));
	}
	if (!e->srv_stream) CUDA_TRY(cudaStreamCreateWithFlags(&e->srv_stream, cudaStreamNonBlocking));
	memset(e->srv_mailbox, 0, SRV_MAILBOX_BYTES);
	e->srv_seq = 0;
	if (expl) {
		ExplicitParams p = e->ex;
		p.custom_cost = 0;
		p.lfh = p.lgh = nullptr;
		explicit_server_kernel<DoubleIntegratorExplicit><<<1, 32, 0, e->srv_stream>>>(p, e->srv_mailbox_dev);
	} else {
		using M = DoubleIntegratorTB;
		TbParams p = e->tb;
		p.custom_cost = 0;
		const size_t smem = sizeof(double) * tb_smem_doubles_per_thread<M, 4>() * TB_THREADS; // stride of the snapshots: a full CTA's
		if (p.sat_mode == SAT_IDENTITY)
			tb_server_kernel<M, 4, SAT_IDENTITY><<<1, 32, smem, e->srv_stream>>>(p, e->srv_mailbox_dev, nullptr);
		else if (p.sat_mode >= SAT_POW2)
			tb_server_kernel<M, 4, SAT_POW2><<<1, 32, smem, e->srv_stream>>>(p, e->srv_mailbox_dev, nullptr);
		else
			tb_server_kernel<M, 4, SAT_GENERAL><<<1, 32, smem, e->srv_stream>>>(p, e->srv_mailbox_dev, nullptr);
	}
	CUDA_TRY(cudaGetLastError());
	e->srv_on = true;
	return ASIF_OK;
}

// one batch of n <= 32 states through the resident warp; returns ASIF_OK, or a negative code (then the caller's arrays are untouched)
int server_call(asif_engine *e, int64_t n, const double *x, const double *u_des, double *u_act, double *relax, int32_t *rc)
{
	const int nx = e->nx, nu = e->nu, nr = e->n_relax;
	double *mb = e->srv_mailbox;
	double *mx = mb + SRV_HDR, *mud = mx + SRV_MAX_STATES * nx, *mua = mud + SRV_MAX_STATES * nu, *mrl = mua + SRV_MAX_STATES * nu;
	int32_t *mrc = reinterpret_cast<int32_t *>(mrl + SRV_MAX_STATES * nr);
	memcpy(mx, x, sizeof(double) * n * nx);
	memcpy(mud, u_des, sizeof(double) * n * nu);
	unsigned long long *hdr = reinterpret_cast<unsigned long long *>(mb);
	const unsigned long long seq = ++e->srv_seq;
	__atomic_store_n(hdr + 1, (unsigned long long)n, __ATOMIC_RELEASE);
	__atomic_store_n(hdr, seq, __ATOMIC_RELEASE);
	const auto t0 = std::chrono::steady_clock::now();
	for (unsigned spins = 0;; spins++) {
		if (__atomic_load_n(hdr + 2, __ATOMIC_ACQUIRE) == seq) break;
		if ((spins & 0xffff) == 0xffff) { // every ~65k polls: has the kernel died, or are we waiting absurdly long?
			const cudaError_t q = cudaStreamQuery(e->srv_stream);
			if (q != cudaErrorNotReady) {
				e->srv_on = false;
				return fail(ASIF_ERR_CUDA, "latency server stopped: %s", cudaGetErrorString(q));
			}
			if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 10.0) {
				server_stop(e);
				return fail(ASIF_ERR_CUDA, "latency server did not answer within 10 s");
			}
		}
	}
	memcpy(u_act, mua, sizeof(double) * n * nu);
	memcpy(relax, mrl, sizeof(double) * n * nr);
	memcpy(rc, mrc, sizeof(int32_t) * n);
	return ASIF_OK;
}

// Allocation calls (cudaMalloc / cudaFree / cudaHostAlloc ...) may wait for the whole device, i.e. for ever while the
// server's warp is resident (observed: cudaHostAlloc of the small-batch scratch never returned).  Every entry point that
// can allocate therefore stops the server for its duration and restarts it on the way out (ServerPause, engine_internal.cuh;
// ~20 us, and such calls move megabytes).  The device-pointer filter path of the two server classes allocates nothing and
// runs beside the server.

// every stream this engine launches on (never cudaDeviceSynchronize: it would wait for a running latency server for ever)
int sync_engine_streams(asif_engine *e)
{
	if (e->stream) CUDA_TRY(cudaStreamSynchronize(e->stream));
	for (Slot &sl : e->slot)
		if (sl.stream) CUDA_TRY(cudaStreamSynchronize(sl.stream));
	return ASIF_OK;
}

// Device alias of a caller's host range when the device can address it (cudaHostAlloc'ed memory, or memory registered
// with cudaHostRegisterMapped; under unified addressing both report a device pointer), else nullptr.
template <class T>
T *mapped_alias(const T *p, size_t count)
{
	if (!p || !count) return nullptr;
	cudaPointerAttributes a0, a1;
	const char *last = reinterpret_cast<const char *>(p) + count * sizeof(T) - 1;
	if (cudaPointerGetAttributes(&a0, p) != cudaSuccess || cudaPointerGetAttributes(&a1, last) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	if (a0.type != cudaMemoryTypeHost || a1.type != cudaMemoryTypeHost || !a0.devicePointer || !a1.devicePointer) return nullptr;
	if (reinterpret_cast<const char *>(a1.devicePointer) - reinterpret_cast<const char *>(a0.devicePointer) != last - reinterpret_cast<const char *>(p))
		return nullptr; // first and last byte are not in one mapping
	return reinterpret_cast<T *>(a0.devicePointer);
}

// plain (unregistered) host memory: what cudaMemcpyAsync would stage through the driver's own bounce buffer
bool is_pageable(const void *p)
{
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	return a.type == cudaMemoryTypeUnregistered;
}

bool small_inplace_enabled()
{
	const char *v = getenv("ASIF_B200_SMALL_INPLACE");
	return !(v && v[0] == '0');
}

bool bounce_enabled()
{
	const char *v = getenv("ASIF_B200_BOUNCE");
	return !(v && v[0] == '0');
}

// ASIF_B200_HOST_IO: how a host-memory batch reaches the kernels.
//   "staged"  H2D copy, kernel, D2H copy per chunk (the only mode for pageable memory)
//   "out"     the kernels store uAct / relax / rc straight into the caller's pinned arrays over PCIe (posted writes,
//             whole 128 / 256 B lines per warp): no D2H copies, no drain; inputs still arrive by chunked H2D copies
//   "inout"   one launch over the caller's pinned arrays, inputs read over PCIe by the kernel as well
//   "auto"    (default) MEASURED per engine: which of "staged" and "inout" is faster depends on the filter class (kernel
//             time per state against link time per state) AND on what else uses the box's host side - with one GPU the
//             in-place launch wins for the trajectory-integrating classes (C2 5.27 vs 6.06 ms per 1e7 states), with four
//             GPUs sharing the host bridge the copy engines' larger PCIe payloads won (round 1, SCALE N = 4).  So the
//             first four large batches alternate between the class default (below) and the other mode (the first batch of
//             a mode pays for its buffers and is not counted), later ones run the faster of the two (wall clock per
//             state, exponentially averaged), and every 32nd batch re-tries the slower one.  Results
//             are the same bits in every mode, so the choice is free.  asif_engine_host_io_stats reports the averages.
//             Small batches (< 2^18 states) keep the class default: "inout" for the trajectory-integrating and realizable
//             filters, "staged" for the explicit and robust ones (ms per batch staged / out / inout, one GPU: C1 0.73 /
//             1.02 / 0.89, C2 6.06 / 5.97 / 5.27, C3a 32.1 / 31.8 / 32.5, C3b 1.02 / 1.30 / 1.43, C4 1.40 / 1.15 / 1.00,
//             C5-filter 12.7 / 12.5 / 12.1; profiles/r01_config_e2e_host_io.jsonl).
// Every mode falls back to "staged" when an array it needs is not device-addressable.
constexpr int64_t HOST_IO_PROBE_MIN_STATES = (int64_t)1 << 18;

int host_io_forced()
{
	const char *v = getenv("ASIF_B200_HOST_IO"); // read per call: a test or a caller may switch it between batches
	if (v && !strcmp(v, "staged")) return 0;
	if (v && !strcmp(v, "out")) return 1;
	if (v && !strcmp(v, "inout")) return 2;
	return -1;
}

int host_io_class_default(const asif_engine *e)
{
	return (e->cfg.filter == ASIF_FILTER_EXPLICIT || e->cfg.filter == ASIF_FILTER_ROBUST) ? 0 : 2;
}

// the mode this batch runs in; *probe says whether its wall time is to be recorded
int host_io_mode(asif_engine *e, int64_t n, bool *probe)
{
	*probe = false;
	const int forced = host_io_forced();
	if (forced >= 0) return forced;
	const int dflt = host_io_class_default(e), other = dflt == 2 ? 0 : 2;
	if (n < HOST_IO_PROBE_MIN_STATES) return dflt;
	*probe = true;
	// two batches per mode before anything is decided: a mode's first batch pays for its buffers (slot allocations, the
	// first touch of the pinned mappings) and its time is replaced, not averaged, by the second
	if (e->io_samples[dflt] == 0) return dflt;
	if (e->io_samples[other] == 0) return other;
	if (e->io_samples[dflt] == 1) return dflt;
	if (e->io_samples[other] == 1) return other;
	const int best = e->io_ms_per_state[other] < e->io_ms_per_state[dflt] ? other : dflt;
	if (++e->io_calls % 32 == 0) return best == dflt ? other : dflt; // conditions change (other ranks start or stop): look again
	return best;
}

void host_io_record(asif_engine *e, int mode, int64_t n, double ms)
{
	const double per = ms / (double)n;
	e->io_ms_per_state[mode] = e->io_samples[mode] >= 2 ? 0.5 * e->io_ms_per_state[mode] + 0.5 * per : per;
	e->io_samples[mode]++;
}

// switches the kernels of this engine between updateCost(uDes) and the caller's linear cost (filter(x, H, c, ...))
void set_custom_cost(asif_engine *e, int on)
{
	e->tb.custom_cost = on;
	e->im.custom_cost = on;
	e->ex.custom_cost = on;
}

// filter_batch proper; cw = doubles per state of the cost argument (nu for uDes, nv for c)
int32_t filter_batch_impl(asif_engine *e, int64_t n, const double *x, const double *u_des, const int cw, double *u_act,
                          double *relax, int32_t *rc, double *diag, int32_t mem, void *stream)
{
	if (!e) return fail(ASIF_ERR_INVALID_ARGUMENT, "engine is NULL");
	if (n < 0) return fail(ASIF_ERR_INVALID_ARGUMENT, "n < 0");
	if (n == 0) return ASIF_OK;
	if (!x || !u_des || !u_act || !relax || !rc) return fail(ASIF_ERR_INVALID_ARGUMENT, "NULL batch pointer");
	CUDA_TRY(cudaSetDevice(e->cfg.device));
	if (mem == ASIF_MEM_DEVICE) {
		cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
		CUDA_TRY(cudaMemsetAsync(e->d_counters, 0, N_COUNTERS * sizeof(unsigned long long), st));
		e->ex.lfh = e->lie_lfh;
		e->ex.lgh = e->lie_lgh;
		int r = launch_filter(e, n, x, u_des, u_act, relax, rc, diag, st);
		if (r) return r;
		if (!stream) CUDA_TRY(cudaStreamSynchronize(st));
		return ASIF_OK;
	}
	if (mem != ASIF_MEM_HOST) return fail(ASIF_ERR_INVALID_ARGUMENT, "mem must be ASIF_MEM_HOST or ASIF_MEM_DEVICE");
	// host memory: chunks rotate over N_SLOTS streams so that the H2D copy of chunk i+1, the kernel of
	// chunk i and the D2H copy of chunk i-1 overlap (they do when the caller's buffers are pinned)
	const int nx = e->nx, nu = e->nu, nr = e->n_relax, nd = e->n_diag;
	// Small batches (the single-state filter() of the host classes above all): five pageable copies and a kernel cost six
	// driver calls that each wait for the one before.  Instead the arrays are copied by the CPU into a pinned scratch the
	// device can address, one launch reads and writes it over PCIe, one synchronise, and the results are copied back.
	// ASIF_B200_SMALL_INPLACE=0 switches it off.
	const size_t o_x = 0, o_ud = o_x + sizeof(double) * n * nx, o_ua = o_ud + sizeof(double) * n * cw,
	             o_relax = o_ua + sizeof(double) * n * nu, o_diag = o_relax + sizeof(double) * n * nr,
	             o_rc = o_diag + (diag ? sizeof(double) * n * nd : 0), total = o_rc + sizeof(int32_t) * n;
	if (e->srv_on && n <= SRV_MAX_STATES && !diag && cw == nu && !e->lie_lfh) { // latency server: no launch at all
		e->last_host_io = ASIF_HOST_IO_INOUT;
		return server_call(e, n, x, u_des, u_act, relax, rc);
	}
	ServerPause pause(e); // everything below may allocate
	{
		if (total <= SMALL_BATCH_BYTES && small_inplace_enabled()) {
			if (!e->small_h) CUDA_TRY(cudaHostAlloc((void **)&e->small_h, SMALL_BATCH_BYTES, cudaHostAllocMapped));
			if (!e->slot[0].stream) CUDA_TRY(cudaStreamCreateWithFlags(&e->slot[0].stream, cudaStreamNonBlocking));
			cudaStream_t st = e->slot[0].stream;
			char *h = e->small_h;
			memcpy(h + o_x, x, sizeof(double) * n * nx);
			memcpy(h + o_ud, u_des, sizeof(double) * n * cw);
			e->ex.lfh = e->lie_lfh;
			e->ex.lgh = e->lie_lgh;
			e->last_host_io = ASIF_HOST_IO_INOUT;
			// latency path: the QP-work statistic is not collected (no counter reset to enqueue, no atomics in the kernel)
			e->ctr = nullptr;
			const int r = launch_filter(e, n, (const double *)(h + o_x), (const double *)(h + o_ud), (double *)(h + o_ua),
			                            (double *)(h + o_relax), (int32_t *)(h + o_rc), diag ? (double *)(h + o_diag) : nullptr, st);
			e->ctr = e->d_counters;
			if (r) return r;
			CUDA_TRY(cudaStreamSynchronize(st));
			memcpy(u_act, h + o_ua, sizeof(double) * n * nu);
			memcpy(relax, h + o_relax, sizeof(double) * n * nr);
			memcpy(rc, h + o_rc, sizeof(int32_t) * n);
			if (diag) memcpy(diag, h + o_diag, sizeof(double) * n * nd);
			return ASIF_OK;
		}
	}
	// batches under four full chunks are cut in four (not below 2^16 states), so that their copies overlap too:
	// 1e6 states, staged, 2^19 -> 2^18: C1 0.73 -> 0.68 ms, C3b 1.02 -> 0.81 ms, C4 1.40 -> 1.19 ms
	int64_t chunk = chunk_states();
	if (n < 4 * chunk) {
		const int64_t quarter = (n + 3) / 4, floor_ = (int64_t)1 << 16;
		chunk = quarter > floor_ ? quarter : (n < floor_ ? n : floor_);
		if (chunk > chunk_states()) chunk = chunk_states();
	}
	// the QP-work counter is zeroed on the engine's own stream and waited for: the slot streams are non-blocking and
	// do not order against the legacy default stream a plain cudaMemset would run on
	CUDA_TRY(cudaMemsetAsync(e->d_counters, 0, N_COUNTERS * sizeof(unsigned long long), e->stream));
	CUDA_TRY(cudaStreamSynchronize(e->stream));
	// device aliases of the caller's output arrays (nullptr: pageable or unmapped memory, staged copies then)
	double *m_ua = nullptr, *m_relax = nullptr, *m_diag = nullptr;
	int32_t *m_rc = nullptr;
	bool probe = false;
	const int io = host_io_mode(e, n, &probe);
	const auto t_io0 = std::chrono::steady_clock::now();
	auto record = [&](int mode) { // wall time of this batch into the engine's per-mode average ("auto" policy)
		if (probe && mode == io)
			host_io_record(e, mode, n, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_io0).count());
	};
	if (io >= 1) {
		m_ua = mapped_alias(u_act, (size_t)n * nu);
		m_relax = mapped_alias(relax, (size_t)n * nr);
		m_rc = mapped_alias(rc, (size_t)n);
		m_diag = diag ? mapped_alias(diag, (size_t)n * nd) : nullptr;
	}
	const bool direct_out = m_ua && m_relax && m_rc && (!diag || m_diag);
	e->last_host_io = direct_out ? ASIF_HOST_IO_OUT : ASIF_HOST_IO_STAGED;
	if (io == 2 && direct_out) {
		const double *m_x = mapped_alias(x, (size_t)n * nx), *m_ud = mapped_alias(u_des, (size_t)n * cw);
		if (m_x && m_ud) {
			if (!e->slot[0].stream) CUDA_TRY(cudaStreamCreateWithFlags(&e->slot[0].stream, cudaStreamNonBlocking));
			e->ex.lfh = e->lie_lfh;
			e->ex.lgh = e->lie_lgh;
			e->last_host_io = ASIF_HOST_IO_INOUT;
			const int r = launch_filter(e, n, m_x, m_ud, m_ua, m_relax, m_rc, m_diag, e->slot[0].stream);
			if (r) return r;
			CUDA_TRY(cudaStreamSynchronize(e->slot[0].stream));
			record(2);
			return ASIF_OK;
		}
	}
	// Pageable caller arrays (plain new[] / std::vector / numpy memory): a pageable cudaMemcpyAsync is staged by the
	// driver on the calling thread (measured 11 GB/s), so large batches bounce through pinned staging buffers of the
	// slots instead, filled and drained by a few host threads (host_copier.hpp); ASIF_B200_BOUNCE=0 switches it off.
	const size_t out_row = sizeof(double) * (nu + nr + (diag ? nd : 0)) + sizeof(int32_t), in_row = sizeof(double) * (nx + cw);
	bool bounce_in = false, bounce_out = false;
	if ((size_t)n * (in_row + out_row) >= ((size_t)8 << 20) && bounce_enabled()) {
		bounce_in = is_pageable(x) && is_pageable(u_des);
		bounce_out = !direct_out && is_pageable(u_act) && is_pageable(relax) && is_pageable(rc) && (!diag || is_pageable(diag));
		if ((bounce_in || bounce_out) && !e->copier) {
			try { // nothing may throw through the C boundary
				e->copier = new HostCopier(e->copy_threads > 0 ? e->copy_threads : HostCopier::default_threads());
			} catch (const std::exception &ex) {
				return fail(ASIF_ERR_INTERNAL, "host copier threads: %s", ex.what());
			}
		}
	}
	const size_t h_x = 0, h_ud = h_x + sizeof(double) * chunk * nx, h_ua = h_ud + sizeof(double) * chunk * e->nv,
	             h_relax = h_ua + sizeof(double) * chunk * nu, h_diag = h_relax + sizeof(double) * chunk * nr,
	             h_rc = h_diag + (diag ? sizeof(double) * chunk * nd : 0), h_bytes = h_rc + sizeof(int32_t) * chunk;
	auto copy_out = [&](Slot &s) { // results of the slot's last chunk: staging -> caller arrays
		if (!s.pend_m) return;
		const int64_t o = s.pend_off, c = s.pend_m;
		HostCopier::Piece jobs[4] = {{u_act + o * nu, s.h + h_ua, sizeof(double) * c * nu},
		                             {relax + o * nr, s.h + h_relax, sizeof(double) * c * nr},
		                             {rc + o, s.h + h_rc, sizeof(int32_t) * c},
		                             {diag ? diag + o * nd : nullptr, s.h + h_diag, diag ? sizeof(double) * c * nd : 0}};
		e->copier->copy(jobs, diag ? 4 : 3);
		s.pend_m = 0;
	};
	for (Slot &s : e->slot) s.pend_m = 0;
	int si = 0;
	// Chunk schedule: the pipeline's fill (first H2D + first kernel) and drain (last kernel + last D2H) are not overlapped
	// with anything, so the first two and the last two chunks are a quarter and a half of the steady-state size
	// (env ASIF_B200_CHUNK_RAMP=0 switches the ramp off).
	static const bool ramp = []() {
		const char *v = getenv("ASIF_B200_CHUNK_RAMP");
		return !(v && v[0] == '0');
	}();
	const int64_t small_ = chunk / 4 > 0 ? chunk / 4 : 1, mid_ = chunk / 2 > 0 ? chunk / 2 : 1;
	const bool use_ramp = ramp && n >= 4 * chunk && chunk == chunk_states(); // not for batches already cut in four
	int64_t m = 0;
	int idx = 0;
	for (int64_t off = 0; off < n; off += m, si = (si + 1) % N_SLOTS, idx++) {
		const int64_t left = n - off;
		m = chunk;
		if (use_ramp) {
			if (idx == 0) m = small_;
			else if (idx == 1) m = mid_;
			else if (left <= small_ + mid_) m = left > small_ ? left - small_ : left; // the last two: about a half, then a quarter
			else if (left < chunk + small_ + mid_) m = left - (small_ + mid_);        // what remains before them
		}
		if (m > left) m = left;
		Slot &s = e->slot[si];
		if (s.stream) CUDA_TRY(cudaStreamSynchronize(s.stream)); // slot buffers free again
		if (bounce_out) copy_out(s);
		int r = ensure_slot(e, s, chunk, diag != nullptr && !direct_out);
		if (r) return r;
		if ((bounce_in || bounce_out) && s.h_cap < h_bytes) {
			if (s.h) cudaFreeHost(s.h);
			s.h = nullptr, s.h_cap = 0;
			CUDA_TRY(cudaHostAlloc((void **)&s.h, h_bytes, cudaHostAllocDefault));
			s.h_cap = h_bytes;
		}
		const double *src_x = x + off * nx, *src_ud = u_des + off * cw;
		if (bounce_in) {
			HostCopier::Piece jobs[2] = {{s.h + h_x, src_x, sizeof(double) * m * nx}, {s.h + h_ud, src_ud, sizeof(double) * m * cw}};
			e->copier->copy(jobs, 2);
			src_x = reinterpret_cast<const double *>(s.h + h_x);
			src_ud = reinterpret_cast<const double *>(s.h + h_ud);
		}
		CUDA_TRY(cudaMemcpyAsync(s.x, src_x, sizeof(double) * m * nx, cudaMemcpyHostToDevice, s.stream));
		CUDA_TRY(cudaMemcpyAsync(s.ud, src_ud, sizeof(double) * m * cw, cudaMemcpyHostToDevice, s.stream));
		e->ex.lfh = e->lie_lfh ? e->lie_lfh + off * e->nc : nullptr; // device copies of the whole batch (filter_batch_lie)
		e->ex.lgh = e->lie_lgh ? e->lie_lgh + off * e->nc * e->nu : nullptr;
		if (direct_out) {
			r = launch_filter(e, m, s.x, s.ud, m_ua + off * nu, m_relax + off * nr, m_rc + off, diag ? m_diag + off * nd : nullptr, s.stream);
			if (r) return r;
			continue;
		}
		r = launch_filter(e, m, s.x, s.ud, s.ua, s.relax, s.rc, diag ? s.diag : nullptr, s.stream);
		if (r) return r;
		if (bounce_out) {
			CUDA_TRY(cudaMemcpyAsync(s.h + h_ua, s.ua, sizeof(double) * m * nu, cudaMemcpyDeviceToHost, s.stream));
			CUDA_TRY(cudaMemcpyAsync(s.h + h_relax, s.relax, sizeof(double) * m * nr, cudaMemcpyDeviceToHost, s.stream));
			CUDA_TRY(cudaMemcpyAsync(s.h + h_rc, s.rc, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, s.stream));
			if (diag) CUDA_TRY(cudaMemcpyAsync(s.h + h_diag, s.diag, sizeof(double) * m * nd, cudaMemcpyDeviceToHost, s.stream));
			s.pend_off = off, s.pend_m = m;
			continue;
		}
		CUDA_TRY(cudaMemcpyAsync(u_act + off * nu, s.ua, sizeof(double) * m * nu, cudaMemcpyDeviceToHost, s.stream));
		CUDA_TRY(cudaMemcpyAsync(relax + off * nr, s.relax, sizeof(double) * m * nr, cudaMemcpyDeviceToHost, s.stream));
		CUDA_TRY(cudaMemcpyAsync(rc + off, s.rc, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, s.stream));
		if (diag) CUDA_TRY(cudaMemcpyAsync(diag + off * nd, s.diag, sizeof(double) * m * nd, cudaMemcpyDeviceToHost, s.stream));
	}
	for (int k = 0; k < N_SLOTS; k++, si = (si + 1) % N_SLOTS) { // oldest chunk first
		Slot &s = e->slot[si];
		if (s.stream) CUDA_TRY(cudaStreamSynchronize(s.stream));
		if (bounce_out) copy_out(s);
	}
	if (!bounce_in && !bounce_out) record(direct_out ? 1 : 0); // pinned arrays only: pageable batches say nothing about the link
	return ASIF_OK;
}
} // namespace

extern "C" {

int32_t asif_engine_filter_batch(asif_engine *e, int64_t n, const double *x, const double *u_des, double *u_act,
                                 double *relax, int32_t *rc, double *diag, int32_t mem, void *stream)
{
	if (!e) return fail(ASIF_ERR_INVALID_ARGUMENT, "engine is NULL");
	set_custom_cost(e, 0);
	return filter_batch_impl(e, n, x, u_des, e->nu, u_act, relax, rc, diag, mem, stream);
}

int32_t asif_engine_set_input_cost(asif_engine *e, const double *H)
{
	if (!e || !H) return fail(ASIF_ERR_INVALID_ARGUMENT, "NULL argument");
	const int nu = e->nu;
	for (int i = 0; i < nu; i++)
		if (!(H[i + i * nu] > 0.0)) return fail(ASIF_ERR_INVALID_ARGUMENT, "H[%d][%d] must be > 0", i, i);
	// updateH (src/asif_implicit_tb.cpp:748-762): the nu x nu block of H_; with diagonalCost = true (the constructors'
	// default, the only mode built here) the QP wrapper reads its diagonal only (src/qpwrapper_osqp.cpp:267-283)
	double *gis[] = {e->tb.gi, e->im.gi, e->ex.gi, e->rb.gi, e->rz.gi};
	double *gihs[] = {e->tb.gih, e->im.gih, e->ex.gih, e->rb.gih, e->rz.gih};
	for (int t = 0; t < 5; t++)
		for (int i = 0; i < nu; i++) {
			gis[t][i] = 1.0 / (2.0 * H[i + i * nu]);
			gihs[t][i] = sqrt(gis[t][i]);
		}
	if (e->srv_on) { // the resident kernel holds its parameters by value: restart it with the new Hessian
		int r = server_stop(e);
		if (!r) r = server_start(e);
		if (r) return r;
	}
	return ASIF_OK;
}

int32_t asif_engine_latency_server(asif_engine *e, int32_t on)
{
	if (!e) return fail(ASIF_ERR_INVALID_ARGUMENT, "engine is NULL");
	return on ? server_start(e) : server_stop(e);
}

int32_t asif_engine_filter_batch_cost(asif_engine *e, int64_t n, const double *x, const double *H, const double *c,
                                      double *u_act, double *relax, int32_t *rc, double *diag, int32_t mem, void *stream)
{
	if (!e) return fail(ASIF_ERR_INVALID_ARGUMENT, "engine is NULL");
	if (e->cfg.filter == ASIF_FILTER_ROBUST || e->cfg.filter == ASIF_FILTER_REALIZABLE)
		return fail(ASIF_ERR_UNSUPPORTED, "filter(x, H, c, ...) of ASIFrobust / ASIFrealizable takes a cost over the LP-dual multipliers "
		            "(nv = 402 / 38), which the reduced formulation solved here does not carry");
	if (H) {
		const int r = asif_engine_set_input_cost(e, H);
		if (r) return r;
	}
	set_custom_cost(e, 1);
	const int32_t r = filter_batch_impl(e, n, x, c, e->nv, u_act, relax, rc, diag, mem, stream);
	set_custom_cost(e, 0);
	return r;
}

int32_t asif_engine_filter_batch_lie(asif_engine *e, int64_t n, const double *x, const double *u_des, const double *Lfh,
                                     const double *Lgh, double *u_act, double *relax, int32_t *rc, double *diag, int32_t mem,
                                     void *stream)
{
	if (!e) return fail(ASIF_ERR_INVALID_ARGUMENT, "engine is NULL");
	if (e->cfg.filter != ASIF_FILTER_EXPLICIT)
		return fail(ASIF_ERR_UNSUPPORTED, "filter(x, uDes, uAct, Lfh, Lgh, relax) exists in ASIF::ASIF only (filter %d given)", e->cfg.filter);
	if (!Lfh || !Lgh) return fail(ASIF_ERR_INVALID_ARGUMENT, "Lfh / Lgh is NULL");
	if (n <= 0) return n < 0 ? fail(ASIF_ERR_INVALID_ARGUMENT, "n < 0") : ASIF_OK;
	CUDA_TRY(cudaSetDevice(e->cfg.device));
	ServerPause pause(e);
	set_custom_cost(e, 0);
	double *dl = nullptr, *dg = nullptr;
	if (mem == ASIF_MEM_DEVICE) {
		e->lie_lfh = Lfh;
		e->lie_lgh = Lgh;
	} else { // host arrays: one device copy of the whole batch (8 (nc + nc nu) bytes per state), chunks index into it
		const size_t bl = sizeof(double) * (size_t)n * e->nc, bg = bl * e->nu;
		cudaError_t ce = cudaMalloc(&dl, bl);
		if (ce == cudaSuccess) ce = cudaMalloc(&dg, bg);
		if (ce == cudaSuccess) ce = cudaMemcpy(dl, Lfh, bl, cudaMemcpyHostToDevice);
		if (ce == cudaSuccess) ce = cudaMemcpy(dg, Lgh, bg, cudaMemcpyHostToDevice);
		if (ce != cudaSuccess) {
			cudaFree(dl);
			cudaFree(dg);
			return fail(ASIF_ERR_CUDA, "Lfh / Lgh upload failed: %s", cudaGetErrorString(ce));
		}
		e->lie_lfh = dl;
		e->lie_lgh = dg;
	}
	const int32_t r = filter_batch_impl(e, n, x, u_des, e->nu, u_act, relax, rc, diag, mem, stream);
	e->lie_lfh = e->lie_lgh = nullptr;
	e->ex.lfh = e->ex.lgh = nullptr;
	if (dl || dg) {
		sync_engine_streams(e);
		cudaFree(dl);
		cudaFree(dg);
	}
	return r;
}

int32_t asif_engine_rollout(asif_engine *e, int64_t n, int32_t steps, double dt, double *x, const double *u_des,
                            double *u_act_last, int32_t *rc_last, int64_t *rc_hist, int32_t mem, void *stream)
{
	if (!e) return fail(ASIF_ERR_INVALID_ARGUMENT, "engine is NULL");
	if (n < 0 || steps < 0) return fail(ASIF_ERR_INVALID_ARGUMENT, "n < 0 or steps < 0");
	if (n == 0) return ASIF_OK;
	if (!x || !u_des || !u_act_last || !rc_last) return fail(ASIF_ERR_INVALID_ARGUMENT, "NULL batch pointer");
	CUDA_TRY(cudaSetDevice(e->cfg.device));
	const int nx = e->nx, nu = e->nu;
	if (mem == ASIF_MEM_DEVICE) {
		cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
		CUDA_TRY(cudaMemsetAsync(e->d_counters, 0, N_COUNTERS * sizeof(unsigned long long), st));
		int r = launch_rollout(e, n, steps, dt, x, u_des, u_act_last, rc_last, st);
		if (r) return r;
		if (rc_hist || !stream) {
			CUDA_TRY(cudaStreamSynchronize(st));
			if (rc_hist) {
				unsigned long long h[9];
				CUDA_TRY(cudaMemcpy(h, e->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
				for (int i = 0; i < 8; i++) rc_hist[i] = (int64_t)h[1 + i];
			}
		}
		return ASIF_OK;
	}
	if (mem != ASIF_MEM_HOST) return fail(ASIF_ERR_INVALID_ARGUMENT, "mem must be ASIF_MEM_HOST or ASIF_MEM_DEVICE");
	ServerPause pause(e);
	Slot &s = e->slot[0];
	int r = ensure_slot(e, s, n, false);
	if (r) return r;
	CUDA_TRY(cudaMemsetAsync(e->d_counters, 0, N_COUNTERS * sizeof(unsigned long long), s.stream));
	CUDA_TRY(cudaMemcpyAsync(s.x, x, sizeof(double) * n * nx, cudaMemcpyHostToDevice, s.stream));
	CUDA_TRY(cudaMemcpyAsync(s.ud, u_des, sizeof(double) * n * nu, cudaMemcpyHostToDevice, s.stream));
	r = launch_rollout(e, n, steps, dt, s.x, s.ud, s.ua, s.rc, s.stream);
	if (r) return r;
	CUDA_TRY(cudaMemcpyAsync(x, s.x, sizeof(double) * n * nx, cudaMemcpyDeviceToHost, s.stream));
	CUDA_TRY(cudaMemcpyAsync(u_act_last, s.ua, sizeof(double) * n * nu, cudaMemcpyDeviceToHost, s.stream));
	CUDA_TRY(cudaMemcpyAsync(rc_last, s.rc, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s.stream));
	CUDA_TRY(cudaStreamSynchronize(s.stream));
	if (rc_hist) {
		unsigned long long h[9];
		CUDA_TRY(cudaMemcpy(h, e->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
		for (int i = 0; i < 8; i++) rc_hist[i] = (int64_t)h[1 + i];
	}
	return ASIF_OK;
}

int32_t asif_engine_set_learning(asif_engine *e, const asif_learning_data *d)
{
	if (!e) return fail(ASIF_ERR_INVALID_ARGUMENT, "engine is NULL");
	if (e->cfg.filter != ASIF_FILTER_IMPLICIT && e->cfg.filter != ASIF_FILTER_IMPLICIT_RB)
		return fail(ASIF_ERR_UNSUPPORTED, "the learned residual exists in ASIFimplicit and ASIFimplicitRB only (filter %d given)", e->cfg.filter);
	CUDA_TRY(cudaSetDevice(e->cfg.device));
	CUDA_TRY(cudaDeviceSynchronize()); // no launch may still read the previous networks
	cudaFree(e->d_learn);
	e->d_learn = nullptr;
	memset(&e->im.learn, 0, sizeof(e->im.learn));
	if (!d) return ASIF_OK;
	const uint32_t din[2] = {d->d_drift_in, d->d_act_in}, dh1[2] = {d->d_drift_hidden, d->d_act_hidden};
	const uint32_t dh2[2] = {d->d_drift_hidden_2, d->d_act_hidden_2}, dout[2] = {d->d_drift_out, d->d_act_out};
	const double *w[2][6] = {{d->w_1_drift, d->b_1_drift, d->w_2_drift, d->b_2_drift, d->w_3_drift, d->b_3_drift},
	                         {d->w_1_act, d->b_1_act, d->w_2_act, d->b_2_act, d->w_3_act, d->b_3_act}};
	size_t len[2][6], total = 0;
	for (int t = 0; t < 2; t++) {
		if (din[t] < 2u * (uint32_t)e->nx || din[t] > (uint32_t)LEARN_MAX_WIDTH || dh1[t] < 1 || dh1[t] > (uint32_t)LEARN_MAX_WIDTH ||
		    dh2[t] < 1 || dh2[t] > (uint32_t)LEARN_MAX_WIDTH || dout[t] < (t == 0 ? 1u : (uint32_t)e->nu) || dout[t] > (uint32_t)LEARN_MAX_WIDTH)
			return fail(ASIF_ERR_INVALID_ARGUMENT, "learning data: net %d has widths in %u, hidden %u / %u, out %u (need in >= 2 nx, all <= %d)",
			            t, din[t], dh1[t], dh2[t], dout[t], LEARN_MAX_WIDTH);
		len[t][0] = (size_t)dh1[t] * din[t]; len[t][1] = dh1[t];
		len[t][2] = (size_t)dh2[t] * dh1[t]; len[t][3] = dh2[t];
		len[t][4] = (size_t)dout[t] * dh2[t]; len[t][5] = dout[t];
		for (int k = 0; k < 6; k++) {
			if (!w[t][k]) return fail(ASIF_ERR_INVALID_ARGUMENT, "learning data: a weight or bias pointer is NULL");
			total += len[t][k];
		}
	}
	double *host = new (std::nothrow) double[total];
	if (!host) return fail(ASIF_ERR_INVALID_ARGUMENT, "out of host memory");
	size_t pos = 0;
	LearnNets L;
	memset(&L, 0, sizeof(L));
	for (int t = 0; t < 2; t++) {
		L.off[t] = (int32_t)pos;
		L.d_in[t] = (int32_t)din[t]; L.d_h1[t] = (int32_t)dh1[t]; L.d_h2[t] = (int32_t)dh2[t]; L.d_out[t] = (int32_t)dout[t];
		for (int k = 0; k < 6; k++) {
			memcpy(host + pos, w[t][k], sizeof(double) * len[t][k]);
			pos += len[t][k];
		}
	}
	cudaError_t ce = cudaMalloc(&e->d_learn, sizeof(double) * total);
	if (ce == cudaSuccess) ce = cudaMemcpy(e->d_learn, host, sizeof(double) * total, cudaMemcpyHostToDevice);
	delete[] host;
	if (ce != cudaSuccess) {
		cudaFree(e->d_learn);
		e->d_learn = nullptr;
		return fail(ASIF_ERR_CUDA, "learning data upload failed: %s", cudaGetErrorString(ce));
	}
	L.blob = e->d_learn;
	e->im.learn = L;
	return ASIF_OK;
}

int32_t asif_engine_last_host_io(const asif_engine *e, int32_t *mode)
{
	if (!e || !mode) return fail(ASIF_ERR_INVALID_ARGUMENT, "NULL argument");
	*mode = e->last_host_io;
	return ASIF_OK;
}

int32_t asif_engine_host_io_stats(const asif_engine *e, double ms_per_1e6_states[3], int32_t samples[3])
{
	if (!e || !ms_per_1e6_states || !samples) return fail(ASIF_ERR_INVALID_ARGUMENT, "NULL argument");
	for (int i = 0; i < 3; i++) {
		ms_per_1e6_states[i] = e->io_ms_per_state[i] * 1e6;
		samples[i] = e->io_samples[i];
	}
	return ASIF_OK;
}

int32_t asif_host_alloc(void **p, uint64_t bytes)
{
	if (!p) return fail(ASIF_ERR_INVALID_ARGUMENT, "NULL argument");
	*p = nullptr;
	if (!bytes) return ASIF_OK;
	CUDA_TRY(cudaHostAlloc(p, (size_t)bytes, cudaHostAllocPortable | cudaHostAllocMapped));
	return ASIF_OK;
}

int32_t asif_host_free(void *p)
{
	if (p) CUDA_TRY(cudaFreeHost(p));
	return ASIF_OK;
}

int32_t asif_host_register(void *p, uint64_t bytes)
{
	if (!p || !bytes) return fail(ASIF_ERR_INVALID_ARGUMENT, "NULL or empty range");
	CUDA_TRY(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
	return ASIF_OK;
}

int32_t asif_host_unregister(void *p)
{
	if (!p) return fail(ASIF_ERR_INVALID_ARGUMENT, "NULL argument");
	CUDA_TRY(cudaHostUnregister(p));
	return ASIF_OK;
}

int32_t asif_engine_last_qp_iterations(asif_engine *e, uint64_t *rows_processed)
{
	if (!e || !rows_processed) return fail(ASIF_ERR_INVALID_ARGUMENT, "NULL argument");
	CUDA_TRY(cudaSetDevice(e->cfg.device));
	{
		const int r = sync_engine_streams(e);
		if (r) return r;
	}
	unsigned long long part[QP_CTR_SPREAD], v = 0;
	CUDA_TRY(cudaMemcpy(part, e->d_counters + QP_CTR_BASE, sizeof(part), cudaMemcpyDeviceToHost));
	for (int i = 0; i < QP_CTR_SPREAD; i++) v += part[i];
	*rows_processed = (uint64_t)v;
	return ASIF_OK;
}

} // extern "C"

// -------------------------------------------------------------------------------------------------
namespace {
template <int NV>
int launch_qp(int64_t n, int nc, int diag_cost, const double *H, const double *c, const double *A, const double *b,
              const double *lb, const double *ub, const uint8_t *be, double *sol, int32_t *status, int share, cudaStream_t st)
{
	int threads = QPB_THREADS;
	size_t smem = sizeof(double) * (size_t)(nc * NV + nc) * (threads + 1);
	if (smem > 200 * 1024) { // fewer problems per CTA
		threads = 32;
		smem = sizeof(double) * (size_t)(nc * NV + nc) * (threads + 1);
	}
	if (smem > 200 * 1024) return fail(ASIF_ERR_UNSUPPORTED, "qp_solve_batch: nc = %d too large for the shared-memory slab", nc);
	auto k = qp_batch_kernel<NV>;
	int r = set_smem(k, smem);
	if (r) return r;
	const unsigned blocks = (unsigned)((n + threads - 1) / threads);
	k<<<blocks, threads, smem, st>>>(n, nc, diag_cost, H, c, A, b, lb, ub, be, sol, status, share & ASIF_QP_SHARED_H,
	                                      share & ASIF_QP_SHARED_BOUNDS);
	CUDA_TRY(cudaGetLastError());
	return ASIF_OK;
}

// nv <= MAX_NV: one problem per thread, exact dual active-set method (qp_gi.cuh); nv > MAX_NV: one cluster per problem,
// operator splitting + polish (qp_admm.cuh / qp_admm.cu)
int launch_qp_any(int device, int nv, int64_t n, int nc, int diag_cost, const double *H, const double *c, const double *A,
                  const double *b, const double *lb, const double *ub, const uint8_t *be, double *sol, int32_t *status, int share,
                  cudaStream_t st, bool host_call)
{
	switch (nv) {
	case 1: return launch_qp<1>(n, nc, diag_cost, H, c, A, b, lb, ub, be, sol, status, share, st);
	case 2: return launch_qp<2>(n, nc, diag_cost, H, c, A, b, lb, ub, be, sol, status, share, st);
	case 3: return launch_qp<3>(n, nc, diag_cost, H, c, A, b, lb, ub, be, sol, status, share, st);
	case 4: return launch_qp<4>(n, nc, diag_cost, H, c, A, b, lb, ub, be, sol, status, share, st);
	default: return asifb::launch_qp_admm(device, nv, nc, n, diag_cost, H, c, A, b, lb, ub, be, sol, status, share, st, host_call);
	}
}
} // namespace

extern "C" int32_t asif_qp_solve_batch(int32_t device, int32_t nv, int32_t nc, int64_t n, int32_t diagonal_cost, const double *H,
                            const double *c, const double *A, const double *b, const double *lb, const double *ub,
                            const uint8_t *be, double *sol, int32_t *status, int32_t share_flags, int32_t mem, void *stream)
{
	if (nv < 1) return fail(ASIF_ERR_UNSUPPORTED, "qp_solve_batch needs nv >= 1 (got %d)", nv);
	if (nc < 0 || n < 0) return fail(ASIF_ERR_INVALID_ARGUMENT, "nc < 0 or n < 0");
	if (n == 0) return ASIF_OK;
	if (!H || !c || (nc > 0 && (!A || !b)) || !lb || !ub || !sol || !status) return fail(ASIF_ERR_INVALID_ARGUMENT, "NULL pointer");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
		return fail(ASIF_ERR_NO_DEVICE, "no CUDA device; this engine has no CPU fallback");
	if (device < 0 || device >= ndev) return fail(ASIF_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)", device, ndev);
	CUDA_TRY(cudaSetDevice(device));
	cudaStream_t st = (cudaStream_t)stream;
	const bool shH = share_flags & ASIF_QP_SHARED_H, shB = share_flags & ASIF_QP_SHARED_BOUNDS;
	const double *dH = H, *dc = c, *dA = A, *db = b, *dlb = lb, *dub = ub;
	const uint8_t *dbe = be;
	double *dsol = sol;
	int32_t *dstatus = status;
	void *scratch = nullptr;
	if (mem == ASIF_MEM_HOST && n <= 8) {
		// Latency path (the QPWrapper interface solves ONE problem per call): the problem is packed into a
		// pinned, device-mapped buffer which the kernel reads and writes directly - one launch and one
		// synchronisation, no cudaMemcpy at all.
		static thread_local void *tl_pin = nullptr;
		static thread_local size_t tl_pin_cap = 0;
		const size_t sH = (shH ? 1 : n) * (size_t)nv * nv, sc = (size_t)n * nv, sA = (size_t)n * nc * nv, sb = (size_t)n * nc;
		const size_t sB = (shB ? 1 : n) * (size_t)nv, ssol = (size_t)n * nv;
		const size_t bytes = (sH + sc + sA + sb + 2 * sB + ssol) * sizeof(double) + (size_t)n * sizeof(int32_t) + (be ? (size_t)nc : 0) + 64;
		if (bytes > tl_pin_cap) {
			if (tl_pin) cudaFreeHost(tl_pin);
			tl_pin = nullptr;
			tl_pin_cap = 0;
			const size_t want = bytes < (1u << 16) ? (1u << 16) : bytes;
			CUDA_TRY(cudaHostAlloc(&tl_pin, want, cudaHostAllocMapped | cudaHostAllocPortable));
			tl_pin_cap = want;
		}
		double *p = (double *)tl_pin;
		double *hH = p; p += sH;
		double *hc = p; p += sc;
		double *hA = p; p += sA;
		double *hb = p; p += sb;
		double *hlb = p; p += sB;
		double *hub = p; p += sB;
		double *hsol = p; p += ssol;
		int32_t *hst = (int32_t *)p;
		uint8_t *hbe = (uint8_t *)(hst + n);
		memcpy(hH, H, sH * 8);
		memcpy(hc, c, sc * 8);
		if (nc) {
			memcpy(hA, A, sA * 8);
			memcpy(hb, b, sb * 8);
		}
		memcpy(hlb, lb, sB * 8);
		memcpy(hub, ub, sB * 8);
		if (be) memcpy(hbe, be, nc);
		int r;
		// unified addressing: the mapped host pointers are valid device pointers
		r = launch_qp_any(device, nv, n, nc, diagonal_cost, hH, hc, hA, hb, hlb, hub, be ? hbe : nullptr, hsol, hst, share_flags, st, true);
		if (r) return r;
		CUDA_TRY(cudaStreamSynchronize(st));
		memcpy(sol, hsol, ssol * 8);
		memcpy(status, hst, (size_t)n * sizeof(int32_t));
		return ASIF_OK;
	}
	if (mem == ASIF_MEM_HOST) {
		// one allocation, carved up; sizes in doubles
		const size_t sH = (shH ? 1 : n) * (size_t)nv * nv, sc = (size_t)n * nv, sA = (size_t)n * nc * nv, sb = (size_t)n * nc;
		const size_t sB = (shB ? 1 : n) * (size_t)nv, ssol = (size_t)n * nv;
		const size_t nd = sH + sc + sA + sb + 2 * sB + ssol;
		const size_t bytes = nd * sizeof(double) + (size_t)n * sizeof(int32_t) + (be ? (size_t)nc : 0) + 64;
		// per-thread, per-device scratch that only grows: the single-problem QPWrapper path (n = 1, one call per
		// filter() of an unmodified reference user) must not pay a cudaMalloc / cudaFree per solve
		{
			static thread_local void *tl_buf[16] = {nullptr};
			static thread_local size_t tl_cap[16] = {0};
			static thread_local int tl_dev[16] = {0};
			const int slot = device & 15;
			if (tl_buf[slot] && tl_dev[slot] != device) { // an ordinal >= 16 shares the slot of another device: never its memory
				cudaSetDevice(tl_dev[slot]);
				cudaFree(tl_buf[slot]);
				cudaSetDevice(device);
				tl_buf[slot] = nullptr;
				tl_cap[slot] = 0;
			}
			tl_dev[slot] = device;
			if (bytes > tl_cap[slot]) {
				if (tl_buf[slot]) cudaFree(tl_buf[slot]);
				tl_buf[slot] = nullptr;
				tl_cap[slot] = 0;
				const size_t want = bytes < (1u << 16) ? (1u << 16) : bytes;
				CUDA_TRY(cudaMalloc(&tl_buf[slot], want));
				tl_cap[slot] = want;
			}
			scratch = tl_buf[slot];
		}
		double *p = (double *)scratch;
		double *hH = p; p += sH;
		double *hc = p; p += sc;
		double *hA = p; p += sA;
		double *hb = p; p += sb;
		double *hlb = p; p += sB;
		double *hub = p; p += sB;
		dsol = p; p += ssol;
		dstatus = (int32_t *)p;
		uint8_t *hbe = (uint8_t *)(dstatus + n);
		cudaError_t err = cudaMemcpyAsync(hH, H, sH * 8, cudaMemcpyHostToDevice, st);
		if (err == cudaSuccess) err = cudaMemcpyAsync(hc, c, sc * 8, cudaMemcpyHostToDevice, st);
		if (err == cudaSuccess && nc) err = cudaMemcpyAsync(hA, A, sA * 8, cudaMemcpyHostToDevice, st);
		if (err == cudaSuccess && nc) err = cudaMemcpyAsync(hb, b, sb * 8, cudaMemcpyHostToDevice, st);
		if (err == cudaSuccess) err = cudaMemcpyAsync(hlb, lb, sB * 8, cudaMemcpyHostToDevice, st);
		if (err == cudaSuccess) err = cudaMemcpyAsync(hub, ub, sB * 8, cudaMemcpyHostToDevice, st);
		if (err == cudaSuccess && be) err = cudaMemcpyAsync(hbe, be, nc, cudaMemcpyHostToDevice, st);
		if (err != cudaSuccess) {
			return fail(ASIF_ERR_CUDA, "qp_solve_batch H2D: %s", cudaGetErrorString(err));
		}
		dH = hH; dc = hc; dA = hA; db = hb; dlb = hlb; dub = hub;
		dbe = be ? hbe : nullptr;
	} else if (mem != ASIF_MEM_DEVICE) {
		return fail(ASIF_ERR_INVALID_ARGUMENT, "mem must be ASIF_MEM_HOST or ASIF_MEM_DEVICE");
	}
	int r;
	r = launch_qp_any(device, nv, n, nc, diagonal_cost, dH, dc, dA, db, dlb, dub, dbe, dsol, dstatus, share_flags, st, mem == ASIF_MEM_HOST);
	if (mem == ASIF_MEM_HOST) {
		cudaError_t err = cudaSuccess;
		if (!r) {
			err = cudaMemcpyAsync(sol, dsol, sizeof(double) * n * nv, cudaMemcpyDeviceToHost, st);
			if (err == cudaSuccess) err = cudaMemcpyAsync(status, dstatus, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st);
		}
		cudaError_t e2 = cudaStreamSynchronize(st);
		if (r) return r;
		if (err != cudaSuccess || e2 != cudaSuccess)
			return fail(ASIF_ERR_CUDA, "qp_solve_batch D2H: %s", cudaGetErrorString(err != cudaSuccess ? err : e2));
	}
	return r;
}

// -------------------------------------------------------------------------------------------------
namespace {
// 8 independent dependent-FMA chains per thread: enough ILP to saturate the FP64 pipe at any occupancy
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b)
{
	double r0 = threadIdx.x * 1e-3, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
	for (int i = 0; i < iters; i++) {
		r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
		r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
	}
	const double s = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
	if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
} // namespace

extern "C" int32_t asif_measure_fp64_peak(int32_t device, double *tflops, double *sm_clock_mhz_est)
{
	if (!tflops) return fail(ASIF_ERR_INVALID_ARGUMENT, "tflops is NULL");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(ASIF_ERR_NO_DEVICE, "no CUDA device");
	CUDA_TRY(cudaSetDevice(device));
	cudaDeviceProp prop;
	CUDA_TRY(cudaGetDeviceProperties(&prop, device));
	const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
	double *out = nullptr;
	CUDA_TRY(cudaMalloc(&out, sizeof(double) * blocks * threads));
	cudaEvent_t e0, e1;
	CUDA_TRY(cudaEventCreate(&e0));
	CUDA_TRY(cudaEventCreate(&e1));
	double best = 0.0;
	for (int rep = 0; rep < 5; rep++) {
		CUDA_TRY(cudaEventRecord(e0));
		fp64_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
		CUDA_TRY(cudaEventRecord(e1));
		CUDA_TRY(cudaEventSynchronize(e1));
		float ms = 0;
		CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
		const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
		const double tf = fl / (ms * 1e-3) / 1e12;
		if (rep > 0 && tf > best) best = tf;
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(out);
	*tflops = best;
	if (sm_clock_mhz_est) // 64 DFMA lanes per SM per clock on sm_100
		*sm_clock_mhz_est = best * 1e12 / (2.0 * 64.0 * prop.multiProcessorCount) / 1e6;
	return ASIF_OK;
}
