// closed_loop.cu -- asif_engine_closed_loop: the example main loops for a fleet, state resident on the device.
// Host orchestration only: per control sample one filter launch (any filter class) and one
// closed_loop_step_kernel launch, all on one stream, nothing synchronised until the end.
#define ASIF_QP_POLISH_INLINE_ROWS 1 // qp_gi.cuh: the form of the vertex polish that costs this unit's kernels least
#include "closed_loop_kernel.cuh"
#include "engine_internal.cuh"

#include <cstring>

using namespace asifb;

namespace {

struct DevBuf {
	void *p = nullptr;
	~DevBuf() { cudaFree(p); }
	cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 8); }
	template <class T>
	T *as() const { return static_cast<T *>(p); }
};

int log_width(const asif_engine *e) { return 1 + 2 * e->nx + 3 * e->nu + e->n_relax + 1 + 2 + 3; }

int64_t log_records(const asif_loop_config *c) { return c->log_stride > 0 ? (c->steps + c->log_stride - 1) / c->log_stride : 0; }

template <class P>
int launch_step(const LoopParams &lp, int64_t n, double *x, const double *xe, const double *ud, const double *uf, double *uh,
                double *ua, const double *relax, const int32_t *rc, double *smooth, const double *diag, double *log,
                unsigned long long *hist, cudaStream_t st)
{
	const unsigned blocks = (unsigned)((n + LOOP_THREADS - 1) / LOOP_THREADS);
	closed_loop_step_kernel<P><<<blocks, LOOP_THREADS, 0, st>>>(lp, n, x, xe, ud, uf, uh, ua, relax, rc, smooth, diag, log, hist);
	CUDA_TRY(cudaGetLastError());
	return ASIF_OK;
}

int dispatch_step(const asif_engine *e, const LoopParams &lp, int64_t n, double *x, const double *xe, const double *ud,
                  const double *uf, double *uh, double *ua, const double *relax, const int32_t *rc, double *smooth, const double *diag,
                  double *log, unsigned long long *hist, cudaStream_t st)
{
	switch (e->cfg.model) {
	case ASIF_MODEL_DOUBLE_INTEGRATOR:
		return launch_step<ModelPlant<DoubleIntegratorExplicit>>(lp, n, x, xe, ud, uf, uh, ua, relax, rc, smooth, diag, log, hist, st);
	case ASIF_MODEL_DOUBLE_INTEGRATOR_TB:
		return launch_step<ModelPlant<DoubleIntegratorTB>>(lp, n, x, xe, ud, uf, uh, ua, relax, rc, smooth, diag, log, hist, st);
	case ASIF_MODEL_INVERTED_PENDULUM:
		return launch_step<ModelPlant<InvertedPendulumImplicit>>(lp, n, x, xe, ud, uf, uh, ua, relax, rc, smooth, diag, log, hist, st);
	case ASIF_MODEL_INVERTED_PENDULUM_TABLE:
	case ASIF_MODEL_INVERTED_PENDULUM_KERNEL:
		return launch_step<PendulumExactPlant>(lp, n, x, xe, ud, uf, uh, ua, relax, rc, smooth, diag, log, hist, st);
	case ASIF_MODEL_SEGWAY:
	case ASIF_MODEL_SEGWAY_SHIPPED: // same plant, the two differ in the backup set only
		return launch_step<ModelPlant<SegwayTB<true>>>(lp, n, x, xe, ud, uf, uh, ua, relax, rc, smooth, diag, log, hist, st);
	}
	return fail(ASIF_ERR_UNSUPPORTED, "closed loop: no plant for model %d", e->cfg.model);
}

} // namespace

extern "C" {

int32_t asif_loop_config_init(asif_loop_config *c)
{
	if (!c) return fail(ASIF_ERR_INVALID_ARGUMENT, "cfg is NULL");
	memset(c, 0, sizeof(*c));
	c->struct_size = (uint32_t)sizeof(*c);
	c->steps = 0;
	c->dt = 1e-3;
	c->steps_per_sample = 1;
	c->smooth_lb = -20.0;
	c->smooth_ub = 20.0;
	c->smooth_rate = 20.0 * 1 * 0.001;
	c->plant_gain = 1.0;
	c->log_after_step = 1;
	return ASIF_OK;
}

int32_t asif_engine_loop_log_dims(const asif_engine *e, const asif_loop_config *c, int64_t dims[2])
{
	if (!e || !c || !dims) return fail(ASIF_ERR_INVALID_ARGUMENT, "NULL argument");
	dims[0] = log_width(e);
	dims[1] = log_records(c);
	return ASIF_OK;
}

int32_t asif_engine_closed_loop(asif_engine *e, int64_t n, const asif_loop_config *c, double *x, const double *u_des,
                                double *u_act_last, double *relax_last, int32_t *rc_last, int64_t *rc_hist, double *log,
                                int32_t mem, void *stream)
{
	if (!e || !c) return fail(ASIF_ERR_INVALID_ARGUMENT, "engine or loop config is NULL");
	if (c->struct_size != sizeof(asif_loop_config)) return fail(ASIF_ERR_INVALID_ARGUMENT, "asif_loop_config: struct_size mismatch (use asif_loop_config_init)");
	if (n < 0 || c->steps < 0 || c->steps_per_sample < 1 || !(c->dt > 0.0)) return fail(ASIF_ERR_INVALID_ARGUMENT, "closed loop: bad n / steps / steps_per_sample / dt");
	if (c->log_stride < 0 || c->log_agents < 0) return fail(ASIF_ERR_INVALID_ARGUMENT, "closed loop: negative log_stride or log_agents");
	if (n == 0) return ASIF_OK;
	if (!x || !u_des || !u_act_last || !relax_last || !rc_last) return fail(ASIF_ERR_INVALID_ARGUMENT, "NULL batch pointer");
	if (mem != ASIF_MEM_HOST && mem != ASIF_MEM_DEVICE) return fail(ASIF_ERR_INVALID_ARGUMENT, "mem must be ASIF_MEM_HOST or ASIF_MEM_DEVICE");
	const int64_t n_log = c->log_stride > 0 ? (c->log_agents < n ? c->log_agents : n) : 0;
	const int64_t n_rec = log_records(c);
	if (n_log > 0 && !log) return fail(ASIF_ERR_INVALID_ARGUMENT, "closed loop: log buffer is NULL");
	CUDA_TRY(cudaSetDevice(e->cfg.device));
	ServerPause pause(e); // scratch buffers are allocated and freed per call
	const int nx = e->nx, nu = e->nu, nr = e->n_relax, W = log_width(e);
	const bool host = mem == ASIF_MEM_HOST;
	cudaStream_t st = (!host && stream) ? (cudaStream_t)stream : e->stream;
	const bool tb = e->cfg.filter == ASIF_FILTER_IMPLICIT_TB;

	DevBuf bx, bud, bua, brl, brc, bxe, buf, bufh, bsm, bdg, blog, bhist;
	CUDA_TRY(bxe.alloc(sizeof(double) * n * nx));
	CUDA_TRY(buf.alloc(sizeof(double) * n * nu));
	CUDA_TRY(bufh.alloc(sizeof(double) * n * nu));
	CUDA_TRY(bsm.alloc(sizeof(double) * n * 2));
	CUDA_TRY(bhist.alloc(sizeof(unsigned long long) * 8));
	if (tb && n_log > 0) CUDA_TRY(bdg.alloc(sizeof(double) * n_log * e->n_diag));
	double *dx = x, *dua = u_act_last, *drl = relax_last, *dlog = log;
	const double *dud = u_des;
	int32_t *drc = rc_last;
	if (host) {
		CUDA_TRY(bx.alloc(sizeof(double) * n * nx));
		CUDA_TRY(bud.alloc(sizeof(double) * n * nu));
		CUDA_TRY(bua.alloc(sizeof(double) * n * nu));
		CUDA_TRY(brl.alloc(sizeof(double) * n * nr));
		CUDA_TRY(brc.alloc(sizeof(int32_t) * n));
		if (n_log > 0) CUDA_TRY(blog.alloc(sizeof(double) * n_log * n_rec * W));
		dx = bx.as<double>();
		dua = bua.as<double>();
		drl = brl.as<double>();
		drc = brc.as<int32_t>();
		dlog = blog.as<double>();
		CUDA_TRY(cudaMemcpyAsync(dx, x, sizeof(double) * n * nx, cudaMemcpyHostToDevice, st));
		CUDA_TRY(cudaMemcpyAsync(bud.p, u_des, sizeof(double) * n * nu, cudaMemcpyHostToDevice, st));
		dud = bud.as<double>();
	}
	if (n_log > 0) CUDA_TRY(cudaMemsetAsync(dlog, 0, sizeof(double) * n_log * n_rec * W, st));
	CUDA_TRY(cudaMemsetAsync(bhist.p, 0, sizeof(unsigned long long) * 8, st));
	CUDA_TRY(cudaMemsetAsync(bufh.p, 0, sizeof(double) * n * nu, st)); // uActNow = {0.0} before the first call
	CUDA_TRY(cudaMemsetAsync(e->d_counters, 0, N_COUNTERS * sizeof(unsigned long long), st));
	// smoothBounds start at the rate limiter's own limits (examples/DoubleIntegrator_RealizableSampled.cpp:107)
	fill_pairs_kernel<<<(unsigned)((n + LOOP_THREADS - 1) / LOOP_THREADS), LOOP_THREADS, 0, st>>>(bsm.as<double>(), n, c->smooth_lb, c->smooth_ub);
	CUDA_TRY(cudaGetLastError());

	LoopParams lp;
	memset(&lp, 0, sizeof(lp));
	lp.dt = c->dt;
	lp.smooth = c->smooth_bounds ? 1 : 0;
	lp.hold_on_failure = (e->cfg.filter == ASIF_FILTER_EXPLICIT || e->cfg.filter == ASIF_FILTER_ROBUST || e->cfg.filter == ASIF_FILTER_REALIZABLE) ? 1 : 0;
	lp.log_stride = n_log > 0 ? c->log_stride : 0;
	lp.log_after = c->log_after_step ? 1 : 0;
	lp.n_relax = nr;
	lp.n_diag = e->n_diag;
	lp.tb_diag = (tb && n_log > 0) ? 1 : 0;
	lp.log_width = W;
	lp.log_records = (int32_t)n_rec;
	lp.log_agents = n_log;
	lp.smooth_lb = c->smooth_lb;
	lp.smooth_ub = c->smooth_ub;
	lp.smooth_rate = c->smooth_rate;
	lp.plant_gain = c->plant_gain;
	double t = 0.0; // tNow accumulates += dt as in the examples
	for (int32_t s = 0; s < c->steps; s += c->steps_per_sample) {
		const int32_t k = (c->steps - s) < c->steps_per_sample ? (c->steps - s) : c->steps_per_sample;
		CUDA_TRY(cudaMemcpyAsync(bxe.p, dx, sizeof(double) * n * nx, cudaMemcpyDeviceToDevice, st)); // xEstim = xNow
		const double *xe = bxe.as<double>();
		int r;
		if (lp.tb_diag) { // the logged agents also produce the TB diagnostics (TTS_, BTorthoBS_, backTrajCritIdx_)
			r = launch_filter_any(e, n_log, xe, dud, buf.as<double>(), drl, drc, bdg.as<double>(), st);
			if (!r && n > n_log)
				r = launch_filter_any(e, n - n_log, xe + n_log * nx, dud + n_log * nu, buf.as<double>() + n_log * nu, drl + n_log * nr,
				                      drc + n_log, nullptr, st);
		} else {
			r = launch_filter_any(e, n, xe, dud, buf.as<double>(), drl, drc, nullptr, st);
		}
		if (r) return r;
		lp.t0 = t;
		lp.k = k;
		lp.step0 = s;
		r = dispatch_step(e, lp, n, dx, xe, dud, buf.as<double>(), bufh.as<double>(), dua, drl, drc, bsm.as<double>(), bdg.as<double>(), dlog,
		                  bhist.as<unsigned long long>(), st);
		if (r) return r;
		for (int32_t j = 0; j < k; j++) t += c->dt;
	}
	if (host) {
		CUDA_TRY(cudaMemcpyAsync(x, dx, sizeof(double) * n * nx, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaMemcpyAsync(u_act_last, dua, sizeof(double) * n * nu, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaMemcpyAsync(relax_last, drl, sizeof(double) * n * nr, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaMemcpyAsync(rc_last, drc, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
		if (n_log > 0) CUDA_TRY(cudaMemcpyAsync(log, dlog, sizeof(double) * n_log * n_rec * W, cudaMemcpyDeviceToHost, st));
	}
	// the scratch buffers are freed on return, so this call always completes before it returns
	CUDA_TRY(cudaStreamSynchronize(st));
	if (rc_hist) {
		unsigned long long h[8];
		CUDA_TRY(cudaMemcpy(h, bhist.p, sizeof(h), cudaMemcpyDeviceToHost));
		for (int i = 0; i < 8; i++) rc_hist[i] = (int64_t)h[i];
	}
	return ASIF_OK;
}

} // extern "C"
