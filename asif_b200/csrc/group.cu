// group.cu -- one caller batch over several GPUs of one box (include/asif_b200.h, "engine groups").
// states are independent, so a batch is cut into contiguous slices [g*ceil(n/G), (g+1)*ceil(n/G)), one per
// device, with no collective anywhere: every device has its own engine (options and tables replicated), its own host
// worker thread and its own streams; each worker runs the ordinary single-device host-memory path on its slice of the
// CALLER's arrays, so the results land where the caller wants them and nothing is gathered afterwards.
// Host code only; no kernels in this unit.
#include "engine_internal.cuh"

#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

// One persistent host thread per device: CUDA calls for a device always come from the same thread (its context
// stays current there), and the devices' copies and launches are issued concurrently.
class DeviceWorker {
public:
	DeviceWorker() : th_([this]() { run(); }) {}
	~DeviceWorker()
	{
		{
			std::lock_guard<std::mutex> lk(m_);
			stop_ = true;
		}
		cv_.notify_all();
		th_.join();
	}
	DeviceWorker(const DeviceWorker &) = delete;
	DeviceWorker &operator=(const DeviceWorker &) = delete;

	void submit(std::function<int32_t()> job)
	{
		{
			std::lock_guard<std::mutex> lk(m_);
			job_ = std::move(job);
			has_job_ = true;
			done_ = false;
		}
		cv_.notify_all();
	}
	// result of the submitted job; err receives the worker thread's asif_last_error() text when it failed
	int32_t wait(std::string &err)
	{
		std::unique_lock<std::mutex> lk(m_);
		cv_done_.wait(lk, [this]() { return done_; });
		if (rc_ != ASIF_OK) err = err_;
		return rc_;
	}

private:
	void run()
	{
		for (;;) {
			std::function<int32_t()> job;
			{
				std::unique_lock<std::mutex> lk(m_);
				cv_.wait(lk, [this]() { return stop_ || has_job_; });
				if (stop_) return;
				job = std::move(job_);
				has_job_ = false;
			}
			int32_t rc;
			std::string err;
			try {
				rc = job();
				if (rc != ASIF_OK) err = asif_last_error(); // thread local: read it on the thread that set it
			} catch (const std::exception &ex) {
				rc = ASIF_ERR_INTERNAL;
				err = ex.what();
			} catch (...) {
				rc = ASIF_ERR_INTERNAL;
				err = "unknown exception in a device worker";
			}
			{
				std::lock_guard<std::mutex> lk(m_);
				rc_ = rc;
				err_ = err;
				done_ = true;
			}
			cv_done_.notify_all();
		}
	}
	std::mutex m_;
	std::condition_variable cv_, cv_done_;
	std::function<int32_t()> job_;
	bool has_job_ = false, done_ = true, stop_ = false;
	int32_t rc_ = ASIF_OK;
	std::string err_;
	std::thread th_; // last member: the thread starts after everything it touches exists
};

} // namespace

struct asif_engine_group {
	std::vector<asif_engine *> engines;
	std::vector<DeviceWorker *> workers;
	std::mutex call; // one batch at a time per group
};

namespace {

void slice_bounds(int64_t n, int G, int g, int64_t &lo, int64_t &hi)
{
	const int64_t per = (n + G - 1) / G;
	lo = per * g < n ? per * g : n;
	hi = lo + per < n ? lo + per : n;
}

// runs job(g, lo, hi) for every device with a non-empty slice, concurrently; first failure wins
int32_t for_each_slice(asif_engine_group *grp, int64_t n, const std::function<int32_t(int, int64_t, int64_t)> &job)
{
	std::lock_guard<std::mutex> lk(grp->call);
	const int G = (int)grp->engines.size();
	std::vector<char> used(G, 0);
	for (int g = 0; g < G; g++) {
		int64_t lo, hi;
		slice_bounds(n, G, g, lo, hi);
		if (hi <= lo) continue;
		used[g] = 1;
		grp->workers[g]->submit([=]() { return job(g, lo, hi); });
	}
	int32_t rc = ASIF_OK;
	std::string err, first;
	for (int g = 0; g < G; g++) {
		if (!used[g]) continue;
		const int32_t r = grp->workers[g]->wait(err);
		if (r != ASIF_OK && rc == ASIF_OK) {
			rc = r;
			first = "device slice " + std::to_string(g) + ": " + err;
		}
	}
	return rc == ASIF_OK ? ASIF_OK : fail(rc, "%s", first.c_str());
}

} // namespace

extern "C" {

int32_t asif_engine_group_create(const asif_engine_config *cfg, const int32_t *devices, int32_t n_devices, asif_engine_group **out)
{
	if (!cfg || !out) return fail(ASIF_ERR_INVALID_ARGUMENT, "cfg/out is NULL");
	*out = nullptr;
	const int32_t ndev = asif_device_count();
	if (ndev <= 0) return fail(ASIF_ERR_NO_DEVICE, "no CUDA device; this engine has no CPU fallback");
	std::vector<int32_t> devs;
	if (!devices || n_devices <= 0) {
		for (int32_t d = 0; d < ndev; d++) devs.push_back(d); // every visible device
	} else {
		for (int32_t i = 0; i < n_devices; i++) {
			if (devices[i] < 0 || devices[i] >= ndev) return fail(ASIF_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)", devices[i], ndev);
			for (int32_t j = 0; j < i; j++)
				if (devices[j] == devices[i]) return fail(ASIF_ERR_INVALID_ARGUMENT, "device %d listed twice", devices[i]);
			devs.push_back(devices[i]);
		}
	}
	asif_engine_group *grp = nullptr;
	try {
		grp = new asif_engine_group();
		const unsigned hw = std::thread::hardware_concurrency();
		for (size_t g = 0; g < devs.size(); g++) {
			asif_engine_config c = *cfg;
			c.device = devs[g];
			asif_engine *e = nullptr;
			const int32_t r = asif_engine_create(&c, &e);
			if (r != ASIF_OK) {
				asif_engine_group_destroy(grp);
				return r; // message already set by asif_engine_create on this thread
			}
			// pageable caller arrays: the slices' bounce copies share the box's cores
			int t = hw ? (int)(hw / devs.size()) : 2;
			e->copy_threads = t < 2 ? 2 : (t > 8 ? 8 : t);
			grp->engines.push_back(e);
			grp->workers.push_back(new DeviceWorker());
		}
	} catch (const std::exception &ex) {
		if (grp) asif_engine_group_destroy(grp);
		return fail(ASIF_ERR_INTERNAL, "group creation failed: %s", ex.what());
	}
	*out = grp;
	return ASIF_OK;
}

int32_t asif_engine_group_destroy(asif_engine_group *grp)
{
	if (!grp) return ASIF_OK;
	for (DeviceWorker *w : grp->workers) delete w; // joins
	for (asif_engine *e : grp->engines) asif_engine_destroy(e);
	delete grp;
	return ASIF_OK;
}

int32_t asif_engine_group_size(const asif_engine_group *grp)
{
	if (!grp) return fail(ASIF_ERR_INVALID_ARGUMENT, "group is NULL");
	return (int32_t)grp->engines.size();
}

asif_engine *asif_engine_group_engine(asif_engine_group *grp, int32_t i)
{
	if (!grp || i < 0 || i >= (int32_t)grp->engines.size()) {
		fail(ASIF_ERR_INVALID_ARGUMENT, "no such group member");
		return nullptr;
	}
	return grp->engines[i];
}

int32_t asif_engine_group_slice(const asif_engine_group *grp, int64_t n, int32_t i, int64_t bounds[2])
{
	if (!grp || !bounds || i < 0 || i >= (int32_t)grp->engines.size() || n < 0) return fail(ASIF_ERR_INVALID_ARGUMENT, "bad argument");
	slice_bounds(n, (int)grp->engines.size(), i, bounds[0], bounds[1]);
	return ASIF_OK;
}

int32_t asif_engine_group_filter_batch(asif_engine_group *grp, int64_t n, const double *x, const double *u_des, double *u_act,
                                       double *relax, int32_t *rc, double *diag)
{
	if (!grp) return fail(ASIF_ERR_INVALID_ARGUMENT, "group is NULL");
	if (n < 0) return fail(ASIF_ERR_INVALID_ARGUMENT, "n < 0");
	if (n == 0) return ASIF_OK;
	if (!x || !u_des || !u_act || !relax || !rc) return fail(ASIF_ERR_INVALID_ARGUMENT, "NULL batch pointer");
	return for_each_slice(grp, n, [=](int g, int64_t lo, int64_t hi) {
		asif_engine *e = grp->engines[g];
		return asif_engine_filter_batch(e, hi - lo, x + lo * e->nx, u_des + lo * e->nu, u_act + lo * e->nu, relax + lo * e->n_relax, rc + lo,
		                                diag ? diag + lo * e->n_diag : nullptr, ASIF_MEM_HOST, nullptr);
	});
}

int32_t asif_engine_group_filter_batch_cost(asif_engine_group *grp, int64_t n, const double *x, const double *H, const double *c,
                                            double *u_act, double *relax, int32_t *rc, double *diag)
{
	if (!grp) return fail(ASIF_ERR_INVALID_ARGUMENT, "group is NULL");
	if (n < 0) return fail(ASIF_ERR_INVALID_ARGUMENT, "n < 0");
	if (!x || !c || !u_act || !relax || !rc) return fail(ASIF_ERR_INVALID_ARGUMENT, "NULL batch pointer");
	if (H) // updateH is sticky: every member takes it, also those whose slice is empty
		for (asif_engine *e : grp->engines) {
			const int32_t r = asif_engine_set_input_cost(e, H);
			if (r != ASIF_OK) return r;
		}
	if (n == 0) return ASIF_OK;
	return for_each_slice(grp, n, [=](int g, int64_t lo, int64_t hi) {
		asif_engine *e = grp->engines[g];
		return asif_engine_filter_batch_cost(e, hi - lo, x + lo * e->nx, nullptr, c + lo * e->nv, u_act + lo * e->nu, relax + lo * e->n_relax,
		                                     rc + lo, diag ? diag + lo * e->n_diag : nullptr, ASIF_MEM_HOST, nullptr);
	});
}

int32_t asif_engine_group_rollout(asif_engine_group *grp, int64_t n, int32_t steps, double dt, double *x, const double *u_des,
                                  double *u_act_last, int32_t *rc_last, int64_t *rc_hist)
{
	if (!grp) return fail(ASIF_ERR_INVALID_ARGUMENT, "group is NULL");
	if (n < 0 || steps < 0) return fail(ASIF_ERR_INVALID_ARGUMENT, "n < 0 or steps < 0");
	if (rc_hist)
		for (int i = 0; i < 8; i++) rc_hist[i] = 0;
	if (n == 0) return ASIF_OK;
	if (!x || !u_des || !u_act_last || !rc_last) return fail(ASIF_ERR_INVALID_ARGUMENT, "NULL batch pointer");
	const int G = (int)grp->engines.size();
	std::vector<int64_t> hist((size_t)G * 8, 0);
	int64_t *hp = hist.data();
	const int32_t r = for_each_slice(grp, n, [=](int g, int64_t lo, int64_t hi) {
		asif_engine *e = grp->engines[g];
		return asif_engine_rollout(e, hi - lo, steps, dt, x + lo * e->nx, u_des + lo * e->nu, u_act_last + lo * e->nu, rc_last + lo,
		                           rc_hist ? hp + 8 * g : nullptr, ASIF_MEM_HOST, nullptr);
	});
	if (r == ASIF_OK && rc_hist)
		for (int g = 0; g < G; g++)
			for (int i = 0; i < 8; i++) rc_hist[i] += hist[(size_t)g * 8 + i];
	return r;
}

} // extern "C"
