// asif_b200.hpp -- C++ host layer over the C ABI (include/asif_b200.h); header only.
//
// Mirrors the reference's C++ surface for the filter path so that it drops in:
//   * ASIF::QPWrapperAbstract            the solver interface (include/qpwrapper_abstract.h:16-51);
//                                        declared here only when the reference header is absent
//   * ASIF::QPWrapperB200                a QPWrapperAbstract backend on the B200 kernels, the
//                                        sibling of QPWrapperOsqp (src/qpwrapper_osqp.cpp)
//   * ASIF::b200::FilterBatchImplicitTB / FilterBatchImplicit / FilterBatchExplicit / FilterBatchRobust /
//     FilterBatchRealizable
//                                        the batched counterparts of ASIFimplicitTB / ASIF with the
//                                        same Options structs, initialize / updateOptions semantics
//                                        and return codes, plus filterBatch(n, X, UDes, UAct, Relax, rc)
// Models are device functors selected by id (the reference passes std::function callbacks, which
// cannot run on the device); asif_b200/csrc/models.cuh keeps their five callback signatures.
// Everything computes on the GPU; there is no CPU fallback - errors surface as negative return
// values with the message available from lastError().
#ifndef ASIF_B200_HPP
#define ASIF_B200_HPP

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/asif_b200.h"

namespace ASIF
{

#ifndef _QPWRAPPER_ABSTRACT_H_
#define _QPWRAPPER_ABSTRACT_H_
	// min (x'Hx + c'x)  s.t.  A x >= b,  lb <= x <= ub      (no 1/2; dense column-major A[i + j*nc])
	class QPWrapperAbstract
	{
	public:
		enum class SOLVER_STATUS : int32_t { INFEASIBLE = 0, FEASIBLE = 1 };

		QPWrapperAbstract(const uint32_t nv, const uint32_t nc, const bool diagonalCost)
		    : nv_(nv), nc_(nc), diagonalCost_(diagonalCost)
		{
			be_ = new bool[nc];
			for (uint32_t i = 0; i < nc_; i++) be_[i] = false;
		}
		virtual ~QPWrapperAbstract(void) { delete[] be_; }

		virtual int32_t initialize(const double H[], const double c[], const double A[], const double b[],
		                           const double lb[], const double ub[], const bool be[] = nullptr) = 0;
		virtual int32_t updateCost(const double H[], const double c[]) = 0;
		virtual int32_t updateA(const double A[]) = 0;
		virtual int32_t updateb(const double b[]) = 0;
		virtual int32_t updateBounds(const double lb[], const double ub[]) = 0;
		virtual int32_t solve(void) = 0;
		virtual int32_t getSolution(double sol[]) = 0;

	protected:
		const uint32_t nv_;
		const uint32_t nc_;
		const bool diagonalCost_;
		bool *be_;
	};
#endif

	namespace b200
	{
		// Pinned, device-addressable host array for the filterBatch arguments (asif_host_alloc): with such arrays the
		// engine overlaps transfers with the kernels or lets the kernels read and write them in place.  Move-only.
		template <class T>
		class PinnedBuffer
		{
		public:
			PinnedBuffer(void) : p_(nullptr), n_(0) {}
			explicit PinnedBuffer(const size_t n) : p_(nullptr), n_(0) { resize(n); }
			PinnedBuffer(const PinnedBuffer &) = delete;
			PinnedBuffer &operator=(const PinnedBuffer &) = delete;
			PinnedBuffer(PinnedBuffer &&o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr, o.n_ = 0; }
			PinnedBuffer &operator=(PinnedBuffer &&o) noexcept
			{
				if (this != &o) {
					release();
					p_ = o.p_, n_ = o.n_;
					o.p_ = nullptr, o.n_ = 0;
				}
				return *this;
			}
			~PinnedBuffer(void) { release(); }
			// returns 0, or the negative code of asif_host_alloc (contents are not kept)
			int32_t resize(const size_t n)
			{
				release();
				void *q = nullptr;
				const int32_t r = asif_host_alloc(&q, (uint64_t)n * sizeof(T));
				if (r == 0) p_ = static_cast<T *>(q), n_ = n;
				return r;
			}
			T *data(void) { return p_; }
			const T *data(void) const { return p_; }
			size_t size(void) const { return n_; }
			T &operator[](const size_t i) { return p_[i]; }
			const T &operator[](const size_t i) const { return p_[i]; }

		private:
			void release(void)
			{
				if (p_) asif_host_free(p_);
				p_ = nullptr, n_ = 0;
			}
			T *p_;
			size_t n_;
		};
	} // namespace b200

	// A QPWrapperAbstract whose solve() runs on the B200 (asif_qp_solve_batch with n = 1).
	// Argument meaning, copies and return values follow QPWrapperOsqp (src/qpwrapper_osqp.cpp:55-261):
	// initialize() returns 0 on success (the polarity of osqp_setup, :121), the update*() calls return 1,
	// solve() returns 1 (FEASIBLE) or the OSQP status code the reference would pass through (:225-238),
	// getSolution() copies nv values.  nullptr arguments of updateCost / updateBounds mean "unchanged".
	class QPWrapperB200 : public QPWrapperAbstract
	{
	public:
		QPWrapperB200(const uint32_t nv, const uint32_t nc, const bool diagonalCost, const int32_t device = 0)
		    : QPWrapperAbstract(nv, nc, diagonalCost), device_(device), H_(nv * nv, 0.0), c_(nv, 0.0), A_((size_t)nv * nc, 0.0),
		      b_(nc, 0.0), lb_(nv, 0.0), ub_(nv, 0.0), sol_(nv, 0.0), beU8_(nc, 0), hasEq_(false), status_(-10)
		{
		}
		virtual ~QPWrapperB200(void) {}

		virtual int32_t initialize(const double H[], const double c[], const double A[], const double b[],
		                           const double lb[], const double ub[], const bool be[] = nullptr)
		{
			// nv <= 4 runs the per-thread exact dual active-set solver; nv > 4 (the LP-dual formulations of ASIFrobust /
			// ASIFrealizable: nv = 402 / 38, Hessian only semi-definite) the cluster-cooperative operator-splitting solver
			// with polish (csrc/qp_admm.cuh) - asif_qp_solve_batch picks by nv.
			if (nv_ < 1) return ASIF_ERR_UNSUPPORTED;
			if (be != nullptr) {
				for (uint32_t i = 0; i < nc_; i++) {
					be_[i] = be[i];
					beU8_[i] = be[i] ? 1 : 0;
					hasEq_ = hasEq_ || be[i];
				}
			}
			std::memcpy(H_.data(), H, sizeof(double) * nv_ * nv_);
			std::memcpy(c_.data(), c, sizeof(double) * nv_);
			std::memcpy(A_.data(), A, sizeof(double) * nv_ * nc_);
			std::memcpy(b_.data(), b, sizeof(double) * nc_);
			std::memcpy(lb_.data(), lb, sizeof(double) * nv_);
			std::memcpy(ub_.data(), ub, sizeof(double) * nv_);
			return asif_device_count() > 0 ? 0 : ASIF_ERR_NO_DEVICE;
		}
		virtual int32_t updateCost(const double H[], const double c[])
		{
			if (H != nullptr) std::memcpy(H_.data(), H, sizeof(double) * nv_ * nv_);
			if (c != nullptr) std::memcpy(c_.data(), c, sizeof(double) * nv_);
			return 1;
		}
		virtual int32_t updateA(const double A[])
		{
			std::memcpy(A_.data(), A, sizeof(double) * nv_ * nc_);
			return 1;
		}
		virtual int32_t updateb(const double b[])
		{
			std::memcpy(b_.data(), b, sizeof(double) * nc_);
			return 1;
		}
		virtual int32_t updateBounds(const double lb[], const double ub[])
		{
			if (lb != nullptr) std::memcpy(lb_.data(), lb, sizeof(double) * nv_);
			if (ub != nullptr) std::memcpy(ub_.data(), ub, sizeof(double) * nv_);
			return 1;
		}
		virtual int32_t solve(void)
		{
			int32_t st = -10; // OSQP_UNSOLVED
			const int32_t r = asif_qp_solve_batch(device_, (int32_t)nv_, (int32_t)nc_, 1, diagonalCost_ ? 1 : 0, H_.data(),
			                                      c_.data(), A_.data(), b_.data(), lb_.data(), ub_.data(),
			                                      hasEq_ ? beU8_.data() : nullptr, sol_.data(), &st,
			                                      ASIF_QP_SHARED_H | ASIF_QP_SHARED_BOUNDS, ASIF_MEM_HOST, nullptr);
			status_ = (r == ASIF_OK) ? st : r;
			return status_;
		}
		virtual int32_t getSolution(double sol[])
		{
			for (uint32_t i = 0; i < nv_; i++) sol[i] = sol_[i];
			return 1;
		}

	protected:
		int32_t device_;
		std::vector<double> H_, c_, A_, b_, lb_, ub_, sol_;
		std::vector<uint8_t> beU8_;
		bool hasEq_;
		int32_t status_;
	};

	namespace b200
	{
		enum class Model : int32_t {
			DoubleIntegrator = ASIF_MODEL_DOUBLE_INTEGRATOR,
			DoubleIntegratorTB = ASIF_MODEL_DOUBLE_INTEGRATOR_TB,
			InvertedPendulum = ASIF_MODEL_INVERTED_PENDULUM,
			Segway = ASIF_MODEL_SEGWAY,
			SegwayShipped = ASIF_MODEL_SEGWAY_SHIPPED,
			InvertedPendulumTable = ASIF_MODEL_INVERTED_PENDULUM_TABLE,
			InvertedPendulumKernel = ASIF_MODEL_INVERTED_PENDULUM_KERNEL
		};

		inline std::string lastError(void) { return std::string(asif_last_error()); }

		// ASIF::LearningData (include/asif_learning_utils.h:8-32): same fields (without the two output pointers Lfh_diff /
		// Lgh_diff, which the reference allocates anew on every call and never frees, :145-146)
		typedef asif_learning_data LearningData;

		// common plumbing of the batched filters
		class FilterBatchBase
		{
		public:
			virtual ~FilterBatchBase(void) { release(); }
			// Device list: call before initialize().  With more than one device every host batch is cut into
			// contiguous slices, one per GPU (asif_engine_group_*), and the results land in the caller's arrays; an empty
			// list means "every visible device".  Without this call the filter runs on the constructor's single device.
			void setDevices(const std::vector<int32_t> &devices)
			{
				devices_ = devices;
				useGroup_ = true;
			}
			// Latency server for the single-state filter() calls (asif_engine_latency_server): call after initialize();
			// returns 0, or ASIF_ERR_UNSUPPORTED for the classes that keep the launch path.  updateOptions() rebuilds the
			// engine: switch it on again afterwards.
			int32_t setLowLatency(const bool on)
			{
				if (engine_ == nullptr) return ASIF_ERR_INVALID_ARGUMENT;
				return asif_engine_latency_server(engine_, on ? 1 : 0);
			}
			uint32_t nDevices(void) const { return group_ ? (uint32_t)asif_engine_group_size(group_) : (engine_ ? 1u : 0u); }
			// filter(x, uDes, uAct, relax) of the reference class on n states; host arrays
			// X[n*nx], UDes[n*nu], UAct[n*nu], Relax[n*nRelax], rc[n]; returns 0 or a negative ASIF_ERR_* (<= -101)
			int32_t filterBatch(const int64_t n, const double X[], const double UDes[], double UAct[], double Relax[],
			                    int32_t rc[], double diag[] = nullptr)
			{
				if (engine_ == nullptr) return ASIF_ERR_INVALID_ARGUMENT;
				if (group_) return asif_engine_group_filter_batch(group_, n, X, UDes, UAct, Relax, rc, diag);
				return asif_engine_filter_batch(engine_, n, X, UDes, UAct, Relax, rc, diag, ASIF_MEM_HOST, nullptr);
			}
			// the filter(x, H, c, uAct, relax) overloads on n states: C[n*nv] is the caller's linear cost of the whole decision
			// vector, H (nu x nu, column-major) replaces the input block of the Hessian and stays in force (updateH), or nullptr
			int32_t filterBatch(const int64_t n, const double X[], const double H[], const double C[], double UAct[], double Relax[],
			                    int32_t rc[], double diag[] = nullptr)
			{
				if (engine_ == nullptr) return ASIF_ERR_INVALID_ARGUMENT;
				if (group_) return asif_engine_group_filter_batch_cost(group_, n, X, H, C, UAct, Relax, rc, diag);
				return asif_engine_filter_batch_cost(engine_, n, X, H, C, UAct, Relax, rc, diag, ASIF_MEM_HOST, nullptr);
			}
			// single state, as every reference class declares it (e.g. include/asif_implicit_tb.h:103-110); relax has nRelax() entries.
			// Returns the reference's code, or an ASIF_ERR_* (<= -101, outside the reference's range) when the engine itself
			// failed - uAct and relax are then left untouched.  The classes that leave uAct untouched when the QP fails
			// (ASIF, ASIFrobust, ASIFrealizable: src/asif.cpp:207-209) do so here too.
			int32_t filter(const double x[], const double H[], const double c[], double uAct[], double relax[])
			{
				int32_t rc = 0;
				double ua[2] = {0.0, 0.0}, rl[2] = {0.0, 0.0};
				const int32_t r = filterBatch(1, x, H, c, ua, rl, &rc);
				if (r != ASIF_OK) return r;
				if (rc == 1 || !keepOnFailure_) {
					for (uint32_t i = 0; i < nu(); i++) uAct[i] = ua[i];
					for (uint32_t i = 0; i < nRelax(); i++) relax[i] = rl[i];
				}
				return rc;
			}
			// same with device pointers on the caller's CUDA stream (cudaStream_t); returns without synchronising
			int32_t filterBatchDevice(const int64_t n, const double *X, const double *UDes, double *UAct, double *Relax,
			                          int32_t *rc, void *stream, double *diag = nullptr)
			{
				if (engine_ == nullptr) return ASIF_ERR_INVALID_ARGUMENT;
				return asif_engine_filter_batch(engine_, n, X, UDes, UAct, Relax, rc, diag, ASIF_MEM_DEVICE, stream);
			}
			// The loop around filter() of the example programs for n agents (asif_engine_closed_loop): X is advanced in
			// place by loop.steps plant steps; see asif_loop_config for sampling, rate limiter and log options.
			// log: logRecordWidth() * logRecords(loop) doubles per recorded agent, or nullptr when loop.log_stride == 0.
			int32_t closedLoop(const int64_t n, const asif_loop_config &loop, double X[], const double UDes[], double UActLast[],
			                   double RelaxLast[], int32_t rcLast[], int64_t rcHist[8] = nullptr, double log[] = nullptr)
			{
				if (engine_ == nullptr) return ASIF_ERR_INVALID_ARGUMENT;
				return asif_engine_closed_loop(engine_, n, &loop, X, UDes, UActLast, RelaxLast, rcLast, rcHist, log, ASIF_MEM_HOST,
				                               nullptr);
			}
			static asif_loop_config loopDefaults(void)
			{
				asif_loop_config c;
				asif_loop_config_init(&c);
				return c;
			}
			int64_t logRecordWidth(void) const
			{
				asif_loop_config c = loopDefaults();
				int64_t d[2] = {0, 0};
				asif_engine_loop_log_dims(engine_, &c, d);
				return d[0];
			}
			int64_t logRecords(const asif_loop_config &loop) const
			{
				int64_t d[2] = {0, 0};
				asif_engine_loop_log_dims(engine_, &loop, d);
				return d[1];
			}
			uint32_t nx(void) const { return dims_[0]; }
			uint32_t nu(void) const { return dims_[1]; }
			uint32_t nRelax(void) const { return dims_[2]; }
			uint32_t nDiag(void) const { return dims_[5]; }

		protected:
			FilterBatchBase(void) : engine_(nullptr), group_(nullptr), useGroup_(false), keepOnFailure_(false) { std::memset(dims_, 0, sizeof(dims_)); }
			void release(void)
			{
				if (group_) asif_engine_group_destroy(group_); // owns its engines
				else asif_engine_destroy(engine_);
				group_ = nullptr;
				engine_ = nullptr;
			}
			int32_t create(const asif_engine_config &cfg)
			{
				if (useGroup_) {
					asif_engine_group *g = nullptr;
					const int32_t r = asif_engine_group_create(&cfg, devices_.empty() ? nullptr : devices_.data(), (int32_t)devices_.size(), &g);
					if (r != ASIF_OK) return r;
					release();
					group_ = g;
					engine_ = asif_engine_group_engine(group_, 0); // borrowed: dims, single-state calls, device-pointer calls
				} else {
					asif_engine *e = nullptr;
					const int32_t r = asif_engine_create(&cfg, &e);
					if (r != ASIF_OK) return r;
					release();
					engine_ = e;
				}
				asif_engine_dims(engine_, dims_);
				return 1;
			}
			// single-state filter(x, uDes, ...) shared by the classes: locals first, so that an engine error leaves the caller's
			// uAct / relax untouched (an ASIF_ERR_* must not read as "backup action is in uAct")
			int32_t filterOne(const double x[], const double uDes[], double uAct[], double relax[])
			{
				int32_t rc = 0;
				double ua[2] = {0.0, 0.0}, rl[2] = {0.0, 0.0};
				const int32_t r = filterBatch(1, x, uDes, ua, rl, &rc);
				if (r != ASIF_OK) return r;
				if (rc == 1 || !keepOnFailure_) {
					for (uint32_t i = 0; i < nu(); i++) uAct[i] = ua[i];
					for (uint32_t i = 0; i < nRelax(); i++) relax[i] = rl[i];
				}
				return rc;
			}
			asif_engine *engine_;
			asif_engine_group *group_;
			std::vector<int32_t> devices_;
			bool useGroup_;
			bool keepOnFailure_; // ASIF / ASIFrobust / ASIFrealizable: uAct, relax untouched unless rc == 1
			int32_t dims_[6];
		};

		// Batched ASIF::ASIF (include/asif.h): explicit CBF filter.
		class FilterBatchExplicit : public FilterBatchBase
		{
		public:
			using FilterBatchBase::filter; // filter(x, H, c, uAct, relax)
			typedef struct {
				double relaxLb = 5.0;
				double relaxCost = 50.0;
				double satSharpness = 5.0;
				double inf = 1e20;
			} Options; // include/asif.h:11-17

			// npSSmax as in the ASIF::ASIF constructor (include/asif.h:39-44): rows kept after sorting by h, default all
			explicit FilterBatchExplicit(const Model model, const int32_t device = 0, const uint32_t npSSmax = (uint32_t)-1)
			    : model_(model), device_(device), npSSmax_(npSSmax)
			{
				keepOnFailure_ = true;
			}
			int32_t initialize(const double lb[], const double ub[]) { return initialize(lb, ub, Options()); }
			int32_t initialize(const double lb[], const double ub[], const Options &options)
			{
				options_ = options;
				asif_engine_config cfg;
				int32_t r = asif_engine_config_init(&cfg, ASIF_FILTER_EXPLICIT, (int32_t)model_);
				if (r != ASIF_OK) return r;
				cfg.device = device_;
				cfg.lb[0] = lb[0];
				cfg.ub[0] = ub[0];
				lb_ = lb[0];
				ub_ = ub[0];
				cfg.relaxLb = options.relaxLb;
				cfg.relaxCost = options.relaxCost;
				cfg.inf = options.inf;
				cfg.npSSmax = (npSSmax_ > 0x7fffffffu) ? 0 : (int32_t)npSSmax_;
				return create(cfg);
			}
			// src/asif.cpp:213-231 moves only the LOWER bound of the relax variable and leaves the upper
			// bound at the value of initialize(): lowering relaxLb frees the variable on [new, old], raising
			// it makes OSQP reject the bound update (l > u) so the old pin stays.  That asymmetry is an
			// accident of the wrapper; the batched engine pins both bounds to the new value (documented
			// deviation, DESIGN.md "quirks").
			int32_t updateOptions(const Options &options)
			{
				const double lb[1] = {lb_}, ub[1] = {ub_};
				return initialize(lb, ub, options);
			}
			// filter(x, uDes, uAct, Lfh, Lgh, relax) (include/asif.h:43-58): caller-supplied Lie derivatives; for this call only
			int32_t filter(const double x[], const double uDes[], double uAct[], double Lfh[], double Lgh[], double &relax)
			{
				int32_t rc = 0;
				if (engine_ == nullptr) return ASIF_ERR_INVALID_ARGUMENT;
				double ua[2] = {0.0, 0.0}, rl = 0.0;
				const int32_t r = asif_engine_filter_batch_lie(engine_, 1, x, uDes, Lfh, Lgh, ua, &rl, &rc, nullptr, ASIF_MEM_HOST, nullptr);
				if (r != ASIF_OK) return r;
				if (rc == 1) { // untouched on failure (src/asif.cpp:207-209)
					for (uint32_t i = 0; i < nu(); i++) uAct[i] = ua[i];
					relax = rl;
				}
				return rc;
			}
			int32_t filterBatchLie(const int64_t n, const double X[], const double UDes[], const double Lfh[], const double Lgh[],
			                       double UAct[], double Relax[], int32_t rc[], double diag[] = nullptr)
			{
				if (engine_ == nullptr) return ASIF_ERR_INVALID_ARGUMENT;
				return asif_engine_filter_batch_lie(engine_, n, X, UDes, Lfh, Lgh, UAct, Relax, rc, diag, ASIF_MEM_HOST, nullptr);
			}
			// the reference leaves uAct / relax untouched on failure (src/asif.cpp:207-209)
			int32_t filter(const double x[], const double uDes[], double uAct[], double &relax) { return filterOne(x, uDes, uAct, &relax); }

		protected:
			Model model_;
			int32_t device_;
			uint32_t npSSmax_;
			Options options_;
			double lb_, ub_;
		};

		// Batched ASIF::ASIFimplicitTB (include/asif_implicit_tb.h).
		class FilterBatchImplicitTB : public FilterBatchBase
		{
		public:
			using FilterBatchBase::filter; // filter(x, H, c, uAct, relax)
			typedef struct {
				double relaxCost = 50.0;
				double relaxSafeLb = 5.0;
				double relaxTTS = 5.0;
				double relaxMinOrtho = 5.0;
				double backTrajHorizon = 1.0;
				double backTrajExtend = 0.05;
				double backTrajDt = 0.01;
				double backTrajMinOrtho = 0.01;
				double backTrajAbsTol = 1.0e-6;
				double backTrajRelTol = 1.0e-6;
				double satSharpness = 0.1;
				double inf = 1e20;
			} Options; // include/asif_implicit_tb.h:19-33

			FilterBatchImplicitTB(const Model model, const uint32_t npBTSS = 4, const int32_t device = 0)
			    : model_(model), npBTSS_(npBTSS), device_(device)
			{
			}
			int32_t initialize(const double lb[], const double ub[]) { return initialize(lb, ub, Options()); }
			int32_t initialize(const double lb[], const double ub[], const Options &options)
			{
				options_ = options;
				lb_ = lb[0];
				ub_ = ub[0];
				return build(options_, options_.backTrajExtend);
			}
			// src/asif_implicit_tb.cpp:365-405: the trajectory length is recomputed WITHOUT backTrajExtend (:377),
			// satSharpness is clamped to [0.01, 2] with return codes 3 / 2 (:391-404), otherwise 1.
			int32_t updateOptions(const Options &options)
			{
				options_ = options;
				int32_t code = 1;
				if (options_.satSharpness > 2) {
					options_.satSharpness = 2;
					code = 2;
				} else if (options_.satSharpness < 0.01) {
					options_.satSharpness = 0.01;
					code = 3;
				}
				const int32_t r = build(options_, 0.0);
				return r == 1 ? code : r;
			}
			int32_t filter(const double x[], const double uDes[], double uAct[], double &relax) { return filterOne(x, uDes, uAct, &relax); }
			// closed-loop rollout of n independent agents (example main loops), state resident on the device(s)
			int32_t rollout(const int64_t n, const int32_t steps, const double dt, double X[], const double UDes[],
			                double UActLast[], int32_t rcLast[], int64_t rcHist[8] = nullptr)
			{
				if (engine_ == nullptr) return ASIF_ERR_INVALID_ARGUMENT;
				if (group_) return asif_engine_group_rollout(group_, n, steps, dt, X, UDes, UActLast, rcLast, rcHist);
				return asif_engine_rollout(engine_, n, steps, dt, X, UDes, UActLast, rcLast, rcHist, ASIF_MEM_HOST, nullptr);
			}
			// src/asif_implicit_tb.cpp:911-933
			static std::string filterErrorMsgString(const int32_t rc)
			{
				switch (rc) {
				case 1: return "Success";
				case 2: return "Inside backup set";
				case -1: return "Inside backup set QP failed";
				case -2: return "QP failed";
				case -3: return "Backup set not reached";
				default: return "Unkown";
				}
			}

		protected:
			int32_t build(const Options &o, const double extend)
			{
				asif_engine_config cfg;
				int32_t r = asif_engine_config_init(&cfg, ASIF_FILTER_IMPLICIT_TB, (int32_t)model_);
				if (r != ASIF_OK) return r;
				cfg.device = device_;
				cfg.npBTSS = (int32_t)npBTSS_;
				cfg.lb[0] = lb_;
				cfg.ub[0] = ub_;
				cfg.relaxCost = o.relaxCost;
				cfg.relaxLb = o.relaxSafeLb;
				cfg.relaxTTS = o.relaxTTS;
				cfg.relaxMinOrtho = o.relaxMinOrtho;
				cfg.backTrajHorizon = o.backTrajHorizon;
				cfg.backTrajExtend = extend;
				cfg.backTrajDt = o.backTrajDt;
				cfg.backTrajMinOrtho = o.backTrajMinOrtho;
				cfg.satSharpness = o.satSharpness;
				cfg.inf = o.inf;
				return create(cfg);
			}
			Model model_;
			uint32_t npBTSS_;
			int32_t device_;
			Options options_;
			double lb_, ub_;
		};

		// Batched ASIF::ASIFimplicit (include/asif_implicit.h): relax is double[2] per state (safe, reach).
		class FilterBatchImplicit : public FilterBatchBase
		{
		public:
			using FilterBatchBase::filter; // filter(x, H, c, uAct, relax)
			typedef struct {
				double relaxCost = 50.0;
				double relaxReachLb = 5.0;
				double relaxSafeLb = 5.0;
				double backTrajHorizon = 1.0;
				double backTrajDt = 0.01;
				double backTrajAbsTol = 1.0e-6;
				double backTrajRelTol = 1.0e-6;
				double satSharpness = 0.1;
				double inf = 1e20;
				bool use_learning = false; // adds the residual of learning_data_ to Lfh[0] / Lgh[0..nu-1] (src/asif_implicit.cpp:585-588)
			} Options; // include/asif_implicit.h:20-34 (x0 / n_debug are not on the batched path)
			LearningData learning_data_ = LearningData(); // public member as in the reference (include/asif_implicit.h:125); read at initialize / updateOptions

			FilterBatchImplicit(const Model model, const uint32_t npBTSS = 10, const int32_t device = 0)
			    : model_(model), npBTSS_(npBTSS), device_(device)
			{
			}
			int32_t initialize(const double lb[], const double ub[]) { return initialize(lb, ub, Options()); }
			int32_t initialize(const double lb[], const double ub[], const Options &options)
			{
				options_ = options;
				lb_ = lb[0];
				ub_ = ub[0];
				return build();
			}
			// src/asif_implicit.cpp:369-401: same satSharpness clamp and return codes as the TB class
			int32_t updateOptions(const Options &options)
			{
				options_ = options;
				int32_t code = 1;
				if (options_.satSharpness > 2) {
					options_.satSharpness = 2;
					code = 2;
				} else if (options_.satSharpness < 0.01) {
					options_.satSharpness = 0.01;
					code = 3;
				}
				const int32_t r = build();
				return r == 1 ? code : r;
			}
			int32_t filter(const double x[], const double uDes[], double uAct[], double relax[2]) { return filterOne(x, uDes, uAct, relax); }

			// src/asif_implicit.cpp:829-842 (same in src/asif_implicit_robust.cpp:967-980)
			static std::string filterErrorMsgString(const int32_t rc)
			{
				switch (rc) {
				case 1: return "Success";
				case -1: return "QP failed";
				default: return "Unkown";
				}
			}
		protected:
			int32_t build(void)
			{
				asif_engine_config cfg;
				int32_t r = asif_engine_config_init(&cfg, ASIF_FILTER_IMPLICIT, (int32_t)model_);
				if (r != ASIF_OK) return r;
				cfg.device = device_;
				cfg.npBTSS = (int32_t)npBTSS_;
				cfg.lb[0] = lb_;
				cfg.ub[0] = ub_;
				cfg.relaxCost = options_.relaxCost;
				cfg.relaxLb = options_.relaxSafeLb;
				cfg.relaxReachLb = options_.relaxReachLb;
				cfg.backTrajHorizon = options_.backTrajHorizon;
				cfg.backTrajDt = options_.backTrajDt;
				cfg.satSharpness = options_.satSharpness;
				cfg.inf = options_.inf;
				r = create(cfg);
				if (r == 1 && options_.use_learning) {
					const int32_t rl = asif_engine_set_learning(engine_, &learning_data_);
					if (rl != ASIF_OK) return rl;
				}
				return r;
			}
			Model model_;
			uint32_t npBTSS_;
			int32_t device_;
			Options options_;
			double lb_, ub_;
		};

		// Batched ASIF::ASIFimplicitRB (include/asif_implicit_robust.h:19-160, src/asif_implicit_robust.cpp): ASIFimplicit with
		// a zero-order-hold backup controller (Options.backContDt) and the h entries of the safety rows replaced by the
		// lower bound of the interval safety set over x_i +- x_unc.  The interval callbacks of the reference constructor
		// (safetySet_int ...) are compiled into the device model functor (safety_set_lower, asif_b200/csrc/models.cuh).
		class FilterBatchImplicitRB : public FilterBatchBase
		{
		public:
			using FilterBatchBase::filter; // filter(x, H, c, uAct, relax)
			typedef struct {
				double *x0 = nullptr;    // unused on the batched path (the reference uses it for the set-up solve only)
				double *x_unc = nullptr; // [nx] state uncertainty box; nullptr = zeros, as in the reference (:276-279)
				int n_debug = -1;        // diagnostics of the single-state class; not on the batched path
				double relaxCost = 50.0;
				double relaxReachLb = 5.0;
				double relaxSafeLb = 5.0;
				double backTrajHorizon = 1.0;
				double backContDt = 0.01;
				double backTrajDt = 0.01;
				double backTrajAbsTol = 1.0e-6;
				double backTrajRelTol = 1.0e-6;
				double satSharpness = 0.1;
				double inf = 1e20;
				bool use_learning = false; // adds the residual of learning_data_ (src/asif_implicit_robust.cpp:705-708)
			} Options; // include/asif_implicit_robust.h:22-38
			LearningData learning_data_ = LearningData(); // include/asif_implicit_robust.h:149

			FilterBatchImplicitRB(const Model model, const uint32_t npBTSS = 10, const int32_t device = 0)
			    : model_(model), npBTSS_(npBTSS), device_(device)
			{
			}
			int32_t initialize(const double lb[], const double ub[]) { return initialize(lb, ub, Options()); }
			int32_t initialize(const double lb[], const double ub[], const Options &options)
			{
				options_ = options;
				lb_ = lb[0];
				ub_ = ub[0];
				return build();
			}
			// src/asif_implicit_robust.cpp:443-479: satSharpness clamp with return codes 2 / 3
			int32_t updateOptions(const Options &options)
			{
				options_ = options;
				int32_t code = 1;
				if (options_.satSharpness > 2) {
					options_.satSharpness = 2;
					code = 2;
				} else if (options_.satSharpness < 0.01) {
					options_.satSharpness = 0.01;
					code = 3;
				}
				const int32_t r = build();
				return r == 1 ? code : r;
			}
			int32_t filter(const double x[], const double uDes[], double uAct[], double relax[2]) { return filterOne(x, uDes, uAct, relax); }

			// src/asif_implicit.cpp:829-842 (same in src/asif_implicit_robust.cpp:967-980)
			static std::string filterErrorMsgString(const int32_t rc)
			{
				switch (rc) {
				case 1: return "Success";
				case -1: return "QP failed";
				default: return "Unkown";
				}
			}
		protected:
			int32_t build(void)
			{
				asif_engine_config cfg;
				int32_t r = asif_engine_config_init(&cfg, ASIF_FILTER_IMPLICIT_RB, (int32_t)model_);
				if (r != ASIF_OK) return r;
				cfg.device = device_;
				cfg.npBTSS = (int32_t)npBTSS_;
				cfg.lb[0] = lb_;
				cfg.ub[0] = ub_;
				cfg.relaxCost = options_.relaxCost;
				cfg.relaxLb = options_.relaxSafeLb;
				cfg.relaxReachLb = options_.relaxReachLb;
				cfg.backTrajHorizon = options_.backTrajHorizon;
				cfg.backTrajDt = options_.backTrajDt;
				cfg.backContDt = options_.backContDt;
				cfg.satSharpness = options_.satSharpness;
				cfg.inf = options_.inf;
				const int nx = (model_ == Model::Segway || model_ == Model::SegwayShipped) ? 4 : 2;
				for (int i = 0; i < nx; i++) cfg.x_unc[i] = options_.x_unc ? options_.x_unc[i] : 0.0;
				r = create(cfg);
				if (r == 1 && options_.use_learning) {
					const int32_t rl = asif_engine_set_learning(engine_, &learning_data_);
					if (rl != ASIF_OK) return rl;
				}
				return r;
			}
			Model model_;
			uint32_t npBTSS_;
			int32_t device_;
			Options options_;
			double lb_, ub_;
		};

		// Batched ASIF::ASIFrobust (include/asif_robust.h) on a table of half-planes h_k = 1 - a_k.x with the
		// InvertedPendulum interval dynamics g1 in [pMin, pMax] (examples/InvertedPendulum_Robust.cpp:53-69).
		class FilterBatchRobust : public FilterBatchBase
		{
		public:
			using FilterBatchBase::filter; // filter(x, H, c, uAct, relax)
			typedef struct {
				double relaxLb = 5.0;
				double relaxCost = 50.0;
				double inf = 1e20;
			} Options; // include/asif_robust.h:14-19

			// halfplanes: npSS rows {a0, a1}; the table is copied
			FilterBatchRobust(const uint32_t npSS, const double halfplanes[], const double pMin, const double pMax,
			                  const int32_t device = 0)
			    : table_(halfplanes, halfplanes + 2 * npSS), pMin_(pMin), pMax_(pMax), device_(device)
			{
				keepOnFailure_ = true;
			}
			int32_t initialize(const double lb[], const double ub[]) { return initialize(lb, ub, Options()); }
			int32_t initialize(const double lb[], const double ub[], const Options &options)
			{
				asif_engine_config cfg;
				int32_t r = asif_engine_config_init(&cfg, ASIF_FILTER_ROBUST, ASIF_MODEL_INVERTED_PENDULUM_TABLE);
				if (r != ASIF_OK) return r;
				cfg.device = device_;
				cfg.lb[0] = lb[0];
				cfg.ub[0] = ub[0];
				cfg.relaxLb = options.relaxLb;
				cfg.relaxCost = options.relaxCost;
				cfg.inf = options.inf;
				cfg.dynParam[0] = pMin_;
				cfg.dynParam[1] = pMax_;
				cfg.halfplanes = table_.data();
				cfg.n_halfplanes = (int32_t)(table_.size() / 2);
				return create(cfg);
			}
			// uAct / relax untouched on failure (src/asif_robust.cpp:249-251)
			int32_t filter(const double x[], const double uDes[], double uAct[], double &relax) { return filterOne(x, uDes, uAct, &relax); }

		protected:
			std::vector<double> table_;
			double pMin_, pMax_;
			int32_t device_;
		};

		// Batched ASIF::ASIFrealizable (include/asif_realizable.h) for a 2-D polytope kernel.  The x-independent
		// interval Lie derivatives over each facet (facetLie[facet][active][LfLo, LfHi, LgLo, LgHi]) are inputs:
		// INTEGRATION.md shows the loop that evaluates them with the user's interval dynamics callback at initialize.
		class FilterBatchRealizable : public FilterBatchBase
		{
		public:
			using FilterBatchBase::filter; // filter(x, H, c, uAct, relax)
			typedef struct {
				double relaxDes = 5.0;
				double relaxOffset = 5.0;
				double relaxCost = 50.0;
				double inf = 1e20;
			} Options; // include/asif_realizable.h:14-20

			FilterBatchRealizable(const double uncertaintyBounds[2], const uint32_t nVertices, const double vertices[],
			                      const uint32_t nFacets, const double normals[], const int32_t facetVertices[],
			                      const int32_t facetActive[], const double facetLie[], const uint32_t maxCriticalFacets,
			                      const uint32_t maxActiveConstraints, const double pMin, const double pMax,
			                      const uint32_t npSSmax = 0, const int32_t device = 0)
			    : v_(vertices, vertices + 2 * nVertices), n_(normals, normals + 2 * nFacets),
			      fv_(facetVertices, facetVertices + 2 * nFacets), fa_(facetActive, facetActive + maxActiveConstraints * nFacets),
			      lie_(facetLie, facetLie + 4 * maxActiveConstraints * nFacets), maxCrit_(maxCriticalFacets),
			      maxAct_(maxActiveConstraints), npSSmax_(npSSmax), pMin_(pMin), pMax_(pMax), device_(device)
			{
				unc_[0] = uncertaintyBounds[0];
				unc_[1] = uncertaintyBounds[1];
				keepOnFailure_ = true;
			}
			int32_t initialize(const double lb[], const double ub[]) { return initialize(lb, ub, Options()); }
			int32_t initialize(const double lb[], const double ub[], const Options &options)
			{
				asif_engine_config cfg;
				int32_t r = asif_engine_config_init(&cfg, ASIF_FILTER_REALIZABLE, ASIF_MODEL_INVERTED_PENDULUM_KERNEL);
				if (r != ASIF_OK) return r;
				cfg.device = device_;
				cfg.lb[0] = lb[0];
				cfg.ub[0] = ub[0];
				cfg.relaxDes = options.relaxDes;
				cfg.relaxOffset = options.relaxOffset;
				cfg.relaxCost = options.relaxCost;
				cfg.inf = options.inf;
				cfg.npSSmax = (int32_t)npSSmax_;
				cfg.uncertaintyBounds[0] = unc_[0];
				cfg.uncertaintyBounds[1] = unc_[1];
				cfg.dynParam[0] = pMin_;
				cfg.dynParam[1] = pMax_;
				cfg.kernel_vertices = v_.data();
				cfg.n_vertices = (int32_t)(v_.size() / 2);
				cfg.facet_normals = n_.data();
				cfg.n_facets = (int32_t)(n_.size() / 2);
				cfg.facet_vertices = fv_.data();
				cfg.facet_active = fa_.data();
				cfg.facet_lie = lie_.data();
				cfg.max_critical_facets = (int32_t)maxCrit_;
				cfg.max_active_constraints = (int32_t)maxAct_;
				return create(cfg);
			}
			// relax[0] of the reference is a multiplier of its LP-dual formulation (not unique): reported as 0;
			// relax[1] is the barrier relaxation eps.  rc: 1, -1 (QP infeasible), -2 (outside the kernel, no critical facet)
			int32_t filter(const double x[], const double uDes[], double uAct[], double relax[2]) { return filterOne(x, uDes, uAct, relax); }

		protected:
			std::vector<double> v_, n_;
			std::vector<int32_t> fv_, fa_;
			std::vector<double> lie_;
			uint32_t maxCrit_, maxAct_, npSSmax_;
			double pMin_, pMax_, unc_[2];
			int32_t device_;
		};
	} // namespace b200
} // namespace ASIF
#endif
