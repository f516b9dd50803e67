/*
 * asif_b200.h -- C ABI of the B200 batched safety-filter engine (libasif_b200.so).
 *
 * This is the drop-in boundary: plain C types, pointers and sizes only.  The C++ host layer
 * (asif_b200/host/asif_b200.hpp: ASIF::QPWrapperB200, ASIF::FilterBatch*) and any foreign
 * binding (ctypes, cgo, JNI ...) sit on top of exactly these entry points.  Every entry point
 * names the reference interface it replaces (paths relative to DrewSingletary/asif).
 *
 * Conventions (same as the reference):
 *   - all reals are FP64; matrices are dense column-major (A[i + j*nc]), include/qpwrapper_abstract.h:11-15
 *   - a batch is an array of states: x[n][nx], u_des[n][nu], u_act[n][nu], relax[n][n_relax], rc[n]
 *   - per-state outcomes are reported only through rc[] with the reference's codes
 *     (src/asif_implicit_tb.cpp:300-361,911-933; src/asif.cpp:199-209); a failing state never
 *     aborts the batch.  Function return values are 0 on success or a negative ASIF_ERR_* code;
 *     asif_last_error() gives the message (thread local).
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef ASIF_B200_H
#define ASIF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASIF_B200_ABI_VERSION 3

/* error codes (function return values).  They start at -101 so that a wrapper which hands a function's failure on in place
 * of a per-state return code (the single-state filter() of the host classes, QPWrapperB200::solve) can never be mistaken
 * for one of the reference's own codes: -1 QP failed / backup input applied, -2 OSQP iteration limit, -3 backup set not
 * reached or primal infeasible, -4 dual infeasible, -7 non-convex, -10 unsolved (src/asif_implicit_tb.cpp:300-361,
 * OSQP 0.6 constants.h). */
#define ASIF_OK 0
#define ASIF_ERR_INVALID_ARGUMENT (-101)
#define ASIF_ERR_UNSUPPORTED (-102)
#define ASIF_ERR_CUDA (-103)
#define ASIF_ERR_NO_DEVICE (-104)
#define ASIF_ERR_INTERNAL (-105) /* host-side failure (allocation, thread creation) caught at the C boundary */

/* filter classes of the reference */
#define ASIF_FILTER_EXPLICIT 1    /* ASIF::ASIF            include/asif.h:39-65           */
#define ASIF_FILTER_IMPLICIT_TB 2 /* ASIF::ASIFimplicitTB  include/asif_implicit_tb.h:93-110 */
#define ASIF_FILTER_IMPLICIT 3    /* ASIF::ASIFimplicit    include/asif_implicit.h:96-113 */
#define ASIF_FILTER_ROBUST 4      /* ASIF::ASIFrobust      include/asif_robust.h:41-59    */
#define ASIF_FILTER_REALIZABLE 5  /* ASIF::ASIFrealizable  include/asif_realizable.h:58-76 */
#define ASIF_FILTER_IMPLICIT_RB 6 /* ASIF::ASIFimplicitRB  include/asif_implicit_robust.h:124-146 */

/* models: the reference passes std::function callbacks; on the device a model is a compiled-in
 * functor with the same five callback signatures (asif_b200/csrc/models.cuh). */
#define ASIF_MODEL_DOUBLE_INTEGRATOR 1    /* examples/DoubleIntegrator.cpp:12-61              */
#define ASIF_MODEL_DOUBLE_INTEGRATOR_TB 2 /* examples/DoubleIntegrator_implicit_tb.cpp:13-85  */
#define ASIF_MODEL_INVERTED_PENDULUM 3    /* examples/InvertedPendulum_Implicit.cpp:13-80     */
#define ASIF_MODEL_INVERTED_PENDULUM_TABLE 4 /* examples/InvertedPendulum_Robust.cpp:53-79 + half-plane table */
#define ASIF_MODEL_INVERTED_PENDULUM_KERNEL 5 /* examples/InvertedPendulum_RealizableSampled.cpp:46-62 + polytope kernel table */
#define ASIF_MODEL_SEGWAY 6               /* examples/segway_implicit_tb.cpp:13-212, backup set centred on the controller equilibrium */
#define ASIF_MODEL_SEGWAY_SHIPPED 7       /* same, backup set exactly as shipped (:41-55)     */

/* memory-space flags for the batch calls */
#define ASIF_MEM_HOST 0   /* all batch pointers are host memory.  Pageable arrays are staged through chunked copies; pinned
                           * (cudaHostAlloc / cudaHostRegisterMapped) arrays overlap the copies with the kernels and, for
                           * the filter classes whose kernels outlast their transfers, are read and written by the kernel
                           * directly over PCIe (env ASIF_B200_HOST_IO = staged | out | inout | auto; results are the same
                           * bits in every mode; "auto" measures staged against in-place on the first large batches of
                           * an engine and keeps the faster, see asif_engine_host_io_stats) */
#define ASIF_MEM_DEVICE 1 /* all batch pointers are device memory on the engine's device; nothing is copied */

/* how a host-memory batch reached the kernels (asif_engine_last_host_io) */
#define ASIF_HOST_IO_STAGED 0 /* chunked H2D copy, kernel, D2H copy */
#define ASIF_HOST_IO_OUT 1    /* chunked H2D copies; the kernels store into the caller's pinned output arrays */
#define ASIF_HOST_IO_INOUT 2  /* one launch reading and writing the caller's pinned arrays over PCIe */

typedef struct asif_engine asif_engine;

/*
 * Engine configuration = the reference's per-class Options struct + the constructor/initialize
 * arguments (dims, lb, ub), flattened.  Fill with asif_engine_config_init() first: it sets every
 * option to the reference default of the filter class (include/asif.h:11-17,
 * include/asif_implicit_tb.h:19-33, include/asif_implicit.h:20-34, include/asif_robust.h:14-19,
 * include/asif_realizable.h:14-20) and lb/ub/npBTSS to the values of the model's example program.
 */
typedef struct asif_engine_config {
	uint32_t struct_size; /* sizeof(asif_engine_config), set by asif_engine_config_init */
	int32_t filter;       /* ASIF_FILTER_* */
	int32_t model;        /* ASIF_MODEL_*  */
	int32_t device;       /* CUDA device ordinal */
	int32_t npBTSS;       /* critical trajectory points (implicit filters) */
	int32_t npSSmax;      /* explicit/realizable: rows kept after sorting by h; <= 0 means all */
	double lb[2], ub[2];  /* input bounds passed to initialize(lb, ub) */
	/* Options */
	double relaxCost;
	double relaxLb;      /* ASIF/ASIFrobust relaxLb; implicit classes: relaxSafeLb */
	double relaxReachLb; /* ASIFimplicit */
	double relaxTTS, relaxMinOrtho;                                        /* ASIFimplicitTB */
	double backTrajHorizon, backTrajExtend, backTrajDt, backTrajMinOrtho;  /* implicit classes */
	double satSharpness;
	double inf;
	double relaxDes, relaxOffset;    /* ASIFrealizable */
	double uncertaintyBounds[4];     /* ASIFrealizable */
	double dynParam[4];              /* model parameters, e.g. IP input gain interval [pMin, pMax] */
	/* tables (host pointers, copied at create): half-planes {a0,a1} for ASIFrobust
	 * (include/KernelData_70-135kg.h) or the polytope kernel for ASIFrealizable */
	const double *halfplanes;
	int32_t n_halfplanes;
	const double *kernel_vertices; /* [nVertices][nx] */
	int32_t n_vertices;
	const double *facet_normals;   /* [nFacets][nx] */
	const int32_t *facet_vertices; /* [nFacets][nx] vertex ids */
	const int32_t *facet_active;   /* [nFacets][maxActiveConstraints] facet ids, -1 = absent */
	/* [nFacets][maxActiveConstraints][4] = LfLo, LfHi, LgLo, LgHi: the interval Lie derivatives of the active
	 * constraint over the facet, which src/asif_realizable.cpp:470-500 recomputes on every call although they do
	 * not depend on x.  The caller evaluates its interval dynamics callback once per facet at initialize
	 * (INTEGRATION.md shows the loop) and hands the numbers over. */
	const double *facet_lie;
	int32_t n_facets, max_critical_facets, max_active_constraints;
	/* ASIFimplicitRB (include/asif_implicit_robust.h:22-38): zero-order-hold period of the backup controller and the
	 * state-uncertainty box x_unc[nx] (Options.x_unc; the reference borrows the pointer, here the values are copied).
	 * The interval safety set the class evaluates on x +- x_unc is compiled into the model functor
	 * (safety_set_lower in asif_b200/csrc/models.cuh). */
	double backContDt;
	double x_unc[4];
} asif_engine_config;

int32_t asif_b200_abi_version(void);
const char *asif_last_error(void);
/* number of CUDA devices visible, or a negative error */
int32_t asif_device_count(void);

/* defaults for (filter, model); replaces the Options() default constructors cited above */
int32_t asif_engine_config_init(asif_engine_config *cfg, int32_t filter, int32_t model);

/* constructor + initialize(lb, ub, options) of the filter classes
 * (e.g. src/asif_implicit_tb.cpp:89-154,169-232) */
int32_t asif_engine_create(const asif_engine_config *cfg, asif_engine **out);
/* An engine is NOT re-entrant: like the reference's filter objects (member scratch, SURVEY 8b "Threading") it keeps
 * per-call state (staging slots, the cost mode of the running call), so use one engine per host thread, or a group. */
int32_t asif_engine_destroy(asif_engine *e);

/* dims[0..5] = nx, nu, n_relax, nc (rows of A), nv, n_diag (doubles per state of the diag record) */
int32_t asif_engine_dims(const asif_engine *e, int32_t dims[6]);

/*
 * filter(x, uDes, uAct, relax) of the selected class on n independent states
 * (src/asif_implicit_tb.cpp:243-363, src/asif.cpp:144-210, ...).  New surface: the reference has no
 * batched call.  diag may be NULL; when given it receives n_diag doubles per state:
 *   TB       : TTS_, BTorthoBS_, hSafetyNow_, hBackupEnd_, backTrajCritIdx_[npBTSS] (-1 = absent),
 *              A_[nc*nv] column-major, b_[nc]                 (include/asif_implicit_tb.h:112-117)
 *   explicit : A_[nc*nv], b_[nc]
 * mem = ASIF_MEM_HOST | ASIF_MEM_DEVICE.  stream: a cudaStream_t (device-memory calls are
 * enqueued on it and return without synchronising) or NULL (engine's own stream; host-memory
 * calls always return with the outputs complete).
 */
int32_t asif_engine_filter_batch(asif_engine *e, int64_t n, const double *x, const double *u_des, double *u_act,
                                 double *relax, int32_t *rc, double *diag, int32_t mem, void *stream);

/*
 * Pinned host memory for the batch arrays of ASIF_MEM_HOST calls.  The reference copies caller arrays into its own
 * new[] buffers (src/asif.cpp:27-34, src/qpwrapper_osqp.cpp:14-29); a batch of 1e7 states is 440 MB per call, so the
 * caller's arrays are used in place instead, and they are fastest when the device can address them:
 * asif_host_alloc returns such memory (cudaHostAlloc, portable + mapped), asif_host_register pins and maps a range the
 * caller already owns (page-aligned ranges register fastest; unregister before freeing it).  Pageable arrays work too,
 * through staged copies.  asif_engine_last_host_io reports which ASIF_HOST_IO_* path the last host batch took
 * (-1 before the first one).
 */
int32_t asif_host_alloc(void **p, uint64_t bytes);
int32_t asif_host_free(void *p);
int32_t asif_host_register(void *p, uint64_t bytes);
int32_t asif_host_unregister(void *p);
int32_t asif_engine_last_host_io(const asif_engine *e, int32_t *mode);
/* what the "auto" policy has measured on this engine: wall time per 1e6 states of large pinned batches in each
 * ASIF_HOST_IO_* mode (exponential average; 0 with samples 0 = not tried yet) */
int32_t asif_engine_host_io_stats(const asif_engine *e, double ms_per_1e6_states[3], int32_t samples[3]);

/*
 * Latency server (the single-state user of the example mains, one filter() per control step).  on = 1
 * keeps one warp resident on the device; host batches of up to 32 states (filter(x, uDes, ...), no diag) then go through a
 * pinned mailbox instead of a kernel launch: the same per-state code, the same bits, about 10 us less per call.  Built for
 * ASIF / DoubleIntegrator and ASIFimplicitTB / DoubleIntegrator (npBTSS 4); ASIF_ERR_UNSUPPORTED otherwise (the launch path
 * stays).  on = 0 stops it; asif_engine_destroy stops it too.  While it runs, do not call cudaDeviceSynchronize on this
 * device from the same thread of control (it waits for the resident kernel): synchronise streams instead.
 */
int32_t asif_engine_latency_server(asif_engine *e, int32_t on);

/*
 * The filter(x, H, c, uAct[, relax]) overloads (src/asif_implicit_tb.cpp:252-363, src/asif.cpp:153-210,
 * src/asif_implicit.cpp:296-356, src/asif_implicit_robust.cpp:384-441): the caller supplies the linear cost c of the
 * whole decision vector, c[n][nv] per state (the input part AND the relax entries, which updateCost(uDes) would
 * leave at -2 relaxCost relaxLb), and optionally the nu x nu input block H of the Hessian (column-major, host
 * memory; NULL keeps the current one).  As in the reference (updateH) a given H stays in force for later calls,
 * including filter_batch(uDes).  Only the diagonal of H is used: diagonalCost = true, the constructors' default.
 * Classes: ASIF, ASIFimplicitTB, ASIFimplicit, ASIFimplicitRB.
 */
int32_t asif_engine_filter_batch_cost(asif_engine *e, int64_t n, const double *x, const double *H, const double *c,
                                      double *u_act, double *relax, int32_t *rc, double *diag, int32_t mem, void *stream);
/*
 * ASIF::filter(x, uDes, uAct, Lfh, Lgh[, relax]) (include/asif.h:43-58, src/asif.cpp:125-141, 287-292): the caller supplies
 * the Lie derivatives of the nc = min(npSSmax, npSS) selected safety functions, Lfh[n][nc] and Lgh[n][nc x nu] (per state
 * column-major), in place of Dh f and Dh g; h still comes from the safety set.  Unlike the reference (which keeps the
 * caller's pointers and stays in this mode for every later call) the arrays apply to this call only.  Explicit filter only.
 */
int32_t asif_engine_filter_batch_lie(asif_engine *e, int64_t n, const double *x, const double *u_des, const double *Lfh,
                                     const double *Lgh, double *u_act, double *relax, int32_t *rc, double *diag, int32_t mem,
                                     void *stream);
/* updateH alone (host pointer, nu x nu column-major) */
int32_t asif_engine_set_input_cost(asif_engine *e, const double *H);

/*
 * Engine groups: ONE caller batch over several GPUs of one box (SURVEY 8b "device list", 8e).  States are independent, so
 * the batch is cut into contiguous slices [g*ceil(n/G), (g+1)*ceil(n/G)), one per device, and there is no collective: each
 * device has its own engine (options and tables replicated), host worker thread and streams, and runs the single-device
 * host-memory path on its slice of the CALLER's arrays - results land in the caller's u_act / relax / rc, nothing is
 * gathered afterwards.  devices = NULL (or n_devices <= 0) takes every visible device; cfg->device is ignored.
 * All batch pointers are host memory (pinned arrays from asif_host_alloc are fastest, as for a single engine).
 * A group runs one batch at a time; calls from several threads are serialised.  asif_engine_group_engine returns a
 * borrowed handle of member i for the per-engine queries (asif_engine_dims, asif_engine_last_host_io, ...).
 * Replaces, for the batched surface, the loop over filter() a multi-GPU caller of the reference API would write
 * (include/asif_implicit_tb.h:93-110 and siblings).
 */
typedef struct asif_engine_group asif_engine_group;
int32_t asif_engine_group_create(const asif_engine_config *cfg, const int32_t *devices, int32_t n_devices, asif_engine_group **out);
int32_t asif_engine_group_destroy(asif_engine_group *g);
int32_t asif_engine_group_size(const asif_engine_group *g); /* devices in the group, or a negative error */
asif_engine *asif_engine_group_engine(asif_engine_group *g, int32_t i);
/* bounds[0..1] = the slice [lo, hi) of an n-state batch that member i processes */
int32_t asif_engine_group_slice(const asif_engine_group *g, int64_t n, int32_t i, int64_t bounds[2]);
int32_t asif_engine_group_filter_batch(asif_engine_group *g, int64_t n, const double *x, const double *u_des, double *u_act,
                                       double *relax, int32_t *rc, double *diag);
int32_t asif_engine_group_filter_batch_cost(asif_engine_group *g, int64_t n, const double *x, const double *H, const double *c,
                                            double *u_act, double *relax, int32_t *rc, double *diag);
/* rc_hist[8] is summed over the members */
int32_t asif_engine_group_rollout(asif_engine_group *g, int64_t n, int32_t steps, double dt, double *x, const double *u_des,
                                  double *u_act_last, int32_t *rc_last, int64_t *rc_hist);

/*
 * Closed-loop rollout as in the example main loops (examples/segway_implicit_tb.cpp:251-283):
 * steps x { filter ; x += dt*(f(x) + g(x) uAct) } with the state resident on the device.
 * x is updated in place; u_des is held per state.  rc_hist[8] (host memory, optional) counts
 * return codes over all state-steps: index rc+3 for rc in [-3,2], index 7 = other.
 */
int32_t asif_engine_rollout(asif_engine *e, int64_t n, int32_t steps, double dt, double *x, const double *u_des,
                            double *u_act_last, int32_t *rc_last, int64_t *rc_hist, int32_t mem, void *stream);

/*
 * The loop around filter() in the example programs, for a fleet of n independent agents with the state resident on
 * the device: every steps_per_sample plant steps the state is sampled (xEstim), filtered, optionally passed through
 * the smoothBounds rate limiter, and held while the plant  x += dt (f(x) + g(x) uAct)  advances
 * (examples/InvertedPendulum_Implicit.cpp:113-147, examples/InvertedPendulum_RealizableSampled.cpp:258-304 `dtPerSample`,
 * examples/DoubleIntegrator_RealizableSampled.cpp:110-205 `smoothBounds`, examples/segway_implicit_tb.cpp:251-283).
 * Works for every filter class; asif_engine_rollout is the fused fast path for ASIFimplicitTB without sampling or log.
 */
typedef struct asif_loop_config {
	uint32_t struct_size;     /* set by asif_loop_config_init */
	int32_t steps;            /* plant steps */
	double dt;                /* plant Euler step (examples: 1e-3) */
	int32_t steps_per_sample; /* dtPerSample >= 1 */
	int32_t smooth_bounds;    /* 1: rate limiter on input 0 */
	double smooth_lb, smooth_ub, smooth_rate; /* examples: -20, 20, 20*dtPerSample*0.001 */
	double plant_gain;        /* input gain p of dynamicsExact for the pendulum table/kernel models */
	int32_t log_stride;       /* record every log_stride-th plant step; 0: no log */
	int32_t log_after_step;   /* 1: record (t+dt, x after the step) as the segway/implicit examples do; 0: (t, x before) */
	int64_t log_agents;       /* agents [0, log_agents) are recorded */
} asif_loop_config;

int32_t asif_loop_config_init(asif_loop_config *cfg);
/* dims[0] = doubles per log record, dims[1] = records per agent for this config.  Record layout:
 * t, x[nx], xEstim[nx], uDes[nu], uFilter[nu], uAct[nu], relax[n_relax], rc, smoothLo, smoothHi, TTS, BTorthoBS, critIdx0
 * (the last three are filled for ASIFimplicitTB, 0 otherwise); log[agent][record][field]. */
int32_t asif_engine_loop_log_dims(const asif_engine *e, const asif_loop_config *cfg, int64_t dims[2]);
/* x is updated in place; u_act_last / relax_last / rc_last receive the last filter call's outputs; rc_hist[8] (host,
 * optional) counts return codes over all filter calls as asif_engine_rollout does; log may be NULL when log_stride is 0.
 * All batch pointers and log live in the memory space `mem`. */
int32_t asif_engine_closed_loop(asif_engine *e, int64_t n, const asif_loop_config *cfg, double *x, const double *u_des,
                                double *u_act_last, double *relax_last, int32_t *rc_last, int64_t *rc_hist, double *log,
                                int32_t mem, void *stream);

/*
 * Learned residual of the implicit classes (Options.use_learning + the public member learning_data_,
 * include/asif_implicit.h:33,125, include/asif_implicit_robust.h:37,149; arithmetic include/asif_learning_utils.h:34-155):
 * two MLPs with two ReLU hidden layers on [x ; Dh_index_[0..nx-1] ; 0...]; the drift net's first output is added to
 * Lfh[0], the actuation net's first nu outputs to Lgh[0..nu-1], before A_ and b_ are filled.  The struct mirrors
 * ASIF::LearningData field for field (weights are column-major [rows x cols], as matrixVectorMultiply reads them);
 * the arrays are host memory and are copied.  data == NULL switches the residual off.  Every layer width <= 64.
 * ASIF_FILTER_IMPLICIT and ASIF_FILTER_IMPLICIT_RB engines only.
 */
typedef struct asif_learning_data {
	uint32_t d_drift_in, d_act_in, d_drift_hidden, d_act_hidden, d_drift_hidden_2, d_act_hidden_2, d_drift_out, d_act_out;
	const double *w_1_drift, *w_2_drift, *w_3_drift, *b_1_drift, *b_2_drift, *b_3_drift;
	const double *w_1_act, *w_2_act, *w_3_act, *b_1_act, *b_2_act, *b_3_act;
} asif_learning_data;
int32_t asif_engine_set_learning(asif_engine *e, const asif_learning_data *data);

/* mean QP work of the last filter_batch/rollout call: rows processed by the active-set solver,
 * summed over states (the "K-bar" of SURVEY 8d is this / states) */
int32_t asif_engine_last_qp_iterations(asif_engine *e, uint64_t *rows_processed);

/*
 * The QPWrapper backend: n independent dense QPs of identical shape
 *      min v'Hv + c'v   s.t.  A v >= b (row i equality where be[i]),  lb <= v <= ub
 * replacing QPWrapperOsqp::initialize/updateCost/updateA/updateb/solve/getSolution
 * (src/qpwrapper_osqp.cpp:55-261).  Layout per problem as in the reference:
 * H[nv*nv] (diagonal_cost: only the diagonal is read, src/qpwrapper_osqp.cpp:267-283), c[nv],
 * A[nc*nv] column-major, b[nc], lb[nv], ub[nv]; problems are consecutive.  H, lb, ub, be may be
 * shared by all problems (stride flags).  status[k] = 1 (QPWrapperAbstract::SOLVER_STATUS::FEASIBLE)
 * or the OSQP code the reference would pass through (-3 primal infeasible, -2 iteration limit, -4 dual
 * infeasible, -7 not convex, 3 / 4 the "inaccurate" infeasibility verdicts).
 * Two solvers, picked by nv:
 *   nv <= 4  (every QP the batched filter classes pose, H positive definite): one problem per THREAD, exact
 *            dual active-set method in registers (csrc/qp_gi.cuh);
 *   nv  > 4  (the LP-dual formulations of ASIFrobust / ASIFrealizable, nv = 402 / 38, src/asif_robust.cpp:21-22;
 *            H may be positive semi-definite): one problem per thread-block CLUSTER, the OSQP algorithm (ADMM
 *            with cached explicit inverse, adaptive rho, infeasibility certificates, active-set polish) at
 *            eps_abs = eps_rel = 1e-8 (csrc/qp_admm.cuh); nv <= 1024, nc <= 4096; solution is NaN where OSQP
 *            returns none.  asif_qp_configure changes its accuracy settings (process-wide; arguments <= 0 / < 0
 *            restore the defaults 1e-8, 20000, polish on, 10 refinement steps); asif_qp_last_info returns
 *            {ADMM iterations, rho updates, polish 1 accepted / -1 rejected / 0 not run, active rows, and the
 *            microseconds spent in equilibration, factorisations, iterations, polish} of the first problem of
 *            the calling thread's last ASIF_MEM_HOST call.
 */
int32_t asif_qp_configure(double eps_abs_rel, int32_t max_iter, int32_t polish, int32_t polish_refine_iter);
int32_t asif_qp_last_info(int32_t info[8]);
#define ASIF_QP_SHARED_H 1      /* one H for all problems */
#define ASIF_QP_SHARED_BOUNDS 2 /* one lb/ub for all problems */
int32_t asif_qp_solve_batch(int32_t device, int32_t nv, int32_t nc, int64_t n, int32_t diagonal_cost,
                            const double *H, const double *c, const double *A, const double *b, const double *lb,
                            const double *ub, const uint8_t *be, double *sol, int32_t *status, int32_t share_flags,
                            int32_t mem, void *stream);

/* measured FP64 FMA rate of the device (dependent-chain DFMA microbenchmark), TFLOP/s; the
 * roofline denominator bench.py reports next to the filter throughput */
int32_t asif_measure_fp64_peak(int32_t device, double *tflops, double *sm_clock_mhz_est);

#ifdef __cplusplus
}
#endif
#endif
