// qp_admm.cu -- device instantiation and launcher of the cluster-cooperative QP solver for nv > 4 (qp_admm.cuh).
// One thread-block cluster per problem (8 CTAs x 512 threads on 8 SMs of one die, cluster.sync() as the phase barrier)
// for the large LP-dual problems, one CTA for small ones; clusters are persistent and stride over the batch, each with
// its own slice of an L2-resident workspace.  Compiled with FMA contraction: nothing here is compared bit for bit.
#include "../../include/asif_b200.h"
#include "qp_admm.cuh"

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace cg = cooperative_groups;

namespace asifb {
int fail(int code, const char *fmt, ...); // engine.cu
}
using asifb::fail;

namespace {

#define QA_CUDA_TRY(expr)                                                                                       \
	do {                                                                                                        \
		cudaError_t e__ = (expr);                                                                               \
		if (e__ != cudaSuccess)                                                                                 \
			return fail(ASIF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
	} while (0)

constexpr int QA_THREADS = 512;
constexpr int QA_CLUSTER = 8;

struct ClusterTeam {
	static constexpr int LANES = 32;
	int tid, nthreads, lane, warp, nwarps;
	bool single;
	double *scratch; // this warp's slice of the CTA's dynamic shared memory (scratch_len doubles)
	int slen;
	__device__ __forceinline__ double *warp_scratch() const { return scratch; }
	__device__ __forceinline__ int scratch_len() const { return slen; }
	__device__ __forceinline__ ClusterTeam(double *smem, int scratch_len)
	{
		scratch = smem + (size_t)(threadIdx.x >> 5) * scratch_len;
		slen = scratch_len;
		cg::cluster_group cl = cg::this_cluster();
		const unsigned cr = cl.block_rank(), cs = cl.num_blocks();
		tid = (int)(cr * blockDim.x + threadIdx.x);
		nthreads = (int)(cs * blockDim.x);
		lane = (int)(threadIdx.x & 31u);
		warp = tid >> 5;
		nwarps = nthreads >> 5;
		single = cs == 1;
	}
	// phase barrier of the team; cluster.sync() is arrive.release + wait.acquire at cluster scope, so the global-memory
	// writes of one phase are visible to every CTA of the cluster in the next
	__device__ __forceinline__ void sync() const
	{
		if (single) __syncthreads();
		else cg::this_cluster().sync();
	}
	__device__ __forceinline__ void warp_sync() const { __syncwarp(); }
	__device__ __forceinline__ double warp_bcast(double v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
	__device__ __forceinline__ double warp_sum(double v) const
	{
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
		return v;
	}
	__device__ __forceinline__ double warp_max(double v) const
	{
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
		return v;
	}
};

struct BatchArgs {
	int32_t nv, nc, diag_cost, share_h, share_bounds;
	int64_t n;
	const double *H, *c, *A, *b, *lb, *ub;
	const uint8_t *be;
	double *sol;
	int32_t *status, *info;
	double *work;
	size_t work_stride; // doubles per cluster
	int32_t near_in_smem; // one-CTA teams: the near part of the workspace (everything but the KKT factor) is in shared memory
	int32_t scratch_len;  // doubles of warp scratch per warp
};

__global__ void __launch_bounds__(QA_THREADS, 1) qp_admm_kernel(const __grid_constant__ BatchArgs a, const __grid_constant__ qpadmm::Settings st)
{
	extern __shared__ double qa_smem[];
	ClusterTeam tm(qa_smem, a.scratch_len);
	cg::cluster_group cl = cg::this_cluster();
	const unsigned csize = cl.num_blocks();
	const int64_t cluster_id = blockIdx.x / csize, n_clusters = gridDim.x / csize;
	__shared__ qpadmm::Work w;
	if (threadIdx.x == 0) {
		double *glob = a.work + (size_t)cluster_id * a.work_stride;
		if (a.near_in_smem) w = qpadmm::carve(qa_smem + (size_t)a.scratch_len * (QA_THREADS / 32), glob, a.nv, a.nc, QA_THREADS / 32);
		else w = qpadmm::carve(glob, a.nv, a.nc);
	}
	__syncthreads();
	for (int64_t k = cluster_id; k < a.n; k += n_clusters) {
		qpadmm::Problem pb;
		pb.nv = a.nv;
		pb.nc = a.nc;
		pb.diag_cost = a.diag_cost;
		pb.H = a.H + (a.share_h ? 0 : k * (int64_t)a.nv * a.nv);
		pb.c = a.c + k * a.nv;
		pb.A = a.A + k * (int64_t)a.nv * a.nc;
		pb.b = a.b + k * a.nc;
		pb.lb = a.lb + (a.share_bounds ? 0 : k * a.nv);
		pb.ub = a.ub + (a.share_bounds ? 0 : k * a.nv);
		pb.be = a.be;
		pb.sol = a.sol + k * a.nv;
		pb.status = a.status + k;
		pb.info = a.info ? a.info + qpadmm::NINFO * k : nullptr;
		qpadmm::Solver<ClusterTeam> s(tm, st, a.nv, a.nc, w);
		const int32_t code = s.solve(pb);
		// the QPWrapper convention (src/qpwrapper_osqp.cpp:217-237): 1 = FEASIBLE for solved / solved inaccurate, else
		// the solver's own status value
		if (tm.tid == 0 && (code == qpadmm::ST_SOLVED || code == qpadmm::ST_SOLVED_INACCURATE)) *pb.status = 1;
		tm.sync();
	}
}

std::mutex g_cfg_mu;
qpadmm::Settings g_settings = qpadmm::default_settings();

struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	int dev = -1;
};
thread_local DevBuf tl_work, tl_info;
thread_local int32_t tl_last_info[qpadmm::NINFO] = {0};

int ensure(DevBuf &b, int device, size_t bytes)
{
	if (b.p && (b.dev != device || b.cap < bytes)) {
		cudaSetDevice(b.dev);
		cudaFree(b.p);
		cudaSetDevice(device);
		b.p = nullptr;
		b.cap = 0;
	}
	if (!b.p) {
		QA_CUDA_TRY(cudaMalloc(&b.p, bytes));
		b.cap = bytes;
		b.dev = device;
	}
	return ASIF_OK;
}

} // namespace

namespace asifb {

// called by asif_qp_solve_batch (engine.cu) for nv > MAX_NV with DEVICE pointers (or mapped host pointers); the launch
// is asynchronous on `st`, the workspace is a per-host-thread buffer that only grows
int launch_qp_admm(int device, int nv, int nc, int64_t n, int diag_cost, const double *H, const double *c, const double *A,
                   const double *b, const double *lb, const double *ub, const uint8_t *be, double *sol, int32_t *status,
                   int share_flags, cudaStream_t st, bool fetch_info)
{
	if (nv > 1024 || nc > 4096) return fail(ASIF_ERR_UNSUPPORTED, "qp_solve_batch: nv = %d / nc = %d beyond the dense solver's range (1024 / 4096)", nv, nc);
	qpadmm::Settings cfg;
	{
		std::lock_guard<std::mutex> lk(g_cfg_mu);
		cfg = g_settings;
	}
	// small problems: one CTA per problem; large ones: a cluster, so that eight SMs' L2 bandwidth serves the factorisations
	const size_t mat = (size_t)nv * nv + (size_t)nv * nc;
	// (measured on the 402-variable QP: 2 / 4 / 8 / 16 CTAs = 61 / 35 / 19.3 / 15.8 ms per solve); 16 needs the non-portable
	// cluster size and a GPC with 16 free SMs: asked for first, 8 is the fallback
	int want = mat > 16384 ? (mat > 200000 ? 16 : QA_CLUSTER) : 1;
	if (const char *e = getenv("ASIF_B200_QP_CLUSTER")) {
		const int v = atoi(e);
		if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) want = v;
	}
	// warp scratch of Solver::factor: four columns per warp when 16 warps' worth fits beside nothing else (cluster teams), one otherwise
	const int ld_col = (nv + 2) & ~1;
	const int scratch_len = (want > 1 && (size_t)4 * ld_col * (QA_THREADS / 32) * sizeof(double) <= 180 * 1024) ? 4 * ld_col : ld_col;
	// per device, cheap: set at every launch
	QA_CUDA_TRY(cudaFuncSetAttribute(qp_admm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
	if (want > 8 && cudaFuncSetAttribute(qp_admm_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
		cudaGetLastError();
		want = 8;
	}
	cudaLaunchConfig_t lc = {};
	cudaLaunchAttribute attr[1];
	int csize = want, max_clusters = 0;
	size_t smem = 0;
	bool near_in_smem = false;
	for (;; csize /= 2) {
		attr[0].id = cudaLaunchAttributeClusterDimension;
		attr[0].val.clusterDim.x = (unsigned)csize;
		attr[0].val.clusterDim.y = 1;
		attr[0].val.clusterDim.z = 1;
		smem = sizeof(double) * (size_t)scratch_len * (QA_THREADS / 32);
		const size_t near_bytes = sizeof(double) * qpadmm::work_near_doubles(nv, nc, QA_THREADS / 32);
		near_in_smem = csize == 1 && smem + near_bytes <= 200 * 1024 && !getenv("ASIF_B200_QP_NO_SMEM");
		if (near_in_smem) smem += near_bytes;
		lc.blockDim = dim3(QA_THREADS);
		lc.dynamicSmemBytes = smem;
		lc.stream = st;
		lc.attrs = attr;
		lc.numAttrs = 1;
		lc.gridDim = dim3((unsigned)csize);
		max_clusters = 0;
		const cudaError_t qe = cudaOccupancyMaxActiveClusters(&max_clusters, qp_admm_kernel, &lc);
		if (qe == cudaSuccess && max_clusters >= 1) break;
		cudaGetLastError();
		if (csize == 1) return fail(ASIF_ERR_CUDA, "qp_solve_batch: the solver kernel does not fit on this device (%s)", cudaGetErrorString(qe));
	}
	const int64_t n_clusters = n < max_clusters ? n : max_clusters;
	const size_t stride = (qpadmm::work_doubles(nv, nc) + 31) & ~(size_t)31;
	int r = ensure(tl_work, device, stride * sizeof(double) * (size_t)n_clusters);
	if (r) return r;
	r = ensure(tl_info, device, sizeof(int32_t) * qpadmm::NINFO * (size_t)n);
	if (r) return r;
	BatchArgs a;
	a.nv = nv;
	a.nc = nc;
	a.diag_cost = diag_cost;
	a.share_h = share_flags & ASIF_QP_SHARED_H;
	a.share_bounds = share_flags & ASIF_QP_SHARED_BOUNDS;
	a.n = n;
	a.H = H; a.c = c; a.A = A; a.b = b; a.lb = lb; a.ub = ub; a.be = be;
	a.sol = sol;
	a.status = status;
	a.info = (int32_t *)tl_info.p;
	a.work = (double *)tl_work.p;
	a.work_stride = stride;
	a.near_in_smem = near_in_smem ? 1 : 0;
	a.scratch_len = scratch_len;
	lc.gridDim = dim3((unsigned)(n_clusters * csize));
	QA_CUDA_TRY(cudaLaunchKernelEx(&lc, qp_admm_kernel, a, cfg));
	QA_CUDA_TRY(cudaGetLastError());
	// statistics of the first problem, for asif_qp_last_info (the QPWrapper path solves one problem per call)
	if (fetch_info) QA_CUDA_TRY(cudaMemcpyAsync(tl_last_info, tl_info.p, sizeof(tl_last_info), cudaMemcpyDeviceToHost, st));
	return ASIF_OK;
}

} // namespace asifb

extern "C" int32_t asif_qp_configure(double eps_abs_rel, int32_t max_iter, int32_t polish, int32_t polish_refine_iter)
{
	std::lock_guard<std::mutex> lk(g_cfg_mu);
	const qpadmm::Settings d = qpadmm::default_settings();
	g_settings.eps_abs = g_settings.eps_rel = eps_abs_rel > 0 ? eps_abs_rel : d.eps_abs;
	g_settings.max_iter = max_iter > 0 ? max_iter : d.max_iter;
	g_settings.polish = polish >= 0 ? (polish ? 1 : 0) : d.polish;
	g_settings.polish_refine_iter = polish_refine_iter >= 0 ? polish_refine_iter : d.polish_refine_iter;
	return ASIF_OK;
}

extern "C" int32_t asif_qp_last_info(int32_t info[8])
{
	if (!info) return fail(ASIF_ERR_INVALID_ARGUMENT, "info is NULL");
	for (int i = 0; i < qpadmm::NINFO; i++) info[i] = tl_last_info[i];
	return ASIF_OK;
}
