// latency_server.cuh -- persistent one-warp kernels that serve single-state filter() calls without a kernel launch.
//
// SURVEY 8f rank 1 asks for a latency path for the unmodified single-state user (the example mains call filter() once per
// control step, examples/DoubleIntegrator_implicit_tb.cpp:117-131).  A launch plus a stream synchronisation costs ~11 us
// on this box before the kernel does anything (host_check --latency, round 2: 12 us for the explicit filter whose kernel
// is a few hundred instructions); the reference's CPU filter() takes 16.5 us per call on the same box.  The server keeps
// one warp resident: the host writes the inputs into a pinned, device-mapped mailbox and bumps a sequence number; lane 0
// polls that number over PCIe, the warp filters up to 32 states with the SAME per-state functions the batch kernels use
// (tb_filter_tile / explicit_filter_tile: identical bits), stores the outputs into the mailbox, fences system-wide and
// writes the sequence number back, which the host is spinning on.  Round trip = two PCIe latencies + the kernel's serial
// chain.  Opt-in (asif_engine_latency_server): a resident warp is a cost, and cudaDeviceSynchronize would never return
// while it runs - the engine stops the server around the few calls that need a device-wide synchronisation.
#pragma once
#include "explicit_kernel.cuh"
#include "tb_kernel.cuh"

namespace asifb {

constexpr int SRV_MAX_STATES = 32;                  // one warp, one state per lane
constexpr unsigned long long SRV_EXIT = ~0ull;      // sequence number that ends the kernel
// mailbox layout (8-byte words): [0] sequence number from the host, [1] n, [2] sequence number echoed by the device, then the arrays
constexpr int SRV_HDR = 8;

__device__ __forceinline__ unsigned long long srv_poll(const volatile unsigned long long *p)
{
	unsigned long long v;
	asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
	return v;
}

__device__ __forceinline__ double srv_poll_f64(const double *p)
{
	double v;
	asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
	return v;
}

// Mailbox arrays behind the header: x[32][NX], cost[32][NU] (uDes), u_act[32][NU], relax[32][NR], rc[32] (int32).
// The inputs are read with volatile loads into shared memory first: the per-state functions take their inputs through
// const __restrict__ pointers, i.e. possibly through the read-only data path, which must never see a mailbox that the
// host rewrites while the kernel is running.  Outputs go through shared memory the same way and are stored by the server.
template <int NX, int NU, int NR>
struct SrvStage {
	double x[SRV_MAX_STATES * NX], ud[SRV_MAX_STATES * NU], ua[SRV_MAX_STATES * NU], rl[SRV_MAX_STATES * NR];
	int32_t rc[SRV_MAX_STATES];
};

template <int NX, int NU, int NR>
__device__ __forceinline__ void srv_load(SrvStage<NX, NU, NR> &st, const double *mailbox, const int n)
{
	const double *x = mailbox + SRV_HDR, *ud = x + SRV_MAX_STATES * NX;
	for (int i = threadIdx.x; i < n * NX; i += 32) st.x[i] = srv_poll_f64(x + i);
	for (int i = threadIdx.x; i < n * NU; i += 32) st.ud[i] = srv_poll_f64(ud + i);
	__syncwarp();
}

template <int NX, int NU, int NR>
__device__ __forceinline__ void srv_store(const SrvStage<NX, NU, NR> &st, double *mailbox, const int n, const unsigned long long seq)
{
	__syncwarp();
	double *ua = mailbox + SRV_HDR + SRV_MAX_STATES * (NX + NU), *rl = ua + SRV_MAX_STATES * NU;
	int32_t *rc = reinterpret_cast<int32_t *>(rl + SRV_MAX_STATES * NR);
	for (int i = threadIdx.x; i < n * NU; i += 32) ua[i] = st.ua[i];
	for (int i = threadIdx.x; i < n * NR; i += 32) rl[i] = st.rl[i];
	for (int i = threadIdx.x; i < n; i += 32) rc[i] = st.rc[i];
	__threadfence_system();
	__syncwarp();
	if (threadIdx.x == 0) {
		reinterpret_cast<volatile unsigned long long *>(mailbox)[2] = seq;
		__threadfence_system();
	}
}

__device__ __forceinline__ bool srv_wait(const double *mailbox, unsigned long long &seen, int &n)
{
	const unsigned long long *hdr = reinterpret_cast<const unsigned long long *>(mailbox);
	unsigned long long seq = 0;
	if (threadIdx.x == 0) {
		do {
			seq = srv_poll(hdr);
		} while (seq == seen);
	}
	seq = __shfl_sync(0xffffffffu, seq, 0);
	if (seq == SRV_EXIT) return false;
	seen = seq;
	unsigned long long nn = 0;
	if (threadIdx.x == 0) nn = srv_poll(hdr + 1);
	n = (int)__shfl_sync(0xffffffffu, nn, 0);
	n = n < 0 ? 0 : (n > SRV_MAX_STATES ? SRV_MAX_STATES : n);
	return true;
}

template <class M, int NPBTSS, int SATMODE>
__global__ void __launch_bounds__(32, 1) tb_server_kernel(const TbParams p, double *mailbox, double *gsnap)
{
	constexpr int NX = M::NX, NU = M::NU;
	extern __shared__ double smem[];
	__shared__ SrvStage<NX, NU, 1> st;
	double *snap = tb_global_snapshots<M>() ? gsnap + threadIdx.x : smem + threadIdx.x;
	unsigned long long seen = 0;
	int n = 0;
	while (srv_wait(mailbox, seen, n)) {
		srv_load(st, mailbox, n);
		// snapshot stride TB_THREADS: the row functor addresses its snapshots with that compile-time stride (tb_kernel.cuh),
		// so the scratch is sized for a full CTA although one warp uses it
		tb_filter_tile<M, NPBTSS, false, SATMODE>(p, (int64_t)n, (int64_t)threadIdx.x, snap, TB_THREADS, st.x, st.ud, st.ua, st.rl, st.rc, nullptr,
		                                          nullptr);
		srv_store(st, mailbox, n, seen);
	}
}

template <class M>
__global__ void __launch_bounds__(32, 1) explicit_server_kernel(const ExplicitParams p, double *mailbox)
{
	constexpr int NX = M::NX, NU = M::NU;
	__shared__ SrvStage<NX, NU, 1> st;
	unsigned long long seen = 0;
	int n = 0;
	while (srv_wait(mailbox, seen, n)) {
		srv_load(st, mailbox, n);
		explicit_filter_tile<M, false>(p, (int64_t)n, (int64_t)threadIdx.x, st.x, st.ud, st.ua, st.rl, st.rc, nullptr, nullptr);
		srv_store(st, mailbox, n, seen);
	}
}

} // namespace asifb
