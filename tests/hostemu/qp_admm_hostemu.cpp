/*
 * qp_admm_hostemu.cpp -- TEST INFRASTRUCTURE ONLY.  Never part of libasif_b200.so, never loaded by the product.
 *
 * The device solver for nv > 4 (asif_b200/csrc/qp_admm.cuh) is written against a Team concept; this file instantiates
 * the same text with a team of ONE host thread (every barrier a no-op, every warp reduction the identity) and exports
 * the two symbols oracle/_ref/libasif_ref_b200.so imports (asif_device_count, asif_qp_solve_batch).  Loaded ahead of
 * that library by tests/test_qp_admm_hostemu.py, it lets the CPU suite (no GPU in this container) run the reference's
 * own ASIFrobust / ASIFrealizable classes on the ALGORITHM of the device solver and compare with the OSQP stand-in
 * build.  What it cannot show is the parallel execution (races, barriers): that is tests/test_gpu_qp_admm.py's job.
 */
#include "../../asif_b200/csrc/qp_admm.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace {
struct HostTeam {
	static constexpr int LANES = 1;
	int tid = 0, nthreads = 1, lane = 0, warp = 0, nwarps = 1;
	std::vector<double> scratch_;
	double *warp_scratch() { return scratch_.data(); }
	int scratch_len() const { return (int)scratch_.size(); }
	void sync() {}
	void warp_sync() {}
	double warp_bcast(double v, int) const { return v; }
	double warp_sum(double v) const { return v; }
	double warp_max(double v) const { return v; }
};
int32_t g_info[qpadmm::NINFO] = {0};
} // namespace

extern "C" {
int32_t asif_device_count(void) { return 1; }
void hostemu_last_info(int32_t *info)
{
	for (int i = 0; i < 4; i++) info[i] = g_info[i];
}
int32_t asif_qp_solve_batch(int32_t, int32_t nv, int32_t nc, int64_t n, int32_t diagonal_cost, const double *H, const double *c,
                            const double *A, const double *b, const double *lb, const double *ub, const uint8_t *be, double *sol,
                            int32_t *status, int32_t share_flags, int32_t, void *)
{
	/* nv <= 4 (the facet feasibility QP of ASIFrealizable) goes through the same algorithm here; on the device those take the
	 * per-thread active-set solver, which is device code only */
	std::vector<double> ws(qpadmm::work_doubles(nv, nc));
	qpadmm::Settings st = qpadmm::default_settings();
	if (const char *e = getenv("HOSTEMU_REFINE")) st.polish_refine_iter = atoi(e);
	for (int64_t k = 0; k < n; k++) {
		qpadmm::Problem pb;
		pb.nv = nv;
		pb.nc = nc;
		pb.diag_cost = diagonal_cost;
		pb.H = H + ((share_flags & 1) ? 0 : k * (int64_t)nv * nv);
		pb.c = c + k * nv;
		pb.A = A + k * (int64_t)nv * nc;
		pb.b = b + k * nc;
		pb.lb = lb + ((share_flags & 2) ? 0 : k * nv);
		pb.ub = ub + ((share_flags & 2) ? 0 : k * nv);
		pb.be = be;
		pb.sol = sol + k * nv;
		int32_t stt = 0;
		pb.status = &stt;
		pb.info = g_info;
		HostTeam tm;
		tm.scratch_.resize(2 * (nv + 2)); /* two columns at a time: exercises the grouped inverse */
		const qpadmm::Work wk = qpadmm::carve(ws.data(), nv, nc);
		qpadmm::Solver<HostTeam> s(tm, st, nv, nc, wk);
		s.solve(pb);
		status[k] = (stt == qpadmm::ST_SOLVED || stt == qpadmm::ST_SOLVED_INACCURATE) ? 1 : stt;
		if (getenv("HOSTEMU_TRACE"))
			fprintf(stderr, "hostemu: nv %d nc %d status %d iters %d rho_updates %d polish %d active %d\n", nv, nc, stt, g_info[0],
			        g_info[1], g_info[2], g_info[3]);
	}
	return 0;
}
}
